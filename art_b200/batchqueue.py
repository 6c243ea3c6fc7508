"""Host-side mirror of the reference's batch loop over the C-ABI batch entry.

Reference: rtgui/batchqueue.cc BatchQueue::startProcessing / processNext (L586-676) hands one job after the other to
rtengine::startBatchProcessing (rtengine/simpleprocess.cc), which develops it on the calling thread.  Here a rank (one process
per GPU) takes every `world`-th job (`dist.frames_for_rank`) and keeps two of its frames in flight through
art_hp_develop_submit / art_hp_develop_wait, so the PCIe copies of one frame hide behind the kernels of its neighbours.
There is no exchange between ranks: frames are independent.
"""
from . import dist as adist


class BatchQueue:
    """`jobs`: a sequence of callables (or arrays) giving each job's (H, W) float32 CFA plane; `params`: DevelopParams (one for all
    jobs or one per job).  Iterating yields (job_index, [red, green, blue]) in submission order; the yielded planes are views of
    the queue's pinned buffers and are valid only until the next iteration -- resuming the generator writes the next raw frame into
    the same slot and queues kernels that overwrite its planes (copy them out, as the reference saves the file)."""

    def __init__(self, hot_path, jobs, params, rank=0, world=1):
        self.hp, self.jobs, self.params = hot_path, jobs, params
        self.mine = adist.frames_for_rank(len(jobs), rank, world)
        self._slots = None

    def _slot(self, k, H, W, params):
        Ho, Wo = params.out_shape(H, W)
        if self._slots is None or self._slots[0][0].array.shape != (H, W) or self._slots[0][1].array.shape != (Ho, Wo):
            # frames in flight keep their own slots alive through `inflight`; new shapes get new buffers
            self._slots = [[self.hp.pinned(H, W)] + [self.hp.pinned(Ho, Wo) for _ in range(3)] for _ in range(2)]
        return self._slots[k & 1]

    def _params_of(self, j):
        return self.params[j] if isinstance(self.params, (list, tuple)) else self.params

    def __iter__(self):
        inflight = []                                   # job indices, oldest first
        for k, j in enumerate(self.mine):
            job = self.jobs[j]
            raw = job() if callable(job) else job
            H, W = raw.shape
            if len(inflight) == 2:                      # the slot about to be reused still holds the oldest frame
                self.hp.develop_wait()
                done = inflight.pop(0)
                yield done[0], [p.array for p in done[1][1:]]
            slot = self._slot(k, H, W, self._params_of(j))
            slot[0].array[:] = raw
            self.hp.develop_submit(slot[0].array, self._params_of(j), slot[1].array, slot[2].array, slot[3].array)
            inflight.append((j, slot))
        while inflight:
            self.hp.develop_wait()
            done = inflight.pop(0)
            yield done[0], [p.array for p in done[1][1:]]
