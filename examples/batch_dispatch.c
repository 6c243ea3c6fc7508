/* An N-GPU job dispatcher for the drop-in boundary: what replaces the serial loop of rtgui/batchqueue.cc (BatchQueue::startProcessing ->
 * rtengine::startBatchProcessing -> batchProcessingThread, rtengine/simpleprocess.cc L586-612: one job after the other on one thread) once the
 * hot path runs behind include/art_hotpath.h.  Plain C99 + pthreads, no CUDA headers.
 *
 *   gcc -std=c99 -O2 -pthread -Iinclude examples/batch_dispatch.c -o examples/batch_dispatch -Lart_b200 -lart_hotpath -Wl,-rpath,$PWD/art_b200 -lm
 *   examples/batch_dispatch [jobs] [width] [height] [gpus]
 *
 * One worker thread per GPU, each with its own context (a context belongs to one device and one thread at a time) and two pinned slots.
 * Workers pull job numbers from a shared counter -- the queue's "next job" -- so a slow job does not hold the others back; within a worker
 * two jobs are in flight (art_hp_develop_submit_packed / art_hp_develop_wait), so upload, kernels and download of consecutive jobs overlap.
 * There is no collective and no exchange between the GPUs: jobs are independent, which is how SURVEY.md 8(e) says configs[4] should run.
 * Every job's result is a function of its job number alone: the per-job checksums do not depend on the number of GPUs or on which GPU
 * took the job, and the program prints their sum so that a run on one GPU and a run on eight can be compared. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "art_hotpath.h"

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* deterministic stand-in for a decoded raw frame (decoding is out of scope), in scaleColors' 0..65535 domain */
static void synth_frame(float* raw, int W, int H, unsigned seed)
{
    unsigned s = seed * 2654435761u + 12345u;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            s = s * 1664525u + 1013904223u;
            const float noise = (float)(s >> 8) * (1.0f / 16777216.0f) - 0.5f;
            const float scene = 0.5f + 0.35f * sinf(0.013f * x + 0.001f * seed) * cosf(0.009f * y) + 0.1f * sinf(0.21f * (x + y));
            const int c = ((y & 1) << 1) | (x & 1);                     /* RGGB */
            const float gain = c == 0 ? 0.55f : c == 3 ? 0.7f : 1.0f;
            float v = 65535.0f * gain * scene * (1.0f + 0.04f * noise);
            raw[(size_t)y * W + x] = v < 0.f ? 0.f : v;
        }
}

typedef struct {
    pthread_mutex_t lock;
    int next, jobs;                 /* the queue: job numbers next .. jobs - 1 are waiting */
    unsigned long long* checksum;   /* per job */
    int* taken_by;                  /* per job: the GPU that developed it */
    int W, H;
} queue_t;

typedef struct { queue_t* q; int device; int done; int rc; char err[256]; } worker_t;

static int take_job(queue_t* q)
{
    pthread_mutex_lock(&q->lock);
    const int j = q->next < q->jobs ? q->next++ : -1;
    pthread_mutex_unlock(&q->lock);
    return j;
}

static unsigned long long checksum16(const uint16_t* p, size_t n)
{
    unsigned long long s = 0;
    for (size_t i = 0; i < n; i += 97) s += p[i];
    return s;
}

#define WCHECK(call)                                                                                         \
    do {                                                                                                     \
        int rc_ = (call);                                                                                    \
        if (rc_ != ART_HP_OK) {                                                                              \
            w->rc = rc_;                                                                                     \
            snprintf(w->err, sizeof w->err, "%s -> %d: %s", #call, rc_, ctx ? art_hp_last_error(ctx) : ""); \
            goto out;                                                                                        \
        }                                                                                                    \
    } while (0)

static void* worker(void* arg)
{
    worker_t* w = (worker_t*)arg;
    queue_t* q = w->q;
    const int W = q->W, H = q->H;
    art_hp_ctx* ctx = NULL;
    float* raw[2] = {NULL, NULL};
    uint16_t* out[2] = {NULL, NULL};
    float** rows[2] = {NULL, NULL};
    int job_in_slot[2] = {-1, -1};
    static const double cam2work[9] = {0.82, 0.15, 0.03, 0.07, 0.96, -0.03, 0.02, -0.10, 1.08};
    static const double prophoto[9] = {0.7976749, 0.1351917, 0.0313534, 0.2880402, 0.7118741, 0.0000857, 0.0, 0.0, 0.8252100};
    art_hp_denoise_params dn;
    art_hp_develop_params p;
    int Wo = 0, Ho = 0, border = 0;
    memset(&dn, 0, sizeof dn);
    dn.luminance = 30; dn.luminanceDetail = 50; dn.chrominance = 15; dn.gamma = 1.7; dn.scale = 1.0;
    memset(&p, 0, sizeof p);
    p.method = ART_HP_BAYER_AMAZE; p.filters = 0x94949494u; p.initialGain = 1.0; p.border = 4;
    p.mul[0] = 1.9f; p.mul[1] = 1.0f; p.mul[2] = 1.6f; p.doClip = 1; p.cam2work = cam2work;
    p.denoise = &dn; p.wprof = prophoto;
    p.fattal_enabled = 1; p.fattal_threshold = 30; p.fattal_amount = 20;

    WCHECK(art_hp_create(&ctx, w->device));
    WCHECK(art_hp_develop_size(&p, W, H, &Wo, &Ho, &border));
    for (int s = 0; s < 2; ++s) {
        raw[s] = (float*)art_hp_host_alloc((size_t)W * H * sizeof(float));
        out[s] = (uint16_t*)art_hp_host_alloc((size_t)Wo * Ho * 3 * sizeof(uint16_t));
        rows[s] = (float**)malloc((size_t)H * sizeof(float*));
        if (!raw[s] || !out[s] || !rows[s]) { w->rc = ART_HP_ERR_NOMEM; snprintf(w->err, sizeof w->err, "host allocation failed"); goto out; }
        for (int y = 0; y < H; ++y) rows[s][y] = raw[s] + (size_t)y * W;      /* array2D<float>'s row table */
    }
    for (int k = 0;; ++k) {
        const int s = k & 1;
        if (job_in_slot[s] >= 0) {                 /* the job submitted two rounds ago is complete in out[s]: a writer would take it here */
            WCHECK(art_hp_develop_wait(ctx));
            q->checksum[job_in_slot[s]] = checksum16(out[s], (size_t)Wo * Ho * 3);
            q->taken_by[job_in_slot[s]] = w->device;
            job_in_slot[s] = -1;
            w->done++;
        }
        const int j = take_job(q);
        if (j < 0) break;
        synth_frame(raw[s], W, H, 1000u + (unsigned)j);          /* "load and decode job j" */
        WCHECK(art_hp_develop_submit_packed(ctx, &p, W, H, rows[s], 16, 0, out[s], (size_t)Wo * 3 * sizeof(uint16_t)));
        job_in_slot[s] = j;
    }
    for (int k = 0; k < 2; ++k) {                  /* drain, oldest first */
        int s = -1;
        for (int t = 0; t < 2; ++t)
            if (job_in_slot[t] >= 0 && (s < 0 || job_in_slot[t] < job_in_slot[s])) s = t;
        if (s < 0) break;
        WCHECK(art_hp_develop_wait(ctx));
        q->checksum[job_in_slot[s]] = checksum16(out[s], (size_t)Wo * Ho * 3);
        q->taken_by[job_in_slot[s]] = w->device;
        job_in_slot[s] = -1;
        w->done++;
    }
out:
    for (int s = 0; s < 2; ++s) { if (raw[s]) art_hp_host_free(raw[s]); if (out[s]) art_hp_host_free(out[s]); free(rows[s]); }
    if (ctx) art_hp_destroy(ctx);
    return NULL;
}

int main(int argc, char** argv)
{
    const int jobs = argc > 1 ? atoi(argv[1]) : 8, W = argc > 2 ? atoi(argv[2]) : 2048, H = argc > 3 ? atoi(argv[3]) : 1536;
    const int have = art_hp_device_count();
    if (have < 1) { fprintf(stderr, "no CUDA device: the hot path has no CPU fallback\n"); return 2; }
    int gpus = argc > 4 ? atoi(argv[4]) : have;
    if (gpus < 1 || gpus > have) gpus = have;
    if (jobs < 1 || W < 64 || H < 64) { fprintf(stderr, "usage: batch_dispatch [jobs] [width] [height] [gpus]\n"); return 1; }

    queue_t q;
    pthread_mutex_init(&q.lock, NULL);
    q.next = 0; q.jobs = jobs; q.W = W; q.H = H;
    q.checksum = (unsigned long long*)calloc((size_t)jobs, sizeof *q.checksum);
    q.taken_by = (int*)calloc((size_t)jobs, sizeof *q.taken_by);
    worker_t* ws = (worker_t*)calloc((size_t)gpus, sizeof *ws);
    pthread_t* th = (pthread_t*)calloc((size_t)gpus, sizeof *th);
    if (!q.checksum || !q.taken_by || !ws || !th) return 1;

    const double t0 = now();
    for (int g = 0; g < gpus; ++g) {
        ws[g].q = &q; ws[g].device = g;
        if (pthread_create(&th[g], NULL, worker, &ws[g])) { fprintf(stderr, "pthread_create failed\n"); return 1; }
    }
    int failed = 0, done = 0;
    for (int g = 0; g < gpus; ++g) {
        pthread_join(th[g], NULL);
        if (ws[g].rc) { fprintf(stderr, "GPU %d: %s\n", g, ws[g].err); failed = 1; }
        done += ws[g].done;
    }
    const double dt = now() - t0;
    unsigned long long sum = 0;
    for (int j = 0; j < jobs; ++j) sum += q.checksum[j] * (unsigned long long)(j + 1);
    printf("batch_dispatch: %d of %d jobs of %dx%d on %d GPU(s) in %.3f s (%.1f Mpixel/s incl. frame synthesis and context set-up), per GPU:", done, jobs, W, H,
           gpus, dt, done * (double)W * H / dt / 1e6);
    for (int g = 0; g < gpus; ++g) printf(" %d", ws[g].done);
    printf(", checksum %llu\n", sum);
    free(q.checksum); free(q.taken_by); free(ws); free(th);
    pthread_mutex_destroy(&q.lock);
    return failed || done != jobs;
}
