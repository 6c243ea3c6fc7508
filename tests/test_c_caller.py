"""The boundary from a compiled caller: examples/batch_develop.c (plain C99, no CUDA headers) builds against include/art_hotpath.h and
libart_hotpath.so; without a GPU it refuses to run (no CPU fallback), on the GPU box it develops a batch through the batch-queue entries."""
import os
import subprocess

import pytest

import art_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "batch_develop")


def build():
    art_b200.load_library()
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-O2", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "batch_develop.c"), "-o", EXE, "-L", os.path.join(ROOT, "art_b200"), "-lart_hotpath",
                    "-Wl,-rpath," + os.path.join(ROOT, "art_b200"), "-lm"], check=True)


def test_c_caller_builds_and_refuses_without_a_gpu():
    build()
    if art_b200.load_library().art_hp_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "2", "256", "256"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_caller_develops_a_batch():
    build()
    r = subprocess.run([EXE, "5", "1024", "768"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "5 frames of 1024x768 -> 1016x760" in r.stdout
    again = subprocess.run([EXE, "5", "1024", "768"], capture_output=True, text=True, timeout=300)
    assert again.stdout.split("checksum")[1] == r.stdout.split("checksum")[1]          # the same frames, the same bits
    print("\n" + r.stdout.strip())


DISPATCH = os.path.join(ROOT, "examples", "batch_dispatch")


def build_dispatch():
    art_b200.load_library()
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-O2", "-pthread", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "batch_dispatch.c"), "-o", DISPATCH, "-L", os.path.join(ROOT, "art_b200"), "-lart_hotpath",
                    "-Wl,-rpath," + os.path.join(ROOT, "art_b200"), "-lm"], check=True)


def test_dispatcher_builds_and_refuses_without_a_gpu():
    build_dispatch()
    if art_b200.load_library().art_hp_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([DISPATCH, "2", "256", "256"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_dispatcher_runs_the_queue_on_every_gpu():
    """examples/batch_dispatch.c: one worker thread and one context per GPU pulling jobs from a shared queue (the N-GPU replacement of the
    reference's serial batchProcessingThread loop).  The per-job results do not depend on how many GPUs share the queue."""
    build_dispatch()
    n = art_b200.load_library().art_hp_device_count()
    one = subprocess.run([DISPATCH, "7", "1024", "768", "1"], capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr + one.stdout
    assert "7 of 7 jobs" in one.stdout
    if n > 1:
        many = subprocess.run([DISPATCH, "7", "1024", "768", str(n)], capture_output=True, text=True, timeout=600)
        assert many.returncode == 0, many.stderr + many.stdout
        assert many.stdout.split("checksum")[1] == one.stdout.split("checksum")[1]
        print("\n" + many.stdout.strip())
    print("\n" + one.stdout.strip())
