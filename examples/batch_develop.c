/* A compiled caller of the drop-in boundary: the loop rtgui/batchqueue.cc runs over a directory of raw files
 * (BatchQueue::startProcessing -> rtengine::startBatchProcessing, one job after the other, L586-676) re-pointed at
 * include/art_hotpath.h.  Plain C99, no CUDA headers, no Python: this is what a maintainer's patch compiles down to.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/batch_develop.c -o examples/batch_develop -Lart_b200 -lart_hotpath -Wl,-rpath,$PWD/art_b200 -lm
 *   examples/batch_develop [frames] [width] [height]
 *
 * Every job: a synthetic RGGB frame (stands in for the decoded raw file; decoding is out of scope) is developed --
 * AMaZE, getImage gains / camera->working matrix, RGB_denoise, Fattal -- through the batch-queue entry points
 * (art_hp_develop_submit_packed / art_hp_develop_wait: two frames in flight, upload / kernels / download overlapped) and
 * arrives in host memory as the 16-bit interleaved scanlines the reference's TIFF / PNG writers take
 * (Imagefloat::getScanline).  Prints frames per second and a checksum of each frame. */
#include <math.h>
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

/* deterministic stand-in for a decoded raw frame, in scaleColors' 0..65535 domain */
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

#define CHECK(call)                                                                           \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ != ART_HP_OK) {                                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? art_hp_last_error(ctx) : ""); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

int main(int argc, char** argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 6, W = argc > 2 ? atoi(argv[2]) : 2048, H = argc > 3 ? atoi(argv[3]) : 1536;
    art_hp_ctx* ctx = NULL;
    if (art_hp_device_count() < 1) { fprintf(stderr, "no CUDA device: the hot path has no CPU fallback\n"); return 2; }
    CHECK(art_hp_create(&ctx, 0));

    static const double cam2work[9] = {0.82, 0.15, 0.03, 0.07, 0.96, -0.03, 0.02, -0.10, 1.08};
    static const double prophoto[9] = {0.7976749, 0.1351917, 0.0313534, 0.2880402, 0.7118741, 0.0000857, 0.0, 0.0, 0.8252100};
    art_hp_denoise_params dn;
    memset(&dn, 0, sizeof dn);
    dn.luminance = 30; dn.luminanceDetail = 50; dn.chrominance = 15; dn.gamma = 1.7; dn.scale = 1.0;
    art_hp_develop_params p;
    memset(&p, 0, sizeof p);
    p.method = ART_HP_BAYER_AMAZE; p.filters = 0x94949494u; p.initialGain = 1.0; p.border = 4;
    p.mul[0] = 1.9f; p.mul[1] = 1.0f; p.mul[2] = 1.6f; p.doClip = 1; p.cam2work = cam2work;
    p.denoise = &dn; p.wprof = prophoto;
    p.fattal_enabled = 1; p.fattal_threshold = 30; p.fattal_amount = 20;

    int Wo, Ho, border;
    CHECK(art_hp_develop_size(&p, W, H, &Wo, &Ho, &border));
    /* two slots: frame k is filled / collected while frame k-1 is on the device */
    float* raw[2];
    uint16_t* out[2];
    float** rows[2];
    for (int s = 0; s < 2; ++s) {
        raw[s] = (float*)art_hp_host_alloc((size_t)W * H * sizeof(float));
        out[s] = (uint16_t*)art_hp_host_alloc((size_t)Wo * Ho * 3 * sizeof(uint16_t));
        rows[s] = (float**)malloc((size_t)H * sizeof(float*));
        if (!raw[s] || !out[s] || !rows[s]) { fprintf(stderr, "host allocation failed\n"); return 1; }
        for (int y = 0; y < H; ++y) rows[s][y] = raw[s] + (size_t)y * W;      /* array2D<float>'s row table */
    }
    const double t0 = now();
    for (int k = 0; k < frames; ++k) {
        const int s = k & 1;
        if (k >= 2) CHECK(art_hp_develop_wait(ctx));       /* frame k-2 is complete in out[s]: a writer would take it here, then the slot is reused */
        synth_frame(raw[s], W, H, 1000u + (unsigned)k);
        CHECK(art_hp_develop_submit_packed(ctx, &p, W, H, rows[s], 16, 0, out[s], (size_t)Wo * 3 * sizeof(uint16_t)));
    }
    while (art_hp_develop_pending(ctx) > 0) CHECK(art_hp_develop_wait(ctx));
    const double dt = now() - t0;
    unsigned long long sum = 0;
    for (int s = 0; s < 2; ++s)
        for (size_t i = 0; i < (size_t)Wo * Ho * 3; i += 97) sum += out[s][i];
    printf("batch_develop: %d frames of %dx%d -> %dx%d 16-bit scanlines in %.3f s (%.1f Mpixel/s incl. frame synthesis), checksum %llu\n", frames, W, H, Wo, Ho,
           dt, frames * (double)W * H / dt / 1e6, sum);
    for (int s = 0; s < 2; ++s) { art_hp_host_free(raw[s]); art_hp_host_free(out[s]); free(rows[s]); }
    art_hp_destroy(ctx);
    return 0;
}
