// Hot / dead pixel filter for sm_100a.
//
// Replaces RawImageSource::findHotDeadPixels (reference rtengine/badpixels.cc L477-627, sum5x5 L36-58) and
// RawImageSource::interpolateBadPixelsBayer (L66-180), the two steps of RawImageSource::preprocess's hot / dead pixel filter
// (rawimagesource.cc L1394-1420).  The reference walks the frame with a five-row ring of (sample - median of its same-colour
// neighbourhood) per thread; the ring is an optimisation of a plane that is zero outside rows 2 .. H-3 / columns 2 .. W-3, and that
// plane is what is built here:
//   k_bp_dev    one thread per pixel: Bayer -- median of the nine samples at distance 2 (a 19-exchange selection network: the reference's
//               own network returns the same element for NaN-free input); X-Trans -- the first 9 / 7 / 5 / 3 / 1 same-colour samples of
//               the 5x5 window in raster order (L496-518), median by rank counting
//   k_bp_mark   one thread per pixel: |dev| against varthresh x the 5x5 sum of |dev|; the sum keeps the reference's SSE2 association, which
//               follows the RING slots (row % 5), not the rows: ((s0 + s1) + (s2 + s3)) + s4 per column, four columns through vhadd, the
//               fifth after them.  Marks are OR-ed into a byte map (PixelsMap::set); the count leaves through one atomic per warp.
//   k_bp_interp gradient-weighted mean of the good same-colour pairs, plain mean as the fallback; reads good pixels, writes bad ones: race-free
//               in place exactly as the reference's parallel loop is.
// 4 B read + 4 B written, then 4 B (+ the 5x5 through L1 / L2) read per pixel and a byte written per bad pixel.
// Compiled with -fmad=false, IEEE division.
#include "ctx.h"

namespace {

__device__ __forceinline__ unsigned fc_bp(unsigned filters, int row, int col) { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }

__device__ __forceinline__ void cswap(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
__device__ __forceinline__ float median9(float p0, float p1, float p2, float p3, float p4, float p5, float p6, float p7, float p8)
{   // Paeth's 19-exchange median-of-9 selection network
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    return p4;
}
// median of the first n (odd) entries by rank: the element with exactly n / 2 elements ordered before it (ties by index)
__device__ __forceinline__ float median_rank(const float* v, int n)
{
    float out = v[0];
    for (int i = 0; i < n; ++i) {
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (v[j] < v[i] || (v[j] == v[i] && j < i)) ? 1 : 0;
        if (rank == n / 2) out = v[i];
    }
    return out;
}

struct BpCfa { int m[36]; };

__global__ void __launch_bounds__(256) k_bp_dev(const float* __restrict__ raw, size_t rp, float* __restrict__ dev, size_t dp, int W, int H,
                                                int is_xtrans, const __grid_constant__ BpCfa cfa)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= W) return;
    float out = 0.f;
    if (i >= 2 && i < H - 2 && j >= 2 && j < W - 2) {
#define RAW(r, c) raw[(size_t)(r) * rp + (c)]
        float med;
        if (!is_xtrans) {
            med = median9(RAW(i - 2, j - 2), RAW(i - 2, j), RAW(i - 2, j + 2), RAW(i, j - 2), RAW(i, j), RAW(i, j + 2),
                          RAW(i + 2, j - 2), RAW(i + 2, j), RAW(i + 2, j + 2));
        } else {
            const int c = cfa.m[(i % 6) * 6 + (j % 6)];
            float m[9];
            int n = 0;
            for (int y = i - 2; y < i + 3; ++y)
                for (int x = j - 2; x < j + 3; ++x)
                    if (cfa.m[(y % 6) * 6 + (x % 6)] == c) { if (n < 9) m[n] = RAW(y, x); ++n; }
            med = n >= 9 ? median_rank(m, 9) : n >= 7 ? median_rank(m, 7) : n >= 5 ? median_rank(m, 5) : n >= 3 ? median_rank(m, 3) : m[0];
        }
        out = RAW(i, j) - med;
#undef RAW
    }
    dev[(size_t)i * dp + j] = out;
}

__global__ void __launch_bounds__(256) k_bp_mark(const float* __restrict__ dev, size_t dp, unsigned char* __restrict__ map, size_t mp, int W, int H,
                                                 float varthresh, int hot, int dead, int* __restrict__ count)
{
    const int cc = blockIdx.x * blockDim.x + threadIdx.x, rr = blockIdx.y + 2;
    bool bad = false;
    if (cc >= 2 && cc < W - 2 && rr < H - 2) {
        float pixdev = dev[(size_t)rr * dp + cc];
        const bool skip = (!dead && pixdev <= 0.f) || (!hot && pixdev >= 0.f);
        if (!skip) {
            pixdev = fabsf(pixdev);
            // ring slot k holds the row r of rr-2 .. rr+2 with r % 5 == k
            const int base = rr - 2, s0 = base % 5;
            const float* in[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int r = base + ((k - s0) + 5) % 5;
                in[k] = dev + (size_t)r * dp;
            }
            float col[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int x = cc - 2 + k;
                col[k] = ((fabsf(in[0][x]) + fabsf(in[1][x])) + (fabsf(in[2][x]) + fabsf(in[3][x]))) + fabsf(in[4][x]);
            }
            float hfnbrave = -pixdev;
            hfnbrave += (col[0] + col[2]) + (col[1] + col[3]);
            hfnbrave += col[4];
            bad = pixdev > varthresh * hfnbrave;
        }
    }
    if (bad) map[(size_t)rr * mp + cc] = 1;
    const unsigned ballot = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count, __popc(ballot));
}

__global__ void __launch_bounds__(256) k_bp_interp_bayer(float* __restrict__ raw, size_t rp, const unsigned char* __restrict__ map, size_t mp, int W, int H,
                                                         unsigned filters, int* __restrict__ count)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y + 2;
    bool done = false;
    if (col >= 2 && col < W - 2 && row < H - 2 && map[(size_t)row * mp + col]) {
#define RAW(r, c) raw[(size_t)(r) * rp + (c)]
#define BAD(x, y) (map[(size_t)(y) * mp + (x)] != 0)
        const float eps = 1.f;
        float wtdsum = 0.f, norm = 0.f;
        if (fc_bp(filters, row & 1, col & 1) == 1) {
            for (int dx = -1; dx <= 1; dx += 2) {
                if (BAD(col + dx, row - 1) || BAD(col - dx, row + 1)) continue;
                const float dirwt = 0.70710678f / (fabsf(RAW(row - 1, col + dx) - RAW(row + 1, col - dx)) + eps);
                wtdsum += dirwt * (RAW(row - 1, col + dx) + RAW(row + 1, col - dx));
                norm += dirwt;
            }
        } else {
            for (int dx = -2; dx <= 2; dx += 4) {
                if (BAD(col + dx, row - 2) || BAD(col - dx, row + 2)) continue;
                const float dirwt = 0.35355339f / (fabsf(RAW(row - 2, col + dx) - RAW(row + 2, col - dx)) + eps);
                wtdsum += dirwt * (RAW(row - 2, col + dx) + RAW(row + 2, col - dx));
                norm += dirwt;
            }
        }
        if (!(BAD(col - 2, row) || BAD(col + 2, row))) {
            const float dirwt = 0.5f / (fabsf(RAW(row, col - 2) - RAW(row, col + 2)) + eps);
            wtdsum += dirwt * (RAW(row, col - 2) + RAW(row, col + 2));
            norm += dirwt;
        }
        if (!(BAD(col, row - 2) || BAD(col, row + 2))) {
            const float dirwt = 0.5f / (fabsf(RAW(row - 2, col) - RAW(row + 2, col)) + eps);
            wtdsum += dirwt * (RAW(row - 2, col) + RAW(row + 2, col));
            norm += dirwt;
        }
        if (norm > 0.f) {
            RAW(row, col) = wtdsum / (2.f * norm);
            done = true;
        } else {
            int tot = 0;
            float sum = 0.f;
            for (int dy = -2; dy <= 2; dy += 2)
                for (int dx = -2; dx <= 2; dx += 2) {
                    if (BAD(col + dx, row + dy)) continue;
                    sum += RAW(row + dy, col + dx);
                    tot++;
                }
            if (tot > 0) { RAW(row, col) = sum / tot; done = true; }
        }
#undef RAW
#undef BAD
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, done);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count, __popc(ballot));
}

}  // namespace

// d_count: one device int, zeroed here, receives the number of pixels marked by this call
int art_find_hot_dead_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans36, const float* raw, size_t rp, float thresh, int hot, int dead,
                          unsigned char* map, size_t mp, int* d_count)
{
    cudaStream_t st = ctx->stream;
    ART_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), st));
    if (W < 5 || H < 5) return ART_HP_OK;
    const size_t dp = round_up((size_t)W, 32);
    void* blk = nullptr;
    int rc = art_pool_alloc(ctx, dp * H * sizeof(float), &blk);
    if (rc) return rc;
    float* dev = (float*)blk;
    BpCfa cfa{};
    if (xtrans36) for (int i = 0; i < 36; ++i) cfa.m[i] = xtrans36[i];
    const float varthresh = (20.f * (thresh / 100.f) + 1.f) / 24.f * (xtrans36 ? 0.25f : 1.f);      // L481
    art_prof_begin(ctx, "k_bp_dev");
    k_bp_dev<<<dim3((W + 255) / 256, H), 256, 0, st>>>(raw, rp, dev, dp, W, H, xtrans36 != nullptr, cfa);
    art_prof_end(ctx);
    art_prof_begin(ctx, "k_bp_mark");
    k_bp_mark<<<dim3((W + 255) / 256, H - 4), 256, 0, st>>>(dev, dp, map, mp, W, H, varthresh, hot, dead, d_count);
    art_prof_end(ctx);
    ctx->launches += 2;
    art_pool_free(ctx, blk);
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_interpolate_bad_bayer_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t rp, const unsigned char* map, size_t mp, int* d_count)
{
    cudaStream_t st = ctx->stream;
    ART_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), st));
    if (W < 5 || H < 5) return ART_HP_OK;
    art_prof_begin(ctx, "k_bp_interp_bayer");
    k_bp_interp_bayer<<<dim3((W + 255) / 256, H - 4), 256, 0, st>>>(raw, rp, map, mp, W, H, filters, d_count);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
