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

// RawImageSource::interpolateBadPixelsXtrans (badpixels.cc L288-475) in the reference's one-thread (raster) order.  Every pair the function weighs is
// checked against the map, so those reads see samples no pass ever writes; the "virtual pixel" of a red / blue site and its distance-2 partner are
// NOT checked (L446-459): when one of them is itself a bad pixel EARLIER in raster order, the serial loop has already rewritten it.  `upd` holds the
// values bad pixels end up with (a copy of the frame to start with); a pass recomputes every bad pixel reading earlier bad neighbours from `upd` and
// everything else from the untouched frame, and the host repeats it until a pass changes nothing: the dependencies follow raster order (a DAG), so
// the fixed point is the serial result, reached after as many passes as the longest chain of bad pixels feeding each other (one or two in practice).
__global__ void __launch_bounds__(256) k_bp_interp_xtrans(const float* __restrict__ raw, size_t rp, float* upd, size_t up, unsigned char* __restrict__ ok, const unsigned char* __restrict__ map,
                                                          size_t mp, int W, int H, const __grid_constant__ BpCfa cfa, int* __restrict__ changed)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y + 2;
    if (!(col >= 2 && col < W - 2 && row < H - 2 && map[(size_t)row * mp + col])) return;
#define RAW(r, c) raw[(size_t)(r) * rp + (c)]
#define BAD(x, y) (map[(size_t)(y) * mp + (x)] != 0)
#define XFC(r, c) cfa.m[((r) % 6) * 6 + ((c) % 6)]
    // an unchecked read: the rewritten value when (r, c) is a bad pixel the raster order reaches before (row, col)
    auto cur = [&](int r, int c) -> float {
        const bool earlier = r < row || (r == row && c < col);
        if (earlier && r >= 2 && c >= 2 && c < W - 2 && BAD(c, r)) return *(volatile const float*)(upd + (size_t)r * up + c);
        return RAW(r, c);
    };
    const float eps = 1.f;
    float wtdsum = 0.f, norm = 0.f;
    const int pixelColor = XFC(row, col);
    if (pixelColor == 1) {
        if (XFC(row, col - 1) == XFC(row, col + 1)) {
            for (int dx = -1; dx <= 1; dx += 2) {
                if (BAD(col + dx, row - 1) || BAD(col - dx, row + 1)) continue;
                const float dirwt = 0.70710678f / (fabsf(RAW(row - 1, col + dx) - RAW(row + 1, col - dx)) + eps);
                wtdsum += dirwt * (RAW(row - 1, col + dx) + RAW(row + 1, col - dx));
                norm += dirwt;
            }
            for (int dx = -1; dx <= 1; dx += 2) {
                if (BAD(col + dx, row - 2) || BAD(col - dx, row + 2)) continue;
                const float dirwt = 0.44721359f / (fabsf(RAW(row - 2, col + dx) - RAW(row + 2, col - dx)) + eps);
                wtdsum += dirwt * (RAW(row - 2, col + dx) + RAW(row + 2, col - dx));
                norm += dirwt;
            }
            for (int dx = -2; dx <= 2; dx += 4) {
                if (BAD(col + dx, row - 1) || BAD(col - dx, row + 1)) continue;
                const float dirwt = 0.44721359f / (fabsf(RAW(row - 1, col + dx) - RAW(row + 1, col - dx)) + eps);
                wtdsum += dirwt * (RAW(row - 1, col + dx) + RAW(row + 1, col - dx));
                norm += dirwt;
            }
        } else {
            const int offset1 = XFC(row - 1, col - 1) == XFC(row + 1, col + 1) ? 1 : -1;
            if (!(BAD(col - offset1, row - 1) || BAD(col + offset1, row + 1))) {
                const float dirwt = 0.70710678f / (fabsf(RAW(row - 1, col - offset1) - RAW(row + 1, col + offset1)) + eps);
                wtdsum += dirwt * (RAW(row - 1, col - offset1) + RAW(row + 1, col + offset1));
                norm += dirwt;
            }
            int offsety = XFC(row - 1, col) != 1 ? 1 : -1;
            int offsetx = offset1 * offsety;
            if (!(BAD(col + offsetx, row) || BAD(col, row + offsety))) {
                const float dirwt = 1.f / (fabsf(RAW(row, col + offsetx) - RAW(row + offsety, col)) + eps);
                wtdsum += dirwt * (RAW(row, col + offsetx) + RAW(row + offsety, col));
                norm += dirwt;
            }
            const int offsety2 = -offsety, offsetx2 = -offsetx;
            offsetx *= 2; offsety *= 2;
            if (!(BAD(col + offsetx, row + offsety2) || BAD(col + offsetx2, row + offsety))) {
                const float dirwt = 0.44721359f / (fabsf(RAW(row + offsety2, col + offsetx) - RAW(row + offsety, col + offsetx2)) + eps);
                wtdsum += dirwt * (RAW(row + offsety2, col + offsetx) + RAW(row + offsety, col + offsetx2));
                norm += dirwt;
            }
        }
    } else {
        for (int d1 = -2, offsety = 3; d1 <= 2; d1 += 4, offsety -= 6)
            for (int d2 = -1, offsetx = 3; d2 < 1; d2 += 2, offsetx -= 6)        // d2 = -1 only, as the reference's loop bound has it (L414)
                if (XFC(row + d1, col + d2) == pixelColor && !(BAD(col + d2, row + d1) || BAD(col + d2 + offsetx, row + d1 + offsety))) {
                    const float dirwt = 0.44721359f / (fabsf(RAW(row + d1, col + d2) - RAW(row + d1 + offsety, col + d2 + offsetx)) + eps);
                    wtdsum += dirwt * (RAW(row + d1, col + d2) + RAW(row + d1 + offsety, col + d2 + offsetx));
                    norm += dirwt;
                }
        bool found = false;
        int dx, dy;
        for (dx = -2, dy = 0; dx <= 2; dx += 4)
            if (XFC(row, col + dx) == pixelColor) { found = true; break; }
        if (!found)
            for (dx = 0, dy = -2; dy <= 2; dy += 4)
                if (XFC(row + dy, col) == pixelColor) { found = true; break; }
        if (found) {        // always, on an X-Trans layout (the host entry checks the 6x6 table)
            float virtualPixel;
            if (dy == 0) virtualPixel = 0.5f * (cur(row - 1, col - dx) + cur(row + 1, col - dx));
            else virtualPixel = 0.5f * (cur(row - dy, col - 1) + cur(row - dy, col + 1));
            const float partner = cur(row + dy, col + dx);
            const float dirwt = 0.5f / (fabsf(virtualPixel - partner) + eps);
            wtdsum += dirwt * (virtualPixel + partner);
            norm += dirwt;
        }
    }
#undef RAW
#undef BAD
#undef XFC
    if (norm > 0.f) {
        const float v = wtdsum / (2.f * norm);
        float* o = upd + (size_t)row * up + col;
        if (__float_as_uint(v) != __float_as_uint(*(volatile float*)o)) { *(volatile float*)o = v; *changed = 1; }
        ok[(size_t)row * mp + col] = 1;
    }
}

__global__ void __launch_bounds__(256) k_bp_commit_xtrans(float* __restrict__ raw, size_t rp, const float* __restrict__ upd, size_t up, const unsigned char* __restrict__ ok,
                                                          const unsigned char* __restrict__ map, size_t mp, int W, int H, int* __restrict__ count)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y + 2;
    const bool done = col >= 2 && col < W - 2 && row < H - 2 && map[(size_t)row * mp + col] && ok[(size_t)row * mp + col];
    if (done) raw[(size_t)row * rp + col] = upd[(size_t)row * up + col];
    const unsigned ballot = __ballot_sync(0xffffffffu, done);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count, __popc(ballot));
}

}  // namespace

// d_count: two device ints (the count, the pass's "changed" flag).  Synchronises the stream once per pass.
int art_interpolate_bad_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans36, float* raw, size_t rp, const unsigned char* map, size_t mp, int* d_count)
{
    cudaStream_t st = ctx->stream;
    ART_CUDA(ctx, cudaMemsetAsync(d_count, 0, 2 * sizeof(int), st));
    if (W < 5 || H < 5) return ART_HP_OK;
    BpCfa cfa{};
    for (int i = 0; i < 36; ++i) {
        if (xtrans36[i] < 0 || xtrans36[i] > 2) return ctx->fail(ART_HP_ERR_INVALID, "xtrans[%d] = %d", i, xtrans36[i]);
        cfa.m[i] = xtrans36[i];
    }
    // every red / blue site of an X-Trans layout has a same-colour sample at distance 2 in its row or its column; the reference's scan (L432-450)
    // runs off the frame when that does not hold
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
            const int col = xtrans36[r * 6 + c];
            if (col == 1) continue;
            const bool has = xtrans36[r * 6 + (c + 4) % 6] == col || xtrans36[r * 6 + (c + 2) % 6] == col || xtrans36[((r + 4) % 6) * 6 + c] == col || xtrans36[((r + 2) % 6) * 6 + c] == col;
            if (!has) return ctx->fail(ART_HP_ERR_INVALID, "not an X-Trans layout: no same-colour sample at distance 2 of (%d, %d)", r, c);
        }
    const size_t up = round_up((size_t)W, 32);
    void* blk = nullptr;
    int rc = art_pool_alloc(ctx, up * H * sizeof(float) + mp * (size_t)H, &blk);
    if (rc) return rc;
    float* upd = (float*)blk;
    unsigned char* ok = (unsigned char*)(upd + up * H);
    auto fail_cuda = [&](cudaError_t e, const char* what) { art_pool_free(ctx, blk); return ctx->fail(ART_HP_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e)); };
    cudaError_t e = cudaMemcpy2DAsync(upd, up * sizeof(float), raw, rp * sizeof(float), (size_t)W * sizeof(float), (size_t)H, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(ok, 0, mp * (size_t)H, st);
    if (e != cudaSuccess) return fail_cuda(e, "staging");
    const dim3 grid((W + 255) / 256, H - 4);
    int* d_changed = d_count + 1;
    for (int pass = 0;; ++pass) {
        if (pass) { e = cudaMemsetAsync(d_changed, 0, sizeof(int), st); if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync"); }
        art_prof_begin(ctx, "k_bp_interp_xtrans");
        k_bp_interp_xtrans<<<grid, 256, 0, st>>>(raw, rp, upd, up, ok, map, mp, W, H, cfa, d_changed);
        art_prof_end(ctx);
        ctx->launches++;
        int changed = 0;
        e = cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail_cuda(e, "pass");
        if (!changed) break;
        if (pass > 65536) { art_pool_free(ctx, blk); return ctx->fail(ART_HP_ERR_CUDA, "bad pixel interpolation did not settle"); }
    }
    k_bp_commit_xtrans<<<grid, 256, 0, st>>>(raw, rp, upd, up, ok, map, mp, W, H, d_count);
    ctx->launches++;
    art_pool_free(ctx, blk);
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

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
