// RGB_denoise detail recovery, the block stage: tcgen05 / TMEM version.
//
// Replaces (reference) rtengine/FTblockDN.cc detail_recovery L1479-1635: per 64 x 64 block at stride 25 -- gather of the
// residual with mirrored edges and the input window (L1547-1567), fftwf REDFT10 along both axes (L1604), RGBtile_denoise
// L494-525 (boxabsblur of |coefficients|, boxblur.h L745-888, and the shrink factor), fftwf REDFT01 (L1614).  The overlap-add
// (RGBoutput_tile_row L531-558) stays in k_dn_gather (denoise.cu), in the reference's one-thread order.
//
// One persistent CTA (128 threads, two per SM) handles PAIRS of horizontally adjacent blocks stacked into M = 128, so each of the four
// 64^3 products of a block pair is a 128 x 64 x 64 tcgen05.mma.kind::tf32 with the accumulator in TMEM.  fp32 accuracy comes
// from the 3xTF32 split (x = big + small; small * big + big * small + big * big, fp32 accumulation): every operand is split
// ONCE, when it is written to shared memory (the mma.sync version re-split per fragment load, which was 5 of every 7
// instructions of its products).  Thread t owns TMEM lane t = row t of every product's result, i.e. (block t / 64, index t % 64).
//
// The chain is arranged so that the moving operand is always A and every hand-over is the same transposing write:
//   product 1  T [i][k'] = sum_j  X[i][j]  C[k'][j]      A(row = i,  K = j)  written by the gather thread of column j
//   product 2  Y'[k'][k] = sum_i  T[i][k'] C[k][i]       A(row = k', K = i)  written by the owner of row i of T
//   product 3  U [k][x]  = sum_k' Y[k][k'] D[x][k']      A(row = k,  K = k') written by the owner of row k' of Y' (after the shrink)
//   product 4  Z'[x][y]  = sum_k  U[k][x]  D[y][k]       A(row = x,  K = k)  written by the owner of row k of U
// so B is the forward matrix C[n][k] for products 1-2 and the backward matrix D[n][k] for 3-4, both pre-split on the host
// and brought in by one bulk copy (TMA engine) each, overlapped with the phases that do not need them.  A thread writing
// "its" K index for 64 rows hits 32 distinct banks per warp because the K-chunk stride (LBO) is padded by 16 bytes.
// The box blur runs in the reference's order (along k' first, then along k) through a pitch-65 scratch in the A buffer.
#include <cmath>
#include "dn_blocks.h"
#include "umma_sm100.cuh"

namespace {

constexpr int TS = 64, OFFSET = 25, BLKRAD = 1;
constexpr unsigned A_SBO = 128, A_LBO = 16 * 128 + 16;     // 16 row groups per K chunk, + 16 bytes: chunk stride == 4 words mod 32
constexpr unsigned B_SBO = 128, B_LBO = 8 * 128;
constexpr int SP = 65;                                     // pitch of the blur scratch (floats)
constexpr unsigned A_BYTES = 128 * SP * 4;                 // 33280 >= 16 * A_LBO = 33024
constexpr unsigned B_BYTES = 16 * B_LBO;                   // 16384 per half
constexpr unsigned SMEM_BYTES = 2 * A_BYTES + 2 * B_BYTES; // 99328: two CTAs per SM
static_assert(A_BYTES >= 16 * A_LBO, "operand image must fit the A buffer");
constexpr unsigned TMEM_COLS = 64;

__device__ __forceinline__ float compute_detail(float d)
{   // L1481-1485
    const float a = (float)(((100. - d) * (100. - d)) + 50. * (100. - d)) * TS * 0.5f;
    return a * a;
}

// element (row = blk * 64 + m, K = q): koff carries the thread's (blk, q) part
__device__ __forceinline__ void put(unsigned char* big, unsigned char* small, unsigned koff, int m, float v)
{
    unsigned b, s;
    umma::split_tf32(v, b, s);
    const unsigned off = koff + (m & 7) * 16 + (m >> 3) * A_SBO;
    *reinterpret_cast<unsigned*>(big + off) = b;
    *reinterpret_cast<unsigned*>(small + off) = s;
}

// this thread's 64 accumulator columns
__device__ __forceinline__ void ld64(unsigned taddr, float (&v)[64])
{
    unsigned r0[16], r1[16], r2[16], r3[16];
    umma::tmem_ld16_nowait(taddr, r0);
    umma::tmem_ld16_nowait(taddr + 16, r1);
    umma::tmem_ld16_nowait(taddr + 32, r2);
    umma::tmem_ld16_nowait(taddr + 48, r3);
    umma::wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i] = __uint_as_float(r0[i]);
        v[16 + i] = __uint_as_float(r1[i]);
        v[32 + i] = __uint_as_float(r2[i]);
        v[48 + i] = __uint_as_float(r3[i]);
    }
}

// 24 instructions: three passes (small * big, big * small, big * big) of eight K steps
__device__ __forceinline__ void issue_product(unsigned tmem, unsigned a_big, unsigned a_small, unsigned b_big, unsigned b_small, unsigned bar)
{
    constexpr unsigned idesc = umma::idesc_tf32(128, 64);
    unsigned acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const unsigned pa = pass == 0 ? a_small : a_big, pb = pass == 1 ? b_small : b_big;
#pragma unroll
        for (int ks = 0; ks < TS / 8; ++ks) {
            umma::mma_tf32(tmem, umma::smem_desc(pa + ks * 2 * A_LBO, A_LBO, A_SBO), umma::smem_desc(pb + ks * 2 * B_LBO, B_LBO, B_SBO), idesc, acc);
            acc = 1;
        }
    }
    umma::mma_commit(bar);
}

// one pass of boxabsblur (boxblur.h L745-888) over a 64-vector (the |.| already taken), same expressions in both directions
template <int RAD>
__device__ __forceinline__ void box64(const float (&s)[64], float (&o)[64])
{
    float len = (float)(RAD + 1);
    float v = s[0];
#pragma unroll
    for (int j = 1; j <= RAD; ++j) v = v + s[j];
    v = v / len;
    o[0] = v;
#pragma unroll
    for (int c = 1; c <= RAD; ++c) { const float lp1 = len + 1.f; v = (v * len + s[c + RAD]) / lp1; o[c] = v; len = lp1; }
    const float rlen = 1.f / len;
#pragma unroll
    for (int c = RAD + 1; c < TS - RAD; ++c) { v = v + (s[c + RAD] - s[c - RAD - 1]) * rlen; o[c] = v; }
#pragma unroll
    for (int c = TS - RAD; c < TS; ++c) { const float lm1 = len - 1.f; v = (v * len - s[c - RAD - 1]) / lm1; o[c] = v; len = lm1; }
}

template <int RAD>
__global__ void __launch_bounds__(128, 2) k_dn_blocks5(DnBlocksArgs a, int npw, int npairs)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ unsigned tmem_slot;
    unsigned char* a_big = sm;
    unsigned char* a_small = sm + A_BYTES;
    unsigned char* b_big = sm + 2 * A_BYTES;
    float* S = reinterpret_cast<float*>(a_big);
    const int t = threadIdx.x, blk = t >> 6, q = t & 63;
    const unsigned bar_b = umma::smem_addr(&bars[0]), bar_m = umma::smem_addr(&bars[1]);
    const unsigned sa_big = umma::smem_addr(a_big), sa_small = umma::smem_addr(a_small);
    const unsigned sb_big = umma::smem_addr(b_big), sb_small = sb_big + B_BYTES;
    if (t == 0) {
        umma::mbar_init(bar_b, 1);
        umma::mbar_init(bar_m, 1);
        umma::mbar_fence_init();
    }
    if (t < 32) umma::tmem_alloc(umma::smem_addr(&tmem_slot), TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const unsigned tmem = tmem_slot;
    const unsigned tm = tmem + ((unsigned)(((t >> 5) & 3) * 32) << 16);       // this warp's lane quadrant
    unsigned par_m = 0, par_b = 0;
    const unsigned koff = (unsigned)blk * 8 * A_SBO + (unsigned)(q >> 2) * A_LBO + (unsigned)(q & 3) * 4;
    const float inv_hi = -1.4426950408889634f / a.detail_hi, inv_lo = -1.4426950408889634f / a.detail_lo;      // exp(x) = 2^(x log2 e)
    if (t == 0 && (int)blockIdx.x < npairs) {
        umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
        umma::bulk_g2s(sb_big, a.fwd_split, 2 * B_BYTES, bar_b);
    }

    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int vblk = pair / npw, hblk = 2 * (pair - vblk * npw) + blk;
        const bool valid = hblk < a.nbw;         // an odd block count leaves the last pair half empty: computed, not stored
        const int top = (vblk - BLKRAD) * OFFSET, left = (hblk - BLKRAD) * OFFSET;

        // ---- phase 0: residual x window for column q of the block (L1547-1567: mirror without repeating the edge, clamped)
        {
            int col = left + q;
            if (col < 0) col = min(-col, a.width - 1);
            else if (col >= a.width) col = max(0, 2 * a.width - 2 - col);
            // sixteen rows of loads in flight before the first use (the row mirror is branch-free so that the loads can be batched)
#pragma unroll 1
            for (int i0 = 0; i0 < TS; i0 += 16) {
                float lin[16], lo[16], win[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int row = top + i0 + u;
                    const int below = min(-row, a.height - 1), above = max(0, 2 * a.height - 2 - row);
                    const int rr = row < 0 ? below : (row >= a.height ? above : row);
                    const size_t p = (size_t)rr * a.width + col;
                    lin[u] = __ldg(a.Lin + p);
                    lo[u] = __ldg(a.L + p);
                    win[u] = __ldg(a.tin + (i0 + u) * TS + q);
                }
                unsigned char* pb = a_big + (i0 >> 3) * A_SBO;
                unsigned char* ps = a_small + (i0 >> 3) * A_SBO;
#pragma unroll
                for (int u = 0; u < 16; ++u) put(pb, ps, koff, u, win[u] * (lin[u] - lo[u]));
            }
        }
        // ---- product 1 (needs the forward matrices)
        umma::fence_before_sync();
        umma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            umma::mbar_wait(bar_b, par_b); par_b ^= 1;
            umma::fence_after_sync();
            issue_product(tmem, sa_big, sa_small, sb_big, sb_small, bar_m);
        }
        umma::mbar_wait(bar_m, par_m); par_m ^= 1;
        umma::fence_after_sync();
        // ---- phase 1: row i = q of T, transposed into the operand of product 2
        {
            float v[64];
            ld64(tm, v);
#pragma unroll
            for (int n = 0; n < TS; ++n) put(a_big, a_small, koff, n, v[n]);
        }
        umma::fence_before_sync();
        umma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            issue_product(tmem, sa_big, sa_small, sb_big, sb_small, bar_m);
        }
        umma::mbar_wait(bar_m, par_m); par_m ^= 1;
        umma::fence_after_sync();
        if (t == 0) {       // the forward matrices are consumed: the backward ones stream in under the blur
            umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
            umma::bulk_g2s(sb_big, a.bwd_split, 2 * B_BYTES, bar_b);
        }
        // ---- phase 2: this thread holds Y[k][k' = q] for all k.  boxabsblur: along k' (threads) through the scratch, then along k (registers)
        {
            float* srow = S + (size_t)(blk * TS + q) * SP;       // scratch row q of this block: element k
            float* scol = S + (size_t)(blk * TS) * SP + q;       // scratch column q: element k' at stride SP
            {
                float y[64];
                ld64(tm, y);
#pragma unroll
                for (int k = 0; k < TS; ++k) srow[k] = fabsf(y[k]);
            }
            __syncthreads();
            {   // as the owner of coefficient row k = q: the horizontal pass, in place
                float s[64], h[64];
#pragma unroll
                for (int c = 0; c < TS; ++c) s[c] = scol[c * SP];
                box64<RAD>(s, h);
#pragma unroll
                for (int c = 0; c < TS; ++c) scol[c * SP] = h[c];
            }
            __syncthreads();
            float h[64];
#pragma unroll
            for (int k = 0; k < TS; ++k) h[k] = srow[k];
            __syncthreads();         // the scratch (= the big operand buffer) is free again
            float nb[64];
            box64<RAD>(h, nb);       // the vertical pass
            float y[64];
            ld64(tm, y);
            // RGBtile_denoise L511-514 with the per-sample detail factor of L1571-1595.  The coefficients already differ from the reference's in
            // their last bits (tensor-core products here, FFTW there), so the factor uses the fast reciprocal / exp2 instead of sleef exp + IEEE division.
            const int icol = left + q;
            const bool col_in = valid && icol >= 0 && icol < a.width;
#pragma unroll
            for (int k = 0; k < TS; ++k) {
                const int row = top + k;
                float idf = inv_lo;
                if (col_in && row >= 0 && row < a.height)
                    idf = a.use_mask ? __fdividef(-1.4426950408889634f, compute_detail(a.params_Ldetail * a.mask[(size_t)row * a.width + icol])) : inv_hi;
                put(a_big, a_small, koff, k, y[k] * (1.0f - exp2f((nb[k] * nb[k]) * idf)));
            }
        }
        // ---- product 3 (needs the backward matrices)
        umma::fence_before_sync();
        umma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            umma::mbar_wait(bar_b, par_b); par_b ^= 1;
            umma::fence_after_sync();
            issue_product(tmem, sa_big, sa_small, sb_big, sb_small, bar_m);
        }
        umma::mbar_wait(bar_m, par_m); par_m ^= 1;
        umma::fence_after_sync();
        // ---- phase 3: row k = q of U, transposed into the operand of product 4
        {
            float v[64];
            ld64(tm, v);
#pragma unroll
            for (int n = 0; n < TS; ++n) put(a_big, a_small, koff, n, v[n]);
        }
        umma::fence_before_sync();
        umma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            issue_product(tmem, sa_big, sa_small, sb_big, sb_small, bar_m);
        }
        umma::mbar_wait(bar_m, par_m); par_m ^= 1;
        umma::fence_after_sync();
        if (t == 0 && pair + (int)gridDim.x < npairs) {       // forward matrices for the next pair
            umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
            umma::bulk_g2s(sb_big, a.fwd_split, 2 * B_BYTES, bar_b);
        }
        // ---- phase 4: this thread holds Z[y][x = q] for all y: rows of 32 consecutive floats per warp
        {
            float z[64];
            ld64(tm, z);
            if (valid) {
                float* out = a.blocks + ((size_t)vblk * a.nbw + hblk) * (TS * TS) + q;
#pragma unroll
                for (int y = 0; y < TS; ++y) out[y * TS] = z[y];
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (t < 32) umma::tmem_free(tmem, TMEM_COLS);
}

}  // namespace

void art_dn_blocks_split_tables(const float* dctf, const float* dctb, unsigned* host)
{
    const float* src[2] = {dctf, dctb};
    for (int m = 0; m < 2; ++m) {
        unsigned* big = host + (size_t)m * 2 * (B_BYTES / 4);
        unsigned* small = big + B_BYTES / 4;
        for (int n = 0; n < TS; ++n)
            for (int k = 0; k < TS; ++k) {
                const float x = src[m][n * TS + k];
                unsigned xb, rb;
                std::memcpy(&xb, &x, 4);
                const unsigned b = (xb + 0x1000u) & 0xffffe000u;
                float bf;
                std::memcpy(&bf, &b, 4);
                const float r = x - bf;
                std::memcpy(&rb, &r, 4);
                const unsigned off = ((n & 7) * 16 + (n >> 3) * B_SBO + (k >> 2) * B_LBO + (k & 3) * 4) / 4;
                big[off] = b;
                small[off] = (rb + 0x1000u) & 0xffffe000u;
            }
    }
}

int art_dn_blocks_launch(art_hp_ctx* ctx, const DnBlocksArgs& a)
{
    const int npw = (a.nbw + 1) / 2, npairs = npw * a.nbh;
    const int grid = std::min(npairs, 2 * ctx->sm_count);
    if (!(ctx->attrs_set & art_hp_ctx::ATTR_DN_BLOCKS)) {
        ART_CUDA(ctx, cudaFuncSetAttribute(k_dn_blocks5<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        ART_CUDA(ctx, cudaFuncSetAttribute(k_dn_blocks5<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        ART_CUDA(ctx, cudaFuncSetAttribute(k_dn_blocks5<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        ctx->attrs_set |= art_hp_ctx::ATTR_DN_BLOCKS;
    }
    switch (a.blur_rad) {
    case 1: k_dn_blocks5<1><<<grid, 128, SMEM_BYTES, ctx->stream>>>(a, npw, npairs); break;
    case 2: k_dn_blocks5<2><<<grid, 128, SMEM_BYTES, ctx->stream>>>(a, npw, npairs); break;
    case 3: k_dn_blocks5<3><<<grid, 128, SMEM_BYTES, ctx->stream>>>(a, npw, npairs); break;
    default: return ctx->fail(ART_HP_ERR_INVALID, "detail recovery: blur radius %d outside 1..3", a.blur_rad);
    }
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
