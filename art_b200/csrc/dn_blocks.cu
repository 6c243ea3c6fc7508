// RGB_denoise detail recovery, the block stage: tcgen05 / TMEM version.
//
// Replaces (reference) rtengine/FTblockDN.cc detail_recovery L1479-1635: per 64 x 64 block at stride 25 -- gather of the
// residual with mirrored edges and the input window (L1547-1567), fftwf REDFT10 along both axes (L1604), RGBtile_denoise
// L494-525 (boxabsblur of |coefficients|, boxblur.h L745-888, and the shrink factor), fftwf REDFT01 (L1614).  The overlap-add
// (RGBoutput_tile_row L531-558) stays in k_dn_gather (denoise.cu), in the reference's one-thread order.
//
// One persistent CTA (256 threads, two per SM) handles PAIRS of horizontally adjacent blocks stacked into M = 128, so each of the four
// 64^3 products of a block pair is a 128 x 64 x 64 tcgen05.mma.kind::tf32 with the accumulator in TMEM.  fp32 accuracy comes
// from the 3xTF32 split (x = big + small; small * big + big * small + big * big, fp32 accumulation): every operand is split
// ONCE, when it is written to shared memory (the mma.sync version re-split per fragment load, which was 5 of every 7
// instructions of its products).  Threads t and t + 128 share TMEM lane t = row t of every product's result, i.e. (block t / 64,
// index t % 64), and take 32 of its 64 columns each.
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

// 2^x, x <= 0, by the special-function unit (results below 2^-126 flush to zero: the factor is then 1 either way)
__device__ __forceinline__ float ex2_fast(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// element (row = blk * 64 + m, K = q): koff carries the thread's (blk, q) part
__device__ __forceinline__ void put(unsigned char* big, unsigned char* small, unsigned koff, int m, float v)
{
    // big = v truncated to TF32 (what the tensor core would read anyway), small = the exact remainder rounded to nearest: the only
    // rounding of the pair is the small part's (2^-22 of v, unbiased)
    const unsigned b = __float_as_uint(v) & 0xffffe000u;
    const unsigned s = umma::rn_tf32(v - __uint_as_float(b));
    const unsigned off = koff + (m & 7) * 16 + (m >> 3) * A_SBO;
    *reinterpret_cast<unsigned*>(big + off) = b;
    *reinterpret_cast<unsigned*>(small + off) = s;
}

// 32 accumulator columns of this thread's lane
__device__ __forceinline__ void ld32(unsigned taddr, float (&v)[32])
{
    unsigned r0[16], r1[16];
    umma::tmem_ld16_nowait(taddr, r0);
    umma::tmem_ld16_nowait(taddr + 16, r1);
    umma::wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i] = __uint_as_float(r0[i]);
        v[16 + i] = __uint_as_float(r1[i]);
    }
}

// The CTA waits for its product: one thread on the mbarrier, the others at the hardware barrier (no issue slots spent spinning)
__device__ __forceinline__ void mma_done(unsigned bar, unsigned& parity)
{
    if (threadIdx.x == 0) umma::mbar_wait(bar, parity);
    parity ^= 1;
    __syncthreads();
    umma::fence_after_sync();
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

// One pass of boxabsblur (boxblur.h L745-888) over a 64-vector (the |.| already taken), same expressions in both directions, cut
// in two halves of 32 outputs so that two threads share a line.  The lower half is the reference's running mean as it stands; the
// upper half restarts the running mean at element 32 from the 2 RAD + 1 samples under the window (the same value up to the
// rounding the running form has accumulated by then, ~1e-7 relative) and finishes with the reference's shrinking window.
// lo: s[j] = sample j, j = 0 .. 31 + RAD.   hi: s[j] = sample 31 - RAD + j, j = 0 .. 32 + RAD.
template <int RAD>
__device__ __forceinline__ void box_lo(const float (&s)[32 + RAD], float (&o)[32])
{
    float v = s[0];
#pragma unroll
    for (int j = 1; j <= RAD; ++j) v = v + s[j];
    v = v * (1.f / (float)(RAD + 1));         // the window lengths are compile-time constants: reciprocals, not IEEE divisions
    o[0] = v;
#pragma unroll
    for (int c = 1; c <= RAD; ++c) { v = (v * (float)(RAD + c) + s[c + RAD]) * (1.f / (float)(RAD + c + 1)); o[c] = v; }
    constexpr float rlen = 1.f / (float)(2 * RAD + 1);
#pragma unroll
    for (int c = RAD + 1; c < 32; ++c) { v = v + (s[c + RAD] - s[c - RAD - 1]) * rlen; o[c] = v; }
}
template <int RAD>
__device__ __forceinline__ void box_hi(const float (&s)[33 + RAD], float (&o)[32])
{
    constexpr int B = 31 - RAD;              // sample index of s[0]
    constexpr float rlen = 1.f / (float)(2 * RAD + 1);
    float v = s[32 - RAD - B];
#pragma unroll
    for (int j = 33 - RAD; j <= 32 + RAD; ++j) v = v + s[j - B];
    v = v * rlen;
    o[0] = v;
#pragma unroll
    for (int c = 33; c < TS - RAD; ++c) { v = v + (s[c + RAD - B] - s[c - RAD - 1 - B]) * rlen; o[c - 32] = v; }
#pragma unroll
    for (int c = TS - RAD; c < TS; ++c) {          // window length TS - 1 - c + RAD + 1 shrinking from 2 RAD + 1
        const int len = TS - c + RAD + 1;
        v = (v * (float)len - s[c - RAD - 1 - B]) * (1.f / (float)(len - 1));
        o[c - 32] = v;
    }
}

template <int RAD>
__global__ void __launch_bounds__(256, 2) k_dn_blocks5(DnBlocksArgs a, int npw, int npairs)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ unsigned tmem_slot;
    unsigned char* a_big = sm;
    unsigned char* a_small = sm + A_BYTES;
    unsigned char* b_big = sm + 2 * A_BYTES;
    float* S = reinterpret_cast<float*>(a_big);
    // thread = (row owner r = TMEM lane, half hh of the 64 columns); r = (block of the pair, index q)
    const int t = threadIdx.x, r = t & 127, hh = t >> 7, blk = r >> 6, q = r & 63, c0 = 32 * hh;
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
    const unsigned tm = tmem + ((unsigned)(((t >> 5) & 3) * 32) << 16) + (unsigned)c0;       // this warp's lane quadrant, this thread's column half
    unsigned par_m = 0, par_b = 0;
    // operand element (row = blk * 64 + m, K = q) for m = c0 + u: everything but u in koff
    const unsigned koff = (unsigned)(blk * 8 + 4 * hh) * A_SBO + (unsigned)(q >> 2) * A_LBO + (unsigned)(q & 3) * 4;
    const float inv_hi = -1.4426950408889634f / a.detail_hi, inv_lo = -1.4426950408889634f / a.detail_lo;      // exp(x) = 2^(x log2 e)
    if (t == 0 && (int)blockIdx.x < npairs) {
        umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
        umma::bulk_g2s(sb_big, a.fwd_split, 2 * B_BYTES, bar_b);
    }

    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int vblk = pair / npw, hblk = 2 * (pair - vblk * npw) + blk;
        const bool valid = hblk < a.nbw;         // an odd block count leaves the last pair half empty: computed, not stored
        const int top = (vblk - BLKRAD) * OFFSET, left = (hblk - BLKRAD) * OFFSET;

        // ---- phase 0: residual x window, column q of the block, rows c0 .. c0 + 31 (L1547-1567: mirror without repeating the edge, clamped)
        {
            int col = left + q;
            if (col < 0) col = min(-col, a.width - 1);
            else if (col >= a.width) col = max(0, 2 * a.width - 2 - col);
            // sixteen rows of loads in flight before the first use; blocks that do not touch the frame's top / bottom (all but two block rows)
            // walk the column by pointer, the others mirror each row (branch-free, so that the loads still batch)
            const float* wp = a.tin + c0 * TS + q;
            if (top >= 0 && top + TS <= a.height) {
                const float* pl = a.Lin + (size_t)(top + c0) * a.width + col;
                const float* po = a.L + (size_t)(top + c0) * a.width + col;
#pragma unroll
                for (int i0 = 0; i0 < 32; i0 += 16) {
                    float lin[16], lo[16], win[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        lin[u] = __ldg(pl + (size_t)(i0 + u) * a.width);
                        lo[u] = __ldg(po + (size_t)(i0 + u) * a.width);
                        win[u] = __ldg(wp + (i0 + u) * TS);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) put(a_big, a_small, koff, i0 + u, win[u] * (lin[u] - lo[u]));
                }
            } else {
#pragma unroll
                for (int i0 = 0; i0 < 32; i0 += 16) {
                    float lin[16], lo[16], win[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int row = top + c0 + i0 + u;
                        const int below = min(-row, a.height - 1), above = max(0, 2 * a.height - 2 - row);
                        const int rr = row < 0 ? below : (row >= a.height ? above : row);
                        const size_t p = (size_t)rr * a.width + col;
                        lin[u] = __ldg(a.Lin + p);
                        lo[u] = __ldg(a.L + p);
                        win[u] = __ldg(wp + (i0 + u) * TS);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) put(a_big, a_small, koff, i0 + u, win[u] * (lin[u] - lo[u]));
                }
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
        mma_done(bar_m, par_m);
        // ---- phase 1: row i = q of T, transposed into the operand of product 2
        {
            float v[32];
            ld32(tm, v);
#pragma unroll
            for (int n = 0; n < 32; ++n) put(a_big, a_small, koff, n, v[n]);
        }
        umma::fence_before_sync();
        umma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            issue_product(tmem, sa_big, sa_small, sb_big, sb_small, bar_m);
        }
        mma_done(bar_m, par_m);
        if (t == 0) {       // the forward matrices are consumed: the backward ones stream in under the blur
            umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
            umma::bulk_g2s(sb_big, a.bwd_split, 2 * B_BYTES, bar_b);
        }
        // ---- phase 2: this thread holds Y[k][k' = q] for k = c0 .. c0 + 31.  boxabsblur: along k' (threads) through the scratch, then along k (registers)
        {
            float* srow = S + (size_t)(blk * TS + q) * SP;       // scratch row q of this block: element k
            float* scol = S + (size_t)(blk * TS) * SP + q;       // scratch column q: element k' at stride SP
            {
                float y[32];
                ld32(tm, y);
#pragma unroll
                for (int k = 0; k < 32; ++k) srow[c0 + k] = fabsf(y[k]);
            }
            __syncthreads();
            {   // as the owner of coefficient row k = q: the horizontal pass over k' = c0 .. c0 + 31, in place
                float h[32];
                if (hh == 0) {
                    float s[32 + RAD];
#pragma unroll
                    for (int c = 0; c < 32 + RAD; ++c) s[c] = scol[c * SP];
                    box_lo<RAD>(s, h);
                } else {
                    float s[33 + RAD];
#pragma unroll
                    for (int c = 0; c < 33 + RAD; ++c) s[c] = scol[(31 - RAD + c) * SP];
                    box_hi<RAD>(s, h);
                }
                __syncthreads();        // every window is read before any element is replaced
#pragma unroll
                for (int c = 0; c < 32; ++c) scol[(c0 + c) * SP] = h[c];
            }
            __syncthreads();
            float nb[32];
            if (hh == 0) {
                float s[32 + RAD];
#pragma unroll
                for (int k = 0; k < 32 + RAD; ++k) s[k] = srow[k];
                box_lo<RAD>(s, nb);       // the vertical pass
            } else {
                float s[33 + RAD];
#pragma unroll
                for (int k = 0; k < 33 + RAD; ++k) s[k] = srow[31 - RAD + k];
                box_hi<RAD>(s, nb);
            }
            __syncthreads();         // the scratch (= the big operand buffer) is free again
            float y[32];
            ld32(tm, y);
            // RGBtile_denoise L511-514 with the per-sample detail factor of L1571-1595.  The coefficients already differ from the reference's in
            // their last bits (tensor-core products here, FFTW there), so the factor uses the fast reciprocal / exp2 instead of sleef exp + IEEE division.
            const int icol = left + q;
            const bool col_in = valid && icol >= 0 && icol < a.width;
            if (!a.use_mask) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const int row = top + c0 + k;
                    const float idf = (col_in && row >= 0 && row < a.height) ? inv_hi : inv_lo;
                    put(a_big, a_small, koff, k, y[k] * (1.0f - ex2_fast((nb[k] * nb[k]) * idf)));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const int row = top + c0 + k;
                    float idf = inv_lo;
                    if (col_in && row >= 0 && row < a.height)
                        idf = __fdividef(-1.4426950408889634f, compute_detail(a.params_Ldetail * __ldg(a.mask + (size_t)row * a.width + icol)));
                    put(a_big, a_small, koff, k, y[k] * (1.0f - ex2_fast((nb[k] * nb[k]) * idf)));
                }
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
        mma_done(bar_m, par_m);
        // ---- phase 3: row k = q of U, transposed into the operand of product 4
        {
            float v[32];
            ld32(tm, v);
#pragma unroll
            for (int n = 0; n < 32; ++n) put(a_big, a_small, koff, n, v[n]);
        }
        umma::fence_before_sync();
        umma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            issue_product(tmem, sa_big, sa_small, sb_big, sb_small, bar_m);
        }
        mma_done(bar_m, par_m);
        if (t == 0 && pair + (int)gridDim.x < npairs) {       // forward matrices for the next pair
            umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
            umma::bulk_g2s(sb_big, a.fwd_split, 2 * B_BYTES, bar_b);
        }
        // ---- phase 4: this thread holds Z[y][x = q] for y = c0 .. c0 + 31: rows of 32 consecutive floats per warp
        {
            float z[32];
            ld32(tm, z);
            if (valid) {
                float* out = a.blocks + ((size_t)vblk * a.nbw + hblk) * (TS * TS) + (size_t)c0 * TS + q;
#pragma unroll
                for (int y = 0; y < 32; ++y) out[y * TS] = z[y];
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
    case 1: k_dn_blocks5<1><<<grid, 256, SMEM_BYTES, ctx->stream>>>(a, npw, npairs); break;
    case 2: k_dn_blocks5<2><<<grid, 256, SMEM_BYTES, ctx->stream>>>(a, npw, npairs); break;
    case 3: k_dn_blocks5<3><<<grid, 256, SMEM_BYTES, ctx->stream>>>(a, npw, npairs); break;
    default: return ctx->fail(ART_HP_ERR_INVALID, "detail recovery: blur radius %d outside 1..3", a.blur_rad);
    }
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
