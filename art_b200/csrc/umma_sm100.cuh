// tcgen05 / TMEM / mbarrier wrappers for sm_100a (inline PTX, no CUTLASS).
//
// Used by the block-DCT kernel of RGB_denoise's detail recovery (denoise.cu: k_dn_blocks5) and by tools/umma_probe.cu, which checks
// every assumption below (descriptor fields, the canonical shared-memory layout, the TMEM lane / column mapping) on the device
// against a double-precision product before the kernel relies on it.
//
// Operand layout: K-major, no swizzle.  An operand with MN rows and K 32-bit elements per row is stored as "core matrices" of
// 8 rows x 16 bytes (4 elements), 128 contiguous bytes each:
//     byte(row, k) = (row % 8) * 16 + (row / 8) * SBO + (k / 4) * LBO + (k % 4) * 4
// SBO = byte stride between 8-row groups, LBO = byte stride between 16-byte chunks along K; both multiples of 16.  One
// tcgen05.mma.kind::tf32 consumes K = 8 elements (two chunks); the next K step starts 2 * LBO further.
// Accumulator layout (cta_group::1, M = 128): row m of D is TMEM lane m, column n is TMEM column n (32-bit each); warp w of a
// CTA can read lanes 32 * (w % 4) .. + 32 only, so thread t of a 128-thread CTA owns row t.
#pragma once
#include <cstdint>

namespace umma {

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a barrier that never completes is a bug in the pipeline, and a trap (the launch fails, the caller sees the
// CUDA error) is better than a hung device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    for (unsigned spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
// 1-D bulk copy global -> shared by the TMA engine, completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------------------
// one full warp; the base address (lane 0, first column) lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(unsigned slot, unsigned cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(unsigned base, unsigned cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 consecutive columns of this thread's lane (warp-wide: lanes 32 * (warp % 4) ..)
__device__ __forceinline__ void tmem_ld16(unsigned addr, float (&v)[16])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr) : "memory");
    wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the same, without the wait: the caller issues several loads and waits once
__device__ __forceinline__ void tmem_ld16_nowait(unsigned addr, unsigned (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, no swizzle (layout type 0), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor for kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0),
// N / 8 in bits 17-22, M / 16 in bits 24-28
__host__ __device__ constexpr unsigned idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_tf32(unsigned d_tmem, uint64_t a_desc, uint64_t b_desc, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of every tcgen05.mma this thread issued so far -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// ---- 3xTF32 -----------------------------------------------------------------------------------------------------------
// x = big + small, both rounded to TF32 (10 mantissa bits) to nearest by integer arithmetic on the bit pattern; x - big is exact.
// (cvt.rna.tf32.f32 is emulated with a dozen instructions on sm_100a.)  The tensor core ignores the 13 low mantissa bits, which are zero here.
__device__ __forceinline__ unsigned rn_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, unsigned& big, unsigned& small)
{
    big = rn_tf32(x);
    small = rn_tf32(x - __uint_as_float(big));
}

}  // namespace umma
