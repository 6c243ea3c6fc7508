// Device check of the tcgen05 building blocks art_b200/csrc/umma_sm100.cuh documents: the K-major no-swizzle operand layout (LBO / SBO),
// the kind::tf32 instruction descriptor, the M = 128 accumulator layout in TMEM and the 3xTF32 split, on one 128 x 64 x 64 product
// against a double-precision reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
// Run: tools/umma_probe [variant]   variant 0 = the documented layout, 1 = LBO / SBO fields exchanged, 2 = one pass (plain TF32)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../art_b200/csrc/umma_sm100.cuh"

constexpr int M = 128, N = 64, K = 64;
constexpr unsigned A_SBO = 128, A_LBO = 16 * 128 + 16;      // 16 row groups of 128 bytes per K chunk, + 16 bytes against bank conflicts
constexpr unsigned B_SBO = 128, B_LBO = 8 * 128;
constexpr unsigned A_BYTES = 16 * A_LBO, B_BYTES = 16 * B_LBO;

__global__ void __launch_bounds__(128) k_probe(const float* A, const unsigned* Bsplit, float* out, int variant)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ unsigned tmem_slot;
    unsigned char* a_big = sm;
    unsigned char* a_small = a_big + A_BYTES;
    unsigned char* b_big = a_small + A_BYTES;
    unsigned char* b_small = b_big + B_BYTES;
    const int t = threadIdx.x;
    const unsigned bar_b = umma::smem_addr(&bars[0]), bar_m = umma::smem_addr(&bars[1]);
    if (t == 0) {
        umma::mbar_init(bar_b, 1);
        umma::mbar_init(bar_m, 1);
        umma::mbar_fence_init();
    }
    if (t < 32) umma::tmem_alloc(umma::smem_addr(&tmem_slot), 64);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const unsigned tmem = tmem_slot;
    if (t == 0) {
        umma::mbar_expect_tx(bar_b, 2 * B_BYTES);
        umma::bulk_g2s(umma::smem_addr(b_big), Bsplit, 2 * B_BYTES, bar_b);
    }
    // thread (blk, q) holds A[(blk, m)][q] for m = 0 .. 63 and writes element (row = blk * 64 + m, k = q)
    const int blk = t >> 6, q = t & 63;
    for (int m = 0; m < 64; ++m) {
        const int row = blk * 64 + m;
        unsigned big, small;
        umma::split_tf32(A[row * K + q], big, small);
        const unsigned off = (row & 7) * 16 + (row >> 3) * A_SBO + (q >> 2) * A_LBO + (q & 3) * 4;
        *(unsigned*)(a_big + off) = big;
        *(unsigned*)(a_small + off) = small;
    }
    umma::fence_async_smem();
    __syncthreads();
    if (t == 0) {
        umma::mbar_wait(bar_b, 0);
        umma::fence_after_sync();
        const unsigned idesc = umma::idesc_tf32(M, N);
        const unsigned la = variant == 1 ? A_SBO : A_LBO, sa = variant == 1 ? A_LBO : A_SBO;
        const unsigned lb = variant == 1 ? B_SBO : B_LBO, sb = variant == 1 ? B_LBO : B_SBO;
        unsigned acc = 0;
        for (int pass = (variant == 2 ? 2 : 0); pass < 3; ++pass) {      // small * big, big * small, big * big
            const unsigned char* pa = pass == 0 ? a_small : a_big;
            const unsigned char* pb = pass == 1 ? b_small : b_big;
            for (int ks = 0; ks < K / 8; ++ks) {
                umma::mma_tf32(tmem, umma::smem_desc(umma::smem_addr(pa) + ks * 2 * A_LBO, la, sa),
                               umma::smem_desc(umma::smem_addr(pb) + ks * 2 * B_LBO, lb, sb), idesc, acc);
                acc = 1;
            }
        }
        umma::mma_commit(bar_m);
    }
    umma::mbar_wait(bar_m, 0);
    umma::fence_after_sync();
    const unsigned lane_base = (unsigned)((t >> 5) & 3) * 32u;
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tmem + (lane_base << 16) + c0, v);
        for (int i = 0; i < 16; ++i) out[t * N + c0 + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (t < 32) umma::tmem_free(tmem, 64);
}

int main(int argc, char** argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<float> A(M * K), B(N * K), out(M * N, -1.f);
    srand(7);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 2000.f - 1000.f;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) B[n * K + k] = (float)(2.0 * cos(M_PI * (k + 0.5) * n / 64));
    // B pre-split in the canonical layout: element (n, k) -> (n % 8) * 16 + (n / 8) * SBO + (k / 4) * LBO + (k % 4) * 4
    std::vector<unsigned> Bs(2 * B_BYTES / 4, 0);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float x = B[n * K + k];
            unsigned xb; memcpy(&xb, &x, 4);
            const unsigned big = (xb + 0x1000u) & 0xffffe000u;
            float bf; memcpy(&bf, &big, 4);
            const float r = x - bf;
            unsigned rb; memcpy(&rb, &r, 4);
            const unsigned small = (rb + 0x1000u) & 0xffffe000u;
            const unsigned off = ((n & 7) * 16 + (n >> 3) * B_SBO + (k >> 2) * B_LBO + (k & 3) * 4) / 4;
            Bs[off] = big;
            Bs[B_BYTES / 4 + off] = small;
        }
    float *dA, *dout; unsigned* dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bs.size() * 4); cudaMalloc(&dout, out.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bs.data(), Bs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, out.size() * 4);
    const size_t smem = 2 * A_BYTES + 2 * B_BYTES;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(dA, dB, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("umma_probe variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
            const double err = fabs(ref - (double)out[m * N + n]);
            if (!(err <= worst)) worst = err;
            scale = fmax(scale, fabs(ref));
            if (!(err < 1e-2 * 64000)) ++bad;
        }
    printf("umma_probe variant %d: max |err| = %.6g on values up to %.6g (relative %.3g), %d of %d gross mismatches -> %s\n", variant, worst, scale,
           worst / scale, bad, M * N, worst / scale < (variant == 2 ? 3e-3 : 3e-6) ? "OK" : "MISMATCH");
    printf("  out[0][0..3] = %g %g %g %g;  out[65][1] = %g\n", out[0], out[1], out[2], out[3], out[65 * N + 1]);
    return worst / scale < (variant == 2 ? 3e-3 : 3e-6) ? 0 : 1;
}
