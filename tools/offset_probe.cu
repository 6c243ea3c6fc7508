// Micro-probe: how does the byte distance between concurrently streamed planes affect HBM throughput?
// (motivated by the AMaZE split pass: load A[k]; store A[k+51200B]; store A[k] ran 20x slower than either store alone)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

// pattern 1: the split: per tile slab, rows of 80 floats, every other row
__global__ void k_split_like(char* base, size_t slab, size_t offA, size_t dist, int mode)
{
    const int t = blockIdx.y;
    const int rr = 13 + 2 * (blockIdx.x * 4 + threadIdx.y);
    if (rr >= 148) return;
    float* A = reinterpret_cast<float*>(base + (size_t)t * slab + offA);
    float* B = reinterpret_cast<float*>(base + (size_t)t * slab + offA + dist);
    const int k = rr * 80 + threadIdx.x;
    if (threadIdx.x < 6 || threadIdx.x >= 74) return;
    const float v = A[k];
    if (mode & 1) B[k] = v;
    if (mode & 2) A[k] = 0.f;
}

// pattern 2: N planes spaced `stride` bytes apart inside each slab: read 3, write 6, thread per pixel
__global__ void k_multi(char* base, size_t slab, size_t stride, int nread, int nwrite)
{
    const int t = blockIdx.y;
    const int i = (blockIdx.x * 4 + threadIdx.y) * 160 + threadIdx.x;
    char* s = base + (size_t)t * slab;
    float acc = 0.f;
    for (int p = 0; p < nread; ++p) acc += reinterpret_cast<const float*>(s + p * stride)[i];
    for (int p = 0; p < nwrite; ++p) reinterpret_cast<float*>(s + (nread + p) * stride)[i] = acc + p;
}

int main()
{
    const int ntiles = 2795;
    const size_t slab = 1448704;
    char* d;
    cudaMalloc(&d, ntiles * slab + (64 << 20));
    cudaMemset(d, 0, ntiles * slab + (64 << 20));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto&& f) { f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int r = 0; r < 5; ++r) f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5; };
    printf("split-like pattern (load A, store A+dist, store A): ms\n");
    for (size_t dist : {51200ul, 51200ul + 128, 51200ul + 256, 51200ul + 512, 51200ul + 1024, 51200ul + 2048, 51200ul + 4096, 102528ul, 65536ul, 40960ul, 32768ul + 128})
        for (int mode : {1, 2, 3}) {
            float ms = time([&] { k_split_like<<<dim3(20, ntiles), dim3(80, 4)>>>(d, slab, 6 * 102528, dist, mode); });
            printf("  dist %7zu mode %d : %.3f ms\n", dist, mode, ms);
        }
    printf("multi-plane pattern: 3 reads + 6 writes of a 160x160 fp32 plane per tile, planes `stride` bytes apart\n");
    for (size_t stride : {102400ul, 102528ul, 102400ul + 256, 102400ul + 512, 102400ul + 1024, 102400ul + 2048, 102400ul + 4096, 102400ul + 8192, 131072ul, 131072ul + 128, 106496ul})
    {
        float ms = time([&] { k_multi<<<dim3(40, ntiles), dim3(160, 4)>>>(d, slab, stride, 3, 6); });
        const double gb = 9.0 * 102400 * ntiles / 1e9;
        printf("  stride %7zu : %.3f ms  %.0f GB/s\n", stride, ms, gb / (ms * 1e-3));
    }
    // slab stride effect
    printf("slab stride effect (stride 102528 planes)\n");
    for (size_t sl : {1448704ul, 1448704ul + 256, 1448704ul + 1024, 1448704ul + 4096, 1450000ul / 128 * 128, 1507328ul, 1572864ul}) {
        float ms = time([&] { k_multi<<<dim3(40, ntiles), dim3(160, 4)>>>(d, sl, 102528, 3, 6); });
        printf("  slab %8zu : %.3f ms\n", sl, ms);
    }
    return 0;
}
