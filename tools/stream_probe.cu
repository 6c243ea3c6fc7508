// Micro-probe: tile-space streaming efficiency vs thread mapping (3 plane reads + 6 plane writes per tile pixel,
// the k_dirinterp traffic shape) on 2795 slabs of 1,448,704 B.
#include <cstdio>
#include <cuda_runtime.h>

constexpr size_t SLAB = 1448704, PL = 102528;

template <int ROWS>
__global__ void k_scalar(char* base)     // block (160, ROWS): one float per thread
{
    const int t = blockIdx.y;
    const int i = (blockIdx.x * ROWS + threadIdx.y) * 160 + threadIdx.x;
    char* s = base + (size_t)t * SLAB;
    float acc = 0.f;
    for (int p = 0; p < 3; ++p) acc += reinterpret_cast<const float*>(s + p * PL)[i];
    for (int p = 0; p < 6; ++p) reinterpret_cast<float*>(s + (3 + p) * PL)[i] = acc + p;
}

template <int ROWS>
__global__ void k_vec4(char* base)       // block (40, ROWS): one float4 per thread
{
    const int t = blockIdx.y;
    const int i = (blockIdx.x * ROWS + threadIdx.y) * 40 + threadIdx.x;
    char* s = base + (size_t)t * SLAB;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int p = 0; p < 3; ++p) { const float4 v = reinterpret_cast<const float4*>(s + p * PL)[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    for (int p = 0; p < 6; ++p) reinterpret_cast<float4*>(s + (3 + p) * PL)[i] = make_float4(acc.x + p, acc.y, acc.z, acc.w);
}

// one CTA per tile, loops over the whole 160x160 plane (persistent within the tile)
template <int NT>
__global__ void k_tile_vec4(char* base)
{
    char* s = base + (size_t)blockIdx.x * SLAB;
    for (int i = threadIdx.x; i < 6400; i += NT) {
        float4 acc = make_float4(0, 0, 0, 0);
        for (int p = 0; p < 3; ++p) { const float4 v = reinterpret_cast<const float4*>(s + p * PL)[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        for (int p = 0; p < 6; ++p) reinterpret_cast<float4*>(s + (3 + p) * PL)[i] = make_float4(acc.x + p, acc.y, acc.z, acc.w);
    }
}

// plain copy for reference: 1 read + 1 write of the whole buffer
__global__ void k_copy(const float4* a, float4* b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main()
{
    const int nt = 2795;
    char* d;
    cudaMalloc(&d, nt * SLAB * 2);
    cudaMemset(d, 0, nt * SLAB * 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = 9.0 * 102400 * nt / 1e9;
    auto run = [&](const char* name, auto&& f, double g) { f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int r = 0; r < 5; ++r) f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5; printf("%-28s %.3f ms  %.0f GB/s\n", name, ms, g / (ms * 1e-3)); };
    run("scalar 160x4", [&] { k_scalar<4><<<dim3(40, nt), dim3(160, 4)>>>(d); }, gb);
    run("scalar 160x2", [&] { k_scalar<2><<<dim3(80, nt), dim3(160, 2)>>>(d); }, gb);
    run("vec4 40x4", [&] { k_vec4<4><<<dim3(40, nt), dim3(40, 4)>>>(d); }, gb);
    run("vec4 40x8", [&] { k_vec4<8><<<dim3(20, nt), dim3(40, 8)>>>(d); }, gb);
    run("vec4 40x16", [&] { k_vec4<16><<<dim3(10, nt), dim3(40, 16)>>>(d); }, gb);
    run("tile vec4 256thr", [&] { k_tile_vec4<256><<<nt, 256>>>(d); }, gb);
    run("tile vec4 512thr", [&] { k_tile_vec4<512><<<nt, 512>>>(d); }, gb);
    run("tile vec4 1024thr", [&] { k_tile_vec4<1024><<<nt, 1024>>>(d); }, gb);
    const size_t n4 = nt * SLAB / 16;
    run("copy 4GB (1R+1W)", [&] { k_copy<<<148 * 16, 512>>>((const float4*)d, (float4*)(d + nt * SLAB), n4); }, 2.0 * nt * SLAB / 1e9);
    return 0;
}
