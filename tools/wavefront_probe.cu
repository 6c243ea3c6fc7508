// Micro-probe: cost per step of the fp32 integral-image wavefront (nlmeans.cu phase B) in a few formulations.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o wavefront_probe wavefront_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int TS = 150, NT = 160;
__device__ __forceinline__ float addz(float x, float y) { float r; asm("add.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float subz(float x, float y) { float r; asm("sub.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y)); return r; }

// V: 0 flags+membar, 1 flags without membar, 2 block barrier per step, 3 one warp does 32-row blocks (other warps idle)
template <int V>
__global__ void __launch_bounds__(NT, 2) probe(float* out, long long* clk, int reps)
{
    extern __shared__ float St[];
    __shared__ volatile int prog[NT / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int TH = TS, TW = TS;
    if (t < NT / 32) prog[t] = 0;
    for (int i = t; i < TS * TS; i += NT) St[i] = (float)((i * 7 + blockIdx.x) % 13) * 0.01f;
    __syncthreads();
    long long c0 = clock64();
    for (int shift = 0; shift < reps; ++shift) {
        if (V == 3) {
            if (warp == 0) {
                for (int b = 0; b * 32 < TH; ++b) {
                    const int r = b * 32 + lane;
                    const bool rowvalid = r < TH;
                    float* myrow = St + min(r, TS - 1) * TS;
                    const float* seam = St + max(b * 32 - 1, 0) * TS;
                    float left = 0.f, up = 0.f, upl, val = 0.f, sc = myrow[0];
                    float seam_cur = b ? seam[0] : 0.f;
                    const int nsteps = TW + 31;
                    for (int s = 0; s < nsteps; ++s) {
                        const int xx = s - lane;
                        const bool act = rowvalid && (unsigned)xx < (unsigned)TW;
                        upl = up;
                        up = __shfl_up_sync(0xffffffffu, val, 1);
                        if (lane == 0) up = seam_cur;
                        seam_cur = b ? seam[min(s + 1, TW - 1)] : 0.f;
                        const float v = subz(addz(left, up), subz(upl, sc));
                        if (act) { myrow[xx] = v; val = v; left = v; }
                        if (rowvalid && (unsigned)(xx + 1) < (unsigned)TW) sc = myrow[xx + 1];
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            continue;
        }
        const bool rowvalid = t < TH;
        float* myrow = St + min(t, TS - 1) * TS;
        const volatile float* seam = St + max(t - 1, 0) * TS;
        float left = 0.f, up = 0.f, upl, val = 0.f, sc = myrow[0];
        const int nsteps = TW + TH - 1;
        const int base = shift * 256;
        int seen = 0;
        for (int s = 0; s < nsteps; ++s) {
            const int xx = s - t;
            const bool act = rowvalid && (unsigned)xx < (unsigned)TW;
            upl = up;
            if (V == 2) {
                up = (t > 0 && act) ? seam[xx] : 0.f;
            } else {
                up = __shfl_up_sync(0xffffffffu, val, 1);
                if (lane == 0) {
                    up = 0.f;
                    if (warp > 0 && act) {
                        if (seen < base + xx + 1) {
                            do { seen = prog[warp - 1]; } while (seen < base + xx + 1);
                            if (V == 0) __threadfence_block();
                        }
                        up = seam[xx];
                    }
                }
            }
            const float v = subz(addz(left, up), subz(upl, sc));
            if (act) {
                myrow[xx] = v; val = v; left = v;
                if (V != 2 && lane == 31 && ((xx & 3) == 3 || xx == TW - 1)) {
                    if (V == 0) __threadfence_block();
                    prog[warp] = base + xx + 1;
                }
            }
            if (rowvalid && (unsigned)(xx + 1) < (unsigned)TW) sc = myrow[xx + 1];
            if (V == 2) __syncthreads();
        }
        __syncthreads();
    }
    long long c1 = clock64();
    if (t == 0) clk[blockIdx.x] = c1 - c0;
    if (t < TS) out[blockIdx.x * TS + t] = St[t * TS + TS - 1];
}

// V4: every warp owns its own 36-row ring and its own tile (10 warps/SM when launched with 2 CTAs of 5 warps): 32-row blocks
__global__ void __launch_bounds__(NT, 2) probe_ring(float* out, long long* clk, int reps)
{
    extern __shared__ float Sall[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float* R = Sall + warp * 36 * TS;
    for (int i = lane; i < 36 * TS; i += 32) R[i] = (float)((i * 7 + blockIdx.x) % 13) * 0.01f;
    __syncwarp();
    const int TH = TS, TW = TS;
    long long c0 = clock64();
    for (int shift = 0; shift < reps; ++shift) {
        for (int b = 0; b * 32 < TH; ++b) {
            const int r = b * 32 + lane;
            const bool rowvalid = r < TH;
            float* myrow = R + (r % 36) * TS;
            const float* seam = R + ((b * 32 + 35) % 36) * TS;
            float left = 0.f, up = 0.f, upl, val = 0.f, sc = myrow[0];
            float seam_cur = b ? seam[0] : 0.f;
            const int nsteps = TW + 31;
            for (int s = 0; s < nsteps; ++s) {
                const int xx = s - lane;
                const bool act = rowvalid && (unsigned)xx < (unsigned)TW;
                upl = up;
                up = __shfl_up_sync(0xffffffffu, val, 1);
                if (lane == 0) up = seam_cur;
                seam_cur = b ? seam[min(s + 1, TW - 1)] : 0.f;
                const float v = subz(addz(left, up), subz(upl, sc));
                if (act) { myrow[xx] = v; val = v; left = v; }
                if (rowvalid && (unsigned)(xx + 1) < (unsigned)TW) sc = myrow[xx + 1];
            }
            __syncwarp();
        }
    }
    long long c1 = clock64();
    if (t == 0) clk[blockIdx.x] = c1 - c0;
    if (lane == 0) out[blockIdx.x * 8 + warp] = R[TS - 1];
}

template <int V> void run(const char* name, int reps)
{
    float* out; long long* clk;
    cudaMalloc(&out, 296 * TS * 4); cudaMalloc(&clk, 296 * 8);
    const size_t smem = TS * TS * 4;
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<V><<<296, NT, smem>>>(out, clk, reps);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<V><<<296, NT, smem>>>(out, clk, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %8.3f ms  %9lld clk/CTA  %7.1f clk per integral image  err=%s\n", name, ms, h, (double)h / reps, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main()
{
    const int reps = 121;
    run<0>("5 warps, flags + membar", reps);
    run<1>("5 warps, flags, no membar", reps);
    run<2>("5 warps, block barrier per step", reps);
    run<3>("1 warp, 32-row blocks", reps);
    {
        float* out; long long* clk;
        cudaMalloc(&out, 296 * 8 * 4); cudaMalloc(&clk, 296 * 8);
        const size_t smem = 5 * 36 * TS * 4;
        cudaFuncSetAttribute(probe_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_ring<<<296, NT, smem>>>(out, clk, reps);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe_ring<<<296, NT, smem>>>(out, clk, reps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %8.3f ms  %9lld clk/CTA  %7.1f clk per integral image per warp (10 warps/SM, each its own tile)  err=%s\n",
               "per-warp ring, 32-row blocks", ms, h, (double)h / reps, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
