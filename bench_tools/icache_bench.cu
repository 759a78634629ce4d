// icache_bench.cu -- does a straight-line loop body larger than the L0 instruction cache throttle
// instruction issue on B200?  Loop bodies of N independent FFMAs, w warps per SM (1-warp CTAs).
// skew > 0: every warp first waits (hash of its block index) x skew clocks, so that the warps of one scheduler run
// the SAME loop at DIFFERENT program counters (the situation of the filter kernels, whose warps drift apart).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o icache_bench icache_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int N>
__global__ void __launch_bounds__(32) k_body(int iters, float* out, float s, int skew)
{
    if (skew > 0) {
        const long long t0 = clock64(), wait = (long long)((blockIdx.x * 2654435761u >> 24) & 63) * skew;
        while (clock64() - t0 < wait) { }
    }
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < N / 16; ++j) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 1.0f + j);
        }
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += a[i];
    if (t == 123.456f) *out = t;
}

template <int N>
void run(int sms, float* d_out, int skew)
{
    for (int w : {1, 4, 8, 16}) {
        const int iters = 4000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_body<N><<<sms * w, 32>>>(iters, d_out, 0.999f, skew);
        cudaEventRecord(e0);
        k_body<N><<<sms * w, 32>>>(iters, d_out, 0.999f, skew);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        int clk = 0;
        cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        double cycles = ms * 1e-3 * clk * 1e3;
        printf("{\"body_instr\": %d, \"warps_per_sm\": %d, \"skew_clk\": %d, \"ipc_sm\": %.3f}\n", N, w, skew, (double)N * iters * w / cycles);
    }
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* d_out;
    cudaMalloc(&d_out, 4);
    for (int skew : {0, 37, 301}) {
        run<256>(p.multiProcessorCount, d_out, skew);
        run<320>(p.multiProcessorCount, d_out, skew);
        run<384>(p.multiProcessorCount, d_out, skew);
        run<512>(p.multiProcessorCount, d_out, skew);
        run<768>(p.multiProcessorCount, d_out, skew);
        run<1024>(p.multiProcessorCount, d_out, skew);
        run<2048>(p.multiProcessorCount, d_out, skew);
    }
    return 0;
}
