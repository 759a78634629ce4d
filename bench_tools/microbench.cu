// microbench.cu -- measures the B200 limits the guided-filter kernels are designed against:
// HBM copy GB/s, L2-resident read GB/s, shared-memory LDS.32/LDS.128 bytes/clk/SM, warp-shuffle
// rate, FADD/FFMA/FFMA2 rate, and whether shuffles and LDS overlap.  Prints one JSON object.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_read(const float4* __restrict__ a, size_t n, int reps, float* out)
{
    float s = 0;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = __ldcg(a + i);
            s += v.x + v.y + v.z + v.w;
        }
    if (s == 12345.678f) *out = s;
}

template <int MODE>   // 0 shfl, 1 lds32, 2 lds128, 3 shfl+lds32, 4 fadd, 5 ffma, 6 ffma2, 7 shfl+fadd
__global__ void __launch_bounds__(1024) k_pipe(int iters, long long* cycles, float* out)
{
    __shared__ float4 sm4[2048];
    float* sm = (float*)sm4;
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = (float)i;
    float a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3, a4 = tid + 4, a5 = tid + 5, a6 = tid + 6, a7 = tid + 7;
    float b0 = 1, b1 = 2, b2 = 3, b3 = 4, b4 = 5, b5 = 6, b6 = 7, b7 = 8;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3 || MODE == 7) {
            a0 = __shfl_xor_sync(~0u, a0, 1); a1 = __shfl_xor_sync(~0u, a1, 2); a2 = __shfl_xor_sync(~0u, a2, 4); a3 = __shfl_xor_sync(~0u, a3, 8);
            a4 = __shfl_xor_sync(~0u, a4, 16); a5 = __shfl_xor_sync(~0u, a5, 3); a6 = __shfl_xor_sync(~0u, a6, 5); a7 = __shfl_xor_sync(~0u, a7, 7);
        }
        if (MODE == 1 || MODE == 3) {
            const int o = (it * 32) & 4095;
            b0 += sm[(tid + o) & 8191]; b1 += sm[(tid + o + 1024) & 8191]; b2 += sm[(tid + o + 2048) & 8191]; b3 += sm[(tid + o + 3072) & 8191];
            b4 += sm[(tid + o + 4096) & 8191]; b5 += sm[(tid + o + 5120) & 8191]; b6 += sm[(tid + o + 6144) & 8191]; b7 += sm[(tid + o + 7168) & 8191];
        }
        if (MODE == 2) {
            const int o = it & 1023;
            float4 v0 = sm4[(tid + o) & 2047], v1 = sm4[(tid + o + 512) & 2047], v2 = sm4[(tid + o + 1024) & 2047], v3 = sm4[(tid + o + 1536) & 2047];
            b0 += v0.x; b1 += v0.w; b2 += v1.y; b3 += v1.z; b4 += v2.x; b5 += v2.w; b6 += v3.y; b7 += v3.z;
        }
        if (MODE == 4 || MODE == 7) {
            b0 += a0; b1 += a1; b2 += a2; b3 += a3; b4 += a4; b5 += a5; b6 += a6; b7 += a7;
        }
        if (MODE == 5) {
            b0 = fmaf(b0, a0, a1); b1 = fmaf(b1, a1, a2); b2 = fmaf(b2, a2, a3); b3 = fmaf(b3, a3, a4);
            b4 = fmaf(b4, a4, a5); b5 = fmaf(b5, a5, a6); b6 = fmaf(b6, a6, a7); b7 = fmaf(b7, a7, a0);
        }
        if (MODE == 6) {
            asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%4, %5}; mov.b64 z, {%6, %7};\n"
                         "fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;\n"
                         "mov.b64 x, {%2, %3}; fma.rn.f32x2 x, x, y, z; mov.b64 {%2, %3}, x; }\n"
                         : "+f"(b0), "+f"(b1), "+f"(b2), "+f"(b3) : "f"(a0), "f"(a1), "f"(a2), "f"(a3));
            asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%4, %5}; mov.b64 z, {%6, %7};\n"
                         "fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;\n"
                         "mov.b64 x, {%2, %3}; fma.rn.f32x2 x, x, y, z; mov.b64 {%2, %3}, x; }\n"
                         : "+f"(b4), "+f"(b5), "+f"(b6), "+f"(b7) : "f"(a4), "f"(a5), "f"(a6), "f"(a7));
        }
    }
    long long t1 = clock64();
    float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + b0 + b1 + b2 + b3 + b4 + b5 + b6 + b7;
    if (s == 12345.678f) *out = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
double run_pipe(int threads, int blocks_per_sm, int sms, int iters, long long* d_cyc, float* d_out)
{
    const int grid = sms * blocks_per_sm;
    k_pipe<MODE><<<grid, threads>>>(iters, d_cyc, d_out);
    k_pipe<MODE><<<grid, threads>>>(iters, d_cyc, d_out);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(grid);
    CK(cudaMemcpy(c.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (auto v : c) avg += (double)v;
    avg /= grid;
    // warp-instructions of the measured kind per clock per SM (8 per iteration per warp)
    const double warps_per_sm = (double)threads / 32 * blocks_per_sm;
    return 8.0 * iters * warps_per_sm / avg;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    // HBM copy: 1 GiB -> 1 GiB
    const size_t nbytes = (size_t)1 << 30;
    float4 *a, *b;
    CK(cudaMalloc(&a, nbytes)); CK(cudaMalloc(&b, nbytes));
    CK(cudaMemset(a, 1, nbytes)); CK(cudaMemset(b, 0, nbytes));
    double best_copy = 0;
    for (int i = 0; i < 6; ++i) {
        CK(cudaEventRecord(e0));
        k_copy<<<sms * 16, 512>>>(a, b, nbytes / 16);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double g = 2.0 * nbytes / ms / 1e6;
        if (g > best_copy) best_copy = g;
    }
    double best_memcpy = 0;
    for (int i = 0; i < 4; ++i) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(b, a, nbytes, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double g = 2.0 * nbytes / ms / 1e6;
        if (g > best_memcpy) best_memcpy = g;
    }
    // HBM read-only 1 GiB
    float* d_out; CK(cudaMalloc(&d_out, 4));
    double best_read = 0;
    for (int i = 0; i < 4; ++i) {
        CK(cudaEventRecord(e0));
        k_read<<<sms * 16, 512>>>(a, nbytes / 16, 1, d_out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double g = 1.0 * nbytes / ms / 1e6;
        if (g > best_read) best_read = g;
    }
    // L2-resident read: 32 MiB x 64 reps
    double l2[3];
    size_t l2sizes[3] = {(size_t)16 << 20, (size_t)32 << 20, (size_t)64 << 20};
    for (int k = 0; k < 3; ++k) {
        const size_t l2bytes = l2sizes[k];
        double best = 0;
        for (int i = 0; i < 3; ++i) {
            CK(cudaEventRecord(e0));
            k_read<<<sms * 8, 1024>>>(a, l2bytes / 16, 64, d_out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double g = 64.0 * l2bytes / ms / 1e6;
            if (g > best) best = g;
        }
        l2[k] = best;
    }
    long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 4));
    const int it = 4096;
    double shfl = run_pipe<0>(1024, 2, sms, it, d_cyc, d_out);
    double lds32 = run_pipe<1>(1024, 2, sms, it, d_cyc, d_out);
    double lds128 = run_pipe<2>(1024, 2, sms, it, d_cyc, d_out) / 2.0;   // 4 LDS.128 per iteration
    double both = run_pipe<3>(1024, 2, sms, it, d_cyc, d_out);           // 8 shfl + 8 lds per iteration
    double fadd = run_pipe<4>(1024, 2, sms, it, d_cyc, d_out);
    double ffma = run_pipe<5>(1024, 2, sms, it, d_cyc, d_out);
    double ffma2 = run_pipe<6>(1024, 2, sms, it, d_cyc, d_out) / 2.0;    // 4 FFMA2 per iteration
    double shfl_fadd = run_pipe<7>(1024, 2, sms, it, d_cyc, d_out);
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"hbm_copy_gbs\": %.1f, \"hbm_memcpy_gbs\": %.1f, \"hbm_read_gbs\": %.1f, "
           "\"l2_read_gbs_16M\": %.1f, \"l2_read_gbs_32M\": %.1f, \"l2_read_gbs_64M\": %.1f, "
           "\"shfl_warpinst_per_clk_sm\": %.3f, \"lds32_warpinst_per_clk_sm\": %.3f, \"lds128_warpinst_per_clk_sm\": %.3f, "
           "\"shfl_plus_lds32_pairs_per_clk_sm\": %.3f, \"fadd_warpinst_per_clk_sm\": %.3f, \"ffma_warpinst_per_clk_sm\": %.3f, "
           "\"ffma2_warpinst_per_clk_sm\": %.3f, \"shfl_plus_fadd_pairs_per_clk_sm\": %.3f}\n",
           prop.name, sms, clk, best_copy, best_memcpy, best_read, l2[0], l2[1], l2[2], shfl, lds32, lds128, both, fadd, ffma, ffma2, shfl_fadd);
    return 0;
}
