// ws_micro.cu -- primitive costs behind the design of gf_ws (developer tool; build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o bench_tools/ws_micro bench_tools/ws_micro.cu; prints one JSON object).
//   nanosleep_clk         : SM clocks one __nanosleep(32) takes
//   ld_*_B_per_clk_sm     : global-load throughput per SM for the access shapes a K-columns-per-lane strip can use,
//                           on an L2-resident buffer, 8 warps per SM:
//                             k12_s48   3 x LDG.128, lane stride 48 B (lane owns 12 adjacent columns)
//                             k12_coal  3 x LDG.128, lane stride 16 B (coalesced; columns interleaved over lanes)
//                             k8_256    1 x LDG.256, lane stride 32 B
//                             k8_128    2 x LDG.128, lane stride 32 B
//                             k16_256   2 x LDG.256, lane stride 64 B
//   bulk_B_per_clk_sm     : cp.async.bulk global->shared of 1536-byte rows + 3 x LDS.128 (lane stride 48 B) per lane
//   mbar_roundtrip_clk    : producer warp arrives on an mbarrier, consumer warp wakes and arrives back (two hops)
//   flag_roundtrip_clk    : the same with st.release / ld.acquire spinning on a shared-memory word
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void k_nanosleep(long long* out)
{
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) __nanosleep(32);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / 256;
}

__device__ __forceinline__ void ld256(const float* p, float (&v)[8])
{
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

// MODE 0 k12_s48, 1 k12_coal, 2 k8_256, 3 k8_128, 4 k16_256.  Every warp walks `rows` rows of its own column window.
template <int MODE>
__global__ void __launch_bounds__(256) k_load(const float* buf, int stride, int rows, float* sink, long long* clk)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    constexpr int WCOLS = MODE <= 1 ? 384 : (MODE == 4 ? 512 : 256);
    const int nwin = stride / WCOLS;
    const float* base = buf + (size_t)(warp % nwin) * WCOLS + (size_t)((warp / nwin) * 7 % 64) * stride;
    float acc = 0.f;
    long long t0 = clock64();
#pragma unroll 2
    for (int y = 0; y < rows; ++y) {
        const float* rp = base + (size_t)y * stride;
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { float4 t = *reinterpret_cast<const float4*>(rp + lane * 12 + 4 * c); acc += t.x + t.y + t.z + t.w; }
        } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { float4 t = *reinterpret_cast<const float4*>(rp + lane * 4 + 128 * c); acc += t.x + t.y + t.z + t.w; }
        } else if (MODE == 2) {
            float v[8]; ld256(rp + lane * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += v[i];
        } else if (MODE == 3) {
#pragma unroll
            for (int c = 0; c < 2; ++c) { float4 t = *reinterpret_cast<const float4*>(rp + lane * 8 + 4 * c); acc += t.x + t.y + t.z + t.w; }
        } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[8]; ld256(rp + lane * 16 + 8 * c, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += v[i];
            }
        }
    }
    long long t1 = clock64();
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// one warp per stream: lane 0 issues bulk copies of 1536-byte rows into a 4-deep shared-memory ring, all lanes read them
__global__ void __launch_bounds__(256) k_bulk(const float* buf, int stride, int rows, float* sink, long long* clk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int D = 4, ROWB = 1536;
    float* st = reinterpret_cast<float*>(smem) + (size_t)wl * D * (ROWB / 4);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 8 * D * ROWB) + wl * D;
    const int nwin = stride / 384;
    const float* base = buf + (size_t)(warp % nwin) * 384 + (size_t)((warp / nwin) * 7 % 64) * stride;
    if (lane == 0)
        for (int i = 0; i < D; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](int y) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + y % D);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(st + (size_t)(y % D) * (ROWB / 4));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWB) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(base + (size_t)y * stride), "r"(ROWB), "r"(bar) : "memory");
    };
    float acc = 0.f;
    long long t0 = clock64();
    if (lane == 0) for (int y = 0; y < D - 1 && y < rows; ++y) issue(y);
    for (int y = 0; y < rows; ++y) {
        if (lane == 0 && y + D - 1 < rows) issue(y + D - 1);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + y % D);
        const unsigned par = (y / D) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(bar), "r"(par) : "memory");
        const float* rp = st + (size_t)(y % D) * (ROWB / 4) + lane * 12;
#pragma unroll
        for (int c = 0; c < 3; ++c) { float4 t = *reinterpret_cast<const float4*>(rp + 4 * c); acc += t.x + t.y + t.z + t.w; }
        __syncwarp();           // the slot is free again once every lane has read it (it is refilled D-1 rows later)
    }
    long long t1 = clock64();
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void k_pingpong(long long* out, int use_mbar)
{
    __shared__ unsigned long long bars[2];
    __shared__ volatile int flags[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bars + i)));
        flags[0] = flags[1] = 0;
    }
    __syncthreads();
    const int N = 2000;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
        const int mine = warp, other = warp ^ 1;
        if (warp == 0) {
            __syncwarp();
            if (use_mbar) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bars + mine)) : "memory"); }
            else if (lane == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared((int*)flags + mine)), "r"(i + 1) : "memory");
        }
        if (use_mbar) {
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}"
                         ::"r"((unsigned)__cvta_generic_to_shared(bars + other)), "r"((unsigned)(i & 1)) : "memory");
        } else {
            int v;
            do { asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared((int*)flags + other)) : "memory"); } while (v < i + 1);
        }
        if (warp == 1) {
            __syncwarp();
            if (use_mbar) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bars + mine)) : "memory"); }
            else if (lane == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared((int*)flags + mine)), "r"(i + 1) : "memory");
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / N;
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long *d_clk, h_clk[1024];
    float *buf, *sink;
    const int stride = 384 * 256 * 4 / 4 * 1;       // 98304 floats per row: divisible by 384, 256 and 512
    const int rows_total = 96;                       // 37.7 MB: L2-resident
    CK(cudaMalloc(&buf, (size_t)stride * rows_total * 4));
    CK(cudaMemset(buf, 0, (size_t)stride * rows_total * 4));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMalloc(&d_clk, 1024 * 8));
    printf("{\"sms\": %d", sms);
    k_nanosleep<<<1, 32>>>(d_clk);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_clk, d_clk, 8, cudaMemcpyDeviceToHost));
    printf(", \"nanosleep32_clk\": %lld", h_clk[0]);
    const int rows = 24;
    auto report = [&](const char* name, double bytes_per_cta) {
        cudaMemcpy(h_clk, d_clk, sms * 8, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int i = 0; i < sms; ++i) mx = h_clk[i] > mx ? h_clk[i] : mx;
        printf(", \"%s_B_per_clk_sm\": %.1f", name, bytes_per_cta / mx);
    };
    for (int rep = 0; rep < 2; ++rep) {     // second pass: warm L2
        k_load<0><<<sms, 256>>>(buf, stride, rows, sink, d_clk); CK(cudaDeviceSynchronize());
        if (rep) report("ld_k12_s48", 8.0 * rows * 1536);
        k_load<1><<<sms, 256>>>(buf, stride, rows, sink, d_clk); CK(cudaDeviceSynchronize());
        if (rep) report("ld_k12_coal", 8.0 * rows * 1536);
        k_load<2><<<sms, 256>>>(buf, stride, rows, sink, d_clk); CK(cudaDeviceSynchronize());
        if (rep) report("ld_k8_256", 8.0 * rows * 1024);
        k_load<3><<<sms, 256>>>(buf, stride, rows, sink, d_clk); CK(cudaDeviceSynchronize());
        if (rep) report("ld_k8_128", 8.0 * rows * 1024);
        k_load<4><<<sms, 256>>>(buf, stride, rows, sink, d_clk); CK(cudaDeviceSynchronize());
        if (rep) report("ld_k16_256", 8.0 * rows * 2048);
        const size_t sm = 8 * 4 * 1536 + 8 * 4 * 8;
        CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_bulk<<<sms, 256, sm>>>(buf, stride, rows, sink, d_clk); CK(cudaDeviceSynchronize());
        if (rep) report("bulk_k12", 8.0 * rows * 1536);
    }
    k_pingpong<<<1, 64>>>(d_clk, 1); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_clk, d_clk, 8, cudaMemcpyDeviceToHost));
    printf(", \"mbar_roundtrip_clk\": %lld", h_clk[0]);
    k_pingpong<<<1, 64>>>(d_clk, 0); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_clk, d_clk, 8, cudaMemcpyDeviceToHost));
    printf(", \"flag_roundtrip_clk\": %lld}\n", h_clk[0]);
    return 0;
}
