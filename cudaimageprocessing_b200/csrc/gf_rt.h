// gf_rt.h -- the few runtime calls the API layer needs, so that the same host logic builds
// against the CUDA runtime (product) or the test-only SIMT emulator (tests/emu, -DGF_CPU_EMU).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef GF_CPU_EMU
#include <cstdlib>
#include <cstring>

#include "cuda_emu.h"
#define GF_LAUNCH(kernel, grid, block, smem, stream, ...) GF_EMU_LAUNCH(kernel, grid, block, smem, __VA_ARGS__)
static inline const char* gf_rt_launch_error() { return nullptr; }
template <class K> static inline const char* gf_rt_set_smem(K, size_t) { return nullptr; }
template <class K> static inline int gf_rt_ctas_per_sm(K, int, size_t) { return 4; }
static inline const char* gf_rt_alloc_async(void** p, size_t n, void*)
{   // 256-byte aligned like cudaMallocAsync (the tuned kernels check the alignment of every plane)
    *p = std::aligned_alloc(256, (n + 255) / 256 * 256);
    return *p ? nullptr : "malloc failed";
}
static inline void gf_rt_free_async(void* p, void*) { std::free(p); }
static inline const char* gf_rt_copy2d_async(void* d, size_t dp, const void* s, size_t sp, size_t wb, size_t rows, void*)
{
    for (size_t y = 0; y < rows; ++y) std::memcpy((char*)d + y * dp, (const char*)s + y * sp, wb);
    return nullptr;
}
static inline const char* gf_rt_device_info(int* sms, int* major, int* minor) { *sms = 4; *major = 0; *minor = 0; return nullptr; }
static inline size_t gf_rt_max_smem() { return 200 * 1024; }
#else
#include <cuda_runtime.h>

#include <mutex>
#define GF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__)
static inline const char* gf_rt_launch_error()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
template <class K> static inline const char* gf_rt_set_smem(K kernel, size_t bytes)
{
    if (bytes <= 48 * 1024) return nullptr;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
// resident CTAs per SM of a kernel (registers, shared memory, threads): the wave size of the band choosers
template <class K> static inline int gf_rt_ctas_per_sm(K kernel, int threads, size_t smem)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}
// Stream-ordered temporaries come from a pool of the library's own (one per device) that keeps up to 256 MiB cached:
// the default pool returns everything to the driver at every synchronisation (release threshold 0), which turns the
// first temporary after each cudaDeviceSynchronize into a driver allocation.
inline cudaMemPool_t gf_rt_pool()
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    static bool failed[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev] && !failed[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) {
            pools[dev] = nullptr; failed[dev] = true;
            cudaGetLastError();
        } else {
            uint64_t keep = 256ull << 20;
            cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return pools[dev];
}
static inline const char* gf_rt_alloc_async(void** p, size_t n, void* stream)
{
    cudaMemPool_t pool = gf_rt_pool();
    cudaError_t e = pool ? cudaMallocFromPoolAsync(p, n, pool, (cudaStream_t)stream) : cudaMallocAsync(p, n, (cudaStream_t)stream);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
static inline void gf_rt_free_async(void* p, void* stream) { cudaFreeAsync(p, (cudaStream_t)stream); }
static inline const char* gf_rt_copy2d_async(void* d, size_t dp, const void* s, size_t sp, size_t wb, size_t rows, void* stream)
{
    cudaError_t e = cudaMemcpy2DAsync(d, dp, s, sp, wb, rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
static inline const char* gf_rt_device_info(int* sms, int* major, int* minor)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
static inline size_t gf_rt_max_smem() { return 227 * 1024; }
#endif
