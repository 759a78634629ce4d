// cuda_utils.h -- the small helper set that the reference's headers pull in
// (GuidedFilter/cuda_utils.h: CHECK, CheckMsg, CUDA_SAFE_FREE, initDevice, cpuTimer, GpuTimer,
// iAlignUp, iDivUp), re-implemented for the drop-in shims.  Same names and behaviour so that
// host code written against the reference compiles unchanged.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#define CHECK(err) gf_shim::check((err), __FILE__, __LINE__)
#define CheckMsg(msg) gf_shim::check_msg((msg), __FILE__, __LINE__)
#define CUDA_SAFE_FREE(a) \
    if ((a) != nullptr) CHECK(cudaFree(a))

namespace gf_shim {
inline void check(cudaError_t err, const char* file, int line)
{
    if (err == cudaSuccess) return;
    std::fprintf(stderr, "CHECK() Runtime API error in file <%s>, line %i : %s.\n", file, line, cudaGetErrorString(err));
    std::exit(-1);
}
inline void check_msg(const char* msg, const char* file, int line)
{
    const cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) return;
    std::fprintf(stderr, "CheckMsg() CUDA error: %s in file <%s>, line %i : %s.\n", msg, file, line, cudaGetErrorString(err));
    std::exit(-1);
}
}  // namespace gf_shim

// Selects device `dev` (clamped to the devices present) and reports it on stderr.
inline bool initDevice(int dev)
{
    int n = 0;
    CHECK(cudaGetDeviceCount(&n));
    if (n == 0) {
        std::fprintf(stderr, "CUDA error: no devices supporting CUDA.\n");
        return false;
    }
    dev = std::max(0, std::min(dev, n - 1));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, dev));
    CHECK(cudaSetDevice(dev));
    int drv = 0, rt = 0;
    CHECK(cudaDriverGetVersion(&drv));
    CHECK(cudaRuntimeGetVersion(&rt));
    std::fprintf(stderr, "Using Device %d: %s, CUDA Driver Version: %d.%d, Runtime Version: %d.%d\n", dev, prop.name,
                 drv / 1000, drv % 1000, rt / 1000, rt % 1000);
    return true;
}

// Wall clock in microseconds.
inline long long cpuTimer()
{
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// cudaEvent stopwatch on a stream; read() returns elapsed milliseconds since construction.
class GpuTimer {
public:
    explicit GpuTimer(cudaStream_t s = 0) : stream_(s)
    {
        cudaEventCreate(&t0_);
        cudaEventCreate(&t1_);
        cudaEventRecord(t0_, stream_);
    }
    ~GpuTimer()
    {
        cudaEventDestroy(t0_);
        cudaEventDestroy(t1_);
    }
    float read()
    {
        float ms = 0.f;
        cudaEventRecord(t1_, stream_);
        cudaEventSynchronize(t1_);
        cudaEventElapsedTime(&ms, t0_, t1_);
        return ms;
    }

private:
    cudaEvent_t t0_, t1_;
    cudaStream_t stream_;
};

inline int iAlignUp(const int a, const int b) { return (a % b != 0) ? (a - a % b + b) : a; }
inline int iDivUp(const int a, const int b) { return (a + b - 1) / b; }
