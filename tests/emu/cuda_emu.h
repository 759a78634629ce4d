// cuda_emu.h -- a small SIMT emulator used ONLY by the CPU test-suite to debug kernel logic
// (indexing, halos, barriers, shuffles) in a container that has nvcc but no GPU.
//
// It is TEST INFRASTRUCTURE: the product package never builds, loads or falls back to it.  The
// kernels under cudaimageprocessing_b200/csrc are compiled a second time by g++ with
// -DGF_CPU_EMU into tests/emu/libgf_emu.so, where every CUDA thread of a block is a ucontext
// fiber, __syncthreads()/__shfl_*_sync() are cooperative yields, and blocks run one after the
// other (different OpenMP threads take different blocks).  A barrier that not every live
// thread reaches is reported as a deadlock instead of hanging.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) longlong2 { long long x, y; };
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
struct int2 { int x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

typedef void* cudaStream_t;

namespace emu {

enum State { RUNNABLE = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };

struct Fiber {
    ucontext_t ctx;
    State st = RUNNABLE;
    uint3 tid;
    char* stack = nullptr;
};

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = -1;
    std::function<void()> body;
    uint32_t shfl[64][32];  // per-warp exchange buffer
    char* dyn_smem = nullptr;
};

extern thread_local Block* g_block;
extern thread_local uint3 g_threadIdx, g_blockIdx;
extern thread_local dim3 g_blockDim, g_gridDim;

void yield_to_scheduler(State s);
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

inline int linear_tid() {
    return (int)(g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z));
}
inline void* dyn_smem() { return g_block->dyn_smem; }

}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { emu::yield_to_scheduler(emu::AT_BLOCK); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_to_scheduler(emu::AT_WARP); }

template <class T>
static inline T emu_shfl(T v, int src_lane_delta_kind, int arg, int width) {
    static_assert(sizeof(T) == 4, "4-byte shuffles only");
    const int t = emu::linear_tid(), w = t >> 5, lane = t & 31;
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    emu::g_block->shfl[w][lane] = bits;
    emu::yield_to_scheduler(emu::AT_WARP);
    int src = lane;
    const int seg = lane & ~(width - 1);
    if (src_lane_delta_kind == 0) { src = lane - arg; if (src < seg) src = lane; }               // up
    else if (src_lane_delta_kind == 1) { src = lane + arg; if (src >= seg + width) src = lane; }  // down
    else if (src_lane_delta_kind == 2) { src = seg + (arg & (width - 1)); }                       // idx
    else { src = lane ^ arg; if (src >= seg + width) src = lane; }                                // xor
    uint32_t r = emu::g_block->shfl[w][src];
    // a lane whose source thread has exited or does not exist reads its own value
    if (w * 32 + src >= (int)emu::g_block->fibers.size()) r = bits;
    emu::yield_to_scheduler(emu::AT_WARP);
    T out;
    std::memcpy(&out, &r, 4);
    return out;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int w = 32) { return emu_shfl(v, 0, (int)d, w); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int w = 32) { return emu_shfl(v, 1, (int)d, w); }
template <class T> static inline T __shfl_sync(unsigned, T v, int l, int w = 32) { return emu_shfl(v, 2, l, w); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int w = 32) { return emu_shfl(v, 3, m, w); }

static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __int2float_rn(int a) { return (float)a; }
template <class T> static inline T __ldg(const T* p) { return *p; }
#ifndef __CUDACC__
using std::max;
using std::min;
#endif

// kernel<<<grid, block, smem, stream>>>(args...)
#define GF_EMU_LAUNCH(kernel, grid, block, smem, ...) \
    emu::run_grid((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
