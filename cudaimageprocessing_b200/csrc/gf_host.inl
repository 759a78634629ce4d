// gf_host.inl -- host-buffer entry points (included at the end of gf_api.cu).
//
// gf_guided_gray_host is the end-to-end call: it uploads guide and src, filters, and downloads
// dst, pipelined in row bands over three streams so that H2D of band b+1, the kernel of band b
// and D2H of band b-1 overlap (PCIe is full duplex; the kernel is ~1% of the copy time).
#include <mutex>
#include <thread>

#include "gf_copy_pool.h"

#ifdef GF_CPU_EMU
extern "C" {
int gf_host_alloc(void** ptr, size_t bytes) { *ptr = std::malloc(bytes); return *ptr ? GF_OK : fail(GF_ERR_NOMEM, "malloc"); }
int gf_host_free(void* ptr) { std::free(ptr); return GF_OK; }
int gf_host_register(void* ptr, size_t) { return ptr ? GF_OK : fail(GF_ERR_INVALID, "null pointer"); }
int gf_host_unregister(void* ptr) { return ptr ? GF_OK : fail(GF_ERR_INVALID, "null pointer"); }
int gf_guided_gray_host(const float* guide, const float* src, float* dst, int width, int height, int r, float eps, int border)
{
    return gf_guided_gray(guide, src, dst, nullptr, nullptr, width, height, 0, 0, 0, 0, r, eps, border, nullptr);
}
}
#else
namespace {
const int kMaxBands = 16;
const int kMaxDevices = 64;
// One pipe per DEVICE (streams, events and the staging planes live on it), each with its own lock: callers on
// different GPUs never serialise, callers on the same GPU share its staging planes and take turns.
struct HostPipe {
    std::mutex mu;
    float* dev = nullptr;      // guide | src | dst planes
    size_t cap = 0;            // floats per plane
    cudaStream_t up = nullptr, comp = nullptr, down = nullptr;
    cudaEvent_t ev_up[kMaxBands], ev_k[kMaxBands], ev_down[kMaxBands];
    float* pin = nullptr;      // PAGEABLE callers only: pinned guide | src | dst planes the copies are staged through
    size_t pin_cap = 0;
    bool init = false;
};
HostPipe g_pipes[kMaxDevices];

#define GF_CU(call)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) return fail(GF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
// inside the band loop: copies touching the caller's buffers may be in flight -> drain the pipe before returning
#define GF_CU_DRAIN(call)                                                            \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) { drain(P); return fail(GF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)

// Pageable host buffers (cv::Mat data, main.cpp:229-230): cudaMemcpyAsync from pageable memory is a single-threaded
// staged copy inside the driver (4K frame: 6.9 ms against 1.5 ms from pinned memory).  The host call stages such
// buffers itself: a few threads copy each row band between the caller's memory and pinned planes while the DMA engine
// moves the previous band (profiles/r2_e2e_pageable_staged.jsonl).
bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeUnregistered;
}

void drain(HostPipe& P)
{
    cudaStreamSynchronize(P.up);
    cudaStreamSynchronize(P.comp);
    cudaStreamSynchronize(P.down);
}
}  // namespace

extern "C" {

int gf_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return fail(GF_ERR_INVALID, "null pointer");
    GF_CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return GF_OK;
}

int gf_host_free(void* ptr)
{
    GF_CU(cudaFreeHost(ptr));
    return GF_OK;
}

int gf_host_register(void* ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return fail(GF_ERR_INVALID, "null pointer");
    GF_CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return GF_OK;
}

int gf_host_unregister(void* ptr)
{
    if (!ptr) return fail(GF_ERR_INVALID, "null pointer");
    GF_CU(cudaHostUnregister(ptr));
    return GF_OK;
}

int gf_guided_gray_host(const float* guide, const float* src, float* dst, int width, int height, int r, float eps, int border)
{
    if (!guide || !src || !dst) return fail(GF_ERR_INVALID, "null image pointer");
    if (width <= 0 || height <= 0 || r < 0) return fail(GF_ERR_INVALID, "bad geometry %dx%d r=%d", width, height, r);
    int dev = 0;
    GF_CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(GF_ERR_UNSUPPORTED, "device index %d", dev);
    HostPipe& P = g_pipes[dev];
    std::lock_guard<std::mutex> lock(P.mu);
    if (!P.init) {
        GF_CU(cudaStreamCreateWithFlags(&P.up, cudaStreamNonBlocking));
        GF_CU(cudaStreamCreateWithFlags(&P.comp, cudaStreamNonBlocking));
        GF_CU(cudaStreamCreateWithFlags(&P.down, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxBands; ++i) {
            GF_CU(cudaEventCreateWithFlags(&P.ev_up[i], cudaEventDisableTiming));
            GF_CU(cudaEventCreateWithFlags(&P.ev_k[i], cudaEventDisableTiming));
            GF_CU(cudaEventCreateWithFlags(&P.ev_down[i], cudaEventDisableTiming));
        }
        P.init = true;
    }
    const size_t n = (size_t)width * height;
    if (n > P.cap) {               // grow-only staging planes; nothing is in flight here (every call drains before it returns)
        if (P.dev) GF_CU(cudaFree(P.dev));
        P.dev = nullptr;
        P.cap = 0;
        cudaError_t e = cudaMalloc((void**)&P.dev, 3 * n * sizeof(float));
        if (e != cudaSuccess) { P.dev = nullptr; return fail(GF_ERR_NOMEM, "staging planes (%zu bytes): %s", 3 * n * sizeof(float), cudaGetErrorString(e)); }
        P.cap = n;
    }
    float *dI = P.dev, *dP = P.dev + P.cap, *dQ = P.dev + 2 * P.cap;
    // pageable buffers are staged through pinned planes by the copy pool; pinned / registered ones are copied directly
    // copy threads: half of the machine's cores, 2..8.  4K frame on the 16-core GPU box (profiles/r2_e2e_pageable_staged.jsonl):
    // 4 threads 3.9-4.5 ms, 8: 3.0-3.5, 12: 2.85-3.3, 16: 2.8; the driver's own pageable copies: 6.9-8.0 ms
    static const int dflt_threads = [] { const int c = (int)std::thread::hardware_concurrency() / 2; return c < 2 ? 2 : (c > 8 ? 8 : c); }();
    const int copy_threads = GF_KNOB("GF_HOST_COPY_THREADS", dflt_threads);
    const bool staged = GF_KNOB("GF_HOST_STAGED", 1) && copy_threads > 0 && (is_pageable(guide) || is_pageable(src) || is_pageable(dst));
    if (staged && n > P.pin_cap) {
        if (P.pin) GF_CU(cudaFreeHost(P.pin));
        P.pin = nullptr;
        P.pin_cap = 0;
        cudaError_t e = cudaHostAlloc((void**)&P.pin, 3 * n * sizeof(float), cudaHostAllocDefault);
        if (e != cudaSuccess) { P.pin = nullptr; return fail(GF_ERR_NOMEM, "pinned staging planes (%zu bytes): %s", 3 * n * sizeof(float), cudaGetErrorString(e)); }
        P.pin_cap = n;
    }
    float *hI = staged ? P.pin : nullptr, *hP = staged ? P.pin + P.pin_cap : nullptr, *hQ = staged ? P.pin + 2 * P.pin_cap : nullptr;
    int nb = height / (4 * r + 64);
    // measured on B200 / PCIe Gen5 (4K frame; bench_tools/e2e_check.py, pcie_pipe.py): 1 band 1.89 ms, 2: 1.67, 4: 1.61,
    // 8: 1.75, 16: 1.65 -- chunked copies lose duplex efficiency, so few large bands win (ideal duplex: 1.27 ms)
    nb = nb < 1 ? 1 : (nb > 4 ? 4 : nb);
    // staged copies are bound by the CPU copies, which only overlap the DMA of OTHER bands: more, equal bands
    if (staged) nb = height / (4 * r + 64) < 1 ? 1 : (height / (4 * r + 64) > 12 ? 12 : height / (4 * r + 64));
    nb = GF_KNOB(staged ? "GF_HOST_STAGED_BANDS" : "GF_HOST_BANDS", nb); nb = nb < 1 ? 1 : (nb > kMaxBands ? kMaxBands : nb);
    // band b is taper % as tall as band b-1: what stays exposed at the end is the kernel and the download of the LAST band.
    // Measured (4K frame, profiles/r2_e2e_bands_taper.jsonl): 4 uniform bands 1.60 ms, taper 80 %: 1.57, 65 %: 1.54,
    // 50-55 %: 1.50, 35 %: 1.54, 25 %: 1.62 (1 band, no overlap: 1.87; upload alone ~1.2).
    int taper = staged ? 100 : GF_KNOB("GF_HOST_TAPER_PCT", 50);
    taper = taper < 20 ? 20 : (taper > 100 ? 100 : taper);
    int cut[kMaxBands + 1];
    {
        double wsum = 0.0, w = 1.0, acc = 0.0;
        for (int b = 0; b < nb; ++b) { wsum += w; w *= taper / 100.0; }
        w = 1.0;
        cut[0] = 0;
        for (int b = 0; b < nb; ++b) {
            acc += w; w *= taper / 100.0;
            cut[b + 1] = b == nb - 1 ? height : (int)(height * (acc / wsum));
            if (cut[b + 1] < cut[b]) cut[b + 1] = cut[b];
        }
    }
    int up_to = 0;
    int out_done = 0;               // staged: bands [0, out_done) have been copied from the pinned dst plane to the caller's
    auto copy_out = [&](int upto_band, bool wait) -> cudaError_t {
        for (; out_done < upto_band; ++out_done) {
            const int a = cut[out_done], e = cut[out_done + 1];
            if (e <= a) continue;
            if (!wait && cudaEventQuery(P.ev_down[out_done]) != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }
            const cudaError_t rc = cudaEventSynchronize(P.ev_down[out_done]);
            if (rc != cudaSuccess) return rc;
            GfCopyPool::get().copy(dst + (size_t)a * width, hQ + (size_t)a * width, (size_t)(e - a) * width * sizeof(float), copy_threads);
        }
        return cudaSuccess;
    };
    for (int b = 0; b < nb; ++b) {
        const int y0 = cut[b], y1 = cut[b + 1];
        if (y1 <= y0) continue;
        int need = y1 + 2 * r;
        if (need > height || b == nb - 1) need = height;
        if (need > up_to) {
            const size_t off = (size_t)up_to * width, cnt = (size_t)(need - up_to) * width * sizeof(float);
            if (staged) {
                GfCopyPool::get().copy(hI + off, guide + off, cnt, copy_threads);
                GF_CU_DRAIN(cudaMemcpyAsync(dI + off, hI + off, cnt, cudaMemcpyHostToDevice, P.up));
                GfCopyPool::get().copy(hP + off, src + off, cnt, copy_threads);
                GF_CU_DRAIN(cudaMemcpyAsync(dP + off, hP + off, cnt, cudaMemcpyHostToDevice, P.up));
            } else {
                GF_CU_DRAIN(cudaMemcpyAsync(dI + off, guide + off, cnt, cudaMemcpyHostToDevice, P.up));
                GF_CU_DRAIN(cudaMemcpyAsync(dP + off, src + off, cnt, cudaMemcpyHostToDevice, P.up));
            }
            up_to = need;
        }
        GF_CU_DRAIN(cudaEventRecord(P.ev_up[b], P.up));
        GF_CU_DRAIN(cudaStreamWaitEvent(P.comp, P.ev_up[b], 0));
        // rows [0, up_to) are resident; the band reads at most rows [y0-2r, y1+2r) after mapping
        int rc = gf_guided_gray_strip(dI, dP, dQ + (size_t)y0 * width, width, height, 0, up_to, y0, y1 - y0, width, width, width,
                                      r, eps, border, P.comp);
        if (rc) { drain(P); return rc; }
        GF_CU_DRAIN(cudaEventRecord(P.ev_k[b], P.comp));
        GF_CU_DRAIN(cudaStreamWaitEvent(P.down, P.ev_k[b], 0));
        GF_CU_DRAIN(cudaMemcpyAsync((staged ? hQ : dst) + (size_t)y0 * width, dQ + (size_t)y0 * width, (size_t)(y1 - y0) * width * sizeof(float),
                              cudaMemcpyDeviceToHost, P.down));
        if (staged) {
            GF_CU_DRAIN(cudaEventRecord(P.ev_down[b], P.down));
            GF_CU_DRAIN(copy_out(b, false));        // bands whose download has already landed, without waiting
        }
    }
    if (staged) GF_CU_DRAIN(copy_out(nb, true));
    GF_CU(cudaStreamSynchronize(P.down));
    GF_CU(cudaStreamSynchronize(P.comp));
    return GF_OK;
}

}  // extern "C"
#endif
