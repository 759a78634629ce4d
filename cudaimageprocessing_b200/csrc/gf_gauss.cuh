// gf_gauss.cuh -- separable Gaussian blur of a float32 gray image (the reference's
// GaussianFilter/ module, gGaussSplit / gGaussOptim<RADIUS,KX>, GaussianFilter/gaussian.cu:129-306, driven
// by gaussianComparasion, :409-660: kernel = cv::getGaussianKernel(2r+1, sigma), border REFLECT101
// (reflectBorder, gaussian.h), result compared with cv::GaussianBlur).
//
// One pass over the image, both 1-D convolutions on chip (8 B/px of HBM traffic: read 4, write 4):
//   CTA = 256 threads = one column strip of 256 - 2r output columns, a band of rows, thread per column.
//   per input row: coalesced load (border columns mirrored) into a double-buffered row in shared
//   memory -> horizontal convolution in the symmetric form  w0 x0 + sum_k wk (x-k + x+k)  -> the
//   thread's own column of a (2r+1)-row ring in shared memory -> vertical convolution of that
//   column, same symmetric form -> coalesced store.  One barrier per row (the ring is thread-private).
// Any radius up to 64, any width / stride / alignment.
#pragma once
#include "gf_common.cuh"
#include "gf_rt.h"

#define GF_GAUSS_MAX_R 64
#define GF_GAUSS_THREADS 256

struct GfGaussArgs {
    const float* src; float* dst;
    int64_t ss, ds;
    int width, height, r, hb, nstrips, nbands;
    float w[GF_GAUSS_MAX_R + 1];      // w[k] = weight of taps -k and +k
};

__device__ __forceinline__ int gf_gauss_reflect(int x, int n)
{
    // REFLECT101, repeated for images narrower than the radius
    if (n == 1) return 0;
    while (x < 0 || x >= n) x = x < 0 ? -x : 2 * n - 2 - x;
    return x;
}

__global__ void __launch_bounds__(GF_GAUSS_THREADS) gf_gauss_kernel(const GF_GRID_CONSTANT GfGaussArgs a)
{
    GF_DYN_SMEM(float, smem);
    const int r = a.r, K = 2 * r + 1, TW = GF_GAUSS_THREADS - 2 * r;
    float* rowbuf = smem;                               // [2][256]
    float* ring = smem + 2 * GF_GAUSS_THREADS;          // [K][TW]
    const int tid = threadIdx.x;
    const int strip = (int)(blockIdx.x % a.nstrips), band = (int)(blockIdx.x / a.nstrips);
    const int xo = strip * TW + tid;                    // output column of threads tid < TW
    const int xi = gf_gauss_reflect(strip * TW - r + tid, a.width);     // column this thread loads
    const int y0 = band * a.hb, y1 = min(a.height, y0 + a.hb);
    const bool out_thread = tid < TW && xo < a.width;
    const int steps = (y1 - y0) + 2 * r;
    float nxt = a.src[(int64_t)gf_gauss_reflect(y0 - r, a.height) * a.ss + xi];
    for (int i = 0; i < steps; ++i) {
        float* rb = rowbuf + (i & 1) * GF_GAUSS_THREADS;
        rb[tid] = nxt;
        if (i + 1 < steps) nxt = a.src[(int64_t)gf_gauss_reflect(y0 - r + i + 1, a.height) * a.ss + xi];
        __syncthreads();
        if (tid < TW) {
            const float* c = rb + tid + r;
            float h = a.w[0] * c[0];
            for (int k = 1; k <= r; ++k) h = fmaf(a.w[k], c[-k] + c[k], h);
            ring[(i % K) * TW + tid] = h;
            if (i >= 2 * r && out_thread) {
                // rows i-2r .. i of the ring are the 2r+1 taps of output row y0 + i - 2r; centre = i - r
                const int ctr = (i - r) % K;
                float v = a.w[0] * ring[ctr * TW + tid];
                for (int k = 1; k <= r; ++k) {
                    int up = ctr - k, dn = ctr + k;
                    up += up < 0 ? K : 0;
                    dn -= dn >= K ? K : 0;
                    v = fmaf(a.w[k], ring[up * TW + tid] + ring[dn * TW + tid], v);
                }
                a.dst[(int64_t)(y0 + i - 2 * r) * a.ds + xo] = v;
            }
        }
    }
}

// Radii 1..16, 16-byte aligned planes: 4 adjacent columns per thread.  The row goes through shared memory
// once (one STS.128, 3 or 5 conflict-free LDS.128 per thread for the 4 + 2R window columns), the vertical
// convolution never touches memory: every horizontally blurred value is scattered into the 2R+1 partial
// output rows it contributes to, which live in registers (the loop is unrolled over one period of 2R+1
// rows so that the accumulator index is static); a row is stored when its last tap has arrived.
#define GF_GAUSS4_THREADS 128
template <int R>
__global__ void __launch_bounds__(GF_GAUSS4_THREADS) gf_gauss4_kernel(const GF_GRID_CONSTANT GfGaussArgs a)
{
    constexpr int K = 2 * R + 1, HR = (R + 3) / 4 * 4, HG = HR / 4, WIN = 4 * GF_GAUSS4_THREADS, TW = WIN - 2 * HR;
    __shared__ __align__(16) float rowbuf[2][WIN];
    const int tid = threadIdx.x;
    const int strip = (int)(blockIdx.x % a.nstrips), band = (int)(blockIdx.x / a.nstrips);
    const int x0 = strip * TW - HR + 4 * tid;           // first of this thread's 4 columns
    const int y0 = band * a.hb, y1 = min(a.height, y0 + a.hb);
    const bool vec_in = x0 >= 0 && x0 + 3 < a.width;
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = gf_gauss_reflect(x0 + j, a.width);
    const bool conv = tid >= HG && tid < GF_GAUSS4_THREADS - HG;       // threads whose window is inside the CTA's columns
    const bool out_vec = conv && x0 + 3 < a.width, out_any = conv && x0 < a.width;
    auto load = [&](int y) {
        const float* row = a.src + (int64_t)gf_gauss_reflect(y, a.height) * a.ss;
        if (vec_in) return *reinterpret_cast<const float4*>(row + x0);
        return make_float4(row[xs[0]], row[xs[1]], row[xs[2]], row[xs[3]]);
    };
    float acc[K][4];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
    const int steps = (y1 - y0) + 2 * R;
    float4 nxt = load(y0 - R);      // one row ahead (two rows ahead measured 0-8 % slower: more registers, same stalls)
    for (int base = 0; base < steps; base += K) {
#pragma unroll
        for (int ph = 0; ph < K; ++ph) {
            const int i = base + ph;
            if (i >= steps) break;
            float* rb = rowbuf[i & 1];
            *reinterpret_cast<float4*>(rb + 4 * tid) = nxt;
            if (i + 1 < steps) nxt = load(y0 - R + i + 1);
            __syncthreads();
            if (conv) {
                float v[4 * (2 * HG + 1)];
#pragma unroll
                for (int m = 0; m < 2 * HG + 1; ++m) {
                    const float4 t = *reinterpret_cast<const float4*>(rb + 4 * (tid - HG + m));
                    v[4 * m] = t.x; v[4 * m + 1] = t.y; v[4 * m + 2] = t.z; v[4 * m + 3] = t.w;
                }
                float h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h[j] = a.w[0] * v[HR + j];
#pragma unroll
                    for (int k = 1; k <= R; ++k) h[j] = fmaf(a.w[k], v[HR + j - k] + v[HR + j + k], h[j]);
                }
                // scatter: blurred row i is tap d of output row i - d (d = 0..2R, weight w[|R - d|])
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    constexpr int KK = K;
                    const int s = (ph - d + KK) % KK;
                    const float wd = a.w[d < R ? R - d : d - R];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[s][j] = fmaf(wd, h[j], acc[s][j]);
                }
                // output row i - 2R is complete (slot (ph + 1) % K); it becomes the slot of row i + 1
                {
                    constexpr int KK = K;
                    const int s = (ph + 1) % KK;
                    if (i >= 2 * R) {
                        float* po = a.dst + (int64_t)(y0 + i - 2 * R) * a.ds + x0;
                        if (out_vec) *reinterpret_cast<float4*>(po) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
                        else if (out_any) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (x0 + j < a.width) po[j] = acc[s][j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
                }
            }
        }
    }
}

#ifndef GF_NO_HOST
#include <math.h>
// cv::getGaussianKernel(2r+1, sigma, CV_32F) restated (what gaussian.cu:437 feeds every kernel with):
// sigma <= 0 -> the fixed binomial tables for 1/3/5/7 taps, else sigma = 0.3*((n-1)*0.5 - 1) + 0.8;
// taps exp(-x^2 / (2 sigma^2)) in double, summed in double, scaled by 1/sum, rounded to float once (OpenCV 4.x).
static inline void gf_gauss_weights(int r, double sigma, float* w /* r+1: w[k] = taps -k and +k */)
{
    const int n = 2 * r + 1;
    static const float tab[4][4] = {{1.f}, {0.5f, 0.25f}, {0.375f, 0.25f, 0.0625f}, {0.28125f, 0.21875f, 0.109375f, 0.03125f}};
    if (sigma <= 0 && n <= 7) {
        for (int k = 0; k <= r; ++k) w[k] = tab[r][k];
        return;
    }
    if (sigma <= 0) sigma = 0.3 * ((n - 1) * 0.5 - 1.0) + 0.8;
    const double s2 = -0.5 / (sigma * sigma);
    double sum = 0;
    for (int i = 0; i < n; ++i) { const double x = i - r; sum += exp(s2 * x * x); }
    sum = 1.0 / sum;
    for (int k = 0; k <= r; ++k) w[k] = (float)(exp(s2 * (double)k * k) * sum);
}

static const char* gf_gauss_launch(const float* src, float* dst, int w, int h, int64_t ss, int64_t ds, int r, double sigma, void* stream,
                                   bool* fast)
{
    *fast = false;
    int sms = 148, mj = 0, mn = 0;
    gf_rt_device_info(&sms, &mj, &mn);
    GfGaussArgs a;
    a.src = src; a.dst = dst; a.ss = ss; a.ds = ds; a.width = w; a.height = h; a.r = r;
    gf_gauss_weights(r, sigma, a.w);
    for (int k = r + 1; k <= GF_GAUSS_MAX_R; ++k) a.w[k] = 0.f;
    const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0 && (ss & 3) == 0 && (ds & 3) == 0;
    if (r >= 1 && r <= 16 && aligned && !GF_KNOB("GF_GAUSS_GENERIC", 0)) {
        const int HR4 = (r + 3) / 4 * 4, TW4 = 4 * GF_GAUSS4_THREADS - 2 * HR4;
        a.nstrips = (w + TW4 - 1) / TW4;
        int cta_sm = 4;
        switch (r) {
#define GF_G4_CASE(RR) case RR: cta_sm = gf_rt_ctas_per_sm(gf_gauss4_kernel<RR>, GF_GAUSS4_THREADS, 0); break;
        GF_G4_CASE(1) GF_G4_CASE(2) GF_G4_CASE(3) GF_G4_CASE(4) GF_G4_CASE(5) GF_G4_CASE(6) GF_G4_CASE(7) GF_G4_CASE(8)
        GF_G4_CASE(9) GF_G4_CASE(10) GF_G4_CASE(11) GF_G4_CASE(12) GF_G4_CASE(13) GF_G4_CASE(14) GF_G4_CASE(15) GF_G4_CASE(16)
#undef GF_G4_CASE
        }
        // Bands.  A CTA walks its rows one barrier at a time, so the launch takes  waves x (hb + 2r)  row times:
        // measured on B200 (profiles/r1_gauss_band_sweep.jsonl) the best split is the ONE wave of resident CTAs
        // with the shortest bands (4K r=8: 37 us at hb=32, 56 us at hb=72, 53 us at hb=24 = two waves).
        const long slots = (long)sms * cta_sm;
        int hb4 = h;
        double best = 1e300;
        for (int nb = 1; nb <= h; ++nb) {
            const int hb = (h + nb - 1) / nb;
            if (hb < 8 && nb > 1) break;
            const long ctas = (long)a.nstrips * ((h + hb - 1) / hb);
            const double cost = (double)((ctas + slots - 1) / slots) * (hb + 2 * r);
            if (cost < best * 0.999) { best = cost; hb4 = hb; }
        }
        hb4 = GF_KNOB("GF_GAUSS_HB", hb4);
        if (hb4 < 1) hb4 = 1;
        if (hb4 > h) hb4 = h;
        a.hb = hb4;
        a.nbands = (h + hb4 - 1) / hb4;
        dim3 grid4((unsigned)((long)a.nstrips * a.nbands)), block4(GF_GAUSS4_THREADS);
        switch (r) {
#define GF_G4_CASE(RR) case RR: { auto k4 = gf_gauss4_kernel<RR>; GF_LAUNCH(k4, grid4, block4, 0, stream, a); } break;
        GF_G4_CASE(1) GF_G4_CASE(2) GF_G4_CASE(3) GF_G4_CASE(4) GF_G4_CASE(5) GF_G4_CASE(6) GF_G4_CASE(7) GF_G4_CASE(8)
        GF_G4_CASE(9) GF_G4_CASE(10) GF_G4_CASE(11) GF_G4_CASE(12) GF_G4_CASE(13) GF_G4_CASE(14) GF_G4_CASE(15) GF_G4_CASE(16)
#undef GF_G4_CASE
        }
        *fast = true;
        return gf_rt_launch_error();
    }
    const int TW = GF_GAUSS_THREADS - 2 * r;
    a.nstrips = (w + TW - 1) / TW;
    const size_t smem = ((size_t)2 * GF_GAUSS_THREADS + (size_t)(2 * r + 1) * TW) * sizeof(float);
    // bands: about two waves of resident CTAs, never shorter than 4r rows (the 2r warm-up rows of a band are redundant work)
    int cta_sm = (int)(gf_rt_max_smem() / (smem + 1024));
    if (cta_sm > 8) cta_sm = 8;
    if (cta_sm < 1) cta_sm = 1;
    long nb = (long)sms * cta_sm * 2 / a.nstrips;
    if (nb < 1) nb = 1;
    int hb = (int)((h + nb - 1) / nb);
    if (hb < 4 * r + 8) hb = 4 * r + 8;
    if (hb > h) hb = h;
    a.hb = hb;
    a.nbands = (h + hb - 1) / hb;
    dim3 grid((unsigned)((long)a.nstrips * a.nbands)), block(GF_GAUSS_THREADS);
    auto k = gf_gauss_kernel;
    if (const char* e = gf_rt_set_smem(k, smem)) return e;
    GF_LAUNCH(k, grid, block, smem, stream, a);
    return gf_rt_launch_error();
}
#endif  // GF_NO_HOST
