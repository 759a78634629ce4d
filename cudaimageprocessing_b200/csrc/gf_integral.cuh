// gf_integral.cuh -- summed-area table of a uint8 image (the reference's Integral/
// module, hIntegral / hAligned4Integral, Integral/integral_d.cu:863-930).
//   integral[y][x] = sum_{i<=y, j<=x} src[i][j]        (inclusive, W x H, no zero row/column)
// Exact integer arithmetic: int32 output wraps modulo 2^32 exactly like the reference's (and NPPI's)
// int accumulators, the int64 variant never overflows (255 * W * H < 2^63).
//
// B200 shape: three launches, every global access coalesced, warps as independent workers.
//   1. gf_sat_cols:  (256-column strip, hb-row band) tiles.  A lane owns 8 adjacent columns (one
//      8-byte load per row), keeps their running column sums in registers and writes them: the
//      band-local VERTICAL prefix.  The band's column totals go to aux[band][x].
//   2. gf_sat_band_scan: exclusive scan of aux over the bands (one thread per column; tiny).
//   3. gf_sat_rows:  one warp per row.  Per 256-column chunk: add the carry of the bands above,
//      lane-local prefix of 8 + 5-step shuffle scan of the lane totals + the row carry, store.
// HBM traffic: 1 + 4 (pass 1) + 4 + 4 (pass 3) = 13 B/px for int32 (the algorithmic minimum, read
// 1 write 4, needs a 2-D decoupled look-back; round 2).
#pragma once
#include "gf_common.cuh"
#include "gf_rt.h"

template <class T>
struct GfSatArgs {
    const unsigned char* src; T* out; T* aux;
    int64_t ss, ds;                 // row strides in elements
    int sw, sh;                     // source size (pixels outside contribute zero)
    int w, h;                       // output size (>= source size for the aligned variant)
    int hb, nstrips, nbands;
};

template <class T, bool ALIGNED>
__global__ void __launch_bounds__(32) gf_sat_cols_kernel(const GfSatArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int strip = (int)(blockIdx.x % a.nstrips), band = (int)(blockIdx.x / a.nstrips);
    const int x0 = strip * 256 + 8 * lane;
    const int y0 = band * a.hb, y1 = min(a.h, y0 + a.hb);
    T v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0;
    const bool full_in = ALIGNED && x0 + 7 < a.sw;       // 8 source bytes, 8-byte aligned
    const bool full_out = ALIGNED && x0 + 7 < a.w;
#pragma unroll 4
    for (int y = y0; y < y1; ++y) {
        unsigned b[8];
        if (y < a.sh && full_in) {
            const uint2 t = *reinterpret_cast<const uint2*>(a.src + (int64_t)y * a.ss + x0);
            b[0] = t.x & 255u; b[1] = (t.x >> 8) & 255u; b[2] = (t.x >> 16) & 255u; b[3] = t.x >> 24;
            b[4] = t.y & 255u; b[5] = (t.y >> 8) & 255u; b[6] = (t.y >> 16) & 255u; b[7] = t.y >> 24;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = (y < a.sh && x0 + j < a.sw) ? a.src[(int64_t)y * a.ss + x0 + j] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += (T)b[j];
        T* po = a.out + (int64_t)y * a.ds + x0;
        if (full_out) {
            if (sizeof(T) == 4) {
                reinterpret_cast<int4*>(po)[0] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
                reinterpret_cast<int4*>(po)[1] = make_int4((int)v[4], (int)v[5], (int)v[6], (int)v[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j += 2) reinterpret_cast<longlong2*>(po)[j / 2] = make_longlong2((long long)v[j], (long long)v[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (x0 + j < a.w) po[j] = v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (x0 + j < a.w) a.aux[(int64_t)band * a.w + x0 + j] = v[j];
}

template <class T>
__global__ void gf_sat_band_scan_kernel(const GfSatArgs<T> a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.w) return;
    T run = 0;
    for (int b = 0; b < a.nbands; ++b) {
        const T t = a.aux[(int64_t)b * a.w + x];
        a.aux[(int64_t)b * a.w + x] = run;
        run += t;
    }
}

template <class T>
__device__ __forceinline__ T gf_sat_shfl_up(T v, int d)
{
    if (sizeof(T) == 4) return (T)__shfl_up_sync(0xffffffffu, (int)v, d);
    const long long x = (long long)v;
    const int lo = __shfl_up_sync(0xffffffffu, (int)(x & 0xffffffffll), d);
    const int hi = __shfl_up_sync(0xffffffffu, (int)(x >> 32), d);
    return (T)(((long long)hi << 32) | (unsigned)lo);
}
template <class T>
__device__ __forceinline__ T gf_sat_shfl(T v, int l)
{
    if (sizeof(T) == 4) return (T)__shfl_sync(0xffffffffu, (int)v, l);
    const long long x = (long long)v;
    const int lo = __shfl_sync(0xffffffffu, (int)(x & 0xffffffffll), l);
    const int hi = __shfl_sync(0xffffffffu, (int)(x >> 32), l);
    return (T)(((long long)hi << 32) | (unsigned)lo);
}

template <class T, bool ALIGNED>
__global__ void __launch_bounds__(128) gf_sat_rows_kernel(const GfSatArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= a.h) return;
    const T* up = a.aux + (int64_t)(y / a.hb) * a.w;     // carry of the bands above, per column
    T* row = a.out + (int64_t)y * a.ds;
    T carry = 0;
    // chunk loads run one chunk ahead of the scan (the scan of chunk k only needs the carry of chunk k-1)
    auto load = [&](int xb, T (&v)[8]) {
        const int x0 = xb + 8 * lane;
        if (ALIGNED && x0 + 7 < a.w) {
            if (sizeof(T) == 4) {
                const int4 t0 = reinterpret_cast<const int4*>(row + x0)[0], t1 = reinterpret_cast<const int4*>(row + x0)[1];
                const int4 u0 = reinterpret_cast<const int4*>(up + x0)[0], u1 = reinterpret_cast<const int4*>(up + x0)[1];
                v[0] = (T)(t0.x + u0.x); v[1] = (T)(t0.y + u0.y); v[2] = (T)(t0.z + u0.z); v[3] = (T)(t0.w + u0.w);
                v[4] = (T)(t1.x + u1.x); v[5] = (T)(t1.y + u1.y); v[6] = (T)(t1.z + u1.z); v[7] = (T)(t1.w + u1.w);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = row[x0 + j] + up[x0 + j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = x0 + j < a.w ? row[x0 + j] + up[x0 + j] : (T)0;
        }
    };
    T nv[8];
    load(0, nv);
    for (int xb = 0; xb < a.w; xb += 256) {
        const int x0 = xb + 8 * lane;
        T v[8];
        const bool full = ALIGNED && x0 + 7 < a.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = nv[j];
        if (xb + 256 < a.w) load(xb + 256, nv);
#pragma unroll
        for (int j = 1; j < 8; ++j) v[j] += v[j - 1];
        T incl = v[7];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T t = gf_sat_shfl_up<T>(incl, d);
            if (lane >= d) incl += t;
        }
        const T add = incl - v[7] + carry;
        if (full) {
            if (sizeof(T) == 4) {
                reinterpret_cast<int4*>(row + x0)[0] = make_int4((int)(v[0] + add), (int)(v[1] + add), (int)(v[2] + add), (int)(v[3] + add));
                reinterpret_cast<int4*>(row + x0)[1] = make_int4((int)(v[4] + add), (int)(v[5] + add), (int)(v[6] + add), (int)(v[7] + add));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) row[x0 + j] = v[j] + add;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (x0 + j < a.w) row[x0 + j] = v[j] + add;
        }
        carry += gf_sat_shfl<T>(incl, 31);
    }
}

// ---- host side ------------------------------------------------------------------------------------
#ifndef GF_NO_HOST
// scratch: at least nbands * w elements of T (the reference's `buff` of w*h ints is always enough); NULL = stream-ordered temporary
template <class T>
static const char* gf_sat_launch(const unsigned char* src, T* out, T* scratch, int sw, int sh, int w, int h, int64_t ss, int64_t ds,
                                 void* stream)
{
    int sms = 148, mj = 0, mn = 0;
    gf_rt_device_info(&sms, &mj, &mn);
    GfSatArgs<T> a;
    a.src = src; a.out = out; a.ss = ss; a.ds = ds; a.sw = sw; a.sh = sh; a.w = w; a.h = h;
    a.nstrips = (w + 255) / 256;
    // bands: ~16 resident warps per SM in one wave, at least 16 rows per band
    long nb = (long)sms * 16 / a.nstrips;
    if (nb < 1) nb = 1;
    int hb = (int)((h + nb - 1) / nb);
    if (hb < 16) hb = 16;
    if (hb > h) hb = h;
    a.hb = hb;
    a.nbands = (h + hb - 1) / hb;
    void* tmp = nullptr;
    if (!scratch) {
        if (const char* e = gf_rt_alloc_async(&tmp, (size_t)a.nbands * w * sizeof(T), stream)) return e;
        scratch = (T*)tmp;
    }
    a.aux = scratch;
    const bool al_in = (ss % 8 == 0) && ((uintptr_t)src % 8 == 0);
    const bool al_out = (ds % 8 == 0) && ((uintptr_t)out % 32 == 0) && ((uintptr_t)scratch % 32 == 0) && (w % 8 == 0);
    {
        dim3 grid((unsigned)((long)a.nstrips * a.nbands)), block(32);
        if (al_in && al_out) { auto k = gf_sat_cols_kernel<T, true>; GF_LAUNCH(k, grid, block, 0, stream, a); }
        else { auto k = gf_sat_cols_kernel<T, false>; GF_LAUNCH(k, grid, block, 0, stream, a); }
    }
    {
        dim3 grid((unsigned)((w + 255) / 256)), block(256);
        auto k = gf_sat_band_scan_kernel<T>;
        GF_LAUNCH(k, grid, block, 0, stream, a);
    }
    {
        dim3 grid((unsigned)((h + 3) / 4)), block(128);
        if (al_out) { auto k = gf_sat_rows_kernel<T, true>; GF_LAUNCH(k, grid, block, 0, stream, a); }
        else { auto k = gf_sat_rows_kernel<T, false>; GF_LAUNCH(k, grid, block, 0, stream, a); }
    }
    const char* err = gf_rt_launch_error();
    if (tmp) gf_rt_free_async(tmp, stream);
    return err;
}
#endif  // GF_NO_HOST
