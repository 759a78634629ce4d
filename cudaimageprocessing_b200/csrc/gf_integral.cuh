// gf_integral.cuh -- summed-area table of a uint8 image (the reference's Integral/
// module, hIntegral / hAligned4Integral, Integral/integral_d.cu:863-930).
//   integral[y][x] = sum_{i<=y, j<=x} src[i][j]        (inclusive, W x H, no zero row/column)
// Exact integer arithmetic: int32 output wraps modulo 2^32 exactly like the reference's (and NPPI's)
// int accumulators, the int64 variant never overflows (255 * W * H < 2^63).
//
// B200 shape (round 2): REDUCE, then SCAN -- the table is written exactly once and never read back.
//   (256-column strip, hb-row band) tiles, one warp each, a lane owns 8 adjacent columns (one 8-byte load per row).
//   1. gf_sat_reduce:  per tile the band total of every column (colsum[band][x]) and the strip total of every
//      row (rowtot[y][strip], one REDUX per row); reads the image, writes ~1/16 of a table.
//   2. gf_sat_carry:   rowtot -> sum of the strips to the LEFT (exclusive scan along the strips, thread per row);
//                      colsum -> sum of the bands ABOVE (exclusive scan along the bands, thread per column);
//      gf_sat_rows<no carry>: colsum[band][.] -> its prefix along x = the table row just above the band.
//   3. gf_sat_final:   per tile  acc[x] = table row above;  per row  acc[x] += (prefix inside the strip) + rowtot[y][strip],
//      store.  Reads the image a second time (L2 at 4K), writes the table.
// HBM traffic: 1 + 1 + 4 = 6 B/px for int32 against 5 B/px algorithmic (read 1, write 4), with no spin-waits (a
// single-pass decoupled look-back would save the second 1 B/px read at the price of inter-CTA flags).
// The round-1 TWO-PASS form (gf_sat_cols -> gf_sat_band_scan -> gf_sat_rows, 13 B/px: the band-local vertical prefix
// goes out and comes back) stays the choice where it measured faster -- int32 tables on the vector path that fit L2
// (4K: 38.9 against 65 us) -- see gf_sat_launch; GF_SAT_TWO_PASS=0/1 forces a form (profiles/r2_integral.jsonl).
#pragma once
#include "gf_common.cuh"
#include "gf_rt.h"
#include "gf_knobs.h"

template <class T>
struct GfSatArgs {
    const unsigned char* src; T* out; T* aux;
    T* rowtot;                      // [h][nstrips] (reduce-then-scan form)
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

// In-place exclusive scan of n elements `stride` apart, 16 independent loads at a time: the plain load-add-store loop
// costs one L2 round trip per element (measured 0.43 us per band: 58 us for the 135 bands of a 4K image).
template <class T>
__device__ __forceinline__ void gf_sat_excl_scan(T* p, int n, int64_t stride)
{
    T run = 0;
    for (int i0 = 0; i0 < n; i0 += 16) {
        T v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = i0 + k < n ? p[(int64_t)(i0 + k) * stride] : (T)0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const T t = v[k]; v[k] = run; run += t; }
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (i0 + k < n) p[(int64_t)(i0 + k) * stride] = v[k];
    }
}

template <class T>
__global__ void gf_sat_band_scan_kernel(const GfSatArgs<T> a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.w) return;
    gf_sat_excl_scan<T>(a.aux + x, a.nbands, a.w);
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

// UP = true: row y of the table (out) += carry of the bands above (aux), prefix along x.
// UP = false: rows of aux itself ([nbands][w]) become their prefix along x (reduce-then-scan form, step 2).
template <class T, bool ALIGNED, bool UP = true>
__global__ void __launch_bounds__(128) gf_sat_rows_kernel(const GfSatArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= (UP ? a.h : a.nbands)) return;
    const T* up = UP ? a.aux + (int64_t)(y / a.hb) * a.w : nullptr;     // carry of the bands above, per column
    T* row = UP ? a.out + (int64_t)y * a.ds : a.aux + (int64_t)y * a.w;
    T carry = 0;
    // chunk loads run one chunk ahead of the scan (the scan of chunk k only needs the carry of chunk k-1)
    auto load = [&](int xb, T (&v)[8]) {
        const int x0 = xb + 8 * lane;
        if (ALIGNED && x0 + 7 < a.w) {
            if (sizeof(T) == 4) {
                const int4 t0 = reinterpret_cast<const int4*>(row + x0)[0], t1 = reinterpret_cast<const int4*>(row + x0)[1];
                int4 u0 = make_int4(0, 0, 0, 0), u1 = u0;
                if (UP) { u0 = reinterpret_cast<const int4*>(up + x0)[0]; u1 = reinterpret_cast<const int4*>(up + x0)[1]; }
                v[0] = (T)(t0.x + u0.x); v[1] = (T)(t0.y + u0.y); v[2] = (T)(t0.z + u0.z); v[3] = (T)(t0.w + u0.w);
                v[4] = (T)(t1.x + u1.x); v[5] = (T)(t1.y + u1.y); v[6] = (T)(t1.z + u1.z); v[7] = (T)(t1.w + u1.w);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = row[x0 + j] + (UP ? up[x0 + j] : (T)0);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = x0 + j < a.w ? row[x0 + j] + (UP ? up[x0 + j] : (T)0) : (T)0;
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


// ---- reduce-then-scan form --------------------------------------------------------------------------
__device__ __forceinline__ unsigned gf_sat_warp_sum(unsigned v)
{
#ifdef GF_CPU_EMU
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
#else
    return __reduce_add_sync(0xffffffffu, v);
#endif
}

// the 8 source bytes of this lane in row y (zeros outside the source image)
template <class T, bool ALIGNED>
__device__ __forceinline__ void gf_sat_ld8(const GfSatArgs<T>& a, int y, int x0, bool full_in, unsigned (&b)[8])
{
    if (y < a.sh && full_in) {
        const uint2 t = *reinterpret_cast<const uint2*>(a.src + (int64_t)y * a.ss + x0);
        b[0] = t.x & 255u; b[1] = (t.x >> 8) & 255u; b[2] = (t.x >> 16) & 255u; b[3] = t.x >> 24;
        b[4] = t.y & 255u; b[5] = (t.y >> 8) & 255u; b[6] = (t.y >> 16) & 255u; b[7] = t.y >> 24;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = (y < a.sh && x0 + j < a.sw) ? a.src[(int64_t)y * a.ss + x0 + j] : 0u;
    }
}

template <class T, bool ALIGNED>
__global__ void __launch_bounds__(32) gf_sat_reduce_kernel(const GfSatArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int strip = (int)(blockIdx.x % a.nstrips), band = (int)(blockIdx.x / a.nstrips);
    const int x0 = strip * 256 + 8 * lane;
    const int y0 = band * a.hb, y1 = min(a.h, y0 + a.hb);
    const bool full_in = ALIGNED && x0 + 7 < a.sw;
    unsigned c[8];                                       // hb * 255 fits 32 bits
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = 0;
#pragma unroll 4
    for (int y = y0; y < y1; ++y) {
        unsigned b[8];
        gf_sat_ld8<T, ALIGNED>(a, y, x0, full_in, b);
        unsigned s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] += b[j]; s += b[j]; }
        s = gf_sat_warp_sum(s);
        if (lane == 0) a.rowtot[(int64_t)y * a.nstrips + strip] = (T)s;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (x0 + j < a.w) a.aux[(int64_t)band * a.w + x0 + j] = (T)c[j];
}

// thread i: row i of rowtot (exclusive scan along the strips) and column i of colsum (exclusive scan along the bands)
template <class T>
__global__ void __launch_bounds__(256) gf_sat_carry_kernel(const GfSatArgs<T> a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.w) gf_sat_excl_scan<T>(a.aux + i, a.nbands, a.w);
    if (i < a.h) gf_sat_excl_scan<T>(a.rowtot + (int64_t)i * a.nstrips, a.nstrips, 1);
}

template <class T, bool ALIGNED>
__global__ void __launch_bounds__(32) gf_sat_final_kernel(const GfSatArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int strip = (int)(blockIdx.x % a.nstrips), band = (int)(blockIdx.x / a.nstrips);
    const int x0 = strip * 256 + 8 * lane;
    const int y0 = band * a.hb, y1 = min(a.h, y0 + a.hb);
    const bool full_in = ALIGNED && x0 + 7 < a.sw;
    const bool full_out = ALIGNED && x0 + 7 < a.w;
    T acc[8];                                            // the table row above the band, then the running table row
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = x0 + j < a.w ? a.aux[(int64_t)band * a.w + x0 + j] : (T)0;
#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
        unsigned b[8];
        gf_sat_ld8<T, ALIGNED>(a, y, x0, full_in, b);
        const T left = a.rowtot[(int64_t)y * a.nstrips + strip];        // one address per warp: broadcast
#pragma unroll
        for (int j = 1; j < 8; ++j) b[j] += b[j - 1];
        unsigned incl = b[7];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const T add = (T)(incl - b[7]) + left;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += (T)b[j] + add;
        T* po = a.out + (int64_t)y * a.ds + x0;
        if (full_out) {
            if (sizeof(T) == 4) {
                reinterpret_cast<int4*>(po)[0] = make_int4((int)acc[0], (int)acc[1], (int)acc[2], (int)acc[3]);
                reinterpret_cast<int4*>(po)[1] = make_int4((int)acc[4], (int)acc[5], (int)acc[6], (int)acc[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j += 2) reinterpret_cast<longlong2*>(po)[j / 2] = make_longlong2((long long)acc[j], (long long)acc[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (x0 + j < a.w) po[j] = acc[j];
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------
#ifndef GF_NO_HOST
// scratch: optional, >= ceil(h/16) * w elements (the reference's w*h `buff` is always enough); used when the carries
// fit (two-pass form: nbands * w; reduce-then-scan form: nbands * w + h * nstrips), else a stream-ordered temporary.
// *launches = kernels launched.
template <class T>
static const char* gf_sat_launch(const unsigned char* src, T* out, T* scratch, int sw, int sh, int w, int h, int64_t ss, int64_t ds,
                                 void* stream, int* launches = nullptr)
{
    int sms = 148, mj = 0, mn = 0;
    gf_rt_device_info(&sms, &mj, &mn);
    GfSatArgs<T> a;
    a.src = src; a.out = out; a.ss = ss; a.ds = ds; a.sw = sw; a.sh = sh; a.w = w; a.h = h;
    a.nstrips = (w + 255) / 256;
    const bool al_in0 = (ss % 8 == 0) && ((uintptr_t)src % 8 == 0);
    const bool al_out0 = (ds % 8 == 0) && ((uintptr_t)out % 32 == 0) && (w % 8 == 0);
    // Which form: measured on B200 (profiles/r2_integral.jsonl).  The two-pass form wins where its 13 B/px stay in L2 and
    // three launches beat four (int32, vector path: 4K 38.9 vs 65 us, 8K 96.5 vs 117 us; anything below ~4 Mpx); the
    // reduce-then-scan form wins for int64 tables (4K 74 vs 94 us, 8K 180 vs 314, 16K^2 954 vs 2243), unaligned rows
    // (5910 x 5941: 155 vs 231 us) and tables far larger than L2 (16K^2 int32: 611 vs 707 us).
    const double mpx = (double)w * h * 1e-6;
    const bool rts_dflt = mpx >= 4.0 && (sizeof(T) == 8 || !(al_in0 && al_out0) || mpx >= 128.0);
    const bool two_pass = GF_KNOB("GF_SAT_TWO_PASS", rts_dflt ? 0 : 1) != 0;
    // bands: ~16 resident warps per SM in one wave, at least 16 rows per band; the reduce-then-scan tiles are latency
    // bound per row, so more, shorter tiles (<= 64 rows) keep more rows in flight (16K^2: 611 us at 64 rows, 844 at 443)
    long nb = (long)sms * 16 / a.nstrips;
    if (nb < 1) nb = 1;
    int hb = (int)((h + nb - 1) / nb);
    if (hb < 16) hb = 16;
    if (!two_pass && hb > 64) hb = 64;
    hb = GF_KNOB("GF_SAT_HB", hb);
    if (hb < 1) hb = 1;
    if (hb > h) hb = h;
    a.hb = hb;
    a.nbands = (h + hb - 1) / hb;
    const size_t n_col = ((size_t)a.nbands * w + 7) / 8 * 8;                 // rowtot starts 32-byte aligned
    const size_t need = two_pass ? (size_t)a.nbands * w : n_col + (size_t)h * a.nstrips;
    void* tmp = nullptr;
    // a caller's scratch is documented as >= ceil(h/16) * w elements (include/gf_b200.h)
    if (!scratch || need > (size_t)((h + 15) / 16) * w) {
        if (const char* e = gf_rt_alloc_async(&tmp, need * sizeof(T), stream)) return e;
        scratch = (T*)tmp;
    }
    a.aux = scratch;
    a.rowtot = scratch + n_col;
    const bool al_in = (ss % 8 == 0) && ((uintptr_t)src % 8 == 0);
    const bool al_out = (ds % 8 == 0) && ((uintptr_t)out % 32 == 0) && ((uintptr_t)scratch % 32 == 0) && (w % 8 == 0);
    const dim3 tiles((unsigned)((long)a.nstrips * a.nbands)), warp(32);
    if (two_pass) {
        if (al_in && al_out) { auto k = gf_sat_cols_kernel<T, true>; GF_LAUNCH(k, tiles, warp, 0, stream, a); }
        else { auto k = gf_sat_cols_kernel<T, false>; GF_LAUNCH(k, tiles, warp, 0, stream, a); }
        {
            dim3 grid((unsigned)((w + 255) / 256)), block(256);
            auto k = gf_sat_band_scan_kernel<T>;
            GF_LAUNCH(k, grid, block, 0, stream, a);
        }
        {
            dim3 grid((unsigned)((h + 3) / 4)), block(128);
            if (al_out) { auto k = gf_sat_rows_kernel<T, true, true>; GF_LAUNCH(k, grid, block, 0, stream, a); }
            else { auto k = gf_sat_rows_kernel<T, false, true>; GF_LAUNCH(k, grid, block, 0, stream, a); }
        }
        if (launches) *launches = 3;
    } else {
        if (al_in && al_out) { auto k = gf_sat_reduce_kernel<T, true>; GF_LAUNCH(k, tiles, warp, 0, stream, a); }
        else { auto k = gf_sat_reduce_kernel<T, false>; GF_LAUNCH(k, tiles, warp, 0, stream, a); }
        {
            const int n = w > h ? w : h;
            dim3 grid((unsigned)((n + 255) / 256)), block(256);
            auto k = gf_sat_carry_kernel<T>;
            GF_LAUNCH(k, grid, block, 0, stream, a);
        }
        {
            dim3 grid((unsigned)((a.nbands + 3) / 4)), block(128);
            if (al_out) { auto k = gf_sat_rows_kernel<T, true, false>; GF_LAUNCH(k, grid, block, 0, stream, a); }
            else { auto k = gf_sat_rows_kernel<T, false, false>; GF_LAUNCH(k, grid, block, 0, stream, a); }
        }
        if (al_in && al_out) { auto k = gf_sat_final_kernel<T, true>; GF_LAUNCH(k, tiles, warp, 0, stream, a); }
        else { auto k = gf_sat_final_kernel<T, false>; GF_LAUNCH(k, tiles, warp, 0, stream, a); }
        if (launches) *launches = 4;
    }
    const char* err = gf_rt_launch_error();
    if (tmp) gf_rt_free_async(tmp, stream);
    return err;
}
#endif  // GF_NO_HOST
