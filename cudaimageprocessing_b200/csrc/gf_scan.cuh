// gf_scan.cuh -- the SCAN path: box means of any radius from row prefix sums (north_star: "an integral-image prefix
// scan ... or a sliding-window column-sum box filter; the choice is picked from measurement").
//
// What the reference does (hBoxFilter, guided_filter_d.cu:868-924): a whole-image float32 integral image (Blelloch
// row scans :9-149, one block per COLUMN for the column scan :152-238) and four gathers per pixel (:241-270); at 4K the
// float32 table is off by ~1e-2.  The B200 form keeps the prefix one-dimensional and in float64:
//   pass 1  gf_rowprefix_kernel   P[y][x] = sum_{i<=x} v[y][i]  (v = a plane or the product of two planes: I*p and
//                                 I*I are never materialised); one CTA per row, 8 columns per thread, block scan
//   pass 2  gf_boxcols_kernel     thread per column, a band of rows: horizontal window sum = two (border: up to six)
//                                 gathers from P, vertical window by a running sum over the row map; mean = sum / count
// 4 + 8 + 16 + 4 = 32 B/px of HBM/L2 traffic per box mean (the sliding-window kernels move 12 B/px for the WHOLE
// filter), so this path is the any-radius fallback (r >= 248, where the streaming kernels run out of threads for their
// 4r halo) and the measured alternative of BASELINE configs[3] -- see profiles/r2_scan_vs_sliding.jsonl.
// Single reflections only (r < width, r < height); anything else stays GF_ERR_UNSUPPORTED.
#pragma once
#include "gf_common.cuh"

template <class TIn, class TAcc>
struct GfScanPrefixArgsT {
    const TIn* a; const TIn* b;         // MODE 1: v = a * b
    TAcc* P;
    int width, height;
    int64_t sa, sb, sp;                 // row strides (elements)
};
typedef GfScanPrefixArgsT<float, double> GfScanPrefixArgs;
// uint8 planes: exact integer prefixes (north_star: "bit-exact integral sums for uint8 input"); 64-bit accumulators
// never overflow (255^2 x 2^31 columns < 2^63)
typedef GfScanPrefixArgsT<unsigned char, long long> GfScanPrefixArgsU8;

// One CTA scans one row at a time (grid-stride over rows): 256 threads x 8 columns per chunk, thread totals scanned
// through shared memory (Hillis-Steele, float64), carry from chunk to chunk.
template <int MODE, class TIn, class TAcc>
__global__ void __launch_bounds__(256) gf_rowprefix_kernel(const GfScanPrefixArgsT<TIn, TAcc> g)
{
    __shared__ TAcc tot[2][256];
    const int t = threadIdx.x;
    for (int y = blockIdx.x; y < g.height; y += gridDim.x) {
        const TIn* ra = g.a + (int64_t)y * g.sa;
        const TIn* rb = MODE == 1 ? g.b + (int64_t)y * g.sb : nullptr;
        TAcc* rp = g.P + (int64_t)y * g.sp;
        TAcc carry = 0;
        for (int x0 = 0; x0 < g.width; x0 += 2048) {
            const int x = x0 + 8 * t;
            TAcc v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = 0;
                if (x + i < g.width) {
                    if (sizeof(TIn) == 1) v[i] = MODE == 1 ? (TAcc)((int)ra[x + i] * (int)rb[x + i]) : (TAcc)ra[x + i];   // exact integers
                    else v[i] = (TAcc)(MODE == 1 ? ra[x + i] * rb[x + i] : ra[x + i]);                                   // float32 product, as the reference's gMultiply
                }
            }
#pragma unroll
            for (int i = 1; i < 8; ++i) v[i] += v[i - 1];
            int cur = 0;
            tot[0][t] = v[7];
            __syncthreads();
#pragma unroll 1
            for (int d = 1; d < 256; d <<= 1) {
                const TAcc s = tot[cur][t] + (t >= d ? tot[cur][t - d] : (TAcc)0);
                tot[cur ^ 1][t] = s;
                cur ^= 1;
                __syncthreads();
            }
            const TAcc before = carry + (t > 0 ? tot[cur][t - 1] : (TAcc)0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (x + i < g.width) rp[x + i] = before + v[i];
            carry += tot[cur][255];
            __syncthreads();
        }
    }
}

template <class TAcc, class TOut>
struct GfScanBoxArgsT {
    const TAcc* P; TOut* out;
    int width, height, r, border, hb;
    int64_t sp, so;
};
typedef GfScanBoxArgsT<double, float> GfScanBoxArgs;            // means of float planes
typedef GfScanBoxArgsT<long long, long long> GfScanBoxArgsU8;  // exact window SUMS of uint8 planes

// sum of v[y][x-r .. x+r] under the border rule, from the row prefix (single reflection: r < n)
template <class TAcc>
__device__ __forceinline__ TAcc gf_scan_hsum(const TAcc* __restrict__ P, int x, int r, int n, int border)
{
    const int lo = x - r, hi = x + r;
    const int l = lo < 0 ? 0 : lo, h = hi > n - 1 ? n - 1 : hi;
    TAcc s = P[h] - (l > 0 ? P[l - 1] : (TAcc)0);
    if (border != GF_TRUNCATE) {
        const int e = border == GF_REFLECT ? 1 : 0;
        if (lo < 0) {                       // indices lo..-1 mirror to [1-e, -lo-e]
            const int b0 = 1 - e, b1 = -lo - e;
            if (b1 >= b0) s += P[b1] - (b0 > 0 ? P[b0 - 1] : (TAcc)0);
        }
        if (hi > n - 1) {                   // indices n..hi mirror to [2n-2+e-hi, n-2+e]
            const int b0 = 2 * n - 2 + e - hi, b1 = n - 2 + e;
            if (b1 >= b0) s += P[b1] - (b0 > 0 ? P[b0 - 1] : (TAcc)0);
        }
    }
    return s;
}

__device__ __forceinline__ int gf_scan_map(int y, int n, int border)      // -1: contributes nothing
{
    if (y >= 0 && y < n) return y;
    if (border == GF_TRUNCATE) return -1;
    const int e = border == GF_REFLECT ? 1 : 0;
    return y < 0 ? -y - e : 2 * n - 2 + e - y;
}

// MEAN: out = sum / (in-image pixel count of the window) as float; otherwise the raw window sum (exact for integers)
template <class TAcc, class TOut, bool MEAN>
__global__ void __launch_bounds__(128) gf_boxcols_kernel(const GfScanBoxArgsT<TAcc, TOut> g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.width) return;
    const int y0 = blockIdx.y * g.hb;
    const int y1 = y0 + g.hb < g.height ? y0 + g.hb : g.height;
    const int r = g.r;
    auto h = [&](int y) -> TAcc {
        const int m = gf_scan_map(y, g.height, g.border);
        return m < 0 ? (TAcc)0 : gf_scan_hsum<TAcc>(g.P + (int64_t)m * g.sp, x, r, g.width, g.border);
    };
    TAcc s = 0;
    for (int y = y0 - r; y < y0 + r; ++y) s += h(y);
    const float cx = gf_count(x, g.width, r, g.border);
    for (int y = y0; y < y1; ++y) {
        s += h(y + r);
        if (MEAN) {
            const float cnt = cx * gf_count(y, g.height, r, g.border);
            g.out[(int64_t)y * g.so + x] = (TOut)((double)s / (double)cnt);
        } else {
            g.out[(int64_t)y * g.so + x] = (TOut)s;
        }
        s -= h(y - r);
    }
}

// a = (ipm - im pm) / (iim - im^2 + eps), b = pm - a im  (guided_filter_d.cu:306-323, 349-362 on the four means)
struct GfScanAbArgs {
    const float* im; const float* pm; float* ipm_a; float* iim_b;     // a overwrites ipm, b overwrites iim
    int width, height; int64_t s;
    float eps;
};
__global__ void __launch_bounds__(256) gf_scan_ab_kernel(const GfScanAbArgs g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.width) return;
    for (int y = blockIdx.y; y < g.height; y += gridDim.y) {
        const int64_t i = (int64_t)y * g.s + x;
        const float im = g.im[i], pm = g.pm[i];
        const float num = fmaf(pm, -im, g.ipm_a[i]);
        const float den = fmaf(-im, im, g.iim_b[i] + g.eps);
        const float a = num / den;
        g.ipm_a[i] = a;
        g.iim_b[i] = fmaf(a, -im, pm);
    }
}
