// gf_generic.cuh -- the any-radius, any-border fused guided-filter kernel and the box filter.
//
// One CTA owns a column strip of `wc` output columns and a band of `hb` output rows and
// streams down the rows ONCE: thread t <-> extended column x0 - 2r + t.
//   stage 1  vertical running sums of {I, p, I*p, I*I} (registers, add new row / subtract the
//            row that left the window), horizontal (2r+1)-sums across threads -> means ->
//            a, b for the row 'r' behind the newest input row;
//   stage 2  horizontal sums of a, b across threads, vertical running sums through a
//            (2r+1)-row ring -> mean_a, mean_b for the row '2r' behind -> q = mean_a*I+mean_b.
// a and b never leave the SM.  HBM traffic is read I, p (+ the halo overlap, served by L2) and
// write q.  The reference does the same maths with 23 launches through 8 scratch planes
// (guided_filter.cpp:28-66) or 2 launches through A, B planes (guided_filter_d.cu:1047-1093).
//
// Horizontal sums are O(1) in r: an inclusive warp-shuffle scan per 32-column chunk, published
// to shared memory; a window is  pre[x+r] - pre[x-r-1] + (totals of the chunks in between).
// Partial sums never exceed 32*(2r+1) terms' worth, which keeps float32 error ~1e-7 relative
// after the 1/(2r+1)^2 normalisation (the reference's float32 integral image loses 1e-2 at 4K,
// SURVEY fact 4).  For r <= GF_DIRECT_R the window is summed directly (pure additions: the
// reference's fused path range r <= 7 keeps its ~1e-7 accuracy, which the uint8-level KAT needs).
#pragma once
#include "gf_common.cuh"

#define GF_DIRECT_R 8

// ---------------------------------------------------------------------------------------------
// Horizontal (2r+1)-window sums of NQ per-thread values across the CTA's threads.
// Valid for threads tid in [r, win - r).  Contains exactly one __syncthreads().
template <int NQ>
__device__ __forceinline__ void gf_hsum(const float (&v)[NQ], float (&out)[NQ], float* s_pre,
                                        float* s_tot, int r, int tid, int win)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int hi = min(tid + r, win - 1);
    if (r <= GF_DIRECT_R) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) s_pre[q * win + tid] = v[q];
        __syncthreads();
        const int lo = max(tid - r, 0);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float s = 0.f;
            for (int j = lo; j <= hi; ++j) s += s_pre[q * win + j];
            out[q] = s;
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float x = v[q];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        s_pre[q * win + tid] = x;
        if (lane == 31) s_tot[q * 32 + warp] = x;
    }
    __syncthreads();
    const int lm = tid - r - 1;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float s = s_pre[q * win + hi];
        int c0 = 0;
        if (lm >= 0) {
            s -= s_pre[q * win + lm];
            c0 = lm >> 5;
        }
        for (int c = c0; c < (hi >> 5); ++c) s += s_tot[q * 32 + c];
        out[q] = s;
    }
}

__device__ __forceinline__ float gf_ld(const GfPlane& p, int64_t frame, int row, int x)
{
    return p.ptr[frame * p.frame_stride + (int64_t)row * p.stride + (int64_t)x * p.xstep + p.coff];
}
__device__ __forceinline__ void gf_st(const GfPlane& p, int64_t frame, int row, int x, float v)
{
    p.ptr[frame * p.frame_stride + (int64_t)row * p.stride + (int64_t)x * p.xstep + p.coff] = v;
}

// ---------------------------------------------------------------------------------------------
// Models: what is summed in stage 1, how (a, b) come out of the means, how q comes out.

// Gray guide.  a = cov(I,p)/(var(I)+eps), b = mean(p) - a*mean(I)  (main.cpp:248-249;
// gCalcA/gCalcB, guided_filter_d.cu:306-323, 349-362).
struct GfGrayModel {
    static const int NI = 1, NQ1 = 4, NQ2 = 2;
    __device__ static __forceinline__ void load(const GfArgs& a, int64_t f, int row, int x, float (&v)[NQ1])
    {
        const float I = gf_ld(a.guide, f, row, x), p = gf_ld(a.src, f, row, x);
        v[0] = I; v[1] = p; v[2] = I * p; v[3] = I * I;
    }
    __device__ static __forceinline__ void solve(const float (&s)[NQ1], const GfNorm& norm, float eps, float (&ab)[NQ2])
    {
        const float mi = gf_norm_apply(s[0], norm), mp = gf_norm_apply(s[1], norm), mip = gf_norm_apply(s[2], norm),
                    mii = gf_norm_apply(s[3], norm);
        const float var = fmaf(-mi, mi, mii);
        const float cov = fmaf(-mi, mp, mip);
        const float aa = cov / (var + eps);
        ab[0] = aa;
        ab[1] = fmaf(-aa, mi, mp);
    }
    __device__ static __forceinline__ float apply(const GfArgs& a, int64_t f, int row, int x,
                                                  const float (&s)[NQ2], const GfNorm& norm)
    {
        const float I = gf_ld(a.guide, f, row, x);
        return gf_norm_apply(fmaf(s[0], I, s[1]), norm);      // gLinearTransform, guided_filter_d.cu:382-395
    }
};

// Colour guide (He et al. 2013 eqs. 19-21): 13 box planes, symmetric 3x3 inverse by cofactors.
struct GfColorModel {
    static const int NI = 3, NQ1 = 13, NQ2 = 4;
    __device__ static __forceinline__ void load(const GfArgs& a, int64_t f, int row, int x, float (&v)[NQ1])
    {
        const float* g = a.guide.ptr + f * a.guide.frame_stride + (int64_t)row * a.guide.stride + (int64_t)x * 3;
        const float i0 = g[0], i1 = g[1], i2 = g[2], p = gf_ld(a.src, f, row, x);
        v[0] = i0; v[1] = i1; v[2] = i2; v[3] = p;
        v[4] = i0 * p; v[5] = i1 * p; v[6] = i2 * p;
        v[7] = i0 * i0; v[8] = i0 * i1; v[9] = i0 * i2; v[10] = i1 * i1; v[11] = i1 * i2; v[12] = i2 * i2;
    }
    __device__ static __forceinline__ void solve(const float (&s)[NQ1], const GfNorm& norm, float eps, float (&ab)[NQ2])
    {
        float m[NQ1];
#pragma unroll
        for (int q = 0; q < NQ1; ++q) m[q] = gf_norm_apply(s[q], norm);
        const float m0 = m[0], m1 = m[1], m2 = m[2], mp = m[3];
        const float c0 = fmaf(-m0, mp, m[4]), c1 = fmaf(-m1, mp, m[5]), c2 = fmaf(-m2, mp, m[6]);
        const float s00 = fmaf(-m0, m0, m[7]) + eps, s01 = fmaf(-m0, m1, m[8]),
                    s02 = fmaf(-m0, m2, m[9]), s11 = fmaf(-m1, m1, m[10]) + eps,
                    s12 = fmaf(-m1, m2, m[11]), s22 = fmaf(-m2, m2, m[12]) + eps;
        const float i00 = s11 * s22 - s12 * s12, i01 = s02 * s12 - s01 * s22, i02 = s01 * s12 - s02 * s11,
                    i11 = s00 * s22 - s02 * s02, i12 = s01 * s02 - s00 * s12, i22 = s00 * s11 - s01 * s01;
        const float inv = 1.0f / (s00 * i00 + s01 * i01 + s02 * i02);
        const float a0 = (i00 * c0 + i01 * c1 + i02 * c2) * inv;
        const float a1 = (i01 * c0 + i11 * c1 + i12 * c2) * inv;
        const float a2 = (i02 * c0 + i12 * c1 + i22 * c2) * inv;
        ab[0] = a0; ab[1] = a1; ab[2] = a2;
        ab[3] = mp - (a0 * m0 + a1 * m1 + a2 * m2);
    }
    __device__ static __forceinline__ float apply(const GfArgs& a, int64_t f, int row, int x,
                                                  const float (&s)[NQ2], const GfNorm& norm)
    {
        const float* g = a.guide.ptr + f * a.guide.frame_stride + (int64_t)row * a.guide.stride + (int64_t)x * 3;
        return gf_norm_apply(s[0] * g[0] + s[1] * g[1] + s[2] * g[2] + s[3], norm);
    }
};

template <class M>
__host__ __device__ inline size_t gf_generic_smem_floats(int win, int r, bool ring_in_smem)
{
    size_t n = (size_t)(M::NQ1 + M::NQ2) * (win + 32);
    if (ring_in_smem) n += (size_t)M::NQ2 * (2 * r + 1) * win;
    return n;
}

// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(1024) gf_generic_kernel(const GfArgs a)
{
    constexpr int NQ1 = M::NQ1, NQ2 = M::NQ2;
    GF_DYN_SMEM(float, smem);
    const int tid = threadIdx.x, win = blockDim.x;
    const int r = a.r, k = 2 * r + 1;
    const int64_t f = blockIdx.z;

    float* s_cs = smem;                       // [NQ1][win]
    float* s_tot1 = s_cs + NQ1 * win;         // [NQ1][32]
    float* s_ab = s_tot1 + NQ1 * 32;          // [NQ2][win]
    float* s_tot2 = s_ab + NQ2 * win;         // [NQ2][32]
    float* ring = a.ring ? a.ring + ((size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)) *
                                        ((size_t)NQ2 * k * win)
                         : s_tot2 + NQ2 * 32; // [k][NQ2][win], column `tid` is private to the thread

    const int xo0 = blockIdx.x * a.wc;
    const int xe = xo0 - 2 * r + tid;         // extended column of this thread
    const int sx = gf_map(xe, a.width, a.border);
    const bool x_in = a.border != GF_TRUNCATE || (xe >= 0 && xe < a.width);
    const float cnt_x = gf_count(xe, a.width, r, a.border);
    const int yo0 = a.out_y0 + blockIdx.y * a.hb;
    const int yo1 = min(a.out_y0 + a.out_rows, yo0 + a.hb);
    const bool out_col = tid >= 2 * r && tid < 2 * r + a.wc && xe < a.width;
    const bool ab_lane = tid >= r && tid < win - r;   // threads whose stage-1 window is complete

    float cs1[NQ1], cs2[NQ2];
#pragma unroll
    for (int q = 0; q < NQ1; ++q) cs1[q] = 0.f;
#pragma unroll
    for (int q = 0; q < NQ2; ++q) cs2[q] = 0.f;
    for (int s = 0; s < k; ++s)
#pragma unroll
        for (int q = 0; q < NQ2; ++q) ring[((size_t)s * NQ2 + q) * win + tid] = 0.f;

    const int steps = (yo1 - yo0) + 4 * r;
    int slot = 0;
    for (int t = 0; t < steps; ++t) {
        const int yi = yo0 - 2 * r + t;       // newest (extended) input row
        // ---- stage 1, vertical: add the new row, drop the one that left the window
        if (sx >= 0) {
            const int sy = gf_map(yi, a.height, a.border);
            if (sy >= 0) {
                float v[NQ1];
                M::load(a, f, sy - a.buf_y0, sx, v);
#pragma unroll
                for (int q = 0; q < NQ1; ++q) cs1[q] += v[q];
            }
            if (t >= k) {
                const int so = gf_map(yi - k, a.height, a.border);
                if (so >= 0) {
                    float v[NQ1];
                    M::load(a, f, so - a.buf_y0, sx, v);
#pragma unroll
                    for (int q = 0; q < NQ1; ++q) cs1[q] -= v[q];
                }
            }
        }
        if (t < 2 * r) continue;              // uniform across the CTA

        // ---- stage 1, horizontal -> a, b of row yc
        const int yc = yi - r;
        float h1[NQ1], ab[NQ2];
        gf_hsum<NQ1>(cs1, h1, s_cs, s_tot1, r, tid, win);
        const bool y_in = a.border != GF_TRUNCATE || (yc >= 0 && yc < a.height);
        if (ab_lane && x_in && y_in) {
            M::solve(h1, gf_norm_make(cnt_x * gf_count(yc, a.height, r, a.border)), a.eps, ab);
        } else {
#pragma unroll
            for (int q = 0; q < NQ2; ++q) ab[q] = 0.f;   // outside the image / incomplete window
        }
        if (NQ2 == 2 && a.A.ptr != nullptr && out_col && yc >= yo0 && yc < yo1) {
            gf_st(a.A, f, yc - a.out_y0, xe, ab[0]);
            gf_st(a.B, f, yc - a.out_y0, xe, ab[1]);
        }

        // ---- stage 2, horizontal then vertical through the ring
        float h2[NQ2];
        gf_hsum<NQ2>(ab, h2, s_ab, s_tot2, r, tid, win);
#pragma unroll
        for (int q = 0; q < NQ2; ++q) {
            float* cell = ring + ((size_t)slot * NQ2 + q) * win + tid;
            cs2[q] += h2[q] - *cell;
            *cell = h2[q];
        }
        slot = slot + 1 == k ? 0 : slot + 1;

        if (t >= 4 * r && out_col) {
            const int yo = yi - 2 * r;
            const GfNorm norm = gf_norm_make(cnt_x * gf_count(yo, a.height, r, a.border));
            gf_st(a.dst, f, yo - a.out_y0, xe, M::apply(a, f, yo - a.buf_y0, xe, cs2, norm));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone box mean (hBoxFilter, guided_filter_d.cu:868-924) of a C-channel interleaved
// image: the stage-1 half of the kernel above.  thread <-> extended pixel column x0 - r + t.
template <int C>
__global__ void __launch_bounds__(1024) gf_box_kernel(const GfArgs a)
{
    GF_DYN_SMEM(float, smem);
    const int tid = threadIdx.x, win = blockDim.x;
    const int r = a.r, k = 2 * r + 1;
    const int64_t f = blockIdx.z;
    float* s_cs = smem;
    float* s_tot = s_cs + C * win;
    const int xe = blockIdx.x * a.wc - r + tid;
    const int sx = gf_map(xe, a.width, a.border);
    const float cnt_x = gf_count(xe, a.width, r, a.border);
    const int yo0 = a.out_y0 + blockIdx.y * a.hb;
    const int yo1 = min(a.out_y0 + a.out_rows, yo0 + a.hb);
    const bool out_col = tid >= r && tid < r + a.wc && xe < a.width;

    float cs[C];
#pragma unroll
    for (int c = 0; c < C; ++c) cs[c] = 0.f;
    const int steps = (yo1 - yo0) + 2 * r;
    for (int t = 0; t < steps; ++t) {
        const int yi = yo0 - r + t;
        if (sx >= 0) {
            const int sy = gf_map(yi, a.height, a.border);
            if (sy >= 0) {
                const float* p = a.src.ptr + f * a.src.frame_stride + (int64_t)(sy - a.buf_y0) * a.src.stride + (int64_t)sx * C;
#pragma unroll
                for (int c = 0; c < C; ++c) cs[c] += p[c];
            }
            if (t >= k) {
                const int so = gf_map(yi - k, a.height, a.border);
                if (so >= 0) {
                    const float* p = a.src.ptr + f * a.src.frame_stride + (int64_t)(so - a.buf_y0) * a.src.stride + (int64_t)sx * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) cs[c] -= p[c];
                }
            }
        }
        if (t < 2 * r) continue;
        float h[C];
        gf_hsum<C>(cs, h, s_cs, s_tot, r, tid, win);
        __syncthreads();                      // s_cs is rewritten next step (single stage: 2nd barrier)
        if (out_col) {
            const int yo = yi - r;
            const GfNorm norm = gf_norm_make(cnt_x * gf_count(yo, a.height, r, a.border));
            float* d = a.dst.ptr + f * a.dst.frame_stride + (int64_t)(yo - a.out_y0) * a.dst.stride + (int64_t)xe * C;
#pragma unroll
            for (int c = 0; c < C; ++c) d[c] = gf_norm_apply(h[c], norm);
        }
    }
}
