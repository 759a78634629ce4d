// gf_fast.cuh -- the tuned fused guided-filter kernel for gray float32 planes (sm_100a).
//
// Same dataflow as gf_generic_kernel (one pass down a column strip, a/b never leave the SM)
// but built around what bounds it on B200 (bench_tools/microbench.cu): shared-memory LDS/STS
// and warp shuffles share one ~1 warp-instruction/clk/SM pipe and shared memory caps the
// resident warps (the stage-2 row ring is 2*(2R+1) floats per column), so the kernel keeps
// everything it can in registers, moves data between lanes as rarely as possible and hides
// latency inside each warp instead of across many warps:
//   * each thread owns K=4 ADJACENT columns: 128-bit loads/stores, four independent running
//     column sums per quantity in registers;
//   * horizontal (2R+1)-window sums with compile-time R are assembled from per-thread block
//     prefixes/suffixes/totals of the 4 columns: window(4l+j) = suf_{l+dL}[oL] + totals in
//     between + pre_{l+dR}[oR].  Only additions (float32 error stays ~1e-7 relative, the
//     quality of the reference's fused path) and 8 + (block span - 2) shuffles per quantity
//     per 128 pixels instead of a 5-step scan per pixel; all shuffles stay inside the warp;
//   * a warp's 32 lanes overlap its neighbour warps by H1 = ceil(R/4) lanes on each side, so
//     stage 1 needs nothing from other warps; the a, b of the H1 edge lanes (8 floats each)
//     come from the neighbour through a double-buffered shared-memory exchange: ONE barrier
//     per row, no divergent code around the shuffles;
//   * software pipeline: stage 2 of row t-1 and stage 1 of row t are independent chains in one
//     loop iteration, and the global loads of row t+1 are issued a full iteration early;
//   * the row that leaves the vertical window is re-read from global memory (an L1/L2 hit
//     2R+1 rows later) instead of being parked in shared memory; only the stage-2 ring lives
//     in shared memory, one float4 per thread and quantity.
// HBM traffic stays at read I, p once + write q once; the halo columns/rows between warps,
// strips and bands are L1/L2 hits.
#pragma once
#include "gf_common.cuh"
#include "gf_job.h"
#include "gf_rt.h"

#define GF_FAST_K 4

__device__ __forceinline__ float gf_rcp(float d)
{
    float r = __fdividef(1.0f, d);          // MUFU.RCP (1 ulp)
    return fmaf(fmaf(-d, r, 1.0f), r, r);   // one Newton step -> ~0.5 ulp
}

template <int R>
struct GfFastGeom {
    static constexpr int H1 = (R + 3) / 4;            // halo lanes per side per stage
    static constexpr int OL = 32 - 2 * H1;            // lanes of a warp that produce output
    __host__ __device__ static constexpr int floordiv4(int v) { return v >= 0 ? v / 4 : -((-v + 3) / 4); }
    __host__ __device__ static constexpr int dL(int j) { return floordiv4(j - R); }
    __host__ __device__ static constexpr int oL(int j) { return (j - R) - 4 * dL(j); }
    __host__ __device__ static constexpr int dR(int j) { return floordiv4(j + R); }
    __host__ __device__ static constexpr int oR(int j) { return (j + R) - 4 * dR(j); }
    static constexpr int dLmin = floordiv4(0 - R);
    static constexpr int dRmax = floordiv4(3 + R);
};

// a, b of the H1 lanes next to each warp edge, for the neighbour warp; double-buffered by row parity.
//   side 0: written by lanes [H1, 2H1)          -> read by the LEFT  neighbour's lanes [32-H1, 32)
//   side 1: written by lanes [32-2H1, 32-H1)    -> read by the RIGHT neighbour's lanes [0, H1)
template <int R, int NW>
struct GfExchange {
    float4 a[2][NW][2][GfFastGeom<R>::H1];
    float4 b[2][NW][2][GfFastGeom<R>::H1];
};

// Block sums of the thread's 4 columns.
struct GfBlock {
    float pre[4];   // pre[o] = c0 + .. + co     (pre[3] = total)
    float suf[4];   // suf[o] = co + .. + c3     (suf[0] = total)
};

__device__ __forceinline__ GfBlock gf_block(const float (&c)[4])
{
    GfBlock b;
    b.pre[0] = c[0];
    b.pre[1] = c[0] + c[1];
    b.pre[2] = b.pre[1] + c[2];
    b.pre[3] = b.pre[2] + c[3];
    b.suf[3] = c[3];
    b.suf[2] = c[2] + c[3];
    b.suf[1] = c[1] + b.suf[2];
    b.suf[0] = b.pre[3];
    return b;
}

// register x of the lane D lanes away (wraps inside the warp: lanes whose source is outside
// produce a value nobody uses)
template <int D>
__device__ __forceinline__ float gf_fetch(float x, int lane)
{
    return __shfl_sync(0xffffffffu, x, (lane + D) & 31);
}

template <int R, int D, int DEND>
struct GfMidLoop {   // tn[d - dLmin] = total of the block d lanes away, d in [D, DEND)
    __device__ static __forceinline__ void run(float total, float (&tn)[32], int lane)
    {
        tn[D - GfFastGeom<R>::dLmin] = (D == 0) ? total : gf_fetch<D>(total, lane);
        GfMidLoop<R, D + 1, DEND>::run(total, tn, lane);
    }
};
template <int R, int DEND>
struct GfMidLoop<R, DEND, DEND> {
    __device__ static __forceinline__ void run(float, float (&)[32], int) {}
};

template <int R, int J>
__device__ __forceinline__ float gf_window_j(const float (&c)[4], const GfBlock& b, const float (&tn)[32], int lane)
{
    using G = GfFastGeom<R>;
    constexpr int dl = G::dL(J), ol = G::oL(J), dr = G::dR(J), orr = G::oR(J);
    if (dl == 0 && dr == 0) {             // window inside the thread's own block (R <= 1)
        float s = c[ol];
#pragma unroll
        for (int o = ol + 1; o <= orr; ++o) s += c[o];
        return s;
    }
    const float left = (dl == 0) ? b.suf[ol] : gf_fetch<dl>(b.suf[ol], lane);
    const float right = (dr == 0) ? b.pre[orr] : gf_fetch<dr>(b.pre[orr], lane);
    float s = left;
#pragma unroll
    for (int d = dl + 1; d <= dr - 1; ++d) s += tn[d - G::dLmin];
    return s + right;
}

// (2R+1)-window sums of the 4 columns of every lane; complete for lanes [H1, 32-H1).
template <int R>
__device__ __forceinline__ void gf_window(const float (&c)[4], float (&out)[4], int lane)
{
    using G = GfFastGeom<R>;
    const GfBlock b = gf_block(c);
    float tn[32];
    GfMidLoop<R, G::dLmin + 1, (G::dRmax - 1 >= G::dLmin + 1 ? G::dRmax : G::dLmin + 1)>::run(b.pre[3], tn, lane);
    out[0] = gf_window_j<R, 0>(c, b, tn, lane);
    out[1] = gf_window_j<R, 1>(c, b, tn, lane);
    out[2] = gf_window_j<R, 2>(c, b, tn, lane);
    out[3] = gf_window_j<R, 3>(c, b, tn, lane);
}

struct GfFastArgs {
    const float* guide; const float* src; float* dst; float* A; float* B;
    int64_t gs, ss, ds, abs_;            // row strides (floats)
    int64_t gfs, sfs, dfs, abfs;         // frame strides (floats)
    int width, height, buf_y0, buf_rows, out_y0, out_rows, border, hb;
    float eps;
    float* ring;                         // global ring scratch (RING_GLOBAL) or nullptr
};

// Everything one thread carries down its strip.
template <int R, int NW>
struct GfFastCtx {
    static constexpr int H1 = GfFastGeom<R>::H1, NT = NW * 32, KW = 2 * R + 1;
    const float* gI; const float* gP; float* gQ; float* gA; float* gB;   // frame bases, already at column x0
    int64_t gs, ss, ds, abs_;
    GfExchange<R, NW>* xch;
    float4* ring;            // this thread's ring cells: ring[(slot*2+q)*NT]
    int lane, warp, x0, width, height, border, buf_y0, out_y0, yo0, yo1;
    bool vec_ok, trunc, s1_lane, out_lane, has_ab;
    float eps;
    GfNorm nk;               // 1 / (2R+1)^2
    float cnt_x[4];
    bool x_in[4];
    int sx[4];
    // state
    float cI[4], cP[4], cIP[4], cII[4], sA[4], sB[4], fA[4], fB[4], va[4], vb[4];
    float nI[4], nP[4], oI[4], oP[4], ctr[4];
    int slot;
};

// 4 pixels of buffer row `row` (< 0: nothing) for this thread; VEC: all 4 columns inside, aligned.
template <bool VEC, int R, int NW>
__device__ __forceinline__ void gf_ld_row(const GfFastCtx<R, NW>& c, const float* base, int64_t stride, int row, float (&v)[4])
{
    if (row < 0) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
    const float* p = base + (int64_t)row * stride;
    if (VEC || c.vec_ok) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = c.sx[j] >= 0 ? p[c.sx[j] - c.x0] : 0.f;
    }
}

// One loop iteration t: phase A = stage 2 (+ output) of the row handled by the previous
// iteration, phase B = stage 1 of input row yi = yo0 - 2R + t.  STEADY: every stage is active,
// every row touched is an interior row (no border mapping, y-normalisation = 1/(2R+1)) and
// the whole warp loads with 128-bit accesses -- straight-line code, no branches but the store.
template <bool STEADY, int R, int NW>
__device__ __forceinline__ void gf_fast_iter(GfFastCtx<R, NW>& c, int t, int steps)
{
    using G = GfFastGeom<R>;
    constexpr int H1 = G::H1, NT = NW * 32, KW = 2 * R + 1;
    const int yi = c.yo0 - 2 * R + t;
    const int lane = c.lane;

    // ================= phase A: stage 2 of centre row yi-1-R =================
    if (STEADY || t - 1 >= 2 * R) {
        if (NW > 1) {
            // a, b of the H1 edge lanes come from the neighbour warps (published last iteration).
            // Branch-free: every lane reads some valid cell, edge lanes keep what they read.
            const int buf = (t - 1) & 1;
            const bool le = lane < H1, re = lane >= 32 - H1;
            const int side = le ? 1 : 0;
            const int w = le ? c.warp - 1 : c.warp + 1;
            const bool has = (le && c.warp > 0) || (re && c.warp < NW - 1);
            const int wi = has ? w : c.warp;
            const int hi = le ? lane : (re ? lane - (32 - H1) : 0);
            const float4 ta = c.xch->a[buf][wi][side][hi], tb = c.xch->b[buf][wi][side][hi];
            const bool edge = le || re;
            c.va[0] = edge ? (has ? ta.x : 0.f) : c.va[0]; c.va[1] = edge ? (has ? ta.y : 0.f) : c.va[1];
            c.va[2] = edge ? (has ? ta.z : 0.f) : c.va[2]; c.va[3] = edge ? (has ? ta.w : 0.f) : c.va[3];
            c.vb[0] = edge ? (has ? tb.x : 0.f) : c.vb[0]; c.vb[1] = edge ? (has ? tb.y : 0.f) : c.vb[1];
            c.vb[2] = edge ? (has ? tb.z : 0.f) : c.vb[2]; c.vb[3] = edge ? (has ? tb.w : 0.f) : c.vb[3];
        }
        float hA[4], hB[4];
        gf_window<R>(c.va, hA, lane);
        gf_window<R>(c.vb, hB, lane);
        {
            float4* ca = c.ring + (size_t)(c.slot * 2 + 0) * NT;
            float4* cb = c.ring + (size_t)(c.slot * 2 + 1) * NT;
            const float4 oa = *ca, ob = *cb;
            c.sA[0] += hA[0] - oa.x; c.sA[1] += hA[1] - oa.y; c.sA[2] += hA[2] - oa.z; c.sA[3] += hA[3] - oa.w;
            c.sB[0] += hB[0] - ob.x; c.sB[1] += hB[1] - ob.y; c.sB[2] += hB[2] - ob.z; c.sB[3] += hB[3] - ob.w;
            *ca = make_float4(hA[0], hA[1], hA[2], hA[3]);
            *cb = make_float4(hB[0], hB[1], hB[2], hB[3]);
            // re-seed the running sums from pure additions every 2R+1 rows (see gf_wp.cuh)
#pragma unroll
            for (int j = 0; j < 4; ++j) { c.fA[j] += hA[j]; c.fB[j] += hB[j]; }
            c.slot = c.slot + 1;
            if (c.slot == KW) {
                c.slot = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) { c.sA[j] = c.fA[j]; c.sB[j] = c.fB[j]; c.fA[j] = 0.f; c.fB[j] = 0.f; }
            }
        }
        if (STEADY || t - 1 >= 4 * R) {            // q of row yo = yi-1-2R; its guide row is in ctr
            const int yo = yi - 1 - 2 * R;
            float q[4];
            if (STEADY || !c.trunc) {
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = gf_norm_apply(fmaf(c.sA[j], c.ctr[j], c.sB[j]), c.nk);
            } else {
                const float cnt_y = gf_count(yo, c.height, R, c.border);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    q[j] = gf_norm_apply(fmaf(c.sA[j], c.ctr[j], c.sB[j]), gf_norm_fast(c.cnt_x[j] * cnt_y));
            }
            float* pq = c.gQ + (int64_t)(yo - c.out_y0) * c.ds;
            if (c.out_lane) {
                if (STEADY || c.vec_ok) {
                    *reinterpret_cast<float4*>(pq) = make_float4(q[0], q[1], q[2], q[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c.x0 + j >= 0 && c.x0 + j < c.width) pq[j] = q[j];
                }
            }
        }
    }
    if (!STEADY && t == steps) return;

    // ================= phase B: stage 1 of row yi =================
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        c.cI[j] += c.nI[j]; c.cP[j] += c.nP[j];
        c.cIP[j] = fmaf(c.nI[j], c.nP[j], c.cIP[j]);
        c.cII[j] = fmaf(c.nI[j], c.nI[j], c.cII[j]);
    }
    if (STEADY || t >= KW) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c.cI[j] -= c.oI[j]; c.cP[j] -= c.oP[j];
            c.cIP[j] = fmaf(-c.oI[j], c.oP[j], c.cIP[j]);
            c.cII[j] = fmaf(-c.oI[j], c.oI[j], c.cII[j]);
        }
    }
    // loads of the next iteration, consumed a full iteration later
    if (STEADY) {
        gf_ld_row<true>(c, c.gI, c.gs, yi + 1 - c.buf_y0, c.nI);
        gf_ld_row<true>(c, c.gP, c.ss, yi + 1 - c.buf_y0, c.nP);
        gf_ld_row<true>(c, c.gI, c.gs, yi + 1 - KW - c.buf_y0, c.oI);
        gf_ld_row<true>(c, c.gP, c.ss, yi + 1 - KW - c.buf_y0, c.oP);
        gf_ld_row<true>(c, c.gI, c.gs, yi - 2 * R - c.buf_y0, c.ctr);
    } else {
        const int sy = gf_map(yi + 1, c.height, c.border);
        gf_ld_row<false>(c, c.gI, c.gs, sy < 0 ? -1 : sy - c.buf_y0, c.nI);
        gf_ld_row<false>(c, c.gP, c.ss, sy < 0 ? -1 : sy - c.buf_y0, c.nP);
        if (t + 1 >= KW) {
            const int so = gf_map(yi + 1 - KW, c.height, c.border);
            gf_ld_row<false>(c, c.gI, c.gs, so < 0 ? -1 : so - c.buf_y0, c.oI);
            gf_ld_row<false>(c, c.gP, c.ss, so < 0 ? -1 : so - c.buf_y0, c.oP);
        }
        if (t >= 4 * R) gf_ld_row<false>(c, c.gI, c.gs, yi - 2 * R - c.buf_y0, c.ctr);
    }
    if (STEADY || t >= 2 * R) {
        // horizontal -> a, b of row yc = yi - R
        const int yc = yi - R;
        float hI[4], hP[4], hIP[4], hII[4];
        gf_window<R>(c.cI, hI, lane);
        gf_window<R>(c.cP, hP, lane);
        gf_window<R>(c.cIP, hIP, lane);
        gf_window<R>(c.cII, hII, lane);
        const bool y_in = STEADY || !c.trunc || (yc >= 0 && yc < c.height);
        if (STEADY) {
            // every window is full, N = (2R+1)^2 exact:
            // a = (N S_Ip - S_I S_p) / (N S_II - S_I^2 + eps N^2),  b = (S_p - a S_I) / N
            const float N = (float)(KW * KW), epsN2 = c.eps * N * N;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float num = fmaf(hIP[j], N, -(hI[j] * hP[j]));
                const float den = fmaf(hII[j], N, fmaf(-hI[j], hI[j], epsN2));
                const float aa = num * gf_rcp(den);
                c.va[j] = c.s1_lane ? aa : 0.f;
                c.vb[j] = c.s1_lane ? gf_norm_apply(fmaf(-aa, hI[j], hP[j]), c.nk) : 0.f;
            }
        } else {
            const float cnt_y = gf_count(yc, c.height, R, c.border);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const GfNorm norm = c.trunc ? gf_norm_fast(c.cnt_x[j] * cnt_y) : c.nk;
                const float mi = gf_norm_apply(hI[j], norm), mp = gf_norm_apply(hP[j], norm);
                const float var = fmaf(-mi, mi, gf_norm_apply(hII[j], norm));
                const float cov = fmaf(-mi, mp, gf_norm_apply(hIP[j], norm));
                const float aa = cov * gf_rcp(var + c.eps);
                const bool ok = c.s1_lane && y_in && c.x_in[j];
                c.va[j] = ok ? aa : 0.f;
                c.vb[j] = ok ? fmaf(-aa, mi, mp) : 0.f;
            }
        }
        if (c.has_ab && c.out_lane && yc >= c.yo0 && yc < c.yo1) {
            float* pa = c.gA + (int64_t)(yc - c.out_y0) * c.abs_;
            float* pb = c.gB + (int64_t)(yc - c.out_y0) * c.abs_;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c.x0 + j < c.width) { pa[j] = c.va[j]; pb[j] = c.vb[j]; }
        }
        if (NW > 1) {
            const int buf = t & 1;
            if (lane >= H1 && lane < 2 * H1) {
                c.xch->a[buf][c.warp][0][lane - H1] = make_float4(c.va[0], c.va[1], c.va[2], c.va[3]);
                c.xch->b[buf][c.warp][0][lane - H1] = make_float4(c.vb[0], c.vb[1], c.vb[2], c.vb[3]);
            }
            if (lane >= 32 - 2 * H1 && lane < 32 - H1) {
                c.xch->a[buf][c.warp][1][lane - (32 - 2 * H1)] = make_float4(c.va[0], c.va[1], c.va[2], c.va[3]);
                c.xch->b[buf][c.warp][1][lane - (32 - 2 * H1)] = make_float4(c.vb[0], c.vb[1], c.vb[2], c.vb[3]);
            }
        }
    }
    if (NW > 1) __syncthreads();
}

template <int R, int NW, bool RING_GLOBAL>
__global__ void __launch_bounds__(NW * 32) gf_fast_gray_kernel(const GfFastArgs a)
{
    using G = GfFastGeom<R>;
    constexpr int H1 = G::H1, OL = G::OL, NT = NW * 32, KW = 2 * R + 1;
    constexpr int WOUT = 4 * (NW * OL - 2 * H1);
    GF_DYN_SMEM(float, smem);
    constexpr size_t xch_floats = sizeof(GfExchange<R, NW>) / 4;
    const int g = threadIdx.x;
    const int64_t f = blockIdx.z;
    GfFastCtx<R, NW> c;
    c.lane = g & 31; c.warp = g >> 5;
    // first of this thread's 4 extended columns: warps overlap by 2*H1 lanes
    c.x0 = (int)blockIdx.x * WOUT - 8 * H1 + c.warp * (4 * OL) + 4 * c.lane;
    c.gI = a.guide + f * a.gfs + c.x0; c.gP = a.src + f * a.sfs + c.x0; c.gQ = a.dst + f * a.dfs + c.x0;
    c.has_ab = a.A != nullptr;
    c.gA = c.has_ab ? a.A + f * a.abfs + c.x0 : nullptr;
    c.gB = c.has_ab ? a.B + f * a.abfs + c.x0 : nullptr;
    c.gs = a.gs; c.ss = a.ss; c.ds = a.ds; c.abs_ = a.abs_;
    c.xch = reinterpret_cast<GfExchange<R, NW>*>(smem);
    // ring of the last 2R+1 rows of (sum_x a, sum_x b): [KW][2][NT] float4, column g private
    c.ring = (RING_GLOBAL
        ? reinterpret_cast<float4*>(a.ring) + ((size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)) * ((size_t)KW * 2 * NT)
        : reinterpret_cast<float4*>(smem + xch_floats)) + g;
    c.width = a.width; c.height = a.height; c.border = a.border; c.buf_y0 = a.buf_y0; c.out_y0 = a.out_y0;
    c.vec_ok = c.x0 >= 0 && c.x0 + 3 < a.width;
    c.yo0 = a.out_y0 + (int)blockIdx.y * a.hb;
    c.yo1 = min(a.out_y0 + a.out_rows, c.yo0 + a.hb);
    c.trunc = a.border == GF_TRUNCATE;
    c.s1_lane = c.lane >= H1 && c.lane < 32 - H1;               // stage-1 windows complete: owns its a, b
    c.out_lane = c.s1_lane && !(c.warp == 0 && c.lane < 2 * H1) && !(c.warp == NW - 1 && c.lane >= 32 - 2 * H1) &&
                 c.x0 < a.width;
    c.eps = a.eps;
    c.nk = gf_norm_make((float)(KW * KW));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        c.cnt_x[j] = gf_count(c.x0 + j, a.width, R, a.border);
        c.x_in[j] = !c.trunc || (c.x0 + j >= 0 && c.x0 + j < a.width);
        c.sx[j] = gf_map(c.x0 + j, a.width, a.border);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        c.cI[j] = c.cP[j] = c.cIP[j] = c.cII[j] = c.sA[j] = c.sB[j] = c.fA[j] = c.fB[j] = c.va[j] = c.vb[j] = c.oI[j] = c.oP[j] = c.ctr[j] = 0.f;
    for (int s = 0; s < KW * 2; ++s) c.ring[(size_t)s * NT] = make_float4(0.f, 0.f, 0.f, 0.f);
    c.slot = 0;

    const int steps = (c.yo1 - c.yo0) + 4 * R;
    {   // row of iteration 0
        const int sy = gf_map(c.yo0 - 2 * R, a.height, a.border);
        gf_ld_row<false>(c, c.gI, c.gs, sy < 0 ? -1 : sy - a.buf_y0, c.nI);
        gf_ld_row<false>(c, c.gP, c.ss, sy < 0 ? -1 : sy - a.buf_y0, c.nP);
    }
    // steady range [ts, te): all stages on, all rows touched interior, y-normalisation constant,
    // and the whole CTA inside the image horizontally (128-bit accesses for every thread)
    const int xl = (int)blockIdx.x * WOUT - 8 * H1, xr = xl + (NW * OL + 2 * H1) * 4;
    const bool cta_inside = xl >= 0 && xr <= a.width;
    int ts = 4 * R + 1;
    if (3 * R + 1 - (c.yo0 - 2 * R) > ts) ts = 3 * R + 1 - (c.yo0 - 2 * R);      // yi >= 3R+1
    if (a.buf_y0 + 2 * R - (c.yo0 - 2 * R) > ts) ts = a.buf_y0 + 2 * R - (c.yo0 - 2 * R);   // yi+1-KW inside the buffer
    int te = min(a.height, a.buf_y0 + a.buf_rows) - 1 - (c.yo0 - 2 * R);          // yi + 1 <= last row present
    if (te > steps) te = steps;
    if (!cta_inside || te < ts) { ts = steps + 1; te = steps + 1; }
    int t = 0;
    for (; t < ts && t <= steps; ++t) gf_fast_iter<false>(c, t, steps);
    for (; t < te; ++t) gf_fast_iter<true>(c, t, steps);
    for (; t <= steps; ++t) gf_fast_iter<false>(c, t, steps);
}

// ---- host side ------------------------------------------------------------------------------------
#if !defined(GF_NO_HOST) && !defined(GF_FAST_NO_TRY)   // (stand-alone SASS builds define GF_NO_HOST; other translation units GF_FAST_NO_TRY)
template <int R, int NW>
struct GfFastLaunch {
    static const char* go(const Job& j, const char** name)
    {
        using G = GfFastGeom<R>;
        constexpr int NT = NW * 32, KW = 2 * R + 1, WOUT = 4 * (NW * G::OL - 2 * G::H1);
        static_assert(WOUT >= 4, "strip has no output columns");
        constexpr size_t mb_bytes = sizeof(GfExchange<R, NW>);
        constexpr size_t ring_bytes = (size_t)KW * 2 * NT * 16;
        const bool ring_global = mb_bytes + ring_bytes > gf_rt_max_smem();
        int sms = 148, mj = 0, mn = 0;
        gf_rt_device_info(&sms, &mj, &mn);
        GfFastArgs a;
        a.guide = j.guide.ptr; a.src = j.src.ptr; a.dst = const_cast<float*>(j.dst.ptr);
        a.A = const_cast<float*>(j.A.ptr); a.B = const_cast<float*>(j.B.ptr);
        a.gs = j.guide.stride; a.ss = j.src.stride; a.ds = j.dst.stride; a.abs_ = j.A.stride;
        a.gfs = j.guide.frame_stride; a.sfs = j.src.frame_stride; a.dfs = j.dst.frame_stride; a.abfs = j.A.frame_stride;
        a.width = j.width; a.height = j.height; a.buf_y0 = j.buf_y0; a.buf_rows = j.buf_rows; a.out_y0 = j.out_y0;
        a.out_rows = j.out_rows; a.border = j.border; a.eps = j.eps; a.ring = nullptr;
        const int nstrips = (j.width + WOUT - 1) / WOUT;
        // band height: enough CTAs to fill the machine, but warm-up (4R rows per band) kept small
        const size_t smem = ring_global ? mb_bytes : mb_bytes + ring_bytes;
        int per_sm = (int)(gf_rt_max_smem() / (smem + 1024));
        const int by_threads = 2048 / NT;
        if (per_sm > by_threads) per_sm = by_threads;
        if (per_sm < 1) per_sm = 1;
        int target = sms * per_sm;
        if (GF_KNOB_SET("GF_FAST_CTAS_PER_SM")) target = sms * GF_KNOB("GF_FAST_CTAS_PER_SM", 1);
        int nb = target / (nstrips * j.count);
        if (nb < 1) nb = 1;
        int hb = (j.out_rows + nb - 1) / nb;
        int hb_min = 6 * R;
        hb_min = GF_KNOB("GF_FAST_HB_MIN", hb_min);
        if (hb < hb_min) hb = hb_min;
        if (hb > j.out_rows) hb = j.out_rows;
        a.hb = hb;
        const int nbands = (j.out_rows + hb - 1) / hb;
        void* ring = nullptr;
        if (ring_global) {
            const size_t n = ring_bytes * nstrips * nbands * j.count;
            if (const char* e = gf_rt_alloc_async(&ring, n, j.stream)) return e;
            a.ring = (float*)ring;
        }
        dim3 grid(nstrips, nbands, j.count), block(NT);
        const char* err = nullptr;
        if (ring_global) {
            auto k = gf_fast_gray_kernel<R, NW, true>;
            err = gf_rt_set_smem(k, smem);
            if (!err) { GF_LAUNCH(k, grid, block, smem, j.stream, a); err = gf_rt_launch_error(); }
        } else {
            auto k = gf_fast_gray_kernel<R, NW, false>;
            err = gf_rt_set_smem(k, smem);
            if (!err) { GF_LAUNCH(k, grid, block, smem, j.stream, a); err = gf_rt_launch_error(); }
        }
        if (ring) gf_rt_free_async(ring, j.stream);
        (void)name;
        return err;
    }
};

// Tries the tuned kernel.  *done=false means "not applicable, use the generic kernel".
// Returns an error string only when a launch was attempted and failed.
static const char* gf_fast_try(const Job& j, bool* done, const char** name)
{
    *done = false;
    if (j.color) return nullptr;
    if (GF_KNOB("GF_DISABLE_FAST", 0)) return nullptr;
    const Plane* pl[3] = {&j.guide, &j.src, &j.dst};
    for (int i = 0; i < 3; ++i)
        if (pl[i]->channels != 1 || pl[i]->coff != 0 || (pl[i]->stride & 3) || (pl[i]->frame_stride & 3) ||
            ((uintptr_t)pl[i]->ptr & 15))
            return nullptr;
    if (j.A.ptr && (j.A.channels != 1 || j.A.coff != 0)) return nullptr;
#define GF_FAST_CASE(RR, NWW, NAME)                      \
    case RR:                                             \
        *done = true;                                    \
        *name = NAME;                                    \
        return GfFastLaunch<RR, NWW>::go(j, name);
    // r <= 16 is served by the warp-private kernel (gf_wp.cuh, 2x faster at r=16: no barrier,
    // smaller ring); this CTA-wide kernel remains for radii whose halo no longer fits one warp.
    switch (j.r) {
        GF_FAST_CASE(20, 4, "fast_r20")
        GF_FAST_CASE(24, 8, "fast_r24")
        GF_FAST_CASE(32, 8, "fast_r32")
    default:
        return nullptr;
    }
#undef GF_FAST_CASE
}
#endif  // GF_NO_HOST
