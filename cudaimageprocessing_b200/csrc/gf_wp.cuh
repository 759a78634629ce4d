// gf_wp.cuh -- warp-private fused guided-filter kernel, gray float32, r <= 16 (the headline path).
//
// Same arithmetic as gf_fast.cuh (K adjacent columns per thread, block prefix/suffix window
// sums over warp shuffles, stage-2 row ring in shared memory, rows loaded one iteration ahead)
// but every WARP is an independent worker: it owns a 32*K-column window of which the middle
// 32-4*H1 lanes (H1 = ceil(R/K) halo lanes per side and stage) produce output, and a band of
// rows.  Nothing is exchanged between warps, so the kernel has no barrier at all, warps drift
// freely and the ring is only 2*(2R+1)*K floats per OUTPUT lane.  The redundant halo lanes are
// L1/L2 hits and cheap FADDs; what they buy back is the barrier stall that dominated the
// exchange version (profiles/).
//
// K = 4 is what this file launches: one float4 per lane and plane, 4 warps per CTA, 12-16 warps per SM
// (first-generation kernel; the fallback for A/B outputs, 16-byte-only alignment and TRUNCATE jobs
// the s8 kernel does not take).  The K-generic window sums below also serve gf_s8.cuh (K = 8) for radii
// that are not multiples of 8.
//
// The row loop is cut into phases with compile-time stage flags so that a warp whose band and
// columns are interior runs straight-line code: no border mapping, no predicates on t, constant
// normalisation.  Border warps (first/last band, first/last strip) run the generic body.
#pragma once
#include "gf_fast.cuh"

// ---- K-generic window sums ----------------------------------------------------------------------
template <int R, int K>
struct GfGeomK {
    static constexpr int H1 = (R + K - 1) / K;         // halo lanes per side per stage
    __host__ __device__ static constexpr int fdiv(int v) { return v >= 0 ? v / K : -((-v + K - 1) / K); }
    __host__ __device__ static constexpr int dL(int j) { return fdiv(j - R); }
    __host__ __device__ static constexpr int oL(int j) { return (j - R) - K * dL(j); }
    __host__ __device__ static constexpr int dR(int j) { return fdiv(j + R); }
    __host__ __device__ static constexpr int oR(int j) { return (j + R) - K * dR(j); }
    static constexpr int dLmin = fdiv(0 - R);
    static constexpr int dRmax = fdiv(K - 1 + R);
};

template <int K>
struct GfBlockK {
    float pre[K];   // pre[o] = c0 + .. + co     (pre[K-1] = total)
    float suf[K];   // suf[o] = co + .. + c(K-1) (suf[0]   = total)
};

template <int K>
__device__ __forceinline__ GfBlockK<K> gf_block_k(const float (&c)[K])
{
    GfBlockK<K> b;
    b.pre[0] = c[0];
#pragma unroll
    for (int o = 1; o < K; ++o) b.pre[o] = b.pre[o - 1] + c[o];
    b.suf[K - 1] = c[K - 1];
#pragma unroll
    for (int o = K - 2; o >= 1; --o) b.suf[o] = c[o] + b.suf[o + 1];
    b.suf[0] = b.pre[K - 1];
    return b;
}

template <int R, int K, int D, int DEND>
struct GfMidLoopK {   // tn[d - dLmin] = total of the block d lanes away, d in [D, DEND)
    __device__ static __forceinline__ void run(float total, float (&tn)[40], int lane)
    {
        tn[D - GfGeomK<R, K>::dLmin] = (D == 0) ? total : gf_fetch<D>(total, lane);
        GfMidLoopK<R, K, D + 1, DEND>::run(total, tn, lane);
    }
};
template <int R, int K, int DEND>
struct GfMidLoopK<R, K, DEND, DEND> {
    __device__ static __forceinline__ void run(float, float (&)[40], int) {}
};

template <int R, int K, int J>
struct GfWindowLoopK {
    __device__ static __forceinline__ void run(const float (&c)[K], const GfBlockK<K>& b, const float (&tn)[40],
                                               float (&out)[K], int lane)
    {
        using G = GfGeomK<R, K>;
        constexpr int dl = G::dL(J), ol = G::oL(J), dr = G::dR(J), orr = G::oR(J);
        if (dl == 0 && dr == 0) {          // window inside the thread's own block
            float s = c[ol];
#pragma unroll
            for (int o = ol + 1; o <= orr; ++o) s += c[o];
            out[J] = s;
        } else {
            const float left = (dl == 0) ? b.suf[ol] : gf_fetch<dl>(b.suf[ol], lane);
            const float right = (dr == 0) ? b.pre[orr] : gf_fetch<dr>(b.pre[orr], lane);
            float s = left;
#pragma unroll
            for (int d = dl + 1; d <= dr - 1; ++d) s += tn[d - G::dLmin];
            out[J] = s + right;
        }
        GfWindowLoopK<R, K, J + 1>::run(c, b, tn, out, lane);
    }
};
template <int R, int K>
struct GfWindowLoopK<R, K, K> {
    __device__ static __forceinline__ void run(const float (&)[K], const GfBlockK<K>&, const float (&)[40], float (&)[K], int) {}
};

// (2R+1)-window sums of the K columns of every lane; complete for lanes [H1, 32-H1).
template <int R, int K>
__device__ __forceinline__ void gf_window_k(const float (&c)[K], float (&out)[K], int lane)
{
    using G = GfGeomK<R, K>;
    const GfBlockK<K> b = gf_block_k<K>(c);
    float tn[40];
    GfMidLoopK<R, K, G::dLmin + 1, (G::dRmax - 1 >= G::dLmin + 1 ? G::dRmax : G::dLmin + 1)>::run(b.pre[K - 1], tn, lane);
    GfWindowLoopK<R, K, 0>::run(c, b, tn, out, lane);
}

// ---- the kernel ------------------------------------------------------------------------------------
template <int R, int K>
struct GfWpGeom {
    static constexpr int H1 = GfGeomK<R, K>::H1;
    static constexpr int VL = 32 - 4 * H1;          // lanes that produce output
    static constexpr int WOUT = K * VL;             // output columns per warp
    static constexpr int WIN = 32 * K;              // columns a warp loads
    static constexpr int KW = 2 * R + 1;
    static constexpr int V4 = K / 4;                // float4 per lane and plane
    static constexpr int RING_CELLS = VL + 1;       // + one dump cell shared by the halo lanes
    static constexpr int WARPS = K == 4 ? 4 : 1;    // warps per CTA
    static constexpr size_t ring_bytes_per_warp = (size_t)KW * 2 * V4 * RING_CELLS * 16;
};

struct GfWpArgs {
    const float* guide; const float* src; float* dst; float* A; float* B;
    int64_t gs, ss, ds, abs_;
    int64_t gfs, sfs, dfs, abfs;
    int width, height, buf_y0, buf_rows, out_y0, out_rows, border, hb;
    int nstrips, nbands, count;
    int hb_e, nbands_e;      // gf_s8 only: band height / band count of the first and last strip (0 = same as the others)
    float eps;
    // Tape scheduling (gf_s8 / gf_c4, tape_piece > 0; see gf_tape_run): one CTA per piece of the tape
    long long tape_piece;    // piece length in cost units (a row of an interior strip = 100 units)
    int tape_rho, tape_we;   // ramp of a band in rows; cost of a row of the first / last strip (>= 100)
};

// Tape scheduling (experiment, off by default: see gf_tape_plan for the measured outcome).  The uniform split (nbands bands of hb rows per strip) leaves the machine badly
// filled whenever strips x bands is not just below a multiple of the resident warps (32 frames of
// 1080p colour, r = 16: 960 strips on 888 warp slots -> 3 waves of half-height bands, 33 % ramp).
// Instead all (frame, strip) columns are laid end to end on a tape, every column preceded by the
// cost of one ramp (rho rows: the 4R warm-up iterations of a band cost about 2.3R + 1 row times),
// rows of the first and last strip weighted by their slower code, and the tape is cut into equal
// pieces, one per resident warp.  A warp walks its piece: every column it touches is one band
// (its own ramp + its rows), so every warp does the same work to within one row and the whole job
// is ONE wave:   time ~ tape / slots + ramp   instead of   waves x (hb + ramp).
// body(frame, strip, first output row, end output row) runs one band.
// With tape_piece == 0 the item is one band of the uniform split: per frame the first and the last
// strip come first in nbands_e bands of hb_e rows (nbands_e > 0, nstrips >= 3: gf_s8's shorter bands
// for the slower edge code), the other strips follow in nbands bands of hb rows.
template <class F>
__device__ __forceinline__ void gf_tape_run(const GfWpArgs& a, long long item, F&& body)
{
    const bool tape = a.tape_piece > 0;
    const int rho = a.tape_rho, we = a.tape_we, ns = a.nstrips;
    const bool edges = ns >= 3 && we != 100;
    const long long zi = (long long)(rho + a.out_rows) * 100, ze = (long long)(rho + a.out_rows) * we;
    const long long frame = edges ? 2 * ze + (ns - 2) * zi : ns * zi;
    long long pos = tape ? item * a.tape_piece : 0;
    long long end = tape ? pos + a.tape_piece : 1;
    if (tape && end > frame * a.count) end = frame * a.count;
    while (pos < end) {
        int64_t f;
        int strip, r0, r1;
        if (tape) {
            const long long fb = (pos / frame) * frame, rem = pos - fb;
            f = pos / frame;
            int w;
            long long z0;
            if (!edges) { strip = (int)(rem / zi); z0 = strip * zi; w = 100; }
            else if (rem < ze) { strip = 0; z0 = 0; w = we; }
            else if (rem < ze + (ns - 2) * zi) { const int k = (int)((rem - ze) / zi); strip = 1 + k; z0 = ze + k * zi; w = 100; }
            else { strip = ns - 1; z0 = ze + (ns - 2) * zi; w = we; }
            const long long zlen = (long long)(rho + a.out_rows) * w, lead = (long long)rho * w;
            const long long u0 = rem - z0, u1 = end - fb - z0 < zlen ? end - fb - z0 : zlen;
            // first row whose cost interval starts at or after u: pieces tile every column exactly
            r0 = u0 <= lead ? 0 : (int)((u0 - lead + w - 1) / w);
            r1 = u1 <= lead ? 0 : (int)((u1 - lead + w - 1) / w);
            pos = fb + z0 + zlen;
        } else {
            const bool two = a.nbands_e > 0 && ns >= 3;
            const long n_edge = two ? 2L * a.nbands_e : 0, n_int = two ? (long)(ns - 2) * a.nbands : (long)ns * a.nbands;
            const long per_frame = n_edge + n_int, idx = (long)(item % per_frame);
            f = item / per_frame;
            int band, hbw;
            if (idx < n_edge) { band = (int)(idx >> 1); strip = (idx & 1) ? ns - 1 : 0; hbw = a.hb_e; }
            else if (two) { const long k = idx - n_edge; band = (int)(k / (ns - 2)); strip = 1 + (int)(k % (ns - 2)); hbw = a.hb; }
            else { band = (int)(idx / ns); strip = (int)(idx % ns); hbw = a.hb; }
            r0 = band * hbw;
            r1 = r0 + hbw < a.out_rows ? r0 + hbw : a.out_rows;
            pos = end;
        }
        if (r1 > r0) body(f, strip, a.out_y0 + r0, a.out_y0 + r1);
    }
}

template <int R, int K>
struct GfWpCtx {
    const float* gI; const float* gP; float* gQ; float* gA; float* gB;   // frame bases at column x0
    int64_t gs, ss, ds, abs_;
    float4* ring;            // this lane's ring cells: ring[((slot*2+q)*V4 + v)*RING_CELLS]
    int lane, x0, width, height, border, buf_y0, out_y0, yo0, yo1;
    bool vec_ok, trunc, s1_lane, out_lane, has_ab;
    float eps;
    GfNorm nk;               // 1 / (2R+1)^2
    float cnt_x[K];
    bool x_in[K];
    int sx[K];
    float cI[K], cP[K], cIP[K], cII[K], sA[K], sB[K], va[K], vb[K];
    float nI[K], nP[K], oI[K], oP[K];   // oI doubles as the guide row of the next output (yi+1-KW == yi-2R)
    int slot;
};

template <int K>
__device__ __forceinline__ void gf_wp_ldv(const float* p, float (&v)[K])
{
#pragma unroll
    for (int i = 0; i < K / 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

template <int R, int K>
__device__ __forceinline__ void gf_wp_ld(const GfWpCtx<R, K>& c, const float* base, int64_t stride, int row, float (&v)[K])
{
    if (row < 0) {
#pragma unroll
        for (int j = 0; j < K; ++j) v[j] = 0.f;
        return;
    }
    const float* p = base + (int64_t)row * stride;
    if (c.vec_ok) {
        gf_wp_ldv<K>(p, v);
    } else {
#pragma unroll
        for (int j = 0; j < K; ++j) v[j] = c.sx[j] >= 0 ? p[c.sx[j] - c.x0] : 0.f;
    }
}

// Iteration t.  PH selects which stages are compiled in:
//   0  t in [0, 2R)        vertical add only
//   2  t in (2R, 4R)       stage 2 without output + stage 1 with subtraction
//   3  t in (4R, steps)    everything (steady state)
//   5  any t               generic: run-time stage flags, border mapping, masks
// PH 0/2/3 require an INTERIOR warp: all rows and columns it touches are inside the image.
template <int PH, int R, int K>
__device__ __forceinline__ void gf_wp_iter(GfWpCtx<R, K>& c, int t, int steps)
{
    using W = GfWpGeom<R, K>;
    constexpr int KW = W::KW, RC = W::RING_CELLS, V4 = W::V4;
    constexpr bool GEN = PH == 5;
    const int yi = c.yo0 - 2 * R + t;
    const int lane = c.lane;
    const bool a_on = GEN ? (t - 1 >= 2 * R) : (PH >= 2);
    const bool out_on = GEN ? (t - 1 >= 4 * R) : (PH == 3);
    const bool sub_on = GEN ? (t >= KW) : (PH >= 2);
    const bool s1_on = GEN ? (t >= 2 * R) : (PH >= 2);

    // ================= phase A: stage 2 of centre row yi-1-R =================
    if (a_on) {
        float hA[K], hB[K];
        gf_window_k<R, K>(c.va, hA, lane);
        gf_window_k<R, K>(c.vb, hB, lane);
#pragma unroll
        for (int v = 0; v < V4; ++v) {
            float4* ca = c.ring + (size_t)((c.slot * 2 + 0) * V4 + v) * RC;
            float4* cb = c.ring + (size_t)((c.slot * 2 + 1) * V4 + v) * RC;
            const float4 oa = *ca, ob = *cb;
            c.sA[4 * v] += hA[4 * v] - oa.x; c.sA[4 * v + 1] += hA[4 * v + 1] - oa.y;
            c.sA[4 * v + 2] += hA[4 * v + 2] - oa.z; c.sA[4 * v + 3] += hA[4 * v + 3] - oa.w;
            c.sB[4 * v] += hB[4 * v] - ob.x; c.sB[4 * v + 1] += hB[4 * v + 1] - ob.y;
            c.sB[4 * v + 2] += hB[4 * v + 2] - ob.z; c.sB[4 * v + 3] += hB[4 * v + 3] - ob.w;
            *ca = make_float4(hA[4 * v], hA[4 * v + 1], hA[4 * v + 2], hA[4 * v + 3]);
            *cb = make_float4(hB[4 * v], hB[4 * v + 1], hB[4 * v + 2], hB[4 * v + 3]);
        }
        c.slot = c.slot + 1 == KW ? 0 : c.slot + 1;
        if (out_on) {                               // q of row yo = yi-1-2R; its guide row is oI
            const int yo = yi - 1 - 2 * R;
            float q[K];
            if (!GEN || !c.trunc) {
#pragma unroll
                for (int j = 0; j < K; ++j) q[j] = gf_norm_apply(fmaf(c.sA[j], c.oI[j], c.sB[j]), c.nk);
            } else {
                const float cnt_y = gf_count(yo, c.height, R, c.border);
#pragma unroll
                for (int j = 0; j < K; ++j)
                    q[j] = gf_norm_apply(fmaf(c.sA[j], c.oI[j], c.sB[j]), gf_norm_fast(c.cnt_x[j] * cnt_y));
            }
            float* pq = c.gQ + (int64_t)(yo - c.out_y0) * c.ds;
            if (c.out_lane) {
                if (!GEN || c.vec_ok) {
#pragma unroll
                    for (int v = 0; v < V4; ++v)
                        reinterpret_cast<float4*>(pq)[v] = make_float4(q[4 * v], q[4 * v + 1], q[4 * v + 2], q[4 * v + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j)
                        if (c.x0 + j >= 0 && c.x0 + j < c.width) pq[j] = q[j];
                }
            }
        }
    }
    if (GEN && t == steps) return;

    // ================= phase B: stage 1 of row yi =================
#pragma unroll
    for (int j = 0; j < K; ++j) {
        c.cI[j] += c.nI[j]; c.cP[j] += c.nP[j];
        c.cIP[j] = fmaf(c.nI[j], c.nP[j], c.cIP[j]);
        c.cII[j] = fmaf(c.nI[j], c.nI[j], c.cII[j]);
    }
    if (sub_on) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            c.cI[j] -= c.oI[j]; c.cP[j] -= c.oP[j];
            c.cIP[j] = fmaf(-c.oI[j], c.oP[j], c.cIP[j]);
            c.cII[j] = fmaf(-c.oI[j], c.oI[j], c.cII[j]);
        }
    }
    // loads of the next iteration, consumed a full iteration later
    if (!GEN) {
        const int64_t rn = (int64_t)(yi + 1 - c.buf_y0);
        gf_wp_ldv<K>(c.gI + rn * c.gs, c.nI);
        gf_wp_ldv<K>(c.gP + rn * c.ss, c.nP);
        if (PH >= 2 || t + 1 >= KW) {
            gf_wp_ldv<K>(c.gI + (rn - KW) * c.gs, c.oI);
            gf_wp_ldv<K>(c.gP + (rn - KW) * c.ss, c.oP);
        }
    } else {
        const int sy = gf_map(yi + 1, c.height, c.border);
        gf_wp_ld(c, c.gI, c.gs, sy < 0 ? -1 : sy - c.buf_y0, c.nI);
        gf_wp_ld(c, c.gP, c.ss, sy < 0 ? -1 : sy - c.buf_y0, c.nP);
        if (t + 1 >= KW) {
            const int so = gf_map(yi + 1 - KW, c.height, c.border);
            gf_wp_ld(c, c.gI, c.gs, so < 0 ? -1 : so - c.buf_y0, c.oI);
            gf_wp_ld(c, c.gP, c.ss, so < 0 ? -1 : so - c.buf_y0, c.oP);
        }
    }
    if (s1_on) {
        // horizontal -> a, b of row yc = yi - R
        float hI[K], hP[K], hIP[K], hII[K];
        gf_window_k<R, K>(c.cI, hI, lane);
        gf_window_k<R, K>(c.cP, hP, lane);
        gf_window_k<R, K>(c.cIP, hIP, lane);
        gf_window_k<R, K>(c.cII, hII, lane);
        if (!GEN) {
            // interior: every window is full, N = (2R+1)^2 exact:
            // a = (N S_Ip - S_I S_p) / (N S_II - S_I^2 + eps N^2),  b = (S_p - a S_I) / N
            const float N = (float)(KW * KW), epsN2 = c.eps * N * N;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float num = fmaf(hIP[j], N, -(hI[j] * hP[j]));
                const float den = fmaf(hII[j], N, fmaf(-hI[j], hI[j], epsN2));
                const float aa = num * gf_rcp(den);
                c.va[j] = aa;
                c.vb[j] = gf_norm_apply(fmaf(-aa, hI[j], hP[j]), c.nk);
            }
        } else {
            const int yc = yi - R;
            const bool y_in = !c.trunc || (yc >= 0 && yc < c.height);
            const float cnt_y = gf_count(yc, c.height, R, c.border);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const GfNorm norm = c.trunc ? gf_norm_fast(c.cnt_x[j] * cnt_y) : c.nk;
                const float mi = gf_norm_apply(hI[j], norm), mp = gf_norm_apply(hP[j], norm);
                const float var = fmaf(-mi, mi, gf_norm_apply(hII[j], norm));
                const float cov = fmaf(-mi, mp, gf_norm_apply(hIP[j], norm));
                const float aa = cov * gf_rcp(var + c.eps);
                const bool ok = c.s1_lane && y_in && c.x_in[j];
                c.va[j] = ok ? aa : 0.f;
                c.vb[j] = ok ? fmaf(-aa, mi, mp) : 0.f;
            }
            if (c.has_ab && c.out_lane && yc >= c.yo0 && yc < c.yo1) {
                float* pa = c.gA + (int64_t)(yc - c.out_y0) * c.abs_;
                float* pb = c.gB + (int64_t)(yc - c.out_y0) * c.abs_;
#pragma unroll
                for (int j = 0; j < K; ++j)
                    if (c.x0 + j < c.width) { pa[j] = c.va[j]; pb[j] = c.vb[j]; }
            }
        }
    }
}

// MINB = resident CTAs per SM the register allocation is sized for.
template <int R, int K, int MINB>
__global__ void __launch_bounds__(GfWpGeom<R, K>::WARPS * 32, MINB) gf_wp_gray_kernel(const GfWpArgs a)
{
    using W = GfWpGeom<R, K>;
    constexpr int H1 = W::H1, KW = W::KW, RC = W::RING_CELLS, V4 = W::V4;
    GF_DYN_SMEM(float, smem);
    const int warp = threadIdx.x >> 5;
    // work item of this warp: (strip, band, frame)
    const long item = (long)blockIdx.x * W::WARPS + warp;
    const long per_frame = (long)a.nstrips * a.nbands;
    if (item >= per_frame * a.count) return;
    const int64_t f = item / per_frame;
    const int band = (int)((item % per_frame) / a.nstrips), strip = (int)(item % a.nstrips);

    GfWpCtx<R, K> c;
    c.lane = threadIdx.x & 31;
    c.x0 = strip * W::WOUT - 2 * H1 * K + K * c.lane;
    c.gI = a.guide + f * a.gfs + c.x0; c.gP = a.src + f * a.sfs + c.x0; c.gQ = a.dst + f * a.dfs + c.x0;
    c.has_ab = a.A != nullptr;
    c.gA = c.has_ab ? a.A + f * a.abfs + c.x0 : nullptr;
    c.gB = c.has_ab ? a.B + f * a.abfs + c.x0 : nullptr;
    c.gs = a.gs; c.ss = a.ss; c.ds = a.ds; c.abs_ = a.abs_;
    c.out_lane = c.lane >= 2 * H1 && c.lane < 32 - 2 * H1 && c.x0 < a.width;
    c.s1_lane = c.lane >= H1 && c.lane < 32 - H1;
    {
        const bool ring_lane = c.lane >= 2 * H1 && c.lane < 32 - 2 * H1;
        float4* base = reinterpret_cast<float4*>(smem) + (size_t)warp * (W::ring_bytes_per_warp / 16);
        c.ring = base + (ring_lane ? c.lane - 2 * H1 : W::VL);
        for (int s = 0; s < KW * 2 * V4; ++s) c.ring[(size_t)s * RC] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    c.width = a.width; c.height = a.height; c.border = a.border; c.buf_y0 = a.buf_y0; c.out_y0 = a.out_y0;
    c.vec_ok = c.x0 >= 0 && c.x0 + K - 1 < a.width;
    c.yo0 = a.out_y0 + band * a.hb;
    c.yo1 = min(a.out_y0 + a.out_rows, c.yo0 + a.hb);
    c.trunc = a.border == GF_TRUNCATE;
    c.eps = a.eps;
    c.nk = gf_norm_make((float)(KW * KW));
#pragma unroll
    for (int j = 0; j < K; ++j) {
        c.cnt_x[j] = gf_count(c.x0 + j, a.width, R, a.border);
        c.x_in[j] = !c.trunc || (c.x0 + j >= 0 && c.x0 + j < a.width);
        c.sx[j] = gf_map(c.x0 + j, a.width, a.border);
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
        c.cI[j] = c.cP[j] = c.cIP[j] = c.cII[j] = c.sA[j] = c.sB[j] = c.va[j] = c.vb[j] = c.oI[j] = c.oP[j] = 0.f;
    c.slot = 0;
    __syncwarp();

    const int steps = (c.yo1 - c.yo0) + 4 * R;
    {   // row of iteration 0
        const int sy = gf_map(c.yo0 - 2 * R, a.height, a.border);
        gf_wp_ld(c, c.gI, c.gs, sy < 0 ? -1 : sy - a.buf_y0, c.nI);
        gf_wp_ld(c, c.gP, c.ss, sy < 0 ? -1 : sy - a.buf_y0, c.nP);
    }
    // interior warp: its 32*K columns and all rows [yo0-2R, yo1+2R] lie inside the image and the
    // buffer (the last iteration loads one row beyond the band's halo, which must exist)
    const int xl = strip * W::WOUT - 2 * H1 * K;
    const int y_end = min(a.height, a.buf_y0 + a.buf_rows);
    const bool interior = xl >= 0 && xl + W::WIN <= a.width && c.yo0 - 2 * R >= max(0, a.buf_y0) && c.yo1 + 2 * R < y_end &&
                          !c.has_ab;
    int t = 0;
    if (interior) {
        for (; t < 2 * R; ++t) gf_wp_iter<0>(c, t, steps);
        gf_wp_iter<5>(c, t, steps); ++t;                 // t = 2R: first stage 1, nothing to subtract yet
        for (; t < 4 * R; ++t) gf_wp_iter<2>(c, t, steps);
        gf_wp_iter<5>(c, t, steps); ++t;                 // t = 4R: stage 2 still without output
        for (; t < steps; ++t) gf_wp_iter<3>(c, t, steps);
        gf_wp_iter<5>(c, t, steps);                      // t = steps: last output row
    } else {
        for (; t <= steps; ++t) gf_wp_iter<5>(c, t, steps);
    }
}

// ---- host side ------------------------------------------------------------------------------------
#if !defined(GF_NO_HOST) && !defined(GF_WP_NO_TRY)   // (stand-alone SASS builds define GF_NO_HOST; other translation units GF_WP_NO_TRY)
template <int R, int K>
static const char* gf_wp_launch(const Job& j)
{
    using W = GfWpGeom<R, K>;
    static_assert(W::VL >= 4, "no output lanes left");
    int sms = 148, mj = 0, mn = 0;
    gf_rt_device_info(&sms, &mj, &mn);
    GfWpArgs a;
    a.guide = j.guide.ptr; a.src = j.src.ptr; a.dst = const_cast<float*>(j.dst.ptr);
    a.A = const_cast<float*>(j.A.ptr); a.B = const_cast<float*>(j.B.ptr);
    a.gs = j.guide.stride; a.ss = j.src.stride; a.ds = j.dst.stride; a.abs_ = j.A.stride;
    a.gfs = j.guide.frame_stride; a.sfs = j.src.frame_stride; a.dfs = j.dst.frame_stride; a.abfs = j.A.frame_stride;
    a.width = j.width; a.height = j.height; a.buf_y0 = j.buf_y0; a.buf_rows = j.buf_rows; a.out_y0 = j.out_y0;
    a.out_rows = j.out_rows; a.border = j.border; a.eps = j.eps; a.count = j.count;
    a.hb_e = 0; a.nbands_e = 0; a.tape_piece = 0; a.tape_rho = 0; a.tape_we = 100;
    a.nstrips = (j.width + W::WOUT - 1) / W::WOUT;
    const size_t smem = W::WARPS * W::ring_bytes_per_warp;
    // resident warps per SM: shared memory (ring) and registers (<= 168 per thread at 12 warps)
    int warps_sm = (int)(gf_rt_max_smem() / (smem + 1024)) * W::WARPS;
    const int reg_cap = K == 4 ? 12 : 8;
    if (warps_sm > reg_cap) warps_sm = reg_cap;
    if (warps_sm < 1) warps_sm = 1;
    // the 128-register build of the K=4 kernel holds 16 warps: better when the job spans many waves
    const long min_items = (long)a.nstrips * ((j.out_rows + 255) / 256) * j.count;
    bool big = K == 4 && R <= 8 && min_items > (long)sms * 12;
    if (GF_KNOB_SET("GF_WP_BIG")) big = K == 4 && R <= 8 && GF_KNOB("GF_WP_BIG", 0) != 0;
    int warps_target = sms * (big ? 16 : warps_sm);
    if (GF_KNOB_SET("GF_WP_WARPS_PER_SM")) warps_target = sms * GF_KNOB("GF_WP_WARPS_PER_SM", 1);
    // Bands: as many as fit in ONE wave of resident warps (a partial second wave costs more than
    // its share), but never so short that the 4R warm-up rows dominate; large jobs get many
    // waves of hb_max-row bands instead.
    int nb = warps_target / (a.nstrips * j.count);
    if (nb < 1) nb = 1;
    int hb = (j.out_rows + nb - 1) / nb;
    int hb_min = 4 * R, hb_max = 256;
    hb_min = GF_KNOB("GF_WP_HB_MIN", hb_min);
    if (hb < hb_min) hb = hb_min;
    if (hb > hb_max) hb = hb_max;
    if (hb > j.out_rows) hb = j.out_rows;
    a.hb = hb;
    a.nbands = (j.out_rows + hb - 1) / hb;
    const long items = (long)a.nstrips * a.nbands * j.count;
    dim3 grid((unsigned)((items + W::WARPS - 1) / W::WARPS)), block(W::WARPS * 32);
    if (big) {
        auto k = gf_wp_gray_kernel<R, K, (K == 4 && R <= 8 ? 4 : 3)>;
        if (const char* e = gf_rt_set_smem(k, smem)) return e;
        GF_LAUNCH(k, grid, block, smem, j.stream, a);
    } else {
        auto k = gf_wp_gray_kernel<R, K, (K == 4 ? 3 : 7)>;
        if (const char* e = gf_rt_set_smem(k, smem)) return e;
        GF_LAUNCH(k, grid, block, smem, j.stream, a);
    }
    return gf_rt_launch_error();
}

static const char* gf_wp_try(const Job& j, bool* done, const char** name)
{
    *done = false;
    if (j.color || j.r < 1 || j.r > 16) return nullptr;
    if (GF_KNOB("GF_DISABLE_WP", 0) || GF_KNOB("GF_DISABLE_FAST", 0)) return nullptr;
    const Plane* pl[3] = {&j.guide, &j.src, &j.dst};
    for (int i = 0; i < 3; ++i)
        if (pl[i]->channels != 1 || pl[i]->coff != 0 || (pl[i]->stride & 3) || (pl[i]->frame_stride & 3) ||
            ((uintptr_t)pl[i]->ptr & 15))
            return nullptr;
    if (j.A.ptr && (j.A.channels != 1 || j.A.coff != 0)) return nullptr;
    *done = true;
    switch (j.r) {
    case 1: *name = "wp_r1"; return gf_wp_launch<1, 4>(j);
    case 2: *name = "wp_r2"; return gf_wp_launch<2, 4>(j);
    case 3: *name = "wp_r3"; return gf_wp_launch<3, 4>(j);
    case 4: *name = "wp_r4"; return gf_wp_launch<4, 4>(j);
    case 5: *name = "wp_r5"; return gf_wp_launch<5, 4>(j);
    case 6: *name = "wp_r6"; return gf_wp_launch<6, 4>(j);
    case 7: *name = "wp_r7"; return gf_wp_launch<7, 4>(j);
    case 8: *name = "wp_r8"; return gf_wp_launch<8, 4>(j);
    case 9: *name = "wp_r9"; return gf_wp_launch<9, 4>(j);
    case 10: *name = "wp_r10"; return gf_wp_launch<10, 4>(j);
    case 11: *name = "wp_r11"; return gf_wp_launch<11, 4>(j);
    case 12: *name = "wp_r12"; return gf_wp_launch<12, 4>(j);
    case 13: *name = "wp_r13"; return gf_wp_launch<13, 4>(j);
    case 14: *name = "wp_r14"; return gf_wp_launch<14, 4>(j);
    case 15: *name = "wp_r15"; return gf_wp_launch<15, 4>(j);
    default: *name = "wp_r16"; return gf_wp_launch<16, 4>(j);
    }
}
#endif  // GF_NO_HOST
