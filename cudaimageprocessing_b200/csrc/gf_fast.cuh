// gf_fast.cuh -- the tuned fused guided-filter kernel for gray float32 planes (sm_100a).
//
// Same dataflow as gf_generic_kernel (one pass down a column strip, a/b never leave the SM)
// but built around what bounds it on B200 (bench_tools/microbench.cu): shared-memory LDS/STS
// and warp shuffles share one ~1 warp-instruction/clk/SM pipe, so the kernel keeps everything it
// can in registers and moves data between lanes as rarely as possible:
//   * each thread owns K=4 ADJACENT columns: 128-bit loads/stores, four independent
//     running column sums per quantity in registers;
//   * horizontal (2R+1)-window sums with compile-time R are assembled from per-thread block
//     prefixes/suffixes/totals of the 4 columns: window(4l+j) = suf_{l+dL}[oL] + totals in
//     between + pre_{l+dR}[oR].  Only additions (float32 error stays ~1e-7 relative, the
//     reference's fused path quality) and 8 + (block span - 2) shuffles per quantity per
//     128 pixels instead of a 5-step scan per pixel;
//   * warps of a CTA share their strip: lanes at a warp edge take their neighbour warp's block
//     sums from a small shared-memory mailbox (two barriers per row);
//   * the row that leaves the vertical window is re-read from global memory (it is an L1/L2
//     hit 2R+1 rows later) instead of being parked in shared memory; only the stage-2 ring
//     (mean-of-a/b rows) lives in shared memory, as float4 per thread.
// HBM traffic stays at read I, p once + write q once; the halo columns/rows between strips and
// bands are L2 hits.
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
    static constexpr int NSLOT = 6;                   // mailbox slots: 4 stage-1 + 2 stage-2 quantities
    __host__ __device__ static constexpr int floordiv4(int v) { return v >= 0 ? v / 4 : -((-v + 3) / 4); }
    __host__ __device__ static constexpr int dL(int j) { return floordiv4(j - R); }
    __host__ __device__ static constexpr int oL(int j) { return (j - R) - 4 * dL(j); }
    __host__ __device__ static constexpr int dR(int j) { return floordiv4(j + R); }
    __host__ __device__ static constexpr int oR(int j) { return (j + R) - 4 * dR(j); }
    static constexpr int dLmin = floordiv4(0 - R);
    static constexpr int dRmax = floordiv4(3 + R);
};

// Mailbox: per quantity slot, per warp, per edge lane h: {v0, v1, v2, v3, total}.
//   left  box (written by lanes h < H1)        : prefix sums, read by the warp to the LEFT
//   right box (written by lanes 32-H1+h)       : suffix sums, read by the warp to the RIGHT
template <int R, int NW>
struct GfMailbox {
    static constexpr int H1 = GfFastGeom<R>::H1;
    float left[GfFastGeom<R>::NSLOT][NW][H1][5];
    float right[GfFastGeom<R>::NSLOT][NW][H1][5];
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

template <int R, int NW>
__device__ __forceinline__ void gf_publish(GfMailbox<R, NW>* mb, int slot, const GfBlock& b, int lane, int warp)
{
    constexpr int H1 = GfFastGeom<R>::H1;
    if (NW == 1) return;
    if (lane < H1) {
        float* d = mb->left[slot][warp][lane];
        d[0] = b.pre[0]; d[1] = b.pre[1]; d[2] = b.pre[2]; d[3] = b.pre[3]; d[4] = b.pre[3];
    }
    if (lane >= 32 - H1) {
        float* d = mb->right[slot][warp][lane - (32 - H1)];
        d[0] = b.suf[0]; d[1] = b.suf[1]; d[2] = b.suf[2]; d[3] = b.suf[3]; d[4] = b.pre[3];
    }
}

// value `x` (register index `idx` of the block: 0..3, or 4 = total) of the thread D lanes away
template <int R, int NW, int D>
__device__ __forceinline__ float gf_fetch(float x, int idx, const GfMailbox<R, NW>* mb, int slot, int lane, int warp)
{
    constexpr int H1 = GfFastGeom<R>::H1;
    float v = __shfl_sync(0xffffffffu, x, (lane + D) & 31);
    if (D < 0) {
        if (lane + D < 0) v = (NW > 1 && warp > 0) ? mb->right[slot][warp - 1][lane + D + H1][idx] : 0.f;
    } else if (D > 0) {
        if (lane + D >= 32) v = (NW > 1 && warp < NW - 1) ? mb->left[slot][warp + 1][lane + D - 32][idx] : 0.f;
    }
    return v;
}

template <int R, int NW, int D, int DEND>
struct GfMidLoop {
    // T_{l+d} for d in [D, DEND], own total for d == 0
    __device__ static __forceinline__ void run(float total, float (&tn)[32], const GfMailbox<R, NW>* mb, int slot, int lane, int warp)
    {
        constexpr int base = GfFastGeom<R>::dLmin;
        tn[D - base] = (D == 0) ? total : gf_fetch<R, NW, D>(total, 4, mb, slot, lane, warp);
        GfMidLoop<R, NW, D + 1, DEND>::run(total, tn, mb, slot, lane, warp);
    }
};
template <int R, int NW, int DEND>
struct GfMidLoop<R, NW, DEND, DEND> {
    __device__ static __forceinline__ void run(float, float (&)[32], const GfMailbox<R, NW>*, int, int, int) {}
};

template <int R, int NW, int J>
__device__ __forceinline__ float gf_window_j(const float (&c)[4], const GfBlock& b, const float (&tn)[32],
                                             const GfMailbox<R, NW>* mb, int slot, int lane, int warp)
{
    using G = GfFastGeom<R>;
    constexpr int dl = G::dL(J), ol = G::oL(J), dr = G::dR(J), orr = G::oR(J);
    if (dl == 0 && dr == 0) {             // window inside the thread's own block (R <= 1)
        float s = c[ol];
#pragma unroll
        for (int o = ol + 1; o <= orr; ++o) s += c[o];
        return s;
    }
    const float left = (dl == 0) ? b.suf[ol] : gf_fetch<R, NW, dl>(b.suf[ol], ol, mb, slot, lane, warp);
    const float right = (dr == 0) ? b.pre[orr] : gf_fetch<R, NW, dr>(b.pre[orr], orr, mb, slot, lane, warp);
    float s = left;
#pragma unroll
    for (int d = dl + 1; d <= dr - 1; ++d) s += tn[d - G::dLmin];
    return s + right;
}

// (2R+1)-window sums of the 4 columns of every thread.  The block sums must have been published
// (gf_publish) and a barrier passed before this is called.
template <int R, int NW>
__device__ __forceinline__ void gf_window(const float (&c)[4], const GfBlock& b, float (&out)[4],
                                          const GfMailbox<R, NW>* mb, int slot, int lane, int warp)
{
    using G = GfFastGeom<R>;
    float tn[32];
    // totals of the blocks strictly between the end blocks of any of the 4 windows
    GfMidLoop<R, NW, G::dLmin + 1, (G::dRmax - 1 >= G::dLmin + 1 ? G::dRmax : G::dLmin + 1)>::run(b.pre[3], tn, mb, slot, lane, warp);
    out[0] = gf_window_j<R, NW, 0>(c, b, tn, mb, slot, lane, warp);
    out[1] = gf_window_j<R, NW, 1>(c, b, tn, mb, slot, lane, warp);
    out[2] = gf_window_j<R, NW, 2>(c, b, tn, mb, slot, lane, warp);
    out[3] = gf_window_j<R, NW, 3>(c, b, tn, mb, slot, lane, warp);
}

struct GfFastArgs {
    const float* guide; const float* src; float* dst; float* A; float* B;
    int64_t gs, ss, ds, abs_;            // row strides (floats)
    int64_t gfs, sfs, dfs, abfs;         // frame strides (floats)
    int width, height, buf_y0, out_y0, out_rows, border, hb;
    float eps;
    float* ring;                         // global ring scratch (RING_GLOBAL) or nullptr
};

// 4 adjacent pixels of extended row `yi` starting at extended column x0 (multiple of 4).
__device__ __forceinline__ void gf_load4(const float* __restrict__ base, int64_t stride, int row, int x0, int width,
                                         int border, bool vec_ok, float (&v)[4])
{
    if (row < 0) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
    const float* p = base + (int64_t)row * stride;
    if (vec_ok) {
        const float4 t = *reinterpret_cast<const float4*>(p + x0);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sx = gf_map(x0 + j, width, border);
            v[j] = sx >= 0 ? p[sx] : 0.f;
        }
    }
}

template <int R, int NW, bool RING_GLOBAL>
__global__ void __launch_bounds__(NW * 32) gf_fast_gray_kernel(const GfFastArgs a)
{
    using G = GfFastGeom<R>;
    constexpr int H1 = G::H1, NT = NW * 32, KW = 2 * R + 1;
    constexpr int WOUT = NT * 4 - 16 * H1;
    GF_DYN_SMEM(float, smem);
    GfMailbox<R, NW>* mb = reinterpret_cast<GfMailbox<R, NW>*>(smem);
    constexpr size_t mb_floats = (sizeof(GfMailbox<R, NW>) + 15) / 16 * 4;
    const int g = threadIdx.x, lane = g & 31, warp = g >> 5;
    const int64_t f = blockIdx.z;
    const float* __restrict__ gI = a.guide + f * a.gfs;
    const float* __restrict__ gP = a.src + f * a.sfs;
    float* __restrict__ gQ = a.dst + f * a.dfs;

    // ring of the last 2R+1 rows of (sum_x a, sum_x b): [KW][2][NT] float4, column g private
    float4* ring = RING_GLOBAL
        ? reinterpret_cast<float4*>(a.ring) + ((size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)) * ((size_t)KW * 2 * NT)
        : reinterpret_cast<float4*>(smem + mb_floats);

    const int x0 = (int)blockIdx.x * WOUT - 8 * H1 + 4 * g;          // first of this thread's 4 extended columns
    const bool vec_ok = x0 >= 0 && x0 + 3 < a.width;
    const int yo0 = a.out_y0 + (int)blockIdx.y * a.hb;
    const int yo1 = min(a.out_y0 + a.out_rows, yo0 + a.hb);
    const bool trunc = a.border == GF_TRUNCATE;
    const bool s1_lane = g >= H1 && g < NT - H1;                     // stage-1 windows complete
    const bool out_lane = g >= 2 * H1 && g < NT - 2 * H1 && x0 < a.width;
    float inv_nx[4];
    bool x_in[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        inv_nx[j] = gf_inv_count(x0 + j, a.width, R, a.border);
        x_in[j] = !trunc || (x0 + j >= 0 && x0 + j < a.width);
    }

    float cI[4], cP[4], cIP[4], cII[4], sA[4], sB[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cI[j] = cP[j] = cIP[j] = cII[j] = sA[j] = sB[j] = 0.f;
    for (int s = 0; s < KW * 2; ++s) ring[(size_t)s * NT + g] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int steps = (yo1 - yo0) + 4 * R;
    int slot = 0;
    for (int t = 0; t < steps; ++t) {
        const int yi = yo0 - 2 * R + t;
        // ---- stage 1, vertical: add row yi, drop row yi - (2R+1)
        {
            const int sy = gf_map(yi, a.height, a.border);
            float vi[4], vp[4];
            gf_load4(gI, a.gs, sy < 0 ? -1 : sy - a.buf_y0, x0, a.width, a.border, vec_ok, vi);
            gf_load4(gP, a.ss, sy < 0 ? -1 : sy - a.buf_y0, x0, a.width, a.border, vec_ok, vp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cI[j] += vi[j]; cP[j] += vp[j];
                cIP[j] = fmaf(vi[j], vp[j], cIP[j]);
                cII[j] = fmaf(vi[j], vi[j], cII[j]);
            }
            if (t >= KW) {
                const int so = gf_map(yi - KW, a.height, a.border);
                gf_load4(gI, a.gs, so < 0 ? -1 : so - a.buf_y0, x0, a.width, a.border, vec_ok, vi);
                gf_load4(gP, a.ss, so < 0 ? -1 : so - a.buf_y0, x0, a.width, a.border, vec_ok, vp);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cI[j] -= vi[j]; cP[j] -= vp[j];
                    cIP[j] = fmaf(-vi[j], vp[j], cIP[j]);
                    cII[j] = fmaf(-vi[j], vi[j], cII[j]);
                }
            }
        }
        if (t < 2 * R) continue;

        // ---- stage 1, horizontal -> a, b of row yc = yi - R
        const int yc = yi - R;
        const GfBlock bI = gf_block(cI), bP = gf_block(cP), bIP = gf_block(cIP), bII = gf_block(cII);
        gf_publish<R, NW>(mb, 0, bI, lane, warp);
        gf_publish<R, NW>(mb, 1, bP, lane, warp);
        gf_publish<R, NW>(mb, 2, bIP, lane, warp);
        gf_publish<R, NW>(mb, 3, bII, lane, warp);
        if (NW > 1) __syncthreads();
        float hI[4], hP[4], hIP[4], hII[4];
        gf_window<R, NW>(cI, bI, hI, mb, 0, lane, warp);
        gf_window<R, NW>(cP, bP, hP, mb, 1, lane, warp);
        gf_window<R, NW>(cIP, bIP, hIP, mb, 2, lane, warp);
        gf_window<R, NW>(cII, bII, hII, mb, 3, lane, warp);

        float va[4], vb[4];
        {
            const bool y_in = !trunc || (yc >= 0 && yc < a.height);
            const float inv_ny = gf_inv_count(yc, a.height, R, a.border);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float norm = inv_nx[j] * inv_ny;
                const float mi = hI[j] * norm, mp = hP[j] * norm;
                const float var = fmaf(-mi, mi, hII[j] * norm);
                const float cov = fmaf(-mi, mp, hIP[j] * norm);
                const float aa = cov * gf_rcp(var + a.eps);
                const bool ok = s1_lane && y_in && x_in[j];
                va[j] = ok ? aa : 0.f;
                vb[j] = ok ? fmaf(-aa, mi, mp) : 0.f;
            }
        }
        if (a.A != nullptr && out_lane && yc >= yo0 && yc < yo1) {
            float* pa = a.A + f * a.abfs + (int64_t)(yc - a.out_y0) * a.abs_ + x0;
            float* pb = a.B + f * a.abfs + (int64_t)(yc - a.out_y0) * a.abs_ + x0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x0 + j < a.width) { pa[j] = va[j]; pb[j] = vb[j]; }
        }

        // ---- stage 2, horizontal
        const GfBlock bA = gf_block(va), bB = gf_block(vb);
        gf_publish<R, NW>(mb, 4, bA, lane, warp);
        gf_publish<R, NW>(mb, 5, bB, lane, warp);
        if (NW > 1) __syncthreads();
        float hA[4], hB[4];
        gf_window<R, NW>(va, bA, hA, mb, 4, lane, warp);
        gf_window<R, NW>(vb, bB, hB, mb, 5, lane, warp);

        // ---- stage 2, vertical through the ring
        {
            float4* ca = ring + ((size_t)slot * 2 + 0) * NT + g;
            float4* cb = ring + ((size_t)slot * 2 + 1) * NT + g;
            const float4 oa = *ca, ob = *cb;
            sA[0] += hA[0] - oa.x; sA[1] += hA[1] - oa.y; sA[2] += hA[2] - oa.z; sA[3] += hA[3] - oa.w;
            sB[0] += hB[0] - ob.x; sB[1] += hB[1] - ob.y; sB[2] += hB[2] - ob.z; sB[3] += hB[3] - ob.w;
            *ca = make_float4(hA[0], hA[1], hA[2], hA[3]);
            *cb = make_float4(hB[0], hB[1], hB[2], hB[3]);
            slot = slot + 1 == KW ? 0 : slot + 1;
        }

        // ---- q of row yo = yi - 2R
        if (t >= 4 * R && out_lane) {
            const int yo = yi - 2 * R;
            const float inv_ny = gf_inv_count(yo, a.height, R, a.border);
            float vi[4], q[4];
            gf_load4(gI, a.gs, yo - a.buf_y0, x0, a.width, a.border, vec_ok, vi);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float norm = inv_nx[j] * inv_ny;
                q[j] = fmaf(sA[j] * norm, vi[j], sB[j] * norm);
            }
            float* pq = gQ + (int64_t)(yo - a.out_y0) * a.ds + x0;
            if (vec_ok) {
                *reinterpret_cast<float4*>(pq) = make_float4(q[0], q[1], q[2], q[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x0 + j >= 0 && x0 + j < a.width) pq[j] = q[j];
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------
template <int R, int NW>
struct GfFastLaunch {
    static const char* go(const Job& j, const char** name)
    {
        using G = GfFastGeom<R>;
        constexpr int NT = NW * 32, KW = 2 * R + 1, WOUT = NT * 4 - 16 * G::H1;
        static_assert(WOUT >= 4, "strip has no output columns");
        constexpr size_t mb_bytes = (sizeof(GfMailbox<R, NW>) + 15) / 16 * 16;
        constexpr size_t ring_bytes = (size_t)KW * 2 * NT * 16;
        const bool ring_global = mb_bytes + ring_bytes > gf_rt_max_smem();
        int sms = 148, mj = 0, mn = 0;
        gf_rt_device_info(&sms, &mj, &mn);
        GfFastArgs a;
        a.guide = j.guide.ptr; a.src = j.src.ptr; a.dst = const_cast<float*>(j.dst.ptr);
        a.A = const_cast<float*>(j.A.ptr); a.B = const_cast<float*>(j.B.ptr);
        a.gs = j.guide.stride; a.ss = j.src.stride; a.ds = j.dst.stride; a.abs_ = j.A.stride;
        a.gfs = j.guide.frame_stride; a.sfs = j.src.frame_stride; a.dfs = j.dst.frame_stride; a.abfs = j.A.frame_stride;
        a.width = j.width; a.height = j.height; a.buf_y0 = j.buf_y0; a.out_y0 = j.out_y0; a.out_rows = j.out_rows;
        a.border = j.border; a.eps = j.eps; a.ring = nullptr;
        const int nstrips = (j.width + WOUT - 1) / WOUT;
        // band height: enough CTAs to fill the machine, but warm-up (4R rows per band) kept small
        const size_t smem = ring_global ? mb_bytes : mb_bytes + ring_bytes;
        int per_sm = (int)(gf_rt_max_smem() / (smem + 1024));
        const int by_threads = 2048 / NT;
        if (per_sm > by_threads) per_sm = by_threads;
        if (per_sm < 1) per_sm = 1;
        int target = sms * per_sm;
        if (const char* e = getenv("GF_FAST_CTAS_PER_SM")) target = sms * atoi(e);
        int nb = target / (nstrips * j.count);
        if (nb < 1) nb = 1;
        int hb = (j.out_rows + nb - 1) / nb;
        int hb_min = 6 * R;
        if (const char* e = getenv("GF_FAST_HB_MIN")) hb_min = atoi(e);
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
    if (getenv("GF_DISABLE_FAST")) return nullptr;
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
    switch (j.r) {
        GF_FAST_CASE(1, 4, "fast_r1")
        GF_FAST_CASE(2, 4, "fast_r2")
        GF_FAST_CASE(3, 4, "fast_r3")
        GF_FAST_CASE(4, 4, "fast_r4")
        GF_FAST_CASE(5, 4, "fast_r5")
        GF_FAST_CASE(6, 4, "fast_r6")
        GF_FAST_CASE(7, 4, "fast_r7")
        GF_FAST_CASE(8, 4, "fast_r8")
        GF_FAST_CASE(12, 4, "fast_r12")
        GF_FAST_CASE(16, 4, "fast_r16")
        GF_FAST_CASE(32, 8, "fast_r32")
    default:
        return nullptr;
    }
#undef GF_FAST_CASE
}
