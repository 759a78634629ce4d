/*
 * gf_oracle.c -- CPU restatement of the reference guided filter.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.  The
 * product library (libgf_b200.so) never links or calls it.
 *
 * What it follows (paths relative to /root/reference):
 *   - gf_oracle_guided_gray_f32: GuidedFilter/main.cpp:236-252 (the "mycv" composition of six
 *     cv::blur calls and Mat arithmetic; identical maths at main.cpp:30-47).  cv::blur on
 *     CV_32F is OpenCV's separable box filter (third-party, un-vendored; any OpenCV >= 3):
 *     a RowSum<float,double> running sum along x followed by a ColumnSum<double,float>
 *     running sum along y, one multiply by 1/(kw*kh) in double, cast to float; border
 *     pixels come from borderInterpolate (BORDER_REFLECT_101 by default).  box_rows() and
 *     box_cols() below restate those two loops in the same operation order, so that with
 *     nthreads == 1 the result is the sequential cv::blur result.
 *   - border == GF_ORACLE_TRUNCATE restates `GuidedFilter::run` (guided_filter.cpp:28-66):
 *     window clipped to the image and divided by the true pixel count
 *     (guided_filter_d.cu:251-262), a/b/q formulas at :306-323, :349-362, :382-395.
 *   - gf_oracle_guided_color_*: not in the reference's own code; He et al.
 *     TPAMI 2013 eqs. (19)-(21).  PARITY UNPINNED.
 *
 * Pinned by tests/test_oracle_golden.py against GuidedFilter/data/adobe_image_4_myres.png.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GF_ORACLE_REFLECT101 0
#define GF_ORACLE_TRUNCATE 1
#define GF_ORACLE_REFLECT 2

/* reflectBorder (guided_filter_d.cu:415-418), extended periodically like cv::borderInterpolate. */
static int border_index(int64_t i, int n, int mode)
{
    if (mode == GF_ORACLE_TRUNCATE) return (i >= 0 && i < n) ? (int)i : -1;
    if (n == 1) return 0;
    if (mode == GF_ORACLE_REFLECT101) {
        int64_t period = 2 * (int64_t)n - 2;
        int64_t m = i % period; if (m < 0) m += period;
        return (int)(m < n ? m : period - m);
    } else {
        int64_t period = 2 * (int64_t)n;
        int64_t m = i % period; if (m < 0) m += period;
        return (int)(m < n ? m : period - 1 - m);
    }
}

/* RowSum<T,double>: dst[y][x] = sum_{k=-r..r} src[y][map(x+k)], running sum along x. */
#define DEFINE_BOX_ROWS(NAME, T)                                                              \
static void NAME(const T* src, int64_t sstride, double* dst, int w, int h, int r, int mode,   \
                 const int* xmap, int nthreads)                                               \
{                                                                                             \
    const int k = 2 * r + 1;                                                                  \
    _Pragma("omp parallel num_threads(nthreads)")                                             \
    {                                                                                         \
        T* ext = (T*)malloc(sizeof(T) * (size_t)(w + 2 * r));                                 \
        _Pragma("omp for schedule(static)")                                                   \
        for (int y = 0; y < h; ++y) {                                                         \
            const T* s = src + (int64_t)y * sstride;                                          \
            for (int x = 0; x < w + 2 * r; ++x) ext[x] = xmap[x] < 0 ? (T)0 : s[xmap[x]];     \
            double acc = 0;                                                                   \
            for (int i = 0; i < k; ++i) acc += (double)ext[i];                                \
            double* d = dst + (int64_t)y * w;                                                 \
            d[0] = acc;                                                                       \
            for (int x = 0; x < w - 1; ++x) {                                                 \
                acc += (double)ext[x + k] - (double)ext[x];                                   \
                d[x + 1] = acc;                                                               \
            }                                                                                 \
        }                                                                                     \
        free(ext);                                                                            \
    }                                                                                         \
    (void)mode;                                                                               \
}
DEFINE_BOX_ROWS(box_rows_f32, float)
DEFINE_BOX_ROWS(box_rows_f64, double)

/* ColumnSum<double,T>: running sum along y of the row sums, scaled, cast. Bands per thread. */
#define DEFINE_BOX_COLS(NAME, T)                                                              \
static void NAME(const double* rows, T* dst, int64_t dstride, int w, int h, int r, int mode,  \
                 const int* ymap, const double* xinv, const double* yinv, int nthreads)       \
{                                                                                             \
    const int k = 2 * r + 1;                                                                  \
    const double scale = 1.0 / ((double)k * (double)k);                                       \
    int nb = nthreads < 1 ? 1 : nthreads;                                                     \
    if (nb > h) nb = h;                                                                       \
    _Pragma("omp parallel for num_threads(nthreads) schedule(static)")                        \
    for (int b = 0; b < nb; ++b) {                                                            \
        const int y0 = (int)((int64_t)h * b / nb), y1 = (int)((int64_t)h * (b + 1) / nb);     \
        double* sum = (double*)calloc((size_t)w, sizeof(double));                             \
        for (int i = 0; i < k - 1; ++i) {                                                     \
            int sy = ymap[y0 + i];                                                            \
            if (sy < 0) continue;                                                             \
            const double* sp = rows + (int64_t)sy * w;                                        \
            for (int x = 0; x < w; ++x) sum[x] += sp[x];                                      \
        }                                                                                     \
        for (int y = y0; y < y1; ++y) {                                                       \
            int sy = ymap[y + k - 1], my = ymap[y];                                           \
            const double* sp = sy < 0 ? NULL : rows + (int64_t)sy * w;                        \
            const double* sm = my < 0 ? NULL : rows + (int64_t)my * w;                        \
            T* d = dst + (int64_t)y * dstride;                                                \
            for (int x = 0; x < w; ++x) {                                                     \
                double s0 = sum[x] + (sp ? sp[x] : 0.0);                                      \
                double sc = (mode == GF_ORACLE_TRUNCATE) ? xinv[x] * yinv[y] : scale;         \
                d[x] = (T)(s0 * sc);                                                          \
                sum[x] = s0 - (sm ? sm[x] : 0.0);                                             \
            }                                                                                 \
        }                                                                                     \
        free(sum);                                                                            \
    }                                                                                         \
}
DEFINE_BOX_COLS(box_cols_f32, float)
DEFINE_BOX_COLS(box_cols_f64, double)

typedef struct {
    int w, h, r, mode, nthreads;
    int *xmap, *ymap;
    double *xinv, *yinv, *rows;
} box_ctx;

static int box_ctx_init(box_ctx* c, int w, int h, int r, int mode, int nthreads)
{
    memset(c, 0, sizeof(*c));
    if (w <= 0 || h <= 0 || r < 0) return -1;
    c->w = w; c->h = h; c->r = r; c->mode = mode; c->nthreads = nthreads < 1 ? 1 : nthreads;
    c->xmap = (int*)malloc(sizeof(int) * (size_t)(w + 2 * r));
    c->ymap = (int*)malloc(sizeof(int) * (size_t)(h + 2 * r));
    c->xinv = (double*)malloc(sizeof(double) * (size_t)w);
    c->yinv = (double*)malloc(sizeof(double) * (size_t)h);
    c->rows = (double*)malloc(sizeof(double) * (size_t)w * (size_t)h);
    if (!c->xmap || !c->ymap || !c->xinv || !c->yinv || !c->rows) return -2;
    for (int x = 0; x < w + 2 * r; ++x) c->xmap[x] = border_index((int64_t)x - r, w, mode);
    for (int y = 0; y < h + 2 * r; ++y) c->ymap[y] = border_index((int64_t)y - r, h, mode);
    for (int x = 0; x < w; ++x) {
        int lo = x - r < 0 ? 0 : x - r, hi = x + r > w - 1 ? w - 1 : x + r;
        c->xinv[x] = 1.0 / (double)(hi - lo + 1);
    }
    for (int y = 0; y < h; ++y) {
        int lo = y - r < 0 ? 0 : y - r, hi = y + r > h - 1 ? h - 1 : y + r;
        c->yinv[y] = 1.0 / (double)(hi - lo + 1);
    }
    return 0;
}

static void box_ctx_free(box_ctx* c)
{
    free(c->xmap); free(c->ymap); free(c->xinv); free(c->yinv); free(c->rows);
}

static void box_f32(box_ctx* c, const float* src, int64_t ss, float* dst, int64_t ds)
{
    box_rows_f32(src, ss, c->rows, c->w, c->h, c->r, c->mode, c->xmap, c->nthreads);
    box_cols_f32(c->rows, dst, ds, c->w, c->h, c->r, c->mode, c->ymap, c->xinv, c->yinv, c->nthreads);
}

static void box_f64(box_ctx* c, const double* src, int64_t ss, double* dst, int64_t ds)
{
    box_rows_f64(src, ss, c->rows, c->w, c->h, c->r, c->mode, c->xmap, c->nthreads);
    box_cols_f64(c->rows, dst, ds, c->w, c->h, c->r, c->mode, c->ymap, c->xinv, c->yinv, c->nthreads);
}

int gf_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* cv::blur(src, dst, Size(2r+1, 2r+1)) for one float32 plane (main.cpp:241). */
int gf_oracle_box_mean_f32(const float* src, float* dst, int w, int h, int64_t sstride,
                           int64_t dstride, int r, int border, int nthreads)
{
    box_ctx c;
    int rc = box_ctx_init(&c, w, h, r, border, nthreads);
    if (rc == 0) box_f32(&c, src, sstride, dst, dstride);
    box_ctx_free(&c);
    return rc;
}

/*
 * main.cpp:236-252, float32 Mats.  A and B (optional, may be NULL) receive h_A_mycv / h_B_mycv.
 * All planes are w x h, row stride in floats.
 */
int gf_oracle_guided_gray_f32(const float* I, const float* p, float* q, float* A, float* B, int w,
                              int h, int64_t stride, int r, float eps, int border, int nthreads)
{
    box_ctx c;
    int rc = box_ctx_init(&c, w, h, r, border, nthreads);
    if (rc) { box_ctx_free(&c); return rc; }
    const size_t n = (size_t)w * (size_t)h;
    float* buf = (float*)malloc(sizeof(float) * n * 7);
    if (!buf) { box_ctx_free(&c); return -2; }
    float *pm = buf, *im = buf + n, *ipm = buf + 2 * n, *iim = buf + 3 * n, *t = buf + 4 * n,
          *a = buf + 5 * n, *b = buf + 6 * n;
    const int nt = c.nthreads;
    box_f32(&c, p, stride, pm, w);                         /* cv::blur(h_src, h_Pm_mycv, wsz)      */
    box_f32(&c, I, stride, im, w);                         /* cv::blur(h_guided, h_Im_mycv, wsz)   */
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) t[(size_t)y * w + x] = p[y * stride + x] * I[y * stride + x];
    box_f32(&c, t, w, ipm, w);                             /* cv::blur(h_src.mul(h_guided), ...)   */
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) t[(size_t)y * w + x] = I[y * stride + x] * I[y * stride + x];
    box_f32(&c, t, w, iim, w);                             /* cv::blur(h_guided.mul(h_guided), ..) */
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        /* h_A = (IPm - Pm.mul(Im)) / (IIm - Im.mul(Im) + eps);  h_B = Pm - A.mul(Im) */
        const float pmim = pm[i] * im[i];
        const float imim = im[i] * im[i];
        float num = ipm[i] - pmim;
        float den = (iim[i] - imim) + eps;
        a[i] = num / den;
        const float aim = a[i] * im[i];
        b[i] = pm[i] - aim;
    }
    box_f32(&c, a, w, pm, w);                              /* cv::blur(h_A_mycv, am, wsz) */
    box_f32(&c, b, w, im, w);                              /* cv::blur(h_B_mycv, bm, wsz) */
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            const float t0 = pm[i] * I[y * stride + x];  /* am.mul(h_guided) + bm */
            q[y * stride + x] = t0 + im[i];
            if (A) A[y * stride + x] = a[i];
            if (B) B[y * stride + x] = b[i];
        }
    free(buf);
    box_ctx_free(&c);
    return 0;
}

/* Same composition with double Mats: the ground truth the CUDA path is held to. */
int gf_oracle_guided_gray_f64(const double* I, const double* p, double* q, double* A, double* B,
                              int w, int h, int64_t stride, int r, double eps, int border,
                              int nthreads)
{
    box_ctx c;
    int rc = box_ctx_init(&c, w, h, r, border, nthreads);
    if (rc) { box_ctx_free(&c); return rc; }
    const size_t n = (size_t)w * (size_t)h;
    double* buf = (double*)malloc(sizeof(double) * n * 7);
    if (!buf) { box_ctx_free(&c); return -2; }
    double *pm = buf, *im = buf + n, *ipm = buf + 2 * n, *iim = buf + 3 * n, *t = buf + 4 * n,
           *a = buf + 5 * n, *b = buf + 6 * n;
    const int nt = c.nthreads;
    box_f64(&c, p, stride, pm, w);
    box_f64(&c, I, stride, im, w);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) t[(size_t)y * w + x] = p[y * stride + x] * I[y * stride + x];
    box_f64(&c, t, w, ipm, w);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) t[(size_t)y * w + x] = I[y * stride + x] * I[y * stride + x];
    box_f64(&c, t, w, iim, w);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        a[i] = (ipm[i] - pm[i] * im[i]) / (iim[i] - im[i] * im[i] + eps);
        b[i] = pm[i] - a[i] * im[i];
    }
    box_f64(&c, a, w, pm, w);
    box_f64(&c, b, w, im, w);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            q[y * stride + x] = pm[i] * I[y * stride + x] + im[i];
            if (A) A[y * stride + x] = a[i];
            if (B) B[y * stride + x] = b[i];
        }
    free(buf);
    box_ctx_free(&c);
    return 0;
}

/*
 * Colour guide (He et al. eqs. 19-21), float32 planes, double box sums.
 * I: w x h x 3 interleaved (row stride istride floats), p/q: w x h x 1.
 */
int gf_oracle_guided_color_f32(const float* I, const float* p, float* q, int w, int h,
                               int64_t istride, int64_t pstride, int64_t qstride, int r, float eps,
                               int border, int nthreads)
{
    box_ctx c;
    int rc = box_ctx_init(&c, w, h, r, border, nthreads);
    if (rc) { box_ctx_free(&c); return rc; }
    const size_t n = (size_t)w * (size_t)h;
    /* 13 stage-1 planes + scratch + 4 stage-2 planes */
    float* buf = (float*)malloc(sizeof(float) * n * 19);
    if (!buf) { box_ctx_free(&c); return -2; }
    float* mu[3]; float* ipm[3]; float* ii[6];
    float* pm = buf; float* t = buf + n;
    for (int k = 0; k < 3; ++k) { mu[k] = buf + (2 + k) * n; ipm[k] = buf + (5 + k) * n; }
    for (int k = 0; k < 6; ++k) ii[k] = buf + (8 + k) * n;
    float* av[3] = { buf + 14 * n, buf + 15 * n, buf + 16 * n };
    float* bv = buf + 17 * n; float* t2 = buf + 18 * n;
    const int nt = c.nthreads;
    static const int pi[6] = { 0, 0, 0, 1, 1, 2 }, pj[6] = { 0, 1, 2, 1, 2, 2 };

#define PLANE_LOOP(expr)                                                                      \
    _Pragma("omp parallel for num_threads(nt) schedule(static)")                              \
    for (int y = 0; y < h; ++y)                                                               \
        for (int x = 0; x < w; ++x) {                                                         \
            const float* Ip = I + y * istride + 3 * (int64_t)x; (void)Ip;                     \
            const float pv = p[y * pstride + x]; (void)pv;                                    \
            t[(size_t)y * w + x] = (expr);                                                    \
        }
    box_f32(&c, p, pstride, pm, w);
    for (int k = 0; k < 3; ++k) {
        PLANE_LOOP(Ip[k]);        box_f32(&c, t, w, mu[k], w);
        PLANE_LOOP(Ip[k] * pv);   box_f32(&c, t, w, ipm[k], w);
    }
    for (int k = 0; k < 6; ++k) { PLANE_LOOP(Ip[pi[k]] * Ip[pj[k]]); box_f32(&c, t, w, ii[k], w); }
#undef PLANE_LOOP

#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const double m0 = mu[0][i], m1 = mu[1][i], m2 = mu[2][i], mp = pm[i];
        const double s00 = ii[0][i] - m0 * m0 + eps, s01 = ii[1][i] - m0 * m1, s02 = ii[2][i] - m0 * m2,
                     s11 = ii[3][i] - m1 * m1 + eps, s12 = ii[4][i] - m1 * m2, s22 = ii[5][i] - m2 * m2 + eps;
        const double c0 = ipm[0][i] - m0 * mp, c1 = ipm[1][i] - m1 * mp, c2 = ipm[2][i] - m2 * mp;
        const double i00 = s11 * s22 - s12 * s12, i01 = s02 * s12 - s01 * s22, i02 = s01 * s12 - s02 * s11,
                     i11 = s00 * s22 - s02 * s02, i12 = s01 * s02 - s00 * s12, i22 = s00 * s11 - s01 * s01;
        const double det = s00 * i00 + s01 * i01 + s02 * i02;
        const double a0 = (i00 * c0 + i01 * c1 + i02 * c2) / det,
                     a1 = (i01 * c0 + i11 * c1 + i12 * c2) / det,
                     a2 = (i02 * c0 + i12 * c1 + i22 * c2) / det;
        av[0][i] = (float)a0; av[1][i] = (float)a1; av[2][i] = (float)a2;
        bv[i] = (float)(mp - (a0 * m0 + a1 * m1 + a2 * m2));
    }
    for (int k = 0; k < 3; ++k) { box_f32(&c, av[k], w, t2, w); memcpy(av[k], t2, sizeof(float) * n); }
    box_f32(&c, bv, w, t2, w);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            const float* Ip = I + y * istride + 3 * (int64_t)x;
            q[y * qstride + x] = av[0][i] * Ip[0] + av[1][i] * Ip[1] + av[2][i] * Ip[2] + t2[i];
        }
    free(buf);
    box_ctx_free(&c);
    return 0;
}
