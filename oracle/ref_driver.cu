// extern "C" handles onto the reference's own GPU entry points, so that tests and bench.py can
// run the UNMODIFIED reference (compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libgfref.so) beside the B200 path.  Test/bench infrastructure only.
#include "guided_filter.h"    // /root/reference/GuidedFilter/guided_filter.h
#include "guided_filter_d.h"  // /root/reference/GuidedFilter/guided_filter_d.h

extern "C" {

// Path A: class GuidedFilter (guided_filter.h:5-55).
void* gfref_create(int w, int h, int gch, int sch)
{
    GuidedFilter* g = new GuidedFilter();
    g->init(w, h, gch, sch);
    return g;
}
void gfref_run(void* g, float* guide, float* src, float* dst, int r, float eps)
{
    static_cast<GuidedFilter*>(g)->run(guide, src, dst, r, eps);
}
void gfref_destroy(void* g) { delete static_cast<GuidedFilter*>(g); }

// Path B: hGuidedFilter (guided_filter_d.h:21); a silent no-op for r outside 1..7.
void gfref_hguided(float* guide, float* src, float* dst, float* A, float* B, float eps, int r,
                   int w, int h, int stride)
{
    hGuidedFilter(guide, src, dst, A, B, eps, r, w, h, stride);
}

void gfref_box(float* src, float* dst, float* integral, int w, int h, int c, int stride,
               int istride, int r)
{
    int4 s, i;
    s.w = w; s.x = h; s.y = c; s.z = stride;
    i.w = w + 1; i.x = h + 1; i.y = c; i.z = istride;
    hBoxFilter(src, dst, integral, s, i, r);
}

// The class takes its strides from cudaMallocPitch (guided_filter.cpp:80-92); callers must
// allocate their planes with the same pitch.  Returns that pitch in floats.
int gfref_pitch_floats(int w, int ch, int h)
{
    void* p = nullptr;
    size_t pitch = 0;
    if (cudaMallocPitch(&p, &pitch, (size_t)w * ch * sizeof(float), h) != cudaSuccess) return -1;
    cudaFree(p);
    return (int)(pitch / sizeof(float));
}
}
