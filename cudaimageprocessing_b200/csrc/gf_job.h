// gf_job.h -- host-side description of one filtering request, shared by the API layer and
// the kernel-family dispatchers.
#pragma once
#include "gf_common.cuh"
#include "gf_knobs.h"

struct Plane {
    const float* ptr;
    int64_t stride, frame_stride;
    int channels;  // interleaved channels of the buffer
    int coff;
};

inline GfPlane mk(const Plane& p)
{
    GfPlane g;
    g.ptr = const_cast<float*>(p.ptr);
    g.stride = p.stride;
    g.frame_stride = p.frame_stride;
    g.xstep = p.channels;
    g.coff = p.coff;
    return g;
}

struct Job {
    Plane guide, src, dst, A, B;
    int count = 1;
    int width = 0, height = 0;  // height = global image height
    int buf_y0 = 0, buf_rows = 0, out_y0 = 0, out_rows = 0;
    int r = 0, border = 0;
    float eps = 0.f;
    bool color = false;         // guide has 3 channels used jointly (3x3 covariance)
    void* stream = nullptr;
};

