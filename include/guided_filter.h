// guided_filter.h -- drop-in for the reference's GuidedFilter/guided_filter.h:5-55.
// Same public interface (`init`, `run`, defaults included); the implementation forwards to the
// C ABI of libgf_b200.so (gf_create / gf_run / gf_destroy), which runs ONE fused sm_100a kernel
// per channel instead of the reference's 23 launches through 9 scratch planes.  Link with
// -lgf_b200.
#pragma once
#include "cuda_utils.h"

struct gf_filter;

class GuidedFilter {
public:
    GuidedFilter();
    ~GuidedFilter();

    /*
    Initialization.
    @param:
        _width, _height: size of the images.
        _guided_channel: channels of the guidance image, 1 or 3.
        _src_channel: channels of the source and destination image, 1 or 3.
    The reference allocates 9 pitched scratch planes here and takes its row strides from
    cudaMallocPitch (guided_filter.cpp:70-93), so callers allocate their images with
    cudaMallocPitch of the same row width.  The shim keeps that contract: it derives the same
    strides and allocates nothing.
    */
    void init(const int _width, const int _height, const int _guided_channel = 3, const int _src_channel = 3);

    /*
    Run filter (guidiance, src, dst are device pointers; r = radius; eps = regularisation).
    Border = window clipped to the image (gIntegralToMean semantics).  Asynchronous on stream 0.
    */
    void run(float* guidiance, float* src, float* dst, const int r, const float eps);

private:
    int4 swhcs;  // source: .w width, .x height, .y channels, .z stride (floats) -- the reference's field order
    int4 gwhcs;  // guide
    gf_filter* impl = nullptr;
};
