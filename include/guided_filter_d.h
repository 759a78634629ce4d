// guided_filter_d.h -- drop-in for the reference's GuidedFilter/guided_filter_d.h:6-21.
// Same six host launchers, same signatures (`int4 whcs` packs .w=width .x=height .y=channels
// .z=stride in floats, guided_filter.cpp:20-22,90-92).  Each forwards to libgf_b200.so.
// Differences a caller can observe: no cudaDeviceSynchronize() after each launch; the
// `integral` scratch of hBoxFilter is ignored (window sums are exact, not a float32 integral
// image); hCalcB with a 1-channel guide computes b = pm - a*im (the reference's gCalcBCN1 does
// not, guided_filter_d.cu:371-372); hGuidedFilter accepts any radius (the reference silently
// does nothing outside 1..7, :1090).  d_A / d_B are filled with a and b as in the reference (main.cpp:283-286 reads
// them back); a caller that never looks can set GF_SHIM_SKIP_AB=1 in the environment to save the 8 B/px of writes.
// A request the library cannot serve (other than the channel combinations the reference refuses too) is fatal, like a
// CUDA error in the reference (cuda_utils.h): the call never returns with dst unwritten.
#pragma once
#include "cuda_utils.h"

/* Box filter */
void hBoxFilter(float* src, float* dst, float* integral, const int4& swhcs, const int4& iwhcs, const int r);

/* Multiplication */
void hMultiply(float* a, float* b, float* c, const int4& awhcs, const int4& bwhcs);

/* Compute A */
void hCalcA(float* a, float* pm, float* im, float* ipm, float* iim, const int4& swhcs, const int4& gwhcs, const float eps);

/* Compute B */
void hCalcB(float* b, float* a, float* pm, float* im, const int4& swhcs, const int4& gwhcs);

/* Linear transform */
void hLinearTransform(float* src, float* dst, float* a, float* b, const int4& swhcs, const int4& dwhcs);

/* Guided filter for one channel */
void hGuidedFilter(float* d_guided, float* d_src, float* d_dst, float* d_A, float* d_B, float eps, int radius, int width, int height, int stride);
