// integral_d.h -- drop-in for the reference's Integral/integral_d.h:5-8 (the two integral-image
// launchers; its curand/compare helpers hInitRand, hRandFill, hCmpMaxAbsDiff are demo utilities of
// Integral/main.cpp and are not part of this library).  Same signatures, forwarded to libgf_b200.so.
// Differences a caller can observe: `buff` is only used as scratch for the per-band / per-strip carries (any
// w*h-int buffer as in Integral/main.cpp:52 is large enough; it may be NULL); calls are asynchronous
// on the default stream like the reference's.
#pragma once
#include "cuda_utils.h"

/* Integral image with extra buffer: integral[y*dstride + x] = sum of src[0..y][0..x] (int32, inclusive) */
void hIntegral(unsigned char* src, int* integral, int* buff, int width, int height, int sstride, int dstride);

/* Integral image without extra buffer; the result is dwidth x dheight (multiples of 4), row stride dwidth */
void hAligned4Integral(unsigned char* src, int* integral, int swidth, int sheight, int sstride, int dwidth, int dheight);
