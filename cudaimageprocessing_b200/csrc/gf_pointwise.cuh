// gf_pointwise.cuh -- the four element-wise launchers of path A (guided_filter_d.h:9-18), kept
// for call-surface compatibility.  The fused kernels do not use them.
#pragma once
#include "gf_common.cuh"

struct GfPwArgs {
    const float* in0; const float* in1; const float* in2; const float* in3;
    float* out;
    int width, height;
    int cs;          // channels of the "source-shaped" operands and of the output
    int cg;          // channels of the "guide-shaped" operands (cs or 1)
    int64_t ss, sg;  // row strides (floats)
    float eps;
};

enum { GF_PW_MUL = 0, GF_PW_CALC_A = 1, GF_PW_CALC_B = 2, GF_PW_LINEAR = 3 };

// One thread per pixel, 2-D grid-stride over rows.  idx_s addresses cs-channel planes,
// idx_g the cg-channel planes (broadcast over channels when cg == 1: the CN1 variants).
template <int OP>
__global__ void __launch_bounds__(256) gf_pointwise_kernel(const GfPwArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.width) return;
    for (int y = blockIdx.y; y < a.height; y += gridDim.y) {
        for (int c = 0; c < a.cs; ++c) {
            const int64_t is = (int64_t)y * a.ss + (int64_t)x * a.cs + c;
            const int64_t ig = (int64_t)y * a.sg + (int64_t)x * a.cg + (a.cg == 1 ? 0 : c);
            if (OP == GF_PW_MUL) {
                // hMultiply: c = a * b (gMultiply / gMultiplyCN1, guided_filter_d.cu:273-303)
                a.out[is] = a.in0[is] * a.in1[ig];
            } else if (OP == GF_PW_CALC_A) {
                // hCalcA: in0=pm in1=im in2=ipm in3=iim (gCalcA, :306-323): eps joins iim first
                const float vim = a.in1[ig];
                const float num = fmaf(a.in0[is], -vim, a.in2[is]);
                const float den = fmaf(-vim, vim, a.in3[ig] + a.eps);
                a.out[is] = num / den;
            } else if (OP == GF_PW_CALC_B) {
                // hCalcB: in0=a in1=im in2=pm (gCalcB, :349-362)
                a.out[is] = fmaf(a.in0[is], -a.in1[ig], a.in2[is]);
            } else {
                // hLinearTransform: in0=src(guide-shaped) in1=a in2=b (gLinearTransform, :382-395)
                a.out[is] = fmaf(a.in0[ig], a.in1[is], a.in2[is]);
            }
        }
    }
}

// ---- channel-interleaved <-> planar (the class API's (1,3)/(3,3) modes run the planar s8 kernel) ----
// TO_PLANAR: planar[c][y][x] = inter[y*stride + x*C + c];  otherwise the reverse.  Thread per pixel,
// 2-D grid-stride over rows; both sides are coalesced across the warp.
struct GfIlArgs {
    float* inter; float* planar;
    int64_t stride;          // row stride of the interleaved buffer (floats)
    int64_t pitch, plane;    // row pitch and plane size of the planar buffer (floats)
    int width, height, channels;
};

template <bool TO_PLANAR>
__global__ void __launch_bounds__(256) gf_interleave_kernel(const GfIlArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.width) return;
    for (int y = blockIdx.y; y < a.height; y += gridDim.y) {
        for (int c = 0; c < a.channels; ++c) {
            float* pi = a.inter + (int64_t)y * a.stride + (int64_t)x * a.channels + c;
            float* pp = a.planar + (int64_t)c * a.plane + (int64_t)y * a.pitch + x;
            if (TO_PLANAR) *pp = *pi;
            else *pi = *pp;
        }
    }
}
