// jinc_weights.cuh -- device helpers shared by the table kernels and the resample kernels: one radial
// weight = LUT[ round_half_even( 1023 * (dx^2+dy^2) / radius^2 ) ]   (src/JincResize.cpp:485-492).
//
// All arithmetic uses explicit round-to-nearest intrinsics: the reference's expression is evaluated without
// FMA contraction, and the LUT index must come out identical.
#ifndef JINC_WEIGHTS_CUH
#define JINC_WEIGHTS_CUH

#include "jinc_b200.h"

// Squared scaled distance along one axis between the (clamped) sample position and integer tap coordinate:
//   d = (clamp(pos, 0, src_n-1) - tap) * filter_step      float subtraction, then double product (:485-486)
__device__ __forceinline__ double jinc_tap_dist2(float pos, int src_n, int tap, double filt_step)
{
    const float hi = (float)(src_n - 1);
    float c = pos > hi ? hi : pos; // upper bound first, as avs/minmax.h clamp does
    c = c < 0.f ? 0.f : c;
    const double d = __dmul_rn((double)__fsub_rn(c, (float)tap), filt_step);
    return __dmul_rn(d, d);
}

// LUT index of squared distance d2.  The reference computes  llround(1023*d2/radius2 + 1.5*2^52)  which is
// round-half-even of the quotient.  A double division per tap is slow on the GPU, so the quotient is first
// estimated with one multiply by 1023/radius2: both results are within 4 ulp of the exact ratio (< 1e-9
// absolute here), so unless the estimate lies within 1e-6 of a rounding boundary the two round to the same
// integer; only then is the exact divide evaluated.
__device__ __forceinline__ int jinc_lut_index(double d2, double radius2, double idx_scale)
{
    const double est = d2 * idx_scale;
    if (est >= 2.0 * JINC_LUT_SAMPLES)
        return JINC_LUT_SAMPLES; // far outside the support: weight 0 either way
    const double r = rint(est);
    if (fabs(fabs(est - r) - 0.5) > 1e-6)
        return (int)r;
    return __double2int_rn(__ddiv_rn(__dmul_rn((double)(JINC_LUT_SAMPLES - 1), d2), radius2));
}

// Lut::GetFactor (:277-282): 0 at or beyond the table end.
__device__ __forceinline__ float jinc_lut_weight(const float* __restrict__ lut, double d2, double radius2, double idx_scale)
{
    const int idx = jinc_lut_index(d2, radius2, idx_scale);
    return idx >= JINC_LUT_SAMPLES ? 0.f : lut[idx];
}

#endif
