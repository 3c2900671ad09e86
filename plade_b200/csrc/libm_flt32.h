// atan2f / asinf as the reference's libm computes them, for host and device.
//
// pcl::getEulerAngles<float> (common/impl/eigen.hpp:664-669), which ClusterTransformation (PLADE/util.cpp:1260)
// calls per hypothesis, resolves to atan2f / asinf.  Those are not correctly rounded: glibc's differ from
// (float) atan2(double) in ~16 % of arguments and from CUDA's own atan2f / asinf as well.  The reference links
// the system libm -- glibc 2.39 in this image (the GPU box runs the same image) -- whose flt-32 versions are the
// classic fdlibm float algorithms (sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c, e_asinf.c; glibc is not part of
// /root/reference, so this restates the published algorithm).  Only IEEE float + - * / sqrt and bit masks are
// used, the library is built without FMA contraction, and tests/test_oracle_cpu.py pins the host build of this
// header against the host libm bit for bit.  Exception flags / errno are not reproduced.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "linalg.h"

namespace plade {

PLADE_HD int32_t flt_word(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
PLADE_HD float word_flt(int32_t i) { float x; memcpy(&x, &i, 4); return x; }

PLADE_HD float atanf_glibc(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                        6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
  int32_t hx = flt_word(x), ix = hx & 0x7fffffff, id;
  if (ix >= 0x4c000000) {                  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;     // NaN
    return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {                   // |x| < 0.4375
    if (ix < 0x31000000) return x;         // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {                 // |x| < 1.1875
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else { id = 3; x = -1.0f / x; }
    }
  }
  float z = x * x, w = z * z;
  float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return hx < 0 ? -z : z;
}

PLADE_HD float atan2f_glibc(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int32_t hx = flt_word(x), ix = hx & 0x7fffffff, hy = flt_word(y), iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return atanf_glibc(y);
  int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m < 2) return y;
    return m == 2 ? pi + tiny : -pi - tiny;
  }
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny; case 2: return 3.0f * pi_o_4 + tiny; default: return -3.0f * pi_o_4 - tiny; }
    } else {
      switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; }
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_glibc(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return word_flt(flt_word(z) ^ (int32_t) 0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

PLADE_HD float asinf_glibc(float x) {
  const float pio2_hi = 1.57079637050628662109375f, pio2_lo = -4.37113900018624283e-8f, pio4_hi = 0.785398185253143310546875f;
  const float p0 = 1.666675248e-1f, p1 = 7.495297643e-2f, p2 = 4.547037598e-2f, p3 = 2.417951451e-2f, p4 = 4.216630880e-2f;
  int32_t hx = flt_word(x), ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return x * pio2_hi + x * pio2_lo;
  if (ix > 0x3f800000) return (x - x) / (x - x);
  if (ix < 0x3f000000) {                   // |x| < 0.5
    if (ix < 0x32000000) return x;         // |x| < 2^-27
    float t = x * x;
    float w = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
    return x + x * w;
  }
  float w = 1.0f - fabsf(x);
  float t = w * 0.5f;
  float p = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
  float s = sqrtf(t);
  if (ix >= 0x3F79999A) {                  // |x| > 0.975
    t = pio2_hi - (2.0f * (s + s * p) - pio2_lo);
  } else {
    w = word_flt(flt_word(s) & (int32_t) 0xfffff000);
    float c = (t - w * w) / (s + w);
    float r = p;
    p = 2.0f * s * r - (pio2_lo - 2.0f * c);
    float q = pio4_hi - 2.0f * w;
    t = pio4_hi - (p - q);
  }
  return hx > 0 ? t : -t;
}

}  // namespace plade
