// Deterministic fp64 log / exp / cos(2*pi*u): the fdlibm algorithms (e_log.c, e_exp.c,
// k_sin.c, k_cos.c) in plain IEEE mul/add/div.  This translation unit is compiled with
// -fmad=false, so no operation is contracted and every result is reproducible bit for bit on
// any IEEE machine — CUDA's libdevice log/exp differ from glibc in the last ulp, which would
// flip UCB arg-max ties (src/solver.jl:117) and fixed-point SIR weights between runs.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#ifdef __CUDACC__
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD inline
#endif

namespace pomcp {

DM_HD uint64_t d2u(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; std::memcpy(&u, &x, 8); return u;
#endif
}
DM_HD double u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x; std::memcpy(&x, &u, 8); return x;
#endif
}

DM_HD double det_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  if (x == 0.0) return -INFINITY;
  if (x < 0.0 || x != x) return NAN;
  if (x == INFINITY) return x;
  uint64_t ux = d2u(x);
  int k = 0;
  if ((ux >> 52) == 0) {
    x *= 18014398509481984.0;
    ux = d2u(x);
    k -= 54;
  }
  uint32_t hx = (uint32_t)(ux >> 32);
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  uint32_t i = (hx + 0x95f64u) & 0x100000u;
  ux = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (ux & 0xffffffffu);
  k += (int)(i >> 20);
  double m = u2d(ux);
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double dk = (double)k;
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

DM_HD double det_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
               P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 709.0) return INFINITY;
  if (x < -745.0) return 0.0;
  int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
  double t = (double)k;
  double hi = x - t * ln2HI;
  double lo = t * ln2LO;
  double r = hi - lo;
  double tt = r * r;
  double c = r - tt * (P1 + tt * (P2 + tt * (P3 + tt * (P4 + tt * P5))));
  double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  if (k >= -1000) return y * u2d((uint64_t)(k + 1023) << 52);
  return (y * u2d((uint64_t)(k + 1000 + 1023) << 52)) * u2d((uint64_t)(-1000 + 1023) << 52);
}

DM_HD double det_ksin(double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = y * y;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return y + y * (z * (S1 + z * r));
}
DM_HD double det_kcos(double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = y * y;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return (1.0 - 0.5 * z) + z * r;
}
DM_HD double det_cos2pi(double u) {
  const double HALF_PI = 1.57079632679489661923;
  double t = u * 4.0;
  int q = (int)t;
  double f = t - (double)q;
  double sn, cs;
  if (f <= 0.5) {
    double y = f * HALF_PI;
    sn = det_ksin(y); cs = det_kcos(y);
  } else {
    double y = (1.0 - f) * HALF_PI;
    sn = det_kcos(y); cs = det_ksin(y);
  }
  switch (q & 3) {
    case 0: return cs;
    case 1: return -sn;
    case 2: return -cs;
    default: return sn;
  }
}
// Box–Muller stand-in for randn(rng) (notebooks/Minimal_Example.ipynb:91,142)
DM_HD double det_normal(double u0, double u1) { return sqrt(-2.0 * det_log(1.0 - u0)) * det_cos2pi(u1); }

}  // namespace pomcp
