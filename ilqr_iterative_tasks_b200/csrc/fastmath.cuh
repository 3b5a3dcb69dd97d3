// fastmath.cuh -- branch-free FP64 exp / reciprocal / sincos for the iLQR inner loops.
//
// The CUDA library versions carry slow-path branches (huge arguments, denormals) that split the stage bodies
// of the backward / forward pass into many small basic blocks and materialise their polynomial coefficients
// through uniform-register moves.  The arguments on this path are tame (barrier exponents <= a few units,
// headings of a few radians, well-scaled 2x2 determinants), so straight-line code with the same ~1 ulp
// accuracy lets ptxas interleave the independent chains of a stage (barrier exponentials, linearisation,
// Riccati products).  Accuracy is pinned on the host by tools/host_core_check.py (max ulp vs glibc).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef I2LQR_HOST_BUILD
#include <cuda_runtime.h>
#define I2_HD __device__ __forceinline__
#else
#define I2_HD inline
#endif

namespace i2lqr {

I2_HD double bits_to_double(long long b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d;
  memcpy(&d, &b, sizeof d);
  return d;
#endif
}
I2_HD long long double_to_bits(double d) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(d);
#else
  long long b;
  memcpy(&b, &d, sizeof b);
  return b;
#endif
}
I2_HD double fma_(double a, double b, double c) { return fma(a, b, c); }

// 1/x for normal, finite x (no denormal / overflow handling): hardware seed + Newton steps, <= 1 ulp.
I2_HD double rcp_fast(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma_(-x, r, 1.0);
  e = fma_(e, e, e);
  r = fma_(r, e, r);
  e = fma_(-x, r, 1.0);
  r = fma_(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

// Polynomial coefficients live in constant memory so that they reach the FP64 pipe as c[bank][offset]
// operands instead of being rebuilt from 32-bit immediates inside the stage loops.
#ifdef __CUDACC__
#define I2_CONST __constant__
#else
#define I2_CONST static const
#endif
I2_CONST double kExpEven[6] = {1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 0.5};
I2_CONST double kExpOdd[6] = {1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0};
I2_CONST double kSinC[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                            -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
I2_CONST double kCosC[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                            2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

// exp(x) for x already inside (-745, 709): k = rint(x log2 e), r = x - k ln2 (two-term Cody-Waite),
// degree-13 Taylor polynomial on |r| <= ln2/2 (truncation < 2^-57, even / odd halves evaluated side by
// side to halve the dependent chain), scaled by 2^k through the exponent field; results below the normal
// range (x < -707.7) are zero.
I2_HD double exp_fast_bounded(double x) {
  const double magic = 6755399441055744.0;  // 2^52 + 2^51
  const double t = fma_(x, 1.4426950408889634, magic);
  const int k = (int)(double_to_bits(t) & 0xffffffffLL);  // low word of the magic sum = rint(x log2 e)
  const double kf = t - magic;
  double r = fma_(kf, -6.93147180369123816490e-01, x);
  r = fma_(kf, -1.90821492927058770002e-10, r);
  const double r2 = r * r;
  double pe = kExpEven[0], po = kExpOdd[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) {
    pe = fma_(pe, r2, kExpEven[i]);
    po = fma_(po, r2, kExpOdd[i]);
  }
  // exp(r) = 1 + r + r^2 (pe + r po), in [0.70, 1.42]
  const double q = fma_(po, r, pe);
  const double pr = fma_(q, r2, r) + 1.0;
  // times 2^k by adding k to the exponent field: exact, i.e. the very bits of the product pr * 2^k, as long as the
  // result is a normal number (k >= -1021 here); below that the result is flushed to zero (the products used to
  // degrade through the denormals -- a barrier term of 1e-308 changes nothing it is added to).  One integer add
  // instead of two FP64 multiplications and the construction of their power-of-two operands.
#ifdef __CUDA_ARCH__
  const double scaled = __hiloint2double(__double2hiint(pr) + (k << 20), __double2loint(pr));
#else
  const double scaled = bits_to_double(double_to_bits(pr) + ((long long)k << 52));
#endif
  return k < -1021 ? 0.0 : scaled;
}
// x <= ~700 guaranteed by the caller, arbitrary below (underflows to 0)
I2_HD double exp_fast_lo(double x) { return exp_fast_bounded(fmax(x, -745.0)); }
I2_HD double exp_fast(double x) { return exp_fast_bounded(fmin(fmax(x, -745.0), 709.0)); }

// sin and cos of x for |x| <= 1e5 (larger or non-finite arguments take the library path): quadrant
// reduction with a three-term pi/2 (k * P1, k * P2 exact for |k| < 2^20), fdlibm kernel polynomials
// (|r| <= pi/4, < 1 ulp each).
struct SinCos {
  double sn, cs;
};
#ifdef __CUDACC__
__device__ __noinline__ SinCos sincos_library(double x) {  // out of line: rare
  SinCos r;
  sincos(x, &r.sn, &r.cs);
  return r;
}
#else
inline SinCos sincos_library(double x) { return SinCos{sin(x), cos(x)}; }
#endif
I2_HD void sincos_fast(double x, double* sn, double* cs) {
  if (!(fabs(x) <= 1.0e5)) {
    const SinCos r = sincos_library(x);
    *sn = r.sn;
    *cs = r.cs;
    return;
  }
  const double magic = 6755399441055744.0;
  const double t = fma_(x, 6.36619772367581382433e-01, magic);
  const int k = (int)(double_to_bits(t) & 0xffffffffLL);
  const double kf = t - magic;
  double r = fma_(kf, -1.57079632673412561417e+00, x);   // pio2_1  (33 bits)
  r = fma_(kf, -6.07710050630396597660e-11, r);          // pio2_2  (next 33 bits)
  r = fma_(kf, -2.02226624879595063154e-21, r);          // pio2_3
  const double z = r * r;
  // __kernel_sin / __kernel_cos polynomials, evaluated side by side
  double ps = kSinC[0], pc = kCosC[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) {
    ps = fma_(ps, z, kSinC[i]);
    pc = fma_(pc, z, kCosC[i]);
  }
  const double s = fma_(ps * z, r, r);
  const double c = fma_(pc * z, z, fma_(-0.5, z, 1.0));
  const bool swap = (k & 1) != 0;
  const double ss = swap ? c : s, cc = swap ? s : c;
  // quadrant signs flipped in the sign bit (an integer operation on the high word instead of an FP64 negation and a select)
#ifdef __CUDA_ARCH__
  *sn = __hiloint2double(__double2hiint(ss) ^ ((k & 2) << 30), __double2loint(ss));
  *cs = __hiloint2double(__double2hiint(cc) ^ (((k + 1) & 2) << 30), __double2loint(cc));
#else
  *sn = (k & 2) ? -ss : ss;
  *cs = ((k + 1) & 2) ? -cc : cc;
#endif
}

// ---- single precision (optional FP32 build): the CUDA library routines are already branch-light in float
I2_HD float rcp_fast(float x) { return 1.0f / x; }
I2_HD float exp_fast_bounded(float x) { return expf(x); }
I2_HD float exp_fast_lo(float x) { return expf(x); }
I2_HD void sincos_fast(float x, float* sn, float* cs) {
#ifdef __CUDA_ARCH__
  sincosf(x, sn, cs);
#else
  *sn = sinf(x);
  *cs = cosf(x);
#endif
}

}  // namespace i2lqr
