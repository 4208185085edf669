// tbk_common.cuh — shared host/device helpers for the B200 tight-binding
// k-mesh engine.  Everything numerical is written as TBK_HD functions so the
// same source is (a) inlined into the sm_100a kernels and (b) compiled by g++
// into tests/hostemu for CPU-side unit tests of the math (the build container
// has no GPU).  The host build is test infrastructure only.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TBK_HD __host__ __device__ __forceinline__
#define TBK_D __device__ __forceinline__
#else
#define TBK_HD inline
#define TBK_D inline
#endif

namespace tbk {

struct alignas(16) cplx {
  double re, im;
};

TBK_HD cplx mk(double re, double im) { cplx z; z.re = re; z.im = im; return z; }
TBK_HD cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
TBK_HD cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
TBK_HD cplx operator-(cplx a) { return mk(-a.re, -a.im); }
TBK_HD cplx operator*(cplx a, cplx b) {
  return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
TBK_HD cplx operator*(double s, cplx a) { return mk(s * a.re, s * a.im); }
TBK_HD cplx operator*(cplx a, double s) { return mk(s * a.re, s * a.im); }
TBK_HD cplx conj(cplx a) { return mk(a.re, -a.im); }
TBK_HD double norm2(cplx a) { return a.re * a.re + a.im * a.im; }
// a * conj(b)
TBK_HD cplx mulc(cplx a, cplx b) {
  return mk(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
}
// conj(a) * b
TBK_HD cplx cmul(cplx a, cplx b) {
  return mk(a.re * b.re + a.im * b.im, a.re * b.im - a.im * b.re);
}
// acc += a*b
TBK_HD void fma_acc(cplx& acc, cplx a, cplx b) {
  acc.re = fma(a.re, b.re, acc.re);
  acc.re = fma(-a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im);
  acc.im = fma(a.im, b.re, acc.im);
}
// acc += conj(a)*b
TBK_HD void fma_acc_conj(cplx& acc, cplx a, cplx b) {
  acc.re = fma(a.re, b.re, acc.re);
  acc.re = fma(a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im);
  acc.im = fma(-a.im, b.re, acc.im);
}
TBK_HD cplx cdiv(cplx a, cplx b) {
  // Smith's algorithm
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, d = b.re + b.im * r;
    return mk((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  }
  double r = b.re / b.im, d = b.re * r + b.im;
  return mk((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
TBK_HD double cabs_(cplx a) { return hypot(a.re, a.im); }

// a * b with a fixed rounding sequence (no compiler-chosen contraction): used where two code paths
// must produce bit-identical products (periodic images written by different paths of the mesh kernel).
TBK_HD cplx mul_fixed(cplx a, cplx b) {
#if defined(__CUDA_ARCH__)
  return mk(fma(a.re, b.re, -__dmul_rn(a.im, b.im)), fma(a.re, b.im, __dmul_rn(a.im, b.re)));
#else
  return mk(fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re));
#endif
}
// 1/sqrt(q) for q in the normal range (callers clamp away 0/denormals): hardware seed
// (MUFU.RSQ64H, ~2^-20) + one cubically convergent step, no range-check branch — the library
// rsqrt()/sqrt() carry a slow-path branch that splits the basic block of a register-resident
// solver.  Relative error <= ~2 ulp.  Host emulation: 1/sqrt.
TBK_HD double rsqrt_fast(double q) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
  const double e = fma(-q * y, y, 1.0);                 // 1 - q y^2
  const double t = fma(e, 0.375, 0.5);                  // y' = y + y e (1/2 + 3/8 e)
  return fma(y * e, t, y);
#else
  return 1.0 / sqrt(q);
#endif
}
// 1/x for |x| in the normal range (callers keep it away from 0 / denormals / overflow): hardware seed (MUFU.RCP64H,
// ~2^-20) + two Newton steps, no range-check branch and no slow path — the IEEE division the compiler emits is a
// ~25-instruction sequence with a fallback call, and the 4 x 4 eigensolver of the mesh kernel has four of them per QL
// iteration (ncu source page: 15 % of the Kane-Mele kernel's stall samples).  Relative error <= ~1 ulp.  Host: 1/x.
TBK_HD double rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}
// sqrt(q) = q * rsqrt(q) with one Newton correction (q normal, > 0); returns the pair (sqrt, rsqrt)
TBK_HD double sqrt_fast(double q, double* rs_out) {
  const double y = rsqrt_fast(q);
  double r = q * y;
  r = fma(fma(-r, r, q), 0.5 * y, r);                   // r + (q - r^2) / (2 sqrt(q))
  *rs_out = y;
  return r;
}

// exp(2*pi*i*x): x in turns.  Device: sincospi (exact range reduction);
// host emulation: explicit reduction to [-1/2, 1/2] turns before sin/cos.
TBK_HD cplx expi_turns(double x) {
  double s, c;
#if defined(__CUDA_ARCH__)
  sincospi(2.0 * x, &s, &c);
#else
  double r = x - nearbyint(x);
  const double two_pi = 6.283185307179586476925286766559;
  s = sin(two_pi * r);
  c = cos(two_pi * r);
#endif
  return mk(c, s);
}

}  // namespace tbk
