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
