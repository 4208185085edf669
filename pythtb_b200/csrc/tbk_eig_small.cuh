// tbk_eig_small.cuh — one-thread-per-matrix Hermitian eigensolvers for the
// tiny matrices of 2-band / spinor models (replaces numpy.linalg.eigh at
// pythtb.py:939/944 for n <= 4).  Everything stays in registers.
//
// Conventions (pythtb.py:947-949): eigenvalues ascending, eigenvector b is
// returned as the ROW w[b][0..n), not conjugated: H * w[b]^T = ev[b] * w[b]^T.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

// ---------------------------------------------------------------------------
// n = 2, closed form.  H = [[h00, conj(h10)], [h10, h11]] (lower triangle in,
// like LAPACK UPLO='L' which numpy uses by default).
// ---------------------------------------------------------------------------
TBK_HD void eigh2(double h00, double h11, cplx h10, double ev[2], cplx w[2][2], bool want_vec) {
  const double mean = 0.5 * (h00 + h11);
  const double delta = 0.5 * (h00 - h11);
  const double b2 = norm2(h10);
  const double r = sqrt(delta * delta + b2);
  ev[0] = mean - r;
  ev[1] = mean + r;
  if (!want_vec) return;
  const cplx b = conj(h10);  // H[0][1]
  if (b2 == 0.0) {
    // diagonal matrix: order the unit vectors by eigenvalue
    const bool swap = h00 > h11;
    w[0][0] = mk(swap ? 0.0 : 1.0, 0.0);
    w[0][1] = mk(swap ? 1.0 : 0.0, 0.0);
    w[1][0] = mk(swap ? 1.0 : 0.0, 0.0);
    w[1][1] = mk(swap ? 0.0 : 1.0, 0.0);
    return;
  }
  // cancellation-free choice of the two null-vector formulas
  const double big = fabs(delta) + r;                  // |delta| + r > 0
#if defined(__CUDA_ARCH__)
  const double inv = rsqrt(b2 + big * big);
#else
  const double inv = 1.0 / sqrt(b2 + big * big);
#endif
  if (delta >= 0.0) {
    w[0][0] = inv * b;        w[0][1] = mk(-big * inv, 0.0);   // lower band
    w[1][0] = mk(big * inv, 0.0); w[1][1] = inv * conj(b);     // upper band
  } else {
    w[0][0] = mk(-big * inv, 0.0); w[0][1] = inv * conj(b);
    w[1][0] = inv * b;        w[1][1] = mk(big * inv, 0.0);
  }
}

// Branch-free variant for the mesh kernels (eigenvectors always wanted).  Same formulas as eigh2
// (cancellation-free null vectors), but: no slow-path sqrt/rsqrt branches, the diagonal /
// fully degenerate matrix is handled by a 1e-290 offset instead of a branch, and the delta >= 0 / < 0
// layouts are chosen by selects, so that several independent matrices interleave in one basic block.
// Accuracy: eigenvalues <= 2 ulp of max|H|, eigenvectors orthonormal to ~4e-16.
TBK_HD void eigh2_fast(double h00, double h11, cplx h10, double ev[2], cplx w[2][2]) {
  const double tiny = 1.0e-290;
  const double mean = 0.5 * (h00 + h11);
  const double delta = 0.5 * (h00 - h11);
  const double b2 = norm2(h10);
  double rs;
  const double r = sqrt_fast(fma(delta, delta, b2) + tiny, &rs);   // + tiny: no-op unless H is exactly degenerate
  ev[0] = mean - r;
  ev[1] = mean + r;
  const double big = fabs(delta) + r;                       // > 0
  const double inv = rsqrt_fast(fma(big, big, b2) + tiny);
  const double x = big * inv;
  const cplx y = mk(h10.re * inv, -h10.im * inv);           // inv * H[0][1]
  const bool pos = delta >= 0.0;
  w[0][0] = mk(pos ? y.re : -x, pos ? y.im : 0.0);
  w[0][1] = mk(pos ? -x : y.re, pos ? 0.0 : -y.im);
  w[1][0] = mk(pos ? x : y.re, pos ? 0.0 : y.im);
  w[1][1] = mk(pos ? y.re : x, pos ? -y.im : 0.0);
}

// ---------------------------------------------------------------------------
// 3 <= N <= 4 (compile time): cyclic complex Jacobi on packed storage.
//   dg[N]                real diagonal
//   lo[N(N-1)/2]         strict lower triangle, lo[r(r-1)/2 + c] = H[r][c], r > c
//   w[N][N]              on exit rows = eigenvectors (if want_vec)
// All loops over matrix indices are fully unrolled so the arrays live in
// registers.
// ---------------------------------------------------------------------------
template <int N>
struct JacobiPacked {
  static constexpr int NL = N * (N - 1) / 2;
  TBK_HD static constexpr int idx(int r, int c) { return r * (r - 1) / 2 + c; }

  TBK_HD static cplx get(const cplx lo[], int r, int c) {
    return r > c ? lo[idx(r, c)] : conj(lo[idx(c, r)]);
  }
  TBK_HD static void set(cplx lo[], int r, int c, cplx v) {
    if (r > c) lo[idx(r, c)] = v; else lo[idx(c, r)] = conj(v);
  }

  TBK_HD static void solve(double dg[N], cplx lo[NL], cplx w[N][N], bool want_vec) {
    if (want_vec) {
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) w[a][b] = mk(a == b ? 1.0 : 0.0, 0.0);
    }
    double fro = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) fro = fma(dg[a], dg[a], fro);
#pragma unroll
    for (int a = 0; a < NL; ++a) fro += 2.0 * norm2(lo[a]);
    const double tol = 1.0e-31 * fro;   // |off|_F <= 3e-16 |H|_F
    for (int sweep = 0; sweep < 24; ++sweep) {
      double off = 0.0;
#pragma unroll
      for (int a = 0; a < NL; ++a) off += norm2(lo[a]);
      if (off <= tol) break;
#pragma unroll
      for (int p = 0; p < N - 1; ++p) {
#pragma unroll
        for (int q = p + 1; q < N; ++q) {
          const cplx apq = conj(lo[idx(q, p)]);       // H[p][q] = g e^{i phi}
          const double g2 = norm2(apq);
          if (g2 > 1.0e-300) {
            // Rotation J = [[c, s e^{i phi}], [-s e^{-i phi}, c]] with t = tan(theta) the smaller root of
            // t^2 + 2 tau t - 1 = 0, tau = (d_q - d_p) / (2 g).  Written without normalising apq:
            //   big = |delta| + sqrt(delta^2 + g^2), delta = (d_q - d_p)/2
            //   t g = sgn(delta) g^2 / big,  c = big / sqrt(big^2 + g^2),  s e^{-i phi} = sgn(delta) conj(apq) / sqrt(big^2 + g^2)
            // (one sqrt, one rsqrt, one division per rotation).
            const double delta = 0.5 * (dg[q] - dg[p]);
            const double big = fabs(delta) + sqrt(fma(delta, delta, g2));
            const double sgn = delta >= 0.0 ? 1.0 : -1.0;
            const double tg = sgn * (g2 / big);
#if defined(__CUDA_ARCH__)
            const double rs = rsqrt(fma(big, big, g2));
#else
            const double rs = 1.0 / sqrt(fma(big, big, g2));
#endif
            const double c = big * rs;
            const cplx se = (sgn * rs) * conj(apq);    // s e^{-i phi}
            const cplx sc = conj(se);                  // s e^{+i phi}
            dg[p] -= tg;
            dg[q] += tg;
            lo[idx(q, p)] = mk(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < N; ++r) {
              if (r != p && r != q) {
                const cplx arp = get(lo, r, p), arq = get(lo, r, q);
                set(lo, r, p, c * arp - arq * se);
                set(lo, r, q, arp * sc + c * arq);
              }
            }
            if (want_vec) {
#pragma unroll
              for (int o = 0; o < N; ++o) {
                const cplx vp = w[p][o], vq = w[q][o];
                w[p][o] = c * vp - vq * se;
                w[q][o] = vp * sc + c * vq;
              }
            }
          }
        }
      }
    }
    // ascending order (selection sort, unrolled; swaps rows of w)
#pragma unroll
    for (int a = 0; a < N - 1; ++a) {
#pragma unroll
      for (int b = a + 1; b < N; ++b) {
        if (dg[b] < dg[a]) {
          const double t = dg[a]; dg[a] = dg[b]; dg[b] = t;
          if (want_vec) {
#pragma unroll
            for (int o = 0; o < N; ++o) { const cplx z = w[a][o]; w[a][o] = w[b][o]; w[b][o] = z; }
          }
        }
      }
    }
  }
};

}  // namespace tbk
