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
  // big >= sqrt(tiny) = 1e-145, so big^2 + b2 >= 1e-290 is a normal number without a further offset (adding
  // `tiny` again halved the norm of both vectors when H is exactly proportional to the identity)
  const double inv = rsqrt_fast(fma(big, big, b2));
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

#if defined(__CUDA_ARCH__)
#define TBK_UNROLL _Pragma("unroll")
#else
#define TBK_UNROLL
#endif

// ---------------------------------------------------------------------------
// 3 <= N <= 4, direct method: Householder tridiagonalisation of the complex Hermitian matrix
// (N - 1 reflectors, the last one a pure phase), implicit-shift QL on the REAL tridiagonal with the
// rotations accumulated in a REAL N x N matrix, back-transformation of its columns by the
// reflectors.  About 3x fewer FP64 operations than the cyclic complex Jacobi above (ncu: the
// Kane-Mele mesh kernel was bound by the ~2800 FP64 instructions of Jacobi per k-point) and half
// the live registers during the iteration.  All matrix indices are compile-time constants (loops
// fully unrolled, data-dependent extents as predicates), so nothing is spilled to local memory.
//   a[N][N]   Hermitian matrix, only the lower triangle is read; destroyed
//   ev[N]     ascending eigenvalues
//   w[N][N]   rows = eigenvectors (not conjugated), H w[b]^T = ev[b] w[b]^T
// Returns false if the QL iteration did not converge in 30 steps (callers fall back to Jacobi).
// ---------------------------------------------------------------------------
// (1) zhetd2 'L': a -> real tridiagonal (d, e), reflectors left in a / tau
template <int N>
TBK_HD void small_hetd2(cplx a[N][N], double d[N], double e[N], cplx tau[N]) {
  // ---- zhetd2 'L': reflector j annihilates a[j+2..N-1][j]; v_j kept in a[j+2..][j], implicit unit at j+1
  TBK_UNROLL
  for (int j = 0; j < N - 1; ++j) {
    double xnorm2 = 0.0;
    TBK_UNROLL
    for (int r = j + 2; r < N; ++r) xnorm2 += norm2(a[r][j]);
    const cplx alpha = a[j + 1][j];
    cplx t = mk(0.0, 0.0);
    double beta = alpha.re;
    if (xnorm2 != 0.0 || alpha.im != 0.0) {
      double rs;
      const double nrm = sqrt_fast(alpha.re * alpha.re + alpha.im * alpha.im + xnorm2 + 1.0e-290, &rs);
      beta = alpha.re >= 0.0 ? -nrm : nrm;
      const double ib = rcp_fast(beta);                          // |beta| >= 1e-145
      t = mk((beta - alpha.re) * ib, -alpha.im * ib);
      const double dr = alpha.re - beta, di = alpha.im;          // 1 / (alpha - beta)
      const double idn = rcp_fast(dr * dr + di * di);           // |alpha - beta|^2 >= beta^2 >= 1e-290
      const cplx scal = mk(dr * idn, -di * idn);
      TBK_UNROLL
      for (int r = j + 2; r < N; ++r) a[r][j] = a[r][j] * scal;
    }
    d[j] = a[j][j].re;
    e[j] = beta;
    tau[j] = t;
    if (j < N - 2) {
      // p = tau A22 v, w = p - (tau/2)(p^H v) v, A22 -= v w^H + w v^H   (lower triangle of A22 only)
      cplx v[N], pv[N];
      TBK_UNROLL
      for (int r = j + 1; r < N; ++r) v[r] = r == j + 1 ? mk(1.0, 0.0) : a[r][j];
      TBK_UNROLL
      for (int r = j + 1; r < N; ++r) {
        cplx acc = mk(0.0, 0.0);
        TBK_UNROLL
        for (int c = j + 1; c < N; ++c) {
          const cplx arc = c < r ? a[r][c] : (c == r ? mk(a[r][r].re, 0.0) : conj(a[c][r]));
          fma_acc(acc, arc, v[c]);
        }
        pv[r] = t * acc;
      }
      cplx dot = mk(0.0, 0.0);
      TBK_UNROLL
      for (int r = j + 1; r < N; ++r) fma_acc_conj(dot, pv[r], v[r]);     // p^H v
      const cplx a2 = (-0.5) * (t * dot);
      TBK_UNROLL
      for (int r = j + 1; r < N; ++r) pv[r] = pv[r] + a2 * v[r];
      TBK_UNROLL
      for (int r = j + 1; r < N; ++r) {
        TBK_UNROLL
        for (int c = j + 1; c <= r; ++c) a[r][c] = a[r][c] - mulc(v[r], pv[c]) - mulc(pv[r], v[c]);
      }
    }
    // j == N-2: the reflector is the scalar 1 - tau on component N-1; a[N-1][N-1] is unchanged (|1 - tau| = 1)
  }
  d[N - 1] = a[N - 1][N - 1].re;
  e[N - 1] = 0.0;
}

// (2) implicit-shift QL on (d, e): on exit d ascending, column c of z the eigenvector of d[c]; false = not converged
// Where the rotations go: Z provides init(), rot(i, sn, cs) (columns i, i + 1 of the accumulated rotation matrix) and
// swap(x, y) (the final ordering).  ZRegs: an N x N register array (the column indices are compile-time constants after
// unrolling); ZNone: eigenvalues only; ZMem: a strided array in (shared) memory + a column permutation instead of swaps.
template <int N>
struct ZRegs {
  double (*z)[N];
  TBK_HD void init() {
    TBK_UNROLL
    for (int r = 0; r < N; ++r) {
      TBK_UNROLL
      for (int c = 0; c < N; ++c) z[r][c] = r == c ? 1.0 : 0.0;
    }
  }
  TBK_HD void rot(int i, double sn, double cs) {
    TBK_UNROLL
    for (int k = 0; k < N; ++k) {
      const double zf = z[k][i + 1];
      z[k][i + 1] = sn * z[k][i] + cs * zf;
      z[k][i] = cs * z[k][i] - sn * zf;
    }
  }
  TBK_HD void swap(int x, int y) {
    TBK_UNROLL
    for (int k = 0; k < N; ++k) { const double tz = z[k][x]; z[k][x] = z[k][y]; z[k][y] = tz; }
  }
};
struct ZNone {
  TBK_HD void init() {}
  TBK_HD void rot(int, double, double) {}
  TBK_HD void swap(int, int) {}
};
template <int N>
struct ZMem {
  double* z;        // element (k, i) at z[(k * N + i) * stride]
  int stride;
  int perm[N];      // on exit: column perm[b] belongs to the b-th smallest eigenvalue
  TBK_HD void init() {
    TBK_UNROLL
    for (int r = 0; r < N; ++r) {
      perm[r] = r;
      TBK_UNROLL
      for (int c = 0; c < N; ++c) z[(r * N + c) * stride] = r == c ? 1.0 : 0.0;
    }
  }
  TBK_HD void rot(int i, double sn, double cs) {
    TBK_UNROLL
    for (int k = 0; k < N; ++k) {
      double* p = z + (k * N + i) * stride;
      const double z0 = p[0], z1 = p[stride];
      p[stride] = sn * z0 + cs * z1;
      p[0] = cs * z0 - sn * z1;
    }
  }
  TBK_HD void swap(int x, int y) { const int t = perm[x]; perm[x] = perm[y]; perm[y] = t; }
};

template <int N, class Z>
TBK_HD bool small_tridiag_ql_t(double d[N], double e[N], Z& z) {
  const double eps = 1.1102230246251565e-16;
  z.init();
  bool ok = true;
  TBK_UNROLL
  for (int l = 0; l < N - 1; ++l) {
    for (int iter = 0;; ++iter) {
      int m = N - 1;
      TBK_UNROLL
      for (int mm = N - 2; mm >= l; --mm)
        if (fabs(e[mm]) <= eps * (fabs(d[mm]) + fabs(d[mm + 1]))) m = mm;
      if (m == l) break;
      if (iter == 30) { ok = false; break; }
      double dm = d[N - 1];
      TBK_UNROLL
      for (int mm = l + 1; mm < N - 1; ++mm)
        if (mm == m) dm = d[mm];
      // e[l] != 0 here (m != l).  The shift only steers the convergence: a reciprocal good to an ulp is plenty.
      // Tiny |e[l]| (the quotient overflows) just gives the shift d[l] - as the IEEE division would.
      const double el = fabs(e[l]) < 1.0e-290 ? (e[l] < 0.0 ? -1.0e-290 : 1.0e-290) : e[l];   // keep the reciprocal finite
      double g = (d[l + 1] - d[l]) * (0.5 * rcp_fast(el));
      double rs;
      double r = sqrt_fast(g * g + 1.0, &rs);
      g = dm - d[l] + e[l] * rcp_fast(g + (g >= 0.0 ? r : -r));
      double sn = 1.0, cs = 1.0, p = 0.0;
      bool dead = false;
      TBK_UNROLL
      for (int i = N - 2; i >= l; --i) {
        if (i < m && !dead) {
          const double f = sn * e[i], b = cs * e[i];
          const double q2 = f * f + g * g;
          if (q2 == 0.0) {
            e[i + 1] = 0.0;
            d[i + 1] -= p;
            dead = true;
          } else {
            r = sqrt_fast(q2, &rs);
            e[i + 1] = r;
            sn = f * rs; cs = g * rs;
            g = d[i + 1] - p;
            r = (d[i] - g) * sn + 2.0 * cs * b;
            p = sn * r;
            d[i + 1] = g + p;
            g = cs * r - b;
            z.rot(i, sn, cs);
          }
        }
      }
      if (!dead) { d[l] -= p; e[l] = g; }
      TBK_UNROLL
      for (int mm = l; mm < N; ++mm)
        if (mm == m) e[mm] = 0.0;
    }
  }
  // ---- ascending order: sort d together with the (real) columns of z
  TBK_UNROLL
  for (int x = 0; x < N - 1; ++x) {
    TBK_UNROLL
    for (int y = x + 1; y < N; ++y) {
      if (d[y] < d[x]) {
        const double td = d[x]; d[x] = d[y]; d[y] = td;
        z.swap(x, y);
      }
    }
  }
  return ok;
}

// WANT_Z = false: eigenvalues only (z is not touched; pass any array)
template <int N, bool WANT_Z = true>
TBK_HD bool small_tridiag_ql(double d[N], double e[N], double z[N][N]) {
  if constexpr (WANT_Z) {
    ZRegs<N> zr{z};
    return small_tridiag_ql_t<N>(d, e, zr);
  } else {
    ZNone zn;
    return small_tridiag_ql_t<N>(d, e, zn);
  }
}

// (3) eigenvectors of H: x = H_0 H_1 ... H_{N-2} z_c, row b of w = eigenvector b
template <int N>
TBK_HD void small_backtransform(const cplx a[N][N], const cplx tau[N], const double d[N], const double z[N][N], double ev[N], cplx w[N][N]) {
  TBK_UNROLL
  for (int b = 0; b < N; ++b) {
    cplx x[N];
    TBK_UNROLL
    for (int k = 0; k < N; ++k) x[k] = mk(z[k][b], 0.0);
    TBK_UNROLL
    for (int j = N - 2; j >= 0; --j) {
      cplx dot = x[j + 1];                                   // v^H x with v[j+1] = 1
      TBK_UNROLL
      for (int r = j + 2; r < N; ++r) fma_acc_conj(dot, a[r][j], x[r]);
      const cplx f = tau[j] * dot;
      x[j + 1] = x[j + 1] - f;
      TBK_UNROLL
      for (int r = j + 2; r < N; ++r) x[r] = x[r] - f * a[r][j];
    }
    ev[b] = d[b];
    TBK_UNROLL
    for (int k = 0; k < N; ++k) w[b][k] = x[k];
  }
}

template <int N>
TBK_HD bool eigh_small_ql(cplx a[N][N], double ev[N], cplx w[N][N]) {
  double d[N], e[N];
  cplx tau[N];
  small_hetd2<N>(a, d, e, tau);
  double z[N][N];
  const bool ok = small_tridiag_ql<N>(d, e, z);
  small_backtransform<N>(a, tau, d, z, ev, w);
  return ok;
}

// ---------------------------------------------------------------------------
// N = 4, direct solver of the real symmetric tridiagonal (d, e) — the hot path of the Kane-Mele mesh kernel.
// The implicit-QL iteration above is a serial dependency chain (shift, rotation, rotation, ... ~7 iterations per
// matrix): ncu r11 showed 56 % of the kernel's FP64 instructions in it at an ILP of about one.  Here instead:
//   eigenvalues   roots of the characteristic quartic in closed form (Ferrari: largest root of the resolvent cubic by
//                 the trigonometric formula — float acos / cos, ~1e-7, is plenty because —) two Newton steps per root
//                 on the Sturm recurrence p_k = (d_k - x) p_{k-1} - e_{k-1}^2 p_{k-2} restore full accuracy
//                 (four independent chains);
//   eigenvectors  the column of adj(T - lambda) with the largest diagonal entry (the twisted factorisation's choice
//                 of the twist index, written with leading / trailing minors: no divisions).
// Valid while the roots are simple and not too close: the function returns false — and the caller takes the QL lane —
// unless min gap > 1e-3 x (largest - smallest root) and everything is finite.  (Kramers pairs at the TRIM points, band
// crossings, diagonal / zero matrices go there; on the 1024 x 1024 Kane-Mele mesh that is ~50 of 10^6 points.)
// Measured on random, clustered and decoupled spectra (tests/hostemu: test_small_direct_solver_n4): eigenvalues to 2e-15 x |T|,
// residuals 1e-15, orthogonality <= 2e-13 at the gap threshold (~ eps / relative gap).
// lam ascending; column c of z = eigenvector of lam[c], as small_tridiag_ql returns them.
// ---------------------------------------------------------------------------
TBK_HD bool small_tridiag4_direct(const double d[4], const double e[4], double lam[4], double z[4][4]) {
  const double mu = 0.25 * ((d[0] + d[1]) + (d[2] + d[3]));
  const double t0 = d[0] - mu, t1 = d[1] - mu, t2 = d[2] - mu, t3 = d[3] - mu;
  const double e0 = e[0], e1 = e[1], e2 = e[2];
  const double f0 = e0 * e0, f1 = e1 * e1, f2 = e2 * e2;
  // x^4 + a2 x^2 + a1 x + a0 (the shift by the mean removes the cubic term)
  const double s01 = t0 * t1, s23 = t2 * t3, u01 = t0 + t1, u23 = t2 + t3;
  const double a2 = s01 + s23 + u01 * u23 - (f0 + f1 + f2);
  const double a1 = f0 * u23 + f1 * (t0 + t3) + f2 * u01 - (s01 * u23 + s23 * u01);
  const double a0 = s01 * s23 - f0 * s23 - f1 * (t0 * t3) - f2 * s01 + f0 * f2;
  // resolvent cubic  zz^3 + 2 a2 zz^2 + (a2^2 - 4 a0) zz - a1^2 = 0, largest root: zz = 2 r cos(theta / 3) - 2 a2 / 3
  const double pp = a2 * a2 * (1.0 / 9.0) + a0 * (4.0 / 3.0);                 // -P/3 of the depressed cubic
  const double qq = a1 * a1 + a2 * a2 * a2 * (2.0 / 27.0) - a2 * a0 * (8.0 / 3.0);   // -Q
  if (!(pp > 1.0e-290)) return false;
  double rs;
  const double r = sqrt_fast(pp, &rs);
  double c = 0.5 * qq * rs * rs * rs;                                          // cos(theta) = -Q / (2 r^3)
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  double y = (double)cosf(acosf((float)c) * (1.0f / 3.0f));                    // in [1/2, 1]
  {
    // one Newton step on 4 y^3 - 3 y = c (denominator 12 y^2 - 3 >= 0, = 0 only at the double root y = 1/2)
    const double den = fma(12.0 * y, y, -3.0);
    if (den > 1.0e-3) y -= (fma(4.0 * y * y, y, -3.0 * y) - c) * rcp_fast(den);
  }
  const double zz = fma(2.0 * r, y, a2 * (-2.0 / 3.0));
  if (!(zz > 1.0e-290)) return false;
  // x^4 + a2 x^2 + a1 x + a0 = (x^2 + s x + tt)(x^2 - s x + uu),  s^2 = zz, tt + uu = a2 + zz, s (uu - tt) = a1
  double rsz;
  const double sgm = sqrt_fast(zz, &rsz);
  const double qd = a1 * rsz;
  const double tt = 0.5 * (a2 + zz - qd), uu = 0.5 * (a2 + zz + qd);
  const double D1 = fma(-4.0, tt, zz), D2 = fma(-4.0, uu, zz);
  const double q1 = sqrt(D1 > 0.0 ? D1 : 0.0), q2 = sqrt(D2 > 0.0 ? D2 : 0.0);
  double x[4] = {0.5 * (-sgm - q1), 0.5 * (-sgm + q1), 0.5 * (sgm - q2), 0.5 * (sgm + q2)};
  // ascending (x[0] <= x[1], x[2] <= x[3] already): three compare-exchanges
  { const double lo = fmin(x[0], x[2]), hi = fmax(x[0], x[2]); x[0] = lo; x[2] = hi; }
  { const double lo = fmin(x[1], x[3]), hi = fmax(x[1], x[3]); x[1] = lo; x[3] = hi; }
  { const double lo = fmin(x[1], x[2]), hi = fmax(x[1], x[2]); x[1] = lo; x[2] = hi; }
  // two Newton steps per root on the Sturm recurrence (p_4 and its derivative).  The size of the SECOND step is the
  // convergence test: at a simple, separated root the first step leaves ~1e-10 and the second is ~1e-18 of the
  // spectrum's width; near a (nearly) multiple root the closed form is only good to ~sqrt(eps), Newton converges
  // linearly, the step stays large — and the roots it would return can be split far enough to pass the gap test below.
  double worst = 0.0;
  TBK_UNROLL
  for (int it = 0; it < 2; ++it) {
    TBK_UNROLL
    for (int b = 0; b < 4; ++b) {
      const double l = x[b];
      const double g0 = t0 - l, g1 = t1 - l, g2 = t2 - l, g3 = t3 - l;
      const double p1 = g0;
      const double p2 = fma(g1, p1, -f0), dp2 = -(p1 + g1);
      const double p3 = fma(g2, p2, -f1 * p1), dp3 = fma(g2, dp2, f1 - p2);
      const double p4 = fma(g3, p3, -f2 * p2), dp4 = fma(g3, dp3, -(p3 + f2 * dp2));
      const double ad = fabs(dp4);
      const double step = ad > 1.0e-290 ? p4 * rcp_fast(dp4) : INFINITY;
      x[b] = l - step;
      if (it == 1) worst = fmax(worst, fabs(step));
    }
  }
  const double spread = x[3] - x[0];
  const double g01 = x[1] - x[0], g12 = x[2] - x[1], g23 = x[3] - x[2];
  const double mingap = fmin(g01, fmin(g12, g23));
  if (!(mingap > 1.0e-3 * spread) || !(spread < 1.0e150) || !(worst < 1.0e-8 * spread)) return false;   // (NaN anywhere: false)
  // eigenvectors: column k of adj(T - lambda), k = argmax |P_k Q_{k+1}| (P: leading, Q: trailing principal minors)
  const double e01 = e0 * e1, e12 = e1 * e2, e012 = e01 * e2;
  TBK_UNROLL
  for (int b = 0; b < 4; ++b) {
    const double l = x[b];
    const double g0 = t0 - l, g1 = t1 - l, g2 = t2 - l, g3 = t3 - l;
    const double p1 = g0, p2 = fma(g1, p1, -f0), p3 = fma(g2, p2, -f1 * p1);
    const double r4 = g3, r3 = fma(g2, r4, -f2), r2 = fma(g1, r3, -f1 * r4);   // trailing minors Q_3, Q_2, Q_1
    const double c0 = r2, c1 = p1 * r3, c2 = p2 * r4, c3 = p3;
    const double m0 = fabs(c0), m1 = fabs(c1), m2 = fabs(c2), m3 = fabs(c3);
    const bool hi = fmax(m2, m3) > fmax(m0, m1);
    const bool odd = hi ? (m3 > m2) : (m1 > m0);
    // candidates: k = 0: ( r2, -e0 r3, e01 r4, -e012 )   k = 1: ( -e0 r3, p1 r3, -p1 e1 r4, p1 e12 )
    //             k = 2: ( e01 r4, -p1 e1 r4, p2 r4, -p2 e2 )   k = 3: ( -e012, p1 e12, -p2 e2, p3 )
    const double v0 = hi ? (odd ? -e012 : e01 * r4) : (odd ? -e0 * r3 : r2);
    const double v1 = hi ? (odd ? p1 * e12 : -p1 * e1 * r4) : (odd ? c1 : -e0 * r3);
    const double v2 = hi ? (odd ? -p2 * e2 : c2) : (odd ? -p1 * e1 * r4 : e01 * r4);
    const double v3 = hi ? (odd ? p3 : -p2 * e2) : (odd ? p1 * e12 : -e012);
    const double n2 = (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
    if (!(n2 > 1.0e-280) || !(n2 < 1.0e280)) return false;
    const double inv = rsqrt_fast(n2);
    z[0][b] = v0 * inv; z[1][b] = v1 * inv; z[2][b] = v2 * inv; z[3][b] = v3 * inv;
    lam[b] = l + mu;
  }
  return true;
}

// the QL lane of the N = 4 solver, out of line (cold): arrays through local memory
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
bool small_tridiag_ql4_cold(double* d, double* e, double* zflat) {
  double dd[4], ee[4], zz[4][4];
  for (int i = 0; i < 4; ++i) { dd[i] = d[i]; ee[i] = e[i]; }
  const bool ok = small_tridiag_ql<4>(dd, ee, zz);
  for (int i = 0; i < 4; ++i) {
    d[i] = dd[i];
    for (int j = 0; j < 4; ++j) zflat[4 * i + j] = zz[i][j];
  }
  return ok;
}

// N = 4 direct solver: Householder, closed-form tridiagonal solver with the QL lane behind it, back-transformation
TBK_HD bool eigh4_direct(cplx a[4][4], double ev[4], cplx w[4][4], int* lane_taken = nullptr) {
  double d[4], e[4];
  cplx tau[4];
  small_hetd2<4>(a, d, e, tau);
  double lam[4], z[4][4];
  bool ok = true;
  const bool fast = small_tridiag4_direct(d, e, lam, z);
  if (lane_taken) *lane_taken = fast ? 0 : 1;
  if (!fast) {
    // (every index below must be a compile-time constant: a rolled loop would index d / z dynamically and move them —
    // and with them the hot path's working set — to local memory)
    double dc[4], ec[4], zc[16];
    TBK_UNROLL
    for (int i = 0; i < 4; ++i) { dc[i] = d[i]; ec[i] = e[i]; }
    ok = small_tridiag_ql4_cold(dc, ec, zc);
    TBK_UNROLL
    for (int i = 0; i < 4; ++i) {
      lam[i] = dc[i];
      TBK_UNROLL
      for (int j = 0; j < 4; ++j) z[i][j] = zc[4 * i + j];
    }
  }
  small_backtransform<4>(a, tau, lam, z, ev, w);
  return ok;
}

// Eigenvalues only, 3 <= N <= 8, one matrix per thread in registers (the lower triangle of `a` is read and destroyed):
// Householder tridiagonalisation + implicit-shift QL without the rotation accumulation.  Replaces the cooperative
// shared-memory solver (8 lanes per matrix, a barrier per phase) for band-structure sweeps of small Wannier models
// (numpy.linalg.eigvalsh at pythtb.py:944): ~2600 SM cycles per n = 8 matrix there, ~25 here.  Returns false if the QL
// iteration did not converge (the caller poisons the eigenvalues, as the larger solvers do).
template <int N>
TBK_HD bool eigvals_small(cplx a[N][N], double ev[N]) {
  double d[N], e[N];
  cplx tau[N];
  small_hetd2<N>(a, d, e, tau);
  double (*nz)[N] = nullptr;
  const bool ok = small_tridiag_ql<N, false>(d, e, nz);
  TBK_UNROLL
  for (int b = 0; b < N; ++b) ev[b] = d[b];
  return ok;
}

// Eigenvalues AND eigenvectors, 3 <= N <= 8, one matrix per thread: the reflectors stay in registers, the N x N real
// rotation matrix of the QL iteration lives in memory (zs: N^2 doubles, `stride` apart — a thread's own column of a
// shared-memory tile), the eigenvectors are back-transformed one at a time and handed to store(b, x) (b-th smallest
// eigenvalue, x[N] its components, NOT conjugated: H x = ev[b] x).  Returns false if the QL iteration did not converge.
template <int N, class Store>
TBK_HD bool eigh_small_mem(cplx a[N][N], double ev[N], double* zs, int stride, Store store) {
  double d[N], e[N];
  cplx tau[N];
  small_hetd2<N>(a, d, e, tau);
  ZMem<N> zm;
  zm.z = zs; zm.stride = stride;
  const bool ok = small_tridiag_ql_t<N>(d, e, zm);
  // the band loop is deliberately NOT unrolled (N copies of the back-transformation and of the caller's store code made
  // the kernel instruction-fetch bound: ncu "no instruction" stalls); the column permutation travels as packed nibbles
  unsigned packed = 0u;
  TBK_UNROLL
  for (int b = 0; b < N; ++b) { ev[b] = d[b]; packed |= (unsigned)zm.perm[b] << (4 * b); }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int b = 0; b < N; ++b) {
    const int col = (int)((packed >> (4 * b)) & 15u);
    cplx x[N];
    TBK_UNROLL
    for (int k = 0; k < N; ++k) x[k] = mk(zs[(k * N + col) * stride], 0.0);
    TBK_UNROLL
    for (int j = N - 2; j >= 0; --j) {
      cplx dot = x[j + 1];                                   // v^H x with v[j+1] = 1
      TBK_UNROLL
      for (int r = j + 2; r < N; ++r) fma_acc_conj(dot, a[r][j], x[r]);
      const cplx f = tau[j] * dot;
      x[j + 1] = x[j + 1] - f;
      TBK_UNROLL
      for (int r = j + 2; r < N; ++r) x[r] = x[r] - f * a[r][j];
    }
    store(b, x);
  }
  return ok;
}

// cold path, kept out of line so that it does not cost the hot kernels registers
template <int N>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
void eigh_small_jacobi_fallback(const double dg_in[N], const cplx lo_in[N * (N - 1) / 2], double ev[N], cplx w[N][N]) {
  double dg[N];
  cplx lo[N * (N - 1) / 2];
  for (int r = 0; r < N; ++r) dg[r] = dg_in[r];
  for (int q = 0; q < N * (N - 1) / 2; ++q) lo[q] = lo_in[q];
  JacobiPacked<N>::solve(dg, lo, w, true);
  for (int r = 0; r < N; ++r) ev[r] = dg[r];
}

// Solver used by the register kernels for N = 3, 4: direct method, Jacobi if QL ever fails to converge.
template <int N>
TBK_HD void eigh_small(const double dg_in[N], const cplx lo_in[N * (N - 1) / 2], double ev[N], cplx w[N][N]) {
  cplx a[N][N];
  TBK_UNROLL
  for (int r = 0; r < N; ++r) {
    TBK_UNROLL
    for (int c = 0; c < N; ++c) a[r][c] = c < r ? lo_in[r * (r - 1) / 2 + c] : mk(c == r ? dg_in[r] : 0.0, 0.0);
  }
  bool ok;
  if constexpr (N == 4) ok = eigh4_direct(a, ev, w);
  else ok = eigh_small_ql<N>(a, ev, w);
  if (!ok) eigh_small_jacobi_fallback<N>(dg_in, lo_in, ev, w);
}

}  // namespace tbk
