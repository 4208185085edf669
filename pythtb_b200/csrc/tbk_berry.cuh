// tbk_berry.cuh — small dense complex kernels behind the Berry-phase engine:
// overlap matrices M = <u_m(k)|u_n(k+b)> (pythtb.py:3793-3817), LU determinant
// phases (np.linalg.det at :3829), the unitary polar factor that the reference
// obtains from an SVD (:3825-3826) and the eigenvalues of the resulting unitary
// Wilson-loop matrix (np.linalg.eigvals at :3834).
//
// "_g" variants are SPMD over a thread group (see tbk_eig_group.cuh); the plain
// variants are serial and are what one-thread-per-link kernels inline.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

// M[m][n] = sum_o conj(a[m][o]) * b[n][o]     (rows are states; _wf_dpr conjugates its first argument)
TBK_HD void overlap_rows(const cplx* a, int lda, const cplx* b, int ldb, int nocc, int n, cplx* M, int ldm) {
  for (int m = 0; m < nocc; ++m)
    for (int q = 0; q < nocc; ++q) {
      cplx acc = mk(0.0, 0.0);
      for (int o = 0; o < n; ++o) fma_acc_conj(acc, a[(size_t)m * lda + o], b[(size_t)q * ldb + o]);
      M[(size_t)m * ldm + q] = acc;
    }
}

template <class G>
TBK_HD void overlap_rows_g(G& g, const cplx* a, int lda, const cplx* b, int ldb, int nocc, int n, cplx* M, int ldm) {
  for (int idx = g.tid(); idx < nocc * nocc; idx += g.size()) {
    const int m = idx / nocc, q = idx - m * nocc;
    cplx acc = mk(0.0, 0.0);
    for (int o = 0; o < n; ++o) fma_acc_conj(acc, a[(size_t)m * lda + o], b[(size_t)q * ldb + o]);
    M[(size_t)m * ldm + q] = acc;
  }
  g.sync();
}

// In-place LU with partial pivoting of a row-major n x n matrix; returns
// det/|det| (or 0 if singular) and log|det|.
TBK_HD cplx lu_det_phase(cplx* M, int n, int ld, double* logabs) {
  cplx u = mk(1.0, 0.0);
  double lg = 0.0;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = norm2(M[(size_t)k * ld + k]);
    for (int r = k + 1; r < n; ++r) {
      const double v = norm2(M[(size_t)r * ld + k]);
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0.0) { *logabs = -INFINITY; return mk(0.0, 0.0); }
    if (piv != k) {
      for (int c = k; c < n; ++c) {
        const cplx t = M[(size_t)k * ld + c];
        M[(size_t)k * ld + c] = M[(size_t)piv * ld + c];
        M[(size_t)piv * ld + c] = t;
      }
      u = -u;
    }
    const cplx p = M[(size_t)k * ld + k];
    const double ap = sqrt(best);
    u = u * mk(p.re / ap, p.im / ap);
    lg += log(ap);
    for (int r = k + 1; r < n; ++r) {
      const cplx f = cdiv(M[(size_t)r * ld + k], p);
      for (int c = k + 1; c < n; ++c) M[(size_t)r * ld + c] = M[(size_t)r * ld + c] - f * M[(size_t)k * ld + c];
    }
  }
  const double nu = sqrt(norm2(u));     // keep |u| = 1 against drift
  *logabs = lg;
  return mk(u.re / nu, u.im / nu);
}

// Group version; ``red`` is a 4-int scratch word array shared by the group.
template <class G>
TBK_HD cplx lu_det_phase_g(G& g, cplx* M, int n, int ld, int* red) {
  cplx u = mk(1.0, 0.0);
  for (int k = 0; k < n; ++k) {
    if (g.tid() == 0) {
      int piv = k;
      double best = norm2(M[(size_t)k * ld + k]);
      for (int r = k + 1; r < n; ++r) {
        const double v = norm2(M[(size_t)r * ld + k]);
        if (v > best) { best = v; piv = r; }
      }
      red[(k & 1) * 2] = piv;
      red[(k & 1) * 2 + 1] = (best == 0.0);
    }
    g.sync();
    const int piv = red[(k & 1) * 2];
    if (red[(k & 1) * 2 + 1]) return mk(0.0, 0.0);
    if (piv != k) {
      for (int c = k + g.tid(); c < n; c += g.size()) {
        const cplx t = M[(size_t)k * ld + c];
        M[(size_t)k * ld + c] = M[(size_t)piv * ld + c];
        M[(size_t)piv * ld + c] = t;
      }
      u = -u;
      g.sync();
    }
    const cplx p = M[(size_t)k * ld + k];
    const double ap = sqrt(norm2(p));
    u = u * mk(p.re / ap, p.im / ap);
    for (int r = k + 1 + g.tid(); r < n; r += g.size()) {
      const cplx f = cdiv(M[(size_t)r * ld + k], p);
      for (int c = k + 1; c < n; ++c) M[(size_t)r * ld + c] = M[(size_t)r * ld + c] - f * M[(size_t)k * ld + c];
    }
    g.sync();
  }
  const double nu = sqrt(norm2(u));
  return mk(u.re / nu, u.im / nu);
}

// C = A * B  (row-major n x n, serial)
TBK_HD void matmul_nn(const cplx* A, const cplx* B, cplx* C, int n) {
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      cplx acc = mk(0.0, 0.0);
      for (int k = 0; k < n; ++k) fma_acc(acc, A[(size_t)r * n + k], B[(size_t)k * n + c]);
      C[(size_t)r * n + c] = acc;
    }
}

// In-place inverse by Gauss-Jordan with partial pivoting (row-major n x n);
// ``w`` is an n x n work matrix.  Returns 0, or 1 if singular.
TBK_HD int invert_gj(cplx* A, int n, cplx* w) {
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) w[(size_t)r * n + c] = mk(r == c ? 1.0 : 0.0, 0.0);
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = norm2(A[(size_t)k * n + k]);
    for (int r = k + 1; r < n; ++r) {
      const double v = norm2(A[(size_t)r * n + k]);
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0.0) return 1;
    if (piv != k)
      for (int c = 0; c < n; ++c) {
        cplx t = A[(size_t)k * n + c]; A[(size_t)k * n + c] = A[(size_t)piv * n + c]; A[(size_t)piv * n + c] = t;
        t = w[(size_t)k * n + c]; w[(size_t)k * n + c] = w[(size_t)piv * n + c]; w[(size_t)piv * n + c] = t;
      }
    const cplx ip = cdiv(mk(1.0, 0.0), A[(size_t)k * n + k]);
    for (int c = 0; c < n; ++c) {
      A[(size_t)k * n + c] = A[(size_t)k * n + c] * ip;
      w[(size_t)k * n + c] = w[(size_t)k * n + c] * ip;
    }
    for (int r = 0; r < n; ++r) {
      if (r == k) continue;
      const cplx f = A[(size_t)r * n + k];
      if (f.re == 0.0 && f.im == 0.0) continue;
      for (int c = 0; c < n; ++c) {
        A[(size_t)r * n + c] = A[(size_t)r * n + c] - f * A[(size_t)k * n + c];
        w[(size_t)r * n + c] = w[(size_t)r * n + c] - f * w[(size_t)k * n + c];
      }
    }
  }
  for (int i = 0; i < n * n; ++i) A[i] = w[i];
  return 0;
}

// Replace M (row-major n x n, nonsingular) by its unitary polar factor
// U = M (M^H M)^{-1/2}  ( = matU @ matV of numpy's SVD, pythtb.py:3825-3826 )
// with the scaled Newton iteration X <- (g X + X^{-H}/g)/2.  w1, w2: n x n work.
// Returns the number of iterations, or -1 if M is singular.
TBK_HD int polar_unitary(cplx* M, int n, cplx* w1, cplx* w2) {
  const int nn = n * n;
  for (int it = 1; it <= 60; ++it) {
    for (int i = 0; i < nn; ++i) w1[i] = M[i];
    if (invert_gj(w1, n, w2)) return -1;         // w1 = X^{-1}
    double nx = 0.0, ni = 0.0;
    for (int i = 0; i < nn; ++i) { nx += norm2(M[i]); ni += norm2(w1[i]); }
    const double gam = sqrt(sqrt(ni / nx));      // (|X^-1|_F / |X|_F)^(1/2)
    double diff = 0.0;
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) {
        const cplx x = M[(size_t)r * n + c];
        const cplx y = conj(w1[(size_t)c * n + r]);          // X^{-H}
        const cplx z = mk(0.5 * (gam * x.re + y.re / gam), 0.5 * (gam * x.im + y.im / gam));
        diff += norm2(z - x);
        w2[(size_t)r * n + c] = z;
      }
    for (int i = 0; i < nn; ++i) M[i] = w2[i];
    if (diff <= 1.0e-30 * (double)n) return it;  // |X_{k+1}-X_k|_F <= 1e-15 |U|_F
  }
  return 60;
}

// Eigenvalues of a general complex matrix (row-major n x n, destroyed):
// Householder reduction to Hessenberg form followed by explicitly shifted
// complex QR with Wilkinson shifts, operating on the active window only.
// Returns 0 on success, >0 = number of eigenvalues not converged.
TBK_HD int comqr_eigvals(cplx* A, int n, int ld, cplx* ev) {
  const double eps = 2.220446049250313e-16;
  // --- Hessenberg reduction
  for (int k = 0; k + 2 < n; ++k) {
    double xn = 0.0;
    for (int r = k + 2; r < n; ++r) xn += norm2(A[(size_t)r * ld + k]);
    const cplx alpha = A[(size_t)(k + 1) * ld + k];
    if (xn == 0.0) continue;
    const double beta = -copysign(sqrt(norm2(alpha) + xn), alpha.re);
    const cplx tau = mk((beta - alpha.re) / beta, -alpha.im / beta);
    const cplx scal = cdiv(mk(1.0, 0.0), mk(alpha.re - beta, alpha.im));
    // v = (1, x*scal) stored in column k below the subdiagonal
    for (int r = k + 2; r < n; ++r) A[(size_t)r * ld + k] = A[(size_t)r * ld + k] * scal;
    A[(size_t)(k + 1) * ld + k] = mk(beta, 0.0);
    // A <- H^H A on rows k+1..n-1, columns k+1..n-1   (H = I - tau v v^H)
    for (int c = k + 1; c < n; ++c) {
      cplx dot = A[(size_t)(k + 1) * ld + c];
      for (int r = k + 2; r < n; ++r) fma_acc_conj(dot, A[(size_t)r * ld + k], A[(size_t)r * ld + c]);
      const cplx f = conj(tau) * dot;
      A[(size_t)(k + 1) * ld + c] = A[(size_t)(k + 1) * ld + c] - f;
      for (int r = k + 2; r < n; ++r) A[(size_t)r * ld + c] = A[(size_t)r * ld + c] - A[(size_t)r * ld + k] * f;
    }
    // A <- A H on all rows, columns k+1..n-1
    for (int r = 0; r < n; ++r) {
      cplx dot = A[(size_t)r * ld + k + 1];
      for (int c = k + 2; c < n; ++c) fma_acc(dot, A[(size_t)r * ld + c], A[(size_t)c * ld + k]);
      const cplx f = dot * tau;
      A[(size_t)r * ld + k + 1] = A[(size_t)r * ld + k + 1] - f;
      for (int c = k + 2; c < n; ++c) A[(size_t)r * ld + c] = A[(size_t)r * ld + c] - mulc(f, A[(size_t)c * ld + k]);
    }
  }
  for (int r = 2; r < n; ++r)
    for (int c = 0; c + 1 < r; ++c) A[(size_t)r * ld + c] = mk(0.0, 0.0);
  // --- shifted QR on the Hessenberg matrix
  int hi = n - 1, iter = 0, bad = 0;
  while (hi >= 0) {
    int l = hi;
    for (; l > 0; --l) {
      const double sub = fabs(A[(size_t)l * ld + l - 1].re) + fabs(A[(size_t)l * ld + l - 1].im);
      double dd = fabs(A[(size_t)(l - 1) * ld + l - 1].re) + fabs(A[(size_t)(l - 1) * ld + l - 1].im) +
                  fabs(A[(size_t)l * ld + l].re) + fabs(A[(size_t)l * ld + l].im);
      if (dd == 0.0) dd = 1.0;
      if (sub <= eps * dd) { A[(size_t)l * ld + l - 1] = mk(0.0, 0.0); break; }
    }
    if (l == hi) { ev[hi] = A[(size_t)hi * ld + hi]; --hi; iter = 0; continue; }
    if (++iter > 60) { ev[hi] = A[(size_t)hi * ld + hi]; --hi; iter = 0; ++bad; continue; }
    // Wilkinson shift from the trailing 2x2 of the window
    cplx mu;
    {
      const cplx a = A[(size_t)(hi - 1) * ld + hi - 1], b = A[(size_t)(hi - 1) * ld + hi];
      const cplx c = A[(size_t)hi * ld + hi - 1], d = A[(size_t)hi * ld + hi];
      if (iter == 10 || iter == 20 || iter == 30) {
        mu = mk(fabs(c.re) + fabs(c.im) + d.re, d.im);    // exceptional shift
      } else {
        const cplx hm = 0.5 * (a - d);
        const cplx disc2 = hm * hm + b * c;
        // principal square root
        const double mod = cabs_(disc2);
        cplx sq = mk(sqrt(0.5 * (mod + disc2.re)), sqrt(0.5 * (mod - disc2.re)));
        if (disc2.im < 0.0) sq.im = -sq.im;
        // pick the root of (x-d)^2 - 2 hm (x-d) - b c closer to d:  x - d = hm -+ sq
        cplx den = hm + sq;
        if (norm2(hm - sq) > norm2(den)) den = hm - sq;
        mu = (den.re == 0.0 && den.im == 0.0) ? d : d - cdiv(b * c, den);
      }
    }
    for (int i = l; i <= hi; ++i) A[(size_t)i * ld + i] = A[(size_t)i * ld + i] - mu;
    // QR sweep: rows (G_k from the left) then columns (G_k^H from the right), window l..hi
    // the rotations are stored in ev[l..hi-1] (c in .re of ev is not enough: keep c,s in two slots)
    // -> apply left rotations first, stash (c, s) in the now-zero subdiagonal + a local pair chain
    {
      // left pass
      for (int k = l; k < hi; ++k) {
        const cplx a = A[(size_t)k * ld + k], b = A[(size_t)(k + 1) * ld + k];
        const double na = cabs_(a), nb = cabs_(b);
        double cs; cplx sn;
        if (nb == 0.0) { cs = 1.0; sn = mk(0.0, 0.0); }
        else if (na == 0.0) { cs = 0.0; sn = mk(b.re / nb, -b.im / nb); }
        else {
          const double rho = hypot(na, nb);
          cs = na / rho;
          const cplx ua = mk(a.re / na, a.im / na);
          sn = mulc(ua, b) * (1.0 / rho);              // (a/|a|) conj(b) / rho
        }
        for (int c = k; c <= hi; ++c) {
          const cplx x = A[(size_t)k * ld + c], y = A[(size_t)(k + 1) * ld + c];
          A[(size_t)k * ld + c] = cs * x + sn * y;
          A[(size_t)(k + 1) * ld + c] = cs * y - cmul(sn, x);
        }
        // stash the rotation in the (now zero) subdiagonal slot and in ev[k]
        A[(size_t)(k + 1) * ld + k] = mk(cs, 0.0);
        ev[k] = sn;
      }
      // right pass: columns k, k+1 <- [x, y] G_k^H ;  x' = c x + conj(s) y ; y' = -s x + c y
      for (int k = l; k < hi; ++k) {
        const double cs = A[(size_t)(k + 1) * ld + k].re;
        const cplx sn = ev[k];
        A[(size_t)(k + 1) * ld + k] = mk(0.0, 0.0);
        const int rmax = (k + 1 <= hi) ? k + 1 : hi;
        for (int r = l; r <= rmax; ++r) {
          const cplx x = A[(size_t)r * ld + k], y = A[(size_t)r * ld + k + 1];
          A[(size_t)r * ld + k] = cs * x + cmul(sn, y);
          A[(size_t)r * ld + k + 1] = cs * y - sn * x;
        }
      }
    }
    for (int i = l; i <= hi; ++i) A[(size_t)i * ld + i] = A[(size_t)i * ld + i] + mu;
  }
  return bad;
}

}  // namespace tbk
