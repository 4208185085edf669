// tbk_eig_group.cuh — one Hermitian matrix per cooperating thread group
// (sub-warp tile, warp, or whole CTA): Householder tridiagonalisation, in-place
// accumulation of Q, implicit-shift QL on the real tridiagonal with the plane
// rotations applied to Q.  Replaces numpy.linalg.eigh/eigvalsh
// (pythtb.py:939/944, 2247/2252) for 5 <= n <= 512.
//
// The code is SPMD over an abstract group ``G``:
//     int  g.tid()        rank of this thread in the group
//     int  g.size()       number of threads in the group
//     void g.sync()       barrier + memory fence for the group
//     double g.sum(x)     all-reduce (every thread gets the sum)
// so the same source runs as a tile of a warp, as a CTA, or — for the CPU
// unit tests of the arithmetic — as a "group" of one host thread.
//
// Matrix storage: column-major A(r,c) = A[r + c*lda] (complex), odd ``lda``
// recommended (conflict-free shared-memory access both along rows and along
// columns with 16-byte elements).  Only the LOWER triangle is read on entry
// (numpy's default UPLO='L').
#pragma once
#include "tbk_common.cuh"

namespace tbk {

struct EigScratch {
  double* d;      // [n]   diagonal of T, then eigenvalues
  double* e;      // [n]   sub-diagonal of T (e[j] couples j, j+1)
  cplx* tau;      // [n]   Householder scalars
  cplx* work;     // [n]
  double* rc;     // [n]   rotation cosines of the current QL sweep
  double* rs;     // [n]   rotation sines
  int* ctl;       // [8]   control words shared by the group (double-buffered)
};

TBK_HD size_t eig_scratch_bytes(int n) {
  // d,e,rc,rs (8n each) + tau,work (16n each) + ctl
  return (size_t)n * (4 * 8 + 2 * 16) + 32;
}

TBK_HD EigScratch eig_scratch_carve(void* base, int n) {
  EigScratch s;
  char* p = (char*)base;
  s.tau = (cplx*)p;  p += (size_t)n * 16;
  s.work = (cplx*)p; p += (size_t)n * 16;
  s.d = (double*)p;  p += (size_t)n * 8;
  s.e = (double*)p;  p += (size_t)n * 8;
  s.rc = (double*)p; p += (size_t)n * 8;
  s.rs = (double*)p; p += (size_t)n * 8;
  s.ctl = (int*)p;
  return s;
}

// Fill the strict upper triangle from the lower one and force a real diagonal.
template <class G>
TBK_HD void herm_fill_upper(G& g, int n, cplx* A, int lda) {
  for (int c = g.tid(); c < n; c += g.size()) {
    A[c + (size_t)c * lda].im = 0.0;
    for (int r = c + 1; r < n; ++r) A[c + (size_t)r * lda] = conj(A[r + (size_t)c * lda]);
  }
  g.sync();
}

// Reduce the full Hermitian A to real symmetric tridiagonal T = Q^H A Q.
// On exit: d,e hold T; the Householder vectors are stored below the first
// sub-diagonal of A (LAPACK zhetd2 'L' layout) with scalars in tau.
template <class G>
TBK_HD void hetrd(G& g, int n, cplx* A, int lda, EigScratch& s) {
  for (int j = 0; j < n - 1; ++j) {
    cplx* col = A + (size_t)j * lda;          // column j
    // --- generate the reflector for x = A(j+1:n, j)  (zlarfg)
    double part = 0.0;
    for (int r = j + 2 + g.tid(); r < n; r += g.size()) part += norm2(col[r]);
    const double xnorm2 = g.sum(part);
    const cplx alpha = col[j + 1];
    cplx tau = mk(0.0, 0.0);
    double beta = alpha.re;
    if (xnorm2 != 0.0 || alpha.im != 0.0) {
      beta = -copysign(sqrt(alpha.re * alpha.re + alpha.im * alpha.im + xnorm2), alpha.re);
      tau = mk((beta - alpha.re) / beta, -alpha.im / beta);
      const cplx scal = cdiv(mk(1.0, 0.0), mk(alpha.re - beta, alpha.im));
      g.sync();                                // everyone has read alpha
      for (int r = j + 2 + g.tid(); r < n; r += g.size()) col[r] = col[r] * scal;
    } else {
      g.sync();
    }
    if (g.tid() == 0) {
      col[j + 1] = mk(1.0, 0.0);
      s.e[j] = beta;
      s.tau[j] = tau;
      s.d[j] = col[j].re;
    }
    g.sync();
    if (tau.re != 0.0 || tau.im != 0.0) {
      // --- p = tau * A22 * v   (rows distributed over the group)
      for (int r = j + 1 + g.tid(); r < n; r += g.size()) {
        cplx acc = mk(0.0, 0.0);
        for (int c = j + 1; c < n; ++c) fma_acc(acc, A[r + (size_t)c * lda], col[c]);
        s.work[r] = tau * acc;
      }
      g.sync();
      // --- w = p - (tau/2) (p^H v) v
      double dre = 0.0, dim = 0.0;
      for (int r = j + 1 + g.tid(); r < n; r += g.size()) {
        const cplx t = cmul(s.work[r], col[r]);
        dre += t.re; dim += t.im;
      }
      dre = g.sum(dre); dim = g.sum(dim);
      const cplx a2 = (-0.5) * (tau * mk(dre, dim));
      g.sync();
      for (int r = j + 1 + g.tid(); r < n; r += g.size()) s.work[r] = s.work[r] + a2 * col[r];
      g.sync();
      // --- A22 -= v w^H + w v^H   (full block, rows distributed)
      for (int r = j + 1 + g.tid(); r < n; r += g.size()) {
        const cplx vr = col[r], wr = s.work[r];
        for (int c = j + 1; c < n; ++c) {
          cplx a = A[r + (size_t)c * lda];
          a = a - mulc(vr, s.work[c]) - mulc(wr, col[c]);
          A[r + (size_t)c * lda] = a;
        }
      }
      g.sync();
    }
  }
  if (g.tid() == 0) {
    s.d[n - 1] = A[(n - 1) + (size_t)(n - 1) * lda].re;
    s.e[n - 1] = 0.0;
    s.tau[n - 1] = mk(0.0, 0.0);
  }
  g.sync();
}

// Overwrite A with Q = H(0) H(1) ... H(n-2) (LAPACK zungtr 'L' / zung2r).
template <class G>
TBK_HD void ungtr(G& g, int n, cplx* A, int lda, EigScratch& s) {
  // shift reflector j from column j to column j+1; first row/column = unit
  for (int c = n - 1; c >= 1; --c) {
    for (int r = c + 1 + g.tid(); r < n; r += g.size()) A[r + (size_t)c * lda] = A[r + (size_t)(c - 1) * lda];
    g.sync();
  }
  for (int r = g.tid(); r < n; r += g.size()) {
    A[r] = mk(r == 0 ? 1.0 : 0.0, 0.0);                // column 0
    if (r > 0) A[(size_t)r * lda] = mk(0.0, 0.0);      // row 0
  }
  g.sync();
  // zung2r on the trailing (n-1)x(n-1) block B(r,c) = A(r+1,c+1), k = n-1
  const int m = n - 1;
  cplx* B = A + 1 + lda;
  for (int i = m - 1; i >= 0; --i) {
    cplx* v = B + (size_t)i * lda;              // reflector i lives in column i, rows i..m-1
    const cplx tau = s.tau[i];
    if (i < m - 1) {
      if (g.tid() == 0) v[i] = mk(1.0, 0.0);
      g.sync();
      // apply H(i) = I - tau v v^H to B(i:m, i+1:m) from the left, one column per thread
      for (int c = i + 1 + g.tid(); c < m; c += g.size()) {
        cplx* bc = B + (size_t)c * lda;
        cplx dot = mk(0.0, 0.0);
        for (int r = i; r < m; ++r) fma_acc_conj(dot, v[r], bc[r]);
        const cplx f = tau * dot;
        for (int r = i; r < m; ++r) bc[r] = bc[r] - v[r] * f;
      }
      g.sync();
      for (int r = i + 1 + g.tid(); r < m; r += g.size()) v[r] = (-1.0) * (tau * v[r]);
    }
    if (g.tid() == 0) v[i] = mk(1.0 - tau.re, -tau.im);
    for (int r = g.tid(); r < i; r += g.size()) v[r] = mk(0.0, 0.0);
    g.sync();
  }
}

// Implicit-shift QL on (d, e); if Z != nullptr the rotations are applied to the
// columns of Z (n x n, column-major, leading dimension ldz), turning Q into the
// eigenvector matrix.  Returns 0, or l+1 if eigenvalue l failed to converge.
template <class G>
TBK_HD int tql_implicit(G& g, int n, cplx* Z, int ldz, EigScratch& s) {
  const double eps = 1.1102230246251565e-16;
  int fail = 0;
  int pass = 0;
  for (int l = 0; l < n; ++l) {
    int iter = 0;
    while (true) {
      // control words are double-buffered: thread 0 may already be writing the
      // next pass while slower threads still read this one
      int* ctl = s.ctl + 4 * (pass & 1);
      ++pass;
      // ---- scalar part: one thread advances the tridiagonal and records the rotations
      if (g.tid() == 0) {
        double* d = s.d; double* e = s.e;
        int m = l;
        for (; m < n - 1; ++m) {
          const double dd = fabs(d[m]) + fabs(d[m + 1]);
          if (fabs(e[m]) <= eps * dd) break;
        }
        int nrot = 0, lo = l;
        if (m != l) {
          double gg = (d[l + 1] - d[l]) / (2.0 * e[l]);
          double r = hypot(gg, 1.0);
          gg = d[m] - d[l] + e[l] / (gg + copysign(r, gg));
          double sn = 1.0, cs = 1.0, p = 0.0;
          int i = m - 1;
          bool early = false;
          for (; i >= l; --i) {
            double f = sn * e[i];
            const double b = cs * e[i];
            r = hypot(f, gg);
            e[i + 1] = r;
            if (r == 0.0) {
              d[i + 1] -= p;
              e[m] = 0.0;
              early = true;
              break;
            }
            sn = f / r;
            cs = gg / r;
            gg = d[i + 1] - p;
            r = (d[i] - gg) * sn + 2.0 * cs * b;
            p = sn * r;
            d[i + 1] = gg + p;
            gg = cs * r - b;
            s.rc[i] = cs;
            s.rs[i] = sn;
          }
          if (!early) {
            d[l] -= p;
            e[l] = gg;
            e[m] = 0.0;
            lo = l;
          } else {
            lo = i + 1;       // rotations recorded for indices m-1 .. i+1
          }
          nrot = m - lo;      // indices lo .. m-1
        }
        ctl[0] = m;
        ctl[1] = nrot;
        ctl[2] = lo;
      }
      g.sync();
      const int m = ctl[0], nrot = ctl[1], lo = ctl[2];
      if (m == l) break;
      if (++iter > 80) { fail = l + 1; break; }
      // ---- vector part: every thread rotates its rows of Z
      if (Z != nullptr && nrot > 0) {
        for (int k = g.tid(); k < n; k += g.size()) {
          cplx hi = Z[k + (size_t)m * ldz];                 // Z(k, i+1), carried
          for (int i = m - 1; i >= lo; --i) {
            const double cs = s.rc[i], sn = s.rs[i];
            const cplx zi = Z[k + (size_t)i * ldz];
            Z[k + (size_t)(i + 1) * ldz] = mk(fma(sn, zi.re, cs * hi.re), fma(sn, zi.im, cs * hi.im));
            hi = mk(fma(cs, zi.re, -sn * hi.re), fma(cs, zi.im, -sn * hi.im));
          }
          Z[k + (size_t)lo * ldz] = hi;
        }
      }
      g.sync();
    }
    if (fail) break;
  }
  g.sync();
  return fail;
}

// rank[i] = position of eigenvalue i in ascending order (stable).
template <class G>
TBK_HD void eig_rank(G& g, int n, const double* d, int* rank) {
  for (int i = g.tid(); i < n; i += g.size()) {
    const double di = d[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (d[j] < di) || (d[j] == di && j < i);
    rank[i] = r;
  }
  g.sync();
}

// Full driver.  A: n x n column-major (lower triangle valid).  On exit
// s.d holds the (unsorted) eigenvalues and, if want_vec, column i of A is the
// eigenvector of s.d[i].  Use eig_rank for the ascending permutation.
template <class G>
TBK_HD int heev_group(G& g, int n, cplx* A, int lda, EigScratch& s, bool want_vec) {
  if (n == 1) {
    if (g.tid() == 0) { s.d[0] = A[0].re; A[0] = mk(1.0, 0.0); }
    g.sync();
    return 0;
  }
  herm_fill_upper(g, n, A, lda);
  hetrd(g, n, A, lda, s);
  if (want_vec) ungtr(g, n, A, lda, s);
  return tql_implicit(g, n, want_vec ? A : (cplx*)nullptr, lda, s);
}

}  // namespace tbk
