// tbk_eig_blocked.cuh — Hermitian eigensolver for the larger supercells / Wannier models
// (33 <= n <= 512): one matrix per CTA, the matrix itself in an L2/HBM workspace, everything
// O(n) in shared memory.  Replaces numpy.linalg.eigh/eigvalsh (pythtb.py:939/944) for the
// BASELINE configs 4 and 5 (ribbons with norb 200-400, the norb-499 slab).
//
// Stages (each one O(n^3) pass over the matrix at most):
//   1. hetrd_blocked   Householder tridiagonalisation in panels of nb columns (LAPACK zhetrd /
//                      zlatrd, lower): per column ONE matrix-vector product with the stored
//                      trailing matrix plus O(n nb) panel corrections, and one rank-2nb update
//                      of the trailing matrix per panel (panels V, W live in shared memory).
//                      Memory traffic ~ (16/3) n^3 bytes instead of ~16 n^3 for the unblocked
//                      column-by-column update.
//   2. tridiag_bisect  all eigenvalues of the real symmetric tridiagonal by Sturm-sequence
//                      bisection, one eigenvalue per thread (embarrassingly parallel; the QL
//                      iteration it replaces is a serial chain of n^2 rotations).
//   3. tridiag_invit   eigenvectors of the tridiagonal by inverse iteration (LAPACK dstein's
//                      scheme: partially pivoted LU of T - lambda I, perturbed pivots, close
//                      eigenvalues separated by 10 eps |T|), one eigenvector per thread, then
//                      modified Gram-Schmidt inside clusters of close eigenvalues, one cluster
//                      per warp.  O(n^2) per matrix instead of 6 n^3 for rotations applied to Q.
//   4. backtransform   x = H_0 H_1 ... H_{n-2} z  for every eigenvector, one column per warp held
//                      in registers, the reflectors streamed from L2 (8 n^3 flops).
// A cluster vector that loses its norm in the Gram-Schmidt step is regenerated (new random start,
// orthogonalised before and after every solve, as dstein does); should that fail too, the matrix is
// reported through the return code and re-solved by the unblocked Householder + implicit-QL solver
// (tbk_eig_group.cuh).
//
// The code is SPMD over the same abstract group as tbk_eig_group.cuh, extended by sub-teams:
//     g.nsub(), g.sub()        number of sub-teams (warps) and this thread's
//     g.lane(), g.subsize()    rank inside the sub-team and its size
//     g.subsum(x)              all-reduce inside the sub-team (no block barrier)
//     g.subsync()              barrier + memory fence for the sub-team
// so that the identical source runs as a CTA on the GPU and as a "group" of one host thread in
// the CPU unit tests (tests/hostemu).
#pragma once
#include "tbk_common.cuh"

namespace tbk {

constexpr int kBlkMaxN = 512;
constexpr int kWarpCluster = 6;       // clusters up to this size are orthogonalised by one sub-team, larger ones by the group

// leading dimension of the shared-memory panels V, W: the smallest value >= n that is 2 mod 8, so that the 16-byte
// fragment reads of the tensor-pipe rank-2nb update (lane -> row g of panel column q) hit eight distinct bank groups
TBK_HD int blk_ldp(int n) { return ((n + 5) / 8) * 8 + 2; }

struct BlkWork {
  int n, lda, nb, ldp;
  cplx* A;         // [n x lda] column-major, lower triangle valid on entry (global)
  cplx* V;         // [nb][n] reflector panel            (shared memory on the device)
  cplx* W;         // [nb][n] zlatrd's W panel           (shared)
  cplx* wcol;      // [n] column part of the symmetric matrix-vector product          (shared)
  cplx* racc;      // [nred][n] row partials of the sub-teams, nred sub-teams per round (shared)
  int nred;
  cplx* dots;      // [2 nb] panel dot products           (shared)
  cplx* tau;       // [n]                                 (shared)
  double* d;       // [n] diagonal of T                   (shared)
  double* e;       // [n] sub-diagonal of T               (shared)
  double* e2;      // [n] e^2                             (shared)
  double* lam;     // [n] eigenvalues, ascending          (shared)
  double* lamp;    // [n] perturbed eigenvalues used by the inverse iteration (shared)
  int* cl;         // [n] first index of the cluster each eigenvalue belongs to (shared)
  int* ctl;        // [4] status words                    (shared)
  double* Z;       // [n][n] row-major: Z[i*n + j] = component i of tridiagonal eigenvector j (global)
  double* lu;      // [4][n][nt] interleaved per-thread LU rows, nt = min(group size, n rounded up) (global)
  int nt;
  unsigned long long* prof;   // -DTBK_HETRD_PROF builds only: cycle counters of the tridiagonalisation's phases
};

// phase counters of the tridiagonalisation (profiles/build_variant.py hetrd_prof -DTBK_HETRD_PROF=1): slot 0 = the
// matrix-vector products, 3 = the rank-2nb updates, 7 = everything else (column update, reflector, corrections)
#if defined(TBK_HETRD_PROF) && defined(__CUDA_ARCH__)
#define TBK_HP_BEGIN long long hp_t0 = (w.prof && g.tid() == 0) ? clock64() : 0; long long hp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TBK_HP(slot) if (w.prof && g.tid() == 0) { const long long hp_t1 = clock64(); hp_acc[slot] += hp_t1 - hp_t0; hp_t0 = hp_t1; }
#define TBK_HP_END if (w.prof && g.tid() == 0) { atomicAdd(w.prof + 0, (unsigned long long)hp_acc[0]); atomicAdd(w.prof + 3, (unsigned long long)hp_acc[3]); atomicAdd(w.prof + 7, (unsigned long long)hp_acc[7]); }
#else
#define TBK_HP_BEGIN
#define TBK_HP(slot)
#define TBK_HP_END
#endif

// wcol + racc: (1 + nred) n complex numbers for the lower-triangle product, and never fewer than `nthreads` (the
// full-matrix variant uses the same region as its per-thread split-product scratch; it needs no racc: nred = 0)
TBK_HD size_t blk_scratch_elems(int n, int nred, int nthreads) {
  const size_t a = (size_t)(1 + nred) * n;
  return a > (size_t)nthreads ? a : (size_t)nthreads;
}
TBK_HD size_t blk_shared_bytes(int n, int nb, int nred, int nthreads) {
  return (size_t)2 * nb * blk_ldp(n) * 16 + blk_scratch_elems(n, nred, nthreads) * 16 + (size_t)2 * nb * 16 + (size_t)n * 16 +
         (size_t)5 * n * 8 + (size_t)n * 4 + 64;
}

// carve the shared part of a BlkWork (n, nb, nred set) out of one 16-byte aligned buffer
TBK_HD void blk_carve_shared(BlkWork& w, void* base, int nthreads) {
  char* p = (char*)base;
  const int n = w.n, nb = w.nb;
  w.ldp = blk_ldp(n);
  w.V = (cplx*)p;    p += (size_t)nb * w.ldp * 16;
  w.W = (cplx*)p;    p += (size_t)nb * w.ldp * 16;
  w.wcol = (cplx*)p;
  w.racc = w.wcol + n;
  p += blk_scratch_elems(n, w.nred, nthreads) * 16;
  w.dots = (cplx*)p; p += (size_t)2 * nb * 16;
  w.tau = (cplx*)p;  p += (size_t)n * 16;
  w.d = (double*)p;  p += (size_t)n * 8;
  w.e = (double*)p;  p += (size_t)n * 8;
  w.e2 = (double*)p; p += (size_t)n * 8;
  w.lam = (double*)p;  p += (size_t)n * 8;
  w.lamp = (double*)p; p += (size_t)n * 8;
  w.cl = (int*)p;    p += (size_t)n * 4;
  w.ctl = (int*)p;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void blk_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// Rank-2nb update of the full trailing matrix on the FP64 tensor pipe (all threads of the CTA):
//   A[r][c] -= sum_kk P[r][kk] conj(Q[c][kk]),  P = [V | W], Q = [W | V]  (K = 2 nb),  j1 <= r, c < n.
// A warp owns 16 x 16 tiles (2 x 2 fragments of mma.sync.m8n8k4.f64, fragment layouts in tbk_eig_wy.cuh); the old
// values of a tile are requested from L2 / HBM first and the DMMAs run while they are on their way.  Operands come
// straight from the shared-memory panels ([k][ldp], ldp == 2 mod 8: conflict-free 16-byte fragment reads); row / column
// indices past n - 1 are clamped for the operand reads and masked for the tile itself.  The scalar version of this
// update (one element per thread: 16-byte load -> 64 FMAs -> store) ran at ~9 % of the FP64 rate.
template <bool LOWER>
__device__ __forceinline__ void blk_rank2k_dmma(cplx* __restrict__ A, int lda, int n, int j1, const cplx* __restrict__ V,
                                                const cplx* __restrict__ W, int ldp, int nb) {
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int m = n - j1;
  const int nt = (m + 15) >> 4;
  const int ksteps = nb >> 1;                         // 2 nb / 4
  // LOWER: only the tiles on and below the diagonal, and of those only the elements r >= c (tile column tc has nt - tc tiles)
  const int ntile = LOWER ? nt * (nt + 1) / 2 : nt * nt;
  for (int t = warp; t < ntile; t += nwarp) {
    int tc, tr;
    if (LOWER) {
      // t = tc nt - tc (tc - 1) / 2 + (tr - tc): invert by a float estimate and one correction either way
      tc = (int)((2.0f * nt + 1.0f - sqrtf((2.0f * nt + 1.0f) * (2.0f * nt + 1.0f) - 8.0f * (float)t)) * 0.5f);
      if (tc < 0) tc = 0;
      while (tc > 0 && tc * nt - tc * (tc - 1) / 2 > t) --tc;
      while ((tc + 1) * nt - (tc + 1) * tc / 2 <= t) ++tc;
      tr = tc + (t - (tc * nt - tc * (tc - 1) / 2));
    } else {
      tc = t / nt; tr = t - tc * nt;
    }
    const int r0 = j1 + tr * 16, c0 = j1 + tc * 16;
    cplx old[2][2][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = r0 + rt * 8 + g, c = c0 + ct * 8 + 2 * q + e;
          old[rt][ct][e] = (r < n && c < n && (!LOWER || r >= c)) ? A[r + (size_t)c * lda] : mk(0.0, 0.0);
        }
    double are[2][2][2], aim[2][2][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) { are[rt][ct][0] = are[rt][ct][1] = 0.0; aim[rt][ct][0] = aim[rt][ct][1] = 0.0; }
    int rr[2], cc[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      rr[x] = r0 + x * 8 + g < n ? r0 + x * 8 + g : n - 1;
      cc[x] = c0 + x * 8 + g < n ? c0 + x * 8 + g : n - 1;
    }
    for (int ks = 0; ks < ksteps; ++ks) {
      const int kk = ks * 4 + q;
      const cplx* Pp = kk < nb ? V + kk * ldp : W + (kk - nb) * ldp;
      const cplx* Qp = kk < nb ? W + kk * ldp : V + (kk - nb) * ldp;
      const cplx p0 = Pp[rr[0]], p1 = Pp[rr[1]], q0 = Qp[cc[0]], q1 = Qp[cc[1]];
#pragma unroll
      for (int rt = 0; rt < 2; ++rt) {
        const cplx pv = rt ? p1 : p0;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          const cplx qv = ct ? q1 : q0;             // p conj(q) = (pr qr + pi qi) + i (pi qr - pr qi)
          blk_dmma(are[rt][ct][0], are[rt][ct][1], pv.re, qv.re);
          blk_dmma(are[rt][ct][0], are[rt][ct][1], pv.im, qv.im);
          blk_dmma(aim[rt][ct][0], aim[rt][ct][1], pv.im, qv.re);
          blk_dmma(aim[rt][ct][0], aim[rt][ct][1], -pv.re, qv.im);
        }
      }
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = r0 + rt * 8 + g, c = c0 + ct * 8 + 2 * q + e;
          if (r < n && c < n && (!LOWER || r >= c))
            A[r + (size_t)c * lda] = mk(old[rt][ct][e].re - are[rt][ct][e], old[rt][ct][e].im - aim[rt][ct][e]);
        }
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// 1. blocked tridiagonalisation.  On exit d, e hold T, the Householder vectors are stored below the
// first sub-diagonal of A (zhetd2 'L' layout, implicit unit at row j+1) with their scalars in tau.
// Only the lower triangle of A is ever read or written: the trailing matrix is streamed from L2 / HBM once per
// column, and the Hermitian product w = A22 v takes both contributions of an element a_rc (to w_r and,
// conjugated, to w_c) from ONE load — half the bytes of a product with the full matrix.  MAXM >= ceil(n / subsize).
// ---------------------------------------------------------------------------------------------
template <int MAXM, class G>
TBK_HD void hetrd_blocked(G& g, const BlkWork& w) {
  const int n = w.n, lda = w.lda, nb = w.nb, ldp = w.ldp;
  const int T = g.size(), tid = g.tid();
  cplx* A = w.A;
  cplx* V = w.V;
  cplx* W = w.W;
  TBK_HP_BEGIN
  for (int c = tid; c < n; c += T) A[c + (size_t)c * lda].im = 0.0;       // real diagonal
  g.sync();
  for (int j0 = 0; j0 < n - 1; j0 += nb) {
    const int nbp = n - 1 - j0 < nb ? n - 1 - j0 : nb;
    // zero the panels (unused panel columns must be exactly zero for the rank-2nb update)
    for (int q = tid; q < nb * ldp; q += T) { V[q] = mk(0.0, 0.0); W[q] = mk(0.0, 0.0); }
    g.sync();
    for (int i = 0; i < nbp; ++i) {
      const int j = j0 + i;
      cplx* col = A + (size_t)j * lda;
      // ---- (1) bring column j up to date with the reflectors of this panel
      if (i > 0) {
        for (int r = j + tid; r < n; r += T) {
          cplx a = col[r];
          for (int k = 0; k < i; ++k) {
            a = a - mulc(V[k * ldp + r], W[k * ldp + j]);
            a = a - mulc(W[k * ldp + r], V[k * ldp + j]);
          }
          if (r == j) a.im = 0.0;
          col[r] = a;
        }
        g.sync();
      }
      // ---- (2) reflector for x = A(j+1:n, j)   (zlarfg)
      double part = 0.0;
      for (int r = j + 2 + tid; r < n; r += T) part += norm2(col[r]);
      const double xnorm2 = g.sum(part);
      const cplx alpha = col[j + 1];
      cplx tau = mk(0.0, 0.0);
      double beta = alpha.re;
      cplx scal = mk(0.0, 0.0);
      if (xnorm2 != 0.0 || alpha.im != 0.0) {
        beta = -copysign(sqrt(alpha.re * alpha.re + alpha.im * alpha.im + xnorm2), alpha.re);
        tau = mk((beta - alpha.re) / beta, -alpha.im / beta);
        scal = cdiv(mk(1.0, 0.0), mk(alpha.re - beta, alpha.im));
      }
      g.sync();                                   // everyone has read alpha
      cplx* v = V + i * ldp;
      for (int r = j + 2 + tid; r < n; r += T) {
        const cplx x = col[r] * scal;             // tau == 0: the column is already zero below j+1
        col[r] = x;
        v[r] = x;
      }
      if (tid == 0) {
        v[j + 1] = mk(1.0, 0.0);
        w.e[j] = beta;
        w.tau[j] = tau;
        w.d[j] = col[j].re;
      }
      g.sync();
      TBK_HP(7)
      // ---- (3) w = A22 v with the stored (panel-start) trailing matrix, rows/cols j+1 .. n-1, LOWER TRIANGLE ONLY.
      // A sub-team (warp) owns the columns c = j+1+sub, j+1+sub+nsub, ...; it streams column c from the diagonal
      // down (contiguous: 512-byte warp loads), lane L holding the rows r = L + S t.  An element a_rc gives
      //   w_r += a_rc v_c            (row part, accumulated in the lane's registers over all its columns) and
      //   w_c += conj(a_rc) v_r      (column part, one reduction over the sub-team per column).
      // The sub-teams' row partials are then summed through shared memory, nred sub-teams per round.
      {
        const int S = g.subsize(), L = g.lane(), nsub = g.nsub(), sub = g.sub();
        cplx acc[MAXM];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < MAXM; ++t) acc[t] = mk(0.0, 0.0);
        for (int c = j + 1 + sub; c < n; c += nsub) {
          const cplx vc = v[c];
          const cplx* acol = A + (size_t)c * lda;
          double sre = 0.0, sim = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
          for (int t = 0; t < MAXM; ++t) {
            if (S * t + S - 1 >= c) {                 // sub-team-uniform: this slot reaches the diagonal of column c
              const int r = L + S * t;
              if (r >= c && r < n) {
                cplx a = acol[r];
                if (r == c) a.im = 0.0;
                fma_acc(acc[t], a, vc);
                if (r > c) {
                  const cplx tt = cmul(a, v[r]);      // conj(a_rc) v_r
                  sre += tt.re; sim += tt.im;
                }
              }
            }
          }
          sre = g.subsum(sre); sim = g.subsum(sim);
          if (L == 0) w.wcol[c] = mk(sre, sim);
        }
        g.sync();                                     // wcol complete
        for (int w0 = 0; w0 < nsub; w0 += w.nred) {
          if (sub >= w0 && sub < w0 + w.nred) {
            cplx* dst = w.racc + (size_t)(sub - w0) * n;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < MAXM; ++t) {
              const int r = L + S * t;
              if (r > j && r < n) dst[r] = acc[t];
            }
          }
          g.sync();
          const int cnt = nsub - w0 < w.nred ? nsub - w0 : w.nred;
          for (int r = j + 1 + tid; r < n; r += T) {
            cplx sacc = w0 == 0 ? w.wcol[r] : W[i * ldp + r];
            for (int q = 0; q < cnt; ++q) sacc = sacc + w.racc[(size_t)q * n + r];
            W[i * ldp + r] = sacc;
          }
          g.sync();
        }
      }
      TBK_HP(0)
      // ---- panel corrections: dots[k] = W_k^H v, dots[nb+k] = V_k^H v, one sub-team per dot product
      if (i > 0) {
        for (int q = g.sub(); q < 2 * i; q += g.nsub()) {
          const cplx* src = q < i ? W + q * ldp : V + (q - i) * ldp;
          double sre = 0.0, sim = 0.0;
          for (int r = j + 1 + g.lane(); r < n; r += g.subsize()) {
            const cplx t = cmul(src[r], v[r]);
            sre += t.re; sim += t.im;
          }
          sre = g.subsum(sre); sim = g.subsum(sim);
          if (g.lane() == 0) w.dots[q < i ? q : nb + (q - i)] = mk(sre, sim);
        }
        g.sync();
        for (int r = j + 1 + tid; r < n; r += T) {
          cplx acc = W[i * ldp + r];
          for (int k = 0; k < i; ++k) {
            acc = acc - V[k * ldp + r] * w.dots[k];
            acc = acc - W[k * ldp + r] * w.dots[nb + k];
          }
          W[i * ldp + r] = acc;
        }
        g.sync();
      }
      // ---- w = tau w;  w += (-tau/2 (w^H v)) v
      double dre = 0.0, dim = 0.0;
      for (int r = j + 1 + tid; r < n; r += T) {
        const cplx wr = tau * W[i * ldp + r];
        W[i * ldp + r] = wr;
        const cplx t = cmul(wr, v[r]);
        dre += t.re; dim += t.im;
      }
      dre = g.sum(dre); dim = g.sum(dim);        // g.sum synchronises: the scaled w is visible
      const cplx a2 = (-0.5) * (tau * mk(dre, dim));
      for (int r = j + 1 + tid; r < n; r += T) W[i * ldp + r] = W[i * ldp + r] + a2 * v[r];
      g.sync();
    }
    TBK_HP(7)
    // ---- rank-2nb update of the trailing matrix, lower triangle only: a_rc -= sum_k (V_kr conj(W_kc) + W_kr conj(V_kc)),
    // c <= r.  A thread takes the row pair (j1 + q, n - 1 - q): together they always have m + 1 columns, so every
    // thread of the group streams the same number of elements (rows alone would leave the last warp with twice the mean).
    const int j1 = j0 + nbp;
    const int m = n - j1;
#if defined(__CUDA_ARCH__)
    if (m > 0) {
      blk_rank2k_dmma<true>(A, lda, n, j1, V, W, ldp, nb);
      g.sync();
    }
#else
    if (m > 0) {
      const int half = (m + 1) / 2;
      int rw = ((half + 31) / 32) * 32;
      if (rw > T) rw = T;
      const int parts = T / rw > 0 ? T / rw : 1;
      const int pr = tid % rw, pp = tid / rw;
      if (pp < parts) {
        for (int q = pr; q < half; q += rw) {
          for (int side = 0; side < 2; ++side) {
            const int r = side == 0 ? j1 + q : n - 1 - q;
            if (side == 1 && r <= j1 + q) break;  // the middle row of an odd m is its own partner
            for (int k0 = 0; k0 < nb; k0 += 8) {  // 8 panel columns at a time in registers
              if (k0 >= nbp) break;
              cplx vr[8], wr[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
              for (int k = 0; k < 8; ++k) {
                const bool in = k0 + k < nb;
                vr[k] = in ? V[(k0 + k) * ldp + r] : mk(0.0, 0.0);
                wr[k] = in ? W[(k0 + k) * ldp + r] : mk(0.0, 0.0);
              }
              for (int c = j1 + pp; c <= r; c += parts) {
                cplx a = A[r + (size_t)c * lda];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k = 0; k < 8; ++k) {
                  if (k0 + k < nb) {
                    const cplx wc = W[(k0 + k) * ldp + c], vc = V[(k0 + k) * ldp + c];
                    // a -= vr * conj(wc) + wr * conj(vc)
                    a.re -= vr[k].re * wc.re + vr[k].im * wc.im + wr[k].re * vc.re + wr[k].im * vc.im;
                    a.im -= vr[k].im * wc.re - vr[k].re * wc.im + wr[k].im * vc.re - wr[k].re * vc.im;
                  }
                }
                A[r + (size_t)c * lda] = a;
              }
            }
          }
        }
      }
      g.sync();
    }
#endif
    TBK_HP(3)
  }
  if (tid == 0) {
    w.d[n - 1] = A[(n - 1) + (size_t)(n - 1) * lda].re;
    w.e[n - 1] = 0.0;
    w.tau[n - 1] = mk(0.0, 0.0);
  }
  TBK_HP(7)
  TBK_HP_END
  g.sync();
}

// ---------------------------------------------------------------------------------------------
// 1'. blocked tridiagonalisation, FULL-MATRIX variant (the round-1 kernel, kept selectable: TBK_HETRD=full).  On exit d, e hold T, the Householder vectors are stored below the
// first sub-diagonal of A (zhetd2 'L' layout, implicit unit at row j+1) with their scalars in tau.
// The strict upper triangle is overwritten (it is filled from the lower one first).
// ---------------------------------------------------------------------------------------------
template <class G>
TBK_HD void hetrd_blocked_full(G& g, const BlkWork& w) {
  const int n = w.n, lda = w.lda, nb = w.nb, ldp = w.ldp;
  const int T = g.size(), tid = g.tid();
  cplx* A = w.A;
  cplx* V = w.V;
  cplx* W = w.W;
  TBK_HP_BEGIN
  // full Hermitian storage: upper from lower, real diagonal
  for (int c = tid; c < n; c += T) A[c + (size_t)c * lda].im = 0.0;
  for (int r = tid; r < n; r += T)
    for (int c = r + 1; c < n; ++c) A[r + (size_t)c * lda] = conj(A[c + (size_t)r * lda]);
  g.sync();
  for (int j0 = 0; j0 < n - 1; j0 += nb) {
    const int nbp = n - 1 - j0 < nb ? n - 1 - j0 : nb;
    // zero the panels (unused panel columns must be exactly zero for the rank-2nb update)
    for (int q = tid; q < nb * ldp; q += T) { V[q] = mk(0.0, 0.0); W[q] = mk(0.0, 0.0); }
    g.sync();
    for (int i = 0; i < nbp; ++i) {
      const int j = j0 + i;
      cplx* col = A + (size_t)j * lda;
      // ---- (1) bring column j up to date with the reflectors of this panel
      if (i > 0) {
        for (int r = j + tid; r < n; r += T) {
          cplx a = col[r];
          for (int k = 0; k < i; ++k) {
            a = a - mulc(V[k * ldp + r], W[k * ldp + j]);
            a = a - mulc(W[k * ldp + r], V[k * ldp + j]);
          }
          if (r == j) a.im = 0.0;
          col[r] = a;
        }
        g.sync();
      }
      // ---- (2) reflector for x = A(j+1:n, j)   (zlarfg)
      double part = 0.0;
      for (int r = j + 2 + tid; r < n; r += T) part += norm2(col[r]);
      const double xnorm2 = g.sum(part);
      const cplx alpha = col[j + 1];
      cplx tau = mk(0.0, 0.0);
      double beta = alpha.re;
      cplx scal = mk(0.0, 0.0);
      if (xnorm2 != 0.0 || alpha.im != 0.0) {
        beta = -copysign(sqrt(alpha.re * alpha.re + alpha.im * alpha.im + xnorm2), alpha.re);
        tau = mk((beta - alpha.re) / beta, -alpha.im / beta);
        scal = cdiv(mk(1.0, 0.0), mk(alpha.re - beta, alpha.im));
      }
      g.sync();                                   // everyone has read alpha
      cplx* v = V + i * ldp;
      for (int r = j + 2 + tid; r < n; r += T) {
        const cplx x = col[r] * scal;             // tau == 0: the column is already zero below j+1
        col[r] = x;
        v[r] = x;
      }
      if (tid == 0) {
        v[j + 1] = mk(1.0, 0.0);
        w.e[j] = beta;
        w.tau[j] = tau;
        w.d[j] = col[j].re;
      }
      g.sync();
      TBK_HP(7)
      // ---- (3) w = A22 v with the stored (panel-start) trailing matrix, rows/cols j+1 .. n-1
      const int m = n - j - 1;
      {
        int rw = ((m + 31) / 32) * 32;
        if (rw > T) rw = T;
        const int parts = T / rw > 0 ? T / rw : 1;
        const int pr = tid % rw, pp = tid / rw;
        if (pp < parts) {
          for (int rb = 0; rb < m; rb += rw) {    // rb > 0 only when m > T
            const int r = j + 1 + rb + pr;
            cplx acc = mk(0.0, 0.0);
            if (r < n) {
              // eight independent accumulators: eight 16-byte loads in flight per thread, no serial FMA chain
              // (measured r13: with one matrix per SM the product moves 16 n^3 / 3 bytes per matrix at 5.7-6 TB/s in
              // aggregate — the HBM rate; a two-deep register pipeline only spilled at the 128-register cap and a TMA ring
              // (cp.async.bulk per column + mbarriers, all free shared memory in flight) was 4x slower: one 16-byte
              // element per thread per column cannot amortise a barrier wait and a slot release)
              const cplx* arow = A + r;
              cplx a0 = mk(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
              int c = j + 1 + pp;
              for (; c + 7 * parts < n; c += 8 * parts) {
                const cplx m0 = arow[(size_t)c * lda], m1 = arow[(size_t)(c + parts) * lda];
                const cplx m2 = arow[(size_t)(c + 2 * parts) * lda], m3 = arow[(size_t)(c + 3 * parts) * lda];
                const cplx m4 = arow[(size_t)(c + 4 * parts) * lda], m5 = arow[(size_t)(c + 5 * parts) * lda];
                const cplx m6 = arow[(size_t)(c + 6 * parts) * lda], m7 = arow[(size_t)(c + 7 * parts) * lda];
                fma_acc(a0, m0, v[c]);
                fma_acc(a1, m1, v[c + parts]);
                fma_acc(a2, m2, v[c + 2 * parts]);
                fma_acc(a3, m3, v[c + 3 * parts]);
                fma_acc(a4, m4, v[c + 4 * parts]);
                fma_acc(a5, m5, v[c + 5 * parts]);
                fma_acc(a6, m6, v[c + 6 * parts]);
                fma_acc(a7, m7, v[c + 7 * parts]);
              }
              for (; c + 3 * parts < n; c += 4 * parts) {
                const cplx m0 = arow[(size_t)c * lda], m1 = arow[(size_t)(c + parts) * lda];
                const cplx m2 = arow[(size_t)(c + 2 * parts) * lda], m3 = arow[(size_t)(c + 3 * parts) * lda];
                fma_acc(a0, m0, v[c]);
                fma_acc(a1, m1, v[c + parts]);
                fma_acc(a2, m2, v[c + 2 * parts]);
                fma_acc(a3, m3, v[c + 3 * parts]);
              }
              for (; c < n; c += parts) fma_acc(a0, arow[(size_t)c * lda], v[c]);
              acc = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
            }
            if (parts == 1) {
              if (r < n) W[i * ldp + r] = acc;
            } else {                              // parts > 1 implies m <= rw: a single row block
              w.wcol[pp * rw + pr] = acc;
            }
          }
        }
        if (parts > 1) {
          g.sync();
          for (int r = tid; r < m; r += T) {
            cplx acc = w.wcol[r];
            for (int q = 1; q < parts; ++q) acc = acc + w.wcol[q * rw + r];
            W[i * ldp + j + 1 + r] = acc;
          }
        }
      }
      g.sync();
      TBK_HP(0)
      // ---- panel corrections: dots[k] = W_k^H v, dots[nb+k] = V_k^H v, one sub-team per dot product
      if (i > 0) {
        for (int q = g.sub(); q < 2 * i; q += g.nsub()) {
          const cplx* src = q < i ? W + q * ldp : V + (q - i) * ldp;
          double sre = 0.0, sim = 0.0;
          for (int r = j + 1 + g.lane(); r < n; r += g.subsize()) {
            const cplx t = cmul(src[r], v[r]);
            sre += t.re; sim += t.im;
          }
          sre = g.subsum(sre); sim = g.subsum(sim);
          if (g.lane() == 0) w.dots[q < i ? q : nb + (q - i)] = mk(sre, sim);
        }
        g.sync();
        for (int r = j + 1 + tid; r < n; r += T) {
          cplx acc = W[i * ldp + r];
          for (int k = 0; k < i; ++k) {
            acc = acc - V[k * ldp + r] * w.dots[k];
            acc = acc - W[k * ldp + r] * w.dots[nb + k];
          }
          W[i * ldp + r] = acc;
        }
        g.sync();
      }
      // ---- w = tau w;  w += (-tau/2 (w^H v)) v
      double dre = 0.0, dim = 0.0;
      for (int r = j + 1 + tid; r < n; r += T) {
        const cplx wr = tau * W[i * ldp + r];
        W[i * ldp + r] = wr;
        const cplx t = cmul(wr, v[r]);
        dre += t.re; dim += t.im;
      }
      dre = g.sum(dre); dim = g.sum(dim);        // g.sum synchronises: the scaled w is visible
      const cplx a2 = (-0.5) * (tau * mk(dre, dim));
      for (int r = j + 1 + tid; r < n; r += T) W[i * ldp + r] = W[i * ldp + r] + a2 * v[r];
      g.sync();
    }
    TBK_HP(7)
    // ---- rank-2nb update of the trailing matrix: A22 -= V W^H + W V^H   (rows/cols >= j1)
    const int j1 = j0 + nbp;
    const int m = n - j1;
    if (m > 0) {
#if defined(__CUDA_ARCH__)
      blk_rank2k_dmma<false>(A, lda, n, j1, V, W, ldp, nb);
#else
      int rw = ((m + 31) / 32) * 32;
      if (rw > T) rw = T;
      const int parts = T / rw > 0 ? T / rw : 1;
      const int pr = tid % rw, pp = tid / rw;
      if (pp < parts) {
        for (int r = j1 + pr; r < n; r += rw) {
          for (int c = j1 + pp; c < n; c += parts) {
            cplx a = A[r + (size_t)c * lda];
            for (int k = 0; k < nb; ++k) {
              const cplx vr = V[k * ldp + r], wr = W[k * ldp + r], wc = W[k * ldp + c], vc = V[k * ldp + c];
              // a -= vr * conj(wc) + wr * conj(vc)
              a.re -= vr.re * wc.re + vr.im * wc.im + wr.re * vc.re + wr.im * vc.im;
              a.im -= vr.im * wc.re - vr.re * wc.im + wr.im * vc.re - wr.re * vc.im;
            }
            A[r + (size_t)c * lda] = a;
          }
        }
      }
#endif
      g.sync();
    }
    TBK_HP(3)
  }
  if (tid == 0) {
    w.d[n - 1] = A[(n - 1) + (size_t)(n - 1) * lda].re;
    w.e[n - 1] = 0.0;
    w.tau[n - 1] = mk(0.0, 0.0);
  }
  TBK_HP(7)
  TBK_HP_END
  g.sync();
}
// Both variants are kept because neither dominates (profiles/README.md r09): the full-matrix product (a thread per
// row, eight independent loads in flight per thread, no shuffle) is latency-friendlier, the lower-triangle one moves
// half the bytes.  w.wcol .. w.racc (>= group size complex numbers) serves as the split-product scratch here.

// ---------------------------------------------------------------------------------------------
// 2. eigenvalues of the tridiagonal (d, e) by bisection on the Sturm count; lam ascending.
// Returns (through every thread) the norm estimate tnorm = max Gershgorin radius.
// ---------------------------------------------------------------------------------------------
TBK_HD int sturm_count(int n, const double* d, const double* e2, double x, double pivmin) {
  int cnt = 0;
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += q < 0.0;
  for (int i = 1; i < n; ++i) {
    q = d[i] - x - e2[i - 1] * rcp_fast(q);        // |q| >= pivmin (a normal number): no IEEE-division slow path in the n-step chain
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;
}

template <class G>
TBK_HD double tridiag_bisect(G& g, const BlkWork& w) {
  const int n = w.n;
  const double eps = 2.220446049250313e-16, safmin = 2.2250738585072014e-308;
  for (int i = g.tid(); i < n; i += g.size()) w.e2[i] = i < n - 1 ? w.e[i] * w.e[i] : 0.0;
  g.sync();
  // Gershgorin interval and pivmin, computed redundantly by every thread (O(n), shared reads)
  double gl = w.d[0] - fabs(w.e[0]), gu = w.d[0] + fabs(w.e[0]), emax = 0.0;
  if (n == 1) { gl = gu = w.d[0]; }
  for (int i = 1; i < n; ++i) {
    const double rad = fabs(w.e[i - 1]) + (i < n - 1 ? fabs(w.e[i]) : 0.0);
    gl = fmin(gl, w.d[i] - rad);
    gu = fmax(gu, w.d[i] + rad);
    emax = fmax(emax, w.e2[i - 1]);
  }
  const double tnorm = fmax(fabs(gl), fabs(gu));
  const double pivmin = safmin * fmax(1.0, emax);
  gl -= 2.1 * tnorm * eps * n + 2.1 * pivmin;
  gu += 2.1 * tnorm * eps * n + 2.1 * pivmin;
  const double atol = 0.25 * eps * tnorm + 2.0 * pivmin;
  for (int j = g.tid(); j < n; j += g.size()) {
    double lo = gl, hi = gu;
    for (int it = 0; it < 120; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (hi - lo <= atol + eps * (fabs(lo) + fabs(hi)) || mid <= lo || mid >= hi) break;
      if (sturm_count(n, w.d, w.e2, mid, pivmin) > j) hi = mid; else lo = mid;
    }
    w.lam[j] = 0.5 * (lo + hi);
  }
  g.sync();
  return tnorm;
}

// ---------------------------------------------------------------------------------------------
// 3. eigenvectors of the tridiagonal by inverse iteration.  Returns 0, or 1 if the spectrum needs
// the fallback solver (huge cluster / dependent cluster vectors).
// ---------------------------------------------------------------------------------------------
// start vectors: a hash of (vector, component, attempt), uniform in (-1, 1).  (A linear congruential stream
// per vector is NOT good enough: streams seeded j * const differ by an affine function of j, so the
// start vectors of a degenerate cluster span only a few dimensions and Gram-Schmidt finds them dependent.)
TBK_HD double invit_rand(unsigned j, unsigned i, unsigned salt) {
  unsigned x = (j + 1u) * 0x9E3779B1u ^ (i + 1u) * 0x85EBCA77u ^ (salt + 1u) * 0xC2B2AE3Du;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return ((double)(x >> 8) + 0.5) * (2.0 / 16777216.0) - 1.0;
}

// One inverse-iteration solve (T - lam I) x = y for the thread that owns column t of the interleaved
// scratch arrays (element i of a per-thread array lives at [i * nt + t]); y is scaled to max-norm 1
// first, x overwrites y.  Partially pivoted elimination fused with the forward substitution
// (LAPACK dlagtf + dlagts), tiny pivots replaced by +-pivtol.
TBK_HD void tridiag_shifted_solve(int n, const double* d, const double* e, double lam, double pivtol,
                                  double* U0, double* U1, double* U2, double* Y, int nt, int t) {
  double mx = 0.0;
  for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(Y[(size_t)i * nt + t]));
  const double sc = mx > 0.0 ? 1.0 / mx : 1.0;
  double ak = d[0] - lam, bk = n > 1 ? e[0] : 0.0, yk = Y[t] * sc;
  // (both sweeps are serial chains in registers fed by loads whose addresses do not depend on the chain: unrolled, the
  // loads of the next steps are issued while the current pivot's division is still in flight)
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
  for (int k = 0; k < n - 1; ++k) {
    const double c = e[k], a1 = d[k + 1] - lam, b1 = k + 1 < n - 1 ? e[k + 1] : 0.0;
    const double y1 = Y[(size_t)(k + 1) * nt + t] * sc;
    const size_t at = (size_t)k * nt + t;
    if (fabs(c) <= fabs(ak)) {
      const double mult = ak != 0.0 ? c / ak : 0.0;
      U0[at] = ak; U1[at] = bk; U2[at] = 0.0; Y[at] = yk;
      ak = a1 - mult * bk; bk = b1; yk = y1 - mult * yk;
    } else {
      const double mult = ak / c;
      U0[at] = c; U1[at] = a1; U2[at] = b1; Y[at] = y1;
      ak = bk - mult * a1; bk = -mult * b1; yk = yk - mult * y1;
    }
  }
  {
    const size_t at = (size_t)(n - 1) * nt + t;
    U0[at] = ak; U1[at] = 0.0; U2[at] = 0.0; Y[at] = yk;
  }
  double x1 = 0.0, x2 = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
  for (int k = n - 1; k >= 0; --k) {
    const size_t at = (size_t)k * nt + t;
    double piv = U0[at];
    if (fabs(piv) < pivtol) piv = piv < 0.0 ? -pivtol : pivtol;
    const double x = (Y[at] - U1[at] * x1 - U2[at] * x2) / piv;
    Y[at] = x;
    x2 = x1; x1 = x;
  }
}

// Returns 0, or 1 if a cluster vector could not be made independent (the caller then uses the
// unblocked QL solver for this matrix).
template <class G>
TBK_HD int tridiag_invit(G& g, const BlkWork& w, double tnorm) {
  const int n = w.n, nt = w.nt;
  const double eps = 2.220446049250313e-16;
  const double scale = tnorm > 0.0 ? tnorm : 1.0;
  const double pertol = 10.0 * eps * scale;       // minimal separation of the shifts (dstein)
  const double ortol = 1.0e-4 * scale;            // closer eigenvalues are orthogonalised explicitly
  const double pivtol = eps * scale;
  // ---- serial pre-pass: separated shifts and clusters
  if (g.tid() == 0) {
    w.lamp[0] = w.lam[0];
    w.cl[0] = 0;
    for (int j = 1; j < n; ++j) {
      double x = w.lam[j];
      if (x - w.lamp[j - 1] < pertol) x = w.lamp[j - 1] + pertol;
      w.lamp[j] = x;
      w.cl[j] = (w.lam[j] - w.lam[j - 1] < ortol) ? w.cl[j - 1] : j;
    }
    w.ctl[1] = 0;
  }
  g.sync();
  double* U0 = w.lu;
  double* U1 = U0 + (size_t)n * nt;
  double* U2 = U1 + (size_t)n * nt;
  double* Y = U2 + (size_t)n * nt;
  // ---- phase 1: every eigenvector independently, one per thread, three solves from a random start
  for (int j0 = 0; j0 < n; j0 += nt) {
    const int t = g.tid();
    const int j = j0 + t;
    if (t < nt && j < n) {
      for (int i = 0; i < n; ++i) Y[(size_t)i * nt + t] = invit_rand((unsigned)j, (unsigned)i, 0u);
      for (int it = 0; it < 3; ++it) tridiag_shifted_solve(n, w.d, w.e, w.lamp[j], pivtol, U0, U1, U2, Y, nt, t);
      // normalise, largest component positive, store as column j of Z
      double nrm = 0.0, big = 0.0;
      for (int i = 0; i < n; ++i) {
        const double x = Y[(size_t)i * nt + t];
        nrm += x * x;
        if (fabs(x) > fabs(big)) big = x;
      }
      const double sc = (big < 0.0 ? -1.0 : 1.0) / sqrt(nrm);
      for (int i = 0; i < n; ++i) w.Z[(size_t)i * n + j] = Y[(size_t)i * nt + t] * sc;
    }
    g.sync();
  }
  // ---- phase 2: modified Gram-Schmidt inside clusters, one cluster per sub-team.  A member that loses
  // its norm against the earlier members (it converged to the same direction of a degenerate
  // eigenspace) is regenerated dstein-style: new random start, orthogonalised BEFORE and after
  // each solve; the solve itself is done by the sub-team's lane 0 in its private scratch column.
  for (int s = g.sub(); s < n; s += g.nsub()) {
    if (w.cl[s] != s) continue;                    // s is not the first member of a cluster
    int last = s;
    while (last + 1 < n && w.cl[last + 1] == s) ++last;
    if (last == s) continue;
    if (last - s + 1 > kWarpCluster) continue;     // large clusters: whole group, below
    const int L = g.lane(), S = g.subsize();
    const int t0 = g.tid() - L;                    // scratch column of this sub-team's lane 0
    const bool have_scratch = t0 < nt;
    for (int b = s + 1; b <= last; ++b) {
      int attempt = 0, solves = 0;
      for (int round = 0; round < 16; ++round) {
        // project out the earlier members (twice), measuring the norm that survives the first pass
        double keep = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
          double before = 0.0;
          for (int i = L; i < n; i += S) { const double x = w.Z[(size_t)i * n + b]; before += x * x; }
          before = sqrt(g.subsum(before));
          for (int a = s; a < b; ++a) {
            double dot = 0.0;
            for (int i = L; i < n; i += S) dot += w.Z[(size_t)i * n + a] * w.Z[(size_t)i * n + b];
            dot = g.subsum(dot);
            for (int i = L; i < n; i += S) w.Z[(size_t)i * n + b] -= dot * w.Z[(size_t)i * n + a];
            g.subsync();
          }
          double nrm = 0.0;
          for (int i = L; i < n; i += S) { const double x = w.Z[(size_t)i * n + b]; nrm += x * x; }
          nrm = sqrt(g.subsum(nrm));
          if (pass == 0) keep = before > 0.0 ? nrm / before : 0.0;
          const double sc = nrm > 0.0 ? 1.0 / nrm : 0.0;
          for (int i = L; i < n; i += S) w.Z[(size_t)i * n + b] *= sc;
          g.subsync();
        }
        // done when the vector kept a fair share of its norm and (if regenerated) went through three solves
        if (keep >= 1.0e-2 && (attempt == 0 || solves >= 3)) break;
        if (!have_scratch || round == 15) {        // cannot iterate here / no independent direction found
          if (L == 0) w.ctl[1] = 1;
          break;
        }
        if (keep < 1.0e-2) {                        // dependent: restart from a fresh random vector
          ++attempt;
          solves = 0;
          for (int i = L; i < n; i += S) w.Z[(size_t)i * n + b] = invit_rand((unsigned)b, (unsigned)i, (unsigned)attempt);
          g.subsync();
          continue;                                 // orthogonalise the start vector first
        }
        // one more solve on the orthogonalised vector
        ++solves;
        if (L == 0) {
          for (int i = 0; i < n; ++i) Y[(size_t)i * nt + t0] = w.Z[(size_t)i * n + b];
          tridiag_shifted_solve(n, w.d, w.e, w.lamp[b], pivtol, U0, U1, U2, Y, nt, t0);
          for (int i = 0; i < n; ++i) w.Z[(size_t)i * n + b] = Y[(size_t)i * nt + t0];
        }
        g.subsync();
      }
    }
  }
  g.sync();
  // ---- phase 3: large clusters (flat bands, high-symmetry k-points), one after another by the WHOLE
  // group: classical Gram-Schmidt applied twice.  The members of a cluster are adjacent columns of the
  // row-major Z, so the projections <z_a, z_b> for all a < b are computed one per thread with coalesced
  // reads, and the update runs one row per thread.  A single sub-team would need O(c^2) serial
  // dot products here (measured: one 200-member cluster took 8x the time of the rest of the matrix).
  for (int s = 0; s < n; ++s) {
    if (w.cl[s] != s) continue;
    int last = s;
    while (last + 1 < n && w.cl[last + 1] == s) ++last;
    if (last - s + 1 <= kWarpCluster) { s = last; continue; }
    double* proj = w.e2;                            // free after the bisection
    for (int b = s + 1; b <= last; ++b) {
      for (int pass = 0; pass < 2; ++pass) {
        double before = 0.0;
        for (int a = s + g.tid(); a < b; a += g.size()) {
          double dot = 0.0;
          for (int i = 0; i < n; ++i) dot += w.Z[(size_t)i * n + a] * w.Z[(size_t)i * n + b];
          proj[a - s] = dot;
        }
        for (int i = g.tid(); i < n; i += g.size()) { const double x = w.Z[(size_t)i * n + b]; before += x * x; }
        before = sqrt(g.sum(before));               // g.sum synchronises: proj is visible
        double part = 0.0;
        for (int i = g.tid(); i < n; i += g.size()) {
          const double* row = w.Z + (size_t)i * n;
          double acc = row[b];
          for (int a = s; a < b; ++a) acc -= proj[a - s] * row[a];
          w.Z[(size_t)i * n + b] = acc;
          part += acc * acc;
        }
        const double nrm = sqrt(g.sum(part));
        // In a c-fold degenerate eigenspace the last members legitimately keep only ~1/sqrt(c) of their norm
        // (and with some probability 10-100x less); what survives still lies in the eigenspace, and its error
        // outside the cluster is amplified by at most 1/keep.  Below 1e-4 the vector is left to the fallback.
        if (pass == 0 && !(nrm >= 1.0e-4 * before) && g.tid() == 0) w.ctl[1] = 1;
        const double sc = nrm > 0.0 ? 1.0 / nrm : 0.0;
        for (int i = g.tid(); i < n; i += g.size()) w.Z[(size_t)i * n + b] *= sc;
        g.sync();
      }
    }
    s = last;
  }
  g.sync();
  return w.ctl[1];
}

// ---------------------------------------------------------------------------------------------
// 4. back-transformation of ONE tridiagonal eigenvector (column c of Z) by a sub-team:
// x = H_0 H_1 ... H_{n-2} z, reflectors in A (zhetd2 'L' layout).  Element r of x is owned by lane
// r % subsize, slot r / subsize; MAXM >= ceil(n / subsize).  The result is handed to `store(r, x_r)`.
// ---------------------------------------------------------------------------------------------
template <int MAXM, class G, class Store>
TBK_HD void backtransform_column(G& g, const BlkWork& w, int c, Store store) {
  const int n = w.n, lda = w.lda;
  const int L = g.lane(), S = g.subsize();
  cplx x[MAXM];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m < MAXM; ++m) {
    const int r = L + S * m;
    x[m] = mk(r < n ? w.Z[(size_t)r * n + c] : 0.0, 0.0);
  }
  for (int j = n - 2; j >= 0; --j) {
    const cplx tau = w.tau[j];
    if (tau.re == 0.0 && tau.im == 0.0) continue;
    const cplx* vcol = w.A + (size_t)j * lda;
    double dre = 0.0, dim = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m = 0; m < MAXM; ++m) {
      const int r = L + S * m;
      if (r > j && r < n) {
        const cplx v = r == j + 1 ? mk(1.0, 0.0) : vcol[r];
        const cplx t = cmul(v, x[m]);             // conj(v) x
        dre += t.re; dim += t.im;
      }
    }
    dre = g.subsum(dre); dim = g.subsum(dim);
    const cplx f = tau * mk(dre, dim);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m = 0; m < MAXM; ++m) {
      const int r = L + S * m;
      if (r > j && r < n) {                       // second read of v: an L1 hit, cheaper than MAXM live registers
        const cplx v = r == j + 1 ? mk(1.0, 0.0) : vcol[r];
        x[m] = x[m] - f * v;
      }
    }
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m < MAXM; ++m) {
    const int r = L + S * m;
    if (r < n) store(r, x[m]);
  }
}

// ---------------------------------------------------------------------------------------------
// 4b. back-transformation of ALL tridiagonal eigenvectors by the whole group: every sub-team owns
// CB columns of a batch of nsub * CB columns (held in registers as above); the reflectors are staged
// RB at a time in `stage` (shared memory, [RB][n], zero above the unit element) by all threads,
// so that the matrix of reflectors is read from L2/HBM once per column batch instead of once per
// column.  A reflector element read from shared memory serves all CB columns of the sub-team: with
// one column the loop is bound by those reads (two 16-byte reads per 8 FMAs), with two or four it is
// bound by the FP64 pipe.  CB is limited by the registers (CB * MAXM complex numbers per thread).
// store(c, r, x_r) receives element r of eigenvector c.
// ---------------------------------------------------------------------------------------------
template <int MAXM, int CB, class G, class Store>
TBK_HD void backtransform_all(G& g, const BlkWork& w, cplx* stage, int RB, Store store) {
  const int n = w.n, lda = w.lda;
  const int L = g.lane(), S = g.subsize(), T = g.size(), tid = g.tid();
  for (int c0 = 0; c0 < n; c0 += g.nsub() * CB) {
    const int cbase = c0 + g.sub() * CB;
    const bool active = cbase < n;
    cplx x[CB][MAXM];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < CB; ++q) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int m = 0; m < MAXM; ++m) {
        const int r = L + S * m;
        x[q][m] = mk((cbase + q < n && r < n) ? w.Z[(size_t)r * n + cbase + q] : 0.0, 0.0);
      }
    }
    for (int jhi = n - 2; jhi >= 0; jhi -= RB) {
      const int jlo = jhi - RB + 1 > 0 ? jhi - RB + 1 : 0;
      const int cnt = jhi - jlo + 1;
      g.sync();                                   // the previous stage has been consumed
      for (int q = tid; q < cnt * n; q += T) {
        const int jj = q / n, r = q - jj * n, j = jlo + jj;
        stage[q] = r > j + 1 ? w.A[r + (size_t)j * lda] : mk(r == j + 1 ? 1.0 : 0.0, 0.0);
      }
      g.sync();
      if (!active) continue;
      for (int j = jhi; j >= jlo; --j) {
        const cplx tau = w.tau[j];
        if (tau.re == 0.0 && tau.im == 0.0) continue;
        const cplx* v = stage + (size_t)(j - jlo) * n;
        double dre[CB], dim[CB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < CB; ++q) { dre[q] = 0.0; dim[q] = 0.0; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int m = 0; m < MAXM; ++m) {
          if (S * m + S - 1 > j) {                // sub-team-uniform: this slot holds rows > j
            const int r = L + S * m;
            if (r < n) {
              const cplx vr = v[r];               // zero for r <= j
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
              for (int q = 0; q < CB; ++q) {
                const cplx t = cmul(vr, x[q][m]); // conj(v) x
                dre[q] += t.re; dim[q] += t.im;
              }
            }
          }
        }
        cplx f[CB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < CB; ++q) {
          const double a = g.subsum(dre[q]), b = g.subsum(dim[q]);
          f[q] = tau * mk(a, b);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int m = 0; m < MAXM; ++m) {
          if (S * m + S - 1 > j) {
            const int r = L + S * m;
            if (r < n) {
              const cplx vr = v[r];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
              for (int q = 0; q < CB; ++q) x[q][m] = x[q][m] - f[q] * vr;
            }
          }
        }
      }
    }
    if (active) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int q = 0; q < CB; ++q) {
        if (cbase + q < n) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
          for (int m = 0; m < MAXM; ++m) {
            const int r = L + S * m;
            if (r < n) store(cbase + q, r, x[q][m]);
          }
        }
      }
    }
  }
  g.sync();
}

}  // namespace tbk
