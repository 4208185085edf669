// tbk_eig_wy.cuh — back-transformation of the blocked eigensolver (n = 33..512) as compact-WY blocks
// on the FP64 tensor pipe.  Included by tbk_solve.cu after tbk_eig_blocked.cuh.
//
// After the tridiagonalisation the eigenvectors of H(k) are  X = H_0 H_1 ... H_{n-2} Z  with Z the (real)
// eigenvectors of the tridiagonal matrix (numpy.linalg.eigh at pythtb.py:939 returns exactly these columns).
// Applying the reflectors one at a time is a BLAS-2 loop (two shared-memory reads per 8 FMAs, ~15 % of the FP64
// rate); here kWyNB = 32 consecutive reflectors are merged into  I - V T V^H  (LAPACK zlarft, forward / columnwise)
// and applied to a panel of kWyNC = 16 eigenvector columns as three GEMMs on mma.sync.m8n8k4.f64 (DMMA):
//     Y  = V^H X      (32 x m)(m x 16)      K = m  streamed in 64-row chunks, split over two warp groups
//     Y' = T Y        (32 x 32)(32 x 16)
//     X -= V Y'       (m x 32)(32 x 16)     64 rows per chunk, 8 rows per warp
// The X panel (n x 16 complex) stays in shared memory for the whole sweep over the reflector blocks, the chunks of
// V are streamed from L2 into a two-stage cp.async ring (one __syncthreads per chunk), so a CTA reads every reflector
// element twice per block and does 2 x 8 x 16 = 256 flops on each 16 bytes it loads.
//
//   blk_wy_kernel              one CTA per (matrix, block): unit-lower-trapezoidal V made explicit in A (ones on the
//                              first sub-diagonal of the block, zeros above), Gram matrix V^H V on the tensor pipe,
//                              the triangular recurrence for T (one row per lane), T stored column-major
//   blk_backtransform_kernel   one CTA per (matrix, 16-column panel): the sweep above, then the Convention-I gauge
//                              factors, periodic images and the closing-row factor on the way out (pythtb.py:2729-2747)
//
// Fragment layouts of mma.sync.aligned.m8n8k4.row.col.f64 (g = lane / 4, q = lane % 4):
//     A (8 x 4): lane holds A[g][q]      B (4 x 8): lane holds B[q][g]      C (8 x 8): lane holds C[g][2q], C[g][2q + 1]
// Shared-memory leading dimensions are chosen so that the 16-byte (re, im) fragment reads of every quarter-warp hit
// eight distinct 16-byte bank groups: a chunk is stored [column][row] with LD = 68 (== 4 mod 8) when it is the A
// operand of V^H X (lane -> column g, row q) and with LD = 66 (== 2 mod 8) when it is the A operand of X -= V Y'
// (lane -> row g, column q); the X panel and the Y blocks are stored [column][row] with LD == 2 mod 8, and fragment
// column x of an 8-column tile is panel column wy_perm(x) so that both the B-operand reads (columns g of a quarter-warp:
// 2p, 2p + 1) and the C-fragment read-modify-writes (columns 2q + e, q = 0..3) are conflict-free.
#pragma once

namespace tbk {

constexpr int kWyNB = 32;        // reflectors per compact-WY block
constexpr int kWyCR = 64;        // rows of V per staged chunk
constexpr int kWyNC = 16;        // eigenvector columns per CTA
constexpr int kWyLD1 = 68;       // chunk leading dimension, V^H X (and the Gram matrix)
constexpr int kWyLD2 = 66;       // chunk leading dimension, X -= V Y' and the T block
constexpr int kWyLDY = 34;       // Y blocks [16][34]
constexpr int kWyThreads = 256;
constexpr int kWyBufElems = kWyNB * kWyLD1;

// per-matrix slot of the staged solver's workspace (tbk_solve.cu: blk_stage_layout)
struct WyArgs {
  int n, lda, nblk;
  size_t slot_bytes;
  size_t off_A, off_Z, off_tau, off_T, off_flag;
  char* ws;
};

TBK_HD int wy_ldx(int n) { return ((n + 5) / 8) * 8 + 2; }            // smallest value >= n that is 2 mod 8
TBK_HD int wy_nblk(int n) { return (n - 1 + kWyNB - 1) / kWyNB; }
TBK_HD size_t wy_bt_smem(int n) {
  return ((size_t)kWyNC * wy_ldx(n) + 2 * (size_t)kWyBufElems + 3 * (size_t)kWyNC * kWyLDY) * 16;
}
TBK_HD size_t wy_t_smem() { return (2 * (size_t)kWyBufElems + 2 * (size_t)kWyNB * (kWyNB + 1)) * 16; }

#if defined(__CUDACC__)

__device__ __forceinline__ int wy_perm(int x) { return (int)((0x57463120u >> (4 * x)) & 7u); }

__device__ __forceinline__ void wy_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void wy_cp16(void* dst, const void* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;                   // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void wy_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wy_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows r0 .. r0 + 63 of the reflector columns j0 .. j0 + 31 -> buf[column * LD + row]; rows >= n and columns
// >= n - 1 (there are n - 1 reflectors) are zero-filled
__device__ __forceinline__ void wy_issue_v(cplx* buf, int LD, const cplx* A, int lda, int n, int j0, int r0) {
  for (int e = threadIdx.x; e < kWyNB * kWyCR; e += kWyThreads) {
    const int col = e >> 6, row = e & 63;
    const int j = j0 + col, r = r0 + row;
    const bool ok = j < n - 1 && r < n;
    wy_cp16(buf + col * LD + row, ok ? A + r + (size_t)j * lda : A, ok);
  }
}

// ---------------------------------------------------------------------------------------------
// T factors.  grid (nblk, matrices of the chunk), 256 threads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWyThreads)
blk_wy_kernel(const WyArgs a) {
  extern __shared__ __align__(16) char smem[];
  cplx* buf[2] = {(cplx*)smem, (cplx*)smem + kWyBufElems};
  cplx* Gs = (cplx*)smem + 2 * kWyBufElems;                 // [32][33]
  cplx* Ts = Gs + kWyNB * (kWyNB + 1);                      // [32][33]
  char* mine = a.ws + (size_t)blockIdx.y * a.slot_bytes;
  if (*(const int*)(mine + a.off_flag) != 0) return;        // solved by the fallback: nothing to back-transform
  const int n = a.n, lda = a.lda, b = blockIdx.x;
  cplx* A = (cplx*)(mine + a.off_A);
  const cplx* tau = (const cplx*)(mine + a.off_tau);
  cplx* Tg = (cplx*)(mine + a.off_T) + (size_t)b * kWyNB * kWyNB;
  const int j0 = b * kWyNB;
  const int kb = n - 1 - j0 < kWyNB ? n - 1 - j0 : kWyNB;
  const int r_first = j0 + 1;
  const int nch = (n - r_first + kWyCR - 1) / kWyCR;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int rt = warp & 3, ctp = warp >> 2;
  double cre[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, cim[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  wy_issue_v(buf[0], kWyLD1, A, lda, n, j0, r_first);
  wy_commit();
  for (int t = 0; t < nch; ++t) {
    cplx* cur = buf[t & 1];
    wy_wait_all();
    __syncthreads();
    if (t + 1 < nch) { wy_issue_v(buf[(t + 1) & 1], kWyLD1, A, lda, n, j0, r_first + (t + 1) * kWyCR); wy_commit(); }
    if (t == 0) {
      // the implicit part of V: ones on the first sub-diagonal of the block, zeros above it — made explicit here and
      // in A (the back-transformation streams A as it is; d, e, tau were saved by the tridiagonalisation)
      for (int e = tid; e < kWyNB * kWyNB; e += kWyThreads) {
        const int col = e >> 5, row = e & 31;
        if (col < kb && row <= col && r_first + row < n) {
          const cplx v = mk(row == col ? 1.0 : 0.0, 0.0);
          cur[col * kWyLD1 + row] = v;
          A[(r_first + row) + (size_t)(j0 + col) * lda] = v;
        }
      }
      __syncthreads();
    }
    // G += V^H V on the chunk: warp (rt, ctp) owns rows 8 rt .. 8 rt + 7, columns 16 ctp .. 16 ctp + 15
#pragma unroll 4
    for (int ks = 0; ks < kWyCR / 4; ++ks) {
      const int rr = ks * 4 + q;
      const cplx v = cur[(rt * 8 + g) * kWyLD1 + rr];
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        const cplx x = cur[((ctp * 2 + ct) * 8 + g) * kWyLD1 + rr];
        wy_dmma(cre[ct][0], cre[ct][1], v.re, x.re);       // conj(v) x
        wy_dmma(cre[ct][0], cre[ct][1], v.im, x.im);
        wy_dmma(cim[ct][0], cim[ct][1], v.re, x.im);
        wy_dmma(cim[ct][0], cim[ct][1], -v.im, x.re);
      }
    }
  }
#pragma unroll
  for (int ct = 0; ct < 2; ++ct)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      Gs[(rt * 8 + g) * (kWyNB + 1) + (ctp * 2 + ct) * 8 + 2 * q + e] = mk(cre[ct][e], cim[ct][e]);
  for (int e = tid; e < kWyNB * (kWyNB + 1); e += kWyThreads) Ts[e] = mk(0.0, 0.0);
  __syncthreads();
  // zlarft, forward / columnwise:  T(i,i) = tau_i,  T(0:i, i) = -tau_i T(0:i, 0:i) (V(:, 0:i)^H v_i).  Row `row` of T
  // depends on nothing but itself and G: one row per lane, no synchronisation inside the recurrence.
  if (tid < kb) {
    const int row = tid;
    cplx* trow = Ts + row * (kWyNB + 1);
    trow[row] = tau[j0 + row];
    for (int i = row + 1; i < kb; ++i) {
      cplx s = mk(0.0, 0.0);
      for (int k = row; k < i; ++k) fma_acc(s, trow[k], Gs[k * (kWyNB + 1) + i]);
      trow[i] = -(tau[j0 + i] * s);
    }
  }
  __syncthreads();
  for (int e = tid; e < kWyNB * kWyNB; e += kWyThreads) {      // column-major: Tg[k * 32 + i] = T[i][k]
    const int k = e >> 5, i = e & 31;
    Tg[e] = Ts[i * (kWyNB + 1) + k];
  }
}

// element o of eigenvector c of mesh point / list entry idx -> the output array (the store of solve_blocked_kernel)
struct BlkStorePoint {
  long long idx, base;
  int zero_mask;
  bool closing;
};
__device__ __forceinline__ void blk_store_vec(const OutSpec& out, int n, const BlkStorePoint& p, int c, int o, cplx v) {
  if (out.mode == 0) {
    out.evec[c * out.vc_sb + p.idx * out.vc_sk + o] = v;
    return;
  }
  if (p.closing) v = v * out.pbc_phase[o];
  const long long at = (long long)c * out.sstride + o;
  out.evec[p.base + at] = v;
  if (p.zero_mask) {
    for (int mm = 1; mm < (1 << out.nd); ++mm) {
      if ((mm & p.zero_mask) != mm) continue;
      long long off = p.base;
      cplx f = v;
      for (int d = 0; d < out.nd; ++d)
        if (mm & (1 << d)) {
          off += (long long)(out.full[d] - 1) * out.gstride[d];
          f = f * out.pbc_phase[d * n + o];
        }
      out.evec[off + at] = f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Back-transformation.  grid (ceil(n / 16), matrices of the chunk), 256 threads, wy_bt_smem(n) bytes.
// ---------------------------------------------------------------------------------------------
struct WyPos {
  int b, phase, t, nch;          // phase 0: chunks of V^H X, 1: the T block, 2: chunks of X -= V Y'
};
__device__ __forceinline__ int wy_chunks(int n, int b) { return (n - (b * kWyNB + 1) + kWyCR - 1) / kWyCR; }
__device__ __forceinline__ void wy_advance(WyPos& p, int n) {
  if (p.phase == 1) { p.phase = 2; p.t = 0; return; }
  if (++p.t < p.nch) return;
  if (p.phase == 0) { p.phase = 1; return; }
  --p.b; p.phase = 0; p.t = 0;
  if (p.b >= 0) p.nch = wy_chunks(n, p.b);
}
__device__ __forceinline__ void wy_issue(const WyPos& p, cplx* buf, const cplx* A, const cplx* Tg, int lda, int n) {
  if (p.phase == 1) {
    const cplx* src = Tg + (size_t)p.b * kWyNB * kWyNB;
    for (int e = threadIdx.x; e < kWyNB * kWyNB; e += kWyThreads) wy_cp16(buf + (e >> 5) * kWyLD2 + (e & 31), src + e, true);
  } else {
    wy_issue_v(buf, p.phase == 0 ? kWyLD1 : kWyLD2, A, lda, n, p.b * kWyNB, p.b * kWyNB + 1 + p.t * kWyCR);
  }
  wy_commit();
}

__global__ void __launch_bounds__(kWyThreads, 1)
blk_backtransform_kernel(const WyArgs a, const PlanView pv, const KSrc ks, const OutSpec out, const int has_h,
                         const long long idx0) {
  extern __shared__ __align__(16) char smem[];
  __shared__ double s_k[TBK_MAX_DIM];
  __shared__ int s_mi[TBK_MAX_DIM];
  const int n = a.n, lda = a.lda, ldx = wy_ldx(n);
  cplx* Xs = (cplx*)smem;                                   // [16][ldx]
  cplx* buf[2] = {Xs + (size_t)kWyNC * ldx, Xs + (size_t)kWyNC * ldx + kWyBufElems};
  cplx* R0 = buf[1] + kWyBufElems;                          // partial of the rows' first half, then Y' = T Y
  cplx* R1 = R0 + kWyNC * kWyLDY;                           // partial of the second half
  cplx* R2 = R1 + kWyNC * kWyLDY;                           // Y = V^H X
  char* mine = a.ws + (size_t)blockIdx.y * a.slot_bytes;
  if (*(const int*)(mine + a.off_flag) != 0) return;
  const cplx* A = (const cplx*)(mine + a.off_A);
  const double* Z = (const double*)(mine + a.off_Z);
  const cplx* Tg = (const cplx*)(mine + a.off_T);
  const int c0 = blockIdx.x * kWyNC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int pg = wy_perm(g), p0 = wy_perm(2 * q), p1 = wy_perm(2 * q + 1);

  WyPos cur;
  cur.b = a.nblk - 1; cur.phase = 0; cur.t = 0; cur.nch = wy_chunks(n, cur.b);
  WyPos nxt = cur;
  wy_issue(nxt, buf[0], A, Tg, lda, n);
  wy_advance(nxt, n);
  // X panel <- columns c0 .. c0 + 15 of Z (real), while the first chunk is on its way
  for (int e = tid; e < kWyNC * n; e += kWyThreads) {
    const int c = e & (kWyNC - 1), r = e >> 4;
    Xs[c * ldx + r] = mk(c0 + c < n ? Z[(size_t)r * n + c0 + c] : 0.0, 0.0);
  }
  double yre[2][2], yim[2][2];                              // V^H X accumulators: warp (rt1, kh)
  const int rt1 = warp & 3, kh = warp >> 2;
  int s = 0;
  while (cur.b >= 0) {
    wy_wait_all();
    __syncthreads();
    if (nxt.b >= 0) { wy_issue(nxt, buf[s ^ 1], A, Tg, lda, n); wy_advance(nxt, n); }
    const cplx* cb = buf[s];
    const int r0 = cur.b * kWyNB + 1 + cur.t * kWyCR;
    if (cur.phase == 0) {
      if (cur.t == 0) {
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) { yre[ct][0] = yre[ct][1] = 0.0; yim[ct][0] = yim[ct][1] = 0.0; }
      }
      // Y[i][c] += sum_r conj(V[r][i]) X[r][c]: rows 32 kh .. 32 kh + 31 of the chunk, i = 8 rt1 + g
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (r0 + kh * 32 + ks * 4 >= n) break;              // warp-uniform: the rest of this half chunk lies past row n - 1
        const int rr = kh * 32 + ks * 4 + q;
        const cplx v = cb[(rt1 * 8 + g) * kWyLD1 + rr];
        const int r = r0 + rr;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          cplx x = mk(0.0, 0.0);
          if (r < n) x = Xs[(ct * 8 + pg) * ldx + r];
          wy_dmma(yre[ct][0], yre[ct][1], v.re, x.re);
          wy_dmma(yre[ct][0], yre[ct][1], v.im, x.im);
          wy_dmma(yim[ct][0], yim[ct][1], v.re, x.im);
          wy_dmma(yim[ct][0], yim[ct][1], -v.im, x.re);
        }
      }
      if (cur.t == cur.nch - 1) {
        cplx* R = kh ? R1 : R0;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          R[(ct * 8 + p0) * kWyLDY + rt1 * 8 + g] = mk(yre[ct][0], yim[ct][0]);
          R[(ct * 8 + p1) * kWyLDY + rt1 * 8 + g] = mk(yre[ct][1], yim[ct][1]);
        }
      }
    } else if (cur.phase == 1) {
      for (int e = tid; e < kWyNC * kWyNB; e += kWyThreads) {
        const int at = (e >> 5) * kWyLDY + (e & 31);
        R2[at] = R0[at] + R1[at];
      }
      __syncthreads();
      // Y' = T Y: warp (rt1, kh) owns rows 8 rt1 .. + 7 of columns 8 kh .. + 7; T is stored [k][i] with LD 66
      double tre[2] = {0.0, 0.0}, tim[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int k = ks * 4 + q;
        const cplx tt = cb[k * kWyLD2 + rt1 * 8 + g];
        const cplx y = R2[(kh * 8 + pg) * kWyLDY + k];
        wy_dmma(tre[0], tre[1], tt.re, y.re);
        wy_dmma(tre[0], tre[1], -tt.im, y.im);
        wy_dmma(tim[0], tim[1], tt.re, y.im);
        wy_dmma(tim[0], tim[1], tt.im, y.re);
      }
      R0[(kh * 8 + p0) * kWyLDY + rt1 * 8 + g] = mk(tre[0], tim[0]);
      R0[(kh * 8 + p1) * kWyLDY + rt1 * 8 + g] = mk(tre[1], tim[1]);
    } else {
      // X[r][c] -= sum_i V[r][i] Y'[i][c]: warp owns rows 8 warp .. + 7 of the chunk
      const int rr = warp * 8 + g, r = r0 + rr;
      if (r0 + warp * 8 < n) {                              // warp-uniform: this warp's eight rows are not all past row n - 1
      double xre[2][2], xim[2][2];
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        cplx x0 = mk(0.0, 0.0), x1 = x0;
        if (r < n) { x0 = Xs[(ct * 8 + p0) * ldx + r]; x1 = Xs[(ct * 8 + p1) * ldx + r]; }
        xre[ct][0] = x0.re; xim[ct][0] = x0.im; xre[ct][1] = x1.re; xim[ct][1] = x1.im;
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int k = ks * 4 + q;
        const cplx v = cb[k * kWyLD2 + rr];
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          const cplx y = R0[(ct * 8 + pg) * kWyLDY + k];
          wy_dmma(xre[ct][0], xre[ct][1], -v.re, y.re);
          wy_dmma(xre[ct][0], xre[ct][1], v.im, y.im);
          wy_dmma(xim[ct][0], xim[ct][1], -v.re, y.im);
          wy_dmma(xim[ct][0], xim[ct][1], -v.im, y.re);
        }
      }
      if (r < n) {
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          Xs[(ct * 8 + p0) * ldx + r] = mk(xre[ct][0], xim[ct][0]);
          Xs[(ct * 8 + p1) * ldx + r] = mk(xre[ct][1], xim[ct][1]);
        }
      }
      }
    }
    wy_advance(cur, n);
    s ^= 1;
  }
  // ---- out: gauge factors (Convention I: component o times conj(d_o(k))), periodic images, closing row
  const long long idx = idx0 + blockIdx.y;
  if (tid == 0) {
    int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
    double k[TBK_MAX_DIM];
    if (out.mode == 1) decode_index(idx, out, mi);
    load_k(ks, idx, mi, k);
    for (int d = 0; d < TBK_MAX_DIM; ++d) { s_k[d] = k[d]; s_mi[d] = mi[d]; }
  }
  __syncthreads();                                          // also: every warp is done with the ring and with X
  cplx* gf = buf[0];
  for (int o = tid; o < n; o += kWyThreads) {
    cplx f = mk(1.0, 0.0);
    if (!has_h && pv.convention == 1 && pv.dim_k > 0) f = conj(plan_gauge(pv, s_k, o));
    gf[o] = f;
  }
  __syncthreads();
  BlkStorePoint pt;
  pt.idx = idx; pt.base = 0; pt.zero_mask = 0; pt.closing = false;
  if (out.mode == 1) {
    pt.closing = is_closing(ks, s_mi);
    for (int d = 0; d < out.nd; ++d) {
      pt.base += s_mi[d] * out.gstride[d];
      if (s_mi[d] == 0 && out.wrap[d]) pt.zero_mask |= 1 << d;
    }
  }
  for (int c = 0; c < kWyNC && c0 + c < n; ++c)
    for (int o = tid; o < n; o += kWyThreads) blk_store_vec(out, n, pt, c0 + c, o, Xs[c * ldx + o] * gf[o]);
}

#endif  // __CUDACC__

}  // namespace tbk
