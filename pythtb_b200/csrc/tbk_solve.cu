// tbk_solve.cu — Bloch-Hamiltonian assembly fused with batched Hermitian
// diagonalisation (the reference's _gen_ham -> _sol_ham loop,
// pythtb.py:1047-1060 and 2475-2525), written for sm_100a.
//
// Kernel families (chosen by nsta):
//   solve_small_kernel<N>   N = 2,3,4: one k-point per thread, H and the
//                           eigenvectors never leave registers (tbk_eig_small.cuh).
//   solve_reg_gemm_kernel   5 <= nsta <= 8: one k-point per thread, H(k) assembled on the
//   solve_reg_kernel        FP64 tensor pipe from a dense coefficient table (or by a scalar
//                           element-major loop), eigensolver in registers.
//   solve_tile_kernel<G>    5 <= nsta <= 32: one k-point per G-lane tile of a
//                           warp (G = 8,16,32), matrix resident in shared memory.
//   solve_blocked_kernel    33 <= nsta <= 512: staged blocked solver (tbk_eig_blocked.cuh,
//                           tbk_eig_wy.cuh).
//   solve_block_kernel      nsta > 32: one k-point per CTA, matrix in shared
//                           memory up to nsta = 112, else in an L2/HBM workspace.
// All of them generate k on the fly (mesh descriptor) or read a k-list, build
// the lower triangle of H from the compiled plan, diagonalise, apply the
// Convention-I gauge, and write the reference's output layouts directly
// (solve_all's eval[band,k] / evec[band,k,orb] or the wf_array grid with its
// periodic images and the running minimum of the direct gaps).
#include <stdlib.h>
#include "tbk_internal.cuh"
#include "tbk_eig_small.cuh"
#include "tbk_eig_group.cuh"
#include "tbk_eig_blocked.cuh"

namespace tbk {

struct KSrc {
  const double* klist;  // [npts][dim_k], or nullptr -> mesh descriptor below
  int dim_k;
  int row0;             // offset added to the axis-0 mesh index (shards)
  int closing_g;        // global axis-0 index solved as the periodic image of row 0 (wrap0 == 2), or -1
  double start[TBK_MAX_DIM];
  double den[TBK_MAX_DIM];   // mesh[d]-1 as double
};

struct OutSpec {
  int mode;  // 0 = list (solve_all layouts), 1 = grid (wf_array slab)
  double* eval; long long ev_sb, ev_sk;
  cplx* evec;   long long vc_sb, vc_sk;
  int nd;
  int cnt[TBK_MAX_DIM];            // solved points per axis
  int full[TBK_MAX_DIM];           // storage extent per axis
  long long gstride[TBK_MAX_DIM];  // storage stride per axis, complex elements
  long long sstride;               // storage stride between the states (bands) of one mesh point (n: [k..., state, orb])
  int wrap[TBK_MAX_DIM];           // write the periodic image along this axis
  const cplx* pbc_phase;           // [nd][n]
  unsigned long long* gaps_bits;   // [n-1] running min of non-negative doubles, or null
};

// tb_model.k_uniform_mesh (pythtb.py:1848-1857): point (i_0/n_0, i_1/n_1, ...) at C-order index of (i_0, i_1, ...)
struct KMeshDesc {
  int nd;
  int n[TBK_MAX_DIM];
};
__global__ void __launch_bounds__(256)
kmesh_uniform_kernel(const KMeshDesc md, long long nk, double* __restrict__ k) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nk) return;
  long long rem = idx;
  double v[TBK_MAX_DIM];
#pragma unroll
  for (int d = TBK_MAX_DIM - 1; d >= 0; --d) {
    if (d < md.nd) {
      const long long q = rem / md.n[d];
      v[d] = (double)(rem - q * md.n[d]) / (double)md.n[d];
      rem = q;
    }
  }
#pragma unroll
  for (int d = 0; d < TBK_MAX_DIM; ++d)
    if (d < md.nd) k[idx * md.nd + d] = v[d];
}

__device__ __forceinline__ void decode_index(long long idx, const OutSpec& o, int mi[TBK_MAX_DIM]) {
#pragma unroll
  for (int d = TBK_MAX_DIM - 1; d >= 0; --d) {
    if (d < o.nd) {
      const long long q = idx / o.cnt[d];
      mi[d] = (int)(idx - q * o.cnt[d]);
      idx = q;
    } else {
      mi[d] = 0;
    }
  }
}

__device__ __forceinline__ void load_k(const KSrc& ks, long long idx, const int mi[TBK_MAX_DIM], double k[TBK_MAX_DIM]) {
#pragma unroll
  for (int d = 0; d < TBK_MAX_DIM; ++d) {
    if (d < ks.dim_k) {
      if (ks.klist) k[d] = ks.klist[idx * ks.dim_k + d];
      else {
        int g = mi[d] + (d == 0 ? ks.row0 : 0);
        if (d == 0 && g == ks.closing_g) g = 0;     // periodic image of row 0, phase applied at the store
        k[d] = ks.start[d] + (double)g / ks.den[d];                                     // pythtb.py:2477
      }
    } else {
      k[d] = 0.0;
    }
  }
}

__device__ __forceinline__ bool is_closing(const KSrc& ks, const int mi[TBK_MAX_DIM]) {
  return ks.klist == nullptr && ks.closing_g >= 0 && mi[0] + ks.row0 == ks.closing_g;
}

__device__ __forceinline__ void atomic_min_nonneg(unsigned long long* addr, double v) {
  atomicMin(addr, (unsigned long long)__double_as_longlong(v));
}

__global__ void fill_u64_kernel(unsigned long long* p, int n, unsigned long long v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace tbk
#include "tbk_mesh_small.cuh"
#include "tbk_eig_wy.cuh"
namespace tbk {

// ===========================================================================
// One k-point per thread, N = 2..4 (k-lists, and meshes the tiled kernel does not take)
// ===========================================================================
template <int N>
__global__ void __launch_bounds__(128)
solve_small_kernel(PlanView pv, KSrc ks, const cplx* __restrict__ hsrc, long long npts, OutSpec out,
                   int want_vec, int stage_plan) {
  constexpr int NP = N * (N + 1) / 2;
  extern __shared__ __align__(16) char smem[];
  // ---- stage the (tiny) plan in shared memory: every thread walks the same term list
  const int nterm = pv.nterm, nph = pv.nph, dk = pv.dim_k;
  const int* pm_ptr = pv.pm_ptr;
  const int* pm_pk = nullptr;          // packed lower index | conj flag
  const cplx* pm_amp = (const cplx*)pv.pm_amp;
  const double* ph_R = pv.ph_R;
  const double* tau = pv.tau;
  if (hsrc == nullptr && stage_plan) {
    cplx* s_amp = (cplx*)smem;
    double* s_R = (double*)(s_amp + nterm);
    double* s_tau = s_R + nph * dk;
    int* s_ptr = (int*)(s_tau + N * dk);
    int* s_pk = s_ptr + (nph + 2);
    for (int t = threadIdx.x; t < nterm; t += blockDim.x) {
      s_amp[t] = pm_amp[t];
      const int e = pv.pm_el[t];
      const int id = e & TBK_PH_MASK;
      const int r = pv.el_row[id], c = pv.el_col[id];
      s_pk[t] = (r * (r + 1) / 2 + c) | (e & TBK_PH_CONJ);
    }
    for (int t = threadIdx.x; t < nph * dk; t += blockDim.x) s_R[t] = ph_R[t];
    for (int t = threadIdx.x; t < N * dk; t += blockDim.x) s_tau[t] = tau[t];
    for (int t = threadIdx.x; t < nph + 2; t += blockDim.x) s_ptr[t] = pm_ptr[t];
    __syncthreads();
    pm_amp = s_amp; ph_R = s_R; tau = s_tau; pm_ptr = s_ptr; pm_pk = s_pk;
  }
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool active = idx < npts;
  int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
  double k[TBK_MAX_DIM] = {0.0, 0.0, 0.0, 0.0};
  double ev[N];
  cplx w[N][N];
#pragma unroll
  for (int b = 0; b < N; ++b) ev[b] = 0.0;
  if (active) {
    cplx acc[NP];
#pragma unroll
    for (int e = 0; e < NP; ++e) acc[e] = mk(0.0, 0.0);
    if (hsrc != nullptr) {
      const cplx* h = hsrc + idx * (long long)(N * N);
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) acc[r * (r + 1) / 2 + c] = h[r * N + c];
    } else {
      if (out.mode == 1) decode_index(idx, out, mi);
      load_k(ks, idx, mi, k);
      // phase-major accumulation: one sincospi per unique lattice vector
      for (int p = 0; p <= nph; ++p) {
        cplx z = mk(1.0, 0.0);
        if (p < nph) {
          double x = 0.0;
          for (int d = 0; d < dk; ++d) x = fma(k[d], ph_R[p * dk + d], x);
          z = expi_turns(x);
        }
        const int t1 = pm_ptr[p + 1];
        for (int t = pm_ptr[p]; t < t1; ++t) {
          int pk;
          if (pm_pk) pk = pm_pk[t];
          else {
            const int e = pv.pm_el[t];
            const int id = e & TBK_PH_MASK;
            const int r = pv.el_row[id], c = pv.el_col[id];
            pk = (r * (r + 1) / 2 + c) | (e & TBK_PH_CONJ);
          }
          const cplx a = pm_amp[t];
          cplx zz = z;
          if (pk & TBK_PH_CONJ) zz.im = -zz.im;
          pk &= TBK_PH_MASK;
#pragma unroll
          for (int e = 0; e < NP; ++e)
            if (pk == e) fma_acc(acc[e], a, zz);
        }
      }
    }
    // ---- diagonalise
    if constexpr (N == 2) {
      eigh2(acc[0].re, acc[2].re, acc[1], ev, w, want_vec != 0);
    } else {
      double dg[N];
      cplx lo[N * (N - 1) / 2];
#pragma unroll
      for (int r = 0; r < N; ++r) {
        dg[r] = acc[r * (r + 1) / 2 + r].re;
#pragma unroll
        for (int c = 0; c < r; ++c) lo[r * (r - 1) / 2 + c] = acc[r * (r + 1) / 2 + c];
      }
      eigh_small<N>(dg, lo, ev, w);
    }
    // ---- Convention I gauge: u_I[b][j] = conj(d_j) u_II[b][j], d_j = exp(2 pi i k.tau_j)
    if (want_vec && hsrc == nullptr && pv.convention == 1 && dk > 0) {
#pragma unroll
      for (int o = 0; o < N; ++o) {
        double x = 0.0;
        for (int d = 0; d < dk; ++d) x = fma(k[d], tau[o * dk + d], x);
        const cplx dj = expi_turns(x);
#pragma unroll
        for (int b = 0; b < N; ++b) w[b][o] = cmul(dj, w[b][o]);
      }
    }
    // ---- write
    if (out.mode == 0) {
      if (out.eval) {
#pragma unroll
        for (int b = 0; b < N; ++b) out.eval[b * out.ev_sb + idx * out.ev_sk] = ev[b];
      }
      if (want_vec) {
#pragma unroll
        for (int b = 0; b < N; ++b) {
          cplx* dst = out.evec + b * out.vc_sb + idx * out.vc_sk;
#pragma unroll
          for (int o = 0; o < N; ++o) dst[o] = w[b][o];
        }
      }
    } else {
      long long base = 0;
#pragma unroll
      for (int d = 0; d < TBK_MAX_DIM; ++d)
        if (d < out.nd) base += mi[d] * out.gstride[d];
      if (is_closing(ks, mi)) {
#pragma unroll
        for (int o = 0; o < N; ++o)
#pragma unroll
          for (int b = 0; b < N; ++b) w[b][o] = w[b][o] * out.pbc_phase[o];
      }
      cplx* dst = out.evec + base;
#pragma unroll
      for (int b = 0; b < N; ++b)
#pragma unroll
        for (int o = 0; o < N; ++o) dst[b * out.sstride + o] = w[b][o];
      // periodic images (impose_pbc, pythtb.py:2729-2747), every subset of the
      // axes on which this point sits at index 0
      int zero_mask = 0;
      for (int d = 0; d < out.nd; ++d)
        if (mi[d] == 0 && out.wrap[d]) zero_mask |= 1 << d;
      if (zero_mask) {
        for (int m = 1; m < (1 << out.nd); ++m) {
          if ((m & zero_mask) != m) continue;
          long long off = base;
          cplx im[N][N];
#pragma unroll
          for (int b = 0; b < N; ++b)
#pragma unroll
            for (int o = 0; o < N; ++o) im[b][o] = w[b][o];
          for (int d = 0; d < out.nd; ++d) {         // one multiply per wrapped axis, in axis order
            if (m & (1 << d)) {
              off += (long long)(out.full[d] - 1) * out.gstride[d];
#pragma unroll
              for (int o = 0; o < N; ++o)
#pragma unroll
                for (int b = 0; b < N; ++b) im[b][o] = im[b][o] * out.pbc_phase[d * N + o];
            }
          }
          cplx* dsti = out.evec + off;
#pragma unroll
          for (int b = 0; b < N; ++b)
#pragma unroll
            for (int o = 0; o < N; ++o) dsti[b * out.sstride + o] = im[b][o];
        }
      }
    }
  }
  // ---- running minimum of the direct gaps (pythtb.py:2484, 2529-2530)
  if (out.mode == 1 && out.gaps_bits != nullptr) {
#pragma unroll
    for (int b = 0; b < N - 1; ++b) {
      double g = active ? (ev[b + 1] - ev[b]) : INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g = fmin(g, __shfl_xor_sync(0xffffffffu, g, o));
      if ((threadIdx.x & 31) == 0) atomic_min_nonneg(out.gaps_bits + b, g);
    }
  }
}

// ===========================================================================
// Eigenvalues only, N = 5..8 (band-structure sweeps of small Wannier models: BASELINE configs[2], silicon on 256^3):
// one k-point per thread, eigensolver in registers (eigvals_small<N>).
// H(k), element-major: a thread first tabulates E_p = exp(2 pi i k.R_p) of its k-point for every unique lattice vector in
// its own column of a shared-memory table ([nph][128]: conflict-free), then accumulates every lower-triangle element in a
// REGISTER over that element's term list (amplitude and phase index are warp-uniform read-only loads, batched by the
// unrolled loop; one 16-byte shared-memory read of the phase per term).  The element loop is unrolled over the
// compile-time (row, column) pairs so that the matrix never leaves registers.  (A phase-major version that accumulated
// into shared memory was measured 3x slower: every term was a load - FMA - store round trip behind an L2-latency load.)
// The cooperative solver this replaces (8 lanes per matrix, matrix in shared memory, a barrier per phase of
// the Householder / QL steps) spent ~2600 SM cycles per n = 8 matrix in the eigensolver alone (profiles/split_cfg3.py).
// ===========================================================================
constexpr int kRegThreads = 128;
template <int N>
__global__ void __launch_bounds__(kRegThreads)
solve_reg_kernel(PlanView pv, KSrc ks, const cplx* __restrict__ hsrc, long long npts, OutSpec out) {
  constexpr int NP = N * (N + 1) / 2;
  extern __shared__ __align__(16) char smem[];
  cplx* ph = (cplx*)smem;                              // [nph][kRegThreads]: phases of this thread's k-point in column tid
  int* elmap = (int*)(ph + (size_t)(hsrc ? 0 : pv.nph) * kRegThreads);   // [NP] plan element of the packed lower index, or -1
  const int tid = threadIdx.x;
  if (hsrc == nullptr) {
    for (int e = tid; e < NP; e += kRegThreads) elmap[e] = -1;
    __syncthreads();
    for (int e = tid; e < pv.nel; e += kRegThreads) { const int r = pv.el_row[e], c = pv.el_col[e]; elmap[r * (r + 1) / 2 + c] = e; }
    __syncthreads();
  }
  const cplx* __restrict__ t_amp = (const cplx*)pv.t_amp;
  const int* __restrict__ t_ph = pv.t_ph;
  for (long long base = (long long)blockIdx.x * kRegThreads; base < npts; base += (long long)gridDim.x * kRegThreads) {
    const long long idx = base + tid;
    const bool active = idx < npts;
    cplx a[N][N];
    if (hsrc != nullptr) {
      if (active) {
        const cplx* h = hsrc + idx * (long long)(N * N);
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) a[r][c] = h[r * N + c];
      }
    } else if (active) {
      int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
      double k[TBK_MAX_DIM] = {0.0, 0.0, 0.0, 0.0};
      load_k(ks, idx, mi, k);
      for (int p = 0; p < pv.nph; ++p) {
        double x = 0.0;
        for (int d = 0; d < pv.dim_k; ++d) x = fma(k[d], pv.ph_R[p * pv.dim_k + d], x);
        ph[p * kRegThreads + tid] = expi_turns(x);
      }
#pragma unroll
      for (int r = 0; r < N; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c) {
          cplx acc = mk(0.0, 0.0);
          const int pe = elmap[r * (r + 1) / 2 + c];
          if (pe >= 0) {
            const int t1 = pv.el_ptr[pe + 1];
#pragma unroll 4
            for (int t = pv.el_ptr[pe]; t < t1; ++t) {
              const int p = __ldg(t_ph + t);
              const double2 av = __ldg(reinterpret_cast<const double2*>(t_amp + t));
              cplx z = mk(1.0, 0.0);
              if (p >= 0) {
                z = ph[(p & TBK_PH_MASK) * kRegThreads + tid];
                if (p & TBK_PH_CONJ) z.im = -z.im;
              }
              fma_acc(acc, mk(av.x, av.y), z);
            }
          }
          a[r][c] = acc;
        }
      }
    }
    if (active) {
#pragma unroll
      for (int r = 0; r < N; ++r) a[r][r].im = 0.0;
      double ev[N];
      const bool ok = eigvals_small<N>(a, ev);
      if (out.eval) {
#pragma unroll
        for (int b = 0; b < N; ++b) out.eval[b * out.ev_sb + idx * out.ev_sk] = ok ? ev[b] : NAN;   // not converged: poisoned
      }
    }
  }
}

// ===========================================================================
// The same sweep with the Hamiltonian assembled on the FP64 TENSOR PIPE.  With the phases as cos / sin pairs the
// assembly is a real matrix product,  [Re H; Im H] (2 NP x k-points) = A (2 NP x (2 nph + 1)) . [cos; sin; 1],  A being the
// model's dense coefficient table (tbk_api.cu, stored in mma.sync.m8n8k4 A-fragment order: one coalesced 256-byte load
// per fragment).  A CTA of 4 warps handles 128 k-points per pass; every thread tabulates the cos / sin column of ITS
// k-point in shared memory (B operand, leading dimension 132 = 4 mod 16: conflict-free 8-byte fragment reads), every WARP
// multiplies the table into the 32 columns of its own threads (MT x 4 accumulator tiles in registers), writes the result
// back over those columns and each thread picks its matrix up from its column — the only synchronisation is __syncwarp.
// A scalar assembly reads one 16-byte phase from shared memory per term: 53 KB per k-point for the silicon model
// (3348 terms), which alone costs the 1.5 ms per 2^20 k-points that the whole kernel now takes; here the phases are read
// 768 bytes per k-point and the 27 DMMA per k-point take ~110 SM cycles.
// ===========================================================================
constexpr int kRegLDB = 132;
// VEC: eigenvectors too (solve_all with eig_vectors, solve_on_grid of 5..8-band models): the real rotation matrix of the
// QL iteration lives in the thread's own column of the tile (free once the matrix has been picked up), the eigenvectors
// are back-transformed one at a time from the reflectors in registers and written with the Convention-I gauge factors,
// periodic images and closing-row factor (blk_store_vec); grid solves also reduce the minimal direct gaps.
template <int N, bool VEC>
__global__ void __launch_bounds__(kRegThreads)
solve_reg_gemm_kernel(PlanView pv, KSrc ks, long long npts, OutSpec out, const double* __restrict__ tab, int KS) {
  constexpr int NP = N * (N + 1) / 2;
  constexpr int MT = (2 * NP + 7) / 8;
  extern __shared__ __align__(16) char smem[];
  double* Bs = (double*)smem;                          // [max(4 KS, 8 MT)][kRegLDB]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int nph = pv.nph;
  // (the lattice vectors stay in global memory: staged in shared memory, the warp-uniform reads of the phase loop made
  // the kernel 16 % slower — 1.85 instead of 1.59 ms per 2^20 k-points)
  for (long long base = (long long)blockIdx.x * kRegThreads; base < npts; base += (long long)gridDim.x * kRegThreads) {
    const bool active = base + tid < npts;
    const long long idx = active ? base + tid : npts - 1;         // idle lanes repeat the last point (they take part in the MMAs)
    {
      int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
      double k[TBK_MAX_DIM] = {0.0, 0.0, 0.0, 0.0};
      if (out.mode == 1) decode_index(idx, out, mi);
      load_k(ks, idx, mi, k);
      for (int p = 0; p < nph; ++p) {
        double x = 0.0;
        for (int d = 0; d < pv.dim_k; ++d) x = fma(k[d], pv.ph_R[p * pv.dim_k + d], x);
        const cplx z = expi_turns(x);
        Bs[(2 * p) * kRegLDB + tid] = z.re;
        Bs[(2 * p + 1) * kRegLDB + tid] = z.im;
      }
      Bs[(2 * nph) * kRegLDB + tid] = 1.0;
      for (int kk = 2 * nph + 1; kk < 4 * KS; ++kk) Bs[kk * kRegLDB + tid] = 0.0;
    }
    __syncwarp();
    double acc[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[mt][j][0] = 0.0; acc[mt][j][1] = 0.0; }
    const double* bcol = Bs + 32 * warp + g;
#pragma unroll 2
    for (int ksi = 0; ksi < KS; ++ksi) {
      double afr[MT], bfr[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) afr[mt] = __ldg(tab + ((size_t)(mt * KS + ksi) * 32 + lane));
#pragma unroll
      for (int j = 0; j < 4; ++j) bfr[j] = bcol[(ksi * 4 + q) * kRegLDB + 8 * j];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j) blk_dmma(acc[mt][j][0], acc[mt][j][1], afr[mt], bfr[j]);
    }
    __syncwarp();                                      // every lane of the warp is done with the warp's columns of Bs
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double* dst = Bs + (mt * 8 + g) * kRegLDB + 32 * warp + 8 * j + 2 * q;
        dst[0] = acc[mt][j][0];
        dst[1] = acc[mt][j][1];
      }
    __syncwarp();
    cplx a[N][N];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) {
        const int pk = r * (r + 1) / 2 + c;
        a[r][c] = mk(Bs[pk * kRegLDB + tid], c == r ? 0.0 : Bs[(NP + pk) * kRegLDB + tid]);
      }
    __syncwarp();                                      // the columns are free for the next pass
    double ev[N];
    if constexpr (!VEC) {
      const bool ok = eigvals_small<N>(a, ev);
      if (active && out.eval) {
#pragma unroll
        for (int b = 0; b < N; ++b) out.eval[b * out.ev_sb + idx * out.ev_sk] = ok ? ev[b] : NAN;
      }
    } else {
      int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
      double k[TBK_MAX_DIM] = {0.0, 0.0, 0.0, 0.0};
      if (out.mode == 1) decode_index(idx, out, mi);
      load_k(ks, idx, mi, k);
      // gauge factors conj(d_o(k)) in rows N^2 .. N^2 + 2N - 1 of the thread's column (registers are scarce here)
      double* gfs = Bs + (size_t)(N * N) * kRegLDB + tid;
#pragma unroll
      for (int o = 0; o < N; ++o) {
        const cplx f = (pv.convention == 1 && pv.dim_k > 0) ? conj(plan_gauge(pv, k, o)) : mk(1.0, 0.0);
        gfs[(2 * o) * kRegLDB] = f.re;
        gfs[(2 * o + 1) * kRegLDB] = f.im;
      }
      BlkStorePoint pt;
      pt.idx = idx; pt.base = 0; pt.zero_mask = 0; pt.closing = false;
      if (out.mode == 1) {
        pt.closing = is_closing(ks, mi);
        for (int d = 0; d < out.nd; ++d) {
          pt.base += mi[d] * out.gstride[d];
          if (mi[d] == 0 && out.wrap[d]) pt.zero_mask |= 1 << d;
        }
      }
      // Grid output (k-major _wfs: a band's N components are one 16 N-byte run, the bands of a mesh point N of them, the
      // points of a warp 16 N^2 bytes apart): written thread by thread, a store instruction touched 32 different lines.
      // Unless a lane sits on a periodic image or the closing row (those go the general way), a band is therefore staged
      // in rows N^2 + 2N .. N^2 + 4N - 1 of the warp's columns and written out with N consecutive lanes per mesh point.
      const bool coalesce = out.mode == 1 && !__any_sync(0xffffffffu, pt.zero_mask != 0 || pt.closing);
      double* stg = Bs + (size_t)(N * N + 2 * N) * kRegLDB;
      // (the rotation matrix goes into this thread's own column: no other lane reads or writes it until the next pass)
      const bool ok = eigh_small_mem<N>(a, ev, Bs + tid, kRegLDB, [&](int b, const cplx (&x)[N]) {
        if (!coalesce) {
          if (active) {
#pragma unroll
            for (int o = 0; o < N; ++o) blk_store_vec(out, N, pt, b, o, x[o] * mk(gfs[(2 * o) * kRegLDB], gfs[(2 * o + 1) * kRegLDB]));
          }
          return;
        }
#pragma unroll
        for (int o = 0; o < N; ++o) {
          const cplx v = x[o] * mk(gfs[(2 * o) * kRegLDB], gfs[(2 * o + 1) * kRegLDB]);
          stg[(2 * o) * kRegLDB + tid] = v.re;
          stg[(2 * o + 1) * kRegLDB + tid] = v.im;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const int e = j * 32 + lane, kp = e / N, o = e - kp * N;
          const long long base_kp = __shfl_sync(0xffffffffu, pt.base, kp);
          const int act_kp = __shfl_sync(0xffffffffu, (int)active, kp);
          const double re = stg[(2 * o) * kRegLDB + 32 * warp + kp], im = stg[(2 * o + 1) * kRegLDB + 32 * warp + kp];
          if (act_kp) out.evec[base_kp + (long long)b * out.sstride + o] = mk(re, im);
        }
        __syncwarp();
      });
      if (out.mode == 0) {
        if (active && out.eval) {
#pragma unroll
          for (int b = 0; b < N; ++b) out.eval[b * out.ev_sb + idx * out.ev_sk] = ok ? ev[b] : NAN;
        }
      } else if (out.gaps_bits != nullptr) {
#pragma unroll
        for (int b = 0; b < N - 1; ++b) {
          double gp = active ? (ok ? ev[b + 1] - ev[b] : 0.0) : INFINITY;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) gp = fmin(gp, __shfl_xor_sync(0xffffffffu, gp, o));
          if (lane == 0) atomic_min_nonneg(out.gaps_bits + b, gp);
        }
      }
      __syncwarp();                                    // the rotation matrices are done with before the next pass
    }
  }
}

// ===========================================================================
// One k-point per thread group, matrix in shared memory or in a workspace
// ===========================================================================
struct GroupShape {
  int n, lda, nph;
  size_t off_scr, off_ph, off_rank, off_k, region;   // byte offsets inside a per-matrix region
  bool a_in_smem;
};

static GroupShape group_shape(int n, int nph, bool a_in_smem) {
  GroupShape s;
  s.n = n; s.lda = n | 1; s.nph = nph; s.a_in_smem = a_in_smem;
  size_t off = a_in_smem ? (size_t)n * s.lda * 16 : 0;
  s.off_scr = off;  off += (eig_scratch_bytes(n) + 15) & ~(size_t)15;
  s.off_ph = off;   off += (size_t)(nph > 0 ? nph : 1) * 16;
  s.off_rank = off; off += ((size_t)n * 4 + 15) & ~(size_t)15;
  s.off_k = off;    off += 64;
  s.region = off;
  return s;
}

template <class G>
__device__ void solve_one_matrix(G& g, const PlanView& pv, const KSrc& ks, const cplx* __restrict__ hsrc,
                                 long long idx, const OutSpec& out, int want_vec, const GroupShape& gs,
                                 cplx* A, char* region) {
  const int n = gs.n, lda = gs.lda;
  EigScratch s = eig_scratch_carve(region + gs.off_scr, n);
  cplx* ph = (cplx*)(region + gs.off_ph);
  int* rank = (int*)(region + gs.off_rank);
  double* kbuf = (double*)(region + gs.off_k);
  int* mibuf = (int*)(kbuf + TBK_MAX_DIM);
  // ---- k-point
  if (g.tid() == 0) {
    int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
    double k[TBK_MAX_DIM];
    if (out.mode == 1) decode_index(idx, out, mi);
    load_k(ks, idx, mi, k);
    for (int d = 0; d < TBK_MAX_DIM; ++d) { kbuf[d] = k[d]; mibuf[d] = mi[d]; }
  }
  for (int i = g.tid(); i < n * lda; i += g.size()) A[i] = mk(0.0, 0.0);
  g.sync();
  // ---- lower triangle of H
  if (hsrc != nullptr) {
    const cplx* h = hsrc + idx * (long long)n * n;
    for (int q = g.tid(); q < n * n; q += g.size()) {
      const int r = q / n, c = q - r * n;
      if (c <= r) A[r + (size_t)c * lda] = h[q];
    }
  } else {
    for (int p = g.tid(); p < pv.nph; p += g.size()) {
      double x = 0.0;
      for (int d = 0; d < pv.dim_k; ++d) x = fma(kbuf[d], pv.ph_R[p * pv.dim_k + d], x);
      ph[p] = expi_turns(x);
    }
    g.sync();
    for (int e = g.tid(); e < pv.nel; e += g.size())
      A[pv.el_row[e] + (size_t)pv.el_col[e] * lda] = plan_element(pv, e, ph, 1);
  }
  g.sync();
  // ---- diagonalise
  const int info = heev_group(g, n, A, lda, s, want_vec != 0);
  eig_rank(g, n, s.d, rank);
  if (info != 0) {   // not converged: poison the eigenvalues so that it cannot go unnoticed
    for (int i = g.tid(); i < n; i += g.size()) s.d[i] = NAN;
    g.sync();
  }
  // ---- gauge factors conj(d_o) -> s.work, sorted eigenvalues -> s.e
  for (int o = g.tid(); o < n; o += g.size()) {
    cplx f = mk(1.0, 0.0);
    if (want_vec && hsrc == nullptr && pv.convention == 1 && pv.dim_k > 0) f = conj(plan_gauge(pv, kbuf, o));
    s.work[o] = f;
    s.e[rank[o]] = s.d[o];
  }
  g.sync();
  // ---- write
  if (out.mode == 0) {
    if (out.eval)
      for (int b = g.tid(); b < n; b += g.size()) out.eval[b * out.ev_sb + idx * out.ev_sk] = s.e[b];
    if (want_vec) {
      for (int q = g.tid(); q < n * n; q += g.size()) {
        const int i = q / n, o = q - i * n;
        out.evec[rank[i] * out.vc_sb + idx * out.vc_sk + o] = A[o + (size_t)i * lda] * s.work[o];
      }
    }
  } else {
    long long base = 0;
    int zero_mask = 0;
    const bool closing = is_closing(ks, mibuf);
    for (int d = 0; d < out.nd; ++d) {
      base += mibuf[d] * out.gstride[d];
      if (mibuf[d] == 0 && out.wrap[d]) zero_mask |= 1 << d;
    }
    for (int q = g.tid(); q < n * n; q += g.size()) {
      const int i = q / n, o = q - i * n;
      cplx v = A[o + (size_t)i * lda] * s.work[o];
      if (closing) v = v * out.pbc_phase[o];      // image of global row 0 (same two-step rounding as the unsharded image)
      const long long at = (long long)rank[i] * out.sstride + o;
      out.evec[base + at] = v;
      if (zero_mask) {
        for (int m = 1; m < (1 << out.nd); ++m) {
          if ((m & zero_mask) != m) continue;
          long long off = base;
          cplx f = v;
          for (int d = 0; d < out.nd; ++d)
            if (m & (1 << d)) {
              off += (long long)(out.full[d] - 1) * out.gstride[d];
              f = f * out.pbc_phase[d * n + o];
            }
          out.evec[off + at] = f;
        }
      }
    }
    if (out.gaps_bits != nullptr)
      for (int b = g.tid(); b < n - 1; b += g.size()) atomic_min_nonneg(out.gaps_bits + b, s.e[b + 1] - s.e[b]);
  }
  g.sync();
}

template <int G>
__global__ void __launch_bounds__(128)
solve_tile_kernel(PlanView pv, KSrc ks, const cplx* __restrict__ hsrc, long long npts, OutSpec out,
                  int want_vec, GroupShape gs) {
  extern __shared__ __align__(16) char smem[];
  constexpr int MATS = 128 / G;
  TileGroup<G> g;
  const int sub = threadIdx.x / G;
  char* region = smem + (size_t)sub * gs.region;
  cplx* A = (cplx*)region;
  for (long long idx = (long long)blockIdx.x * MATS + sub; idx < npts; idx += (long long)gridDim.x * MATS)
    solve_one_matrix(g, pv, ks, hsrc, idx, out, want_vec, gs, A, region);
}

__global__ void __launch_bounds__(256)
solve_block_kernel(PlanView pv, KSrc ks, const cplx* __restrict__ hsrc, long long npts, OutSpec out,
                   int want_vec, GroupShape gs, cplx* gA) {
  extern __shared__ __align__(16) char smem[];
  __shared__ double red[32];
  BlockGroup g(red);
  cplx* A = gs.a_in_smem ? (cplx*)smem : gA + (size_t)blockIdx.x * gs.n * gs.lda;
  for (long long idx = blockIdx.x; idx < npts; idx += gridDim.x)
    solve_one_matrix(g, pv, ks, hsrc, idx, out, want_vec, gs, A, smem);
}

// ===========================================================================
// nsta > 32, blocked solver (tbk_eig_blocked.cuh): one k-point per CTA, H(k) and the
// tridiagonal eigenvectors in an L2/HBM workspace, panels and all O(n) data in shared memory
// ===========================================================================
struct BlkShape {
  int n, lda, nb, nt, threads, nph, nred, sym;
  size_t off_ph, off_gf, off_misc, smem;             // shared-memory layout after the BlkWork part
  size_t ws_A, ws_Z, ws_lu, ws_block;                // per-CTA global workspace, bytes
  GroupShape fallback;                               // unblocked solver's layout inside the same shared buffer
};

static BlkShape blk_shape(int n, int nph) {
  BlkShape s;
  s.n = n; s.lda = n | 1; s.nph = nph;
  s.threads = n <= 256 ? 256 : 512;
  const int nwarps = s.threads / 32;
  // tridiagonalisation variant: TBK_HETRD=sym reads / updates the lower triangle only (half the DRAM traffic),
  // TBK_HETRD=full (default: measured faster on B200, profiles/README.md r09) streams the full trailing matrix
  {
    static int sym = -1;
    if (sym < 0) { const char* e = getenv("TBK_HETRD"); sym = (e && strcmp(e, "sym") == 0) ? 1 : 0; }
    s.sym = sym;
  }
  auto total = [&](int nb, int nred) {
    size_t off = (blk_shared_bytes(n, nb, nred, s.threads) + 15) & ~(size_t)15;
    off += (size_t)(nph > 0 ? nph : 1) * 16 + (size_t)n * 16 + 128;
    return off;
  };
  // panel width: 16 if two CTAs still fit an SM, else 8.  (Measured at n = 400: nb = 4 with two resident
  // CTAs is 15 % slower than nb = 8 with one — the stage is bandwidth-, not latency-bound.)
  // nred (lower-triangle variant only): how many warps' row partials of the symmetric matrix-vector product are
  // summed per round through shared memory — all of them when they fit, fewer (more rounds) for the largest matrices.
  const int want_red = s.sym ? nwarps : 0;
  if (total(16, want_red) * 2 + 2048 <= (size_t)kMaxSmem) { s.nb = 16; s.nred = want_red; }
  else {
    s.nb = 8; s.nred = want_red;
    while (s.nred > 1 && total(8, s.nred) + 1024 > (size_t)kMaxSmem) --s.nred;
  }
  {
    // A/B knob TBK_BLK_SHAPE="threads,nb" (n > 256 only), e.g. "256,4": two 256-thread CTAs per SM with half-width
    // panels, so that one matrix' latency-bound stages overlap the other's HBM-bound matrix-vector products
    static int kt = -1, knb = 0;
    if (kt < 0) {
      kt = 0;
      const char* e = getenv("TBK_BLK_SHAPE");
      if (e) { int a = 0, b = 0; if (sscanf(e, "%d,%d", &a, &b) == 2 && (a == 256 || a == 512) && (b == 4 || b == 8 || b == 16)) { kt = a; knb = b; } }
    }
    static int ksmall = -1;                             // TBK_BLK_SHAPE_ALL=1: the knob also applies to n <= 256
    if (ksmall < 0) { const char* e = getenv("TBK_BLK_SHAPE_ALL"); ksmall = (e && atoi(e) == 1) ? 1 : 0; }
    if (kt > 0 && (n > 256 || ksmall) && !s.sym && total(knb, 0) + 1024 <= (size_t)kMaxSmem) { s.threads = kt; s.nb = knb; s.nred = 0; }
  }
  size_t off = (blk_shared_bytes(n, s.nb, s.nred, s.threads) + 15) & ~(size_t)15;
  s.off_ph = off;   off += (size_t)(nph > 0 ? nph : 1) * 16;
  s.off_gf = off;   off += (size_t)n * 16;
  s.off_misc = off; off += 128;
  s.fallback = group_shape(n, nph, false);
  s.smem = off > s.fallback.region ? off : s.fallback.region;
  s.nt = ((n + 31) / 32) * 32;
  if (s.nt > s.threads) s.nt = s.threads;
  s.ws_A = ((size_t)n * s.lda * 16 + 255) & ~(size_t)255;
  s.ws_Z = ((size_t)n * n * 8 + 255) & ~(size_t)255;
  s.ws_lu = ((size_t)4 * n * s.nt * 8 + 255) & ~(size_t)255;
  s.ws_block = s.ws_A + s.ws_Z + s.ws_lu;
  return s;
}

static long long blk_blocks(const BlkShape& s, long long npts) {
  int per_sm = (int)((size_t)kMaxSmem / (s.smem + 1024));
  if (per_sm < 1) per_sm = 1;
  const int by_threads = 2048 / s.threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > 4) per_sm = 4;
  const long long cap = (long long)kNumSM * per_sm;
  return npts < cap ? npts : cap;
}

// Staged solver (eigenvectors wanted): the matrices of a chunk keep their reflectors (A, tau), tridiagonal eigenvectors
// (Z) and compact-WY T factors in per-matrix SLOTS of the workspace; solve_blocked_kernel fills the slots (H build,
// tridiagonalisation, bisection, inverse iteration), blk_wy_kernel and blk_backtransform_kernel (tbk_eig_wy.cuh) finish
// the chunk on the FP64 tensor pipe.  The inverse-iteration scratch stays per resident CTA.
constexpr int kBlkSlotsPerSM = 4;
struct BlkStage {
  WyArgs wy;            // slot layout (ws filled in at launch)
  long long slots;      // matrices per chunk
  long long ctas;       // resident CTAs of the front kernel
  size_t lu_bytes;      // inverse-iteration scratch per front-kernel CTA
  size_t total;
};
static BlkStage blk_stage_layout(const BlkShape& shp, long long npts) {
  BlkStage s;
  memset(&s, 0, sizeof(s));
  const int n = shp.n;
  s.wy.n = n; s.wy.lda = shp.lda; s.wy.nblk = wy_nblk(n);
  auto r256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  s.wy.off_A = off;    off += shp.ws_A;
  s.wy.off_Z = off;    off += shp.ws_Z;
  s.wy.off_tau = off;  off += r256((size_t)n * 16);
  s.wy.off_T = off;    off += r256((size_t)s.wy.nblk * kWyNB * kWyNB * 16);
  s.wy.off_flag = off; off += 256;
  s.wy.slot_bytes = off;
  // matrices per SM and chunk: as many as ~12 GB of slots allow, between 4 and 16 (TBK_BLK_SLOTS overrides).  A chunk
  // ends with the tail of its slowest matrices (up to 3x the mean) and three launches; measured on the norb-499 slab
  // over a [129, 129] mesh: 5.75 / 5.10 / 4.78 s at 4 / 8 / 16 matrices per SM (profiles/README.md r15).
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("TBK_BLK_SLOTS"); forced = e ? atoi(e) : 0; if (forced < 0 || forced > 32) forced = 0; }
  int per_sm = forced;
  if (per_sm == 0) {
    const double budget = 12.0e9;
    per_sm = (int)(budget / ((double)kNumSM * (double)off));
    if (per_sm < kBlkSlotsPerSM) per_sm = kBlkSlotsPerSM;
    if (per_sm > 16) per_sm = 16;
  }
  const long long cap = (long long)kNumSM * per_sm;
  // equal chunks: 2048 matrices with room for 1776 are two chunks of 1024, not 1776 + a 272-matrix chunk that leaves
  // most SMs idle behind its slowest matrix (a shard of the [129, 129] slab mesh on 8 GPUs)
  const long long want = npts < 1 ? 1 : npts;
  const long long nchunk = (want + cap - 1) / cap;
  s.slots = (want + nchunk - 1) / nchunk;
  s.ctas = blk_blocks(shp, s.slots);
  s.lu_bytes = shp.ws_lu;
  s.total = (size_t)s.slots * s.wy.slot_bytes + (size_t)s.ctas * s.lu_bytes;
  return s;
}
static bool blk_staged_enabled() {
  static int on = -1;                                   // TBK_BACKTR=legacy: the one-kernel solver of rounds 1-2 (A/B knob)
  if (on < 0) { const char* e = getenv("TBK_BACKTR"); on = (e && strcmp(e, "legacy") == 0) ? 0 : 1; }
  return on == 1;
}

// Optional per-stage cycle counters (TBK_PROF=1): pinned host words the kernel adds clock64 deltas to
// [0] hetrd [1] bisect [2] invit [3] backtransform [4] matrices [5] fallbacks [6] slowest matrix;
// read by tbk_debug_profile.
static unsigned long long* g_blk_prof = nullptr;
static unsigned long long* blk_prof() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("TBK_PROF"); on = (e && atoi(e) == 1) ? 1 : 0; }
  if (!on) return nullptr;
  if (!g_blk_prof) {
    if (cudaMallocHost((void**)&g_blk_prof, 8 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    memset(g_blk_prof, 0, 8 * sizeof(unsigned long long));
  }
  return g_blk_prof;
}

// resident CTAs per SM the 256-thread instantiations (n <= 256) are compiled for (A/B knob, profiles/build_variant.py)
#ifndef TBK_BLK_MINB_SMALL
#define TBK_BLK_MINB_SMALL 2
#endif
template <int MAXM>
__global__ void __launch_bounds__(MAXM <= 8 ? 256 : 512, MAXM <= 8 ? TBK_BLK_MINB_SMALL : 1)
solve_blocked_kernel(PlanView pv, KSrc ks, const cplx* __restrict__ hsrc, long long npts, OutSpec out, int want_vec,
                     BlkShape shp, char* __restrict__ gws, unsigned long long* prof, const WyArgs stg, const long long idx0,
                     unsigned* __restrict__ work) {
  extern __shared__ __align__(16) char smem[];
  __shared__ double red[32];
  BlockGroup g(red);
  const int n = shp.n, lda = shp.lda, tid = threadIdx.x, T = blockDim.x;
  BlkWork w;
  w.n = n; w.lda = lda; w.nb = shp.nb; w.nt = shp.nt; w.nred = shp.nred;
  w.prof = prof;
  blk_carve_shared(w, smem, shp.threads);
  cplx* ph = (cplx*)(smem + shp.off_ph);
  cplx* gf = (cplx*)(smem + shp.off_gf);
  double* kbuf = (double*)(smem + shp.off_misc);
  int* mibuf = (int*)(kbuf + TBK_MAX_DIM);
  // stg.ws == nullptr: one kernel does everything, the CTA's private workspace block holds A, Z and the
  // inverse-iteration scratch.  Otherwise (staged, eigenvectors wanted): this kernel handles the matrices
  // idx0 .. idx0 + npts - 1 of a chunk, A / Z / tau / the fallback flag go to the matrix' slot, gws holds the per-CTA scratch.
  const bool staged = stg.ws != nullptr;
  char* mine = gws + (size_t)blockIdx.x * (staged ? shp.ws_lu : shp.ws_block);
  if (staged) {
    w.lu = (double*)mine;
  } else {
    w.A = (cplx*)mine;
    w.Z = (double*)(mine + shp.ws_A);
    w.lu = (double*)(mine + shp.ws_A + shp.ws_Z);
  }
  // work distribution: static stride, or (staged chunks) a self-resetting device counter — matrices with large
  // eigenvalue clusters take up to 3x the mean, and with 4 matrices per CTA a static split made the unlucky CTA the
  // tail of every chunk.  atomicInc wraps to 0 after npts successful + gridDim failing fetches: ready for the next launch.
  __shared__ long long s_next;
  auto fetch = [&](long long cur) -> long long {
    if (work == nullptr) return cur < 0 ? (long long)blockIdx.x : cur + gridDim.x;
    __syncthreads();
    if (tid == 0) s_next = (long long)atomicInc(work, (unsigned)(npts + gridDim.x - 1));
    __syncthreads();
    return s_next;
  };
  for (long long it = fetch(-1); it < npts; it = fetch(it)) {
    const long long idx = idx0 + it;
    char* slot = nullptr;
    if (staged) {
      slot = stg.ws + (size_t)it * stg.slot_bytes;
      w.A = (cplx*)(slot + stg.off_A);
      w.Z = (double*)(slot + stg.off_Z);
    }
    // ---- k-point and the lower triangle of H(k)
    if (tid == 0) {
      int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
      double k[TBK_MAX_DIM];
      if (out.mode == 1) decode_index(idx, out, mi);
      load_k(ks, idx, mi, k);
      for (int d = 0; d < TBK_MAX_DIM; ++d) { kbuf[d] = k[d]; mibuf[d] = mi[d]; }
    }
    for (int i = tid; i < n * lda; i += T) w.A[i] = mk(0.0, 0.0);
    g.sync();
    if (hsrc != nullptr) {
      const cplx* h = hsrc + idx * (long long)n * n;
      for (int q = tid; q < n * n; q += T) {
        const int r = q / n, c = q - r * n;
        if (c <= r) w.A[r + (size_t)c * lda] = h[q];
      }
    } else {
      for (int p = tid; p < pv.nph; p += T) {
        double x = 0.0;
        for (int d = 0; d < pv.dim_k; ++d) x = fma(kbuf[d], pv.ph_R[p * pv.dim_k + d], x);
        ph[p] = expi_turns(x);
      }
      g.sync();
      for (int e = tid; e < pv.nel; e += T) w.A[pv.el_row[e] + (size_t)pv.el_col[e] * lda] = plan_element(pv, e, ph, 1);
    }
    g.sync();
    long long t0 = prof ? clock64() : 0;
    const long long tstart = t0;
#define TBK_PROF_MARK(slot) if (prof && tid == 0) { const long long t1 = clock64(); atomicAdd(prof + slot, (unsigned long long)(t1 - t0)); t0 = t1; }
    if (shp.sym) hetrd_blocked<MAXM>(g, w);
    else hetrd_blocked_full(g, w);
#if defined(TBK_HETRD_PROF)
    if (prof && tid == 0) t0 = clock64();       // the phases of the tridiagonalisation were counted inside (slots 0, 3, 7)
#else
    TBK_PROF_MARK(0)
#endif
    const double tnorm = tridiag_bisect(g, w);
    TBK_PROF_MARK(1)
    if (want_vec) {
      const int fail = tridiag_invit(g, w, tnorm);
      TBK_PROF_MARK(2)
      if (staged && tid == 0) *(int*)(slot + stg.off_flag) = fail ? 1 : 0;
      if (fail) {
        if (prof && tid == 0) atomicAdd(prof + 5, 1ull);
        // a spectrum the inverse iteration should not be trusted with: unblocked Householder + QL
        // (rebuilds H(k), writes every output of this k-point)
        solve_one_matrix(g, pv, ks, hsrc, idx, out, want_vec, shp.fallback, w.A, smem);
        continue;
      }
      if (staged) {
        // the back-transformation is left to blk_wy_kernel / blk_backtransform_kernel
        cplx* tau_g = (cplx*)(slot + stg.off_tau);
        for (int o = tid; o < n; o += T) tau_g[o] = w.tau[o];
        if (prof && tid == 0) { atomicAdd(prof + 4, 1ull); atomicMax(prof + 6, (unsigned long long)(clock64() - tstart)); }
        if (out.mode == 0) {
          if (out.eval)
            for (int b = tid; b < n; b += T) out.eval[b * out.ev_sb + idx * out.ev_sk] = w.lam[b];
        } else if (out.gaps_bits != nullptr) {
          for (int b = tid; b < n - 1; b += T) atomic_min_nonneg(out.gaps_bits + b, w.lam[b + 1] - w.lam[b]);
        }
        g.sync();
        continue;
      }
      for (int o = tid; o < n; o += T) {
        cplx f = mk(1.0, 0.0);
        if (hsrc == nullptr && pv.convention == 1 && pv.dim_k > 0) f = conj(plan_gauge(pv, kbuf, o));
        gf[o] = f;
      }
      g.sync();
      long long base = 0;
      int zero_mask = 0;
      bool closing = false;
      if (out.mode == 1) {
        closing = is_closing(ks, mibuf);
        for (int d = 0; d < out.nd; ++d) {
          base += mibuf[d] * out.gstride[d];
          if (mibuf[d] == 0 && out.wrap[d]) zero_mask |= 1 << d;
        }
      }
      backtransform_all<MAXM, (MAXM <= 4 ? 4 : (MAXM <= 8 ? 2 : 1))>(g, w, w.V, 2 * w.nb, [&](int c, int o, cplx x) {
        cplx v = x * gf[o];
        if (out.mode == 0) {
          out.evec[c * out.vc_sb + idx * out.vc_sk + o] = v;
        } else {
          if (closing) v = v * out.pbc_phase[o];
          const long long at = (long long)c * out.sstride + o;
          out.evec[base + at] = v;
          if (zero_mask) {
            for (int mm = 1; mm < (1 << out.nd); ++mm) {
              if ((mm & zero_mask) != mm) continue;
              long long off = base;
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
      });
      g.sync();
      TBK_PROF_MARK(3)
    }
    if (prof && tid == 0) { atomicAdd(prof + 4, 1ull); atomicMax(prof + 6, (unsigned long long)(clock64() - tstart)); }
#undef TBK_PROF_MARK
    if (out.mode == 0) {
      if (out.eval)
        for (int b = tid; b < n; b += T) out.eval[b * out.ev_sb + idx * out.ev_sk] = w.lam[b];
    } else if (out.gaps_bits != nullptr) {
      for (int b = tid; b < n - 1; b += T) atomic_min_nonneg(out.gaps_bits + b, w.lam[b + 1] - w.lam[b]);
    }
    g.sync();
  }
}

// ===========================================================================
// Full H(k) output (tb_model._gen_ham)
// ===========================================================================
__global__ void __launch_bounds__(128)
gen_ham_kernel(PlanView pv, const double* __restrict__ klist, long long nk, cplx* __restrict__ ham) {
  extern __shared__ __align__(16) char smem[];
  cplx* ph = (cplx*)smem;                       // [nph]
  cplx* dj = ph + (pv.nph > 0 ? pv.nph : 1);    // [n]
  const int n = pv.nsta;
  for (long long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
    double k[TBK_MAX_DIM] = {0.0, 0.0, 0.0, 0.0};
    for (int d = 0; d < pv.dim_k; ++d) k[d] = klist[ik * pv.dim_k + d];
    for (int p = threadIdx.x; p < pv.nph; p += blockDim.x) {
      double x = 0.0;
      for (int d = 0; d < pv.dim_k; ++d) x = fma(k[d], pv.ph_R[p * pv.dim_k + d], x);
      ph[p] = expi_turns(x);
    }
    for (int o = threadIdx.x; o < n; o += blockDim.x)
      dj[o] = (pv.convention == 1 && pv.dim_k > 0) ? plan_gauge(pv, k, o) : mk(1.0, 0.0);
    cplx* h = ham + ik * (long long)n * n;
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) h[q] = mk(0.0, 0.0);
    __syncthreads();
    for (int e = threadIdx.x; e < pv.nel; e += blockDim.x) {
      const int r = pv.el_row[e], c = pv.el_col[e];
      cplx v = plan_element(pv, e, ph, 1);
      v = cmul(dj[r], v) * dj[c];               // H_I = D^H H_II D
      if (r == c) v.im = 0.0;
      h[(size_t)r * n + c] = v;
      if (r != c) h[(size_t)c * n + r] = conj(v);
    }
    __syncthreads();
  }
}

// ===========================================================================
// host-side launch logic
// ===========================================================================
static size_t small_plan_smem(const PlanView& pv, int n) {
  return (size_t)pv.nterm * 16 + (size_t)pv.nph * pv.dim_k * 8 + (size_t)n * pv.dim_k * 8 +
         (size_t)(pv.nph + 2) * 4 + (size_t)pv.nterm * 4 + 64;
}

static int block_threads_for(int n) { return n <= 64 ? 64 : (n <= 128 ? 128 : 256); }

static int launch_solve(const PlanView& pv, const KSrc& ks, const cplx* hsrc, int n, long long npts,
                        const OutSpec& out, int want_vec, void* ws, size_t ws_bytes, cudaStream_t st,
                        const double* gemm_tab = nullptr, int gemm_ks = 0) {
  if (npts <= 0) return TBK_OK;
  if (n <= 4 && n >= 2) {
    const size_t sm = hsrc ? 0 : small_plan_smem(pv, n);
    const int stage = (!hsrc && sm <= 40 * 1024) ? 1 : 0;
    const size_t dyn = stage ? sm : 0;
    const long long blocks = (npts + 127) / 128;
    if (blocks > 0x7fffffffLL) { set_error("too many k-points for one launch"); return TBK_ERR_ARG; }
    if (n == 2) solve_small_kernel<2><<<(unsigned)blocks, 128, dyn, st>>>(pv, ks, hsrc, npts, out, want_vec, stage);
    else if (n == 3) solve_small_kernel<3><<<(unsigned)blocks, 128, dyn, st>>>(pv, ks, hsrc, npts, out, want_vec, stage);
    else solve_small_kernel<4><<<(unsigned)blocks, 128, dyn, st>>>(pv, ks, hsrc, npts, out, want_vec, stage);
    TBK_LAUNCH_CHECK("solve_small_kernel");
    note_kernel("solve_small_kernel");
    return TBK_OK;
  }
  const int nph = hsrc ? 0 : pv.nph;
  {
    // eigenvalues only, n = 5..8: one k-point per thread, eigensolver in registers (TBK_REG_EIGVALS=0: the tile solver)
    static int reg_on = -1;
    if (reg_on < 0) { const char* e = getenv("TBK_REG_EIGVALS"); reg_on = (e && atoi(e) == 0) ? 0 : 1; }
    static int gemm_on = -1;                           // TBK_REG_GEMM=0: scalar assembly (A/B knob)
    if (gemm_on < 0) { const char* e = getenv("TBK_REG_GEMM"); gemm_on = (e && atoi(e) == 0) ? 0 : 1; }
    const bool reg_vals = !want_vec && out.mode == 0 && out.eval != nullptr;
    const bool reg_vecs = want_vec && out.evec != nullptr && (out.mode == 1 || out.mode == 0);
    if (reg_on && gemm_on && gemm_tab != nullptr && hsrc == nullptr && n >= 5 && n <= 8 && (reg_vals || reg_vecs) && npts > 0) {
      const int MT = (n * (n + 1) + 7) / 8;
      int rows = 4 * gemm_ks > 8 * MT ? 4 * gemm_ks : 8 * MT;
      if (want_vec && rows < n * n + 4 * n) rows = n * n + 4 * n;      // rotation matrix, gauge factors, one staged band (eigenvector variant)
      const size_t dyn = (size_t)rows * kRegLDB * 8;
      if (dyn + 1024 <= (size_t)kMaxSmem) {
        int per_sm = (int)((size_t)kMaxSmem / (dyn + 1024));
        if (per_sm > 2) per_sm = 2;
        long long blocks = (npts + kRegThreads - 1) / kRegThreads;
        if (blocks > (long long)kNumSM * per_sm) blocks = (long long)kNumSM * per_sm;
#define TBK_REGG_LAUNCH(NN)                                                                                             \
        do {                                                                                                            \
          if (want_vec) {                                                                                               \
            TBK_CUDA(cudaFuncSetAttribute(solve_reg_gemm_kernel<NN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
            solve_reg_gemm_kernel<NN, true><<<(unsigned)blocks, kRegThreads, dyn, st>>>(pv, ks, npts, out, gemm_tab, gemm_ks);  \
          } else {                                                                                                      \
            TBK_CUDA(cudaFuncSetAttribute(solve_reg_gemm_kernel<NN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
            solve_reg_gemm_kernel<NN, false><<<(unsigned)blocks, kRegThreads, dyn, st>>>(pv, ks, npts, out, gemm_tab, gemm_ks);  \
          }                                                                                                             \
        } while (0)
        switch (n) {
          case 5: TBK_REGG_LAUNCH(5); break;
          case 6: TBK_REGG_LAUNCH(6); break;
          case 7: TBK_REGG_LAUNCH(7); break;
          default: TBK_REGG_LAUNCH(8); break;
        }
#undef TBK_REGG_LAUNCH
        TBK_LAUNCH_CHECK("solve_reg_gemm_kernel");
        note_kernel("solve_reg_gemm_kernel");
        return TBK_OK;
      }
    }
    if (reg_on && n >= 5 && n <= 8 && !want_vec && out.mode == 0 && out.eval != nullptr &&
        (size_t)nph * kRegThreads * 16 + 1024 <= (size_t)kMaxSmem) {
      const size_t dyn = (size_t)nph * kRegThreads * 16 + (size_t)(n * (n + 1) / 2) * 4 + 16;    // phase table + element map
      int per_sm = (int)((size_t)kMaxSmem / (dyn + 1024));
      if (per_sm > 4) per_sm = 4;
      if (per_sm < 1) per_sm = 1;
      long long blocks = (npts + kRegThreads - 1) / kRegThreads;
      if (blocks > (long long)kNumSM * per_sm) blocks = (long long)kNumSM * per_sm;
#define TBK_REG_LAUNCH(NN)                                                                                          \
      do {                                                                                                          \
        TBK_CUDA(cudaFuncSetAttribute(solve_reg_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
        solve_reg_kernel<NN><<<(unsigned)blocks, kRegThreads, dyn, st>>>(pv, ks, hsrc, npts, out);                  \
      } while (0)
      switch (n) {
        case 5: TBK_REG_LAUNCH(5); break;
        case 6: TBK_REG_LAUNCH(6); break;
        case 7: TBK_REG_LAUNCH(7); break;
        default: TBK_REG_LAUNCH(8); break;
      }
#undef TBK_REG_LAUNCH
      TBK_LAUNCH_CHECK("solve_reg_kernel");
      note_kernel("solve_reg_kernel");
      return TBK_OK;
    }
  }
  if (n <= 32) {
    const int G = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    const int mats = 128 / G;
    GroupShape gs = group_shape(n, nph, true);
    const size_t dyn = gs.region * mats;
    if (dyn > (size_t)kMaxSmem) { set_error("model needs %zu bytes of shared memory per CTA", dyn); return TBK_ERR_UNSUPPORTED; }
    int per_sm = (int)((size_t)kMaxSmem / (dyn + 1024));
    if (per_sm > 12) per_sm = 12;
    if (per_sm < 1) per_sm = 1;
    long long blocks = (npts + mats - 1) / mats;
    if (blocks > (long long)kNumSM * per_sm) blocks = (long long)kNumSM * per_sm;
#define TBK_TILE_LAUNCH(GG)                                                                                   \
    do {                                                                                                      \
      TBK_CUDA(cudaFuncSetAttribute(solve_tile_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
      solve_tile_kernel<GG><<<(unsigned)blocks, 128, dyn, st>>>(pv, ks, hsrc, npts, out, want_vec, gs);       \
    } while (0)
    if (G == 8) TBK_TILE_LAUNCH(8);
    else if (G == 16) TBK_TILE_LAUNCH(16);
    else TBK_TILE_LAUNCH(32);
#undef TBK_TILE_LAUNCH
    TBK_LAUNCH_CHECK("solve_tile_kernel");
    note_kernel("solve_tile_kernel");
    return TBK_OK;
  }
  // blocked solver (the default for nsta > 32; TBK_BLOCKED=0 keeps the unblocked one-CTA solver)
  {
    const char* env = getenv("TBK_BLOCKED");
    if (n <= kBlkMaxN && !(env && atoi(env) == 0)) {
      const BlkShape shp = blk_shape(n, nph);
      if (shp.smem <= (size_t)kMaxSmem) {
        WyArgs stg;
        memset(&stg, 0, sizeof(stg));
        unsigned* work = nullptr;
#define TBK_BLK_LAUNCH(MM, BLOCKS, COUNT, GWS, IDX0)                                                              \
        do {                                                                                                      \
          TBK_CUDA(cudaFuncSetAttribute(solve_blocked_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shp.smem)); \
          solve_blocked_kernel<MM><<<(unsigned)(BLOCKS), shp.threads, shp.smem, st>>>(pv, ks, hsrc, COUNT, out, want_vec, shp, GWS, blk_prof(), stg, IDX0, work); \
        } while (0)
#define TBK_BLK_DISPATCH(BLOCKS, COUNT, GWS, IDX0)                                                                \
        do {                                                                                                      \
          if (shp.threads == 512) TBK_BLK_LAUNCH(16, BLOCKS, COUNT, GWS, IDX0);   /* (compiled for 512 threads) */ \
          else if (n <= 128) TBK_BLK_LAUNCH(4, BLOCKS, COUNT, GWS, IDX0);                                         \
          else TBK_BLK_LAUNCH(8, BLOCKS, COUNT, GWS, IDX0);                                                       \
        } while (0)
        if (want_vec && blk_staged_enabled()) {
          BlkStage sg = blk_stage_layout(shp, npts);
          if (ws == nullptr || ws_bytes < sg.total) { set_error("workspace too small: need %zu bytes, have %zu", sg.total, ws_bytes); return TBK_ERR_WORKSPACE; }
          stg = sg.wy;
          stg.ws = (char*)ws;
          char* lu_base = (char*)ws + (size_t)sg.slots * sg.wy.slot_bytes;
          const size_t smem_t = wy_t_smem(), smem_b = wy_bt_smem(n);
          TBK_CUDA(cudaFuncSetAttribute(blk_wy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
          TBK_CUDA(cudaFuncSetAttribute(blk_backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
          for (long long base = 0; base < npts; base += sg.slots) {
            const long long cnt = npts - base < sg.slots ? npts - base : sg.slots;
            const long long blocks = cnt < sg.ctas ? cnt : sg.ctas;
            work = cnt > blocks ? take_ticket() : nullptr;       // more matrices than CTAs: dynamic distribution
            TBK_BLK_DISPATCH(blocks, cnt, lu_base, base);
            TBK_LAUNCH_CHECK("solve_blocked_kernel");
            blk_wy_kernel<<<dim3((unsigned)stg.nblk, (unsigned)cnt), kWyThreads, smem_t, st>>>(stg);
            TBK_LAUNCH_CHECK("blk_wy_kernel");
            blk_backtransform_kernel<<<dim3((unsigned)((n + kWyNC - 1) / kWyNC), (unsigned)cnt), kWyThreads, smem_b, st>>>(
                stg, pv, ks, out, hsrc != nullptr ? 1 : 0, base);
            TBK_LAUNCH_CHECK("blk_backtransform_kernel");
          }
          note_kernel("solve_blocked_kernel");
          return TBK_OK;
        }
        const long long blocks = blk_blocks(shp, npts);
        const size_t need = (size_t)blocks * shp.ws_block;
        if (ws == nullptr || ws_bytes < need) { set_error("workspace too small: need %zu bytes, have %zu", need, ws_bytes); return TBK_ERR_WORKSPACE; }
        TBK_BLK_DISPATCH(blocks, npts, (char*)ws, 0LL);
#undef TBK_BLK_DISPATCH
#undef TBK_BLK_LAUNCH
        TBK_LAUNCH_CHECK("solve_blocked_kernel");
        note_kernel("solve_blocked_kernel");
        return TBK_OK;
      }
    }
  }
  // one matrix per CTA
  GroupShape gs = group_shape(n, nph, true);
  bool in_smem = gs.region <= (size_t)kMaxSmem;
  if (!in_smem) gs = group_shape(n, nph, false);
  if (gs.region > (size_t)kMaxSmem) { set_error("nsta=%d with %d phases does not fit shared memory", n, nph); return TBK_ERR_UNSUPPORTED; }
  const int threads = block_threads_for(n);
  int per_sm = in_smem ? (int)((size_t)kMaxSmem / (gs.region + 1024)) : 2;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long blocks = npts < (long long)kNumSM * per_sm ? npts : (long long)kNumSM * per_sm;
  cplx* gA = nullptr;
  if (!in_smem) {
    const size_t need = (size_t)blocks * n * gs.lda * 16;
    if (ws == nullptr || ws_bytes < need) { set_error("workspace too small: need %zu bytes, have %zu", need, ws_bytes); return TBK_ERR_WORKSPACE; }
    gA = (cplx*)ws;
  }
  TBK_CUDA(cudaFuncSetAttribute(solve_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs.region));
  solve_block_kernel<<<(unsigned)blocks, threads, gs.region, st>>>(pv, ks, hsrc, npts, out, want_vec, gs, gA);
  TBK_LAUNCH_CHECK("solve_block_kernel");
  note_kernel("solve_block_kernel");
  return TBK_OK;
}

constexpr int kMeshGridCap = kNumSM * 8;
static size_t solve_ws_bytes(int n, long long npts) {
  if (n <= 4) return (size_t)kMeshGridCap * 4 * 8;   // per-CTA gap partials of mesh_small_kernel
  if (n <= 32) return 0;   // register / tile kernels: shared-memory resident
  long long blocks = npts < (long long)kNumSM * 2 ? npts : (long long)kNumSM * 2;
  if (blocks < 1) blocks = 1;
  size_t need = n <= 96 ? 0 : (size_t)blocks * n * (n | 1) * 16;        // unblocked solver with A in the workspace
  if (n <= kBlkMaxN) {                                                   // blocked solver (nph does not change the workspace)
    const BlkShape shp = blk_shape(n, 0);
    long long b2 = blk_blocks(shp, npts < 1 ? 1 : npts);
    // the shape used at launch may have fewer resident CTAs (phase table in shared memory), never more
    const size_t n2 = (size_t)b2 * shp.ws_block;
    if (n2 > need) need = n2;
    const size_t n3 = blk_stage_layout(shp, npts < 1 ? 1 : npts).total;   // staged solver (eigenvectors wanted)
    if (n3 > need) need = n3;
  }
  return need;
}

// n == 1 is trivial but must work (single-orbital models)
__global__ void solve_n1_kernel(PlanView pv, KSrc ks, const cplx* __restrict__ hsrc, long long npts, OutSpec out, int want_vec) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= npts) return;
  int mi[TBK_MAX_DIM] = {0, 0, 0, 0};
  double k[TBK_MAX_DIM];
  double e0;
  if (hsrc) e0 = hsrc[idx].re;
  else {
    if (out.mode == 1) decode_index(idx, out, mi);
    load_k(ks, idx, mi, k);
    cplx acc = mk(0.0, 0.0);
    for (int e = 0; e < pv.nel; ++e) {
      for (int t = pv.el_ptr[e]; t < pv.el_ptr[e + 1]; ++t) {
        const cplx a = mk(pv.t_amp[2 * t], pv.t_amp[2 * t + 1]);
        const int p = pv.t_ph[t];
        if (p < 0) acc = acc + a;
        else {
          double x = 0.0;
          for (int d = 0; d < pv.dim_k; ++d) x = fma(k[d], pv.ph_R[(p & TBK_PH_MASK) * pv.dim_k + d], x);
          cplx z = expi_turns(x);
          if (p & TBK_PH_CONJ) z.im = -z.im;
          fma_acc(acc, a, z);
        }
      }
    }
    e0 = acc.re;
  }
  if (out.mode == 0) {
    if (out.eval) out.eval[idx * out.ev_sk] = e0;
    if (want_vec) out.evec[idx * out.vc_sk] = mk(1.0, 0.0);
  } else {
    long long base = 0;
    int zero_mask = 0;
    for (int d = 0; d < out.nd; ++d) {
      base += mi[d] * out.gstride[d];
      if (mi[d] == 0 && out.wrap[d]) zero_mask |= 1 << d;
    }
    const cplx v0 = is_closing(ks, mi) ? out.pbc_phase[0] : mk(1.0, 0.0);
    out.evec[base] = v0;
    for (int m = 1; m < (1 << out.nd); ++m) {
      if ((m & zero_mask) != m) continue;
      long long off = base;
      cplx f = v0;
      for (int d = 0; d < out.nd; ++d)
        if (m & (1 << d)) { off += (long long)(out.full[d] - 1) * out.gstride[d]; f = f * out.pbc_phase[d]; }
      out.evec[off] = f;
    }
  }
}

}  // namespace tbk

using namespace tbk;

extern "C" {

int tbk_gen_ham(const tbk_model* m, const double* k_dev, int64_t nk, double* ham_dev, void* stream) {
  TBK_NVTX("tbk_gen_ham");
  if (!m || !ham_dev || nk < 0 || (m->pv.dim_k > 0 && !k_dev)) { set_error("tbk_gen_ham: bad argument"); return TBK_ERR_ARG; }
  if (nk == 0) return TBK_OK;
  const size_t dyn = (size_t)((m->pv.nph > 0 ? m->pv.nph : 1) + m->pv.nsta) * 16;
  if (dyn > (size_t)kMaxSmem) { set_error("tbk_gen_ham: phase table too large"); return TBK_ERR_UNSUPPORTED; }
  TBK_CUDA(cudaFuncSetAttribute(gen_ham_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  const long long blocks = nk < (long long)kNumSM * 8 ? nk : (long long)kNumSM * 8;
  gen_ham_kernel<<<(unsigned)blocks, 128, dyn, (cudaStream_t)stream>>>(m->pv, k_dev, nk, (cplx*)ham_dev);
  TBK_LAUNCH_CHECK("gen_ham_kernel");
  return TBK_OK;
}

int tbk_kmesh_uniform(const int32_t* mesh, int32_t nd, double* k_dev, void* stream) {
  TBK_NVTX("tbk_kmesh_uniform");
  if (!mesh || !k_dev || nd < 1 || nd > TBK_MAX_DIM) { set_error("tbk_kmesh_uniform: bad argument"); return TBK_ERR_ARG; }
  KMeshDesc md;
  long long nk = 1;
  for (int d = 0; d < TBK_MAX_DIM; ++d) {
    md.n[d] = d < nd ? mesh[d] : 1;
    if (md.n[d] < 1) { set_error("tbk_kmesh_uniform: mesh sizes must be positive"); return TBK_ERR_ARG; }
    nk *= md.n[d];
  }
  md.nd = nd;
  kmesh_uniform_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, (cudaStream_t)stream>>>(md, nk, k_dev);
  TBK_LAUNCH_CHECK("kmesh_uniform_kernel");
  return TBK_OK;
}

int tbk_debug_profile(uint64_t* out8, int32_t reset) {
  if (!out8) { set_error("tbk_debug_profile: null argument"); return TBK_ERR_ARG; }
  for (int i = 0; i < 8; ++i) out8[i] = g_blk_prof ? g_blk_prof[i] : 0;
  if (reset && g_blk_prof) memset(g_blk_prof, 0, 8 * sizeof(unsigned long long));
  return TBK_OK;
}

size_t tbk_eigh_workspace(int32_t n, int64_t batch, int32_t) { return solve_ws_bytes(n, batch); }

int tbk_eigh_batched(const double* ham_dev, int32_t n, int64_t batch, double* eval_dev, double* evec_dev,
                     void* ws_dev, size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_eigh_batched");
  if (!ham_dev || !eval_dev || n < 1 || batch < 0) { set_error("tbk_eigh_batched: bad argument"); return TBK_ERR_ARG; }
  PlanView pv;
  memset(&pv, 0, sizeof(pv));
  pv.nsta = n; pv.convention = 2;
  KSrc ks;
  memset(&ks, 0, sizeof(ks));
  ks.closing_g = -1;
  OutSpec out;
  memset(&out, 0, sizeof(out));
  out.mode = 0;
  out.eval = eval_dev; out.ev_sb = 1; out.ev_sk = n;
  out.evec = (cplx*)evec_dev; out.vc_sb = n; out.vc_sk = (long long)n * n;
  const int want_vec = evec_dev != nullptr;
  if (n == 1) {
    solve_n1_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pv, ks, (const cplx*)ham_dev, batch, out, want_vec);
    TBK_LAUNCH_CHECK("solve_n1_kernel");
    return TBK_OK;
  }
  return launch_solve(pv, ks, (const cplx*)ham_dev, n, batch, out, want_vec, ws_dev, ws_bytes, (cudaStream_t)stream);
}

size_t tbk_solve_workspace(int32_t nsta, int64_t nk, int32_t) { return solve_ws_bytes(nsta, nk); }

int tbk_solve_k(const tbk_model* m, const double* k_dev, int64_t nk, double* eval_dev, int64_t ev_sb, int64_t ev_sk,
                double* evec_dev, int64_t vc_sb, int64_t vc_sk, void* ws_dev, size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_solve_k");
  if (!m || nk < 0 || (!eval_dev && !evec_dev) || (m->pv.dim_k > 0 && !k_dev)) { set_error("tbk_solve_k: bad argument"); return TBK_ERR_ARG; }
  KSrc ks;
  memset(&ks, 0, sizeof(ks));
  ks.klist = k_dev; ks.dim_k = m->pv.dim_k; ks.closing_g = -1;
  OutSpec out;
  memset(&out, 0, sizeof(out));
  out.mode = 0;
  out.eval = eval_dev; out.ev_sb = ev_sb; out.ev_sk = ev_sk;
  out.evec = (cplx*)evec_dev; out.vc_sb = vc_sb; out.vc_sk = vc_sk;
  const int want_vec = evec_dev != nullptr;
  if (m->pv.nsta == 1) {
    solve_n1_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m->pv, ks, nullptr, nk, out, want_vec);
    TBK_LAUNCH_CHECK("solve_n1_kernel");
    return TBK_OK;
  }
  return launch_solve(m->pv, ks, nullptr, m->pv.nsta, nk, out, want_vec, ws_dev, ws_bytes, (cudaStream_t)stream,
                      m->gemm_tab, m->gemm_ks);
}

// mesh_small_kernel dispatch; returns 1 if it took the job, 0 if the shape does not fit, < 0 on error
static bool mesh_small_takes(const tbk_model* m, const OutSpec& out) {
  const int n = m->pv.nsta, nd = out.nd;
  if (!m->dense.valid || n < 2 || n > 4) return false;
  if (((uintptr_t)out.evec & 31) != 0 || ((out.sstride * 16) & 31) != 0) return false;
  if (nd > 1 && out.cnt[nd - 1] < 48) return false;   // too few points along the fastest axis to fill a CTA row
  long long nseg = (out.cnt[nd - 1] + kMeshThreads - 1) / kMeshThreads;
  for (int d = 0; d < nd - 1; ++d) nseg *= out.cnt[d];
  return nseg > 0 && nseg < 0x7fffffffLL;
}

static int launch_mesh_small(const tbk_model* m, const KSrc& ks, const OutSpec& out, double* gaps_dev,
                             void* ws, size_t ws_bytes, tbk_peer* peer, cudaStream_t st) {
  const DenseSmall& ds = m->dense;
  const int n = m->pv.nsta, nd = out.nd;
  if (!ds.valid || n < 2 || n > 4) return 0;
  if (nd > 1 && out.cnt[nd - 1] < 48) return 0;       // too few points along the fastest axis to fill a CTA row
  if (((uintptr_t)out.evec & 31) != 0 || ((out.sstride * 16) & 31) != 0) return 0;      // 256-bit stores need a 32-byte aligned array
  MeshTiling tl;
  long long outer = 1;
  for (int d = 0; d < nd - 1; ++d) outer *= out.cnt[d];
  tl.nbx = (out.cnt[nd - 1] + kMeshThreads - 1) / kMeshThreads;
  const long long nseg = outer * tl.nbx;
  if (nseg <= 0) return 1;
  if (nseg >= 0x7fffffffLL) return 0;                 // 32-bit tile arithmetic in the kernel: leave huge meshes to the generic path
  tl.outer = (int)outer;
  tl.nseg = (unsigned)nseg;
  tl.closing_g = ks.closing_g;
  const bool p4 = ds.nph <= 4;
  // one balanced wave: #SM x (CTAs resident per SM) persistent CTAs, each with an equal share of rows.
  // resident CTAs per SM: n = 2: 4 (x 2 rows in flight per thread; 3, 5 and 6 CTAs and 1 row measured slower, profiles/README.md r02)
  int occ = n == 2 ? 4 : (n == 3 ? 3 : 2);
  if (n == 4 && p4) { const char* e = getenv("TBK_MESH_VARIANT4"); const int v4 = e ? atoi(e) : 0; occ = v4 == 1 ? 2 : (v4 == 2 ? 4 : 3); }
  long long want = (long long)kNumSM * occ;
  // at least ~4 rows per CTA so the per-CTA sincospi prologue stays amortised
  if (want > (nseg + 3) / 4) want = (nseg + 3) / 4;
  if (want < 1) want = 1;
  // few column blocks (2-D meshes): split the CTAs evenly over the blocks so that none straddles two of them
  tl.cpb = 0;
  static int flat = -1;                                  // TBK_MESH_FLAT=1: the flat split (A/B knob)
  if (flat < 0) { const char* e = getenv("TBK_MESH_FLAT"); flat = (e && atoi(e) == 1) ? 1 : 0; }
  if (!flat && n == 2 && tl.nbx <= want && outer >= 1) {   // measured: n = 2 gains 0.5 us of 20, n = 4 loses 3 % (fewer CTAs)
    long long cpb = want / tl.nbx;
    if (cpb > outer) cpb = outer;
    if (cpb >= 1 && cpb * tl.nbx * 10 >= want * 9) { tl.cpb = (int)cpb; want = cpb * tl.nbx; }   // keep >= 90 % of the wave
  }
  const int grid = (int)want;
  double* partial = nullptr;
  unsigned* ticket = nullptr;
  if (gaps_dev) {
    const size_t need = (size_t)grid * (n - 1) * 8;
    if (!ws || ws_bytes < need) { set_error("tbk_solve_grid: workspace too small (%zu < %zu)", ws_bytes, need); return TBK_ERR_WORKSPACE; }
    partial = (double*)ws;
    ticket = take_ticket();
    if (!ticket) { set_error("tbk_solve_grid: cannot allocate the reduction tickets"); return TBK_ERR_CUDA; }
  }
  const int gauge = (m->pv.convention == 1 && m->pv.dim_k > 0) ? 1 : 0;
  // cross-rank minimum of the gaps: synchronous, or (tbk_peer_defer) only posted by this kernel — the grid solve
  // never completes older collectives itself, so that nothing but a few stores sits between its last CTA and
  // the dependent flux kernel
  PeerView pview = peer_none();
  if (gaps_dev) { if (int rc = peer_next(peer, n - 1, 1, gaps_dev, false, st, &pview)) return rc; }
  // a synchronous prepared call waits on a pinned word this kernel's last CTA writes after the gaps
  // (not for a deferred reduction: the gaps are then completed by a later kernel)
  const DoneSignal done = (gaps_dev && (pview.nranks <= 1 || pview.mode == 1)) ? take_done_request() : DoneSignal{nullptr, 0};
#define TBK_MESH_LAUNCH(NN, PP, MB, RP) \
  mesh_small_kernel<NN, PP, MB, RP><<<grid, kMeshThreads, 0, st>>>(ds, ks, out, tl, gauge, partial, ticket, gaps_dev, pview, cta_trace_buffer(), done)
  if (n == 2) {
    switch (ds.nph) {                               // exact phase counts for the headline case (Haldane: 3)
      case 1: mesh_small_kernel<2, 1, 4, 2, true><<<grid, kMeshThreads, 0, st>>>(ds, ks, out, tl, gauge, partial, ticket, gaps_dev, pview, cta_trace_buffer(), done); break;
      case 2: mesh_small_kernel<2, 2, 4, 2, true><<<grid, kMeshThreads, 0, st>>>(ds, ks, out, tl, gauge, partial, ticket, gaps_dev, pview, cta_trace_buffer(), done); break;
      case 3: mesh_small_kernel<2, 3, 4, 2, true><<<grid, kMeshThreads, 0, st>>>(ds, ks, out, tl, gauge, partial, ticket, gaps_dev, pview, cta_trace_buffer(), done); break;
      case 4: mesh_small_kernel<2, 4, 4, 2, true><<<grid, kMeshThreads, 0, st>>>(ds, ks, out, tl, gauge, partial, ticket, gaps_dev, pview, cta_trace_buffer(), done); break;
      default: TBK_MESH_LAUNCH(2, 8, 4, 1); break;
    }
  }
  else if (n == 3) { if (p4) TBK_MESH_LAUNCH(3, 4, 3, 1); else TBK_MESH_LAUNCH(3, 8, 3, 1); }
  else {
    int v4 = 0;
    { const char* e = getenv("TBK_MESH_VARIANT4"); if (e) v4 = atoi(e); }   // tuning knob: resident CTAs per SM
    if (p4) { if (v4 == 1) TBK_MESH_LAUNCH(4, 4, 2, 1); else if (v4 == 2) TBK_MESH_LAUNCH(4, 4, 4, 1); else TBK_MESH_LAUNCH(4, 4, 3, 1); }
    else TBK_MESH_LAUNCH(4, 8, 2, 1);
  }
#undef TBK_MESH_LAUNCH
  TBK_LAUNCH_CHECK("mesh_small_kernel");
  note_kernel("mesh_small_kernel");
  return 1;
}

int tbk_solve_grid(const tbk_model* m, const double* start_k, const int32_t* mesh, int32_t nd, int32_t row0,
                   int32_t nrows, int32_t wrap0, double* wfs_dev, const double* pbc_phase_dev, double* gaps_dev,
                   void* ws_dev, size_t ws_bytes, void* stream) {
  return tbk_solve_grid_x(m, start_k, mesh, nd, row0, nrows, wrap0, wfs_dev, pbc_phase_dev, gaps_dev, ws_dev, ws_bytes,
                          0, nullptr, stream);
}

int tbk_solve_grid_x(const tbk_model* m, const double* start_k, const int32_t* mesh, int32_t nd, int32_t row0,
                     int32_t nrows, int32_t wrap0, double* wfs_dev, const double* pbc_phase_dev, double* gaps_dev,
                     void* ws_dev, size_t ws_bytes, int64_t state_stride, tbk_peer* peer, void* stream) {
  TBK_NVTX("tbk_solve_grid_x");
  if (!m || !start_k || !mesh || !wfs_dev || !pbc_phase_dev || nd < 1 || nd > TBK_MAX_DIM || nd != m->pv.dim_k ||
      nrows < 0 || row0 < 0 || wrap0 < 0 || wrap0 > 2 || state_stride < 0) {
    set_error("tbk_solve_grid: bad argument (nd=%d dim_k=%d)", nd, m ? m->pv.dim_k : -1);
    return TBK_ERR_ARG;
  }
  const int n = m->pv.nsta;
  KSrc ks;
  memset(&ks, 0, sizeof(ks));
  ks.klist = nullptr; ks.dim_k = nd; ks.row0 = row0;
  ks.closing_g = wrap0 == 2 ? mesh[0] - 1 : -1;
  OutSpec out;
  memset(&out, 0, sizeof(out));
  out.mode = 1; out.nd = nd;
  out.evec = (cplx*)wfs_dev;
  out.pbc_phase = (const cplx*)pbc_phase_dev;
  long long npts = 1;
  for (int d = 0; d < nd; ++d) {
    if (mesh[d] < 2) { set_error("tbk_solve_grid: mesh extent must be >= 2"); return TBK_ERR_ARG; }
    ks.start[d] = start_k[d];
    ks.den[d] = (double)(mesh[d] - 1);
    out.cnt[d] = d == 0 ? nrows + (wrap0 == 2 ? 1 : 0) : mesh[d] - 1;
    out.full[d] = d == 0 ? nrows + 1 : mesh[d];
    out.wrap[d] = d == 0 ? (wrap0 == 1) : 1;
    npts *= out.cnt[d];
  }
  // k-major [row][i_1]..[state][orb]: a mesh point is n*n contiguous elements; state-major [state][row][i_1]..[orb]:
  // a mesh point is n elements per state, the states state_stride apart
  long long stride = state_stride > 0 ? (long long)n : (long long)n * n;
  for (int d = nd - 1; d >= 0; --d) { out.gstride[d] = stride; stride *= out.full[d]; }
  out.sstride = state_stride > 0 ? (long long)state_stride : (long long)n;
  if (state_stride > 0 && state_stride < stride) { set_error("tbk_solve_grid: state_stride smaller than one state plane"); return TBK_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 1) {
    if (npts > 0) solve_n1_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(m->pv, ks, nullptr, npts, out, 1);
    TBK_LAUNCH_CHECK("solve_n1_kernel");
    note_kernel("solve_n1_kernel");
    return TBK_OK;
  }
  const bool fused = peer && peer->connected && peer->nranks > 1 && gaps_dev;
  if (fused && !(npts > 0 && mesh_small_takes(m, out))) {
    set_error("tbk_solve_grid_x: the fused cross-rank reduction needs the register-resident mesh kernel (nsta <= 4)");
    return TBK_ERR_UNSUPPORTED;
  }
  if (npts > 0) {
    const int took = launch_mesh_small(m, ks, out, gaps_dev, ws_dev, ws_bytes, peer, st);
    if (took < 0) return took;
    if (took == 1) return TBK_OK;
  }
  if (gaps_dev) {
    out.gaps_bits = (unsigned long long*)gaps_dev;
    fill_u64_kernel<<<(n - 1 + 127) / 128, 128, 0, st>>>(out.gaps_bits, n - 1, 0x7FF0000000000000ULL);
    TBK_LAUNCH_CHECK("fill_u64_kernel");
  }
  return launch_solve(m->pv, ks, nullptr, n, npts, out, 1, ws_dev, ws_bytes, st, m->gemm_tab, m->gemm_ks);
}

}  // extern "C"
