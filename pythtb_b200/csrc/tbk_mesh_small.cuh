// tbk_mesh_small.cuh — wf_array.solve_on_grid for 2 <= nsta <= 4 on a regular
// mesh (pythtb.py:2421-2532): the headline kernel of BASELINE configs[1]
// (Haldane / Kane-Mele, 1024 x 1024).  Included by tbk_solve.cu.
//
// Design (B200): the kernel is bound by the 16 n^2 bytes it must write per
// k-point, so everything else is arranged to stay below that:
//   * one k-point per thread per step, threads along the FASTEST mesh axis, a CTA
//     walks `ti` consecutive indices of the remaining (flattened) axes;
//   * exp(2 pi i k.R) = exp(2 pi i k_outer.R_outer) * exp(2 pi i k_last R_last): the
//     second factor is computed once per thread per tile, the first once per CTA
//     per mesh row (shared memory) — about 0.4 sincospi per k-point instead of
//     nph + nsta;
//   * H(k) from a dense coefficient table held in the kernel-parameter constant
//     bank (DenseSmall): 4 FMA per (element, phase), no indexed accumulators;
//   * closed-form (n = 2) / Householder + QL (n = 3) / Householder + direct quartic solver (n = 4) in registers
//     (tbk_eig_small.cuh);
//   * Convention-I gauge, periodic images and the running minimum of the direct
//     gaps fused in; the gap reduction finishes in the last CTA (ticket), so a
//     grid solve is ONE launch.
#pragma once

namespace tbk {

constexpr int kMeshThreads = 128;
constexpr int kMeshMaxRows = 32;

struct MeshTiling {
  int nbx;              // 128-wide column blocks along the last axis
  int outer;            // product of cnt[d], d < nd-1  ("rows")
  unsigned nseg;        // outer * nbx row segments (< 2^31, checked by the launcher), column block major
  int closing_g;        // global axis-0 index that is the periodic image of row 0 (wrap0 == 2), else -1
  int cpb;              // > 0: CTAs per column block, each with an equal run of rows of ONE block (grid = cpb * nbx);
};                      // 0: flat split of the nseg segments over the grid (a CTA may straddle two blocks)

template <int N>
struct EigRows {
  cplx w[N][N];
};

// periodic images of one mesh point (pythtb.py:2729-2747): every subset of the axes on which the
// point sits at index 0; one multiply per wrapped axis, in axis order (so that a shard's closing row
// and the unsharded image agree bit for bit).  Cold path: kept out of line.
template <int N>
__device__ __noinline__ void mesh_store_images(const EigRows<N>& e, const OutSpec& out, long long base, int zero_mask) {
  const int nd = out.nd;
  for (int m = 1; m < (1 << nd); ++m) {
    if ((m & zero_mask) != m) continue;
    long long off = base;
    cplx im[N][N];
#pragma unroll
    for (int b = 0; b < N; ++b)
#pragma unroll
      for (int o = 0; o < N; ++o) im[b][o] = e.w[b][o];
    for (int d = 0; d < nd; ++d) {
      if (m & (1 << d)) {
        off += (long long)(out.full[d] - 1) * out.gstride[d];
#pragma unroll
        for (int o = 0; o < N; ++o) {
          const cplx ph = out.pbc_phase[d * N + o];
#pragma unroll
          for (int b = 0; b < N; ++b) im[b][o] = mul_fixed(im[b][o], ph);
        }
      }
    }
    cplx* dsti = out.evec + off;
#pragma unroll
    for (int b = 0; b < N; ++b)
#pragma unroll
      for (int o = 0; o < N; ++o) dsti[b * out.sstride + o] = im[b][o];
  }
}

// 32-byte aligned 256-bit store (STG.E.ENL2.256 on sm_100): one full sector per instruction.
__device__ __forceinline__ void st256(cplx* p, cplx a, cplx b) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.re), "d"(a.im), "d"(b.re), "d"(b.im) : "memory");
}
// WIDE = false: plain 16-byte stores (cold paths; ptxas 12.9 truncates the v4.f64 asm store to its first
// element when the operands are reloaded from local memory inside an out-of-line function).
// ss: elements between the states of the point (N for the reference layout; a multiple of 2 for WIDE)
template <int N, bool WIDE>
__device__ __forceinline__ void mesh_store_point(cplx* dst, const cplx (&w)[N][N], long long ss) {
  if constexpr (N == 2 && WIDE) {
    st256(dst, w[0][0], w[0][1]);
    st256(dst + ss, w[1][0], w[1][1]);
  } else if constexpr (N == 4 && WIDE) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      st256(dst + ss * b, w[b][0], w[b][1]);
      st256(dst + ss * b + 2, w[b][2], w[b][3]);
    }
  } else {
#pragma unroll
    for (int b = 0; b < N; ++b)
#pragma unroll
      for (int o = 0; o < N; ++o) dst[b * ss + o] = w[b][o];
  }
}

// cold path of the store: closing-row pbc factor, the point itself, its periodic images
template <int N>
__device__ __noinline__ void mesh_store_special(EigRows<N>& e, const OutSpec& out, const cplx* pbc0, long long at,
                                                int zero_mask, int closing) {
  if (closing) {                                    // axis-0 image of global row 0 (pythtb.py:2729); a second
#pragma unroll                                      // multiply, so the values equal the unsharded image bit for bit
    for (int o = 0; o < N; ++o) {
      const cplx ph = pbc0[o];
#pragma unroll
      for (int b = 0; b < N; ++b) e.w[b][o] = mul_fixed(e.w[b][o], ph);
    }
  }
  mesh_store_point<N, false>(out.evec + at, e.w, out.sstride);
  if (zero_mask) mesh_store_images<N>(e, out, at, zero_mask);
}

// periodic image of one point along one axis: every component times that axis' pbc phase, 256-bit stores
template <int N>
__device__ __forceinline__ void mesh_store_image(cplx* dst, const cplx (&w)[N][N], const cplx* ph, long long ss) {
  cplx im[N][N];
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int o = 0; o < N; ++o) im[b][o] = mul_fixed(w[b][o], ph[o]);
  mesh_store_point<N, true>(dst, im, ss);
}

struct MeshRow {                                    // per mesh row (outer index), shared memory
  long long base;                                   // storage offset of the row, complex elements
  int flags;                                        // bits 0..3 zero_mask, bit 8 closing row
  int pad;
};

// Persistent kernel: the grid is one balanced wave (#SM x resident CTAs); CTA c owns the row
// segments [c*nseg/G, (c+1)*nseg/G) — a contiguous run of rows inside one column block (two at a
// block boundary), processed in chunks of <= kMeshMaxRows rows, RPI rows per loop iteration
// (independent dependency chains in one basic block: the kernel is latency-, not bandwidth-bound).
//
// Gauge: the stored Convention-I eigenvector is  d_0(k) * D(k)^H u_II, i.e. component o carries
// exp(-2 pi i k.(tau_o - tau_0)): the overall phase of an eigenvector is arbitrary (LAPACK's is too),
// and this choice needs N-1 instead of N phase factors per k-point.
// EXACT: the model has exactly NPH phases (compile-time trip count: no per-phase test splits the basic block in
// which the RPI independent rows interleave).
template <int N, int NPH, int MINB, int RPI, bool EXACT = false>
__global__ void __launch_bounds__(kMeshThreads, MINB)
mesh_small_kernel(const __grid_constant__ DenseSmall ds, const __grid_constant__ KSrc ks,
                  const __grid_constant__ OutSpec out, const __grid_constant__ MeshTiling tl, int gauge,
                  double* __restrict__ gap_partial, unsigned* __restrict__ ticket, double* __restrict__ gaps_out,
                  const __grid_constant__ PeerView peer, unsigned long long* __restrict__ trace, const DoneSignal done) {
  constexpr int NP = N * (N + 1) / 2;
  constexpr int NG = N - 1;                         // relative gauge factors of states 1..N-1
  constexpr int NQ = NPH + NG;
  __shared__ cplx s_out[kMeshMaxRows][NQ];
  __shared__ MeshRow s_row[kMeshMaxRows];
  __shared__ double s_red[kMeshThreads / 32][N];
  __shared__ cplx s_pbc0[N];                        // pbc phase of axis 0 (closing rows of a shard)
  __shared__ cplx s_pbc[TBK_MAX_DIM][N];            // pbc phases of every wrapped axis (periodic images)
  __shared__ int s_last;
  __shared__ double s_fin[N];
  const unsigned long long t_begin = cta_trace_begin(trace);
  // let a programmatic dependent (the flux kernel) be scheduled as soon as this grid's CTAs retire: the whole
  // grid is resident from the start (one wave), so the dependent can never take a slot a CTA of this grid needs
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  peer_prologue(peer);                              // first CTA: post the deferred reductions of the previous step (N > 1)
  if (threadIdx.x < N) s_pbc0[threadIdx.x] = tl.closing_g >= 0 ? out.pbc_phase[threadIdx.x] : mk(1.0, 0.0);
  if (threadIdx.x < out.nd * N) {
    const int d = threadIdx.x / N;
    s_pbc[d][threadIdx.x - d * N] = out.wrap[d] ? out.pbc_phase[threadIdx.x] : mk(1.0, 0.0);
  }
  const int nd = out.nd;
  const int last = nd - 1;
  const int nph = EXACT ? NPH : ds.nph;   // phases p >= nph are skipped (uniform predicate)
  const int tid = threadIdx.x;
  const long long gs_last = out.gstride[last];
  const long long ss = out.sstride;
  double gmin[N - 1];
#pragma unroll
  for (int b = 0; b < N - 1; ++b) gmin[b] = INFINITY;

  // balanced split (32-bit arithmetic only).  cpb > 0: the CTA owns rows [i outer/cpb, (i+1) outer/cpb) of column
  // block bx — a CTA that straddled two column blocks paid the per-block prologue (sincospi, phase powers, the
  // per-row phase table) twice and was the tail of the wave.  cpb == 0: the first (nseg % G) CTAs take one segment more.
  unsigned seg, seg_end;
  if (tl.cpb > 0) {
    const unsigned bxi = blockIdx.x / (unsigned)tl.cpb, i = blockIdx.x - bxi * (unsigned)tl.cpb;
    const unsigned per = (unsigned)tl.outer / (unsigned)tl.cpb, extra = (unsigned)tl.outer - per * (unsigned)tl.cpb;
    seg = bxi * (unsigned)tl.outer + per * i + (i < extra ? i : extra);
    seg_end = seg + per + (i < extra ? 1u : 0u);
  } else {
    const unsigned per = tl.nseg / gridDim.x, extra = tl.nseg - per * gridDim.x;
    seg = per * blockIdx.x + (blockIdx.x < extra ? blockIdx.x : extra);
    seg_end = seg + per + (blockIdx.x < extra ? 1u : 0u);
  }
  while (seg < seg_end) {
    const int bx = (int)(seg / (unsigned)tl.outer);
    const int row_lo = (int)(seg - (unsigned)bx * (unsigned)tl.outer);
    int row_hi = row_lo + (int)(seg_end - seg);
    if (row_hi > tl.outer) row_hi = tl.outer;
    seg += (unsigned)(row_hi - row_lo);
    // ---- per-thread factors along the fastest axis: R components are integers, so every phase
    // factor is a power of E1 = exp(2 pi i k_last) (one sincospi); the gauge factors need one
    // sincospi per distinct orbital position (both spin components of an orbital share it)
    const int j = bx * kMeshThreads + tid;
    const bool active = j < out.cnt[last];
    const int gj = j + (last == 0 ? ks.row0 : 0);
    const bool closing_j = (last == 0 && gj == tl.closing_g);
    const double kl = ks.start[last] + (double)(closing_j ? 0 : gj) / ks.den[last];
    cplx fc[NQ];
    {
      const cplx e1 = expi_turns(kl);
#pragma unroll
      for (int p = 0; p < NPH; ++p) {
        const int m = (int)ds.R[p][last];            // 0 for the zero-padded phases p >= nph
        const int am = m < 0 ? -m : m;
        cplx z = mk(1.0, 0.0);
        for (int i = 0; i < am; ++i) z = z * e1;
        if (m < 0) z.im = -z.im;
        fc[p] = z;
      }
#pragma unroll
      for (int o = 1; o < N; ++o) {
        const double dt = ds.tau[o][last] - ds.tau[0][last];
        if (dt == 0.0) fc[NPH + o - 1] = mk(1.0, 0.0);
        else if (o > 1 && ds.tau[o][last] == ds.tau[o - 1][last]) fc[NPH + o - 1] = fc[NPH + o - 2];
        else fc[NPH + o - 1] = expi_turns(-kl * dt);
      }
    }
    const int zlast = (j == 0 && out.wrap[last]) ? (1 << last) : 0;
    const int special_j = zlast | (closing_j ? 256 : 0);
    cplx* const dst_col = out.evec + (long long)j * gs_last;

    for (int chunk = row_lo; chunk < row_hi; chunk += kMeshMaxRows) {
      const int nrows = row_hi - chunk < kMeshMaxRows ? row_hi - chunk : kMeshMaxRows;
      // ---- per-row (outer index) factors, one (row, q) pair per thread
      __syncthreads();                              // previous chunk's readers are done
      for (int t = tid; t < nrows * NQ; t += kMeshThreads) {
        const int r = t / NQ, q = t - r * NQ;
        unsigned o = (unsigned)(chunk + r);
        double x = 0.0;
        bool closing = false;
        long long base = 0;
        int zmask = 0;
        for (int d = last - 1; d >= 0; --d) {       // C-order decode over the outer axes, innermost first
          unsigned idx = o;
          if (d > 0) {
            const unsigned qq = o / (unsigned)out.cnt[d];
            idx = o - qq * (unsigned)out.cnt[d];
            o = qq;
          }
          int g = (int)idx + (d == 0 ? ks.row0 : 0);
          if (d == 0 && g == tl.closing_g) { g = 0; closing = true; }
          const double kd = ks.start[d] + (double)g / ks.den[d];                   // pythtb.py:2477
          const double c = q < NPH ? (q < nph ? ds.R[q][d] : 0.0) : -(ds.tau[q - NPH + 1][d] - ds.tau[0][d]);
          x = fma(kd, c, x);
          base += (long long)idx * out.gstride[d];
          if (idx == 0 && out.wrap[d]) zmask |= 1 << d;
        }
        s_out[r][q] = expi_turns(x);
        if (q == 0) {
          s_row[r].base = base;
          s_row[r].flags = zmask | (closing ? 256 : 0);
        }
      }
      __syncthreads();
      if (active)
      for (int r0 = 0; r0 < nrows; r0 += RPI) {
        // ---- H(k), lower triangle: straight-line over the coefficient table, RPI rows interleaved
        cplx acc[RPI][NP];
        int rr[RPI];
#pragma unroll
        for (int u = 0; u < RPI; ++u) {
          rr[u] = r0 + u < nrows ? r0 + u : nrows - 1;     // tail: recompute the last row (stored once)
#pragma unroll
          for (int e = 0; e < NP; ++e) acc[u][e] = mk(ds.C[e][0], ds.C[e][1]);
        }
#pragma unroll
        for (int p = 0; p < NPH; ++p) {
          if (p < nph) {
#pragma unroll
            for (int u = 0; u < RPI; ++u) {
              const cplx z = s_out[rr[u]][p] * fc[p];      // exp(2 pi i k.R_p)
#pragma unroll
              for (int e = 0; e < NP; ++e) {
                acc[u][e].re = fma(ds.P[p][e][0], z.re, acc[u][e].re);
                acc[u][e].re = fma(ds.Q[p][e][0], z.im, acc[u][e].re);
                acc[u][e].im = fma(ds.P[p][e][1], z.re, acc[u][e].im);
                acc[u][e].im = fma(ds.Q[p][e][1], z.im, acc[u][e].im);
              }
            }
          }
        }
        // ---- diagonalise (rows of w = eigenvectors, ascending eigenvalues)
        cplx w[RPI][N][N];
#pragma unroll
        for (int u = 0; u < RPI; ++u) {
          double ev[N];
          if constexpr (N == 2) {
            eigh2_fast(acc[u][0].re, acc[u][2].re, acc[u][1], ev, w[u]);
          } else {
            double dg[N];
            cplx lo[N * (N - 1) / 2];
#pragma unroll
            for (int a = 0; a < N; ++a) {
              dg[a] = acc[u][a * (a + 1) / 2 + a].re;
#pragma unroll
              for (int c = 0; c < a; ++c) lo[a * (a - 1) / 2 + c] = acc[u][a * (a + 1) / 2 + c];
            }
            eigh_small<N>(dg, lo, ev, w[u]);
          }
#pragma unroll
          for (int b = 0; b < N - 1; ++b) gmin[b] = fmin(gmin[b], ev[b + 1] - ev[b]);
          // ---- Convention-I gauge relative to state 0
          if (gauge) {
#pragma unroll
            for (int o = 1; o < N; ++o) {
              const cplx f = s_out[rr[u]][NPH + o - 1] * fc[NPH + o - 1];
#pragma unroll
              for (int b = 0; b < N; ++b) w[u][b][o] = w[u][b][o] * f;
            }
          }
        }
        // ---- store.  Periodic images (pythtb.py:2729-2747) and the closing-row factor of a shard are
        // applied INLINE with 256-bit stores: a row at index 0 of one wrapped outer axis writes its image
        // row (uniform branch), the lane that owns column 0 writes the image along the fastest axis.  The
        // out-of-line path (local-memory copy + a loop over axis subsets) is left to 1-D meshes and to the
        // edges of 3-D/4-D meshes where two outer axes wrap at once: taking it for every point of row 0 and
        // a read-back pass for column 0 made the CTAs that own them the tail of the single wave.
#pragma unroll
        for (int u = 0; u < RPI; ++u) {
          if (u > 0 && r0 + u >= nrows) break;
          const MeshRow mr = s_row[rr[u]];
          const int special = last == 0 ? special_j : (mr.flags | zlast);
          if (special == 0) {                       // the common case: one test, two 256-bit stores
            mesh_store_point<N, true>(dst_col + mr.base, w[u], ss);
            continue;
          }
          const int zm = mr.flags & 15;
          if (last == 0 || (zm & (zm - 1))) {
            EigRows<N> eg;
#pragma unroll
            for (int b = 0; b < N; ++b)
#pragma unroll
              for (int o = 0; o < N; ++o) eg.w[b][o] = w[u][b][o];
            mesh_store_special<N>(eg, out, s_pbc0, mr.base + (long long)j * gs_last, special & 15, special & 256);
          } else {
            if (mr.flags & 256) {                   // closing row of a shard: the axis-0 image of global row 0
#pragma unroll
              for (int o = 0; o < N; ++o) {
                const cplx ph = s_pbc0[o];
#pragma unroll
                for (int b = 0; b < N; ++b) w[u][b][o] = mul_fixed(w[u][b][o], ph);
              }
            }
            cplx* const dst = dst_col + mr.base;
            mesh_store_point<N, true>(dst, w[u], ss);
            const long long img_last = (long long)(out.full[last] - 1) * gs_last;
            if (zlast) mesh_store_image<N>(dst + img_last, w[u], s_pbc[last], ss);
            if (zm) {                               // image row along the one wrapped outer axis at index 0
              const int d = __ffs(zm) - 1;
              cplx wi[N][N];
#pragma unroll
              for (int b = 0; b < N; ++b)
#pragma unroll
                for (int o = 0; o < N; ++o) wi[b][o] = mul_fixed(w[u][b][o], s_pbc[d][o]);
              cplx* const dsti = dst + (long long)(out.full[d] - 1) * out.gstride[d];
              mesh_store_point<N, true>(dsti, wi, ss);
              if (zlast) mesh_store_image<N>(dsti + img_last, wi, s_pbc[last], ss);
            }
          }
        }
      }
    }
  }
  // ---- minimal direct gaps (pythtb.py:2484, 2529-2530): CTA partial, last CTA finishes
  if (gaps_out == nullptr) { cta_trace_end(trace, t_begin); return; }
#pragma unroll
  for (int b = 0; b < N - 1; ++b) {
    double g = gmin[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g = fmin(g, __shfl_xor_sync(0xffffffffu, g, o));
    if ((tid & 31) == 0) s_red[tid >> 5][b] = g;
  }
  __syncthreads();
  if (tid < N - 1) {
    double g = s_red[0][tid];
    for (int wv = 1; wv < kMeshThreads / 32; ++wv) g = fmin(g, s_red[wv][tid]);
    gap_partial[(size_t)blockIdx.x * (N - 1) + tid] = g;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) { cta_trace_end(trace, t_begin); return; }
  __threadfence();
#pragma unroll
  for (int b = 0; b < N - 1; ++b) {
    double g = INFINITY;
    for (int i = tid; i < (int)gridDim.x; i += kMeshThreads) g = fmin(g, __ldcg(gap_partial + (size_t)i * (N - 1) + b));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g = fmin(g, __shfl_xor_sync(0xffffffffu, g, o));
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5][b] = g;
    __syncthreads();
    if (tid == 0) {
      double t = s_red[0][b];
      for (int wv = 1; wv < kMeshThreads / 32; ++wv) t = fmin(t, s_red[wv][b]);
      if (peer.nranks > 1) s_fin[b] = t;
      else gaps_out[b] = t;
    }
  }
  if (peer.nranks > 1) {                            // minimum over the ranks through the peers' mailboxes: posted here,
    __syncthreads();                                // completed here (synchronous) or by a later kernel (deferred)
    peer_collective(peer, s_fin, N - 1, 1, gaps_out, &s_last);
  }
  if (tid == 0) signal_done(done);                  // single rank: thread 0 wrote gaps_out itself
  cta_trace_end(trace, t_begin);
}

}  // namespace tbk
