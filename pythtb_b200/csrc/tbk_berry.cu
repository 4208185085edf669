// tbk_berry.cu — Berry-phase / Wilson-loop / Berry-flux kernels and the
// position-operator (hybrid Wannier) kernels, for sm_100a.
//
// Reference code replaced (pythtb.py): _wf_dpr 3793-3796, _one_berry_loop
// 3798-3838, _one_flux_plane 3840-3865, wf_array.impose_pbc/impose_loop
// 2674-2791, tb_model.position_matrix/position_hwf 2034-2279.
//
// Formulation (verified against the reference):
//   plaquette phase = -arg[ D(a->b) D(b->c) D(c->d) D(d->a) ],  D = det of one
//   link overlap matrix (det of a product = product of dets), and
//   string phase    = -arg prod_links D.
// For nocc <= 4 a link determinant is evaluated by one thread from registers /
// local memory; for larger nocc one CTA builds the overlap matrix with all its
// threads and runs a cooperative LU.
#include "tbk_internal.cuh"
#include "tbk_berry.cuh"
#include "tbk_eig_group.cuh"

namespace tbk {

struct WfView {
  const cplx* wfs;
  int n, nsta_arr, nocc;
  const int* occ;
  long long ss;          // elements between the states of one mesh point (n for the reference layout)
};

// unit-modulus (or zero) determinant of the overlap between the occupied blocks at two mesh points
template <int NOCC>
__device__ __forceinline__ cplx link_det_small(const WfView& v, const cplx* __restrict__ pa, const cplx* __restrict__ pb) {
  cplx M[NOCC * NOCC];
#pragma unroll
  for (int i = 0; i < NOCC * NOCC; ++i) M[i] = mk(0.0, 0.0);
  const int n = v.n;
  for (int o = 0; o < n; ++o) {
    cplx a[NOCC], b[NOCC];
#pragma unroll
    for (int m = 0; m < NOCC; ++m) {
      const long long so = (long long)v.occ[m] * v.ss + o;
      a[m] = pa[so];
      b[m] = pb[so];
    }
#pragma unroll
    for (int m = 0; m < NOCC; ++m)
#pragma unroll
      for (int q = 0; q < NOCC; ++q) fma_acc_conj(M[m * NOCC + q], a[m], b[q]);
  }
  if (NOCC == 1) return M[0];
  if (NOCC == 2) return M[0] * M[3] - M[1] * M[2];
  double lg;
  return lu_det_phase(M, NOCC, NOCC, &lg);
}

__device__ __forceinline__ double neg_arg(cplx z) {
  // -numpy.angle(z): angle in (-pi, pi], so the result lies in [-pi, pi)
  if (z.re == 0.0 && z.im == 0.0) return 0.0;
  return -atan2(z.im, z.re);
}

__device__ __forceinline__ cplx unit(cplx z) {
  const double a = hypot(z.re, z.im);
  return a > 0.0 ? mk(z.re / a, z.im / a) : z;
}

// ---------------------------------------------------------------------------
// Fused plaquette kernel, nocc <= 4: one plaquette per thread.
// ---------------------------------------------------------------------------
template <int NOCC>
__global__ void __launch_bounds__(256)
flux_small_kernel(WfView v, const long long* __restrict__ slice_off, long long n0, long long stride0,
                  long long n1, long long stride1, double* __restrict__ plaq, double* __restrict__ partial) {
  const long long p1 = n1 - 1, p0 = n0 - 1;
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // fast axis
  const long long i = blockIdx.y;
  const long long s = blockIdx.z;
  double phase = 0.0;
  if (j < p1 && i < p0) {
    const cplx* base = v.wfs + slice_off[s];
    const cplx* a = base + i * stride0 + j * stride1;
    const cplx* b = a + stride0;
    const cplx* c = b + stride1;
    const cplx* d = a + stride1;
    // loop (i,j)->(i+1,j)->(i+1,j+1)->(i,j+1)->(i,j), pythtb.py:3855-3861
    cplx prod = unit(link_det_small<NOCC>(v, a, b));
    prod = prod * unit(link_det_small<NOCC>(v, b, c));
    prod = prod * unit(link_det_small<NOCC>(v, c, d));
    prod = prod * unit(link_det_small<NOCC>(v, d, a));
    phase = neg_arg(prod);
    if (plaq) plaq[(s * p0 + i) * p1 + j] = phase;
  }
  if (partial) {
    // deterministic block partial sum (fixed order): warp shuffle then shared
    __shared__ double red[8];
    double x = phase;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      partial[(s * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
  }
}


// ---------------------------------------------------------------------------
// Row-marching plaquette kernel for register-sized states (nocc <= 2, n <= 4):
// the Berry-flux kernel of BASELINE configs[1].
//
// Link-variable form: with a=(i,j), b=(i+1,j), c=(i+1,j+1),
// d=(i,j+1) and L(x,y) = det <x|y>,
//     phase(i,j) = -arg[ L(a,b) L(b,c) L(c,d) L(d,a) ]
// and L(c,d) = conj(V(i,j+1)), L(d,a) = conj(H(i,j)) where V(i,j) = L((i,j),(i+1,j)) is the
// vertical and H(i,j) = L((i,j),(i,j+1)) the horizontal link.  A WARP owns up to 31 plaquette columns
// and marches down its rows: lane t loads only ITS column's occupied states, computes V(i, col0+t)
// (the right neighbour's comes by shuffle) and H(i+1, col) against the right neighbour's state (also by
// shuffle), which it keeps in a register for the next row — two link determinants per plaquette, no
// block barrier in the row loop, one or two rows of loads in flight while a row is being reduced.
// When only the plane sum is wanted, plaquettes whose loop product z has Re z > 1/2 and
// Re z > 8 |Im z| (|arg z| < 0.125) are multiplied together — at most 24 per product, so
// |sum of args| < pi and arg(prod) = sum(arg) exactly — and ONE atan2 is taken per thread per tile;
// any other plaquette gets its own atan2.  The plane sum is finished by the last CTA (ticket) in a
// fixed order: one launch, deterministic.
// Work is dealt out per WARP, not per CTA: the plane is cut into column strips of (nearly) equal width
// <= 31 and row blocks of (nearly) equal height such that strips x blocks fills the resident warps of one
// wave.  (CTA-sized tiles of 4 x 31 columns left the last column block of a 1024-wide mesh with one busy
// warp in four and 711 tiles for 740 CTA slots: the per-CTA timeline showed CTA times from 14.3 to 18.3 us.)
// ---------------------------------------------------------------------------
constexpr int kFluxThreads = 128;
constexpr int kFluxWarpCols = 31;
constexpr int kFluxCols = kFluxWarpCols * (kFluxThreads / 32);
constexpr int kFluxMaxRows = 24;

constexpr int kFluxWarps = kFluxThreads / 32;
constexpr int kFluxDefaultSlots12 = 4;   // default prefetch rotation of flux_rows_kernel<1, 2> (measured, profiles/README.md)

struct FluxTiling {
  long long nstrip;       // column strips per slice: widths wc or wc + 1 (<= 31), the first cx strips the wider ones
  int wc, cx;
  long long nrb;          // row blocks per slice: heights hr or hr + 1, the first rx blocks the taller ones
  int hr, rx;
  long long nitems;       // nslice * nrb * nstrip warp work items (strip index fastest)
};

// One balanced wave of warp items: `resident` = #SM x CTAs per SM CTAs of kFluxWarps warps.
static FluxTiling flux_tiling(long long nslice, long long n0, long long n1, long long resident) {
  FluxTiling t;
  const long long p0 = n0 - 1, p1 = n1 - 1;
  t.nstrip = (p1 + kFluxWarpCols - 1) / kFluxWarpCols;
  t.wc = (int)(p1 / t.nstrip);
  t.cx = (int)(p1 - t.wc * t.nstrip);
  long long nrb = resident * kFluxWarps / (nslice * t.nstrip);   // row blocks per strip
  if (nrb > p0 / 4) nrb = p0 / 4;                                // keep the per-item first-row loads amortised
  if (nrb < 1) nrb = 1;
  t.nrb = nrb;
  t.hr = (int)(p0 / nrb);
  t.rx = (int)(p0 - t.hr * nrb);
  t.nitems = nslice * t.nrb * t.nstrip;
  return t;
}
// upper bound of nitems over every `resident` the launcher may use (workspace sizing)
static long long flux_tiles_bound(long long nslice, long long n0, long long n1) {
  const long long nstrip = (n1 - 1 + kFluxWarpCols - 1) / kFluxWarpCols;
  const long long a = nslice * nstrip * ((n0 - 1) / 4 + 1);
  const long long b = (long long)kNumSM * 16 * kFluxWarps + nslice * nstrip;
  return a < b ? a : b;
}

template <int NOCC, int N>
struct OccState {
  cplx u[NOCC][N];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int m = 0; m < NOCC; ++m)
#pragma unroll
      for (int o = 0; o < N; ++o) u[m][o] = mk(0.0, 0.0);
  }
  __device__ __forceinline__ void load(const cplx* __restrict__ p, const long long* occ) {   // occ[m]: element offset of state m
#pragma unroll
    for (int m = 0; m < NOCC; ++m) {
      const double2* src = reinterpret_cast<const double2*>(p + occ[m]);
#pragma unroll
      for (int o = 0; o < N; ++o) { const double2 t = __ldg(src + o); u[m][o] = mk(t.x, t.y); }
    }
  }
};

// det <a|b> for NOCC <= 2
template <int NOCC, int N>
__device__ __forceinline__ cplx link_det(const OccState<NOCC, N>& a, const OccState<NOCC, N>& b) {
  cplx M[NOCC][NOCC];
#pragma unroll
  for (int m = 0; m < NOCC; ++m)
#pragma unroll
    for (int q = 0; q < NOCC; ++q) {
      cplx acc = mk(0.0, 0.0);
#pragma unroll
      for (int o = 0; o < N; ++o) fma_acc_conj(acc, a.u[m][o], b.u[q][o]);
      M[m][q] = acc;
    }
  if constexpr (NOCC == 1) return M[0][0];
  else return M[0][0] * M[1][1] - M[0][1] * M[1][0];
}

// the state held by the next lane
template <int NOCC, int N>
__device__ __forceinline__ OccState<NOCC, N> shfl_down_state(const OccState<NOCC, N>& a) {
  OccState<NOCC, N> r;
#pragma unroll
  for (int m = 0; m < NOCC; ++m)
#pragma unroll
    for (int o = 0; o < N; ++o) {
      r.u[m][o].re = __shfl_down_sync(0xffffffffu, a.u[m][o].re, 1);
      r.u[m][o].im = __shfl_down_sync(0xffffffffu, a.u[m][o].im, 1);
    }
  return r;
}

// SLOTS: register slots of the row rotation (rows i, i+1 and SLOTS - 2 rows of loads in flight); 0 = the default
// (4 for states of <= 4 components, 3 above).  With a state-major array a row of the one-band, two-orbital case is
// 32 useful bytes per lane, and deeper rotations (6, 8 slots: 4 / 6 rows in flight) trade occupancy for bytes in flight.
template <int NOCC, int N, bool WANT_PLAQ, int SLOTS = 0>
__global__ void __launch_bounds__(kFluxThreads, (NOCC == 1 && N == 2) ? (SLOTS > 4 ? 4 : 5) : (NOCC * N >= 6 ? 3 : 1))   // 5 CTAs / SM for the Haldane case (<= 102 registers), 3 for the 6- and 8-component states (<= 168)
flux_rows_kernel(WfView v, const long long* __restrict__ slice_off, long long n0, long long stride0, long long n1,
                 long long stride1, FluxTiling tl, long long nslice, double* __restrict__ plaq,
                 double* __restrict__ partial, unsigned* __restrict__ ticket, double* __restrict__ total,
                 const __grid_constant__ PeerView peer, unsigned long long* __restrict__ trace, const DoneSignal done) {
  __shared__ double s_red[kFluxThreads / 32];
  __shared__ double s_fin[kPeerMaxVals];
  __shared__ int s_last;
  const unsigned long long t_begin = cta_trace_begin(trace);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Programmatic dependent launch: this grid may have been scheduled while the previous kernel on the
  // stream (typically the grid solve that writes the array read here) was still draining; everything it
  // wrote is visible after this wait.  A no-op when launched without the attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  peer_prologue(peer);                              // first CTA: post what an earlier kernel of this step kept locally (N > 1)
  long long occ[NOCC];
#pragma unroll
  for (int m = 0; m < NOCC; ++m) occ[m] = (long long)v.occ[m] * v.ss;
  const long long p0 = n0 - 1, p1 = n1 - 1;
  const long long per_slice = tl.nrb * tl.nstrip;
  // one slice (the usual 2-D mesh): a warp keeps adding its items up and the CTA writes ONE partial, so the
  // last CTA has gridDim.x values to sum instead of one per item; several slices: one partial per item
  const bool cta_partial = nslice == 1;
  double wacc = 0.0;
  for (long long item = (long long)blockIdx.x * kFluxWarps + warp; item < tl.nitems; item += (long long)gridDim.x * kFluxWarps) {
    const long long s = item / per_slice;
    const long long rem = item - s * per_slice;
    const long long rb = rem / tl.nstrip, strip = rem - rb * tl.nstrip;
    const int width = tl.wc + (strip < tl.cx ? 1 : 0);                    // plaquette columns of this strip
    const long long col = strip * tl.wc + (strip < tl.cx ? strip : tl.cx) + lane;   // mesh column of this lane's vertical link
    const bool has_col = lane <= width;                                   // col <= strip end <= p1 < n1
    const bool owner = lane < width;                                      // owns plaquette column `col`
    const long long i0 = rb * tl.hr + (rb < tl.rx ? rb : tl.rx);
    const int nrow = tl.hr + (rb < tl.rx ? 1 : 0);
    const cplx* pa = v.wfs + slice_off[s] + col * stride1 + i0 * stride0;  // u(i0, col)
    double acc = 0.0;
    {
      // ---- every lane loads ONLY its own column, kSlots - 2 rows ahead (a rotation of kSlots register
      // slots: rows i, i+1 and the rows in flight), and takes the right-hand neighbour's state for the
      // horizontal link by shuffle.  The kernel is bound by the bytes a warp keeps in flight (per-CTA
      // timeline + Little's law: one row ahead = 2 KB per warp = ~40 KB per SM gave ~3.9 TB/s for the
      // 2-component state); dropping the neighbour-column registers pays for the deeper prefetch.
      // Small states (nocc x n <= 4): two rows ahead; larger ones: one row ahead, the registers go to occupancy.
      constexpr int kSlots = SLOTS > 0 ? SLOTS : ((NOCC * N <= 4) ? 4 : 3);
      OccState<NOCC, N> X[kSlots];
#pragma unroll
      for (int q = 0; q < kSlots; ++q) X[q].zero();        // lanes past the strip carry zeros, never garbage
      if (has_col) {
        X[0].load(pa, occ);
        X[1].load(pa + stride0, occ);
#pragma unroll
        for (int q = 2; q <= kSlots - 2; ++q)              // rows 0 .. kSlots - 2 are in flight before the first step
          if (nrow >= q) X[q].load(pa + q * stride0, occ);
      }
      cplx hda;                                            // L(d,a) = conj(H(i,col))
      {
        const OccState<NOCC, N> d = shfl_down_state(X[0]);
        hda = conj(link_det<NOCC, N>(X[0], d));
      }
      cplx prod = mk(1.0, 0.0);
      double* pq = WANT_PLAQ ? plaq + (s * p0 + i0) * p1 + col : nullptr;
      const cplx* pf = pa + (kSlots - 1) * stride0;        // row i0 + r + kSlots - 1 at step r
#pragma unroll 1
      for (int r = 0; r < nrow; r += kSlots) {
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
          if (r + u < nrow) {
            if (has_col && r + u + kSlots - 1 <= nrow) X[(u + kSlots - 1) % kSlots].load(pf, occ);
            pf += stride0;
            const OccState<NOCC, N>& A = X[u];
            const OccState<NOCC, N>& B = X[(u + 1) % kSlots];
            const cplx lab = link_det<NOCC, N>(A, B);      // V(i, col); junk in lanes without a column
            cplx right;                                    // V(i, col+1) from the next lane
            right.re = __shfl_down_sync(0xffffffffu, lab.re, 1);
            right.im = __shfl_down_sync(0xffffffffu, lab.im, 1);
            const OccState<NOCC, N> C = shfl_down_state(B);
            const cplx lbc = link_det<NOCC, N>(B, C);      // H(i+1, col)
            cplx z = lab * lbc;
            z = mulc(z, right);                            // L(c,d) = conj(V(i, col+1))
            z = z * hda;
            hda = conj(lbc);
            if (owner) {
              if (WANT_PLAQ) {
                const double phase = neg_arg(z);
                pq[0] = phase;
                pq += p1;
                acc += phase;
              } else if (z.re > 0.5 && z.re > 8.0 * fabs(z.im)) {
                prod = prod * z;
              } else {
                acc += neg_arg(z);
              }
            }
          }
        }
        if (!WANT_PLAQ && ((r + kSlots) % 24) == 0) {      // at most 24 small angles per product
          if (owner) acc += neg_arg(prod);
          prod = mk(1.0, 0.0);
        }
      }
      if (!WANT_PLAQ && owner) acc += neg_arg(prod);
    }
    if (partial) {
      // fixed-order warp sum -> partial[item]
      double x = acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (cta_partial) wacc += x;
      else if (lane == 0) partial[item] = x;
    }
  }
  if (partial && cta_partial) {
    if (lane == 0) s_red[warp] = wacc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kFluxWarps; ++w) t += s_red[w];
      partial[blockIdx.x] = t;
    }
  }
  if (!partial) { cta_trace_end(trace, t_begin); return; }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (int)atomicInc(ticket, gridDim.x - 1);          // order of finishing
  __syncthreads();
  const int tick = s_last;
  __syncthreads();
  // the FIRST CTA to finish completes the older deferred cross-rank reductions (N > 1) while the wave drains
  if (tick == 0) peer_complete_pending(peer, &s_last);
  if (tick != (int)gridDim.x - 1) { cta_trace_end(trace, t_begin); return; }
  __threadfence();
  const long long count = cta_partial ? (long long)gridDim.x : per_slice;
  for (long long s = 0; s < nslice; ++s) {
    double x = 0.0;
    for (long long i = tid; i < count; i += kFluxThreads) x += __ldcg(partial + s * count + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kFluxThreads / 32; ++w) t += s_red[w];
      if (peer.nranks > 1) s_fin[s] = t;
      else total[s] = t;
    }
  }
  if (peer.nranks > 1) {                                  // sum over the ranks: synchronous (post + complete here) or
    __syncthreads();                                      // deferred (kept locally, posted by the next kernel's first CTA)
    peer_own(peer, s_fin, (int)nslice, 0, total, &s_last);
  }
  if (tid == 0) signal_done(done);                        // single rank: thread 0 wrote total[] itself
  cta_trace_end(trace, t_begin);
}

template <int NOCC, int N, int SLOTS = 0>
static int launch_flux_rows(const WfView& v, const long long* off, long long nslice, long long n0, long long stride0,
                            long long n1, long long stride1, double* plaq, double* total, double* partial, tbk_peer* peer,
                            cudaStream_t st) {
  static int occ_plaq = 0, occ_sum = 0;                   // resident CTAs per SM of the two variants
  if (occ_plaq == 0) {
    int a = 0, b = 0;
    TBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, flux_rows_kernel<NOCC, N, true, SLOTS>, kFluxThreads, 0));
    TBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, flux_rows_kernel<NOCC, N, false, SLOTS>, kFluxThreads, 0));
    occ_sum = b > 0 ? (b > 16 ? 16 : b) : 1;
    occ_plaq = a > 0 ? (a > 16 ? 16 : a) : 1;
  }
  const long long resident = (long long)kNumSM * (plaq ? occ_plaq : occ_sum);
  const FluxTiling tl = flux_tiling(nslice, n0, n1, resident);
  unsigned* ticket = nullptr;
  if (total) {
    ticket = take_ticket();
    if (!ticket) { set_error("tbk_flux_plane: cannot allocate the reduction tickets"); return TBK_ERR_CUDA; }
  }
  const long long ctas = (tl.nitems + kFluxWarps - 1) / kFluxWarps;
  const int grid = (int)(ctas < resident ? ctas : resident);
  PeerView pview = peer_none();                             // cross-rank sum: synchronous or posted (tbk_peer_defer); either way
  if (total) { if (int rc = peer_next(peer, (int)nslice, 0, total, true, st, &pview)) return rc; }   // it completes older deferred ones
  // launched as a programmatic dependent of the previous kernel on the stream: its CTAs are scheduled as that
  // kernel's CTAs retire and wait at griddepcontrol.wait, which hides this kernel's launch latency and ramp
  // behind the predecessor's tail (TBK_PDL=0 turns it off)
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("TBK_PDL"); pdl = (e && atoi(e) == 0) ? 0 : 1; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kFluxThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  double* part_arg = plaq ? (total ? partial : nullptr) : partial;
  unsigned long long* trace = cta_trace_buffer();
  if (trace) trace += (size_t)4 * kCtaTraceCap;          // the flux kernel's half of the trace buffer
  // a synchronous prepared call waits on a pinned word the last CTA writes after the totals
  const DoneSignal done = (total && (pview.nranks <= 1 || pview.mode == 1)) ? take_done_request() : DoneSignal{nullptr, 0};
  if (plaq)
    TBK_CUDA(cudaLaunchKernelEx(&cfg, flux_rows_kernel<NOCC, N, true, SLOTS>, v, off, n0, stride0, n1, stride1, tl, nslice, plaq,
                                part_arg, ticket, total, pview, trace, done));
  else
    TBK_CUDA(cudaLaunchKernelEx(&cfg, flux_rows_kernel<NOCC, N, false, SLOTS>, v, off, n0, stride0, n1, stride1, tl, nslice, plaq,
                                part_arg, ticket, total, pview, trace, done));
  TBK_LAUNCH_CHECK("flux_rows_kernel");
  return TBK_OK;
}

// ---------------------------------------------------------------------------
// Bulk-copy ring variant of the plaquette kernel (contiguous mesh columns): the rows of a tile
// are streamed into a shared-memory ring by cp.async.bulk (the TMA engine, 1-D form) signalling
// one mbarrier per stage, so the bytes in flight per SM are set by the ring depth (tens of KB)
// instead of by the registers a thread can spare for prefetching — ncu showed the register
// version waiting on loads for 59 % of its cycles with ~20 KB in flight per SM.
// A CTA owns 128 plaquette columns x `ti` rows; thread t owns plaquette column t and reads the
// states of columns t and t+1 from the ring, so every k-point block is fetched from HBM once per
// tile (plus one halo row), and every thread owns a plaquette (no shuffle, no idle lane).
// Per plaquette: V(i,j) = det<a|b>, V(i,j+1) = det<d|c>, H(i+1,j) = det<b|c>, H(i,j) carried:
//     phase = -arg[ V(i,j) H(i+1,j) conj V(i,j+1) conj H(i,j) ]        (pythtb.py:3855-3861)
// ---------------------------------------------------------------------------
constexpr int kRingThreads = 128;
constexpr int kRingMaxStages = 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  const unsigned a = smem_u32(bar);
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}

struct RingTiling {
  int ti;                 // plaquette rows per tile
  int stages;
  unsigned stage_bytes;
  long long nbx, nrb, ntiles;
};

template <int NOCC, int N>
struct RingState {
  cplx u[NOCC][N];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int m = 0; m < NOCC; ++m)
#pragma unroll
      for (int o = 0; o < N; ++o) u[m][o] = mk(0.0, 0.0);
  }
  __device__ __forceinline__ void load(const cplx* p, const int (&occ)[NOCC]) {     // shared memory
#pragma unroll
    for (int m = 0; m < NOCC; ++m)
#pragma unroll
      for (int o = 0; o < N; ++o) u[m][o] = p[occ[m] * N + o];
  }
};
template <int NOCC, int N>
__device__ __forceinline__ cplx ring_link(const RingState<NOCC, N>& a, const RingState<NOCC, N>& b) {
  cplx M[NOCC][NOCC];
#pragma unroll
  for (int m = 0; m < NOCC; ++m)
#pragma unroll
    for (int q = 0; q < NOCC; ++q) {
      cplx acc = mk(0.0, 0.0);
#pragma unroll
      for (int o = 0; o < N; ++o) fma_acc_conj(acc, a.u[m][o], b.u[q][o]);
      M[m][q] = acc;
    }
  if constexpr (NOCC == 1) return M[0][0];
  else return M[0][0] * M[1][1] - M[0][1] * M[1][0];
}

template <int NOCC, int N, bool WANT_PLAQ>
__global__ void __launch_bounds__(kRingThreads)
flux_ring_kernel(WfView v, const long long* __restrict__ slice_off, long long n0, long long stride0, long long n1,
                 RingTiling tl, long long nslice, double* __restrict__ plaq, double* __restrict__ partial,
                 unsigned* __restrict__ ticket, double* __restrict__ total, const __grid_constant__ PeerView peer) {
  extern __shared__ __align__(128) char ring[];
  __shared__ __align__(8) unsigned long long s_bar[kRingMaxStages];
  __shared__ double s_red[kRingThreads / 32];
  __shared__ double s_fin[kPeerMaxVals];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = tl.stages;
  const int blk = v.nsta_arr * v.n;                         // complex elements per mesh point
  int occ[NOCC];
#pragma unroll
  for (int m = 0; m < NOCC; ++m) occ[m] = v.occ[m];
  if (tid == 0) {
    for (int sg = 0; sg < D; ++sg) mbar_init(&s_bar[sg], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  peer_prologue(peer);
  const long long p0 = n0 - 1, p1 = n1 - 1;
  unsigned q_cons = 0;                                      // rows consumed so far by this CTA (stage = q % D)
  for (long long tile = blockIdx.x; tile < tl.ntiles; tile += gridDim.x) {
    const long long s = tile / (tl.nrb * tl.nbx);
    const long long rem = tile - s * tl.nrb * tl.nbx;
    const long long rb = rem / tl.nbx, bx = rem - rb * tl.nbx;
    const long long c0 = bx * kRingThreads;
    const long long col = c0 + tid;
    const bool owner = col < p1;
    const long long i0 = rb * tl.ti;
    const int nrow = (int)((i0 + tl.ti < p0 ? i0 + tl.ti : p0) - i0);     // plaquette rows; nrow + 1 mesh rows
    const int ncl = (int)(n1 - c0 < kRingThreads + 1 ? n1 - c0 : kRingThreads + 1);
    const unsigned row_bytes = (unsigned)ncl * (unsigned)blk * 16u;
    const cplx* g0 = v.wfs + slice_off[s] + c0 * (long long)blk + i0 * stride0;
    // ---- producer prologue: the first min(D, nrow + 1) rows of the tile
    if (tid == 0) {
      const int pre = nrow + 1 < D ? nrow + 1 : D;
      for (int r = 0; r < pre; ++r) {
        const unsigned sg = (q_cons + r) % D;
        mbar_expect_tx(&s_bar[sg], row_bytes);
        bulk_g2s(ring + (size_t)sg * tl.stage_bytes, g0 + (long long)r * stride0, row_bytes, &s_bar[sg]);
      }
    }
    RingState<NOCC, N> A, Ad, B, Bc;
    cplx hda = mk(1.0, 0.0);
    double acc = 0.0;
    cplx prod = mk(1.0, 0.0);
    int nprod = 0;
    double* pq = WANT_PLAQ ? plaq + (s * p0 + i0) * p1 + col : nullptr;
    for (int r = 0; r <= nrow; ++r) {
      const unsigned sg = q_cons % D;
      mbar_wait(&s_bar[sg], (q_cons / D) & 1u);
      if (owner) {
        const cplx* row = reinterpret_cast<const cplx*>(ring + (size_t)sg * tl.stage_bytes) + (size_t)tid * blk;
        B.load(row, occ);
        Bc.load(row + blk, occ);
        const cplx hn = ring_link<NOCC, N>(B, Bc);         // H(i0 + r, col)
        if (r > 0) {
          cplx z = ring_link<NOCC, N>(A, B) * hn;          // V(i, col) H(i+1, col)
          z = mulc(z, ring_link<NOCC, N>(Ad, Bc));         // conj V(i, col+1)
          z = z * hda;                                     // conj H(i, col)
          if (WANT_PLAQ) {
            const double phase = neg_arg(z);
            pq[0] = phase;
            pq += p1;
            acc += phase;
          } else if (z.re > 0.5 && z.re > 8.0 * fabs(z.im)) {
            prod = prod * z;                               // |arg z| < 0.125: at most 24 per product
            if (++nprod == 24) { acc += neg_arg(prod); prod = mk(1.0, 0.0); nprod = 0; }
          } else {
            acc += neg_arg(z);
          }
        }
        hda = conj(hn);
        A = B;
        Ad = Bc;
      }
      ++q_cons;
      __syncthreads();                                     // every thread is done with stage sg
      if (tid == 0 && r + D <= nrow) {                     // refill it with row r + D of this tile
        mbar_expect_tx(&s_bar[sg], row_bytes);
        bulk_g2s(ring + (size_t)sg * tl.stage_bytes, g0 + (long long)(r + D) * stride0, row_bytes, &s_bar[sg]);
      }
    }
    if (!WANT_PLAQ && owner && nprod) acc += neg_arg(prod);
    if (partial) {
      double x = acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) s_red[warp] = x;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kRingThreads / 32; ++w) t += s_red[w];
        partial[tile] = t;
      }
      __syncthreads();
    }
  }
  if (!partial) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const long long per = tl.nrb * tl.nbx;
  for (long long s = 0; s < nslice; ++s) {
    double x = 0.0;
    for (long long i = tid; i < per; i += kRingThreads) x += __ldcg(partial + s * per + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kRingThreads / 32; ++w) t += s_red[w];
      if (peer.nranks > 1) s_fin[s] = t;
      else total[s] = t;
    }
  }
  if (peer.nranks > 1) {                                  // sum over the ranks, through the peers' mailboxes
    __syncthreads();
    peer_collective(peer, s_fin, (int)nslice, 0, total, &s_last);
  }
}

// tiles of the ring kernel: one balanced wave of `resident` CTAs where the mesh allows it
static RingTiling ring_tiling(long long nslice, long long n0, long long n1, int blk, long long resident, int stages,
                              unsigned stage_bytes) {
  RingTiling t;
  t.stages = stages;
  t.stage_bytes = stage_bytes;
  t.nbx = (n1 - 1 + kRingThreads - 1) / kRingThreads;
  const long long p0 = n0 - 1;
  long long per_col = resident / (nslice * t.nbx);
  if (per_col < 1) per_col = 1;
  if (per_col > p0) per_col = p0;
  long long ti = (p0 + per_col - 1) / per_col;
  if (ti < 8 && p0 >= 8) ti = 8;                          // one halo row per tile: keep the re-read below ~12 %
  t.ti = (int)ti;
  t.nrb = (p0 + ti - 1) / ti;
  t.ntiles = nslice * t.nrb * t.nbx;
  (void)blk;
  return t;
}
static long long ring_tiles_bound(long long nslice, long long n0, long long n1) {
  const long long nbx = (n1 - 1 + kRingThreads - 1) / kRingThreads;
  return nslice * nbx * ((n0 - 1 + 7) / 8 + 1) + (long long)kNumSM * 16;
}

template <int NOCC, int N>
static int launch_flux_ring(const WfView& v, const long long* off, long long nslice, long long n0, long long stride0,
                            long long n1, double* plaq, double* total, double* partial, tbk_peer* peer, cudaStream_t st) {
  const int blk = v.nsta_arr * v.n;
  const unsigned stage_bytes = (unsigned)((((size_t)(kRingThreads + 1) * blk * 16) + 127) & ~(size_t)127);
  const int stages = stage_bytes <= 12 * 1024 ? 4 : (stage_bytes <= 40 * 1024 ? 3 : 2);
  const size_t dyn = (size_t)stages * stage_bytes;
  auto kern_p = flux_ring_kernel<NOCC, N, true>;
  auto kern_s = flux_ring_kernel<NOCC, N, false>;
  static int occ_p = 0, occ_s = 0;
  static size_t occ_dyn = 0;
  if (occ_p == 0 || occ_dyn != dyn) {
    TBK_CUDA(cudaFuncSetAttribute(kern_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    TBK_CUDA(cudaFuncSetAttribute(kern_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int a = 0, b = 0;
    TBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, kern_p, kRingThreads, dyn));
    TBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern_s, kRingThreads, dyn));
    occ_p = a > 0 ? (a > 6 ? 6 : a) : 1;
    occ_s = b > 0 ? (b > 6 ? 6 : b) : 1;
    occ_dyn = dyn;
  }
  const long long resident = (long long)kNumSM * (plaq ? occ_p : occ_s);
  const RingTiling tl = ring_tiling(nslice, n0, n1, blk, resident, stages, stage_bytes);
  unsigned* ticket = nullptr;
  if (total) {
    ticket = take_ticket();
    if (!ticket) { set_error("tbk_flux_plane: cannot allocate the reduction tickets"); return TBK_ERR_CUDA; }
  }
  const int grid = (int)(tl.ntiles < resident ? tl.ntiles : resident);
  PeerView pview = peer_none();
  if (total) { if (int rc = peer_next(peer, (int)nslice, 0, total, true, st, &pview)) return rc; }
  if (plaq)
    kern_p<<<grid, kRingThreads, dyn, st>>>(v, off, n0, stride0, n1, tl, nslice, plaq, total ? partial : nullptr, ticket,
                                            total, pview);
  else
    kern_s<<<grid, kRingThreads, dyn, st>>>(v, off, n0, stride0, n1, tl, nslice, plaq, partial, ticket, total, pview);
  TBK_LAUNCH_CHECK("flux_ring_kernel");
  note_kernel("flux_ring_kernel");
  return TBK_OK;
}

// sum partial[s][0..count) in a fixed order -> total[s]
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double* __restrict__ partial, long long count, double* __restrict__ total) {
  __shared__ double red[256];
  const long long s = blockIdx.x;
  double x = 0.0;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) x += partial[s * count + i];
  red[threadIdx.x] = x;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[s] = red[0];
}

// ---------------------------------------------------------------------------
// General nocc: one CTA per link.  Links are enumerated by a LinkMap.
// ---------------------------------------------------------------------------
struct LinkMap {
  // link l -> (string/slice s, position t, direction dir): endpoints
  //   a = off[s] + t0*stride0 + t1*stride1 ,  b = a + (dir ? stride1 : stride0)
  const long long* off;
  long long n0, n1, stride0, stride1;
  int two_dirs;      // 1: plane (links in both directions), 0: strings along stride0
};

// number of links per slice
__host__ __device__ inline long long links_per_slice(const LinkMap& m) {
  return m.two_dirs ? (m.n0 - 1) * m.n1 + m.n0 * (m.n1 - 1) : (m.n0 - 1);
}

__device__ inline void link_endpoints(const LinkMap& m, long long l, long long& a, long long& b) {
  const long long per = links_per_slice(m);
  const long long s = l / per;
  long long r = l - s * per;
  const long long base = m.off[s];
  if (!m.two_dirs) { a = base + r * m.stride0; b = a + m.stride0; return; }
  const long long nx = (m.n0 - 1) * m.n1;        // links along axis 0: index (i, j), i < n0-1
  if (r < nx) {
    const long long i = r / m.n1, j = r - i * m.n1;
    a = base + i * m.stride0 + j * m.stride1; b = a + m.stride0;
  } else {
    r -= nx;                                      // links along axis 1: index (i, j), j < n1-1
    const long long i = r / (m.n1 - 1), j = r - i * (m.n1 - 1);
    a = base + i * m.stride0 + j * m.stride1; b = a + m.stride1;
  }
}

// ---------------------------------------------------------------------------
// Overlap matrix of one link on the FP64 tensor pipe:  M[m][q] = sum_o conj(A[m][o]) B[q][o]
// (_wf_dpr, pythtb.py:3793-3796, for all nocc^2 pairs of a link at once).  A, B are the occupied
// rows (row stride n) of the two mesh points.  One CTA of 8 warps computes M in 64 x 64 output tiles;
// a warp owns 16 x 32 of it as 2 x 4 fragments of mma.sync.m8n8k4.f64 (DMMA); the K (orbital)
// dimension is staged 16 orbitals at a time in shared memory as separate re / im planes with a
// leading dimension of 20 doubles, which makes every 8-byte fragment read conflict-free.  A complex
// product is four real DMMAs:  re += Ar Br, re += Ai Bi, im += Ar Bi, im += (-Ai) Br.
// ---------------------------------------------------------------------------
constexpr int kOvTile = 64, kOvKC = 16, kOvLD = 20;
struct OvSmem {
  double are[kOvTile][kOvLD], aim[kOvTile][kOvLD], bre[kOvTile][kOvLD], bim[kOvTile][kOvLD];
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One operand of the CTA-level complex GEMM: element (row r, contraction index x) lives at
// p[rowmap ? rowmap[r] : r) * rs + x * xs]; conj != 0 uses its complex conjugate.
struct GemmSide {
  const cplx* p;
  long long rs, xs;
  const int* rowmap;
  int conj;
};

// C[m][q] = sum_x a(m, x) * b(q, x) * (wgt ? wgt[x] : 1),  m < ma, q < nb_, x < kdim;  C row-major [ma][ldc].
// Called by all 256 threads of a CTA.  The loader picks the thread -> element mapping that makes the global
// reads contiguous for the operand's unit stride (rows or contraction index).
// Software pipeline: the global loads of K-chunk i + 1 are issued into registers BEFORE the DMMAs of chunk i and
// written to shared memory after them, so a chunk's ~1 us of L2 / HBM latency hides under the previous chunk's
// tensor-pipe work (the single-stage version waited for every chunk: DMMA pipe 18-65 % busy, profiles/README.md r06).
// Not inlined: the callers (link overlap + LU, Newton-Schulz, tree products, position matrices) would otherwise each
// carry the 64 accumulator registers through their own code (254 registers, one CTA per SM).
// A/B knobs (profiles/build_variant.py): resident CTAs per SM the DMMA kernels are compiled for, and the software
// pipeline of the GEMM (0 = the single-stage loop of round 1: load, barrier, DMMAs, barrier)
#ifndef TBK_GEMM_MINB
#define TBK_GEMM_MINB 2
#endif
#ifndef TBK_GEMM_PIPELINE
#define TBK_GEMM_PIPELINE 1
#endif
struct GemmStage {
  cplx a[4], b[4];
};

// Loader mapping (element idx = tid + 256 i of a 64 x 16 chunk, i < 4): unit contraction stride -> row = idx / 16,
// xc = idx % 16 (16 lanes read 256 contiguous bytes of a row, and the shared-memory stores are conflict-free);
// unit row stride -> row = idx % 64, xc = idx / 64.  Everything that does not depend on the K chunk — the four rows, their
// bounds test, the row map (occupied-state index) and the base pointers — is computed once per output tile (GemmTile),
// not once per loaded element: in the first version that index arithmetic was ~45 % of the kernel's instructions and
// the DMMAs 2 % (ncu source page, profiles/r11/lines_large.txt).
struct GemmTile {
  const cplx* pa[4];    // element i of A at chunk 0, nullptr if its row is outside the matrix
  const cplx* pb[4];
  int xa0, dxa, xb0, dxb;   // contraction index of element i inside the chunk: x0 + i * dx
  int ra0, dra, rb0, drb;   // row of element i inside the tile
};

__device__ __forceinline__ GemmTile gemm_tile_setup(const GemmSide& A, int ma, int m0, const GemmSide& B, int nb_, int q0) {
  const int tid = threadIdx.x;
  GemmTile t;
  if (A.xs == 1) { t.ra0 = tid >> 4; t.dra = 16; t.xa0 = tid & 15; t.dxa = 0; }
  else { t.ra0 = tid & 63; t.dra = 0; t.xa0 = tid >> 6; t.dxa = 4; }
  if (B.xs == 1) { t.rb0 = tid >> 4; t.drb = 16; t.xb0 = tid & 15; t.dxb = 0; }
  else { t.rb0 = tid & 63; t.drb = 0; t.xb0 = tid >> 6; t.dxb = 4; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ra = m0 + t.ra0 + i * t.dra, rb = q0 + t.rb0 + i * t.drb;
    t.pa[i] = t.pb[i] = nullptr;
    if (ra < ma) {
      const long long r = A.rowmap ? A.rowmap[ra] : ra;
      t.pa[i] = A.p + r * A.rs + (long long)(t.xa0 + i * t.dxa) * A.xs;
    }
    if (rb < nb_) {
      const long long r = B.rowmap ? B.rowmap[rb] : rb;
      t.pb[i] = B.p + r * B.rs + (long long)(t.xb0 + i * t.dxb) * B.xs;
    }
  }
  return t;
}

__device__ __forceinline__ void gemm_stage_load(GemmStage& st, const GemmTile& t, long long xsa, long long xsb, int kdim, int x0,
                                                const double* __restrict__ wgt) {
  const cplx zero = mk(0.0, 0.0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int xa = x0 + t.xa0 + i * t.dxa, xb = x0 + t.xb0 + i * t.dxb;
    st.a[i] = (t.pa[i] && xa < kdim) ? t.pa[i][(long long)x0 * xsa] : zero;
    cplx b = (t.pb[i] && xb < kdim) ? t.pb[i][(long long)x0 * xsb] : zero;
    if (wgt && xb < kdim) { const double f = wgt[xb]; b.re *= f; b.im *= f; }
    st.b[i] = b;
  }
}

__device__ __forceinline__ void gemm_stage_store(const GemmStage& st, const GemmTile& t, double sa, double sb, OvSmem& sm) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ra = t.ra0 + i * t.dra, xa = t.xa0 + i * t.dxa, rb = t.rb0 + i * t.drb, xb = t.xb0 + i * t.dxb;
    sm.are[ra][xa] = st.a[i].re; sm.aim[ra][xa] = sa * st.a[i].im;
    sm.bre[rb][xb] = st.b[i].re; sm.bim[rb][xb] = sb * st.b[i].im;
  }
}

// herm: the product is Hermitian (B is A's conjugate partner, ma == nb_: X^H X, <u| r |u>): only the tiles on and above
// the diagonal are computed, the others are written as conjugate transposes — half the DMMAs.
__device__ __noinline__ void cta_gemm_dmma(const GemmSide& A, int ma, const GemmSide& B, int nb_, int kdim, const double* __restrict__ wgt,
                                           cplx* __restrict__ C, int ldc, OvSmem& sm, bool herm = false) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp >> 1, wc = warp & 1;
  const double sa = A.conj ? -1.0 : 1.0, sb = B.conj ? -1.0 : 1.0;
  for (int m0 = 0; m0 < ma; m0 += kOvTile) {
    for (int q0 = herm ? m0 : 0; q0 < nb_; q0 += kOvTile) {
      double cre[2][4][2], cim[2][4][2];
#pragma unroll
      for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) { cre[rt][ct][0] = cre[rt][ct][1] = 0.0; cim[rt][ct][0] = cim[rt][ct][1] = 0.0; }
      bool rv[2], cv[4];                         // warp-uniform: does the 8 x 8 fragment touch the matrix at all
#pragma unroll
      for (int rt = 0; rt < 2; ++rt) rv[rt] = m0 + wr * 16 + rt * 8 < ma;
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) cv[ct] = q0 + wc * 32 + ct * 8 < nb_;
      GemmStage st;
      const GemmTile tile = gemm_tile_setup(A, ma, m0, B, nb_, q0);
#if TBK_GEMM_PIPELINE
      gemm_stage_load(st, tile, A.xs, B.xs, kdim, 0, wgt);
#endif
      for (int x0 = 0; x0 < kdim; x0 += kOvKC) {
#if !TBK_GEMM_PIPELINE
        gemm_stage_load(st, tile, A.xs, B.xs, kdim, x0, wgt);
#endif
        __syncthreads();                         // the previous chunk has been consumed
        gemm_stage_store(st, tile, sa, sb, sm);
        __syncthreads();
#if TBK_GEMM_PIPELINE
        if (x0 + kOvKC < kdim) gemm_stage_load(st, tile, A.xs, B.xs, kdim, x0 + kOvKC, wgt);   // in flight under the DMMAs below
#endif
#pragma unroll
        for (int ks = 0; ks < kOvKC / 4; ++ks) {
          double ar[2], ai[2], an[2];
#pragma unroll
          for (int rt = 0; rt < 2; ++rt) {
            const int r = wr * 16 + rt * 8 + g;
            ar[rt] = sm.are[r][ks * 4 + t];
            ai[rt] = sm.aim[r][ks * 4 + t];
            an[rt] = -ai[rt];
          }
#pragma unroll
          for (int ct = 0; ct < 4; ++ct) {
            if (!cv[ct]) continue;
            const int c = wc * 32 + ct * 8 + g;
            const double br = sm.bre[c][ks * 4 + t], bi = sm.bim[c][ks * 4 + t];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
              if (!rv[rt]) continue;
              dmma884(cre[rt][ct][0], cre[rt][ct][1], ar[rt], br);    // (ar + i ai)(br + i bi)
              dmma884(cre[rt][ct][0], cre[rt][ct][1], an[rt], bi);
              dmma884(cim[rt][ct][0], cim[rt][ct][1], ar[rt], bi);
              dmma884(cim[rt][ct][0], cim[rt][ct][1], ai[rt], br);
            }
          }
        }
      }
#pragma unroll
      for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
          const int m = m0 + wr * 16 + rt * 8 + g;
          const int q = q0 + wc * 32 + ct * 8 + 2 * t;
          if (m < ma) {
            if (q < nb_) C[(size_t)m * ldc + q] = mk(cre[rt][ct][0], cim[rt][ct][0]);
            if (q + 1 < nb_) C[(size_t)m * ldc + q + 1] = mk(cre[rt][ct][1], cim[rt][ct][1]);
            if (herm && q0 > m0) {                      // the mirror tile
              if (q < nb_) C[(size_t)q * ldc + m] = mk(cre[rt][ct][0], -cim[rt][ct][0]);
              if (q + 1 < nb_) C[(size_t)(q + 1) * ldc + m] = mk(cre[rt][ct][1], -cim[rt][ct][1]);
            }
          }
        }
    }
  }
  __syncthreads();
}

// overlap of the occupied blocks of two mesh points: M[m][q] = sum_o conj(A[m][o]) B[q][o]
__device__ __forceinline__ void cta_overlap_dmma(const WfView& v, const cplx* __restrict__ pa, const cplx* __restrict__ pb,
                                                 cplx* __restrict__ M, int ld, OvSmem& sm) {
  const GemmSide A{pa, v.ss, 1, v.occ, 1}, B{pb, v.ss, 1, v.occ, 0};
  cta_gemm_dmma(A, v.nocc, B, v.nocc, v.n, nullptr, M, ld, sm);
}

// det(M)/|det(M)| of the row-major n x n matrix M (destroyed) by LU with partial pivoting, all threads of
// the CTA: parallel pivot search, one warp per row in the rank-1 update (coalesced along the row).
// sred/sidx: [32] shared scratch each; fbuf: [n] shared multipliers.
__device__ cplx lu_det_phase_cta(cplx* __restrict__ M, int n, int ld, double* sred, int* sidx, cplx* fbuf) {
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  cplx u = mk(1.0, 0.0);
  for (int k = 0; k < n; ++k) {
    double best = -1.0;
    int bi = n;
    for (int r = k + tid; r < n; r += T) {
      const double x = norm2(M[(size_t)r * ld + k]);
      if (x > best) { best = x; bi = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { sred[warp] = best; sidx[warp] = bi; }
    __syncthreads();
    best = sred[0]; bi = sidx[0];
    for (int w = 1; w < nw; ++w)
      if (sred[w] > best || (sred[w] == best && sidx[w] < bi)) { best = sred[w]; bi = sidx[w]; }
    if (!(best > 0.0)) return mk(0.0, 0.0);
    const int piv = bi;
    if (piv != k) {
      for (int c = k + tid; c < n; c += T) {
        const cplx tmp = M[(size_t)k * ld + c];
        M[(size_t)k * ld + c] = M[(size_t)piv * ld + c];
        M[(size_t)piv * ld + c] = tmp;
      }
      u = -u;
    }
    __syncthreads();                              // swap done; sred/sidx free for the next step
    const cplx p = M[(size_t)k * ld + k];
    const double ap = sqrt(norm2(p));
    u = u * mk(p.re / ap, p.im / ap);
    for (int r = k + 1 + tid; r < n; r += T) fbuf[r] = cdiv(M[(size_t)r * ld + k], p);
    __syncthreads();
    // rank-1 update of the trailing block, one warp per pair of rows.  M lives in global memory (L2): written as a plain
    // `row[c] -= f * rowk[c]` loop, a warp had ONE load in flight and paid an L2 round trip per 32 elements — at n = 200
    // that was 7 of the 10 ms of a link determinant (profiles/README.md r11).  Here a lane keeps its four pivot-row
    // elements of a 128-column chunk in registers and issues the eight loads of two rows before it uses any of them.
    const cplx* rowk = M + (size_t)k * ld;
    for (int cb = k + 1; cb < n; cb += 128) {
      cplx rk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int c = cb + lane + 32 * u; rk[u] = c < n ? rowk[c] : mk(0.0, 0.0); }
      for (int r = k + 1 + 2 * warp; r < n; r += 2 * nw) {
        const bool two = r + 1 < n;
        cplx* row0 = M + (size_t)r * ld;
        cplx* row1 = M + (size_t)(two ? r + 1 : r) * ld;
        const cplx f0 = fbuf[r], f1 = fbuf[two ? r + 1 : r];
        cplx v0[4], v1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = cb + lane + 32 * u;
          if (c < n) { v0[u] = row0[c]; v1[u] = row1[c]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = cb + lane + 32 * u;
          if (c < n) {
            row0[c] = v0[u] - f0 * rk[u];
            if (two) row1[c] = v1[u] - f1 * rk[u];
          }
        }
      }
    }
    __syncthreads();
  }
  const double nu = sqrt(norm2(u));
  return mk(u.re / nu, u.im / nu);
}

constexpr int kWilsonBig = 8;     // from this nocc on the Wilson-loop branch runs on CTA-wide GEMMs

// sum over the CTA (every thread gets it); red: [32] shared doubles
__device__ __forceinline__ double block_sum(double x, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = x;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  __syncthreads();
  return t;
}

__device__ __forceinline__ double block_max(double x, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = x;
  __syncthreads();
  double t = red[0];
  for (int i = 1; i < nw; ++i) t = fmax(t, red[i]);
  __syncthreads();
  return t;
}

// mode 0: out[l] = det/|det| of the overlap.   mode 1: out[l*nocc*nocc ...] = unitary polar factor.
__global__ void __launch_bounds__(256, TBK_GEMM_MINB)
link_matrix_kernel(WfView v, LinkMap map, long long nlinks, int mode, cplx* __restrict__ out, cplx* __restrict__ gws) {
  extern __shared__ __align__(16) char dyn_smem[];
  __shared__ double red[32];
  __shared__ int ired[32];
  OvSmem& ov = *reinterpret_cast<OvSmem*>(dyn_smem);
  cplx* fbuf = reinterpret_cast<cplx*>(dyn_smem + sizeof(OvSmem));
  const int nocc = v.nocc;
  const int ld = nocc | 1;
  // per-CTA workspace: M [nocc*ld] (+ 2 work matrices for the polar factor)
  cplx* M = gws + (size_t)blockIdx.x * (size_t)nocc * ld * 3;
  for (long long l = blockIdx.x; l < nlinks; l += gridDim.x) {
    long long oa, ob;
    link_endpoints(map, l, oa, ob);
    const cplx* pa = v.wfs + oa;
    const cplx* pb = v.wfs + ob;
    cta_overlap_dmma(v, pa, pb, M, ld, ov);
    if (mode == 0) {
      const cplx u = lu_det_phase_cta(M, nocc, ld, red, ired, fbuf);
      if (threadIdx.x == 0) out[l] = u;
      __syncthreads();
    } else {
      if (nocc >= kWilsonBig) {
        // unitary polar factor by Newton-Schulz, X <- X (3 I - X^H X) / 2, both products on the DMMA GEMM.
        // The iteration converges for |X|_2 < sqrt(3).  The overlap of two orthonormal sets has singular values
        // in (0, 1]; whatever else is in the array (user-filled, not normalised) is first scaled by the
        // spectral-norm bound sqrt(|X|_1 |X|_inf) >= |X|_2 whenever that bound exceeds 1.7, so the iteration cannot
        // diverge.  Singular directions of an overlap (sigma = 0: symmetry-enforced orthogonality between the two
        // occupied sets, e.g. a surface band crossing the chosen filling at a high-symmetry k) stay zero: the result
        // is the partial isometry sum_{sigma > 0} u v^H, where the reference's SVD puts an arbitrary unitary
        // completion of the null spaces.  Only a non-finite iterate (NaN / Inf in the array) is poisoned with NaN.
        cplx* X = M;
        cplx* G = M + (size_t)nocc * ld;
        cplx* Xn = G + (size_t)nocc * ld;
        double rmax = 0.0, cmax = 0.0;
        for (int r = threadIdx.x; r < nocc; r += blockDim.x) {
          double rs = 0.0, cs = 0.0;
          for (int c = 0; c < nocc; ++c) {
            rs += sqrt(norm2(X[(size_t)r * ld + c]));
            cs += sqrt(norm2(X[(size_t)c * ld + r]));
          }
          rmax = fmax(rmax, rs); cmax = fmax(cmax, cs);
        }
        rmax = block_max(rmax, red);
        cmax = block_max(cmax, red);
        const double bound = sqrt(rmax * cmax);
        if (bound > 1.7) {
          const double sc = 1.0 / bound;
          for (int idx = threadIdx.x; idx < nocc * nocc; idx += blockDim.x) {
            cplx& x = X[(size_t)(idx / nocc) * ld + idx % nocc];
            x = sc * x;
          }
          __syncthreads();
        }
        bool converged = false;
        for (int it = 0; it < 100; ++it) {
          const GemmSide A1{X, 1, ld, nullptr, 1}, B1{X, 1, ld, nullptr, 0};
          cta_gemm_dmma(A1, nocc, B1, nocc, nocc, nullptr, G, ld, ov, true);    // G = X^H X (Hermitian: half the tiles)
          double dev = 0.0;
          for (int idx = threadIdx.x; idx < nocc * nocc; idx += blockDim.x) {
            const int r = idx / nocc, c = idx - r * nocc;
            cplx gg = G[(size_t)r * ld + c];
            if (r == c) gg.re -= 1.0;
            dev += norm2(gg);
            G[(size_t)r * ld + c] = mk((r == c ? 1.0 : 0.0) - 0.5 * gg.re, -0.5 * gg.im);   // 1.5 I - 0.5 G
          }
          dev = block_sum(dev, red);
          if (!(dev == dev) || dev > 1.0e300) break;                              // NaN / Inf in the input
          converged = true;                                                       // finite: X is the best iterate so far
          if (!(dev > 1.0e-28 * nocc)) break;                                     // |X^H X - I|_F <= 1e-14 sqrt(nocc)
          // quadratic convergence: with delta = |X^H X - I|_F the next iterate has ~(3/4) delta^2, so once
          // delta^2 = dev <= 1e-14 sqrt(nocc) the update below is the last one and needs no check product
          const bool last = !(dev > 1.0e-14 * sqrt((double)nocc));
          const GemmSide A2{X, ld, 1, nullptr, 0}, B2{G, 1, ld, nullptr, 0};
          cta_gemm_dmma(A2, nocc, B2, nocc, nocc, nullptr, Xn, ld, ov);          // Xn = X P
          for (int idx = threadIdx.x; idx < nocc * nocc; idx += blockDim.x) {
            const size_t at = (size_t)(idx / nocc) * ld + idx % nocc;
            X[at] = Xn[at];
          }
          __syncthreads();
          if (last) break;
        }
        if (!converged) {
          for (int idx = threadIdx.x; idx < nocc * nocc; idx += blockDim.x) X[(size_t)(idx / nocc) * ld + idx % nocc] = mk(NAN, NAN);
          __syncthreads();
        }
        cplx* dst = out + (size_t)l * nocc * nocc;
        for (int idx = threadIdx.x; idx < nocc * nocc; idx += blockDim.x) dst[idx] = X[(size_t)(idx / nocc) * ld + idx % nocc];
        __syncthreads();
        continue;
      }
      // small nocc: serial scaled Newton on a compact copy (thread 0)
      if (threadIdx.x == 0) {
        cplx* X = M + (size_t)nocc * ld;          // compact nocc x nocc
        cplx* w1 = X + (size_t)nocc * nocc;
        cplx* w2 = M;                             // reuse (M is consumed into X first)
        for (int r = 0; r < nocc; ++r)
          for (int c = 0; c < nocc; ++c) X[(size_t)r * nocc + c] = M[(size_t)r * ld + c];
        polar_unitary(X, nocc, w1, w2);
        cplx* dst = out + (size_t)l * nocc * nocc;
        for (int i = 0; i < nocc * nocc; ++i) dst[i] = X[i];
      }
      __syncthreads();
    }
  }
}

// small-nocc version of the link kernel: one link per thread
template <int NOCC>
__global__ void __launch_bounds__(128)
link_small_kernel(WfView v, LinkMap map, long long nlinks, int mode, cplx* __restrict__ out) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= nlinks) return;
  long long oa, ob;
  link_endpoints(map, l, oa, ob);
  const cplx* pa = v.wfs + oa;
  const cplx* pb = v.wfs + ob;
  if (mode == 0) {
    out[l] = unit(link_det_small<NOCC>(v, pa, pb));
    return;
  }
  cplx M[NOCC * NOCC], w1[NOCC * NOCC], w2[NOCC * NOCC];
  for (int m = 0; m < NOCC; ++m)
    for (int q = 0; q < NOCC; ++q) {
      cplx acc = mk(0.0, 0.0);
      const cplx* ra = pa + (long long)v.occ[m] * v.ss;
      const cplx* rb = pb + (long long)v.occ[q] * v.ss;
      for (int o = 0; o < v.n; ++o) fma_acc_conj(acc, ra[o], rb[o]);
      M[m * NOCC + q] = acc;
    }
  polar_unitary(M, NOCC, w1, w2);
  cplx* dst = out + (size_t)l * NOCC * NOCC;
  for (int i = 0; i < NOCC * NOCC; ++i) dst[i] = M[i];
}

// plaquette phases from link determinants (general nocc)
__global__ void __launch_bounds__(256)
plaq_from_links_kernel(const cplx* __restrict__ dets, long long nslice, long long n0, long long n1,
                       double* __restrict__ plaq, double* __restrict__ partial) {
  const long long p0 = n0 - 1, p1 = n1 - 1;
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long i = blockIdx.y, s = blockIdx.z;
  double phase = 0.0;
  if (j < p1 && i < p0) {
    const long long per = (n0 - 1) * n1 + n0 * (n1 - 1);
    const cplx* dx = dets + s * per;                 // dx[i*n1 + j]  : (i,j)->(i+1,j)
    const cplx* dy = dx + (n0 - 1) * n1;             // dy[i*(n1-1)+j]: (i,j)->(i,j+1)
    cplx prod = dx[i * n1 + j];
    prod = prod * dy[(i + 1) * (n1 - 1) + j];
    prod = mulc(prod, dx[i * n1 + j + 1]);           // reversed link: conjugate determinant
    prod = mulc(prod, dy[i * (n1 - 1) + j]);
    phase = neg_arg(prod);
    if (plaq) plaq[(s * p0 + i) * p1 + j] = phase;
  }
  if (partial) {
    __shared__ double red[8];
    double x = phase;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      partial[(s * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
  }
}

// string phase = -arg prod_t det_t : one warp per string, fixed-order tree product
__global__ void __launch_bounds__(128)
string_phase_kernel(const cplx* __restrict__ dets, long long nstr, long long nlink, double* __restrict__ out) {
  const long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nstr) return;
  cplx p = mk(1.0, 0.0);
  for (long long t = lane; t < nlink; t += 32) {
    p = p * dets[s * nlink + t];
    if (((t >> 5) & 63) == 63) p = unit(p);      // keep the modulus at 1 on long strings
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cplx q;
    q.re = __shfl_down_sync(0xffffffffu, p.re, o);
    q.im = __shfl_down_sync(0xffffffffu, p.im, o);
    p = p * q;
  }
  if (lane == 0) out[s] = neg_arg(p);
}

// Wilson loop spectrum: ordered product of the unitary link matrices of one string,
// eigenvalues by complex QR, phases sorted ascending.  One thread per string.
__global__ void __launch_bounds__(64)
string_wilson_kernel(const cplx* __restrict__ umats, long long nstr, long long nlink, int nocc,
                     cplx* __restrict__ gws, double* __restrict__ out, cplx* __restrict__ prod_out) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= nstr) return;
  const size_t nn = (size_t)nocc * nocc;
  cplx* P = gws + (size_t)s * (2 * nn + nocc);
  cplx* T = P + nn;
  cplx* ev = T + nn;
  for (int r = 0; r < nocc; ++r)
    for (int c = 0; c < nocc; ++c) P[(size_t)r * nocc + c] = mk(r == c ? 1.0 : 0.0, 0.0);
  for (long long t = 0; t < nlink; ++t) {
    matmul_nn(P, umats + (size_t)(s * nlink + t) * nn, T, nocc);     // prd = prd @ U_t  (pythtb.py:3826)
    for (size_t i = 0; i < nn; ++i) P[i] = T[i];
  }
  if (prod_out) {                                // only the ordered product is wanted (sharded strings)
    for (size_t i = 0; i < nn; ++i) prod_out[(size_t)s * nn + i] = P[i];
    return;
  }
  comqr_eigvals(P, nocc, nocc, ev);
  double* o = out + s * nocc;
  for (int i = 0; i < nocc; ++i) o[i] = neg_arg(ev[i]);
  for (int i = 1; i < nocc; ++i) {              // insertion sort (np.sort, :3837)
    const double x = o[i];
    int j = i - 1;
    while (j >= 0 && o[j] > x) { o[j + 1] = o[j]; --j; }
    o[j + 1] = x;
  }
}

// ---- Wilson-loop spectra for nocc >= kWilsonBig -------------------------------------------------
// (a) ordered product of the link matrices of every string as a binary tree: one launch per level,
//     U[i] <- U[i] U[i + stride] for i = 0, 2 stride, 4 stride, ... (one CTA per product, DMMA GEMM);
//     after ceil(log2 nlink) levels U[0] holds prod_t U[t]  (pythtb.py:3826, same left-to-right order).
__global__ void __launch_bounds__(256, TBK_GEMM_MINB)
string_product_kernel(cplx* __restrict__ umats, long long nstr, long long nlink, long long stride, int nocc,
                      cplx* __restrict__ tmp) {
  __shared__ OvSmem sm;
  const size_t nn = (size_t)nocc * nocc;
  const long long np = (nlink - stride + 2 * stride - 1) / (2 * stride);
  cplx* T = tmp + (size_t)blockIdx.x * nn;
  for (long long wk = blockIdx.x; wk < nstr * np; wk += gridDim.x) {
    const long long s = wk / np, j = wk - s * np;
    cplx* A = umats + (size_t)(s * nlink + 2 * stride * j) * nn;
    const cplx* B = A + (size_t)stride * nn;
    const GemmSide ga{A, nocc, 1, nullptr, 0}, gb{B, 1, nocc, nullptr, 0};
    cta_gemm_dmma(ga, nocc, gb, nocc, nocc, nullptr, T, nocc, sm);
    for (size_t i = threadIdx.x; i < nn; i += blockDim.x) A[i] = T[i];
    __syncthreads();
  }
}

// (b) eigenphases of the unitary W = U[0] of every string through the Hermitian solver: W is normal, so
//     H(phi) = cos(phi) (W + W^H)/2 + sin(phi) (W - W^H)/(2i) has the eigenvectors of W (eigenvalue
//     cos(theta - phi)); the phases are read off the Rayleigh quotients v^H W v.  Two generic angles are
//     run and, per string, the one whose quotients are closer to the unit circle is kept: an accidental
//     collision cos(theta_1 - phi) = cos(theta_2 - phi) with theta_1 != theta_2 mixes eigenvectors for one
//     angle only.
__global__ void __launch_bounds__(256)
unitary_herm_kernel(const cplx* __restrict__ umats, long long nstr, long long nlink, int nocc, double cphi, double sphi,
                    cplx* __restrict__ H) {
  const size_t nn = (size_t)nocc * nocc;
  const long long total = nstr * (long long)nn;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long s = q / (long long)nn;
    const int r = (int)((q - s * (long long)nn) / nocc), c = (int)(q - s * (long long)nn - (long long)r * nocc);
    const cplx* W = umats + (size_t)s * nlink * nn;
    const cplx a = W[(size_t)r * nocc + c], b = conj(W[(size_t)c * nocc + r]);
    const cplx hs = mk(0.5 * (a.re + b.re), 0.5 * (a.im + b.im));             // (W + W^H) / 2
    const cplx ha = mk(0.5 * (a.im - b.im), -0.5 * (a.re - b.re));            // (W - W^H) / (2i)
    H[q] = mk(cphi * hs.re + sphi * ha.re, cphi * hs.im + sphi * ha.im);
  }
}

// one warp per (string, band): rho = v^H W v with v the eigenvector ROW vec[s][b][:] (H v^T = lambda v^T)
__global__ void __launch_bounds__(256)
unitary_rayleigh_kernel(const cplx* __restrict__ umats, const cplx* __restrict__ vec, long long nstr, long long nlink, int nocc,
                        double* __restrict__ phase, double* __restrict__ modulus) {
  const size_t nn = (size_t)nocc * nocc;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long wk = warp0; wk < nstr * nocc; wk += nwarp) {
    const long long s = wk / nocc;
    const int b = (int)(wk - s * nocc);
    const cplx* W = umats + (size_t)s * nlink * nn;
    const cplx* v = vec + ((size_t)s * nocc + b) * nocc;
    cplx acc = mk(0.0, 0.0);
    for (int i = 0; i < nocc; ++i) {
      cplx row = mk(0.0, 0.0);
      for (int j = lane; j < nocc; j += 32) fma_acc(row, W[(size_t)i * nocc + j], v[j]);
      fma_acc_conj(acc, v[i], row);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o);
      acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o);
    }
    if (lane == 0) {
      phase[wk] = neg_arg(acc);
      modulus[wk] = sqrt(norm2(acc));
    }
  }
}

// per string: pick the better of the two runs, sort its phases ascending (np.sort, pythtb.py:3837)
__global__ void __launch_bounds__(64)
unitary_select_kernel(const double* __restrict__ ph0, const double* __restrict__ mod0, const double* __restrict__ ph1,
                      const double* __restrict__ mod1, long long nstr, int nocc, double* __restrict__ out) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= nstr) return;
  double w0 = 0.0, w1 = 0.0;
  for (int b = 0; b < nocc; ++b) {
    w0 = fmax(w0, fabs(mod0[s * nocc + b] - 1.0));
    w1 = fmax(w1, fabs(mod1[s * nocc + b] - 1.0));
  }
  const double* src = (w0 <= w1 ? ph0 : ph1) + s * nocc;
  double* o = out + s * nocc;
  for (int i = 0; i < nocc; ++i) o[i] = src[i];
  for (int i = 1; i < nocc; ++i) {
    const double x = o[i];
    int j = i - 1;
    while (j >= 0 && o[j] > x) { o[j + 1] = o[j]; --j; }
    o[j + 1] = x;
  }
}

// last slice = first slice (* phase[o]) on [outer][len][inner][nsta_arr][n]
__global__ void __launch_bounds__(256)
impose_boundary_kernel(cplx* __restrict__ wfs, long long outer, long long len, long long inner, int nsta_arr, int n,
                       const cplx* __restrict__ phase) {
  const long long blk = (long long)nsta_arr * n;
  const long long per_outer = inner * blk;
  const long long total = outer * per_outer;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long o = q / per_outer, r = q - o * per_outer;
    const int orb = (int)(r % n);
    cplx val = wfs[o * len * per_outer + r];
    if (phase) val = val * phase[orb];
    wfs[(o * len + (len - 1)) * per_outer + r] = val;
  }
}

// dst[q] = src[q] (* phase[q % n]): packs the first local row of a shard for the ring shift
__global__ void __launch_bounds__(256)
halo_pack_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst, long long total, int n, const cplx* __restrict__ phase) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    cplx v = src[q];
    if (phase) v = v * phase[(int)(q % n)];
    dst[q] = v;
  }
}

// X[k][m][q] = sum_o conj(A[k][m][o]) pos[o] A[k][q][o]
__global__ void __launch_bounds__(256)
position_matrix_kernel(const cplx* __restrict__ evec, long long batch, int nocc, int n, const double* __restrict__ pos,
                       cplx* __restrict__ xmat) {
  const long long total = batch * nocc * nocc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long k = idx / ((long long)nocc * nocc);
    const int r = (int)(idx - k * nocc * nocc);
    const int m = r / nocc, q = r - m * nocc;
    const cplx* a = evec + (k * nocc + m) * (long long)n;
    const cplx* b = evec + (k * nocc + q) * (long long)n;
    cplx acc = mk(0.0, 0.0);
    for (int o = 0; o < n; ++o) fma_acc_conj(acc, a[o], pos[o] * b[o]);
    xmat[idx] = acc;
  }
}

// DMMA versions for nocc >= 16: one CTA per k-point
__global__ void __launch_bounds__(256, TBK_GEMM_MINB)
position_matrix_dmma_kernel(const cplx* __restrict__ evec, long long batch, int nocc, int n, const double* __restrict__ pos,
                            cplx* __restrict__ xmat) {
  __shared__ OvSmem sm;
  for (long long k = blockIdx.x; k < batch; k += gridDim.x) {
    const cplx* e = evec + k * (long long)nocc * n;
    const GemmSide A{e, n, 1, nullptr, 1}, B{e, n, 1, nullptr, 0};
    cta_gemm_dmma(A, nocc, B, nocc, n, pos, xmat + k * (long long)nocc * nocc, nocc, sm, true);   // <u| r |u> is Hermitian
  }
}

__global__ void __launch_bounds__(256, TBK_GEMM_MINB)
hwf_to_orbital_dmma_kernel(const cplx* __restrict__ hwf, const cplx* __restrict__ evec, long long batch, int nocc, int n,
                           cplx* __restrict__ out) {
  __shared__ OvSmem sm;
  for (long long k = blockIdx.x; k < batch; k += gridDim.x) {
    // out[i][o] = sum_m hwf[i][m] evec[m][o]: B(q = o, x = m) = evec[m][o] -> row stride 1, contraction stride n
    const GemmSide A{hwf + k * (long long)nocc * nocc, nocc, 1, nullptr, 0};
    const GemmSide B{evec + k * (long long)nocc * n, 1, n, nullptr, 0};
    cta_gemm_dmma(A, nocc, B, n, nocc, nullptr, out + k * (long long)nocc * n, n, sm);
  }
}

// out[k][i][o] = sum_m hwf[k][i][m] * evec[k][m][o]      (pythtb.py:2262-2274)
__global__ void __launch_bounds__(256)
hwf_to_orbital_kernel(const cplx* __restrict__ hwf, const cplx* __restrict__ evec, long long batch, int nocc, int n,
                      cplx* __restrict__ out) {
  const long long total = batch * nocc * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long k = idx / ((long long)nocc * n);
    const int r = (int)(idx - k * nocc * n);
    const int i = r / n, o = r - i * n;
    const cplx* h = hwf + (k * nocc + i) * (long long)nocc;
    const cplx* e = evec + k * (long long)nocc * n + o;
    cplx acc = mk(0.0, 0.0);
    for (int m = 0; m < nocc; ++m) fma_acc(acc, h[m], e[(long long)m * n]);
    out[idx] = acc;
  }
}

static long long link_ws_elems(int nocc) { return (long long)nocc * (nocc | 1) * 3; }
static int product_grid(long long nstr, long long nlink) {
  const long long work = nstr * ((nlink + 1) / 2), cap = (long long)kNumSM * 4;
  return (int)(work < cap ? (work > 0 ? work : 1) : cap);
}
static int link_grid(long long nlinks) {
  const long long cap = (long long)kNumSM * 4;
  return (int)(nlinks < cap ? (nlinks > 0 ? nlinks : 1) : cap);
}

// prod[s] = umats[s][0]  (the root of the product tree)
__global__ void __launch_bounds__(256)
string_root_copy_kernel(const cplx* __restrict__ umats, long long nstr, long long nlink, int nocc, cplx* __restrict__ prod) {
  const size_t nn = (size_t)nocc * nocc;
  const long long total = nstr * (long long)nn;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long s = q / (long long)nn;
    prod[q] = umats[(size_t)s * nlink * nn + (size_t)(q - s * (long long)nn)];
  }
}

static size_t align256_(size_t x) { return (x + 255) & ~(size_t)255; }
size_t wilson_tail_bytes(int nocc, long long nstr, long long nlink);

// ordered product of the nlink unitary matrices of every string (umats [nstr][nlink][nocc][nocc], destroyed)
// and then either its eigenphases (out, sorted ascending) or the product itself (prod_out); ws as sized by
// wilson_tail_bytes.
static int wilson_tail(cplx* umats, long long nstr, long long nlink, int nocc, double* out, cplx* prod_out, char* ws,
                       cudaStream_t st) {
  const size_t nn = (size_t)nocc * nocc;
  if (nocc < kWilsonBig) {
    cplx* sws = (cplx*)ws;
    string_wilson_kernel<<<(unsigned)((nstr + 63) / 64), 64, 0, st>>>(umats, nstr, nlink, nocc, sws, out, prod_out);
    TBK_LAUNCH_CHECK("string_wilson_kernel");
    return TBK_OK;
  }
  // large nocc: tree product, Hermitian solver on two generic combinations, Rayleigh quotients
  const int pgrid = product_grid(nstr, nlink);
  cplx* ptmp = (cplx*)ws;        ws += align256_((size_t)pgrid * nn * 16);
  for (long long stride = 1; stride < nlink; stride *= 2) {
    string_product_kernel<<<pgrid, 256, 0, st>>>(umats, nstr, nlink, stride, nocc, ptmp);
    TBK_LAUNCH_CHECK("string_product_kernel");
  }
  long long eb = (nstr * (long long)nn + 255) / 256;
  if (eb > kNumSM * 16) eb = kNumSM * 16;
  if (prod_out) {
    string_root_copy_kernel<<<(unsigned)eb, 256, 0, st>>>(umats, nstr, nlink, nocc, prod_out);
    TBK_LAUNCH_CHECK("string_root_copy_kernel");
    return TBK_OK;
  }
  cplx* H = (cplx*)ws;           ws += align256_((size_t)nstr * nn * 16);
  cplx* vecs = (cplx*)ws;        ws += align256_((size_t)nstr * nn * 16);
  double* evals = (double*)ws;   ws += align256_((size_t)nstr * nocc * 8);
  double* ph[2]; double* md[2];
  for (int a = 0; a < 2; ++a) {
    ph[a] = (double*)ws; ws += align256_((size_t)nstr * nocc * 8);
    md[a] = (double*)ws; ws += align256_((size_t)nstr * nocc * 8);
  }
  const size_t ews = tbk_eigh_workspace(nocc, nstr, 1);
  const double phis[2] = {0.7390851332151607, 2.0287578381104342};
  long long rb = (nstr * nocc * 32 + 255) / 256;
  if (rb > kNumSM * 16) rb = kNumSM * 16;
  for (int a = 0; a < 2; ++a) {
    unitary_herm_kernel<<<(unsigned)eb, 256, 0, st>>>(umats, nstr, nlink, nocc, cos(phis[a]), sin(phis[a]), H);
    TBK_LAUNCH_CHECK("unitary_herm_kernel");
    const int rc = tbk_eigh_batched((const double*)H, nocc, nstr, evals, (double*)vecs, ws, ews, (void*)st);
    if (rc) return rc;
    unitary_rayleigh_kernel<<<(unsigned)rb, 256, 0, st>>>(umats, vecs, nstr, nlink, nocc, ph[a], md[a]);
    TBK_LAUNCH_CHECK("unitary_rayleigh_kernel");
  }
  unitary_select_kernel<<<(unsigned)((nstr + 63) / 64), 64, 0, st>>>(ph[0], md[0], ph[1], md[1], nstr, nocc, out);
  TBK_LAUNCH_CHECK("unitary_select_kernel");
  return TBK_OK;
}

size_t wilson_tail_bytes(int nocc, long long nstr, long long nlink) {
  const size_t nn = (size_t)nocc * nocc;
  if (nocc < kWilsonBig) return align256_((size_t)nstr * (2 * nn + nocc) * 16);
  size_t bytes = align256_((size_t)product_grid(nstr, nlink) * nn * 16);
  bytes += 2 * align256_((size_t)nstr * nn * 16);
  bytes += 5 * align256_((size_t)nstr * nocc * 8);
  bytes += tbk_eigh_workspace(nocc, nstr, 1) + 256;
  return bytes;
}

static int launch_links(const WfView& v, const LinkMap& map, long long nlinks, int mode, cplx* out, cplx* gws, cudaStream_t st) {
  if (nlinks <= 0) return TBK_OK;
  if (v.nocc <= 4) {
    const unsigned blocks = (unsigned)((nlinks + 127) / 128);
    switch (v.nocc) {
      case 1: link_small_kernel<1><<<blocks, 128, 0, st>>>(v, map, nlinks, mode, out); break;
      case 2: link_small_kernel<2><<<blocks, 128, 0, st>>>(v, map, nlinks, mode, out); break;
      case 3: link_small_kernel<3><<<blocks, 128, 0, st>>>(v, map, nlinks, mode, out); break;
      default: link_small_kernel<4><<<blocks, 128, 0, st>>>(v, map, nlinks, mode, out); break;
    }
    TBK_LAUNCH_CHECK("link_small_kernel");
    return TBK_OK;
  }
  const size_t dyn = sizeof(OvSmem) + (size_t)v.nocc * 16 + 64;
  TBK_CUDA(cudaFuncSetAttribute(link_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  link_matrix_kernel<<<link_grid(nlinks), 256, dyn, st>>>(v, map, nlinks, mode, out, gws);
  TBK_LAUNCH_CHECK("link_matrix_kernel");
  return TBK_OK;
}

}  // namespace tbk

using namespace tbk;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int tbk_impose_boundary(double* wfs_dev, int64_t outer, int64_t len, int64_t inner, int32_t nsta_arr, int32_t n,
                        const double* phase_dev, void* stream) {
  TBK_NVTX("tbk_impose_boundary");
  if (!wfs_dev || outer < 1 || len < 2 || inner < 1 || nsta_arr < 1 || n < 1) { set_error("tbk_impose_boundary: bad argument"); return TBK_ERR_ARG; }
  const long long total = outer * inner * nsta_arr * n;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSM * 16) blocks = kNumSM * 16;
  impose_boundary_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((cplx*)wfs_dev, outer, len, inner, nsta_arr, n, (const cplx*)phase_dev);
  TBK_LAUNCH_CHECK("impose_boundary_kernel");
  return TBK_OK;
}

int tbk_halo_pack(const double* row_dev, double* dst_dev, int64_t npoints, int32_t nsta_arr, int32_t n,
                  const double* phase_dev, void* stream) {
  TBK_NVTX("tbk_halo_pack");
  if (!row_dev || !dst_dev || npoints < 0 || nsta_arr < 1 || n < 1) { set_error("tbk_halo_pack: bad argument"); return TBK_ERR_ARG; }
  const long long total = npoints * nsta_arr * n;
  if (total == 0) return TBK_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSM * 16) blocks = kNumSM * 16;
  halo_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)row_dev, (cplx*)dst_dev, total, n, (const cplx*)phase_dev);
  TBK_LAUNCH_CHECK("halo_pack_kernel");
  return TBK_OK;
}

size_t tbk_flux_workspace(int32_t nocc, int32_t n, int64_t nslice, int64_t n0, int64_t n1) {
  (void)n;
  const long long bx = (n1 - 1 + 255) / 256;
  size_t bytes = align256((size_t)(nslice * (n0 - 1) * (bx > 0 ? bx : 1)) * 8);   // block partial sums
  const long long tb1 = flux_tiles_bound(nslice, n0, n1), tb2 = ring_tiles_bound(nslice, n0, n1);
  bytes += align256((size_t)(tb1 > tb2 ? tb1 : tb2) * 8);
  if (nocc > 4) {
    const long long nlinks = nslice * ((n0 - 1) * n1 + n0 * (n1 - 1));
    bytes += align256((size_t)nlinks * 16);
    bytes += align256((size_t)link_grid(nlinks) * link_ws_elems(nocc) * 16);
  }
  return bytes + 256;
}

int tbk_flux_plane(const tbk_wf_view* view, const int64_t* slice_off_dev, int64_t nslice, int64_t n0, int64_t stride0,
                   int64_t n1, int64_t stride1, double* plaq_dev, double* total_dev, void* ws_dev, size_t ws_bytes,
                   void* stream) {
  return tbk_flux_plane_x(view, slice_off_dev, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, ws_dev, ws_bytes,
                          nullptr, stream);
}

int tbk_flux_plane_x(const tbk_wf_view* view, const int64_t* slice_off_dev, int64_t nslice, int64_t n0, int64_t stride0,
                     int64_t n1, int64_t stride1, double* plaq_dev, double* total_dev, void* ws_dev, size_t ws_bytes,
                     tbk_peer* peer, void* stream) {
  TBK_NVTX("tbk_flux_plane_x");
  if (!view || !view->wfs_dev || !view->occ_dev || !slice_off_dev || nslice < 1 || n0 < 2 || n1 < 2 || view->nocc < 1 ||
      (!plaq_dev && !total_dev)) {
    set_error("tbk_flux_plane: bad argument");
    return TBK_ERR_ARG;
  }
  if (ws_bytes < tbk_flux_workspace(view->nocc, view->n, nslice, n0, n1) || !ws_dev) {
    set_error("tbk_flux_plane: workspace too small");
    return TBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  WfView v{(const cplx*)view->wfs_dev, view->n, view->nsta_arr, view->nocc, view->occ_dev,
           view->state_stride > 0 ? (long long)view->state_stride : (long long)view->n};
  const long long p0 = n0 - 1, p1 = n1 - 1;
  const long long bx = (p1 + 255) / 256;
  if (p0 > 65535 || nslice > 65535) { set_error("tbk_flux_plane: mesh too large for one launch"); return TBK_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)bx, (unsigned)p0, (unsigned)nslice);
  char* ws = (char*)ws_dev;
  double* partial = total_dev ? (double*)ws : nullptr;
  ws += align256((size_t)(nslice * p0 * bx) * 8);
  const long long* off = (const long long*)slice_off_dev;
  const bool rows_kernel = view->nocc <= 2 && view->n >= 2 && view->n <= 4 && view->nocc <= view->n;
  if (peer && peer->connected && peer->nranks > 1 && total_dev && !(rows_kernel && nslice <= kPeerMaxVals)) {
    set_error("tbk_flux_plane_x: the fused cross-rank sum needs nocc <= 2, n <= 4 and at most %d slices", kPeerMaxVals);
    return TBK_ERR_UNSUPPORTED;
  }
  // contiguous mesh columns: the bulk-copy ring kernel is available behind TBK_FLUX_RING=1.  Measured on B200
  // (profiles/r01/README.md) it streams at the same ~5 TB/s as the register-prefetch kernel while the SMs are
  // active and pays one halo row per tile, so on the 1024 x 1024 mesh it is 1-2 us slower: not the default.
  const char* ring_env = getenv("TBK_FLUX_RING");
  const bool ring_on = ring_env && atoi(ring_env) == 1;
  if (rows_kernel && ring_on && v.ss == view->n && stride1 == (long long)view->nsta_arr * view->n && n1 >= 64 &&
      ((uintptr_t)view->wfs_dev & 15) == 0) {
    double* part2 = (double*)ws;
    int rc = TBK_OK;
    const int key = view->nocc * 10 + view->n;
    switch (key) {
      case 12: rc = launch_flux_ring<1, 2>(v, off, nslice, n0, stride0, n1, plaq_dev, total_dev, part2, peer, st); break;
      case 22: rc = launch_flux_ring<2, 2>(v, off, nslice, n0, stride0, n1, plaq_dev, total_dev, part2, peer, st); break;
      case 13: rc = launch_flux_ring<1, 3>(v, off, nslice, n0, stride0, n1, plaq_dev, total_dev, part2, peer, st); break;
      case 23: rc = launch_flux_ring<2, 3>(v, off, nslice, n0, stride0, n1, plaq_dev, total_dev, part2, peer, st); break;
      case 14: rc = launch_flux_ring<1, 4>(v, off, nslice, n0, stride0, n1, plaq_dev, total_dev, part2, peer, st); break;
      default: rc = launch_flux_ring<2, 4>(v, off, nslice, n0, stride0, n1, plaq_dev, total_dev, part2, peer, st); break;
    }
    return rc;
  }
  if (rows_kernel) {
    double* part2 = (double*)ws;
    int rc = TBK_OK;
    const int key = view->nocc * 10 + view->n;
    switch (key) {
      case 12: {
        // prefetch depth of the one-band / two-orbital case (TBK_FLUX_SLOTS = 4 | 6 | 8; A/B knob, see profiles/README.md)
        static int slots = -1;
        if (slots < 0) { const char* e = getenv("TBK_FLUX_SLOTS"); slots = e ? atoi(e) : kFluxDefaultSlots12; }
        const bool deep_ok = v.ss != view->n;              // deeper rotations pay only when a row is one band (state-major)
        if (slots == 8 && deep_ok) rc = launch_flux_rows<1, 2, 8>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st);
        else if (slots == 6 && deep_ok) rc = launch_flux_rows<1, 2, 6>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st);
        else rc = launch_flux_rows<1, 2>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st);
        break;
      }
      case 22: rc = launch_flux_rows<2, 2>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st); break;
      case 13: rc = launch_flux_rows<1, 3>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st); break;
      case 23: rc = launch_flux_rows<2, 3>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st); break;
      case 14: rc = launch_flux_rows<1, 4>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st); break;
      default: rc = launch_flux_rows<2, 4>(v, off, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, part2, peer, st); break;
    }
    return rc;
  }
  if (view->nocc <= 4) {
    switch (view->nocc) {
      case 1: flux_small_kernel<1><<<grid, 256, 0, st>>>(v, off, n0, stride0, n1, stride1, plaq_dev, partial); break;
      case 2: flux_small_kernel<2><<<grid, 256, 0, st>>>(v, off, n0, stride0, n1, stride1, plaq_dev, partial); break;
      case 3: flux_small_kernel<3><<<grid, 256, 0, st>>>(v, off, n0, stride0, n1, stride1, plaq_dev, partial); break;
      default: flux_small_kernel<4><<<grid, 256, 0, st>>>(v, off, n0, stride0, n1, stride1, plaq_dev, partial); break;
    }
    TBK_LAUNCH_CHECK("flux_small_kernel");
  } else {
    LinkMap map{off, n0, n1, stride0, stride1, 1};
    const long long nlinks = nslice * links_per_slice(map);
    cplx* dets = (cplx*)ws;
    ws += align256((size_t)nlinks * 16);
    cplx* gws = (cplx*)ws;
    int rc = launch_links(v, map, nlinks, 0, dets, gws, st);
    if (rc) return rc;
    plaq_from_links_kernel<<<grid, 256, 0, st>>>(dets, nslice, n0, n1, plaq_dev, partial);
    TBK_LAUNCH_CHECK("plaq_from_links_kernel");
  }
  if (total_dev) {
    reduce_partials_kernel<<<(unsigned)nslice, 256, 0, st>>>(partial, p0 * bx, total_dev);
    TBK_LAUNCH_CHECK("reduce_partials_kernel");
  }
  return TBK_OK;
}

size_t tbk_berry_workspace(int32_t nocc, int32_t n, int64_t nstr, int64_t npts, int32_t berry_evals) {
  (void)n;
  const long long nlinks = nstr * (npts - 1);
  size_t bytes = 256;
  if (!berry_evals) bytes += align256((size_t)nlinks * 16);
  else {
    bytes += align256((size_t)nlinks * nocc * nocc * 16);
    bytes += wilson_tail_bytes(nocc, nstr, npts - 1);
  }
  if (nocc > 4) bytes += align256((size_t)link_grid(nlinks) * link_ws_elems(nocc) * 16);
  return bytes;
}

int tbk_berry_strings(const tbk_wf_view* view, const int64_t* string_off_dev, int64_t nstr, int64_t npts, int64_t stride,
                      int32_t berry_evals, double* out_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_berry_strings");
  if (!view || !view->wfs_dev || !view->occ_dev || !string_off_dev || !out_dev || nstr < 1 || npts < 2 || view->nocc < 1) {
    set_error("tbk_berry_strings: bad argument");
    return TBK_ERR_ARG;
  }
  if (!ws_dev || ws_bytes < tbk_berry_workspace(view->nocc, view->n, nstr, npts, berry_evals)) {
    set_error("tbk_berry_strings: workspace too small");
    return TBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  WfView v{(const cplx*)view->wfs_dev, view->n, view->nsta_arr, view->nocc, view->occ_dev,
           view->state_stride > 0 ? (long long)view->state_stride : (long long)view->n};
  LinkMap map{(const long long*)string_off_dev, npts, 1, stride, 0, 0};
  const long long nlink = npts - 1, nlinks = nstr * nlink;
  char* ws = (char*)ws_dev;
  if (!berry_evals) {
    cplx* dets = (cplx*)ws;
    ws += align256((size_t)nlinks * 16);
    int rc = launch_links(v, map, nlinks, 0, dets, (cplx*)ws, st);
    if (rc) return rc;
    const long long threads = nstr * 32;
    string_phase_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(dets, nstr, nlink, out_dev);
    TBK_LAUNCH_CHECK("string_phase_kernel");
    return TBK_OK;
  }
  const size_t nn = (size_t)view->nocc * view->nocc;
  cplx* umats = (cplx*)ws;
  ws += align256((size_t)nlinks * nn * 16);
  cplx* lws = (cplx*)ws;
  if (view->nocc > 4) ws += align256((size_t)link_grid(nlinks) * link_ws_elems(view->nocc) * 16);
  int rc = launch_links(v, map, nlinks, 1, umats, lws, st);
  if (rc) return rc;
  return wilson_tail(umats, nstr, nlink, view->nocc, out_dev, nullptr, ws, st);
}

int tbk_wilson_products(const tbk_wf_view* view, const int64_t* string_off_dev, int64_t nstr, int64_t npts, int64_t stride,
                        double* prod_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_wilson_products");
  if (!view || !view->wfs_dev || !view->occ_dev || !string_off_dev || !prod_dev || nstr < 1 || npts < 2 || view->nocc < 1) {
    set_error("tbk_wilson_products: bad argument");
    return TBK_ERR_ARG;
  }
  if (!ws_dev || ws_bytes < tbk_berry_workspace(view->nocc, view->n, nstr, npts, 1)) {
    set_error("tbk_wilson_products: workspace too small");
    return TBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  WfView v{(const cplx*)view->wfs_dev, view->n, view->nsta_arr, view->nocc, view->occ_dev,
           view->state_stride > 0 ? (long long)view->state_stride : (long long)view->n};
  LinkMap map{(const long long*)string_off_dev, npts, 1, stride, 0, 0};
  const long long nlink = npts - 1, nlinks = nstr * nlink;
  const size_t nn = (size_t)view->nocc * view->nocc;
  char* ws = (char*)ws_dev;
  cplx* umats = (cplx*)ws;
  ws += align256((size_t)nlinks * nn * 16);
  cplx* lws = (cplx*)ws;
  if (view->nocc > 4) ws += align256((size_t)link_grid(nlinks) * link_ws_elems(view->nocc) * 16);
  int rc = launch_links(v, map, nlinks, 1, umats, lws, st);
  if (rc) return rc;
  return wilson_tail(umats, nstr, nlink, view->nocc, nullptr, (cplx*)prod_dev, ws, st);
}

size_t tbk_wilson_workspace(int32_t nocc, int64_t nstr, int64_t nmat) { return wilson_tail_bytes(nocc, nstr, nmat) + 256; }

int tbk_wilson_phases(double* mats_dev, int64_t nstr, int64_t nmat, int32_t nocc, double* out_dev, void* ws_dev,
                      size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_wilson_phases");
  if (!mats_dev || !out_dev || nstr < 1 || nmat < 1 || nocc < 1) { set_error("tbk_wilson_phases: bad argument"); return TBK_ERR_ARG; }
  if (!ws_dev || ws_bytes < tbk_wilson_workspace(nocc, nstr, nmat)) { set_error("tbk_wilson_phases: workspace too small"); return TBK_ERR_WORKSPACE; }
  return wilson_tail((cplx*)mats_dev, nstr, nmat, nocc, out_dev, nullptr, (char*)ws_dev, (cudaStream_t)stream);
}

int tbk_wilson_chain(double* mats_dev, int64_t nstr, int64_t nmat, int32_t nocc, double* prod_dev, void* ws_dev,
                     size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_wilson_chain");
  if (!mats_dev || !prod_dev || nstr < 1 || nmat < 1 || nocc < 1) { set_error("tbk_wilson_chain: bad argument"); return TBK_ERR_ARG; }
  if (!ws_dev || ws_bytes < tbk_wilson_workspace(nocc, nstr, nmat)) { set_error("tbk_wilson_chain: workspace too small"); return TBK_ERR_WORKSPACE; }
  return wilson_tail((cplx*)mats_dev, nstr, nmat, nocc, nullptr, (cplx*)prod_dev, (char*)ws_dev, (cudaStream_t)stream);
}

int tbk_position_matrix(const double* evec_dev, int64_t batch, int32_t nocc, int32_t n, const double* pos_dev,
                        double* xmat_dev, void* stream) {
  TBK_NVTX("tbk_position_matrix");
  if (!evec_dev || !pos_dev || !xmat_dev || batch < 0 || nocc < 1 || n < 1) { set_error("tbk_position_matrix: bad argument"); return TBK_ERR_ARG; }
  if (batch == 0) return TBK_OK;
  if (nocc >= 16) {
    const long long blocks = batch < (long long)kNumSM * 4 ? batch : (long long)kNumSM * 4;
    position_matrix_dmma_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)evec_dev, batch, nocc, n, pos_dev, (cplx*)xmat_dev);
    TBK_LAUNCH_CHECK("position_matrix_dmma_kernel");
    return TBK_OK;
  }
  const long long total = batch * nocc * nocc;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSM * 16) blocks = kNumSM * 16;
  position_matrix_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)evec_dev, batch, nocc, n, pos_dev, (cplx*)xmat_dev);
  TBK_LAUNCH_CHECK("position_matrix_kernel");
  return TBK_OK;
}

size_t tbk_position_hwf_workspace(int32_t nocc, int32_t n, int64_t batch) {
  (void)n;
  return align256((size_t)batch * nocc * nocc * 16) * 2 + tbk_eigh_workspace(nocc, batch, 1) + 512;
}

int tbk_position_hwf(const double* evec_dev, int64_t batch, int32_t nocc, int32_t n, const double* pos_dev,
                     double* hwfc_dev, double* hwf_dev, int32_t orbital_basis, void* ws_dev, size_t ws_bytes, void* stream) {
  TBK_NVTX("tbk_position_hwf");
  if (!evec_dev || !pos_dev || !hwfc_dev || batch < 0 || nocc < 1 || n < 1) { set_error("tbk_position_hwf: bad argument"); return TBK_ERR_ARG; }
  if (!ws_dev || ws_bytes < tbk_position_hwf_workspace(nocc, n, batch)) { set_error("tbk_position_hwf: workspace too small"); return TBK_ERR_WORKSPACE; }
  if (batch == 0) return TBK_OK;
  char* ws = (char*)ws_dev;
  double* xmat = (double*)ws;
  ws += align256((size_t)batch * nocc * nocc * 16);
  double* vecs = (double*)ws;
  ws += align256((size_t)batch * nocc * nocc * 16);
  int rc = tbk_position_matrix(evec_dev, batch, nocc, n, pos_dev, xmat, stream);
  if (rc) return rc;
  const bool direct = hwf_dev && !orbital_basis;
  rc = tbk_eigh_batched(xmat, nocc, batch, hwfc_dev, hwf_dev ? (direct ? hwf_dev : vecs) : nullptr, ws,
                        ws_bytes - (size_t)(ws - (char*)ws_dev), stream);
  if (rc) return rc;
  if (hwf_dev && orbital_basis && nocc >= 16) {
    const long long blocks = batch < (long long)kNumSM * 4 ? batch : (long long)kNumSM * 4;
    hwf_to_orbital_dmma_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)vecs, (const cplx*)evec_dev, batch, nocc, n, (cplx*)hwf_dev);
    TBK_LAUNCH_CHECK("hwf_to_orbital_dmma_kernel");
  } else if (hwf_dev && orbital_basis) {
    const long long total = batch * nocc * n;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSM * 16) blocks = kNumSM * 16;
    hwf_to_orbital_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)vecs, (const cplx*)evec_dev, batch, nocc, n, (cplx*)hwf_dev);
    TBK_LAUNCH_CHECK("hwf_to_orbital_kernel");
  }
  return TBK_OK;
}

}  // extern "C"
