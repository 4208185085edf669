// tbk_api.cu — error plumbing, model upload, misc entry points of libtbk_b200.so.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <functional>
#include <vector>
#include "tbk_internal.cuh"

namespace tbk {

static thread_local char g_err[512] = "";
static thread_local const char* g_last_kernel = "";

static std::atomic<long long> g_launches{0};

void note_kernel(const char* name) { g_last_kernel = name; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-device array of self-resetting tickets for "last CTA finishes the reduction" kernels.
// One-time 4 KB allocation per device (zeroed once); a call takes the next slot round-robin.  A
// ticket is incremented with atomicInc(slot, nblocks-1), which wraps it back to 0 by itself.
static unsigned* g_tickets[64] = {nullptr};
static unsigned g_next_ticket[64] = {0};
static std::mutex g_ticket_mu;

unsigned* take_ticket() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_ticket_mu);
  if (!g_tickets[dev]) {
    if (cudaMalloc(&g_tickets[dev], kTicketSlots * sizeof(unsigned)) != cudaSuccess) return nullptr;
    if (cudaMemset(g_tickets[dev], 0, kTicketSlots * sizeof(unsigned)) != cudaSuccess) return nullptr;
  }
  const unsigned slot = g_next_ticket[dev]++ % kTicketSlots;
  return g_tickets[dev] + slot;
}

static thread_local DoneSignal g_done_req = {nullptr, 0};
void post_done_request(const DoneSignal& d) { g_done_req = d; }
DoneSignal take_done_request() { const DoneSignal d = g_done_req; g_done_req.flag = nullptr; return d; }
bool done_request_pending() { return g_done_req.flag != nullptr; }

unsigned long long* cta_trace_buffer() {
  static int on = -1;
  static unsigned long long* buf = nullptr;
  if (on < 0) { const char* e = getenv("TBK_CTA_TRACE"); on = (e && atoi(e) == 1) ? 1 : 0; }
  if (!on) return nullptr;
  if (!buf) {
    const size_t bytes = (size_t)2 * kCtaTraceCap * 4 * sizeof(unsigned long long);   // [0, cap): solve kernel, [cap, 2 cap): flux kernel
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return nullptr;
    cudaMemset(buf, 0, bytes);
  }
  return buf;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return TBK_ERR_CUDA;
}

// device-side barrier over the peer group (one warp): a synchronous all-reduce of a dummy value
__global__ void peer_barrier_kernel(const __grid_constant__ PeerView pv, double* scratch) {
  __shared__ double s_v[1];
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_v[0] = 1.0;
  __syncthreads();
  peer_collective(pv, s_v, 1, 0, scratch, &s_fail);
}

// completes deferred collectives nobody else picked up (one CTA, posts nothing)
__global__ void peer_collect_kernel(const __grid_constant__ PeerView pv) {
  __shared__ int s_fail;
  peer_prologue(pv);
  __syncthreads();
  peer_collective(pv, nullptr, 0, 0, nullptr, &s_fail);
}

int peer_flush(tbk_peer* p, cudaStream_t st) {
  if (!peer_active(p)) return TBK_OK;
  while (p->nqueue > 0) {
    PeerView v = peer_none();
    peer_view_base(p, v);
    v.epoch = 0;
    peer_attach_posts(p, v);                  // its first (only) CTA posts what nobody posted yet ...
    peer_attach_pends(p, v, true);            // ... and completes everything that is posted, up to the view's capacity
    peer_age_posts(p);
    peer_collect_kernel<<<1, 64, 0, st>>>(v);
    TBK_LAUNCH_CHECK("peer_collect_kernel");
  }
  return TBK_OK;
}

__global__ void flush_l2_kernel(double4* __restrict__ buf, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = make_double4(v, v, v, v);
}

// ---- FP64 peak microkernels: the denominators of the FP64 rooflines are MEASURED in the run that reports them
// (MEASURED_PEAKS.json has no FP64 figure).  kind 0: 16 independent DFMA chains per thread (FMA pipe);
// kind 1: 8 independent mma.sync.m8n8k4.f64 accumulator pairs per warp (the DMMA path of the overlap GEMMs).
constexpr int kPeakThreads = 256, kPeakCtasPerSm = 4;
__global__ void __launch_bounds__(kPeakThreads)
fp64_peak_dfma_kernel(int iters, double seed, double* __restrict__ sink) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + 1.0e-3 * (double)(threadIdx.x + i);
  const double m = 1.0 - 1.0e-9, c = 1.0e-9 * seed;
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += a[i];
  if (t == 123.456) sink[0] = t;                  // never true: keeps the chains alive
}
__global__ void __launch_bounds__(kPeakThreads)
fp64_peak_dmma_kernel(int iters, double seed, double* __restrict__ sink) {
  double d0[8], d1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { d0[i] = 0.0; d1[i] = 0.0; }
  const double a = 1.0e-3 * seed + 1.0e-6 * (double)(threadIdx.x & 31), b = 1.0e-3 - 1.0e-6 * (double)(threadIdx.x & 31);
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(d0[i]), "+d"(d1[i]) : "d"(a), "d"(b));
  }
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += d0[i] + d1[i];
  if (t == 123.456) sink[0] = t;
}

}  // namespace tbk

using namespace tbk;

extern "C" {

int tbk_version(void) { return 100; }

const char* tbk_last_error(void) { return g_err; }

int tbk_model_create(const tbk_model_desc* d, tbk_model** out) {
  TBK_NVTX("tbk_model_create");
  if (!d || !out) { set_error("tbk_model_create: null argument"); return TBK_ERR_ARG; }
  if (d->dim_k < 0 || d->dim_k > TBK_MAX_DIM || d->nsta < 1 || d->nph < 0 || d->nel < 0 || d->nterm < 0 ||
      (d->convention != 1 && d->convention != 2)) {
    set_error("tbk_model_create: bad descriptor (dim_k=%d nsta=%d nph=%d nel=%d nterm=%d convention=%d)",
              d->dim_k, d->nsta, d->nph, d->nel, d->nterm, d->convention);
    return TBK_ERR_ARG;
  }
  const int dk = d->dim_k > 0 ? d->dim_k : 1;
  const int nph = d->nph > 0 ? d->nph : 1, nel = d->nel > 0 ? d->nel : 1, nterm = d->nterm > 0 ? d->nterm : 1;
  struct Seg { const void* src; size_t bytes; size_t off; };
  Seg segs[10] = {
      {d->ph_R, (size_t)nph * dk * 8, 0},        {d->tau, (size_t)d->nsta * dk * 8, 0},
      {d->el_ptr, (size_t)(d->nel + 1) * 4, 0},  {d->el_row, (size_t)nel * 4, 0},
      {d->el_col, (size_t)nel * 4, 0},           {d->t_ph, (size_t)nterm * 4, 0},
      {d->t_amp, (size_t)nterm * 16, 0},         {d->pm_ptr, (size_t)(d->nph + 2) * 4, 0},
      {d->pm_el, (size_t)nterm * 4, 0},          {d->pm_amp, (size_t)nterm * 16, 0}};
  size_t total = 0;
  for (auto& s : segs) {
    if (!s.src) { set_error("tbk_model_create: null array in descriptor"); return TBK_ERR_ARG; }
    s.off = total;
    total += (s.bytes + 255) & ~(size_t)255;
  }
  std::vector<char> host(total, 0);
  for (auto& s : segs) memcpy(host.data() + s.off, s.src, s.bytes);
  // sanity: CSR monotone and indices in range (cheap, protects the kernels)
  for (int e = 0; e < d->nel; ++e) {
    if (d->el_ptr[e] > d->el_ptr[e + 1] || d->el_row[e] < d->el_col[e] || d->el_row[e] >= d->nsta || d->el_col[e] < 0) {
      set_error("tbk_model_create: malformed element table at %d", e);
      return TBK_ERR_ARG;
    }
  }
  int maxpp = 0;
  for (int p = 0; p <= d->nph; ++p) {
    const int c = d->pm_ptr[p + 1] - d->pm_ptr[p];
    if (c < 0) { set_error("tbk_model_create: malformed phase table"); return TBK_ERR_ARG; }
    if (c > maxpp) maxpp = c;
  }
  if (d->nel > 0 && d->el_ptr[d->nel] != d->nterm) { set_error("tbk_model_create: el_ptr/nterm mismatch"); return TBK_ERR_ARG; }
  tbk_model* m = new tbk_model();
  TBK_CUDA(cudaGetDevice(&m->device));
  cudaError_t e = cudaMalloc(&m->blob, total);
  if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaMalloc(model)"); }
  e = cudaMemcpy(m->blob, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(m->blob); delete m; return cuda_fail(e, "cudaMemcpy(model)"); }
  m->blob_bytes = total;
  m->max_terms_per_phase = maxpp;
  char* b = (char*)m->blob;
  PlanView& pv = m->pv;
  pv.dim_k = d->dim_k; pv.nsta = d->nsta; pv.nph = d->nph; pv.nel = d->nel; pv.nterm = d->nterm;
  pv.convention = d->convention;
  pv.ph_R = (const double*)(b + segs[0].off);
  pv.tau = (const double*)(b + segs[1].off);
  pv.el_ptr = (const int*)(b + segs[2].off);
  pv.el_row = (const int*)(b + segs[3].off);
  pv.el_col = (const int*)(b + segs[4].off);
  pv.t_ph = (const int*)(b + segs[5].off);
  pv.t_amp = (const double*)(b + segs[6].off);
  pv.pm_ptr = (const int*)(b + segs[7].off);
  pv.pm_el = (const int*)(b + segs[8].off);
  pv.pm_amp = (const double*)(b + segs[9].off);
  // ---- dense coefficient form for the register-resident small-matrix mesh kernel
  DenseSmall& ds = m->dense;
  memset(&ds, 0, sizeof(ds));
  if (d->nsta >= 2 && d->nsta <= 4 && d->nph <= kDenseMaxPh && d->dim_k >= 1) {
    ds.valid = 1;
    ds.nph = d->nph;
    for (int p = 0; p < d->nph; ++p)
      for (int x = 0; x < d->dim_k; ++x) ds.R[p][x] = d->ph_R[p * d->dim_k + x];
    for (int o = 0; o < d->nsta; ++o)
      for (int x = 0; x < d->dim_k; ++x) ds.tau[o][x] = d->tau[o * d->dim_k + x];
    double A[kDenseMaxPh][kDenseMaxEl][2] = {}, B[kDenseMaxPh][kDenseMaxEl][2] = {};
    for (int p = 0; p <= d->nph; ++p) {
      for (int t = d->pm_ptr[p]; t < d->pm_ptr[p + 1]; ++t) {
        const int e = d->pm_el[t] & TBK_PH_MASK;
        const bool cj = (d->pm_el[t] & TBK_PH_CONJ) != 0;
        const int r = d->el_row[e], c = d->el_col[e];
        const int pk = r * (r + 1) / 2 + c;
        const double ar = d->pm_amp[2 * t], ai = d->pm_amp[2 * t + 1];
        if (p == d->nph) { ds.C[pk][0] += ar; ds.C[pk][1] += ai; }
        else if (!cj) { A[p][pk][0] += ar; A[p][pk][1] += ai; }
        else { B[p][pk][0] += ar; B[p][pk][1] += ai; }
      }
    }
    for (int p = 0; p < d->nph; ++p)
      for (int e = 0; e < kDenseMaxEl; ++e) {
        ds.P[p][e][0] = A[p][e][0] + B[p][e][0];
        ds.P[p][e][1] = A[p][e][1] + B[p][e][1];
        // Q = i (A - B)
        ds.Q[p][e][0] = -(A[p][e][1] - B[p][e][1]);
        ds.Q[p][e][1] = A[p][e][0] - B[p][e][0];
        if (ds.P[p][e][0] != 0.0 || ds.P[p][e][1] != 0.0 || ds.Q[p][e][0] != 0.0 || ds.Q[p][e][1] != 0.0)
          ds.mask[p] |= 1u << e;
      }
  }
  // ---- nsta = 5..8: dense REAL coefficient table for the tensor-pipe Hamiltonian assembly of the eigenvalue sweeps
  // (solve_reg_gemm_kernel):  [Re H_e; Im H_e](k) = A [cos 2 pi k.R_p; sin 2 pi k.R_p; 1], rows m < NP = Re of the packed
  // lower-triangle element m, rows NP + m = Im; columns 2p / 2p + 1 = cos / sin of phase p, column 2 nph = constant terms.
  // amp E = (ar c - ai s) + i (ar s + ai c),  amp conj(E) = (ar c + ai s) + i (ai c - ar s).
  // Stored in mma.sync.m8n8k4 A-fragment order: tab[(mt * KS + ks) * 32 + lane] = A[8 mt + lane / 4][4 ks + lane % 4].
  m->gemm_tab = nullptr; m->gemm_ks = 0;
  if (d->nsta >= 5 && d->nsta <= 8 && d->dim_k >= 1 && (2 * d->nph + 1 + 3) / 4 <= kGemmMaxKS) {
    const int NP = d->nsta * (d->nsta + 1) / 2, MT = (2 * NP + 7) / 8, KS = (2 * d->nph + 1 + 3) / 4, K = 4 * KS;
    std::vector<double> A((size_t)MT * 8 * K, 0.0);
    for (int p = 0; p <= d->nph; ++p) {
      for (int t = d->pm_ptr[p]; t < d->pm_ptr[p + 1]; ++t) {
        const int e = d->pm_el[t] & TBK_PH_MASK;
        const bool cj = (d->pm_el[t] & TBK_PH_CONJ) != 0;
        const int r = d->el_row[e], c = d->el_col[e];
        const int pk = r * (r + 1) / 2 + c;
        const double ar = d->pm_amp[2 * t], ai = d->pm_amp[2 * t + 1];
        double* re = A.data() + (size_t)pk * K;
        double* im = A.data() + (size_t)(NP + pk) * K;
        if (p == d->nph) { re[2 * d->nph] += ar; im[2 * d->nph] += ai; }
        else if (!cj) { re[2 * p] += ar; re[2 * p + 1] -= ai; im[2 * p] += ai; im[2 * p + 1] += ar; }
        else { re[2 * p] += ar; re[2 * p + 1] += ai; im[2 * p] += ai; im[2 * p + 1] -= ar; }
      }
    }
    std::vector<double> frag((size_t)MT * KS * 32);
    for (int mt = 0; mt < MT; ++mt)
      for (int ks = 0; ks < KS; ++ks)
        for (int lane = 0; lane < 32; ++lane)
          frag[((size_t)mt * KS + ks) * 32 + lane] = A[(size_t)(8 * mt + lane / 4) * K + 4 * ks + lane % 4];
    if (cudaMalloc(&m->gemm_tab, frag.size() * 8) == cudaSuccess &&
        cudaMemcpy(m->gemm_tab, frag.data(), frag.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess) {
      m->gemm_ks = KS;
    } else {
      cudaGetLastError();
      if (m->gemm_tab) cudaFree(m->gemm_tab);
      m->gemm_tab = nullptr;                 // not fatal: the scalar assembly is used instead
    }
  }
  *out = m;
  return TBK_OK;
}

int tbk_model_destroy(tbk_model* m) {
  if (!m) return TBK_OK;
  if (m->gemm_tab) cudaFree(m->gemm_tab);
  cudaFree(m->blob);
  delete m;
  return TBK_OK;
}

int tbk_peer_create(int32_t rank, int32_t nranks, tbk_peer** out, void* handle_out) {
  if (!out || !handle_out || nranks < 1 || nranks > kPeerMaxRanks || rank < 0 || rank >= nranks) {
    set_error("tbk_peer_create: bad argument (rank=%d nranks=%d, at most %d ranks)", rank, nranks, kPeerMaxRanks);
    return TBK_ERR_ARG;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  tbk_peer* p = new tbk_peer();
  memset(p, 0, sizeof(*p));
  p->rank = rank; p->nranks = nranks;
  TBK_CUDA(cudaGetDevice(&p->device));
  void* mem = nullptr;
  cudaError_t e = cudaMalloc(&mem, kPeerMailboxBytes + kPeerScratchBytes);
  if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(mailbox)"); }
  e = cudaMemset(mem, 0, kPeerMailboxBytes + kPeerScratchBytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, mem);
  if (e != cudaSuccess) { cudaFree(mem); delete p; return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle_out, &h, sizeof(h));
  p->box[rank] = (double*)mem;
  *out = p;
  return TBK_OK;
}

int tbk_peer_connect(tbk_peer* p, const void* handles) {
  if (!p || !handles) { set_error("tbk_peer_connect: null argument"); return TBK_ERR_ARG; }
  for (int r = 0; r < p->nranks; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* mem = nullptr;
    TBK_CUDA(cudaIpcOpenMemHandle(&mem, h, cudaIpcMemLazyEnablePeerAccess));
    p->box[r] = (double*)mem;
  }
  p->connected = true;
  return TBK_OK;
}

int tbk_peer_destroy(tbk_peer* p) {
  if (!p) return TBK_OK;
  for (int r = 0; r < p->nranks; ++r) {
    if (!p->box[r]) continue;
    if (r == p->rank) cudaFree(p->box[r]);
    else cudaIpcCloseMemHandle(p->box[r]);
  }
  delete p;
  return TBK_OK;
}

int tbk_peer_defer(tbk_peer* p, int32_t on) {
  if (p) p->defer_next = on != 0;
  return TBK_OK;
}

int tbk_peer_flush(tbk_peer* p, void* stream) { return peer_flush(p, (cudaStream_t)stream); }

int tbk_peer_barrier(tbk_peer* p, void* stream) {
  TBK_NVTX("tbk_peer_barrier");
  if (!peer_active(p)) return TBK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // a barrier is a synchronous collective of its own; it leaves deferred collectives alone (tbk_peer_flush
  // completes those) unless the slot-reuse rule forces them out first
  if (p->nqueue > 0 && p->epoch + 1 - p->queue[0].epoch > (unsigned long long)kPeerMaxLag) {
    if (int rc = peer_flush(p, st)) return rc;
  }
  PeerView v = peer_none();
  peer_view_base(p, v);
  v.epoch = ++p->epoch; v.mode = 1;
  peer_barrier_kernel<<<1, 32, 0, st>>>(v, (double*)((char*)p->box[p->rank] + kPeerMailboxBytes + kPeerLocalBytes));
  TBK_LAUNCH_CHECK("peer_barrier_kernel");
  return TBK_OK;
}

int tbk_debug_cta_trace(uint64_t* out, int64_t max_ctas, int32_t reset) {
  if (!out || max_ctas < 0) { set_error("tbk_debug_cta_trace: bad argument"); return TBK_ERR_ARG; }
  unsigned long long* buf = cta_trace_buffer();
  if (!buf) { set_error("tbk_debug_cta_trace: tracing is off (set TBK_CTA_TRACE=1 before the first launch)"); return TBK_ERR_UNSUPPORTED; }
  const size_t n = (size_t)(max_ctas < 2 * kCtaTraceCap ? max_ctas : 2 * kCtaTraceCap) * 4 * sizeof(unsigned long long);
  TBK_CUDA(cudaDeviceSynchronize());
  TBK_CUDA(cudaMemcpy(out, buf, n, cudaMemcpyDeviceToHost));
  if (reset) TBK_CUDA(cudaMemset(buf, 0, (size_t)2 * kCtaTraceCap * 4 * sizeof(unsigned long long)));
  return TBK_OK;
}

// ---- prepared calls: the arguments of a repeated solve_grid / flux_plane call, kept by value, so that
// re-issuing it costs the host one 3-argument call (a parameter sweep or a timing loop re-runs the same
// launch thousands of times and the kernels last ~20 us: the argument marshalling of the host language
// is then a visible share of the step).
int tbk_solve_grid_prepare(const tbk_model* m, const double* start_k, const int32_t* mesh, int32_t nd, int32_t row0,
                           int32_t nrows, int32_t wrap0, double* wfs_dev, const double* pbc_phase_dev, double* gaps_dev,
                           void* ws_dev, size_t ws_bytes, int64_t state_stride, tbk_peer* peer, tbk_prepared** out) {
  if (!m || !start_k || !mesh || !out || nd < 1 || nd > TBK_MAX_DIM) { set_error("tbk_solve_grid_prepare: bad argument"); return TBK_ERR_ARG; }
  std::vector<double> sk(start_k, start_k + nd);
  std::vector<int32_t> ms(mesh, mesh + nd);
  tbk_prepared* p = new tbk_prepared();
  p->done_flag = nullptr; p->done_seq = 0;
  p->run = [=](void* stream) {
    return tbk_solve_grid_x(m, sk.data(), ms.data(), nd, row0, nrows, wrap0, wfs_dev, pbc_phase_dev, gaps_dev, ws_dev,
                            ws_bytes, state_stride, peer, stream);
  };
  *out = p;
  return TBK_OK;
}

int tbk_flux_plane_prepare(const tbk_wf_view* view, const int64_t* slice_off_dev, int64_t nslice, int64_t n0,
                           int64_t stride0, int64_t n1, int64_t stride1, double* plaq_dev, double* total_dev,
                           void* ws_dev, size_t ws_bytes, tbk_peer* peer, tbk_prepared** out) {
  if (!view || !out) { set_error("tbk_flux_plane_prepare: bad argument"); return TBK_ERR_ARG; }
  const tbk_wf_view v = *view;
  tbk_prepared* p = new tbk_prepared();
  p->done_flag = nullptr; p->done_seq = 0;
  p->run = [=](void* stream) {
    return tbk_flux_plane_x(&v, slice_off_dev, nslice, n0, stride0, n1, stride1, plaq_dev, total_dev, ws_dev, ws_bytes,
                            peer, stream);
  };
  *out = p;
  return TBK_OK;
}

int tbk_prepared_run(tbk_prepared* p, void* stream, int32_t sync) {
  TBK_NVTX("tbk_prepared_run");
  if (!p || !p->run) { set_error("tbk_prepared_run: null handle"); return TBK_ERR_ARG; }
  static int spin = -1;                 // TBK_SPIN_SYNC=0: always wait with cudaStreamSynchronize (A/B knob)
  if (spin < 0) { const char* e = getenv("TBK_SPIN_SYNC"); spin = (e && atoi(e) == 0) ? 0 : 1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (sync && spin) {
    if (!p->done_flag) {
      if (cudaHostAlloc((void**)&p->done_flag, sizeof(unsigned long long), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        p->done_flag = nullptr;
      } else {
        *p->done_flag = 0;
      }
    }
    if (p->done_flag) post_done_request(DoneSignal{p->done_flag, ++p->done_seq});
  }
  const int rc = p->run(stream);
  const bool signalled = sync && spin && p->done_flag && !done_request_pending();   // a launcher took the request
  take_done_request();
  if (rc != TBK_OK) return rc;
  if (!sync) return TBK_OK;
  if (signalled) {
    // spin on the pinned word; every few thousand polls make sure the stream is still alive (a failed
    // kernel would never write it)
    volatile unsigned long long* f = p->done_flag;
    const unsigned long long want = p->done_seq;
    for (unsigned n = 1; *f != want; ++n) {
      if ((n & 0x3fff) == 0) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) break;                         // finished (the word is written by now or never)
        if (q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery");
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (*f == want) return TBK_OK;
  }
  TBK_CUDA(cudaStreamSynchronize(st));
  return TBK_OK;
}

int tbk_prepared_destroy(tbk_prepared* p) {
  if (p && p->done_flag) cudaFreeHost(p->done_flag);
  delete p;
  return TBK_OK;
}

double tbk_bench_fp64_flops(int32_t kind, int32_t iters) {
  const double threads = (double)kNumSM * kPeakCtasPerSm * kPeakThreads;
  if (kind == 0) return threads * (double)iters * 16.0 * 2.0;              // 16 FMA per thread per iteration
  return (threads / 32.0) * (double)iters * 8.0 * 512.0;                   // 8 m8n8k4 DMMA (2*8*8*4 flops) per warp per iteration
}

int tbk_bench_fp64(int32_t kind, int32_t iters, double* sink_dev, void* stream) {
  if (!sink_dev || iters < 1 || kind < 0 || kind > 1) { set_error("tbk_bench_fp64: bad argument"); return TBK_ERR_ARG; }
  const int grid = kNumSM * kPeakCtasPerSm;
  if (kind == 0) fp64_peak_dfma_kernel<<<grid, kPeakThreads, 0, (cudaStream_t)stream>>>(iters, 1.0, sink_dev);
  else fp64_peak_dmma_kernel<<<grid, kPeakThreads, 0, (cudaStream_t)stream>>>(iters, 1.0, sink_dev);
  TBK_LAUNCH_CHECK("fp64_peak_kernel");
  return TBK_OK;
}

const char* tbk_last_kernel(void) { return g_last_kernel; }

int64_t tbk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int tbk_stream_sync(void* stream) {
  TBK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return TBK_OK;
}

int tbk_flush_l2(void* buf_dev, size_t bytes, void* stream) {
  if (!buf_dev || bytes < 32) { set_error("tbk_flush_l2: bad buffer"); return TBK_ERR_ARG; }
  static double v = 0.0;
  v += 1.0;
  flush_l2_kernel<<<kNumSM * 8, 256, 0, (cudaStream_t)stream>>>((double4*)buf_dev, bytes / 32, v);
  TBK_LAUNCH_CHECK("flush_l2_kernel");
  return TBK_OK;
}

}  // extern "C"
