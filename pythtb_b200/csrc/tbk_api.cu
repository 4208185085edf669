// tbk_api.cu — error plumbing, model upload, misc entry points of libtbk_b200.so.
#include <stdarg.h>
#include <atomic>
#include <mutex>
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

__global__ void flush_l2_kernel(double4* __restrict__ buf, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = make_double4(v, v, v, v);
}

}  // namespace tbk

using namespace tbk;

extern "C" {

int tbk_version(void) { return 100; }

const char* tbk_last_error(void) { return g_err; }

int tbk_model_create(const tbk_model_desc* d, tbk_model** out) {
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
  *out = m;
  return TBK_OK;
}

int tbk_model_destroy(tbk_model* m) {
  if (!m) return TBK_OK;
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
  cudaError_t e = cudaMalloc(&mem, kPeerMailboxBytes);
  if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(mailbox)"); }
  e = cudaMemset(mem, 0, kPeerMailboxBytes);
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
