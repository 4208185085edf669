// tbk_internal.cuh — shared by the .cu translation units of libtbk_b200.so:
// error plumbing, the device-resident model, thread-group types for the SPMD
// eigensolver / LU code, and small launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <functional>
#include <stdio.h>
#include <string.h>
#include "../../include/tbk.h"
#include "tbk_common.cuh"
#include "tbk_plan.cuh"
#include "tbk_peer.cuh"

namespace tbk {

void set_error(const char* fmt, ...);
void count_launch();                    // every kernel this library launches is counted (tbk_launch_count)
void note_kernel(const char* name);     // remembered for tbk_last_kernel()
constexpr int kTicketSlots = 1024;
unsigned* take_ticket();                // zero-initialised, self-resetting device counter (see tbk_api.cu)
int cuda_fail(cudaError_t e, const char* what);

#define TBK_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

#define TBK_LAUNCH_CHECK(name)                               \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return cuda_fail(e__, name);     \
    tbk::count_launch();                                     \
  } while (0)

// Optional per-CTA timeline of the two headline kernels (TBK_CTA_TRACE=1, profiling only): a device
// buffer of (smid, begin ns, end ns, blockIdx) per CTA of the last traced launch; tbk_debug_cta_trace
// copies it out.  nullptr (one uniform predicate in the kernel) unless enabled.
constexpr int kCtaTraceCap = 4096;
unsigned long long* cta_trace_buffer();
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long cta_trace_begin(const unsigned long long* trace) {
  return (trace && threadIdx.x == 0) ? global_ns() : 0ull;
}
__device__ __forceinline__ void cta_trace_end(unsigned long long* trace, unsigned long long t0) {
  if (trace && threadIdx.x == 0 && blockIdx.x < kCtaTraceCap) {
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    unsigned long long* p = trace + 4 * (size_t)blockIdx.x;
    p[0] = sm; p[1] = t0; p[2] = global_ns(); p[3] = blockIdx.x;
  }
}
#endif

// Completion signal for a synchronous prepared call: the kernel that writes the call's (pinned-host) results
// also writes `seq` to a pinned host word after them (system-scope fence in between), and the host spins on
// that word instead of on cudaStreamSynchronize — the results are usable ~1 us after the last store instead of
// after the stream's completion semaphore has made its way through the driver.  tbk_prepared_run posts a
// request (thread-local); a launcher whose kernel supports the signal takes it.
struct DoneSignal {
  unsigned long long* flag;           // nullptr: no signal
  unsigned long long seq;
};
void post_done_request(const DoneSignal& d);
DoneSignal take_done_request();        // returns the pending request (flag == nullptr if none) and clears it
bool done_request_pending();
#if defined(__CUDACC__)
// by ONE thread, after the results it (or, behind a __syncthreads, its CTA) wrote
__device__ __forceinline__ void signal_done(const DoneSignal& d) {
  if (d.flag) {
    __threadfence_system();
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(d.flag), "l"(d.seq) : "memory");
  }
}
#endif

// NVTX range around a C-ABI entry point (header-only NVTX3: a no-op unless a profiler injects itself)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define TBK_NVTX(name) tbk::NvtxRange nvtx_range__(name)

constexpr int kNumSM = 148;           // B200
constexpr int kMaxSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

// ---------------------------------------------------------------- thread groups
#if defined(__CUDACC__)
template <int G>
struct TileGroup {
  int t;
  unsigned mask;
  __device__ TileGroup() {
    const int lane = threadIdx.x & 31;
    t = lane % G;
    mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - t));
  }
  __device__ int tid() const { return t; }
  __device__ int size() const { return G; }
  __device__ void sync() { __syncwarp(mask); }
  __device__ double sum(double x) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o, G);
    return x;
  }
};

struct BlockGroup {
  double* red;  // [32] shared scratch
  __device__ explicit BlockGroup(double* r) : red(r) {}
  __device__ int tid() const { return threadIdx.x; }
  __device__ int size() const { return blockDim.x; }
  __device__ void sync() { __syncthreads(); }
  // sub-teams = warps (tbk_eig_blocked.cuh)
  __device__ int nsub() const { return blockDim.x >> 5; }
  __device__ int sub() const { return threadIdx.x >> 5; }
  __device__ int lane() const { return threadIdx.x & 31; }
  __device__ int subsize() const { return 32; }
  __device__ void subsync() { __syncwarp(); }
  __device__ double subsum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  }
  __device__ double sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = x;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += red[i];
    __syncthreads();
    return s;
  }
};

struct ThreadGroup {  // a "group" of one thread (serial code paths)
  __device__ int tid() const { return 0; }
  __device__ int size() const { return 1; }
  __device__ void sync() {}
  __device__ double sum(double x) { return x; }
};
#endif

}  // namespace tbk

namespace tbk {
// Dense coefficient form of a small model (nsta <= 4, nph <= kDenseMaxPh), passed BY VALUE as a
// kernel parameter so that every coefficient is a constant-bank operand of the FP64 pipe:
//   H_e(k) = C_e + sum_p ( P_ep cos(2 pi k.R_p) + Q_ep sin(2 pi k.R_p) ),  e = lower-triangle element
// with P = A + B, Q = i (A - B) where A (B) collects the amplitudes multiplying E_p (conj E_p).
constexpr int kDenseMaxPh = 8;
constexpr int kDenseMaxEl = 10;   // 4*5/2
struct DenseSmall {
  int valid;                       // 0: model does not fit this form
  int nph;
  unsigned mask[kDenseMaxPh];      // bit e: (P,Q)[p][e] != 0
  double C[kDenseMaxEl][2];
  double P[kDenseMaxPh][kDenseMaxEl][2];
  double Q[kDenseMaxPh][kDenseMaxEl][2];
  double R[kDenseMaxPh][TBK_MAX_DIM];
  double tau[4][TBK_MAX_DIM];
};
}  // namespace tbk

// The opaque model handle of tbk.h
struct tbk_model {
  tbk::PlanView pv;   // device pointers
  tbk::DenseSmall dense;
  void* blob;         // single device allocation backing every array
  size_t blob_bytes;
  int device;
  int max_terms_per_phase;
};

// The opaque prepared call of tbk.h: the bound arguments of one entry point
struct tbk_prepared {
  std::function<int(void*)> run;     // re-issues the call on the given stream
  unsigned long long* done_flag;     // pinned host word of the completion signal (allocated on first synchronous run)
  unsigned long long done_seq;
};

// The opaque peer group of tbk.h: this rank's mailbox and the IPC mappings of the others
struct tbk_peer {
  int rank, nranks, device;
  unsigned long long epoch;          // advanced by every collective issued through this group
  double* box[tbk::kPeerMaxRanks];   // box[rank] is the local allocation
  bool connected;
  bool defer_next;                   // tbk_peer_defer: the next *_x collective is only POSTED by its kernel
  int npending;                      // posted, not yet completed collectives (oldest first)
  tbk::PeerPending pending[2 * tbk::kPeerMaxPend];
};

namespace tbk {
// completes every pending collective with one-CTA kernels (tbk_api.cu); no-op when nothing is pending
int peer_flush(tbk_peer* p, cudaStream_t st);

inline PeerView peer_none() {
  PeerView v;
  memset(&v, 0, sizeof(v));
  return v;
}
inline bool peer_active(const tbk_peer* p) { return p && p->connected && p->nranks > 1; }

// moves the oldest `count` pending collectives into the view
inline void peer_take_pending(tbk_peer* p, PeerView& v, int count) {
  for (int i = 0; i < count; ++i) v.pend[v.npend++] = p->pending[i];
  for (int i = count; i < p->npending; ++i) p->pending[i - count] = p->pending[i];
  p->npending -= count;
}

// PeerView for the next collective of nv values (advances the epoch).
//   completer: a kernel that may complete older deferred collectives (the flux kernels, the flush / barrier
//   kernels); the grid-solve kernel only posts when deferred, so that nothing but a few stores sits between its
//   last CTA and the dependent flux kernel.
// A synchronous collective (defer_next not set) completes everything pending together with its own result; a
// deferred one is queued on the handle and completes only the queue entries that are >= 2 epochs old (the previous
// step's).  rc != 0: launching a flush failed.
inline int peer_next(tbk_peer* p, int nv, int op, double* out, bool completer, cudaStream_t st, PeerView* view) {
  PeerView v = peer_none();
  if (!peer_active(p)) { if (p) p->defer_next = false; *view = v; return 0; }
  const bool defer = p->defer_next;
  p->defer_next = false;
  // slot-reuse safety / queue capacity: never let a posted epoch lag more than kPeerMaxLag, never overflow the view
  if (p->npending > 0 && (p->epoch + 1 - p->pending[0].epoch > (unsigned long long)kPeerMaxLag ||
                          (!defer && p->npending > kPeerMaxPend) || p->npending >= 2 * kPeerMaxPend - 1)) {
    if (int rc = peer_flush(p, st)) return rc;
  }
  v.rank = p->rank; v.nranks = p->nranks; v.epoch = ++p->epoch;
  for (int r = 0; r < p->nranks; ++r) v.box[r] = p->box[r];
  v.complete_self = defer ? 0 : 1;
  if (!defer) {
    peer_take_pending(p, v, p->npending);                     // <= kPeerMaxPend by the check above
  } else {
    if (completer) {
      int old = 0;
      while (old < p->npending && old < kPeerMaxPend && p->pending[old].epoch + 2 <= v.epoch) ++old;
      peer_take_pending(p, v, old);
    }
    PeerPending& q = p->pending[p->npending++];
    q.epoch = v.epoch; q.nv = nv; q.op = op; q.out = out;
  }
  *view = v;
  return 0;
}
}  // namespace tbk
