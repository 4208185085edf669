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
  // nsta = 5..8: dense real coefficient table of H(k) = A [cos; sin; 1] in DMMA A-fragment order (tbk_api.cu), or nullptr
  double* gemm_tab;
  int gemm_ks;        // K steps of four: 4 gemm_ks >= 2 nph + 1
};
namespace tbk {
constexpr int kGemmMaxKS = 53;   // 2 nph + 1 <= 212: the phase table of 128 k-points still fits one CTA's shared memory
}

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
  bool defer_next;                   // tbk_peer_defer: the next *_x collective is a deferred one
  int nqueue;                        // deferred collectives on their way (oldest first)
  tbk::PeerPending queue[tbk::kPeerQueue];
};

namespace tbk {
// posts and completes every deferred collective with one-CTA kernels (tbk_api.cu); no-op when the queue is empty
int peer_flush(tbk_peer* p, cudaStream_t st);

inline PeerView peer_none() {
  PeerView v;
  memset(&v, 0, sizeof(v));
  return v;
}
inline bool peer_active(const tbk_peer* p) { return p && p->connected && p->nranks > 1; }
inline void peer_view_base(const tbk_peer* p, PeerView& v) {
  v.rank = p->rank; v.nranks = p->nranks;
  for (int r = 0; r < p->nranks; ++r) v.box[r] = p->box[r];
}
inline double* peer_local_slot(const tbk_peer* p, unsigned long long epoch) {
  return (double*)((char*)p->box[p->rank] + kPeerMailboxBytes) + (size_t)(epoch % kPeerDepth) * kPeerMaxVals;
}
// attach every not yet posted queue entry to the view's post list (the kernel's first CTA posts them)
inline void peer_attach_posts(tbk_peer* p, PeerView& v) {
  for (int i = 0; i < p->nqueue && v.npost < kPeerMaxPost; ++i)
    if (!p->queue[i].posted) { v.post[v.npost++] = p->queue[i]; p->queue[i].posted = 2; }     // 2: posted by THIS kernel
}
// move queue entries to the view's completion list: all of them, or only those a PREVIOUS kernel posted
inline void peer_attach_pends(tbk_peer* p, PeerView& v, bool also_fresh) {
  int keep = 0;
  for (int i = 0; i < p->nqueue; ++i) {
    const bool take = (p->queue[i].posted == 1 || (also_fresh && p->queue[i].posted == 2)) && v.npend < kPeerMaxPend;
    if (take) v.pend[v.npend++] = p->queue[i];
    else p->queue[keep++] = p->queue[i];
  }
  p->nqueue = keep;
}
// what this kernel posts counts as "posted by an earlier kernel" from the next kernel on
inline void peer_age_posts(tbk_peer* p) {
  for (int i = 0; i < p->nqueue; ++i)
    if (p->queue[i].posted == 2) p->queue[i].posted = 1;
}

// PeerView for the next collective of nv values (advances the epoch).
//   completer: a kernel whose last CTA may complete older deferred collectives (the flux kernels, the flush
//   kernel); the grid-solve kernel never does, so that nothing sits between its last CTA and the dependent flux kernel.
// Every collective-capable kernel posts (from its FIRST CTA) whatever the queue holds unposted.  A synchronous
// collective (defer_next not set) also completes the whole queue together with its own result; a deferred one is
// queued and its kernel completes only what an EARLIER kernel posted.  rc != 0: launching a flush failed.
inline int peer_next(tbk_peer* p, int nv, int op, double* out, bool completer, cudaStream_t st, PeerView* view) {
  PeerView v = peer_none();
  if (!peer_active(p)) { if (p) p->defer_next = false; *view = v; return 0; }
  const bool defer = p->defer_next;
  p->defer_next = false;
  // slot-reuse safety / capacities: never run more than kPeerMaxLag epochs ahead of the oldest uncompleted one
  if (p->nqueue > 0 && (p->epoch + 1 - p->queue[0].epoch > (unsigned long long)kPeerMaxLag || p->nqueue >= kPeerQueue - 1 ||
                        (!defer && p->nqueue > kPeerMaxPend) )) {
    if (int rc = peer_flush(p, st)) return rc;
  }
  peer_view_base(p, v);
  v.epoch = ++p->epoch;
  peer_attach_posts(p, v);
  if (!defer) {
    v.mode = 1;
    peer_attach_pends(p, v, true);
  } else {
    v.mode = 2;
    v.local = peer_local_slot(p, v.epoch);
    if (completer) peer_attach_pends(p, v, false);
    PeerPending& q = p->queue[p->nqueue++];
    q.epoch = v.epoch; q.nv = nv; q.op = op; q.out = out; q.local = v.local; q.posted = 0;
  }
  peer_age_posts(p);
  *view = v;
  return 0;
}
}  // namespace tbk
