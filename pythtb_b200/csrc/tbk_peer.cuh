// tbk_peer.cuh — all-reduce of a few doubles across the GPUs of one box, INSIDE the kernel that
// produced them, over NVLink peer memory (no NCCL launch, no extra kernel).
//
// Every rank owns a small "mailbox" in its HBM, mapped into all peers through CUDA IPC
// (tbk_peer_create / tbk_peer_connect).  The protocol is the flag-in-data ("LL") scheme: a double
// travels as two 8-byte words, each carrying 32 bits of payload and the 32-bit tag of the call's
// epoch, and an aligned 8-byte store is a single NVLink transaction — so a word is either absent
// (old tag) or complete, and no release fence / separate flag store is needed.
//
// A collective has three steps that need not run in the same kernel:
//   STORE     (deferred collectives only) the CTA that finishes a rank's local reduction keeps the values in a
//             LOCAL scratch slot — no NVLink traffic at the end of the producing kernel;
//   POST      some CTA stores the values into slot [epoch % depth][rank] of EVERY rank's mailbox (one thread
//             per word, fire and forget);
//   COMPLETE  some CTA polls the words of its OWN mailbox until each carries that epoch's tag, combines the
//             nranks vectors in rank order (deterministic) and writes the result.
// A synchronous collective (the public API returns the number to the caller) posts and completes in the
// producing kernel's last CTA: one exposed NVLink round trip plus the skew between the ranks.
// A DEFERRED collective (tbk_peer_defer) is stored locally by its kernel, posted by the FIRST CTA of the next
// collective-capable kernel on the stream at its START, and completed by the last CTA of a later flux kernel (or
// tbk_peer_flush).  Remote stores at the END of a kernel are what costs: the grid does not complete (and a
// programmatic dependent does not pass griddepcontrol.wait) before they are acknowledged over NVLink, ~2 us
// measured per kernel (profiles/README.md r09: N = 2 efficiency 0.876 with posts at the kernel ends).  Posted from
// a kernel's first CTA they drain under the kernel's body, and a device-resident pipeline (a parameter sweep that
// reads its Chern numbers at the end) never waits for a peer inside a step.
//
// Slot reuse: epoch e and e + kPeerDepth share a slot.  The host handle never lets the newest epoch run more than
// kPeerMaxLag ahead of the oldest uncompleted one (it inserts a flush kernel first); rank r posting e + D has
// therefore completed e + D - L, which needed q's words of e + D - L, which q posted after completing
// e + D - 2L >= e for D >= 2L: the receiver has read a slot before anybody overwrites it.
// All ranks must issue the same sequence of collectives (as with NCCL).  A rank that waits longer
// than ~4 s gives up and returns NaN instead of hanging the GPU.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerMaxVals = 16;                     // doubles per contribution
constexpr int kPeerSlotWords = 2 * kPeerMaxVals;     // 8-byte words per (epoch slot, source rank)
constexpr int kPeerDepth = 16;                       // epochs before a slot is reused
constexpr int kPeerMaxLag = 7;                       // the newest epoch runs at most this far ahead of the oldest uncompleted one
constexpr int kPeerMaxPend = 6;                      // deferred collectives one kernel can complete
constexpr int kPeerMaxPost = 4;                      // deferred collectives one kernel can post
constexpr int kPeerQueue = 8;                        // deferred collectives in flight on the host handle
constexpr size_t kPeerMailboxBytes = (size_t)kPeerDepth * kPeerMaxRanks * kPeerSlotWords * sizeof(unsigned long long);
constexpr size_t kPeerLocalBytes = (size_t)kPeerDepth * kPeerMaxVals * sizeof(double);   // local slots of deferred values
constexpr size_t kPeerScratchBytes = kPeerLocalBytes + 256;   // behind the mailbox: local slots, then the barrier result

struct PeerPending {                                 // a deferred collective on its way
  unsigned long long epoch;
  int nv, op;                                        // op 0: sum in rank order, 1: min
  double* out;                                       // where the combined result goes
  double* local;                                     // this rank's values (local scratch slot of the epoch)
  int posted;                                        // host bookkeeping: a kernel that posts it has been enqueued
};

struct PeerView {
  int rank, nranks;                                  // nranks <= 1: no exchange
  unsigned long long epoch;                          // of this kernel's own collective (0: none)
  int mode;                                          // own collective: 1 synchronous (post + complete here), 2 deferred (store locally)
  double* local;                                     // mode 2: the local slot of `epoch`
  int npost, npend;
  PeerPending post[kPeerMaxPost];                    // deferred collectives this kernel's FIRST CTA posts at its start
  PeerPending pend[kPeerMaxPend];                    // deferred collectives this kernel's last CTA completes
  double* box[kPeerMaxRanks];                        // mailbox of every rank (own one included)
};

#if defined(__CUDACC__)
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned peer_tag(unsigned long long epoch) {
  return (unsigned)epoch | 0x80000000u;              // never the zero of a fresh mailbox, never the tag this
}                                                    // slot carried kPeerDepth collectives ago

// POST: vals[nv] (shared or global memory, visible to the CTA) of `epoch` -> every rank's mailbox.  All threads of one CTA.
__device__ __forceinline__ void peer_post(const PeerView& pv, unsigned long long epoch, const double* s_vals, int nv) {
  const unsigned tag = peer_tag(epoch);
  const int nw = 2 * nv;
  const size_t slot = (size_t)((int)(epoch % kPeerDepth) * pv.nranks + pv.rank) * kPeerSlotWords;
  for (int t = threadIdx.x; t < pv.nranks * nw; t += blockDim.x) {
    const int r = t / nw, w = t - r * nw;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s_vals[w >> 1]);
    const unsigned half = (w & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pv.box[r]) + slot + w, ((unsigned long long)tag << 32) | half);
  }
}

// COMPLETE a list of collectives at once: wait for every rank's words of every entry (ONE round of polling loads for
// all of them: the entries are usually long there, and what is paid is a load latency, not a wait), combine in
// rank order, write out[].  All threads of one CTA.  *s_fail (shared) is set when a peer never arrived (~4 s).
__device__ inline void peer_complete_list(const PeerView& pv, const PeerPending* list, int count, int* s_fail) {
  __shared__ unsigned s_half[kPeerMaxPend + 1][kPeerMaxRanks * kPeerSlotWords];
  const int tid = threadIdx.x;
  int total = 0;
  for (int i = 0; i < count; ++i) total += pv.nranks * 2 * list[i].nv;
  for (int t = tid; t < total; t += blockDim.x) {     // one word per thread: a single round of loads in the usual case
    int i = 0, q = t;
    while (q >= pv.nranks * 2 * list[i].nv) { q -= pv.nranks * 2 * list[i].nv; ++i; }
    const int nw = 2 * list[i].nv;
    const int r = q / nw, w = q - r * nw;
    q = r * kPeerSlotWords + w;
    const unsigned long long epoch = list[i].epoch;
    const unsigned tag = peer_tag(epoch);
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(pv.box[pv.rank]) +
                                    (size_t)((int)(epoch % kPeerDepth) * pv.nranks + r) * kPeerSlotWords + w;
    const long long t0 = clock64();
    unsigned long long x = ld_relaxed_sys_u64(src);
    while ((unsigned)(x >> 32) != tag) {
      if (clock64() - t0 > 8000000000LL) { *s_fail = 1; break; }     // ~4 s at 2 GHz: a peer never arrived
      x = ld_relaxed_sys_u64(src);
    }
    s_half[i][q] = (unsigned)x;
  }
  __syncthreads();
  for (int t = tid; t < count * kPeerMaxVals; t += blockDim.x) {
    const int i = t / kPeerMaxVals, k = t - i * kPeerMaxVals;
    if (k >= list[i].nv) continue;
    const int op = list[i].op;
    double acc = op == 0 ? 0.0 : INFINITY;
    for (int r = 0; r < pv.nranks; ++r) {
      const unsigned long long bits = ((unsigned long long)s_half[i][r * kPeerSlotWords + 2 * k + 1] << 32) |
                                      (unsigned long long)s_half[i][r * kPeerSlotWords + 2 * k];
      const double x = __longlong_as_double((long long)bits);
      acc = op == 0 ? acc + x : fmin(acc, x);
    }
    if (*s_fail) acc = NAN;
    list[i].out[k] = acc;
  }
  __syncthreads();                                   // s_half is free again; out[] is written
}

// Kernel prologue, called by ALL threads of the kernel right at its start (after griddepcontrol.wait where the
// kernel is a programmatic dependent): the first CTA posts the deferred collectives attached to the view from
// their local slots.  The remote stores drain under the kernel's body.
__device__ __forceinline__ void peer_prologue(const PeerView& pv) {
  if (pv.nranks > 1 && pv.npost > 0 && blockIdx.x == 0) {
    for (int i = 0; i < pv.npost; ++i) peer_post(pv, pv.post[i].epoch, pv.post[i].local, pv.post[i].nv);
  }
}

// The older deferred collectives attached to the view, completed by all threads of ONE CTA.  The flux kernel hands
// this to the FIRST CTA that finishes its work (ticket 0): the polls then run while the rest of the wave is still
// finishing, instead of after the last CTA's own reduction (measured: 2.3 us at the end of every step, N = 2).
__device__ inline void peer_complete_pending(const PeerView& pv, int* s_fail) {
  if (pv.nranks > 1 && pv.npend > 0) {
    if (threadIdx.x == 0) *s_fail = 0;
    __syncthreads();
    peer_complete_list(pv, pv.pend, pv.npend, s_fail);
  }
}

// The kernel's OWN collective as its LAST CTA sees it, called by all its threads: vals[nv] are this rank's values.
//   mode 1 (synchronous): post them and complete: out[nv] = sum (op 0, rank order) / min (op 1) over the ranks.
//   mode 2 (deferred): keep them in the local slot (a later kernel's first CTA posts them).
__device__ inline void peer_own(const PeerView& pv, const double* vals, int nv, int op, double* out, int* s_fail) {
  __shared__ double s_vals[kPeerMaxVals];
  const int tid = threadIdx.x;
  if (pv.epoch == 0) return;
  if (tid == 0) *s_fail = 0;
  if (tid < nv) {
    const double x = vals[tid];
    if (pv.mode == 2) pv.local[tid] = x;
    else s_vals[tid] = x;
  }
  __syncthreads();
  if (pv.mode == 1) {
    peer_post(pv, pv.epoch, s_vals, nv);
    PeerPending self;
    self.epoch = pv.epoch; self.nv = nv; self.op = op; self.out = out; self.local = nullptr; self.posted = 1;
    __shared__ PeerPending s_self;
    if (tid == 0) s_self = self;
    __syncthreads();
    peer_complete_list(pv, &s_self, 1, s_fail);
  }
}

// Both in one CTA (kernels with a single finishing CTA: the grid solve's last CTA — which has no pending list —,
// the flush and barrier kernels).
__device__ inline void peer_collective(const PeerView& pv, const double* vals, int nv, int op, double* out, int* s_fail) {
  peer_own(pv, vals, nv, op, out, s_fail);
  peer_complete_pending(pv, s_fail);
}
#endif

}  // namespace tbk
