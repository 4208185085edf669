// tbk_peer.cuh — all-reduce of a few doubles across the GPUs of one box, INSIDE the kernel that
// produced them, over NVLink peer memory (no NCCL launch, no extra kernel).
//
// Every rank owns a small "mailbox" in its HBM, mapped into all peers through CUDA IPC
// (tbk_peer_create / tbk_peer_connect).  The protocol is the flag-in-data ("LL") scheme: a double
// travels as two 8-byte words, each carrying 32 bits of payload and the 32-bit tag of the call's
// epoch, and an aligned 8-byte store is a single NVLink transaction — so a word is either absent
// (old tag) or complete, and no release fence / separate flag store is needed.
//
// A collective has two halves that need not run in the same kernel:
//   POST      the CTA that finishes a rank's local reduction stores its words into slot
//             [epoch % depth][rank] of EVERY rank's mailbox (one thread per word, fire and forget);
//   COMPLETE  some CTA polls the words of its OWN mailbox until each carries that epoch's tag,
//             combines the nranks vectors in rank order (deterministic) and writes the result.
// A synchronous collective (the public API returns the number to the caller) does both in the
// producing kernel's last CTA: one exposed NVLink round trip plus the skew between the ranks.
// A DEFERRED collective (tbk_peer_defer) is only posted by its kernel; it is completed by a later
// kernel on the stream — the next flux kernel completes the collectives that are at least two epochs
// old, i.e. those of the PREVIOUS solve + flux step, whose words arrived tens of microseconds ago —
// or by tbk_peer_flush.  A device-resident pipeline (a parameter sweep that reads its Chern numbers
// at the end) therefore never waits for a peer inside a step: no exposed round trip, no skew.
//
// Slot reuse: epoch e and e + kPeerDepth share a slot.  The host handle never lets a posted epoch
// lag more than kPeerMaxLag behind (it inserts a completing kernel first), and a rank posts e only
// after completing everything up to e - kPeerMaxLag; rank r posting e + D has therefore completed
// e + D - L, which needed q's words of e + D - L, which q posted after completing e + D - 2L >= e
// for D >= 2L: the receiver has read a slot before anybody overwrites it.
// All ranks must issue the same sequence of collectives (as with NCCL).  A rank that waits longer
// than ~4 s gives up and returns NaN instead of hanging the GPU.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerMaxVals = 16;                     // doubles per contribution
constexpr int kPeerSlotWords = 2 * kPeerMaxVals;     // 8-byte words per (epoch slot, source rank)
constexpr int kPeerDepth = 8;                        // epochs before a slot is reused
constexpr int kPeerMaxLag = 4;                       // a posted collective is completed at most this many epochs later
constexpr int kPeerMaxPend = 4;                      // deferred collectives one kernel can complete
constexpr size_t kPeerMailboxBytes = (size_t)kPeerDepth * kPeerMaxRanks * kPeerSlotWords * sizeof(unsigned long long);
constexpr size_t kPeerScratchBytes = 512;            // local scratch behind the mailbox (barrier result)

struct PeerPending {                                 // a posted, not yet completed collective
  unsigned long long epoch;
  int nv, op;                                        // op 0: sum in rank order, 1: min
  double* out;
};

struct PeerView {
  int rank, nranks;                                  // nranks <= 1: no exchange
  unsigned long long epoch;                          // of this kernel's own post (0: it posts nothing)
  int complete_self;                                 // 1: also wait for the peers' words of `epoch` and write the result
  int npend;
  PeerPending pend[kPeerMaxPend];                    // older collectives this kernel completes
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

// POST: s_vals[nv] (shared memory, visible to the CTA) -> every rank's mailbox.  All threads of one CTA.
__device__ __forceinline__ void peer_post(const PeerView& pv, const double* s_vals, int nv) {
  const unsigned tag = peer_tag(pv.epoch);
  const int nw = 2 * nv;
  const size_t slot = (size_t)((int)(pv.epoch % kPeerDepth) * pv.nranks + pv.rank) * kPeerSlotWords;
  for (int t = threadIdx.x; t < pv.nranks * nw; t += blockDim.x) {
    const int r = t / nw, w = t - r * nw;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s_vals[w >> 1]);
    const unsigned half = (w & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pv.box[r]) + slot + w, ((unsigned long long)tag << 32) | half);
  }
}

// COMPLETE: wait for every rank's words of `epoch`, combine in rank order, write out[nv].  All threads of one CTA;
// s_half: [kPeerMaxRanks * kPeerSlotWords] shared words; *s_fail (shared) is set when a peer never arrived.
__device__ inline void peer_complete(const PeerView& pv, unsigned long long epoch, int nv, int op, double* out,
                                     unsigned* s_half, int* s_fail) {
  const unsigned tag = peer_tag(epoch);
  const int nw = 2 * nv, tid = threadIdx.x;
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pv.box[pv.rank]) +
                                   (size_t)((int)(epoch % kPeerDepth) * pv.nranks) * kPeerSlotWords;
  for (int t = tid; t < pv.nranks * nw; t += blockDim.x) {
    const int r = t / nw, w = t - r * nw;
    const unsigned long long* src = mine + (size_t)r * kPeerSlotWords + w;
    const long long t0 = clock64();
    unsigned long long x = ld_relaxed_sys_u64(src);
    while ((unsigned)(x >> 32) != tag) {
      if (clock64() - t0 > 8000000000LL) { *s_fail = 1; break; }     // ~4 s at 2 GHz: a peer never arrived
      x = ld_relaxed_sys_u64(src);
    }
    s_half[r * kPeerSlotWords + w] = (unsigned)x;
  }
  __syncthreads();
  if (tid < nv) {
    double acc = op == 0 ? 0.0 : INFINITY;
    for (int r = 0; r < pv.nranks; ++r) {
      const unsigned long long bits = ((unsigned long long)s_half[r * kPeerSlotWords + 2 * tid + 1] << 32) |
                                      (unsigned long long)s_half[r * kPeerSlotWords + 2 * tid];
      const double x = __longlong_as_double((long long)bits);
      acc = op == 0 ? acc + x : fmin(acc, x);
    }
    if (*s_fail) acc = NAN;
    out[tid] = acc;
  }
  __syncthreads();                                   // s_half is free again; out[] is written
}

// The collective as one kernel sees it, called by ALL threads of ONE CTA per rank: post vals[nv] (if this
// kernel has an epoch of its own), complete the older collectives attached to the view, and — for a synchronous
// collective — complete its own: out[nv] = sum (op 0, rank order) / min (op 1) over the ranks.  The post goes
// first: the peers' waits then overlap this rank's completions.
__device__ inline void peer_collective(const PeerView& pv, const double* vals, int nv, int op, double* out, int* s_fail) {
  __shared__ double s_vals[kPeerMaxVals];
  __shared__ unsigned s_half[kPeerMaxRanks * kPeerSlotWords];
  const int tid = threadIdx.x;
  if (tid == 0) *s_fail = 0;
  if (pv.epoch != 0 && tid < nv) s_vals[tid] = vals[tid];
  __syncthreads();
  if (pv.epoch != 0) peer_post(pv, s_vals, nv);
  for (int i = 0; i < pv.npend; ++i) peer_complete(pv, pv.pend[i].epoch, pv.pend[i].nv, pv.pend[i].op, pv.pend[i].out, s_half, s_fail);
  if (pv.epoch != 0 && pv.complete_self) peer_complete(pv, pv.epoch, nv, op, out, s_half, s_fail);
}
#endif

}  // namespace tbk
