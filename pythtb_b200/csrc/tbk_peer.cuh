// tbk_peer.cuh — all-reduce of a few doubles across the GPUs of one box, INSIDE the kernel that
// produced them, over NVLink peer memory (no NCCL launch, no extra kernel).
//
// Every rank owns a small "mailbox" in its HBM, mapped into all peers through CUDA IPC
// (tbk_peer_create / tbk_peer_connect).  The protocol is the flag-in-data ("LL") scheme: a double
// travels as two 8-byte words, each carrying 32 bits of payload and the 32-bit tag of the call's
// epoch, and an aligned 8-byte store is a single NVLink transaction — so a word is either absent
// (old tag) or complete, and no release fence / separate flag store is needed (the first version
// paid a system-scope release per contribution).  The CTA that finishes a rank's local reduction
//   1. stores its words into slot [parity][rank] of EVERY rank's mailbox (one thread per word);
//   2. polls the words of its OWN mailbox until each carries this epoch's tag;
//   3. combines the nranks vectors in rank order (deterministic) and writes the result.
// A collective can also be DEFERRED: the producing kernel keeps its local result in this rank's own
// memory (no NVLink traffic at all) and the host handle remembers it as pending; the next collective
// kernel on the stream appends the pending values to its own message, so a solve + flux step is ONE
// exchange (one exposed NVLink round trip, one kernel that has to drain remote stores) instead of two.
// Slots are reused every kPeerDepth = 4 epochs: a rank posts epoch e+4 only after it completed e+3,
// which needs every rank's e+3 words, which a rank posts only after it has read everything up to e+2.
// All ranks must issue the same sequence of collectives (as with NCCL).  A rank that waits longer
// than ~4 s gives up and returns NaN instead of hanging the GPU.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerMaxVals = 16;                     // doubles per contribution
constexpr int kPeerSlotWords = 2 * kPeerMaxVals;     // 8-byte words per (parity, source rank) slot
constexpr int kPeerDepth = 4;                        // epochs in flight before a slot is reused
constexpr size_t kPeerMailboxBytes = (size_t)kPeerDepth * kPeerMaxRanks * kPeerSlotWords * sizeof(unsigned long long);
constexpr size_t kPeerScratchBytes = 512;            // local scratch behind the mailbox: [0,16) deferred values, [16] barrier result

struct PeerPending {                                 // a deferred collective (nv 0: none): the local values wait in
  int nv, op;                                        // `local` (this rank's memory) for the next exchange
  double* out;
  const double* local;
};

struct PeerView {
  int rank, nranks;                                  // nranks <= 1: no exchange
  unsigned long long epoch;                          // > 0, identical on all ranks for one collective
  int defer;                                         // 1: no exchange now, the kernel stores its local result to `local`
  double* local;                                     // (the host handle remembers it as pending)
  PeerPending pend;                                  // an earlier deferred collective that rides on this one
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

// The collective, called by ALL threads of ONE CTA per rank.  vals[nv] (shared or global, written before
// a __syncthreads by the caller) -> out[nv] = sum (op 0, rank order) / min (op 1) over the ranks; a pending
// deferred collective attached to the view travels in the same message and is combined into pv.pend.out.
// nv + pv.pend.nv <= kPeerMaxVals.  *s_fail (shared) is set when a peer never arrived; the outputs are then NaN.
__device__ inline void peer_allreduce(const PeerView& pv, const double* vals, int nv, int op, double* out, int* s_fail) {
  __shared__ double s_vals[kPeerMaxVals];
  __shared__ unsigned s_half[kPeerMaxRanks * kPeerSlotWords];
  const int tid = threadIdx.x;
  const int np = pv.pend.nv, nt = nv + np;
  if (tid == 0) *s_fail = 0;
  if (tid < nv) s_vals[tid] = vals[tid];
  else if (tid < nt) s_vals[tid] = pv.pend.local[tid - nv];
  __syncthreads();
  const unsigned tag = peer_tag(pv.epoch);
  const int nw = 2 * nt;
  const int depth = (int)(pv.epoch % kPeerDepth);
  const size_t slot = (size_t)(depth * pv.nranks + pv.rank) * kPeerSlotWords;
  for (int t = tid; t < pv.nranks * nw; t += blockDim.x) {    // 1. one word per thread, to every rank
    const int r = t / nw, w = t - r * nw;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s_vals[w >> 1]);
    const unsigned half = (w & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pv.box[r]) + slot + w, ((unsigned long long)tag << 32) | half);
  }
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pv.box[pv.rank]) +
                                   (size_t)(depth * pv.nranks) * kPeerSlotWords;
  for (int t = tid; t < pv.nranks * nw; t += blockDim.x) {    // 2. wait for every rank's words
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
  if (tid < nt) {                                             // 3. rank order
    const int o = tid < nv ? op : pv.pend.op;
    double acc = o == 0 ? 0.0 : INFINITY;
    for (int r = 0; r < pv.nranks; ++r) {
      const unsigned long long bits = ((unsigned long long)s_half[r * kPeerSlotWords + 2 * tid + 1] << 32) |
                                      (unsigned long long)s_half[r * kPeerSlotWords + 2 * tid];
      const double x = __longlong_as_double((long long)bits);
      acc = o == 0 ? acc + x : fmin(acc, x);
    }
    if (*s_fail) acc = NAN;
    if (tid < nv) out[tid] = acc;
    else pv.pend.out[tid - nv] = acc;
  }
}
#endif

}  // namespace tbk
