// tbk_peer.cuh — all-reduce of a few doubles across the GPUs of one box, INSIDE the kernel that
// produced them, over NVLink peer memory (no NCCL launch, no extra kernel).
//
// Every rank owns a small "mailbox" in its HBM, mapped into all peers through CUDA IPC
// (tbk_peer_create / tbk_peer_connect).  The CTA that finishes a rank's local reduction
//   1. stores its vector into slot [parity][rank] of EVERY rank's mailbox (peer stores over NVLink),
//      then the call's epoch into the slot's flag word with st.release.sys;
//   2. spins (ld.acquire.sys) on the nranks flags of its OWN mailbox until all carry this epoch;
//   3. combines the nranks vectors in rank order (deterministic) and writes the result.
// Two parities make slot reuse safe: a rank can only be one collective ahead of the slowest rank,
// because finishing collective e requires everybody's contribution to e.  All ranks must issue the
// same sequence of collectives (as with NCCL).  A rank that waits longer than ~4 s gives up and
// returns NaN instead of hanging the GPU.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerMaxVals = 16;                     // doubles per contribution
constexpr int kPeerSlot = 1 + kPeerMaxVals;          // flag + values, in doubles
constexpr size_t kPeerMailboxBytes = (size_t)2 * kPeerMaxRanks * kPeerSlot * sizeof(double);

struct PeerView {
  int rank, nranks;                                  // nranks <= 1: no exchange
  unsigned long long epoch;                          // > 0, identical on all ranks for one collective
  double* box[kPeerMaxRanks];                        // mailbox of every rank (own one included)
};

#if defined(__CUDACC__)
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Called by ALL threads of ONE CTA per rank (blockDim >= nranks).  vals[nv] (shared or global,
// written before a __syncthreads by the caller) -> out[nv] = sum / min over ranks.
// op: 0 = sum (rank order), 1 = min.
__device__ inline void peer_allreduce(const PeerView& pv, const double* vals, int nv, int op, double* out, int* s_fail) {
  const int tid = threadIdx.x;
  if (tid == 0) *s_fail = 0;
  __syncthreads();
  const int parity = (int)(pv.epoch & 1ull);
  if (tid < pv.nranks) {
    double* dst = pv.box[tid] + (size_t)(parity * pv.nranks + pv.rank) * kPeerSlot;
    for (int v = 0; v < nv; ++v) st_relaxed_sys(dst + 1 + v, vals[v]);
    st_release_sys(reinterpret_cast<unsigned long long*>(dst), pv.epoch);
    const unsigned long long* flag =
        reinterpret_cast<const unsigned long long*>(pv.box[pv.rank] + (size_t)(parity * pv.nranks + tid) * kPeerSlot);
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != pv.epoch) {
      if (clock64() - t0 > 8000000000LL) { *s_fail = 1; break; }     // ~4 s at 2 GHz: a peer never arrived
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (tid < nv) {
    double acc = op == 0 ? 0.0 : INFINITY;
    for (int r = 0; r < pv.nranks; ++r) {
      const double x = ld_relaxed_sys(pv.box[pv.rank] + (size_t)(parity * pv.nranks + r) * kPeerSlot + 1 + tid);
      acc = op == 0 ? acc + x : fmin(acc, x);
    }
    out[tid] = *s_fail ? NAN : acc;
  }
}
#endif

}  // namespace tbk
