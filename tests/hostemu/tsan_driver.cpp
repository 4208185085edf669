// tests/hostemu/tsan_driver.cpp — TEST INFRASTRUCTURE ONLY.
// The blocked eigensolver of pythtb_b200/csrc/tbk_eig_blocked.cuh run by a team of host threads (emu.cpp:
// TeamGroup, one thread per CUDA thread, pthread barriers for __syncthreads / __syncwarp) under
// ThreadSanitizer: a shared-memory or workspace access that is not ordered by a barrier — i.e. a missing
// __syncthreads()/__syncwarp() in the kernel — is reported as a data race.  Built and run by
// tests/test_hostemu_math.py::test_blocked_solver_has_no_unsynchronised_accesses.
//   usage: tsan_driver n nb T S kind seed     (kind 0 random, 1 exactly degenerate pairs, 2 banded "ribbon",
//                                              3 two flat levels of n/2 states each: the CTA-wide cluster path)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "emu.cpp"

int main(int argc, char** argv) {
  if (argc < 7) { std::printf("usage: tsan_driver n nb T S kind seed\n"); return 2; }
  const int n = std::atoi(argv[1]), nb = std::atoi(argv[2]), T = std::atoi(argv[3]), S = std::atoi(argv[4]);
  const int kind = std::atoi(argv[5]);
  std::mt19937_64 rng(std::atoi(argv[6]));
  std::normal_distribution<double> gauss(0.0, 1.0);
  std::vector<cplx> H((size_t)n * n, mk(0.0, 0.0));
  if (kind == 2) {
    for (int i = 0; i + 1 < n; ++i) {
      const cplx t = mk(-1.0 - (i % 2) * std::cos(0.7), -(i % 2) * std::sin(0.7));
      H[(size_t)i * n + i + 1] = t;
      H[(size_t)(i + 1) * n + i] = conj(t);
    }
    for (int i = 0; i < n; ++i) H[(size_t)i * n + i] = mk((i % 2) ? -0.4 : 0.4, 0.0);
  } else {
    for (int r = 0; r < n; ++r)
      for (int c = 0; c <= r; ++c) {
        cplx v = mk(gauss(rng), r == c ? 0.0 : gauss(rng));
        if (kind == 1) {                               // H (x) I_2 structure: every level twice
          const int rr = r / 2, cc = c / 2;
          v = (r % 2 == c % 2) ? mk(std::sin(1.3 * rr + 0.7 * cc) + (rr == cc ? rr : 0.0), rr == cc ? 0.0 : std::cos(rr - 2.1 * cc)) : mk(0.0, 0.0);
          if (rr < cc) v = mk(0.0, 0.0);
        }
        H[(size_t)r * n + c] = v;
        H[(size_t)c * n + r] = conj(v);
      }
    if (kind == 1)                                      // make it exactly Hermitian from the lower triangle
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < r; ++c) H[(size_t)c * n + r] = conj(H[(size_t)r * n + c]);
  }
  if (kind == 3) {                                      // H = 1 - 2 P, P = projector on n/2 random vectors (Gram-Schmidt)
    const int m = n / 2;
    std::vector<cplx> Q((size_t)m * n);
    for (int a = 0; a < m; ++a) {
      for (int r = 0; r < n; ++r) Q[(size_t)a * n + r] = mk(gauss(rng), gauss(rng));
      for (int pass = 0; pass < 2; ++pass)
        for (int b = 0; b < a; ++b) {
          cplx d = mk(0.0, 0.0);
          for (int r = 0; r < n; ++r) d = d + conj(Q[(size_t)b * n + r]) * Q[(size_t)a * n + r];
          for (int r = 0; r < n; ++r) Q[(size_t)a * n + r] = Q[(size_t)a * n + r] - d * Q[(size_t)b * n + r];
        }
      double nrm = 0.0;
      for (int r = 0; r < n; ++r) nrm += norm2(Q[(size_t)a * n + r]);
      nrm = 1.0 / std::sqrt(nrm);
      for (int r = 0; r < n; ++r) Q[(size_t)a * n + r] = nrm * Q[(size_t)a * n + r];
    }
    for (int r = 0; r < n; ++r)
      for (int c = 0; c <= r; ++c) {
        cplx acc = mk(r == c ? 1.0 : 0.0, 0.0);
        for (int a = 0; a < m; ++a) acc = acc - 2.0 * (Q[(size_t)a * n + r] * conj(Q[(size_t)a * n + c]));
        if (r == c) acc.im = 0.0;
        H[(size_t)r * n + c] = acc;
        H[(size_t)c * n + r] = conj(acc);
      }
  }
  const int lda = n | 1;
  std::vector<cplx> A((size_t)n * lda, mk(0.0, 0.0));
  for (int r = 0; r < n; ++r)
    for (int c = 0; c <= r; ++c) A[r + (size_t)c * lda] = H[(size_t)r * n + c];
  std::vector<double> ev(n);
  std::vector<cplx> vec((size_t)n * n);
  // nb == 0 selects the group solver (Householder + implicit QL by one team: solve_tile_kernel / solve_block_kernel)
  const int rc = nb == 0 ? emu_heev_group_team(n, (double*)A.data(), lda, 1, ev.data(), (double*)vec.data(), T)
                         : emu_heev_blocked_team(n, (double*)A.data(), lda, nb, 1, ev.data(), (double*)vec.data(), T, S);
  if (rc != 0) { std::printf("rc %d\n", rc); return rc == 1 ? 0 : 3; }   // 1 = fallback requested: legitimate
  double worst = 0.0, scale = 1.0;
  for (const cplx& h : H) scale = std::fmax(scale, std::fabs(h.re) + std::fabs(h.im));
  for (int b = 0; b < n; ++b)
    for (int r = 0; r < n; ++r) {
      cplx acc = mk(0.0, 0.0);
      for (int c = 0; c < n; ++c) acc = acc + H[(size_t)r * n + c] * vec[(size_t)b * n + c];
      const cplx d = acc - ev[b] * vec[(size_t)b * n + r];
      worst = std::fmax(worst, std::fabs(d.re) + std::fabs(d.im));
    }
  std::printf("n %d nb %d T %d S %d kind %d residual %.3e\n", n, nb, T, S, kind, worst / scale);
  return worst / scale < 1e-11 * n ? 0 : 4;
}
