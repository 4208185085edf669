// tests/hostemu/emu.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the TBK_HD math of pythtb_b200/csrc/*.cuh with g++ so the CPU test
// suite can unit-test the device arithmetic (eigensolvers, Hamiltonian
// assembly, overlap/determinant code) without a GPU.  Never loaded by the
// product package.
#include <vector>
#include <cstring>
#include <thread>
#include <pthread.h>
#include "tbk_common.cuh"
#include "tbk_eig_small.cuh"
#include "tbk_eig_group.cuh"
#include "tbk_plan.cuh"
#include "tbk_berry.cuh"
#include "tbk_eig_blocked.cuh"

using namespace tbk;

struct HostGroup {
  int tid() const { return 0; }
  int size() const { return 1; }
  void sync() {}
  double sum(double x) { return x; }
  int nsub() const { return 1; }
  int sub() const { return 0; }
  int lane() const { return 0; }
  int subsize() const { return 1; }
  double subsum(double x) { return x; }
  void subsync() {}
};

// A CTA emulated by T host threads: the same SPMD code the kernels run, barriers for __syncthreads /
// __syncwarp, "warps" (sub-teams) of S consecutive threads.  Verifies the PARALLEL decomposition of the
// group algorithms (which element each thread owns, what is visible after which barrier), not only their
// arithmetic; a missing barrier shows up as a wrong result or as a ThreadSanitizer report.
struct TeamShared {
  int T, S;
  pthread_barrier_t all;
  std::vector<pthread_barrier_t> sub;
  std::vector<double> red;
  TeamShared(int T_, int S_) : T(T_), S(S_), sub(T_ / S_), red(T_) {
    pthread_barrier_init(&all, nullptr, T);
    for (auto& b : sub) pthread_barrier_init(&b, nullptr, S);
  }
  ~TeamShared() {
    pthread_barrier_destroy(&all);
    for (auto& b : sub) pthread_barrier_destroy(&b);
  }
};
struct TeamGroup {
  TeamShared* sh;
  int t;
  int tid() const { return t; }
  int size() const { return sh->T; }
  void sync() { pthread_barrier_wait(&sh->all); }
  double sum(double x) {
    sh->red[t] = x;
    sync();
    double s = 0.0;
    for (int i = 0; i < sh->T; ++i) s += sh->red[i];
    sync();
    return s;
  }
  int nsub() const { return sh->T / sh->S; }
  int sub() const { return t / sh->S; }
  int lane() const { return t % sh->S; }
  int subsize() const { return sh->S; }
  void subsync() { pthread_barrier_wait(&sh->sub[sub()]); }
  double subsum(double x) {
    sh->red[t] = x;
    subsync();
    double s = 0.0;
    for (int i = sub() * sh->S; i < (sub() + 1) * sh->S; ++i) s += sh->red[i];
    subsync();
    return s;
  }
};

static int g_hetrd_sym = 1;

template <int N>
static int emu_eigvals_n(const cplx* h, double* ev) {
  cplx a[N][N];
  for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) a[r][c] = h[r * N + c];
  return eigvals_small<N>(a, ev) ? 1 : 0;
}

// eigenvectors-through-memory register solver (n = 5..8 with eigenvectors); returns 1 if converged
template <int N>
static int emu_eigh_mem_n(const cplx* h, double* ev, cplx* w) {
  cplx a[N][N];
  for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) a[r][c] = h[r * N + c];
  std::vector<double> zs((size_t)N * N * 3);
  return eigh_small_mem<N>(a, ev, zs.data(), 3, [&](int b, const cplx (&x)[N]) { for (int o = 0; o < N; ++o) w[b * N + o] = x[o]; }) ? 1 : 0;
}

extern "C" {

// which tridiagonalisation the emu_heev_blocked* entry points run: 1 = lower triangle only, 0 = full matrix
void emu_set_hetrd_sym(int on) { g_hetrd_sym = on; }

// The blocked solver run by a team of T threads in sub-teams of S (T a multiple of S), as solve_blocked_kernel
// runs it with T = 256 / 512 and S = 32.  Same contract as emu_heev_blocked.
int emu_heev_blocked_team(int n, double* A, int lda, int nb, int want_vec, double* ev, double* evec, int T, int S) {
  if (T < 1 || S < 1 || T % S) return -1;
  TeamShared shd(T, S);
  BlkWork w;
  w.n = n; w.lda = lda; w.nb = nb; w.A = (cplx*)A;
  w.nred = T / S > 3 ? 3 : T / S;                // fewer slots than sub-teams: several reduction rounds
  std::vector<char> sh(blk_shared_bytes(n, nb, w.nred, T) + 64);
  blk_carve_shared(w, sh.data(), T);
  int nt = ((n + S - 1) / S) * S;
  if (nt > T) nt = T;
  std::vector<double> Z((size_t)n * n), lu((size_t)4 * n * nt);
  w.Z = Z.data(); w.lu = lu.data(); w.nt = nt;
  cplx* out = (cplx*)evec;
  std::vector<int> fail(T, 0);
  auto body = [&](int t) {
    TeamGroup g{&shd, t};
    if (g_hetrd_sym) hetrd_blocked<kBlkMaxN>(g, w);
    else hetrd_blocked_full(g, w);
    const double tnorm = tridiag_bisect(g, w);
    if (t == 0) std::memcpy(ev, w.lam, n * 8);
    if (!want_vec) return;
    fail[t] = tridiag_invit(g, w, tnorm);
    if (fail[t]) return;                         // uniform across the team by construction
    backtransform_all<kBlkMaxN, 2>(g, w, w.V, 2 * nb, [&](int c, int r, cplx x) { out[(size_t)c * n + r] = x; });
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(body, t);
  body(0);
  for (auto& x : th) x.join();
  for (int t = 1; t < T; ++t)
    if (fail[t] != fail[0]) return -2;           // the fallback decision must not diverge inside a CTA
  return fail[0];
}

void emu_eigh2_fast(double h00, double h11, double h10re, double h10im, double* ev, double* w) {
  cplx ww[2][2];
  eigh2_fast(h00, h11, mk(h10re, h10im), ev, ww);
  std::memcpy(w, ww, sizeof(ww));
}

void emu_eigh2(double h00, double h11, double h10re, double h10im, double* ev, double* w) {
  cplx ww[2][2];
  eigh2(h00, h11, mk(h10re, h10im), ev, ww, true);
  std::memcpy(w, ww, sizeof(ww));
}

// H: full n x n row-major complex (interleaved); lower triangle is used.
int emu_jacobi(int n, const double* H, double* ev, double* w) {
  const cplx* h = (const cplx*)H;
  if (n == 3) {
    double dg[3]; cplx lo[3]; cplx ww[3][3];
    for (int r = 0; r < 3; ++r) { dg[r] = h[r * 3 + r].re; for (int c = 0; c < r; ++c) lo[JacobiPacked<3>::idx(r, c)] = h[r * 3 + c]; }
    JacobiPacked<3>::solve(dg, lo, ww, true);
    std::memcpy(ev, dg, sizeof(dg)); std::memcpy(w, ww, sizeof(ww));
    return 0;
  }
  if (n == 4) {
    double dg[4]; cplx lo[6]; cplx ww[4][4];
    for (int r = 0; r < 4; ++r) { dg[r] = h[r * 4 + r].re; for (int c = 0; c < r; ++c) lo[JacobiPacked<4>::idx(r, c)] = h[r * 4 + c]; }
    JacobiPacked<4>::solve(dg, lo, ww, true);
    std::memcpy(ev, dg, sizeof(dg)); std::memcpy(w, ww, sizeof(ww));
    return 0;
  }
  return -1;
}

// direct small solver (Householder + QL), n = 3 or 4; returns 1 if the QL iteration converged
int emu_small_ql(int n, const double* H, double* ev, double* w) {
  const cplx* h = (const cplx*)H;
  if (n == 3) {
    cplx a[3][3]; cplx ww[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = h[r * 3 + c];
    const bool ok = eigh_small_ql<3>(a, ev, ww);
    std::memcpy(w, ww, sizeof(ww));
    return ok ? 1 : 0;
  }
  if (n == 4) {
    cplx a[4][4]; cplx ww[4][4];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) a[r][c] = h[r * 4 + c];
    const bool ok = eigh_small_ql<4>(a, ev, ww);
    std::memcpy(w, ww, sizeof(ww));
    return ok ? 1 : 0;
  }
  return -1;
}

// eigenvalues-only register solver (n = 5..8 band-structure sweeps); returns 1 if the QL iteration converged
int emu_eigh_small_mem(int n, const double* H, double* ev, double* w) {
  const cplx* h = (const cplx*)H;
  cplx* ww = (cplx*)w;
  switch (n) {
    case 3: return emu_eigh_mem_n<3>(h, ev, ww);
    case 5: return emu_eigh_mem_n<5>(h, ev, ww);
    case 6: return emu_eigh_mem_n<6>(h, ev, ww);
    case 7: return emu_eigh_mem_n<7>(h, ev, ww);
    case 8: return emu_eigh_mem_n<8>(h, ev, ww);
  }
  return -1;
}
int emu_eigvals_small(int n, const double* H, double* ev) {
  const cplx* h = (const cplx*)H;
  switch (n) {
    case 3: return emu_eigvals_n<3>(h, ev);
    case 4: return emu_eigvals_n<4>(h, ev);
    case 5: return emu_eigvals_n<5>(h, ev);
    case 6: return emu_eigvals_n<6>(h, ev);
    case 7: return emu_eigvals_n<7>(h, ev);
    case 8: return emu_eigvals_n<8>(h, ev);
  }
  return -1;
}

// the N = 4 solver of the mesh kernels (closed-form tridiagonal solver + QL lane): returns 0 = fast lane, 1 = QL lane,
// -1 = the QL lane did not converge
int emu_eigh4_direct(const double* H, double* ev, double* w) {
  const cplx* h = (const cplx*)H;
  cplx a[4][4]; cplx ww[4][4];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) a[r][c] = h[r * 4 + c];
  int lane = 0;
  const bool ok = eigh4_direct(a, ev, ww, &lane);
  std::memcpy(w, ww, sizeof(ww));
  return ok ? lane : -1;
}

// A: n x n column-major with leading dimension lda (lower triangle valid).
// Outputs: ev[n] ascending, evec[n][n] rows = eigenvectors (reference layout).
int emu_heev_group(int n, double* A, int lda, int want_vec, double* ev, double* evec) {
  std::vector<char> buf(eig_scratch_bytes(n));
  EigScratch s = eig_scratch_carve(buf.data(), n);
  HostGroup g;
  cplx* a = (cplx*)A;
  int info = heev_group(g, n, a, lda, s, want_vec != 0);
  std::vector<int> rank(n);
  eig_rank(g, n, s.d, rank.data());
  cplx* out = (cplx*)evec;
  for (int i = 0; i < n; ++i) {
    ev[rank[i]] = s.d[i];
    if (want_vec)
      for (int o = 0; o < n; ++o) out[(size_t)rank[i] * n + o] = a[o + (size_t)i * lda];
  }
  return info;
}

// The group solver run by a team of T threads (solve_tile_kernel: T = 8 / 16 / 32 lanes of a warp;
// solve_block_kernel: a whole CTA).  Same contract as emu_heev_group.
int emu_heev_group_team(int n, double* A, int lda, int want_vec, double* ev, double* evec, int T) {
  if (T < 1) return -1;
  TeamShared shd(T, T);
  std::vector<char> buf(eig_scratch_bytes(n));
  EigScratch s = eig_scratch_carve(buf.data(), n);
  cplx* a = (cplx*)A;
  std::vector<int> rank(n), info(T, 0);
  auto body = [&](int t) {
    TeamGroup g{&shd, t};
    info[t] = heev_group(g, n, a, lda, s, want_vec != 0);
    eig_rank(g, n, s.d, rank.data());
    g.sync();
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(body, t);
  body(0);
  for (auto& x : th) x.join();
  cplx* out = (cplx*)evec;
  for (int i = 0; i < n; ++i) {
    ev[rank[i]] = s.d[i];
    if (want_vec)
      for (int o = 0; o < n; ++o) out[(size_t)rank[i] * n + o] = a[o + (size_t)i * lda];
  }
  for (int t = 1; t < T; ++t)
    if (info[t] != info[0]) return -2;
  return info[0];
}

// Blocked solver (tbk_eig_blocked.cuh): A n x n column-major, leading dimension lda, lower triangle valid.
// Outputs ev[n] ascending, evec[n][n] rows = eigenvectors; returns 0, 1 = the spectrum asks for the fallback.
int emu_heev_blocked(int n, double* A, int lda, int nb, int want_vec, double* ev, double* evec, double* tri) {
  HostGroup g;
  BlkWork w;
  w.n = n; w.lda = lda; w.nb = nb; w.A = (cplx*)A;
  w.nred = 1;
  std::vector<char> sh(blk_shared_bytes(n, nb, 1, 1) + 64);
  blk_carve_shared(w, sh.data(), 1);
  std::vector<double> Z((size_t)n * n), lu((size_t)4 * n);
  w.Z = Z.data(); w.lu = lu.data(); w.nt = 1;
  if (g_hetrd_sym) hetrd_blocked<kBlkMaxN>(g, w);
  else hetrd_blocked_full(g, w);
  if (tri) { std::memcpy(tri, w.d, n * 8); std::memcpy(tri + n, w.e, n * 8); }
  const double tnorm = tridiag_bisect(g, w);
  std::memcpy(ev, w.lam, n * 8);
  if (!want_vec) return 0;
  if (tridiag_invit(g, w, tnorm)) return 1;
  cplx* out = (cplx*)evec;
  if (nb % 2) {          // odd panel width: exercise the per-column variant
    for (int c = 0; c < n; ++c)
      backtransform_column<kBlkMaxN>(g, w, c, [&](int r, cplx x) { out[(size_t)c * n + r] = x; });
  } else {
    if (nb % 3 == 1) backtransform_all<kBlkMaxN, 3>(g, w, w.V, 2 * nb, [&](int c, int r, cplx x) { out[(size_t)c * n + r] = x; });
    else backtransform_all<kBlkMaxN, 1>(g, w, w.V, 2 * nb, [&](int c, int r, cplx x) { out[(size_t)c * n + r] = x; });
  }
  return 0;
}

// Hamiltonian assembly from a compiled plan (host pointers), Convention I/II.
void emu_gen_ham(const PlanView* pv, const double* k, int64_t nk, double* H) {
  const int n = pv->nsta;
  std::vector<cplx> ph(pv->nph > 0 ? pv->nph : 1);
  for (int64_t ik = 0; ik < nk; ++ik) {
    cplx* h = (cplx*)H + (size_t)ik * n * n;
    plan_phases(*pv, k + ik * pv->dim_k, ph.data(), 1);
    for (int e = 0; e < pv->nel; ++e) {
      const cplx v = plan_element(*pv, e, ph.data(), 1);
      const int r = pv->el_row[e], c = pv->el_col[e];
      h[(size_t)r * n + c] = v;
      if (r != c) h[(size_t)c * n + r] = conj(v);
    }
    if (pv->convention == 1) plan_gauge_matrix(*pv, k + ik * pv->dim_k, h, n);
  }
}

// det of the overlap matrix of two row-blocks a[nocc][n], b[nocc][n]; returns
// the unit-modulus phase of the determinant (re, im) and log|det|.
void emu_link_det(int nocc, int n, const double* a, const double* b, double* out3) {
  std::vector<cplx> M((size_t)nocc * nocc);
  overlap_rows((const cplx*)a, n, (const cplx*)b, n, nocc, n, M.data(), nocc);
  double lg;
  cplx u = lu_det_phase(M.data(), nocc, nocc, &lg);
  out3[0] = u.re; out3[1] = u.im; out3[2] = lg;
}

// polar factor of a small matrix (in place), row-major nocc x nocc
int emu_polar(int nocc, double* M) {
  std::vector<cplx> w1((size_t)nocc * nocc), w2((size_t)nocc * nocc);
  return polar_unitary((cplx*)M, nocc, w1.data(), w2.data());
}

// eigenvalue phases of a unitary (general complex) matrix, row-major
int emu_eigvals(int n, double* M, double* ev_re_im) {
  return comqr_eigvals((cplx*)M, n, n, (cplx*)ev_re_im);
}

}  // extern "C"
