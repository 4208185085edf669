// tbk_plan.cuh — the compiled tight-binding model ("plan") as the kernels see
// it, and the Hamiltonian-assembly arithmetic that replaces
// tb_model._gen_ham (pythtb.py:874-925).
//
// The host-side model compiler (pythtb_b200/_plan.py) flattens
// _site_energies/_hoppings/_orb/_per into:
//   * a table of UNIQUE lattice vectors R (periodic components only, unique up
//     to sign) — one sincospi per entry per k-point instead of one per hopping;
//   * a list of scalar terms  H[row][col] += amp * E_p  (or amp * conj(E_p),
//     or amp for on-site terms), grouped by lower-triangle matrix element in
//     the reference's accumulation order (pythtb.py:894-924), and the same
//     terms grouped by phase;
//   * tau[state][dim_k], the periodic components of the orbital positions.
// The kernels assemble the Convention-II matrix  H_II = sum_R T(R) e^{2 pi i k.R}
// and obtain PythTB's Convention-I objects through the diagonal gauge
// D = diag(e^{2 pi i k.tau_j}):  H_I = D^H H_II D,  u_I = D^H u_II
// (doc/formalism/pythtb-formalism.tex:341-364).  Eigenvalues are identical.
#pragma once
#include "tbk_common.cuh"

namespace tbk {

constexpr int TBK_PH_CONJ = 1 << 30;   // flag in t_ph / pm_el: use conj(E_p)
constexpr int TBK_PH_MASK = TBK_PH_CONJ - 1;
constexpr int TBK_MAX_DIMK = 4;

struct PlanView {
  int dim_k, nsta, nph, nel, nterm, convention;  // convention: 1 = PythTB (tau in phase), 2 = R only
  const double* ph_R;    // [nph][dim_k]
  const double* tau;     // [nsta][dim_k]
  // element-major CSR over lower-triangle elements (row >= col)
  const int* el_ptr;     // [nel+1]
  const int* el_row;     // [nel]
  const int* el_col;     // [nel]
  const int* t_ph;       // [nterm]  phase index | TBK_PH_CONJ, or -1 (no phase)
  const double* t_amp;   // [nterm][2]
  // phase-major CSR: bucket p < nph holds the terms of phase p, bucket nph the constant terms
  const int* pm_ptr;     // [nph+2]
  const int* pm_el;      // [nterm]  element index | TBK_PH_CONJ
  const double* pm_amp;  // [nterm][2]
};

// E_p = exp(2 pi i k.R_p) for every table entry; ph[p*stride].
TBK_HD void plan_phases(const PlanView& pv, const double* k, cplx* ph, int stride) {
  for (int p = 0; p < pv.nph; ++p) {
    double x = 0.0;
    for (int d = 0; d < pv.dim_k; ++d) x = fma(k[d], pv.ph_R[p * pv.dim_k + d], x);
    ph[(size_t)p * stride] = expi_turns(x);
  }
}

// One lower-triangle element of H_II from precomputed phases.
TBK_HD cplx plan_element(const PlanView& pv, int e, const cplx* ph, int stride) {
  cplx acc = mk(0.0, 0.0);
  const int t1 = pv.el_ptr[e + 1];
  for (int t = pv.el_ptr[e]; t < t1; ++t) {
    const cplx a = mk(pv.t_amp[2 * t], pv.t_amp[2 * t + 1]);
    const int p = pv.t_ph[t];
    if (p < 0) {
      acc = acc + a;
    } else {
      cplx z = ph[(size_t)(p & TBK_PH_MASK) * stride];
      if (p & TBK_PH_CONJ) z.im = -z.im;
      fma_acc(acc, a, z);
    }
  }
  return acc;
}

// d_j = exp(2 pi i k.tau_j)
TBK_HD cplx plan_gauge(const PlanView& pv, const double* k, int j) {
  double x = 0.0;
  for (int d = 0; d < pv.dim_k; ++d) x = fma(k[d], pv.tau[j * pv.dim_k + d], x);
  return expi_turns(x);
}

// H_I = D^H H_II D on a full row-major n x n matrix (serial helper).
TBK_HD void plan_gauge_matrix(const PlanView& pv, const double* k, cplx* h, int n) {
  for (int r = 0; r < n; ++r) {
    const cplx dr = plan_gauge(pv, k, r);
    for (int c = 0; c < n; ++c) {
      const cplx dc = plan_gauge(pv, k, c);
      h[(size_t)r * n + c] = cmul(dr, h[(size_t)r * n + c]) * dc;
    }
  }
}

}  // namespace tbk
