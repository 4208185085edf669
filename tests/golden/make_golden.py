"""Generate the parity fixtures by running the UNMODIFIED reference.

Run in the build container (the GPU box has no /root/reference):

    python -m tests.golden.make_golden

It imports ``/root/reference/pythtb.py`` (PythTB 1.8.0) as ``pythtb``, runs
every case of ``tests/cases.py`` through it and stores the results in
``tests/golden/<case>.npz``.  Before trusting them it re-checks the reference
against the reference's own golden data
(``/root/reference/tests/test_examples/*/*/golden_outputs/*.npy``) and records
the deviations in ``tests/golden/golden_log.json``.

It also stores Wannier90 fixtures: the silicon model of
``website/local/w90_example/example_a`` as flat arrays (the ``_hr.dat`` itself is
third-party data and is not copied), its eigenvalues on fixed k-points, and
the reference's parse of a small synthetic Wannier90 data set written by
``tests/w90_synth.py``.
"""
import io
import json
import os
import sys
import contextlib
import datetime
import platform

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import pythtb as ref  # noqa: E402  (the unmodified reference)

from .. import cases  # noqa: E402
from .. import w90_synth  # noqa: E402

REF_GOLD = os.path.join(REF, "tests", "test_examples")

# (our case, our key) -> reference golden file, optional post-processing
CROSS = [
    ("haldane_bands", "evals", "haldane/haldane/golden_outputs/evals.npy"),
    ("haldane_bands", "evals_dos", "haldane/haldane/golden_outputs/evals_dos.npy"),
    ("haldane_bp", "phi_a1", "haldane/haldane_bp/golden_outputs/phi_a1.npy"),
    ("haldane_bp", "phi_b1", "haldane/haldane_bp/golden_outputs/phi_b1.npy"),
    ("haldane_bp", "phi_c1", "haldane/haldane_bp/golden_outputs/phi_c1.npy"),
    ("haldane_bp", "flux_a1", "haldane/haldane_bp/golden_outputs/flux_a1.npy"),
    ("haldane_bp", "flux_a2", "haldane/haldane_bp/golden_outputs/flux_a2.npy"),
    ("cone", "bphase_circ0", "graphene/cone/golden_outputs/bphase_circ0.npy"),
    ("cone", "bphase_circ1", "graphene/cone/golden_outputs/bphase_circ1.npy"),
    ("cone", "bphase_circ01", "graphene/cone/golden_outputs/bphase_circ01.npy"),
    ("cone", "bflux_square_0", "graphene/cone/golden_outputs/bflux_square_0.npy"),
    ("cone", "bflux_square_1", "graphene/cone/golden_outputs/bflux_square_1.npy"),
    ("cone", "bflux_square_01", "graphene/cone/golden_outputs/bflux_square_01.npy"),
    ("cone", "plaq", "graphene/cone/golden_outputs/plaq.npy"),
    ("bn_ribbon", "berry_phase_orig", "boron_nitride/bn_ribbon_berry/golden_outputs/berry_phase_orig.npy"),
    ("bn_ribbon", "berry_phase_perp", "boron_nitride/bn_ribbon_berry/golden_outputs/berry_phase_perp.npy"),
    ("cubic_slab", "evals", "slab/cubic_slab_hwf/golden_outputs/evals.npy"),
    ("cubic_slab", "hwfc", "slab/cubic_slab_hwf/golden_outputs/hwfc.npy"),
    ("cubic_slab", "px", "slab/cubic_slab_hwf/golden_outputs/px.npy"),
    ("three_site", "wann_center", "three_site/3site_cycle/golden_outputs/3site_cycle_wann_centers.npy"),
    ("three_site", "final", "three_site/3site_cycle/golden_outputs/3site_cycle_final.npy"),
    ("misc_bands", "checkerboard", "checkerboard/checkerboard/golden_outputs/evals.npy"),
    ("more_bands", "evals_buckled", "buckling/buckled_layer/golden_outputs/evals.npy"),
    ("more_bands", "evals_trestle", "buckling/trestle/golden_outputs/evals.npy"),
    ("more_bands", "evals_graphene", "graphene/graphene/golden_outputs/evals.npy"),
    ("more_bands", "evals_supercell", "supercell/supercell/golden_outputs/evals.py.npy"),
    ("more_bands", "evals_0dim", "zero_dim/0dim/golden_outputs/evals.npy"),
    ("haldane_finite", "evals_edge", "haldane/edge/golden_outputs/evals.npy"),
    ("haldane_finite", "evals_edge_half", "haldane/edge/golden_outputs/evals_half.npy"),
    ("haldane_finite", "evals_fin_false", "haldane/haldane_fin/golden_outputs/evals_false.npy"),
    ("haldane_finite", "evals_fin_true", "haldane/haldane_fin/golden_outputs/evals_true.npy"),
    ("haldane_hwf", "phi1", "haldane/haldane_hwf/golden_outputs/phi1.npy"),
    ("haldane_hwf", "rib_eval", "haldane/haldane_hwf/golden_outputs/rib_eval.npy"),
    ("haldane_hwf", "jump_k", "haldane/haldane_hwf/golden_outputs/jump_k.npy"),
    ("haldane_hwf", "hwfc_flat", "haldane/haldane_hwf/golden_outputs/hwfcs.npy",
     lambda g: np.concatenate([np.asarray(x, dtype=float).reshape(-1) for x in g])),
    ("haldane_hwf", "pos_exp_sum", "haldane/haldane_hwf/golden_outputs/pos_exps.npy",
     lambda g: np.array([np.sum(np.asarray(x, dtype=float)) for x in g])),
    ("three_site_fin", "fluxes", "three_site/3site_cycle_fin/golden_outputs/3site_cycle_fluxes.npy"),
    ("three_site_fin", "evals_chain", "three_site/3site_cycle_fin/golden_outputs/3site_cycle_fin_evals.npy",
     lambda g: g[:, ::4]),
    ("three_site_fin", "pos_exp_sum", "three_site/3site_cycle_fin/golden_outputs/3site_cycle_fin_xexp.npy",
     lambda g: g[:, ::4].sum(axis=0)),
]


def dump_model(m):
    """Flatten a reference tb_model into plain arrays."""
    nhop = len(m._hoppings)
    blk = (2, 2) if m._nspin == 2 else ()
    amp = np.zeros((nhop,) + blk, dtype=complex)
    hi = np.zeros(nhop, dtype=int)
    hj = np.zeros(nhop, dtype=int)
    hR = np.zeros((nhop, m._dim_r), dtype=int)
    for n, h in enumerate(m._hoppings):
        amp[n] = h[0]
        hi[n], hj[n] = h[1], h[2]
        if m._dim_k > 0:
            hR[n] = h[3]
    return dict(dim_k=m._dim_k, dim_r=m._dim_r, nspin=m._nspin, lat=m._lat, orb=m._orb,
                per=np.array(m._per, dtype=int), site_energies=np.array(m._site_energies),
                hop_amp=amp, hop_i=hi, hop_j=hj, hop_R=hR)


def w90_fixtures():
    out = {}
    # --- silicon (the only complete Wannier90 data set in the reference tree)
    sil = ref.w90(os.path.join(REF, "website/local/w90_example/example_a"), "silicon")
    full = sil.model(zero_energy=0.0)
    small = sil.model(zero_energy=6.2285135, min_hopping_norm=0.01)
    for tag, m in (("full", full), ("small", small)):
        for k, v in dump_model(m).items():
            out["silicon_%s_%s" % (tag, k)] = v
    rng = np.random.RandomState(7)
    kpts = np.vstack([rng.rand(61, 3), [[0, 0, 0], [0.5, 0.0, 0.5], [0.375, 0.375, 0.75]]])
    out["silicon_k"] = kpts
    out["silicon_full_evals"] = full.solve_all(kpts)
    out["silicon_small_evals"] = small.solve_all(kpts)
    with contextlib.redirect_stdout(io.StringIO()):
        (w_kpts, w_ene) = sil.w90_bands_consistency()
    out["silicon_band_kpts"] = w_kpts
    out["silicon_band_ene"] = w_ene
    dev = np.max(np.abs(full.solve_all(w_kpts) - w_ene))
    # --- synthetic data set written by our own writer, parsed by the reference
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        w90_synth.write(tmp, "synth", num_wan=5, seed=11)
        syn = ref.w90(tmp, "synth")
        for tag, kw in (("all", {}), ("cut", dict(min_hopping_norm=0.05, max_distance=4.0,
                                                  ignorable_imaginary_part=0.02, zero_energy=0.3))):
            m = syn.model(**kw)
            for k, v in dump_model(m).items():
                out["synth_%s_%s" % (tag, k)] = v
        k2 = np.random.RandomState(8).rand(9, 3)
        out["synth_k"] = k2
        out["synth_all_evals"] = syn.model().solve_all(k2)
        (dd, hh) = syn.dist_hop()                      # pythtb.py:3590-3645 (no i == j entries at R = 0)
        out["synth_dist_hop_dist"], out["synth_dist_hop_ham"] = dd, hh
        out["synth_shells"] = syn.shells()
    return out, float(dev)


def main():
    log = dict(generated_at=datetime.datetime.now().isoformat(),
               python=platform.python_version(), numpy=np.__version__,
               pythtb=ref.__version__, reference_path=REF, cross_check={}, cases={})
    results = {}
    for name, fn in cases.ALL_CASES.items():
        with contextlib.redirect_stdout(io.StringIO()):
            res = fn(ref)
        res = {k: (np.array([]) if v is None else np.asarray(v)) for k, v in res.items()}
        results[name] = res
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        log["cases"][name] = {k: list(v.shape) for k, v in res.items()}
        print("case %-16s %d arrays" % (name, len(res)))
    worst = 0.0
    for case, key, rel, *post in CROSS:
        gold = np.load(os.path.join(REF_GOLD, rel), allow_pickle=True)
        if post:
            gold = post[0](gold)
        gold = np.asarray(gold, dtype=float)
        ours = results[case][key]
        dev = float(np.max(np.abs(np.asarray(ours, dtype=float).reshape(gold.shape) - gold)))
        log["cross_check"]["%s.%s" % (case, key)] = dict(reference_golden=rel, max_abs_dev=dev)
        worst = max(worst, dev)
        print("cross-check %-28s vs %-70s max|dev| = %.2e" % (case + "." + key, rel, dev))
    km = np.load(os.path.join(REF_GOLD, "kane_mele/kane_mele/golden_outputs/kane_mele_evals.npy"))
    ours = np.array([results["kane_mele"]["evals_even"], results["kane_mele"]["evals_odd"]])
    dev = float(np.max(np.abs(ours - km)))
    log["cross_check"]["kane_mele.evals"] = dict(max_abs_dev=dev)
    worst = max(worst, dev)
    kw = np.load(os.path.join(REF_GOLD, "kane_mele/kane_mele/golden_outputs/kane_mele_wan_cent.npy"))
    ours = np.array([results["kane_mele"]["wan_cent_even"], results["kane_mele"]["wan_cent_odd"]])
    dev = float(np.max(np.abs(ours - kw)))
    log["cross_check"]["kane_mele.wan_cent"] = dict(max_abs_dev=dev)
    worst = max(worst, dev)
    print("cross-check kane_mele evals/wan_cent  max|dev| = %.2e" % dev)
    w90, si_dev = w90_fixtures()
    np.savez_compressed(os.path.join(HERE, "w90.npz"), **w90)
    log["w90"] = dict(note="reference tests hold no w90 goldens (run.py are dummies): parity of the "
                           "w90 path is pinned only to reference outputs generated here",
                      silicon_vs_wannier90_band_dat_max_abs_dev_eV=si_dev)
    log["worst_cross_check_dev"] = worst
    with open(os.path.join(HERE, "golden_log.json"), "w") as f:
        json.dump(log, f, indent=1)
    print("worst deviation from the reference's own golden data: %.2e" % worst)
    assert worst < 1e-10, "reference run here disagrees with the reference's golden outputs"


if __name__ == "__main__":
    main()
