"""The Wannier90 importer (``pythtb_b200.w90``, host code with the interface of pythtb.py:3208-3759).

The reference's tests hold no w90 goldens (tests/test_examples/w90/*/run.py are dummies), so the fixture is
the unmodified reference's own parse of a synthetic data set: ``tests/golden/make_golden.py`` writes it with
``tests/w90_synth.py`` (seed 11), parses it with ``pythtb.w90`` and stores the resulting model arrays and
eigenvalues in ``tests/golden/w90.npz``.  Here the same files are written again, parsed by this package, and
the models must agree entry by entry — hopping list ORDER included, since it fixes the accumulation order of
``_gen_ham`` (pythtb.py:900-924)."""
import os

import numpy as np
import pytest

from oracle import pythtb_oracle as orc
from tests import w90_synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CUT = dict(min_hopping_norm=0.05, max_distance=4.0, ignorable_imaginary_part=0.02, zero_energy=0.3)


def _parse(tmp_path):
    import pythtb_b200
    w90_synth.write(str(tmp_path), "synth", num_wan=5, seed=11)
    return pythtb_b200.w90(str(tmp_path), "synth")


def _check_model(m, z, pre):
    assert (m._dim_k, m._dim_r, m._nspin) == (int(z[pre + "dim_k"]), int(z[pre + "dim_r"]), int(z[pre + "nspin"]))
    assert np.max(np.abs(m._lat - z[pre + "lat"])) < 1e-12
    assert np.max(np.abs(m._orb - z[pre + "orb"])) < 1e-12
    assert list(m._per) == list(z[pre + "per"])
    assert np.max(np.abs(np.asarray(m._site_energies) - z[pre + "site_energies"])) < 1e-12
    assert len(m._hoppings) == len(z[pre + "hop_amp"])
    amp = np.array([h[0] for h in m._hoppings])
    assert np.array_equal(np.array([h[1] for h in m._hoppings]), z[pre + "hop_i"])
    assert np.array_equal(np.array([h[2] for h in m._hoppings]), z[pre + "hop_j"])
    assert np.array_equal(np.array([list(h[3]) for h in m._hoppings]), z[pre + "hop_R"])
    assert np.max(np.abs(amp - z[pre + "hop_amp"])) < 1e-12
    assert m._assume_position_operator_diagonal is False


def test_parser_and_model_match_the_reference_parse(tmp_path):
    z = np.load(os.path.join(GOLD, "w90.npz"))
    w = _parse(tmp_path)
    assert w.num_wan == 5
    _check_model(w.model(), z, "synth_all_")
    _check_model(w.model(**CUT), z, "synth_cut_")
    # the numpy oracle on the parsed model reproduces the reference's eigenvalues
    ev = orc.solve_all(w.model(), z["synth_k"])
    assert np.max(np.abs(ev - z["synth_all_evals"])) < 1e-10
    # helper methods keep the reference's shapes (pythtb.py:3590-3685)
    dist, ham = w.dist_hop()
    assert dist.shape == ham.shape == (len(w.ham_r) * 25 - 5,)      # R = 0 contributes no i == j entries (:3624-3636)
    if "synth_dist_hop_dist" in z.files:
        assert np.allclose(dist, z["synth_dist_hop_dist"], atol=1e-12) and np.allclose(ham, z["synth_dist_hop_ham"], atol=1e-12)
    assert np.all(np.diff(w.shells()) > 0)


def test_w90_guards(tmp_path):
    """position operators refuse Wannier models unless told otherwise (pythtb.py:2028-2032, 3952-3974)."""
    m = _parse(tmp_path).model()
    with pytest.raises(Exception):
        m.position_matrix(np.zeros((2, 5), dtype=complex), 0)
    m.ignore_position_operator_offdiagonal()


@pytest.mark.gpu
def test_gpu_w90_parsed_model_eigenvalues(tmp_path):
    z = np.load(os.path.join(GOLD, "w90.npz"))
    w = _parse(tmp_path)
    ev = w.model().solve_all(z["synth_k"])
    assert np.max(np.abs(ev - z["synth_all_evals"])) <= 1e-10 * max(1.0, np.max(np.abs(ev)))
    m = w.model(**CUT)
    assert np.max(np.abs(m.solve_all(z["synth_k"]) - orc.solve_all(m, z["synth_k"]))) <= 1e-10 * max(1.0, np.max(np.abs(ev)))


def test_array_native_plan_builder_equals_the_loop(tmp_path):
    """Large models compile their plan with array operations (pythtb_b200/_plan.py): same terms, same order,
    same lattice-vector table as the per-hopping loop — spinless Wannier model, spinor models (2x2 blocks,
    complex on-site blocks), dim_k < dim_r, zero amplitudes and R = 0 hoppings included."""
    import io
    import contextlib
    import pythtb_b200
    from pythtb_b200 import _plan
    from tests import models as M
    zoo = [_parse(tmp_path).model()]
    assert len(zoo[0]._hoppings) > 500
    with contextlib.redirect_stdout(io.StringIO()):
        zoo += [M.random_model(pythtb_b200, norb=6, dim=3, nhop=60, nspin=2, seed=2),
                M.random_model(pythtb_b200, norb=9, dim=2, nhop=80, nspin=1, seed=3),
                M.kane_mele(pythtb_b200, "odd"), M.bn_ribbon(pythtb_b200, 9), M.cubic_slab(pythtb_b200, 6),
                M.haldane(pythtb_b200, 0.0)]
    zoo[1].set_hop(0.0, 0, 1, [0, 0, 0], mode="reset", allow_conjugate_pair=True)      # an explicit zero amplitude
    old = _plan.VECTORISE_FROM
    try:
        for m in zoo:
            _plan.VECTORISE_FROM = 10 ** 9
            a = _plan.compile_plan(m)
            _plan.VECTORISE_FROM = 1
            b = _plan.compile_plan(m)
            assert (a.nph, a.nel, a.nterm) == (b.nph, b.nel, b.nterm)
            for name in ("ph_R", "tau", "el_ptr", "el_row", "el_col", "t_ph", "t_amp", "pm_ptr", "pm_el", "pm_amp"):
                assert np.array_equal(getattr(a, name), getattr(b, name)), name
    finally:
        _plan.VECTORISE_FROM = old
