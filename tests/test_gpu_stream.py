"""Streamed 1-D Berry phase (BASELINE config 4) on the GPU: the chunked
assemble -> diagonalise -> link-overlap pipeline with a carried point against the materialised path and
the oracle, for every solver family, both branches, several chunk sizes (chunk boundaries, a last short
chunk, a one-link chunk), and at the config's orbital count."""
import numpy as np
import pytest

from tests import compare, models as M

pytestmark = pytest.mark.gpu
TWO_PI = 2.0 * np.pi


def _mod():
    import pythtb_b200
    return pythtb_b200


@pytest.mark.parametrize("which,nk,chunks", [
    ("ssh", 1001, (None, 100, 333)),          # n = 2: register mesh kernel, 1-D
    ("three", 257, (None, 64, 255)),          # n = 3
    ("ribbon8", 301, (None, 50, 299)),        # n = 16: tile solver, nocc = 8 (CTA-wide Wilson kernels)
    ("ribbon40", 1001, (None, 148, 999)),     # n = 80: blocked solver, nocc = 40
])
def test_streamed_string_matches_materialised_path(which, nk, chunks):
    from oracle import pythtb_oracle as orc
    mod = _mod()
    if which == "ssh":
        m = M.random_model(mod, norb=2, dim=1, nhop=3, nspin=1, seed=41)
        occ = [0]
    elif which == "three":
        m = M.three_site(mod, 0.2)
        occ = [0, 1]
    else:
        m = M.bn_ribbon(mod, int(which[6:]))
        occ = list(range(m._nsta // 2))
    start = [0.13]
    full = mod.wf_array(m, [nk])
    gaps_ref = full.solve_on_grid(start)
    ph_ref = full.berry_phase(occ)
    ev_ref = full.berry_phase(occ, berry_evals=True) if len(occ) > 1 else None
    if nk <= 301:
        wfs = np.array(full._wfs)
        assert abs(compare.circ_diff(ph_ref, orc.berry_phase(wfs, 1, occ, 0), TWO_PI)) < compare.TOL_PHASE
    ws = mod.wf_array(m, [nk], stream=True)
    gaps = ws.solve_on_grid(start)
    assert np.max(np.abs(gaps - gaps_ref)) < 1e-12
    assert abs(compare.circ_diff(ws.berry_phase(occ), ph_ref, TWO_PI)) < compare.TOL_PHASE
    for chunk in chunks:
        ph, g2 = ws.berry_phase_stream(start, occ, want_gaps=True, chunk=chunk)
        assert abs(compare.circ_diff(ph, ph_ref, TWO_PI)) < compare.TOL_PHASE, chunk
        assert np.max(np.abs(g2 - gaps_ref)) < 1e-12
        if ev_ref is not None:
            ev = ws.berry_phase_stream(start, occ, berry_evals=True, chunk=chunk)
            ok, dev = compare.sets_close(ev, ev_ref, TWO_PI, compare.TOL_PHASE)
            assert ok, (chunk, dev)
    with pytest.raises(Exception, match="stream=True"):
        ws._wfs
    with pytest.raises(Exception, match="stream=True"):
        mod.wf_array(M.haldane(mod), [5, 5], stream=True)


def test_streamed_string_at_config4_orbital_count_in_bounded_memory():
    """norb = 400 (BN ribbon cut_piece(200, 1)), 1201 k-points = 4 chunks: the streamed phase equals the
    materialised one, and the device memory the pass needs does not grow with the string length."""
    import torch
    mod = _mod()
    rib = M.bn_ribbon(mod, 200)
    occ = list(range(200))
    nk = 1201
    full = mod.wf_array(rib, [nk])
    full.solve_on_grid([0.0])
    ph_ref = full.berry_phase(occ)
    del full
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    ws = mod.wf_array(rib, [nk], stream=True)
    ph, gaps = ws.berry_phase_stream([0.0], occ, want_gaps=True)
    peak_1 = torch.cuda.max_memory_allocated() - base
    assert abs(compare.circ_diff(ph, ph_ref, TWO_PI)) < compare.TOL_PHASE
    assert gaps.shape == (399,) and np.all(gaps >= 0)
    assert gaps[199] > 0.05                                     # the ribbon is an insulator at half filling
    torch.cuda.reset_peak_memory_stats()
    ws2 = mod.wf_array(rib, [4 * nk - 3], stream=True)
    ph2 = ws2.berry_phase_stream([0.0], occ)
    peak_4 = torch.cuda.max_memory_allocated() - base
    assert peak_4 <= peak_1 * 1.05 + (1 << 20), (peak_1, peak_4)   # 4 x the string, the same memory
    assert peak_1 < 4 * (1 << 30)
    assert abs(compare.circ_diff(ph2, ph_ref, TWO_PI)) < 1e-2      # the continuum limit: the phase converges in nk
