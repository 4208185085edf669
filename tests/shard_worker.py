"""Worker of tests/test_sharded_gloo.py: one rank of a world_size-N gloo group.
Drives the product's host-side sharding logic (pythtb_b200.wfarray) with the
numpy oracle engine and compares every global result with the unsharded one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from tests import models as M, oracle_api as api, compare
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for halo in ("exchange", "recompute"):
            for model, occ, mesh in ((M.haldane(api, 0.0), [0], [14, 9]), (M.kane_mele(api, "odd"), [0, 1], [9, 8])):
                full = api.wf_array(model, mesh)
                gaps_ref = full.solve_on_grid([-0.5, -0.5])
                w = api.wf_array(model, mesh, shard=(rank, world), halo=halo)
                gaps = w.solve_on_grid([-0.5, -0.5])
                assert np.max(np.abs(gaps - gaps_ref)) < 1e-12, (gaps, gaps_ref)
                sh = w._shard
                # the local slab equals the corresponding rows of the unsharded array
                assert np.max(np.abs(w._wfs - full._wfs[sh.row0:sh.row0 + sh.nrows + 1])) < 1e-12
                for dirs in (None, [1, 0]):
                    f_ref = full.berry_flux(occ, dirs)
                    f = w.berry_flux(occ, dirs)
                    assert abs(f - f_ref) < 1e-10, (f, f_ref)
                    p_ref = full.berry_flux(occ, dirs, individual_phases=True)
                    p = w.berry_flux(occ, dirs, individual_phases=True)
                    assert p.shape == p_ref.shape
                    assert np.max(np.abs(compare.circ_diff(p, p_ref, 2 * np.pi))) < 1e-10
                for d in (0, 1):
                    b_ref = full.berry_phase(occ, d, contin=False)
                    b = w.berry_phase(occ, d, contin=False)
                    assert b.shape == b_ref.shape, (b.shape, b_ref.shape)
                    assert np.max(np.abs(compare.circ_diff(b, b_ref, 2 * np.pi))) < 1e-10
                bc_ref = full.berry_phase(occ, 1, contin=True)
                bc = w.berry_phase(occ, 1, contin=True)
                assert np.max(np.abs(compare.circ_diff(bc, bc_ref, 2 * np.pi))) < 1e-10
                if len(occ) > 1:
                    e_ref = full.berry_phase(occ, 1, contin=False, berry_evals=True)
                    e = w.berry_phase(occ, 1, contin=False, berry_evals=True)
                    assert np.max(np.abs(compare.circ_diff(e, e_ref, 2 * np.pi))) < 1e-10
                    # strings along the sharded axis: ordered product of the per-rank products
                    e0_ref = full.berry_phase(occ, 0, contin=False, berry_evals=True)
                    e0 = w.berry_phase(occ, 0, contin=False, berry_evals=True)
                    assert e0.shape == e0_ref.shape, (e0.shape, e0_ref.shape)
                    ok, dev = compare.sets_close(e0, e0_ref, 2 * np.pi, 1e-9)
                    assert ok, dev
        # impose_pbc / impose_loop along the sharded axis on a manually filled array: a ring shift, not a local copy
        mh = M.haldane(api, 0.0)
        full = api.wf_array(mh, [9, 6])
        full.solve_on_grid([-0.5, -0.5])
        fa = api.wf_array(mh, [9, 6])
        fa._wfs[...] = full._wfs
        fa._wfs[-1] = 0.0
        fa.impose_pbc(0, 0)
        ws = api.wf_array(mh, [9, 6], shard=(rank, world))
        sh = ws._shard
        ws._wfs[...] = full._wfs[sh.row0:sh.row0 + sh.nrows + 1]
        ws._wfs[-1] = 0.0
        ws.impose_pbc(0, 0)
        assert np.max(np.abs(ws._wfs - fa._wfs[sh.row0:sh.row0 + sh.nrows + 1])) < 1e-14
        ws._wfs[-1] = 0.0
        ws.impose_loop(0)
        want = full._wfs[0] if sh.is_last else full._wfs[sh.row0 + sh.nrows]
        assert np.max(np.abs(ws._wfs[-1] - want)) < 1e-14
        # solve_on_slice on a sharded array (parametric axis along the sharded direction and across it)
        lam = np.linspace(0.0, 1.0, 9)
        kx = np.linspace(0.0, 1.0, 6)
        full = api.wf_array(M.three_site(api, 0.0), [9, 6])
        ws = api.wf_array(M.three_site(api, 0.0), [9, 6], shard=(rank, world))
        for il, lm in enumerate(lam):                                   # axis 0 fixed: one global row per call
            mdl = M.three_site(api, lm)
            e_f = full.solve_on_slice({0: il}, kx.reshape(-1, 1), model=mdl)
            e_s = ws.solve_on_slice({0: il}, kx.reshape(-1, 1), model=mdl)
            assert np.max(np.abs(e_f - e_s)) < 1e-12
        sh = ws._shard
        assert np.max(np.abs(np.abs(ws._wfs) - np.abs(full._wfs[sh.row0:sh.row0 + sh.nrows + 1]))) < 1e-12
        m2 = M.haldane(api, 0.2)
        full2 = api.wf_array(m2, [7, 5])
        ws2 = api.wf_array(m2, [7, 5], shard=(rank, world))
        kk = np.stack(np.meshgrid(np.linspace(0, 1, 7), np.linspace(0, 1, 5), indexing="ij"), axis=-1)
        for j in range(5):                                              # axis 0 free: every rank fills its rows
            e_f = full2.solve_on_slice({1: j}, kk[:, j])
            e_s = ws2.solve_on_slice({1: j}, kk[:, j])
            assert e_s.shape == e_f.shape and np.max(np.abs(e_f - e_s)) < 1e-12
        sh = ws2._shard
        assert np.max(np.abs(np.abs(ws2._wfs) - np.abs(full2._wfs[sh.row0:sh.row0 + sh.nrows + 1]))) < 1e-12
        # streamed 1-D string (BASELINE config 4 in miniature): links dealt to the ranks, never materialised
        rib = M.bn_ribbon(api, 5)
        occ_r = list(range(rib._nsta // 2))
        full = api.wf_array(rib, [23])
        gaps_ref = full.solve_on_grid([0.05])
        for kw in (dict(stream=True), dict()):
            ws = api.wf_array(rib, [23], shard=(rank, world), **kw)
            if kw:
                assert np.max(np.abs(ws.solve_on_grid([0.05]) - gaps_ref)) < 1e-12
                ph = ws.berry_phase(occ_r)
                ev = ws.berry_phase(occ_r, berry_evals=True)
            else:
                ph, gp = ws.berry_phase_stream([0.05], occ_r, want_gaps=True)
                assert np.max(np.abs(gp - gaps_ref)) < 1e-12
                ev = ws.berry_phase_stream([0.05], occ_r, berry_evals=True)
            assert abs(compare.circ_diff(ph, full.berry_phase(occ_r), 2 * np.pi)) < 1e-10
            ok, dev = compare.sets_close(ev, full.berry_phase(occ_r, berry_evals=True), 2 * np.pi, 1e-9)
            assert ok, dev
        # Convention II (tb_model.set_convention): the closing row of the last rank is a plain copy of row 0
        m2 = M.haldane(api, 0.0)
        m2.set_convention(2)
        full = api.wf_array(m2, [11, 7])
        gaps_ref = full.solve_on_grid([-0.5, -0.5])
        for halo in ("exchange", "recompute"):
            w = api.wf_array(m2, [11, 7], shard=(rank, world), halo=halo)
            assert np.max(np.abs(w.solve_on_grid([-0.5, -0.5]) - gaps_ref)) < 1e-12
            sh = w._shard
            assert np.max(np.abs(w._wfs - full._wfs[sh.row0:sh.row0 + sh.nrows + 1])) < 1e-12
            if sh.is_last:
                assert np.array_equal(w._wfs[-1], full._wfs[0])
            assert abs(w.berry_flux([0]) - full.berry_flux([0])) < 1e-10
            b, b_ref = w.berry_phase([0], 0, contin=False), full.berry_phase([0], 0, contin=False)
            assert np.max(np.abs(compare.circ_diff(b, b_ref, 2 * np.pi))) < 1e-10
        # 3-D mesh: axis 0 sharded, flux on planes that do / do not contain it
        m3 = M.random_model(api, norb=2, dim=3, nhop=6, nspin=1, seed=5)
        full = api.wf_array(m3, [7, 5, 6])
        full.solve_on_grid([0.0, 0.1, 0.2])
        w = api.wf_array(m3, [7, 5, 6], shard=(rank, world), halo="exchange")
        w.solve_on_grid([0.0, 0.1, 0.2])
        for dirs in ([0, 1], [1, 2], [2, 0]):
            for ind in (False, True):
                a, b = full.berry_flux([0], dirs, ind), w.berry_flux([0], dirs, ind)
                assert np.shape(a) == np.shape(b), (dirs, ind, np.shape(a), np.shape(b))
                assert np.max(np.abs(compare.circ_diff(np.asarray(b), np.asarray(a), 2 * np.pi))) < 1e-10
        print("rank %d ok" % rank)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
