"""One rank of the multi-GPU parity check (launched by torchrun / tests/test_gpu_multi.py):
sharded wf_array over NCCL against the unsharded array computed on the same GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pythtb_b200 as tb
    from tests import models as M, compare
    def log(*a):
        print("[rank %d]" % rank, *a, flush=True)

    try:
        cases = [(M.haldane(tb, 0.0), [0], [130, 97]), (M.kane_mele(tb, "odd"), [0, 1], [67, 70]),
                 (M.random_model(tb, norb=7, dim=2, nhop=12, nspin=1, seed=3), [0, 1, 2], [21, 11])]
        for halo in ("exchange", "recompute", "auto"):
            for model, occ, mesh in cases:
                log("case", halo, model._nsta, mesh)
                full = tb.wf_array(model, mesh)
                gaps_ref = full.solve_on_grid([-0.5, -0.5])
                w = tb.wf_array(model, mesh, shard=(rank, world), halo=halo)
                gaps = w.solve_on_grid([-0.5, -0.5])
                assert np.array_equal(gaps, gaps_ref), (halo, gaps, gaps_ref)
                sh = w._shard
                # slabs equal the rows of the unsharded array, closing row included: bit for bit when the
                # closing row is recomputed; the exchanged image of row 0 applies the two pbc phases of its
                # corner point in the other order (1 ulp)
                got, want = np.array(w._wfs), np.array(full._wfs)[sh.row0:sh.row0 + sh.nrows + 1]
                assert np.array_equal(got[:-1], want[:-1]), (halo, rank)
                if halo == "recompute" or (halo == "auto" and model._nsta <= 16):
                    assert np.array_equal(got[-1], want[-1]), (halo, rank)
                else:
                    assert np.max(np.abs(got[-1] - want[-1])) < 1e-15, (halo, rank, np.max(np.abs(got[-1] - want[-1])))
                f_ref, f = full.berry_flux(occ), w.berry_flux(occ)
                assert abs(f - f_ref) < 1e-9, (f, f_ref)
                p_ref, p = full.berry_flux(occ, individual_phases=True), w.berry_flux(occ, individual_phases=True)
                assert p.shape == p_ref.shape and np.max(np.abs(compare.circ_diff(p, p_ref, 2 * np.pi))) < 1e-10
                for d in (0, 1):
                    b_ref, b = full.berry_phase(occ, d, contin=False), w.berry_phase(occ, d, contin=False)
                    assert b.shape == b_ref.shape and np.max(np.abs(compare.circ_diff(b, b_ref, 2 * np.pi))) < 1e-9
                if len(occ) > 1 and halo == "auto":
                    # Wilson-loop spectra; d = 0 runs along the sharded axis (ordered product across ranks)
                    for d in (0, 1):
                        e_ref = full.berry_phase(occ, d, contin=False, berry_evals=True)
                        e = w.berry_phase(occ, d, contin=False, berry_evals=True)
                        assert e.shape == e_ref.shape, (e.shape, e_ref.shape)
                        ok, dev = compare.sets_close(e, e_ref, 2 * np.pi, 1e-8)
                        assert ok, (d, dev)
        # deferred reductions (tbk_peer_defer): POSTED by the producing kernel, completed by a later collective
        # kernel (a synchronous one completes everything pending; a deferred flux kernel the ones of EARLIER
        # steps) or by an explicit flush; repeated calls exercise the slot reuse and the forced flush
        for model, occ, mesh in cases[:2]:
            full = tb.wf_array(model, mesh)
            gaps_ref = full.solve_on_grid([-0.5, -0.5])
            f_ref = full.berry_flux(occ)
            w = tb.wf_array(model, mesh, shard=(rank, world), halo="recompute")
            eng = model._engine()
            for rep in range(7):
                g = w._solve_on_grid_device([-0.5, -0.5], defer_reduce=True)
                if rep % 3 == 0:
                    f = w._berry_flux_device(occ)                                    # synchronous: completes g as well
                    assert abs(float(f.cpu().reshape(-1)[0]) - f_ref) < 1e-9
                elif rep % 3 == 1:
                    eng.peer_flush()
                else:
                    g2 = w._solve_on_grid_device([-0.5, -0.5], defer_reduce=True)   # two posted reductions in flight
                    eng.peer_barrier()                                                # a barrier leaves them alone ...
                    eng.peer_flush()                                                  # ... the flush completes both
                    assert np.array_equal(g2.cpu().numpy(), gaps_ref), (rep, g2, gaps_ref)
                assert np.array_equal(g.cpu().numpy(), gaps_ref), (rep, g, gaps_ref)
            # the pipelined step of bench.py: both reductions of a step only posted; step i's flux kernel completes
            # the reductions of step i-1; one flush at the end.  33 steps: every mailbox slot is reused 8 times.
            keep = []
            for step in range(33):
                eng.peer_barrier()
                g = w._solve_on_grid_device([-0.5, -0.5], defer_reduce=True)
                f = w._berry_flux_device(occ, defer_reduce=True)
                keep.append((g, f))
                if step >= 2:                       # results of step - 2 were completed by the flux kernel of step - 1
                    g_old, f_old = keep[step - 2]
                    assert np.array_equal(g_old.cpu().numpy(), gaps_ref), (step, g_old)
                    assert abs(float(f_old.cpu().reshape(-1)[0]) - f_ref) < 1e-9, (step, f_old)
            eng.peer_flush()
            for g, f in keep[-2:]:
                assert np.array_equal(g.cpu().numpy(), gaps_ref)
                assert abs(float(f.cpu().reshape(-1)[0]) - f_ref) < 1e-9
            # many posted solves in a row without a completer: the library flushes by itself before a slot is reused
            gs = [w._solve_on_grid_device([-0.5, -0.5], defer_reduce=True) for _ in range(11)]
            eng.peer_flush()
            for g in gs:
                assert np.array_equal(g.cpu().numpy(), gaps_ref)
        # solve_on_slice on a sharded array: a fixed global row per call / every rank fills its rows
        lam = np.linspace(0.0, 1.0, 9)
        kx = np.linspace(0.0, 1.0, 6).reshape(-1, 1)
        full = tb.wf_array(M.three_site(tb, 0.0), [9, 6])
        ws = tb.wf_array(M.three_site(tb, 0.0), [9, 6], shard=(rank, world))
        for il, lm in enumerate(lam):
            mdl = M.three_site(tb, lm)
            assert np.max(np.abs(full.solve_on_slice({0: il}, kx, model=mdl) - ws.solve_on_slice({0: il}, kx, model=mdl))) < 1e-12
        sh = ws._shard
        assert np.array_equal(np.array(ws._wfs), np.array(full._wfs)[sh.row0:sh.row0 + sh.nrows + 1])
        for arr in (full, ws):
            arr.impose_loop(0)
        assert abs(compare.circ_diff(ws.berry_flux([0]), full.berry_flux([0]), 2 * np.pi)) < 1e-9
        m2 = M.haldane(tb, 0.2)
        full2, ws2 = tb.wf_array(m2, [7, 5]), tb.wf_array(m2, [7, 5], shard=(rank, world))
        kk = np.stack(np.meshgrid(np.linspace(0, 1, 7), np.linspace(0, 1, 5), indexing="ij"), axis=-1)
        for j in range(5):
            assert np.max(np.abs(full2.solve_on_slice({1: j}, kk[:, j]) - ws2.solve_on_slice({1: j}, kk[:, j]))) < 1e-12
        sh = ws2._shard
        assert np.array_equal(np.array(ws2._wfs), np.array(full2._wfs)[sh.row0:sh.row0 + sh.nrows + 1])
        # streamed 1-D string dealt to the ranks (BASELINE config 4 in miniature), both branches
        rib = M.bn_ribbon(tb, 20)
        occ_r = list(range(rib._nsta // 2))
        fullr = tb.wf_array(rib, [101])
        gaps_ref = fullr.solve_on_grid([0.05])
        wsr = tb.wf_array(rib, [101], shard=(rank, world), stream=True)
        assert np.max(np.abs(wsr.solve_on_grid([0.05]) - gaps_ref)) < 1e-12
        assert abs(compare.circ_diff(wsr.berry_phase(occ_r), fullr.berry_phase(occ_r), 2 * np.pi)) < 1e-9
        ok, dev = compare.sets_close(wsr.berry_phase_stream([0.05], occ_r, berry_evals=True, chunk=17),
                                     fullr.berry_phase(occ_r, berry_evals=True), 2 * np.pi, 1e-8)
        assert ok, dev
        # a wider occupied set: the CTA-wide Wilson-loop kernels (nocc >= 8) across ranks
        rib = M.random_model(tb, norb=20, dim=2, nhop=40, nspin=1, seed=11)
        full = tb.wf_array(rib, [17, 9])
        full.solve_on_grid([0.0, 0.0])
        w = tb.wf_array(rib, [17, 9], shard=(rank, world))
        w.solve_on_grid([0.0, 0.0])
        occ = list(range(10))
        for d in (0, 1):
            e_ref = full.berry_phase(occ, d, contin=False, berry_evals=True)
            e = w.berry_phase(occ, d, contin=False, berry_evals=True)
            ok, dev = compare.sets_close(e, e_ref, 2 * np.pi, 1e-8)
            assert e.shape == e_ref.shape and ok, (d, dev)
        print("rank %d ok" % rank, flush=True)
    except BaseException:
        # a failed rank must not leave the others waiting inside a collective: report and kill the job
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
