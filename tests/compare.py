"""Tolerances of the parity contract (BASELINE.json north_star):
eigenvalues |dE| <= 1e-10 * max(1, max|E|); gauge-invariant projectors 1e-9;
Berry phases / fluxes 1e-8 rad modulo 2 pi; H(k) elementwise 1e-12."""
import numpy as np

TWO_PI = 2.0 * np.pi
TOL_EVAL = 1.0e-10
TOL_PROJ = 1.0e-9
TOL_PHASE = 1.0e-8
TOL_HAM = 1.0e-12

# keys holding phases divided by 2 pi (Wannier centres)
_TURNS = ("wan_cent_even", "wan_cent_odd", "px", "wann_center")
# keys holding SORTED sets of phases per string (Wilson-loop spectra)
_SETS = ("wan_cent", "wilson")
_PHASE_WORDS = ("phi", "phase", "flux", "plaq", "bphase", "bflux", "final", "px", "wan_cent", "wann_center", "wilson")
_EVAL_WORDS = ("evals", "gaps", "checkerboard", "molecule", "three_site_one", "haldane_one", "haldane_fin", "hwfc",
               "pos_trace", "pos_exp")


def circ_diff(a, b, period):
    return (np.asarray(a) - np.asarray(b) + 0.5 * period) % period - 0.5 * period


def sets_close(a, b, period, tol):
    """Rows of a, b are sets of phases on a circle of the given period."""
    a = np.asarray(a, dtype=float).reshape(-1, np.shape(a)[-1])
    b = np.asarray(b, dtype=float).reshape(-1, np.shape(b)[-1])
    worst = 0.0
    for ra, rb in zip(a, b):
        sa = np.sort(ra % period)
        gaps = np.diff(np.append(sa, sa[0] + period))
        cut = sa[np.argmax(gaps)] + 0.5 * np.max(gaps)      # middle of the largest empty arc
        ua = np.sort((ra - cut) % period)
        ub = np.sort((rb - cut) % period)
        worst = max(worst, float(np.max(np.abs(ua - ub))))
    return worst <= tol, worst


def check(key, got, want):
    """Return (ok, measure, rule) for one named output."""
    got = np.asarray(got)
    want = np.asarray(want)
    if want.size == 0:
        return got.size == 0 or got is None, 0.0, "empty"
    assert got.shape == want.shape or got.size == want.size, (key, got.shape, want.shape)
    got = got.reshape(want.shape)
    if key.startswith("ham_"):
        dev = float(np.max(np.abs(got - want)))
        return dev <= TOL_HAM * max(1.0, float(np.max(np.abs(want)))), dev, "ham"
    if key.startswith("proj") or key.endswith("_proj") or key == "pos_op":
        dev = float(np.max(np.abs(got - want)))
        return dev <= TOL_PROJ, dev, "projector"
    if any(w in key for w in _PHASE_WORDS):
        turns = any(key.startswith(t) for t in _TURNS)
        period = 1.0 if turns else TWO_PI
        tol = TOL_PHASE / TWO_PI if turns else TOL_PHASE
        if any(w in key for w in _SETS) and want.ndim >= 1 and want.shape[-1] > 1 and "contin" not in key:
            ok, dev = sets_close(got, want, period, tol)
            return ok, dev, "phase-set mod period"
        if "contin" in key:
            ok, dev = sets_close(got, want, period, tol)
            return ok, dev, "phase-set mod period (continuity branch)"
        dev = float(np.max(np.abs(circ_diff(got, want, period))))
        return dev <= tol, dev, "phase mod period"
    if any(w in key for w in _EVAL_WORDS):
        dev = float(np.max(np.abs(got - want)))
        return dev <= TOL_EVAL * max(1.0, float(np.max(np.abs(want)))), dev, "eigenvalue"
    dev = float(np.max(np.abs(got - want)))
    return dev <= 1.0e-9 * max(1.0, float(np.max(np.abs(want)))), dev, "generic"


def compare_case(name, got, want):
    """Compare a dict of outputs with the fixture; returns list of failures."""
    bad = []
    for key in want.files if hasattr(want, "files") else want.keys():
        if key.startswith("k_"):
            continue
        ok, dev, rule = check(key, got[key], want[key])
        if not ok:
            bad.append("%s.%s: deviation %.3e (%s)" % (name, key, dev, rule))
    return bad
