"""Model zoo shared by the golden-vector generator and the parity tests.

Every builder takes the *module* that provides ``tb_model`` (either the
unmodified reference ``pythtb`` imported from /root/reference by
``tests/golden/make_golden.py``, or ``pythtb_b200``), so both sides build the
same physics through the same public calls.  Parameters follow the reference
example scripts (cited per builder); the code here is a fresh write-up.
"""
import numpy as np

_HEX_LAT = [[1.0, 0.0], [0.5, np.sqrt(3.0) / 2.0]]
_HEX_ORB = [[1.0 / 3.0, 1.0 / 3.0], [2.0 / 3.0, 2.0 / 3.0]]


def haldane(mod, delta=0.2, t=-1.0, t2=0.15):
    """Haldane model, examples/haldane.py:14-40 (delta=0.2) and
    examples/haldane_bp.py:14-41 (delta=0)."""
    m = mod.tb_model(2, 2, _HEX_LAT, _HEX_ORB)
    t2c = t2 * np.exp(1.0j * np.pi / 2.0)
    m.set_onsite([-delta, delta])
    for R in ([0, 0],):
        m.set_hop(t, 0, 1, R)
    m.set_hop(t, 1, 0, [1, 0])
    m.set_hop(t, 1, 0, [0, 1])
    m.set_hop(t2c, 0, 0, [1, 0])
    m.set_hop(t2c, 1, 1, [1, -1])
    m.set_hop(t2c, 1, 1, [0, 1])
    m.set_hop(t2c.conjugate(), 1, 1, [1, 0])
    m.set_hop(t2c.conjugate(), 0, 0, [1, -1])
    m.set_hop(t2c.conjugate(), 0, 0, [0, 1])
    return m


def kane_mele(mod, topological="odd"):
    """Kane-Mele model, examples/kane_mele.py:14-68."""
    m = mod.tb_model(2, 2, _HEX_LAT, _HEX_ORB, nspin=2)
    esite = {"even": 2.5, "odd": 1.0}[topological]
    thop = 1.0
    spin_orb = 0.6 * thop * 0.5
    rashba = 0.25 * thop
    m.set_onsite([esite, -esite])
    sx = np.array([0.0, 1.0, 0.0, 0.0])
    sy = np.array([0.0, 0.0, 1.0, 0.0])
    sz = np.array([0.0, 0.0, 0.0, 1.0])
    for R in ([0, 0], [0, -1], [-1, 0]):
        m.set_hop(thop, 0, 1, R)
    m.set_hop(-1.0j * spin_orb * sz, 0, 0, [0, 1])
    m.set_hop(1.0j * spin_orb * sz, 0, 0, [1, 0])
    m.set_hop(-1.0j * spin_orb * sz, 0, 0, [1, -1])
    m.set_hop(1.0j * spin_orb * sz, 1, 1, [0, 1])
    m.set_hop(-1.0j * spin_orb * sz, 1, 1, [1, 0])
    m.set_hop(1.0j * spin_orb * sz, 1, 1, [1, -1])
    r3h = np.sqrt(3.0) / 2.0
    m.set_hop(1.0j * rashba * (0.5 * sx - r3h * sy), 0, 1, [0, 0], mode="add")
    m.set_hop(1.0j * rashba * (-1.0 * sx), 0, 1, [0, -1], mode="add")
    m.set_hop(1.0j * rashba * (0.5 * sx + r3h * sy), 0, 1, [-1, 0], mode="add")
    return m


def graphene(mod, delta=-0.1, t=-1.0):
    """Graphene / boron nitride, examples/cone.py:14-35 (delta=-0.1),
    examples/bn_ribbon_berry.py:14-29 (delta=0.4)."""
    m = mod.tb_model(2, 2, _HEX_LAT, _HEX_ORB)
    m.set_onsite([-delta, delta])
    m.set_hop(t, 0, 1, [0, 0])
    m.set_hop(t, 1, 0, [1, 0])
    m.set_hop(t, 1, 0, [0, 1])
    return m


def bn_ribbon(mod, ncell=3, delta=0.4, t=-1.0):
    """BN ribbon, examples/bn_ribbon_berry.py:31 (cut_piece(3,1))."""
    return graphene(mod, delta, t).cut_piece(ncell, 1, glue_edgs=False)


def three_site(mod, lmbd=0.0, delta=2.0, t=-1.0):
    """3-site chain at pump parameter lambda, examples/3site_cycle.py:14-60."""
    m = mod.tb_model(1, 1, [[1.0]], [[0.0], [1.0 / 3.0], [2.0 / 3.0]])
    m.set_hop(t, 0, 1, [0])
    m.set_hop(t, 1, 2, [0])
    m.set_hop(t, 2, 0, [1])
    ons = [-delta * np.cos(2.0 * np.pi * (lmbd - j / 3.0)) for j in range(3)]
    m.set_onsite(ons, mode="reset")
    return m


def checkerboard(mod, delta=1.1, t=0.6):
    """Checkerboard, examples/checkerboard.py:14-33."""
    m = mod.tb_model(2, 2, [[1.0, 0.0], [0.0, 1.0]], [[0.0, 0.0], [0.5, 0.5]])
    m.set_onsite([-delta, delta])
    m.set_hop(t, 1, 0, [0, 0])
    m.set_hop(t, 1, 0, [1, 0])
    m.set_hop(t, 1, 0, [0, 1])
    m.set_hop(t, 1, 0, [1, 1])
    return m


def cubic_bulk(mod, delta=1.0, ta=0.4, tb=0.7):
    """3-D two-orbital cubic model, examples/cubic_slab_hwf.py:13-33."""
    m = mod.tb_model(3, 3, np.identity(3).tolist(), [[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    m.set_onsite([-delta, delta])
    for R in ([-1, 0, 0], [0, 0, -1], [-1, -1, 0], [0, -1, -1]):
        m.set_hop(ta, 0, 1, R)
    for R in ([0, 0, 0], [0, -1, 0], [-1, -1, -1], [-1, 0, -1]):
        m.set_hop(tb, 0, 1, R)
    return m


def cubic_slab(mod, nl=9, **kw):
    """Slab of the cubic model, examples/cubic_slab_hwf.py:35-41."""
    slab = cubic_bulk(mod, **kw).cut_piece(nl, 2, glue_edgs=False)
    return slab.remove_orb(2 * nl - 1)


def molecule(mod):
    """0-D three-site molecule, examples/0dim.py."""
    m = mod.tb_model(0, 2, [[1.0, 0.0], [0.0, 1.0]],
                     [[0.0, 0.0], [0.5, 0.0], [0.25, 0.5]])
    m.set_onsite([0.4, 0.0, -0.3])
    m.set_hop(-1.0, 0, 1)
    m.set_hop(-0.7 + 0.2j, 1, 2)
    m.set_hop(-0.5, 2, 0)
    return m


def random_model(mod, norb=6, dim=2, nhop=20, nspin=1, seed=0):
    """Seeded random periodic model with complex hoppings and generic orbital
    positions (exercises every term of pythtb.py:900-924)."""
    rng = np.random.RandomState(seed)
    lat = np.identity(dim) + 0.1 * rng.rand(dim, dim)
    if np.linalg.det(lat) < 0:
        lat[0] *= -1.0
    orb = rng.rand(norb, dim)
    m = mod.tb_model(dim, dim, lat.tolist(), orb.tolist(), nspin=nspin)
    if nspin == 1:
        m.set_onsite(rng.randn(norb).tolist())
    else:
        m.set_onsite([rng.randn(4).tolist() for _ in range(norb)])
    seen = set()
    while len(seen) < nhop:
        i, j = int(rng.randint(norb)), int(rng.randint(norb))
        R = tuple(int(x) for x in rng.randint(-2, 3, size=dim))
        if i == j and not any(R):
            continue
        if (i, j, R) in seen or (j, i, tuple(-x for x in R)) in seen:
            continue
        seen.add((i, j, R))
        if nspin == 1:
            amp = complex(rng.randn(), rng.randn())
        else:
            amp = (rng.randn(2, 2) + 1.0j * rng.randn(2, 2))
        m.set_hop(amp, i, j, list(R))
    return m
