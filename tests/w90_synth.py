"""Writer for a small synthetic Wannier90 data set (``prefix.win``,
``prefix_hr.dat``, ``prefix_centres.xyz``) in the formats the reference's
``w90`` class consumes (pythtb.py:3336-3445).

Used to exercise the Wannier90 importer without shipping third-party data:
random Hermitian H(R) with exponential decay on a +-R-symmetric star of
lattice vectors, with Wigner-Seitz degeneracies > 1 on some of them.
"""
import os
import numpy as np


def write(path, prefix, num_wan=5, seed=0, rmax=2, bohr=True):
    rng = np.random.RandomState(seed)
    lat = np.array([[3.1, 0.2, 0.0], [0.1, 2.9, 0.3], [0.0, 0.4, 3.3]])
    with open(os.path.join(path, prefix + ".win"), "w") as f:
        f.write("num_wann = %d\n\nBegin Unit_Cell_Cart\n" % num_wan)
        scale = 1.0
        if bohr:
            f.write("Bohr\n")
            scale = 1.0 / 0.5291772108
        for row in lat:
            f.write("  %.10f  %.10f  %.10f\n" % tuple(row * scale))
        f.write("End Unit_Cell_Cart\n")
    # R star: all |R_i| <= rmax with |R|_1 <= rmax+1, first-seen order mixes signs
    rvecs = [(a, b, c) for a in range(-rmax, rmax + 1) for b in range(-rmax, rmax + 1)
             for c in range(-1, 2) if abs(a) + abs(b) + abs(c) <= rmax + 1]
    rng.shuffle(rvecs)
    rvecs = [tuple(int(x) for x in r) for r in rvecs]
    hr = {}
    deg = {}
    for r in rvecs:
        neg = tuple(-x for x in r)
        if neg in hr:
            hr[r] = hr[neg].conj().T
            deg[r] = deg[neg]
            continue
        mat = (rng.randn(num_wan, num_wan) + 1.0j * rng.randn(num_wan, num_wan))
        mat *= np.exp(-1.2 * np.sqrt(sum(x * x for x in r)))
        # sprinkle tiny imaginary parts so ignorable_imaginary_part has work to do
        mask = rng.rand(num_wan, num_wan) < 0.4
        mat = np.where(mask, mat.real + 0.01j * rng.randn(num_wan, num_wan), mat)
        if r == (0, 0, 0):
            mat = 0.5 * (mat + mat.conj().T)
        hr[r] = mat
        deg[r] = 1 if r == (0, 0, 0) else int(rng.choice([1, 1, 2, 4]))
    with open(os.path.join(path, prefix + "_hr.dat"), "w") as f:
        f.write(" synthetic data written by tests/w90_synth.py\n")
        f.write("%12d\n%12d\n" % (num_wan, len(rvecs)))
        degs = [deg[r] for r in rvecs]
        for n in range(0, len(degs), 15):
            f.write("".join("%5d" % d for d in degs[n:n + 15]) + "\n")
        for r in rvecs:
            for j in range(num_wan):
                for i in range(num_wan):
                    v = hr[r][i, j]
                    f.write("%5d%5d%5d%5d%5d%16.10f%16.10f\n" % (r + (i + 1, j + 1, v.real, v.imag)))
    cen = rng.rand(num_wan, 3) @ lat
    with open(os.path.join(path, prefix + "_centres.xyz"), "w") as f:
        f.write("%6d\n synthetic centres\n" % (num_wan + 1))
        for c in cen:
            f.write("X   %16.10f %16.10f %16.10f\n" % tuple(c))
        f.write("Si  0.0 0.0 0.0\n")
