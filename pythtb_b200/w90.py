"""Wannier90 importer with the interface of ``pythtb.w90``
(/root/reference/pythtb.py:3208-3759).

Text parsing is host-side, one-time work (the parser is out of
scope for the GPU); only the resulting hopping list feeds the kernels.  The
parser is vectorised and ``model()`` inserts hoppings in O(nhop) while keeping
the reference's list order (R vectors in first-seen order, then i, then j;
pythtb.py:3532-3584), instead of the reference's O(nhop^2) ``set_hop`` scan.
"""
import numpy as np

from .model import tb_model

__all__ = ["w90"]


class w90(object):
    """``w90(path, prefix)``: reads prefix.win, prefix_hr.dat, prefix_centres.xyz."""

    def __init__(self, path, prefix):
        self.path = path
        self.prefix = prefix
        self.lat = self._read_win()
        self._read_hr()
        self._read_centres()

    def _full(self, suffix):
        return self.path + "/" + self.prefix + suffix

    def _read_win(self):
        """unit_cell_cart block, optional Bohr/Ang line (pythtb.py:3336-3364)."""
        with open(self._full(".win"), "r") as f:
            ln = f.readlines()
        for i, line in enumerate(ln):
            sp = line.split()
            if len(sp) >= 2 and sp[0].lower() == "begin" and sp[1].lower() == "unit_cell_cart":
                unit = ln[i + 1].strip().lower()
                pref, skip = 1.0, 0
                if unit == "bohr":
                    pref, skip = 0.5291772108, 1
                elif unit in ("ang", "angstrom"):
                    skip = 1
                lat = np.zeros((3, 3), dtype=float)
                for j in range(3):
                    sp = ln[i + skip + 1 + j].split()
                    for k in range(3):
                        lat[j, k] = float(sp[k]) * pref
                return lat
        raise Exception("Unable to find unit_cell_cart block in the .win file.")

    def _read_hr(self):
        """num_wan, Wigner-Seitz degeneracies, then rows ``R1 R2 R3 i j Re Im``
        = <0 i|H|R j> (pythtb.py:3367-3426)."""
        with open(self._full("_hr.dat"), "r") as f:
            ln = f.readlines()
        self.num_wan = int(ln[1])
        num_ws = int(ln[2])
        deg = []
        last = None
        for j in range(3, len(ln)):
            deg.extend(int(s) for s in ln[j].split())
            if len(deg) == num_ws:
                last = j
                break
            if len(deg) > num_ws:
                raise Exception("Too many degeneracies for WS points!")
        deg = np.array(deg, dtype=int)
        # one C-level parse of the whole body; the rows of one R form a block of num_wan^2 lines, but nothing
        # below relies on it: rows are grouped by R vector in first-seen order (dict order = the order the
        # reference's loop at pythtb.py:3532 visits them)
        rows = np.loadtxt(ln[last + 1:], dtype=float, ndmin=2)
        if rows.shape[1] != 7:
            raise Exception("Unexpected row format in the _hr.dat file.")
        rvec = rows[:, :3].astype(int)
        ii = rows[:, 3].astype(int) - 1
        jj = rows[:, 4].astype(int) - 1
        val = rows[:, 5] + 1.0j * rows[:, 6]
        uniq, first, inv = np.unique(rvec, axis=0, return_index=True, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        order = np.argsort(first, kind="stable")              # unique R vectors in first-seen order
        rank = np.empty(len(order), dtype=int)
        rank[order] = np.arange(len(order))
        if len(order) > num_ws:
            raise Exception("More R vectors in the file than Wigner-Seitz points announced!")
        ham = np.zeros((len(order), self.num_wan, self.num_wan), dtype=complex)
        ham[rank[inv], ii, jj] = val                          # a repeated (R, i, j) row: the last one wins, as in the reference
        self.ham_r = {}
        for pos, u in enumerate(order):
            self.ham_r[tuple(int(x) for x in uniq[u])] = {"h": ham[pos], "deg": deg[pos]}
        for R in self.ham_r:
            if R != (0, 0, 0) and tuple(-x for x in R) not in self.ham_r:
                raise Exception("Did not find negative R for R = " + str(R) + "!")

    def _read_centres(self):
        """Wannier centres (rows starting with X) -> reduced coordinates,
        NOT wrapped into the home cell (pythtb.py:3429-3445, 3925-3938)."""
        with open(self._full("_centres.xyz"), "r") as f:
            ln = f.readlines()
        xyz = []
        for i in range(2, 2 + self.num_wan):
            sp = ln[i].split()
            if sp[0] != "X":
                raise Exception("Inconsistency in the centres file.")
            xyz.append([float(sp[1]), float(sp[2]), float(sp[3])])
        self.xyz_cen = np.array(xyz, dtype=float)
        cnv = np.linalg.inv(np.array(self.lat).T)
        self.red_cen = np.array([np.dot(cnv, c) for c in self.xyz_cen])

    def model(self, zero_energy=0.0, min_hopping_norm=None, max_distance=None, ignorable_imaginary_part=None):
        """pythtb.py:3448-3586: tb_model from H(R)/deg(R), keeping one of +-R,
        with optional norm / distance / imaginary-part filters."""
        tb = tb_model(3, 3, self.lat, self.red_cen)
        tb._assume_position_operator_diagonal = False
        h0 = self.ham_r[(0, 0, 0)]
        diag = np.diagonal(h0["h"]) / float(h0["deg"])
        if np.any(np.abs(diag.imag) > 1.0e-9):
            raise Exception("Onsite terms should be real!")
        tb.set_onsite(diag.real - zero_energy)
        amps, hi, hj, hR = [], [], [], []
        nw = self.num_wan
        for R, ent in self.ham_r.items():
            first = next((x for x in R if x != 0), 0)
            if first < 0:
                continue                      # keep the lexicographically positive of +-R
            at_origin = (first == 0)
            ham = ent["h"] / float(ent["deg"])
            keep = np.ones((nw, nw), dtype=bool)
            if at_origin:
                keep &= np.triu(np.ones((nw, nw), dtype=bool), 1)   # j > i only
            if max_distance is not None:
                vecR = R[0] * self.lat[0] + R[1] * self.lat[1] + R[2] * self.lat[2]
                dvec = -self.xyz_cen[:, None, :] + self.xyz_cen[None, :, :] + vecR
                dist = np.sqrt(np.sum(dvec * dvec, axis=-1))
                keep &= ~(dist > max_distance)
            if min_hopping_norm is not None:
                keep &= ~(np.abs(ham) < min_hopping_norm)
            if ignorable_imaginary_part is not None:
                ham = np.where(np.abs(ham.imag) < ignorable_imaginary_part, ham.real + 0.0j, ham)
            idx_i, idx_j = np.nonzero(keep)     # row-major: i outer, j inner, as the reference loops
            if len(idx_i):
                amps.append(ham[idx_i, idx_j])
                hi.append(idx_i)
                hj.append(idx_j)
                hR.append(np.tile(np.array(R, dtype=int), (len(idx_i), 1)))
        if amps:
            amps, hi, hj, hR = (np.concatenate(x) for x in (amps, hi, hj, hR))
            amps = amps.tolist()                # Python complex, as set_hop would store them
        tb._bulk_set_hops(amps, hi, hj, hR)
        return tb

    def dist_hop(self):
        """Distances and |H| of all matrix elements (pythtb.py:3590-3645)."""
        ret_ham, ret_dist = [], []
        nw = self.num_wan
        offdiag = ~np.eye(nw, dtype=bool)
        for R, ent in self.ham_r.items():
            vecR = R[0] * self.lat[0] + R[1] * self.lat[1] + R[2] * self.lat[2]
            dvec = -self.xyz_cen[:, None, :] + self.xyz_cen[None, :, :] + vecR
            dist = np.sqrt(np.sum(dvec * dvec, axis=-1))
            ham = ent["h"] / float(ent["deg"])
            if R[0] == 0 and R[1] == 0 and R[2] == 0:
                # the on-site energies of the home cell are not hoppings (pythtb.py:3624-3636, avoid_diagonal)
                dist, ham = dist[offdiag], ham[offdiag]
            ret_dist.append(dist.reshape(-1))
            ret_ham.append(ham.reshape(-1))
        return (np.concatenate(ret_dist), np.concatenate(ret_ham))

    def shells(self, num_digits=2):
        """Distinct rounded hopping distances (pythtb.py:3647-3685)."""
        dist = self.dist_hop()[0]
        return np.array(sorted(set(np.round(dist, num_digits).tolist())))

    def w90_bands_consistency(self):
        """k-points and energies interpolated by Wannier90 itself
        (prefix_band.kpt / prefix_band.dat; pythtb.py:3687-3759)."""
        kpts = np.loadtxt(self._full("_band.kpt"), skiprows=1)[:, :3]
        ene = np.loadtxt(self._full("_band.dat"))[:, 1]
        return (kpts, ene.reshape((self.num_wan, kpts.shape[0])))
