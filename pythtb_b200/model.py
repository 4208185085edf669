"""Host-side tight-binding model with the PythTB 1.8.0 ``tb_model`` interface
(/root/reference/pythtb.py:29-2279).

Model definition and model surgery are host bookkeeping (out of
scope for the GPU) and are re-provided here with the reference's call
signatures, argument meaning and error behaviour, because they define the
input of the hot path.  Everything numerical — ``solve_all``/``solve_one``,
``_gen_ham``/``_sol_ham`` and the position-operator routines — is delegated to
the engine (``pythtb_b200._engine``), i.e. to the sm_100a kernels behind the
C-ABI.  There is no CPU fallback.

Hoppings are kept in the reference's list format ``[amp, i, j, R]`` (attribute
``_hoppings``, pythtb.py:475-478) plus a dictionary index so that ``set_hop``
is O(1) instead of the reference's O(nhop) scan (pythtb.py:450-493).
"""
import copy

import numpy as np

__all__ = ["tb_model"]


class _LazyAttr(object):
    """Data descriptor: an attribute of a lazily reduced model that triggers the materialisation when read."""

    def __init__(self, name):
        self.slot = "_lazy_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        obj._materialise()
        return obj.__dict__[self.slot]

    def __set__(self, obj, value):
        obj.__dict__[self.slot] = value


def _is_int(a):
    return np.issubdtype(type(a), np.integer)


def _offdiag_approximation_warning_and_stop():
    raise Exception("""

----------------------------------------------------------------------

  It looks like you are trying to calculate Berry-like object that
  involves position operator.  However, you are using a tight-binding
  model that was generated from Wannier90.  This procedure introduces
  approximation as it ignores off-diagonal elements of the position
  operator in the Wannier basis.

  If you know what you are doing and wish to continue with the
  calculation despite this approximation, please call the following
  function on your tb_model object

    my_model.ignore_position_operator_offdiagonal()

----------------------------------------------------------------------

""")


class KMesh(object):
    """A uniform k-mesh kept as its descriptor (``tb_model.k_uniform_mesh(mesh, lazy=True)``):
    ``solve_all`` generates the points on the device instead of receiving a host
    list (256^3 points are 403 MB of k-vectors).  ``np.asarray(kmesh)`` gives
    exactly the array the reference's ``k_uniform_mesh`` returns (pythtb.py:1848-1857)."""

    def __init__(self, mesh_size):
        self.mesh_size = tuple(int(n) for n in mesh_size)
        self.shape = (int(np.prod(self.mesh_size)), len(self.mesh_size))
        self.ndim = 2

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        axes = [np.arange(n) / float(n) for n in self.mesh_size]
        grids = np.meshgrid(*axes, indexing="ij")
        out = np.stack([g.reshape(-1) for g in grids], axis=-1)
        return out if dtype is None else out.astype(dtype)


class tb_model(object):
    """Tight-binding model; same constructor as pythtb.tb_model
    (pythtb.py:94-185): ``tb_model(dim_k, dim_r, lat=None, orb=None, per=None, nspin=1)``."""

    # numerical backend; the product engine is resolved lazily so that model
    # building works without a GPU.  tests/oracle_api.py swaps this attribute.
    _engine_factory = None

    def __init__(self, dim_k, dim_r, lat=None, orb=None, per=None, nspin=1):
        if not _is_int(dim_k):
            raise Exception("\n\nArgument dim_k not an integer")
        if dim_k < 0 or dim_k > 4:
            raise Exception("\n\nArgument dim_k out of range. Must be between 0 and 4.")
        if not _is_int(dim_r):
            raise Exception("\n\nArgument dim_r not an integer")
        if dim_r < dim_k or dim_r > 4:
            raise Exception("\n\nArgument dim_r out of range. Must be dim_r>=dim_k and dim_r<=4.")
        self._dim_k, self._dim_r = dim_k, dim_r
        if lat is None or (isinstance(lat, str) and lat == "unit"):
            self._lat = np.identity(dim_r, float)
            print(" Lattice vectors not specified! I will use identity matrix.")
        else:
            self._lat = np.array(lat, dtype=float)
            if self._lat.shape != (dim_r, dim_r):
                raise Exception("\n\nWrong lat array dimensions")
        if dim_r > 0:
            vol = np.linalg.det(self._lat)
            if np.abs(vol) < 1.0e-6:
                raise Exception("\n\nLattice vectors length/area/volume too close to zero, or zero.")
            if vol < 0.0:
                raise Exception("\n\nLattice vectors need to form right handed system.")
        if orb is None or (isinstance(orb, str) and orb == "bravais"):
            self._norb = 1
            self._orb = np.zeros((1, dim_r))
            print(" Orbital positions not specified. I will assume a single orbital at the origin.")
        elif _is_int(orb):
            self._norb = orb
            self._orb = np.zeros((orb, dim_r))
            print(" Orbital positions not specified. I will assume ", orb, " orbitals at the origin")
        else:
            self._orb = np.array(orb, dtype=float)
            if self._orb.ndim != 2:
                raise Exception("\n\nWrong orb array rank")
            self._norb = self._orb.shape[0]
            if self._orb.shape[1] != dim_r:
                raise Exception("\n\nWrong orb array dimensions")
        if per is None:
            self._per = list(range(dim_k))
        else:
            if len(per) != dim_k:
                raise Exception("\n\nWrong choice of periodic/infinite direction!")
            self._per = per
        if nspin not in [1, 2]:
            raise Exception("\n\nWrong value of nspin, must be 1 or 2!")
        self._nspin = nspin
        self._assume_position_operator_diagonal = True
        self._convention = 1
        self._nsta = self._norb * self._nspin
        if nspin == 1:
            self._site_energies = np.zeros(self._norb, dtype=float)
        else:
            self._site_energies = np.zeros((self._norb, 2, 2), dtype=complex)
        self._site_energies_specified = np.zeros(self._norb, dtype=bool)
        self._hoppings = []
        self._hop_index = {}
        self._plan_cache = None

    # ------------------------------------------------------------------ helpers
    def _touch(self):
        self._plan_cache = None
        self.__dict__["_version"] = self.__dict__.get("_version", 0) + 1     # witnessed by lazily reduced views

    def _rper(self, ind_R):
        if self._dim_k == 0:
            return ()
        arr = np.array(ind_R)
        return tuple(int(arr[p]) for p in self._per)

    def _reindex(self):
        self._hop_index = {}
        for n, h in enumerate(self._hoppings):
            key = (h[1], h[2], self._rper(h[3]) if self._dim_k > 0 else ())
            self._hop_index[key] = n

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_plan_cache":
                new._plan_cache = None
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        return new

    def _engine(self):
        fac = type(self)._engine_factory
        if fac is None:
            from . import _engine
            fac = _engine.get_engine
        return fac()

    def _val_to_block(self, val):
        """Scalar / (I,sx,sy,sz) 4-vector / 2x2 matrix -> 2x2 block for
        spinors (pythtb.py:517-560); identity for nspin=1."""
        if self._nspin == 1:
            return val
        use = np.array(val)
        if use.shape == (2, 2):
            return use
        ret = np.zeros((2, 2), dtype=complex)
        if use.shape == ():
            ret[0, 0] += use
            ret[1, 1] += use
        elif use.shape == (4,):
            ret[0, 0] += use[0] + use[3]
            ret[1, 1] += use[0] - use[3]
            ret[0, 1] += use[1] - 1.0j * use[2]
            ret[1, 0] += use[1] + 1.0j * use[2]
        else:
            raise Exception("\n\nWrong format of the on-site or hopping term. Must be single number, or\n"
                            "in the case of a spinfull model can be array of four numbers or 2x2\nmatrix.")
        return ret

    # ---------------------------------------------------------- model definition
    def set_onsite(self, onsite_en, ind_i=None, mode="set"):
        """pythtb.py:186-306."""
        if ind_i is None:
            if len(onsite_en) != self._norb:
                raise Exception("\n\nWrong number of site energies")
            to_check = onsite_en
        else:
            if ind_i < 0 or ind_i >= self._norb:
                raise Exception("\n\nIndex ind_i out of scope.")
            to_check = [onsite_en]
        for ons in to_check:
            a = np.array(ons)
            if a.shape == ():
                if np.abs(a - a.conjugate()) > 1.0e-8:
                    raise Exception("\n\nOnsite energy should not have imaginary part!")
            elif a.shape == (4,):
                if np.max(np.abs(a - a.conjugate())) > 1.0e-8:
                    raise Exception("\n\nOnsite energy or Zeeman field should not have imaginary part!")
            elif a.shape == (2, 2):
                if np.max(np.abs(a - a.T.conjugate())) > 1.0e-8:
                    raise Exception("\n\nOnsite matrix should be Hermitian!")
        m = mode.lower()
        if m not in ("set", "reset", "add"):
            raise Exception("\n\nWrong value of mode parameter")
        sites = range(self._norb) if ind_i is None else [ind_i]
        if m == "set":
            if ind_i is None:
                if self._site_energies_specified.any():
                    raise Exception("\n\nSome or all onsite energies were already specified! Use mode=\"reset\" or mode=\"add\".")
            elif self._site_energies_specified[ind_i]:
                raise Exception("\n\nOnsite energy for this site was already specified! Use mode=\"reset\" or mode=\"add\".")
        for i in sites:
            blk = self._val_to_block(onsite_en if ind_i is not None else onsite_en[i])
            if m == "add":
                self._site_energies[i] += blk
            else:
                self._site_energies[i] = blk
            self._site_energies_specified[i] = True
        self._touch()

    def set_hop(self, hop_amp, ind_i, ind_j, ind_R=None, mode="set", allow_conjugate_pair=False):
        """pythtb.py:308-515.  <i,0|H|j,R> = hop_amp; the conjugate partner is implied."""
        if self._dim_k != 0 and (ind_R is None):
            raise Exception("\n\nNeed to specify ind_R!")
        if self._dim_k == 1 and _is_int(ind_R):
            tmp = np.zeros(self._dim_r, dtype=int)
            tmp[self._per] = ind_R
            ind_R = tmp
        if self._dim_k != 0 and len(ind_R) != self._dim_r:
            raise Exception("\n\nLength of input ind_R vector must equal dim_r! Even if dim_k<dim_r.")
        if ind_i < 0 or ind_i >= self._norb:
            raise Exception("\n\nIndex ind_i out of scope.")
        if ind_j < 0 or ind_j >= self._norb:
            raise Exception("\n\nIndex ind_j out of scope.")
        rper = self._rper(ind_R)
        if ind_i == ind_j and not any(rper):
            raise Exception("\n\nDo not use set_hop for onsite terms. Use set_onsite instead!")
        ind_i, ind_j = int(ind_i), int(ind_j)
        if not allow_conjugate_pair:
            if (ind_j, ind_i, tuple(-x for x in rper)) in self._hop_index:
                raise Exception("\n\nFollowing matrix element was already implicitely specified:\n"
                                "   i=" + str(ind_i) + " j=" + str(ind_j) +
                                ("" if self._dim_k == 0 else " R=" + str(ind_R)) +
                                "\nRemember,specifying <i|H|j+R> automatically specifies <j|H|i-R>.  For\n"
                                "consistency, specify all hoppings for a given bond in the same\n"
                                "direction.  (Or, alternatively, see the documentation on the\n"
                                "'allow_conjugate_pair' flag.)\n")
        amp = self._val_to_block(hop_amp)
        new_hop = [amp, ind_i, ind_j] if self._dim_k == 0 else [amp, ind_i, ind_j, np.array(ind_R)]
        key = (ind_i, ind_j, rper)
        where = self._hop_index.get(key)
        m = mode.lower()
        if m == "set":
            if where is not None:
                raise Exception("\n\nHopping energy for this site was already specified! Use mode=\"reset\" or mode=\"add\".")
        elif m not in ("reset", "add"):
            raise Exception("\n\nWrong value of mode parameter")
        if where is None:
            self._hop_index[key] = len(self._hoppings)
            self._hoppings.append(new_hop)
        elif m == "reset":
            self._hoppings[where] = new_hop
        else:
            self._hoppings[where][0] = self._hoppings[where][0] + new_hop[0]
        self._touch()

    def _bulk_set_hops(self, amps, ii, jj, RR):
        """O(nhop) insertion of hoppings known to be distinct (used by w90.model)."""
        n = len(ii)
        if n == 0:
            return
        ii = np.asarray(ii, dtype=int).tolist()
        jj = np.asarray(jj, dtype=int).tolist()
        RR = np.asarray(RR, dtype=int).reshape(n, -1)
        keys = RR[:, list(self._per)].tolist() if self._dim_k > 0 else [[]] * n      # _rper of every row at once
        base = len(self._hoppings)
        for pos in range(n):
            self._hop_index[(ii[pos], jj[pos], tuple(keys[pos]))] = base + pos
            self._hoppings.append([self._val_to_block(amps[pos]), ii[pos], jj[pos], RR[pos].copy()])
        self._touch()

    # ------------------------------------------------------------------ queries
    def get_num_orbitals(self):
        return self._norb

    def get_orb(self):
        return self._orb.copy()

    def get_lat(self):
        return self._lat.copy()

    def display(self):
        """Text report (pythtb.py:562-634); presentation only."""
        print("---------------------------------------")
        print("report of tight-binding model")
        print("---------------------------------------")
        print("k-space dimension           =", self._dim_k)
        print("r-space dimension           =", self._dim_r)
        print("number of spin components   =", self._nspin)
        print("periodic directions         =", self._per)
        print("number of orbitals          =", self._norb)
        print("number of electronic states =", self._nsta)
        print("lattice vectors:")
        for i, o in enumerate(self._lat):
            print(" #", i, " ===>  ", np.round(o, 4))
        print("positions of orbitals:")
        for i, o in enumerate(self._orb):
            print(" #", i, " ===>  ", np.round(o, 4))
        print("site energies:")
        for i, site in enumerate(self._site_energies):
            print(" #", i, " ===>  ", np.round(site, 4))
        print("hoppings:")
        for h in self._hoppings:
            tail = "" if self._dim_k == 0 else " + " + str(list(h[3]))
            print("<", h[1], "| H |", h[2], tail, ">     ===>  ", np.round(h[0], 4))
        print()

    def visualize(self, *a, **k):
        raise NotImplementedError("visualize() is presentation code outside the B200 hot path "
                                  "(SURVEY.md §2: out of scope); use the reference for plotting")

    # ------------------------------------------------------------- the hot path
    def set_convention(self, convention):
        """Extension (not in PythTB 1.8.0, which implements Convention I only):
        choose the Bloch-phase convention of doc/formalism/pythtb-formalism.tex.
        1 = H_ij(k) = sum_R e^{ik.(R+tau_j-tau_i)} H_ij(R)   (tex:300-304, pythtb.py:912-916),
        2 = H~_ij(k) = sum_R e^{ik.R} H_ij(R)                 (tex:341-344).
        Eigenvalues are the same; eigenvectors are related by C~_j = e^{ik.tau_j} C_j
        (tex:355-364) and are periodic in k, so ``wf_array`` closes a periodic
        mesh direction with a plain copy instead of the e^{-iG.tau_j} factor of
        tex:688-691 / pythtb.py:2729."""
        if convention not in (1, 2):
            raise Exception("\n\nconvention must be 1 (PythTB, orbital positions in the phase) or 2")
        if convention != self._convention:
            self._convention = convention
            self._touch()

    def get_convention(self):
        return self._convention

    def _plan(self, convention=None):
        from ._plan import compile_plan
        if convention is None:
            convention = self._convention
        if self._plan_cache is None or self._plan_cache[0] != convention:
            self._plan_cache = (convention, compile_plan(self, convention))
        return self._plan_cache[1]

    def _check_k(self, k_input):
        if k_input is None:
            if self._dim_k != 0:
                raise Exception("\n\nHave to provide a k-vector!")
            return None
        kpnt = np.array(k_input, dtype=float)
        if kpnt.ndim == 0:
            kpnt = kpnt.reshape(1)
        if kpnt.shape != (self._dim_k,):
            raise Exception("\n\nk-vector of wrong shape!")
        return kpnt

    def _gen_ham(self, k_input=None):
        """H(k), pythtb.py:874-925, assembled on the GPU.  Returns
        ``[norb,norb]`` or ``[norb,2,norb,2]`` complex."""
        kpnt = self._check_k(k_input)
        klist = np.zeros((1, 0)) if kpnt is None else kpnt.reshape(1, -1)
        ham = self._engine().gen_ham(self, klist)[0]
        if self._nspin == 2:
            ham = ham.reshape(self._norb, 2, self._norb, 2)
        return ham

    def _sol_ham(self, ham, eig_vectors=False):
        """pythtb.py:927-953 through the batched GPU eigensolver."""
        ham_use = np.asarray(ham, dtype=complex).reshape(self._nsta, self._nsta)
        if np.max(ham_use - ham_use.T.conj()) > 1.0e-9:
            raise Exception("\n\nHamiltonian matrix is not hermitian?!")
        ev, vec = self._engine().eigh(ham_use[None], eig_vectors)
        if not eig_vectors:
            return ev[0]
        out = vec[0]
        if self._nspin == 2:
            out = out.reshape(self._nsta, self._norb, 2)
        return ev[0], out

    def solve_all(self, k_list=None, eig_vectors=False, device_result=False):
        """pythtb.py:955-1079: ``eval[band,k]`` and ``evec[band,k,orb(,spin)]``
        (k axis dropped for dim_k == 0).  One fused assemble+diagonalise launch
        for the whole list.  Extensions: ``k_list`` may be a ``KMesh`` descriptor
        (``k_uniform_mesh(mesh, lazy=True)``; the points are generated on the device),
        and with it ``device_result=True`` returns torch tensors instead of host arrays."""
        if isinstance(k_list, KMesh):
            if self._dim_k == 0 or k_list.shape[1] != self._dim_k:
                raise Exception("\n\nk-vector of wrong shape!")
            eng = self._engine()
            if hasattr(eng, "solve_all_mesh"):
                return eng.solve_all_mesh(self, k_list.mesh_size, eig_vectors, device_result)
            k_list = np.asarray(k_list)
        if device_result:
            raise Exception("\n\ndevice_result needs a KMesh descriptor (k_uniform_mesh(mesh, lazy=True))")
        if k_list is None:
            if self._dim_k != 0:
                raise Exception("\n\nHave to provide a k-vector!")
            kl = np.zeros((1, 0))
        else:
            if self._dim_k == 0:
                raise Exception("\n\nk-vector of wrong shape!")
            kl = np.array(k_list, dtype=float)
            if kl.ndim == 1 and self._dim_k == 1:
                kl = kl.reshape(-1, 1)
            if kl.ndim != 2 or kl.shape[1] != self._dim_k:
                raise Exception("\n\nk-vector of wrong shape!")
        res = self._engine().solve_all(self, kl, eig_vectors)
        if k_list is None:
            if not eig_vectors:
                return res[:, 0]
            ev, vec = res
            return ev[:, 0], vec[:, 0]
        return res

    def solve_one(self, k_point=None, eig_vectors=False):
        """pythtb.py:1081-1103."""
        if k_point is None:
            return self.solve_all(eig_vectors=eig_vectors)
        if not eig_vectors:
            return self.solve_all([k_point], eig_vectors=False)[:, 0]
        ev, vec = self.solve_all([k_point], eig_vectors=True)
        return ev[:, 0], vec[:, 0]

    # ------------------------------------------------------------ model surgery
    def cut_piece(self, num, fin_dir, glue_edgs=False):
        """pythtb.py:1105-1231: repeat the cell ``num`` times along ``fin_dir``
        and drop that periodicity; orbital i of cell n gets index i+norb*n."""
        if self._dim_k == 0:
            raise Exception("\n\nModel is already finite")
        if not _is_int(num):
            raise Exception("\n\nArgument num not an integer")
        if num < 1:
            raise Exception("\n\nArgument num must be positive!")
        if num == 1 and glue_edgs:
            raise Exception("\n\nCan't have num==1 and glueing of the edges!")
        if list(self._per).count(fin_dir) != 1:
            raise Exception("\n\nCan not make model finite along this direction!")
        shift = np.zeros(self._dim_r)
        shift[fin_dir] = 1.0
        fin_orb = np.concatenate([self._orb + float(c) * shift for c in range(num)], axis=0)
        onsite = np.concatenate([self._site_energies] * num, axis=0)
        fin_per = [p for p in self._per if p != fin_dir]
        fin = self.__class__(self._dim_k - 1, self._dim_r, copy.deepcopy(self._lat), fin_orb, fin_per, self._nspin)
        fin._assume_position_operator_diagonal = self._assume_position_operator_diagonal
        fin._convention = self._convention
        fin.set_onsite(onsite, mode="reset")
        ntot = self._norb * num
        for c in range(num):
            for h in self._hoppings:
                ind_R = np.array(h[3]).copy()
                jump = int(ind_R[fin_dir])
                if fin._dim_k != 0:
                    ind_R[fin_dir] = 0
                hi = h[1] + c * self._norb
                hj = h[2] + (c + jump) * self._norb
                if glue_edgs:
                    hj = int(hj) % int(ntot)
                elif hj < 0 or hj >= ntot:
                    continue
                if fin._dim_k == 0:
                    fin.set_hop(h[0], hi, hj, mode="add", allow_conjugate_pair=True)
                else:
                    fin.set_hop(h[0], hi, hj, ind_R, mode="add", allow_conjugate_pair=True)
        return fin

    def reduce_dim(self, remove_k, value_k, lazy=False):
        """pythtb.py:1233-1311: fix one k component, fold its phase into the amplitudes.

        ``lazy=True`` (extension, SURVEY.md section 8f: parameter sweeps ``for kx in ...: model.reduce_dim(0, kx)``):
        the reduced model is obtained from THIS model's compiled plan with array operations
        (``_plan.reduce_plan``: H_red(k') = H(k', k_remove = value)) — no hopping list is walked and no Python
        model is rebuilt per value; ``_hoppings`` / ``_site_energies`` of the returned model are materialised with
        the reference algorithm only if something reads them."""
        if lazy:
            return _ReducedMixin._make(self, remove_k, value_k)
        if self._dim_k == 0:
            raise Exception("\n\nCan not reduce dimensionality even further!")
        if hasattr(self, "_materialise"):
            self._materialise()                     # a lazily reduced model: from here on it is an ordinary one
        red = copy.deepcopy(self)
        red._per.remove(remove_k)
        red._dim_k = len(red._per)
        if red._dim_k != self._dim_k - 1:
            raise Exception("\n\nSpecified wrong dimension to reduce!")
        red._hoppings = []
        red._hop_index = {}
        for hop in self._hoppings:
            amp = complex(hop[0]) if self._nspin == 1 else np.array(hop[0], dtype=complex)
            i, j = hop[1], hop[2]
            ind_R = np.array(hop[3], dtype=int)
            rv = -red._orb[i, :] + red._orb[j, :] + np.array(ind_R, dtype=float)
            phase = np.exp((2.0j) * np.pi * (value_k * rv[remove_k]))
            if i == j and not np.any(ind_R[red._per]):
                if ind_R[remove_k] == 0:
                    red.set_onsite(amp * phase, i, mode="add")
                elif self._nspin == 1:
                    red.set_onsite(amp * phase + (amp * phase).conj(), i, mode="add")
                else:
                    red.set_onsite(amp * phase + (amp.T * phase).conj(), i, mode="add")
            else:
                ind_R[remove_k] = 0
                red.set_hop(amp * phase, i, j, ind_R, mode="add", allow_conjugate_pair=True)
        red._touch()
        return red

    def change_nonperiodic_vector(self, np_dir, new_latt_vec=None, to_home=True, to_home_suppress_warning=False):
        """pythtb.py:1313-1438: redefine one non-periodic lattice vector keeping
        Cartesian orbital positions."""
        if list(self._per).count(np_dir) == 1:
            print("\nnp_dir =", np_dir)
            raise Exception("Selected direction is not nonperiodic")
        if new_latt_vec is None:
            per_rows = np.zeros_like(self._lat)
            for d in self._per:
                per_rows[d] = self._lat[d]
            coeffs = np.linalg.lstsq(per_rows.T, self._lat[np_dir], rcond=None)[0]
            new_vec = self._lat[np_dir] - np.dot(self._lat.T, coeffs)
        else:
            new_vec = np.array(new_latt_vec)
            if new_vec.shape != (self._dim_r,):
                raise Exception("\n\nNonperiodic vector has wrong length")
        new_lat = copy.deepcopy(self._lat)
        new_lat[np_dir] = new_vec
        new_orb = [np.linalg.solve(new_lat.T, np.dot(self._lat.T, o)) for o in self._orb]
        out = copy.deepcopy(self)
        out._lat = np.array(new_lat, dtype=float)
        out._orb = np.array(new_orb, dtype=float)
        if new_latt_vec is None:
            for i in out._per:
                if np.abs(np.dot(out._lat[i], out._lat[np_dir])) > 1.0e-6:
                    raise Exception("\n\nThis shouldn't happen.  New nonperiodic vector\n"
                                    "is not perpendicular to periodic vectors!?")
        for i in range(self._norb):
            if np.max(np.abs(np.dot(self._lat.T, self._orb[i]) - np.dot(out._lat.T, out._orb[i]))) > 1.0e-6:
                raise Exception("\n\nThis shouldn't happen. New choice of nonperiodic vector\n"
                                "somehow changed Cartesian coordinates of orbitals.")
        if np.abs(np.linalg.det(out._lat)) < 1.0e-6:
            raise Exception("\n\nLattice with new choice of nonperiodic vector has zero volume?!")
        if to_home:
            out._shift_to_home(to_home_suppress_warning)
        out._touch()
        return out

    def make_supercell(self, sc_red_lat, return_sc_vectors=False, to_home=True, to_home_suppress_warning=False):
        """pythtb.py:1440-1637: arbitrary integer supercell of the periodic directions."""
        if self._dim_r == 0:
            raise Exception("\n\nMust have at least one periodic direction to make a super-cell")
        use = np.array(sc_red_lat)
        if use.shape != (self._dim_r, self._dim_r):
            raise Exception("\n\nDimension of sc_red_lat array must be dim_r*dim_r")
        if use.dtype != int:                # pythtb.py:1508-1509: no silent truncation of 2.5 -> 2
            raise Exception("\n\nsc_red_lat array elements must be integers")
        for i in range(self._dim_r):
            for j in range(self._dim_r):
                if i == j:
                    if i not in self._per and use[i, j] != 1:
                        raise Exception("\n\nDiagonal elements of sc_red_lat for non-periodic directions must equal 1.")
                elif (i not in self._per or j not in self._per) and use[i, j] != 0:
                    raise Exception("\n\nOff-diagonal elements of sc_red_lat for non-periodic directions must equal 0.")
        det = np.linalg.det(use)
        if np.abs(det) < 1.0e-6:
            raise Exception("\n\nSuper-cell lattice vectors length/area/volume too close to zero, or zero.")
        if det < 0.0:
            raise Exception("\n\nSuper-cell lattice vectors need to form right handed system.")
        # candidate translations: bounding box of the supercell corners
        max_R = np.max(np.abs(use)) * self._dim_r
        rng = [range(-max_R, max_R + 1)] * self._dim_r
        grid = np.array(np.meshgrid(*rng, indexing="ij")).reshape(self._dim_r, -1).T
        eps_shift = np.sqrt(2.0) * 1.0e-8

        def to_sc(red):
            return np.linalg.solve(np.array(use.T, dtype=float), np.array(red, dtype=float).T).T

        frac = to_sc(grid)
        inside = np.all((frac > -eps_shift) & (frac <= 1.0 - eps_shift), axis=1)
        sc_vec = grid[inside]
        # the reference enumerates candidates in C order of the box, so does meshgrid('ij')
        if int(round(np.abs(det))) != len(sc_vec):
            raise Exception("\n\nSuper-cell generation failed! Wrong number of super-cell vectors found.")
        sc_cart_lat = np.dot(use, self._lat)
        sc_orb = []
        for cur in sc_vec:
            for o in self._orb:
                sc_orb.append(to_sc(o + np.array(cur, dtype=float)))
        sc = self.__class__(self._dim_k, self._dim_r, sc_cart_lat, sc_orb, per=self._per, nspin=self._nspin)
        sc._assume_position_operator_diagonal = self._assume_position_operator_diagonal
        sc._convention = self._convention
        lookup = {tuple(int(x) for x in v): n for n, v in enumerate(sc_vec)}
        for icur, cur in enumerate(sc_vec):
            for i in range(self._norb):
                sc.set_onsite(self._site_energies[i], i + icur * self._norb)
            for h in self._hoppings:
                amp, i, j = h[0], h[1], h[2]
                ind_R = np.array(h[3], dtype=int)
                target = cur + ind_R
                fr = to_sc(target)
                sc_part = np.floor(fr).astype(int)
                orig_part = target - np.dot(sc_part, use)
                pair = lookup.get(tuple(int(x) for x in orig_part))
                if pair is None:
                    raise Exception("\n\nDid not find super cell vector!")
                sc.set_hop(amp, i + icur * self._norb, j + pair * self._norb, sc_part,
                           mode="add", allow_conjugate_pair=True)
        if to_home:
            sc._shift_to_home(to_home_suppress_warning)
        if return_sc_vectors:
            return sc, sc_vec
        return sc

    def _shift_to_home(self, to_home_suppress_warning=False):
        """pythtb.py:1639-1715.  As shipped in 1.8.0 the shift block sits after
        the orbital loop and under ``to_home_suppress_warning==False``, so only
        the LAST orbital is moved into the home cell and only when warnings are
        enabled.  Orbital positions enter every Berry
        phase, so that behaviour is reproduced deliberately."""
        flagged = [[] for _ in range(self._dim_r)]
        disp = np.zeros(self._dim_r, dtype=int)
        for i in range(self._norb):
            disp = np.zeros(self._dim_r, dtype=int)
            for k in range(self._dim_r):
                shift = int(np.floor(self._orb[i, k] + 1.0e-6))
                if k in self._per:
                    disp[k] = shift
                elif shift != 0:
                    flagged[k].append(i)
        if to_home_suppress_warning:
            return
        lines = ["  * Direction %1d : Orbitals " % k + ", ".join(str(e) for e in orbs)
                 for k, orbs in enumerate(flagged) if orbs]
        if lines:
            print("  WARNING from '_shift_to_home': orbitals are not shifted to the home cell along\n"
                  "  non-periodic directions (PythTB >= 1.7.3 behaviour):\n" + "\n".join(lines))
        if self._norb == 0:
            return
        i = self._norb - 1
        self._orb[i] -= disp
        if self._dim_k != 0:
            for h in self._hoppings:
                if h[1] == i:
                    h[3] = h[3] - disp
                if h[2] == i:
                    h[3] = h[3] + disp
        self._reindex()
        self._touch()

    def remove_orb(self, to_remove):
        """pythtb.py:1718-1789."""
        idx = [to_remove] if _is_int(to_remove) else list(copy.deepcopy(to_remove))
        for o in idx:
            if (not _is_int(o)) or o < 0 or o > self._norb - 1:
                raise Exception("\n\nSpecified wrong orbitals to remove!")
        if len(set(idx)) != len(idx):
            raise Exception("\n\nSpecified duplicate orbitals to remove!")
        ret = copy.deepcopy(self)
        keep = np.array([o for o in range(self._norb) if o not in set(idx)], dtype=int)
        newid = -np.ones(self._norb, dtype=int)
        newid[keep] = np.arange(len(keep))
        ret._norb = len(keep)
        ret._nsta = ret._norb * self._nspin
        ret._orb = ret._orb[keep]
        ret._site_energies = ret._site_energies[keep]
        ret._site_energies_specified = ret._site_energies_specified[keep]
        hops = []
        for h in ret._hoppings:
            if newid[h[1]] < 0 or newid[h[2]] < 0:
                continue
            h[1], h[2] = int(newid[h[1]]), int(newid[h[2]])
            hops.append(h)
        ret._hoppings = hops
        ret._reindex()
        ret._touch()
        return ret

    # ------------------------------------------------------------ k-list helpers
    def k_uniform_mesh(self, mesh_size, lazy=False):
        """pythtb.py:1792-1861: all points (i/N1, j/N2, ...) in C order.
        ``lazy=True`` (extension) returns a ``KMesh`` descriptor that ``solve_all``
        expands on the device."""
        use = np.array(list(map(round, mesh_size)), dtype=int)
        if use.shape != (self._dim_k,):
            print(use.shape)
            raise Exception("\n\nIncorrect size of the specified k-mesh!")
        if np.min(use) <= 0:
            raise Exception("\n\nMesh must have positive non-zero number of elements.")
        if self._dim_k not in (1, 2, 3):
            raise Exception("\n\nUnsupported dim_k!")
        if lazy:
            return KMesh(use)
        return np.asarray(KMesh(use))

    def k_path(self, kpts, nk, report=True):
        """pythtb.py:1863-2026: piecewise-linear path with nearly equidistant
        points (distance measured with the reciprocal metric)."""
        if isinstance(kpts, str):
            named = {"full": [[0.0], [0.5], [1.0]], "fullc": [[-0.5], [0.0], [0.5]], "half": [[0.0], [0.5]]}
            k_list = np.array(named[kpts])
        else:
            k_list = np.array(kpts)
        if k_list.ndim == 1 and self._dim_k == 1:
            k_list = np.array([k_list]).T
        if k_list.shape[1] != self._dim_k:
            print("input k-space dimension is", k_list.shape[1])
            print("k-space dimension taken from model is", self._dim_k)
            raise Exception("\n\nk-space dimensions do not match")
        if nk < k_list.shape[0]:
            raise Exception("\n\nMust have more points in the path than number of nodes.")
        n_nodes = k_list.shape[0]
        lat_per = np.copy(self._lat)[self._per]
        k_metric = np.linalg.inv(np.dot(lat_per, lat_per.T))
        k_node = np.zeros(n_nodes, dtype=float)
        for n in range(1, n_nodes):
            dk = k_list[n] - k_list[n - 1]
            k_node[n] = k_node[n - 1] + np.sqrt(np.dot(dk, np.dot(k_metric, dk)))
        node_index = [0]
        for n in range(1, n_nodes - 1):
            node_index.append(int(round(k_node[n] / k_node[-1] * (nk - 1))))
        node_index.append(nk - 1)
        k_dist = np.zeros(nk, dtype=float)
        k_vec = np.zeros((nk, self._dim_k), dtype=float)
        k_vec[0] = k_list[0]
        for n in range(1, n_nodes):
            n_i, n_f = node_index[n - 1], node_index[n]
            for j in range(n_i, n_f + 1):
                frac = float(j - n_i) / float(n_f - n_i)
                k_dist[j] = k_node[n - 1] + frac * (k_node[n] - k_node[n - 1])
                k_vec[j] = k_list[n - 1] + frac * (k_list[n] - k_list[n - 1])
        if report:
            if self._dim_k == 1:
                print(" Path in 1D BZ defined by nodes at " + str(k_list.flatten()))
            else:
                print("----- k_path report begin ----------")
                print("real-space lattice vectors\n", lat_per)
                print("k-space metric tensor\n", k_metric)
                print("internal coordinates of nodes\n", k_list)
                print("node distance list:", k_node)
                print("node index list:   ", np.array(node_index))
                print("----- k_path report end ------------")
            print()
        return (k_vec, k_dist, k_node)

    def ignore_position_operator_offdiagonal(self):
        """pythtb.py:2028-2032."""
        self._assume_position_operator_diagonal = True

    # --------------------------------------------------------- position operator
    def _position_checks(self, dir):
        if dir in self._per:
            raise Exception("Can not compute position matrix elements along periodic direction!")
        if dir < 0 or dir >= self._dim_r:
            raise Exception("Direction out of range!")
        if not self._assume_position_operator_diagonal:
            _offdiag_approximation_warning_and_stop()

    def position_matrix(self, evec, dir):
        """X_mn = <u_m| r_dir |u_n>, pythtb.py:2034-2113 (GPU: one contraction kernel)."""
        self._position_checks(dir)
        ev = np.asarray(evec, dtype=complex)
        flat = ev.reshape(ev.shape[0], -1)
        mat = self._engine().position_matrix(self, flat[None], dir)[0]
        if np.max(mat - mat.T.conj()) > 1.0e-9:
            raise Exception("\n\n Position matrix is not hermitian?!")
        return mat

    def position_expectation(self, evec, dir):
        """pythtb.py:2115-2160."""
        if not self._assume_position_operator_diagonal:
            _offdiag_approximation_warning_and_stop()
        return np.array(np.real(self.position_matrix(evec, dir).diagonal()), dtype=float)

    def position_hwf(self, evec, dir, hwf_evec=False, basis="orbital"):
        """Hybrid Wannier centres/functions, pythtb.py:2162-2279 (GPU:
        contraction + batched eigh + rotation to the orbital basis)."""
        self._position_checks(dir)
        b = basis.lower().strip()
        if hwf_evec and b not in ("wavefunction", "bloch", "orbital"):
            raise Exception("\n\nBasis must be either 'wavefunction', 'bloch', or 'orbital'")
        ev = np.asarray(evec, dtype=complex)
        flat = ev.reshape(ev.shape[0], -1)
        hwfc, hwf = self._engine().position_hwf(self, flat[None], dir, hwf_evec, b == "orbital")
        if not hwf_evec:
            return hwfc[0]
        out = hwf[0]
        if b == "orbital" and self._nspin == 2:
            out = out.reshape(out.shape[0], self._norb, 2)
        return hwfc[0], out


class _ReducedMixin(object):
    """``parent.reduce_dim(remove_k, value_k, lazy=True)``: a ``tb_model`` of one dimension less whose compiled plan
    is derived from the parent's plan (``_plan.reduce_plan``); everything numerical (``solve_all``, ``wf_array``)
    runs from that plan.  The Python-level description (``_hoppings``, ``_site_energies``) is built by the
    reference algorithm (pythtb.py:1268-1310) on first access; a model that is edited afterwards behaves like any
    other ``tb_model``."""
    _hoppings = _LazyAttr("hoppings")
    _site_energies = _LazyAttr("site_energies")
    _site_energies_specified = _LazyAttr("site_energies_specified")
    _hop_index = _LazyAttr("hop_index")

    _classes = {}

    @classmethod
    def _make(cls, parent, remove_k, value_k):
        base = type(parent)
        if not issubclass(base, _ReducedMixin):            # keep the parent's class (and engine) underneath
            if base not in _ReducedMixin._classes:
                _ReducedMixin._classes[base] = type(base.__name__ + "_reduced", (_ReducedMixin, base), {})
            base = _ReducedMixin._classes[base]
        cls = base
        if parent._dim_k == 0:
            raise Exception("\n\nCan not reduce dimensionality even further!")
        if list(parent._per).count(remove_k) != 1:
            raise Exception("\n\nSpecified wrong dimension to reduce!")
        from ._plan import reduce_plan
        snap = parent._plan()
        new = cls.__new__(cls)
        for k, v in parent.__dict__.items():
            if k in ("_hoppings", "_site_energies", "_site_energies_specified", "_hop_index", "_plan_cache") or k.startswith("_lazy_"):
                continue
            new.__dict__[k] = copy.copy(v) if isinstance(v, (list, np.ndarray)) else v
        new._per = [p for p in parent._per if p != remove_k]
        new._dim_k = len(new._per)
        new._lazy = (parent, parent.__dict__.get("_version", 0), remove_k, float(value_k))
        new._red_plan = reduce_plan(snap, list(parent._per).index(remove_k), float(value_k))
        new._plan_cache = None
        return new

    def _materialise(self):
        lazy = self.__dict__.get("_lazy")
        if lazy is None:
            return
        parent, version, remove_k, value_k = lazy
        if parent.__dict__.get("_version", 0) != version:
            raise Exception("\n\nThe model this lazily reduced model was derived from has been modified since;"
                            "\ncall reduce_dim again.")
        self.__dict__["_lazy"] = None
        eager = tb_model.reduce_dim(parent, remove_k, value_k)
        for name in ("hoppings", "site_energies", "site_energies_specified", "hop_index"):
            self.__dict__["_lazy_" + name] = getattr(eager, "_" + name)
        self.__dict__["_red_plan"] = None

    def _plan(self, convention=None):
        red = self.__dict__.get("_red_plan")
        if red is not None and self.__dict__.get("_lazy") is not None and convention in (None, red.convention):
            return red
        return tb_model._plan(self, convention)

    def __deepcopy__(self, memo):
        if self.__dict__.get("_lazy") is None:
            return tb_model.__deepcopy__(self, memo)
        new = self.__class__.__new__(self.__class__)          # the plan is immutable and the parent only a witness:
        memo[id(self)] = new                                  # a copy (wf_array takes one) shares both
        for k, v in self.__dict__.items():
            new.__dict__[k] = v if k in ("_lazy", "_red_plan") else copy.deepcopy(v, memo)
        return new
