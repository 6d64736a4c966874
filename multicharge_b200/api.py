"""Host-side mirror of the reference's model API for the EEQ hot path, on top of the C ABI.

Names, argument meaning and error behaviour follow ``module multicharge`` (src/multicharge.f90:16-27):

    reference (Fortran)                               here
    ------------------------------------------------  -----------------------------------------
    type(structure_type)            (mctc-lib)        Structure
    new_eeq2019_model(mol, model, error)              new_eeq2019_model(mol) -> MchrgModel
    new_eeq_model(self, mol, error, chi, rad, ...)    new_eeq_model(mol, chi, rad, ...)
    model%ncoord%get_coordination_number(...)         model.get_coordination_number(mol, grad)
    model%local_charge(mol, trans, qloc, ...)         model.local_charge(mol, grad)
    model%solve(mol, error, cn, qloc, dcndr, ...)     model.solve(mol, cn, qloc, dcndr, ..., energy=, ...)
    get_charges(model, mol, error, qvec, dqdr, dqdL)  get_charges(model, mol, grad) -> qvec[, dqdr, dqdL]
    get_eeq_charges(mol, error, qvec, ...)            get_eeq_charges(mol, grad)
    allocated(error)                                  MchrgError / FactorizationError raised

Array convention (memory identical to the Fortran arrays; numpy axes reversed):
    xyz[nat,3], lattice[k] = k-th cell vector, cn[nat], dcndr[i,j,c] = d cn_i/d r_jc, dcndL[i,b,a],
    qvec[nat], energy[nat], gradient[nat,3], sigma[3,3], dqdr[i,j,c] = d q_i/d r_jc, dqdL[i,b,a],
    amat[nat+1,nat+1], dadr[k,j,c], dadL[k,b,a], atrace[i,c], dxdr[k,j,c], dxdL[k,b,a].
Inputs/outputs may be numpy arrays (host) or torch CUDA tensors (device, float64, contiguous).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MchrgError, check


class FactorizationError(MchrgError):
    """fatal_error("Bunch-Kaufman factorization failed.") of model/type.F90:222-225: the Coulomb
    block is not positive definite; ``info`` is the 1-based index of the failed pivot."""

    def __init__(self, info):
        RuntimeError.__init__(self, f"Bunch-Kaufman factorization failed. (non-positive pivot {info})")
        self.code = info
        self.info = info


class _Ptr(C.c_void_p):
    """void* that keeps the array it points into alive for as long as the argument object lives (ctypes holds
    the argument objects until the foreign call returns), so temporaries such as `_ptr(_f64(x))` are safe."""


def _ptr(a, dtype=None):
    """Address of a host (numpy) or device (torch) array for the C ABI; `dtype` is enforced when given."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        if dtype is not None and a.dtype != np.dtype(dtype):
            raise TypeError(f"array must have dtype {np.dtype(dtype)}, got {a.dtype}")
        p = _Ptr(a.ctypes.data)
        p._keep = a
        return p
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if dtype is not None and str(a.dtype).replace("torch.", "") != np.dtype(dtype).name:
            raise TypeError(f"tensor must have dtype {np.dtype(dtype)}, got {a.dtype}")
        p = _Ptr(a.data_ptr())
        p._keep = a
        return p
    raise TypeError(f"unsupported array type {type(a)}")


def _f64(a):
    if a is None or hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One GPU + one stream + workspace (mchrg_b200_context)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(_lib.load().mchrg_b200_context_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().mchrg_b200_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return _lib.load().mchrg_b200_stream(self._h)

    @property
    def launch_count(self):
        return int(_lib.load().mchrg_b200_launch_count(self._h))

    def synchronize(self):
        check(_lib.load().mchrg_b200_synchronize(self._h))

    def batch_trace(self):
        """Per-kernel times (ms) and launch counts of the last large batch call made with MCHRG_B200_TRACE=1:
        (pair phase A, factorisation, pair phase C)."""
        ms = (C.c_double * 3)()
        nl = (C.c_int32 * 3)()
        check(_lib.load().mchrg_b200_batch_trace(self._h, ms, nl))
        return list(ms), list(nl)

    @property
    def side_stream(self):
        return _lib.load().mchrg_b200_side_stream(self._h)

    # --- communicator of the multi-GPU single-system path (mchrg_b200_comm_*) -------------------------
    @staticmethod
    def comm_unique_id():
        """ncclGetUniqueId through the library: 128 bytes that rank 0 hands to every rank."""
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        check(_lib.load().mchrg_b200_comm_unique_id(buf))
        return buf.raw

    def comm_init_nccl(self, rank, world, unique_id):
        """One process per GPU: join the NCCL communicator described by `unique_id` as `rank` of `world`."""
        if len(unique_id) != _lib.COMM_ID_BYTES:
            raise ValueError("unique_id must be the 128 bytes of comm_unique_id()")
        check(_lib.load().mchrg_b200_comm_init_nccl(self._h, int(rank), int(world), C.c_char_p(bytes(unique_id))))

    @staticmethod
    def comm_init_local(contexts):
        """One process, one host thread per context: tie `contexts` together (rank = position in the list)."""
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        check(_lib.load().mchrg_b200_comm_init_local(arr, len(contexts)))

    def comm_destroy(self):
        check(_lib.load().mchrg_b200_comm_destroy(self._h))

    @property
    def comm_rank(self):
        return int(_lib.load().mchrg_b200_comm_rank(self._h))

    @property
    def comm_size(self):
        return int(_lib.load().mchrg_b200_comm_size(self._h))

    def comm_bcast(self, t, root):
        """Broadcast the device tensor `t` of rank `root` to every rank, in place."""
        check(_lib.load().mchrg_b200_comm_bcast(self._h, _ptr(t), t.numel() * t.element_size(), int(root)))

    def comm_allreduce_sum(self, t):
        """In-place sum over ranks of a float64 device tensor."""
        check(_lib.load().mchrg_b200_comm_allreduce_sum(self._h, _ptr(t, np.float64), t.numel()))

    def comm_allgather(self, send, recv):
        """recv[r] = send of rank r (device tensors; recv holds comm_size times the bytes of send)."""
        nbytes = send.numel() * send.element_size()
        if recv.numel() * recv.element_size() != nbytes * self.comm_size:
            raise ValueError("recv must hold comm_size * send bytes")
        check(_lib.load().mchrg_b200_comm_allgather(self._h, _ptr(send), _ptr(recv), nbytes))

    # dense building blocks (tests)
    def dpotrf(self, a):
        """In-place lower Cholesky of a column-major n x n matrix given as a[n,n] Fortran-ordered view."""
        n = a.shape[0]
        return check(_lib.load().mchrg_b200_dpotrf(self._h, n, _ptr(a), n))

    def dgemm(self, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        return check(_lib.load().mchrg_b200_dgemm(self._h, transb.encode(), m, n, k, alpha, _ptr(a), lda, _ptr(b), ldb,
                                                  beta, _ptr(c), ldc))

    def dpotrs_right(self, m, n, l, ldl, x, ldx):
        return check(_lib.load().mchrg_b200_dpotrs_right(self._h, m, n, _ptr(l), ldl, _ptr(x), ldx))


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Structure:
    """type(structure_type) of mctc-lib, reduced to what the path reads: nat, nid, id (1-based species),
    num (atomic number per species), xyz (Bohr), charge, lattice, periodic."""

    def __init__(self, num, xyz, charge=0.0, lattice=None, periodic=None, id=None):
        self.xyz = _f64(np.asarray(xyz, dtype=np.float64).reshape(-1, 3))
        num = np.asarray(num, dtype=np.int32)
        if id is None:  # atomic numbers per atom -> species in order of first appearance (mctc-lib new())
            uniq = []
            for z in num.tolist():
                if z not in uniq:
                    uniq.append(z)
            self.id = np.array([uniq.index(z) + 1 for z in num.tolist()], dtype=np.int32)
            self.num = np.array(uniq, dtype=np.int32)
        else:
            self.id = np.ascontiguousarray(id, dtype=np.int32)
            self.num = num
        self.nat = int(self.xyz.shape[0])
        self.nid = int(len(self.num))
        self.charge = float(charge)
        self.lattice = np.zeros((3, 3)) if lattice is None else _f64(np.asarray(lattice, dtype=np.float64).reshape(3, 3))
        if periodic is None:
            periodic = [lattice is not None] * 3
        self.periodic = np.array([bool(p) for p in np.broadcast_to(periodic, (3,))])

    def _c(self, xyz=None, id=None):
        s = _lib.Structure()
        s.nat, s.nid = self.nat, self.nid
        s.id = _ptr(self.id if id is None else id, np.int32)      # the arrays outlive the call: members of self
        s.xyz = _ptr(self.xyz if xyz is None else xyz, np.float64)  # or the caller's own tensors
        s.charge = self.charge
        s.periodic = (C.c_int32 * 3)(*[int(p) for p in self.periodic])
        s.lattice = (C.c_double * 9)(*self.lattice.ravel().tolist())
        return s


class SolveResult(dict):
    __getattr__ = dict.get


class MchrgModel:
    """class(mchrg_model_type) (model/type.F90:42-76): immutable after construction."""

    def __init__(self, handle, ctx, kind):
        self._h = handle
        self.ctx = ctx
        self.kind = kind

    def __del__(self):
        try:
            if self._h:
                _lib.load().mchrg_b200_model_free(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def _status(rc):
        check(rc)
        if rc > 0:
            raise FactorizationError(rc)

    # --- mctc_ncoord ---------------------------------------------------------------------------
    def get_coordination_number(self, mol, grad=False):
        n = mol.nat
        cn = np.zeros(n)
        dcndr = np.zeros((n, n, 3)) if grad else None
        dcndL = np.zeros((n, 3, 3)) if grad else None
        s = mol._c()
        self._status(_lib.load().mchrg_b200_get_coordination_number(self._h, C.byref(s), _ptr(cn), _ptr(dcndr), _ptr(dcndL)))
        return (cn, dcndr, dcndL) if grad else cn

    def local_charge(self, mol, grad=False):
        n = mol.nat
        qloc = np.zeros(n)
        dqlocdr = np.zeros((n, n, 3)) if grad else None
        dqlocdL = np.zeros((n, 3, 3)) if grad else None
        s = mol._c()
        self._status(_lib.load().mchrg_b200_local_charge(self._h, C.byref(s), _ptr(qloc), _ptr(dqlocdr), _ptr(dqlocdL)))
        return (qloc, dqlocdr, dqlocdL) if grad else qloc

    # --- deferred procedures --------------------------------------------------------------------
    def get_xvec(self, mol, cn, qloc=None):
        x = np.zeros(mol.nat + 1)
        s = mol._c()
        self._status(_lib.load().mchrg_b200_get_xvec(self._h, C.byref(s), _ptr(_f64(cn)), _ptr(_f64(qloc)), _ptr(x)))
        return x

    def get_xvec_derivs(self, mol, cn, dcndr, dcndL, qloc=None, dqlocdr=None, dqlocdL=None):
        n = mol.nat
        dxdr = np.zeros((n + 1, n, 3))
        dxdL = np.zeros((n + 1, 3, 3))
        s = mol._c()
        self._status(_lib.load().mchrg_b200_get_xvec_derivs(
            self._h, C.byref(s), _ptr(_f64(cn)), _ptr(_f64(qloc)), _ptr(_f64(dcndr)), _ptr(_f64(dcndL)),
            _ptr(_f64(dqlocdr)), _ptr(_f64(dqlocdL)), _ptr(dxdr), _ptr(dxdL)))
        return dxdr, dxdL

    def get_coulomb_matrix(self, mol, cn=None, qloc=None):
        a = np.zeros((mol.nat + 1, mol.nat + 1))
        s = mol._c()
        self._status(_lib.load().mchrg_b200_get_coulomb_matrix(self._h, C.byref(s), _ptr(_f64(cn)), _ptr(_f64(qloc)), _ptr(a)))
        return a

    def get_coulomb_derivs(self, mol, qvec, cn=None, qloc=None, dcndr=None, dcndL=None, dqlocdr=None, dqlocdL=None):
        n = mol.nat
        dadr = np.zeros((n + 1, n, 3))
        dadL = np.zeros((n + 1, 3, 3))
        atrace = np.zeros((n, 3))
        s = mol._c()
        self._status(_lib.load().mchrg_b200_get_coulomb_derivs(
            self._h, C.byref(s), _ptr(_f64(cn)), _ptr(_f64(qloc)), _ptr(_f64(dcndr)), _ptr(_f64(dcndL)),
            _ptr(_f64(dqlocdr)), _ptr(_f64(dqlocdL)), _ptr(_f64(qvec)), _ptr(dadr), _ptr(dadL), _ptr(atrace)))
        return dadr, dadL, atrace

    # --- solve ----------------------------------------------------------------------------------
    def solve(self, mol, cn, qloc=None, dcndr=None, dcndL=None, dqlocdr=None, dqlocdL=None,
              energy=None, gradient=None, sigma=None, qvec=None, dqdr=None, dqdL=None):
        """mchrg_model_type%solve: outputs are the arrays passed in (None = not present);
        energy and sigma are accumulated into, the others overwritten."""
        s = mol._c()
        rc = _lib.load().mchrg_b200_solve(
            self._h, C.byref(s), _ptr(_f64(cn)), _ptr(_f64(qloc)), _ptr(_f64(dcndr)), _ptr(_f64(dcndL)),
            _ptr(_f64(dqlocdr)), _ptr(_f64(dqlocdL)), _ptr(energy), _ptr(gradient), _ptr(sigma), _ptr(qvec),
            _ptr(dqdr), _ptr(dqdL))
        self._status(rc)

    def singlepoint(self, mol, want=("qvec", "energy", "gradient"), xyz=None, id=None):
        """CN -> local charge -> solve (app/main.f90:114-118); returns freshly allocated host arrays."""
        n = mol.nat
        out = SolveResult(cn=np.zeros(n), qloc=np.zeros(n))
        if "qvec" in want:
            out["qvec"] = np.zeros(n)
        if "energy" in want:
            out["energy"] = np.zeros(n)
        if "gradient" in want:
            out["gradient"] = np.zeros((n, 3))
            out["sigma"] = np.zeros((3, 3))
        if "dqdr" in want:
            out["dqdr"] = np.zeros((n, n, 3))
            out["dqdL"] = np.zeros((n, 3, 3))
        s = mol._c(xyz=xyz, id=id)
        rc = _lib.load().mchrg_b200_singlepoint(
            self._h, C.byref(s), _ptr(out.cn), _ptr(out.qloc), _ptr(out.energy), _ptr(out.gradient), _ptr(out.sigma),
            _ptr(out.qvec), _ptr(out.dqdr), _ptr(out.dqdL))
        self._status(rc)
        return out


def singlepoint_rows(model, mol, row_begin, row_end, dqdr=True, out=None, xyz=None, id=None):
    """One rank's share of a single large system (mchrg_b200_singlepoint_rows): derivative outputs for the atom
    block [row_begin, row_end) only -- gradient[nrows,3], dqdr[nat, nrows, 3] (dqdr[i, j - row_begin, c] =
    d q_i / d r_jc); cn, qvec, energy, sigma, dqdL complete.  `out` may hold caller (device) buffers."""
    n, nj = mol.nat, row_end - row_begin
    if out is None:
        out = SolveResult(cn=np.zeros(n), qvec=np.zeros(n), energy=np.zeros(n), gradient=np.zeros((nj, 3)),
                          sigma=np.zeros((3, 3)), dqdr=np.zeros((n, nj, 3)) if dqdr else None,
                          dqdL=np.zeros((n, 3, 3)) if dqdr else None)
    s = mol._c(xyz=xyz, id=id)
    rc = _lib.load().mchrg_b200_singlepoint_rows(
        model._h, C.byref(s), int(row_begin), int(row_end), _ptr(out.get("cn")), _ptr(out.get("energy")),
        _ptr(out.get("gradient")), _ptr(out.get("sigma")), _ptr(out.get("qvec")), _ptr(out.get("dqdr")), _ptr(out.get("dqdL")))
    MchrgModel._status(rc)
    return out


def new_eeq_model(mol, chi, rad, eta, kcnchi, rcov, cutoff=25.0, cn_exp=7.5, cn_max=8.0, ctx=None):
    """new_eeq_model (model/eeq.f90:62-95); per-species arrays of length mol.nid."""
    ctx = ctx or default_context()
    arrs = [_f64(a) for a in (chi, rad, eta, kcnchi, rcov)]
    if any(len(a) != mol.nid for a in arrs):
        raise ValueError("per-species parameter arrays must have length mol.nid")
    h = C.c_void_p()
    check(_lib.load().mchrg_b200_new_eeq_model(ctx._h, mol.nid, *[_ptr(a) for a in arrs], float(cutoff), float(cn_exp),
                                               float(cn_max), C.byref(h)))
    return MchrgModel(h, ctx, 1)


def new_eeq2019_model(mol, ctx=None):
    """new_eeq2019_model (param.f90:46-73)."""
    ctx = ctx or default_context()
    h = C.c_void_p()
    num = np.ascontiguousarray(mol.num, dtype=np.int32)
    check(_lib.load().mchrg_b200_new_eeq2019_model(ctx._h, mol.nid, _ptr(num), C.byref(h)))
    return MchrgModel(h, ctx, 1)


def get_charges(model, mol, grad=False):
    """get_charges (charge.f90:37-77): returns qvec, or (qvec, dqdr, dqdL) when grad."""
    n = mol.nat
    qvec = np.zeros(n)
    dqdr = np.zeros((n, n, 3)) if grad else None
    dqdL = np.zeros((n, 3, 3)) if grad else None
    s = mol._c()
    MchrgModel._status(_lib.load().mchrg_b200_get_charges(model._h, C.byref(s), _ptr(qvec), _ptr(dqdr), _ptr(dqdL)))
    return (qvec, dqdr, dqdL) if grad else qvec


def get_eeq_charges(mol, grad=False, ctx=None):
    """get_eeq_charges (charge.f90:81-104)."""
    return get_charges(new_eeq2019_model(mol, ctx), mol, grad)


def new_eeq2019_batch_model(ctx=None):
    """EEQ2019 model over all elements (species index == atomic number, Z = 1..103), as the batch
    entry point requires."""
    ctx = ctx or default_context()
    h = C.c_void_p()
    num = np.arange(1, 104, dtype=np.int32)
    check(_lib.load().mchrg_b200_new_eeq2019_model(ctx._h, 103, _ptr(num), C.byref(h)))
    return MchrgModel(h, ctx, 1)


def singlepoint_batch(model, offsets, num, xyz, charge, gradient=True, out=None):
    """CN -> solve(energy, gradient, qvec) for a batch of independent non-periodic molecules
    (one get_charges / CLI single point per molecule, app/main.f90:114-118).

    offsets[nmol+1] int64, num[ntot] int32 atomic numbers, xyz[ntot,3], charge[nmol]; arrays may be
    numpy (host) or torch CUDA tensors (device).  Returns dict(qvec, energy, gradient, status);
    pass `out` (same keys) to reuse buffers / keep results on the device.
    """
    nmol = int(offsets.shape[0]) - 1
    host = isinstance(xyz, np.ndarray)
    if out is None:
        if not host:
            raise ValueError("device inputs need caller-provided `out` tensors")
        ntot = int(offsets[-1])
        out = SolveResult(qvec=np.empty(ntot), energy=np.empty(ntot),
                          gradient=np.empty((ntot, 3)) if gradient else None,
                          status=np.empty(nmol, dtype=np.int32))
    f8, i4, i8 = np.float64, np.int32, np.int64
    rc = _lib.load().mchrg_b200_singlepoint_batch(
        model._h, nmol, _ptr(offsets, i8), _ptr(num, i4), _ptr(xyz, f8), _ptr(charge, f8), _ptr(out["qvec"], f8),
        _ptr(out["energy"], f8), _ptr(out.get("gradient"), f8), _ptr(out["status"], i4))
    check(rc)
    return out


def new_eeqbc_model(mol, chi, rad, eta, kcnchi, kqchi, kqeta, kcnrad, cap, avg_cn, rvdw, rcov, en,
                    kbc=0.65, cutoff=25.0, cn_exp=2.0, norm_exp=1.0, ctx=None):
    """new_eeqbc_model (model/eeqbc.f90:100-176); per-species arrays of length mol.nid, rvdw[nid, nid] per
    species pair (the reference stores the same numbers per atom pair, param.f90:114-115)."""
    ctx = ctx or default_context()
    arrs = {k: _f64(v) for k, v in dict(chi=chi, rad=rad, eta=eta, kcnchi=kcnchi, kqchi=kqchi, kqeta=kqeta, cap=cap,
                                        avg_cn=avg_cn, rcov=rcov, en=en).items()}
    rvdw = _f64(rvdw)
    if any(len(a) != mol.nid for a in arrs.values()) or rvdw.shape != (mol.nid, mol.nid):
        raise ValueError("per-species parameter arrays must have length mol.nid (rvdw: nid x nid)")
    h = C.c_void_p()
    check(_lib.load().mchrg_b200_new_eeqbc_model(
        ctx._h, mol.nid, _ptr(arrs["chi"]), _ptr(arrs["rad"]), _ptr(arrs["eta"]), _ptr(arrs["kcnchi"]),
        _ptr(arrs["kqchi"]), _ptr(arrs["kqeta"]), float(kcnrad), _ptr(arrs["cap"]), _ptr(arrs["avg_cn"]), _ptr(rvdw),
        float(kbc), float(cutoff), float(cn_exp), _ptr(arrs["rcov"]), _ptr(arrs["en"]), float(norm_exp), C.byref(h)))
    return MchrgModel(h, ctx, 2)


def new_eeqbc2025_model(mol, rvdw, en_pauling, ctx=None):
    """new_eeqbc2025_model (param.f90:75-125).  The pairwise D3 vdW radii (get_vdw_rad * autoaa) and the
    Pauling electronegativities live in mctc-lib and must be supplied: rvdw[nid, nid], en_pauling[nid]."""
    ctx = ctx or default_context()
    h = C.c_void_p()
    num = np.ascontiguousarray(mol.num, dtype=np.int32)
    rvdw, en = _f64(rvdw), _f64(en_pauling)
    if rvdw.shape != (mol.nid, mol.nid) or len(en) != mol.nid:
        raise ValueError("rvdw must be nid x nid and en_pauling of length nid")
    check(_lib.load().mchrg_b200_new_eeqbc2025_model(ctx._h, mol.nid, _ptr(num), _ptr(rvdw), _ptr(en), C.byref(h)))
    return MchrgModel(h, ctx, 2)


def get_eeqbc_charges(mol, rvdw, en_pauling, grad=False, ctx=None):
    """get_eeqbc_charges (charge.f90:108-131)."""
    return get_charges(new_eeqbc2025_model(mol, rvdw, en_pauling, ctx), mol, grad)
