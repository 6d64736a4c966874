"""TEST INFRASTRUCTURE -- ctypes front end of the CPU oracle (oracle/mchrg_oracle*.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  It is the checker, never the product path.

Array convention: numpy arrays are C-contiguous with the axes of the reference's Fortran arrays
REVERSED, so the memory is bit-for-bit the Fortran layout:
  xyz[nat,3], lattice[k] = k-th cell vector, dcndr[i,j,c] = d cn_i / d r_jc (Fortran dcndr(c,j,i)),
  dcndL[i,b,a], dqdr[i,j,c], dqdL[i,b,a], amat[ndim,ndim], dadr[k,j,c], dadL[k,b,a], atrace[i,c].
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import build as _build  # noqa: E402

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class _Mol(C.Structure):
    _fields_ = [("nat", C.c_int), ("nid", C.c_int), ("id", c_ip), ("xyz", c_dp), ("charge", C.c_double),
                ("periodic", C.c_int * 3), ("lattice", C.c_double * 9)]


class _Ncoord(C.Structure):
    _fields_ = [("rcov", c_dp), ("en", c_dp), ("kcn", C.c_double), ("norm_exp", C.c_double),
                ("cutoff", C.c_double), ("cut", C.c_double)]


class _EeqModel(C.Structure):
    _fields_ = [("nid", C.c_int), ("chi", c_dp), ("rad", c_dp), ("eta", c_dp), ("kcnchi", c_dp),
                ("ncoord", _Ncoord)]


class _EeqbcModel(C.Structure):
    _fields_ = [("nid", C.c_int), ("chi", c_dp), ("rad", c_dp), ("eta", c_dp), ("kcnchi", c_dp),
                ("kqchi", c_dp), ("kqeta", c_dp), ("cap", c_dp), ("avg_cn", c_dp), ("rvdw", c_dp),
                ("kcnrad", C.c_double), ("kbc", C.c_double), ("norm_exp", C.c_double),
                ("ncoord", _Ncoord), ("ncoord_en", _Ncoord)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_ewald_alpha.restype = C.c_double
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def max_threads():
    return int(lib().orc_get_max_threads())


class Mol:
    """Minimal structure_type: num[nid], id[nat] (1-based), xyz[nat,3] Bohr, charge, lattice, periodic."""

    def __init__(self, num, xyz, charge=0.0, lattice=None, periodic=None, id=None):
        xyz = _f64(xyz).reshape(-1, 3)
        num = np.asarray(num, dtype=np.int32)
        if id is None:  # per-atom atomic numbers given: derive species like mctc-lib's new()
            uniq = []
            for z in num.tolist():
                if z not in uniq:
                    uniq.append(z)
            self.id = np.array([uniq.index(z) + 1 for z in num.tolist()], dtype=np.int32)
            self.num = np.array(uniq, dtype=np.int32)
        else:
            self.id = np.ascontiguousarray(id, dtype=np.int32)
            self.num = num
        self.xyz = xyz
        self.nat = xyz.shape[0]
        self.nid = len(self.num)
        self.charge = float(charge)
        self.lattice = np.zeros((3, 3)) if lattice is None else _f64(lattice).reshape(3, 3)
        if periodic is None:
            periodic = [lattice is not None] * 3
        self.periodic = np.array([bool(p) for p in np.broadcast_to(periodic, (3,))])
        self._c = _Mol(self.nat, self.nid, self.id.ctypes.data_as(c_ip), _dp(self.xyz), self.charge,
                       (C.c_int * 3)(*[int(p) for p in self.periodic]),
                       (C.c_double * 9)(*self.lattice.ravel().tolist()))

    def cref(self):
        return C.byref(self._c)


def _ncoord(rcov, en, kcn, norm_exp, cutoff, cut):
    return _Ncoord(_dp(rcov), _dp(en), kcn, norm_exp, cutoff, cut)


class EeqModel:
    """eeq_model (model/eeq.f90:42-95)."""

    def __init__(self, chi, rad, eta, kcnchi, rcov, cutoff=25.0, cn_exp=7.5, cn_max=8.0):
        self.chi, self.rad, self.eta, self.kcnchi, self.rcov = (_f64(a) for a in (chi, rad, eta, kcnchi, rcov))
        self.cutoff, self.kcn, self.cn_max = cutoff, cn_exp, cn_max
        self._c = _EeqModel(len(self.chi), _dp(self.chi), _dp(self.rad), _dp(self.eta), _dp(self.kcnchi),
                            _ncoord(self.rcov, None, cn_exp, 1.0, cutoff, cn_max))

    def cref(self):
        return C.byref(self._c)


def param_eeq2019(z):
    out = np.zeros(5)
    lib().orc_param_eeq2019(int(z), _dp(out))
    return out  # chi, eta, kcnchi, rad, rcov


def param_eeqbc2025(z):
    out = np.zeros(9)
    lib().orc_param_eeqbc2025(int(z), _dp(out))
    return out  # chi, eta, rad, kcnchi, kqchi, kqeta, cap, cov_radii, avg_cn


def new_eeq2019_model(mol):
    """param.f90:46-73."""
    p = np.array([param_eeq2019(z) for z in mol.num])
    return EeqModel(chi=p[:, 0], eta=p[:, 1], kcnchi=p[:, 2], rad=p[:, 3], rcov=p[:, 4],
                    cutoff=25.0, cn_exp=7.5, cn_max=8.0)


# ---------------- mctc-lib pieces ----------------

def lattice_points_cutoff(mol, cutoff):
    ptr = c_dp()
    ntr = lib().orc_lattice_points_cutoff(mol._c.periodic, mol._c.lattice, C.c_double(cutoff), C.byref(ptr))
    out = np.ctypeslib.as_array(ptr, shape=(ntr, 3)).copy()
    lib().orc_free(ptr)
    return out


def dir_trans(lattice):
    out = np.zeros((125, 3))
    lib().orc_get_dir_trans(_dp(_f64(lattice)), _dp(out))
    return out


def rec_trans(lattice):
    out = np.zeros((124, 3))
    lib().orc_get_rec_trans(_dp(_f64(lattice)), _dp(out))
    return out


def ewald_alpha(lattice):
    return float(lib().orc_ewald_alpha(_dp(_f64(lattice))))


def wsc(mol):
    """new_wignerseitz_cell: returns (trans[ntr,3], nimg[iat,jat], tridx[iat,jat,ntr])."""

    class _Wsc(C.Structure):
        _fields_ = [("ntr", C.c_int), ("trans", c_dp), ("nimg", c_ip), ("tridx", c_ip), ("nimg_max", C.c_int)]

    w = _Wsc()
    lib().orc_wsc_new(mol.cref(), C.byref(w))
    n = mol.nat
    trans = np.ctypeslib.as_array(w.trans, shape=(w.ntr, 3)).copy()
    nimg = np.ctypeslib.as_array(w.nimg, shape=(n, n)).copy()
    tridx = np.ctypeslib.as_array(w.tridx, shape=(n, n, w.ntr)).copy()
    lib().orc_wsc_free(C.byref(w))
    return trans, nimg, tridx


def get_cn(nc_model, mol, grad=False, trans=None, which="ncoord"):
    """ncoord%get_coordination_number(mol, trans, cn[, dcndr, dcndL])."""
    ncs = getattr(nc_model._c, which)
    if trans is None:
        trans = lattice_points_cutoff(mol, ncs.cutoff)
    trans = _f64(trans)
    n = mol.nat
    cn = np.zeros(n)
    dcndr = np.zeros((n, n, 3)) if grad else None
    dcndL = np.zeros((n, 3, 3)) if grad else None
    lib().orc_ncoord_get_cn(C.byref(ncs), mol.cref(), C.c_int(len(trans)), _dp(trans), _dp(cn), _dp(dcndr), _dp(dcndL))
    return (cn, dcndr, dcndL) if grad else cn


# ---------------- EEQ2019 ----------------

def eeq_get_xvec(model, mol, cn):
    x = np.zeros(mol.nat + 1)
    lib().orc_eeq_get_xvec(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(x))
    return x


def eeq_get_xvec_derivs(model, mol, cn, dcndr, dcndL):
    n = mol.nat
    dxdr = np.zeros((n + 1, n, 3))
    dxdL = np.zeros((n + 1, 3, 3))
    lib().orc_eeq_get_xvec_derivs(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(_f64(dcndr)), _dp(_f64(dcndL)),
                                  _dp(dxdr), _dp(dxdL))
    return dxdr, dxdL


def eeq_get_coulomb_matrix(model, mol):
    a = np.zeros((mol.nat + 1, mol.nat + 1))
    lib().orc_eeq_get_coulomb_matrix(model.cref(), mol.cref(), _dp(a))
    return a


def eeq_get_coulomb_derivs(model, mol, qvec):
    n = mol.nat
    dadr = np.zeros((n + 1, n, 3))
    dadL = np.zeros((n + 1, 3, 3))
    atrace = np.zeros((n, 3))
    lib().orc_eeq_get_coulomb_derivs(model.cref(), mol.cref(), _dp(_f64(qvec)), _dp(dadr), _dp(dadL), _dp(atrace))
    return dadr, dadL, atrace


class SolveResult(dict):
    __getattr__ = dict.get


def _outputs(n, want):
    out = SolveResult()
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
    return out


def eeq_solve(model, mol, cn, dcndr=None, dcndL=None, want=("qvec", "energy", "gradient", "dqdr")):
    """mchrg_model_type%solve for EEQ2019 (energy / sigma start from zero here)."""
    out = _outputs(mol.nat, want)
    stat = lib().orc_eeq_solve(model.cref(), mol.cref(), _dp(_f64(cn)), None,
                               _dp(None if dcndr is None else _f64(dcndr)),
                               _dp(None if dcndL is None else _f64(dcndL)),
                               _dp(out.energy), _dp(out.gradient), _dp(out.sigma),
                               _dp(out.qvec), _dp(out.dqdr), _dp(out.dqdL))
    out["stat"] = stat
    return out


def eeq_get_charges(model, mol, grad=False):
    out = _outputs(mol.nat, ("qvec", "dqdr") if grad else ("qvec",))
    out["stat"] = lib().orc_eeq_get_charges(model.cref(), mol.cref(), _dp(out.qvec), _dp(out.dqdr), _dp(out.dqdL))
    return out


def eeq_singlepoint(model, mol, want=("qvec", "energy", "gradient")):
    """CN -> solve, the app/main.f90:114-118 sequence."""
    out = _outputs(mol.nat, want)
    out["cn"] = np.zeros(mol.nat)
    out["stat"] = lib().orc_eeq_singlepoint(model.cref(), mol.cref(), _dp(out.cn), _dp(out.qvec), _dp(out.energy),
                                            _dp(out.gradient), _dp(out.sigma), _dp(out.dqdr), _dp(out.dqdL))
    return out


def eeq2019_singlepoint_batch(offsets, num, xyz, charge, gradient=True, threads=None):
    """CPU baseline for the batch workload: one reference single point per molecule, molecules spread
    over `threads` host threads (default: all)."""
    L = lib()
    nthr = int(threads or os.cpu_count())
    L.orc_set_num_threads(nthr)
    L.scipy_openblas_set_num_threads(1)  # molecule-level parallelism; tiny LAPACK calls stay serial
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    num = np.ascontiguousarray(num, dtype=np.int32)
    xyz = _f64(xyz)
    charge = _f64(charge)
    ntot, nmol = int(offsets[-1]), len(charge)
    out = SolveResult(qvec=np.zeros(ntot), energy=np.zeros(ntot), gradient=np.zeros((ntot, 3)),
                      status=np.zeros(nmol, dtype=np.int32))
    L.orc_eeq2019_singlepoint_batch(C.c_int(nmol), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                    num.ctypes.data_as(c_ip), _dp(xyz), _dp(charge), _dp(out.qvec), _dp(out.energy),
                                    _dp(out.gradient), out.status.ctypes.data_as(c_ip), C.c_int(1 if gradient else 0))
    out["threads"] = nthr
    return out


# ---------------- EEQ-BC (non-periodic) ----------------

class EeqbcModel:
    """eeqbc_model (model/eeqbc.f90:56-176); rvdw is per species pair [nid, nid]."""

    def __init__(self, chi, rad, eta, kcnchi, kqchi, kqeta, kcnrad, cap, avg_cn, rvdw, rcov, en,
                 kbc=0.65, cutoff=25.0, cn_exp=2.0, norm_exp=1.0, cn_max=-1.0):
        (self.chi, self.rad, self.eta, self.kcnchi, self.kqchi, self.kqeta, self.cap, self.avg_cn, self.rvdw,
         self.rcov, self.en) = (_f64(a) for a in (chi, rad, eta, kcnchi, kqchi, kqeta, cap, avg_cn, rvdw, rcov, en))
        nid = len(self.chi)
        assert self.rvdw.shape == (nid, nid)
        self.kcnrad, self.kbc, self.norm_exp, self.cutoff, self.kcn, self.cn_max = kcnrad, kbc, norm_exp, cutoff, cn_exp, cn_max
        self._c = _EeqbcModel(nid, _dp(self.chi), _dp(self.rad), _dp(self.eta), _dp(self.kcnchi), _dp(self.kqchi),
                              _dp(self.kqeta), _dp(self.cap), _dp(self.avg_cn), _dp(self.rvdw), kcnrad, kbc, norm_exp,
                              _ncoord(self.rcov, None, cn_exp, norm_exp, cutoff, cn_max),
                              _ncoord(self.rcov, self.en, cn_exp, norm_exp, cutoff, cn_max))

    def cref(self):
        return C.byref(self._c)


def synthetic_vdw_en(num):
    """Deterministic stand-ins for the mctc-lib tables that are not available offline (pairwise D3 vdW radii
    x autoaa, Pauling EN / 3.98): NOT the reference's numbers, only fixed inputs shared by oracle and device."""
    z = np.asarray(num, dtype=np.float64)
    r = 1.1 + 0.35 * np.log1p(z)  # "Angstrom" scale, grows slowly with Z
    rvdw = 0.5 * (r[:, None] + r[None, :]) * 0.529177210903 * 3.2
    en = (0.8 + 2.6 * np.exp(-((z % 8) - 6.5) ** 2 / 18.0)) / 3.98
    return np.ascontiguousarray(rvdw), en


def new_eeqbc2025_model(mol, rvdw=None, en=None):
    """param.f90:75-125 with the tables of param/eeqbc2025.f90; rvdw/en default to the synthetic stand-ins."""
    p = np.array([param_eeqbc2025(z) for z in mol.num])
    if rvdw is None or en is None:
        rvdw, en = synthetic_vdw_en(mol.num)
    return EeqbcModel(chi=p[:, 0], eta=p[:, 1], rad=p[:, 2], kcnchi=p[:, 3], kqchi=p[:, 4], kqeta=p[:, 5], cap=p[:, 6],
                      rcov=p[:, 7], avg_cn=p[:, 8], rvdw=rvdw, en=en, kcnrad=0.14, kbc=0.60, cutoff=25.0, cn_exp=2.0,
                      norm_exp=0.75)


def eeqbc_local_charge(model, mol, grad=False):
    trans = lattice_points_cutoff(mol, model.cutoff)
    n = mol.nat
    qloc = np.zeros(n)
    dqr = np.zeros((n, n, 3)) if grad else None
    dql = np.zeros((n, 3, 3)) if grad else None
    lib().orc_eeqbc_local_charge(model.cref(), mol.cref(), C.c_int(len(trans)), _dp(trans), _dp(qloc), _dp(dqr), _dp(dql))
    return (qloc, dqr, dql) if grad else qloc


def eeqbc_get_xvec(model, mol, cn, qloc):
    x = np.zeros(mol.nat + 1)
    lib().orc_eeqbc_get_xvec(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(_f64(qloc)), _dp(x))
    return x


def eeqbc_get_coulomb_matrix(model, mol, cn, qloc):
    a = np.zeros((mol.nat + 1, mol.nat + 1))
    lib().orc_eeqbc_get_coulomb_matrix(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(_f64(qloc)), _dp(a))
    return a


def eeqbc_get_cmat(model, mol, grad=False):
    n = mol.nat
    cmat = np.zeros((n + 1, n + 1))
    dcdr = np.zeros((n + 1, n, 3)) if grad else None
    dcdL = np.zeros((n + 1, 3, 3)) if grad else None
    lib().orc_eeqbc_get_cmat(model.cref(), mol.cref(), _dp(cmat), _dp(dcdr), _dp(dcdL))
    return (cmat, dcdr, dcdL) if grad else cmat


def eeqbc_get_xvec_derivs(model, mol, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL):
    n = mol.nat
    dxdr = np.zeros((n + 1, n, 3))
    dxdL = np.zeros((n + 1, 3, 3))
    lib().orc_eeqbc_get_xvec_derivs(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(_f64(qloc)), _dp(_f64(dcndr)),
                                    _dp(_f64(dcndL)), _dp(_f64(dqlocdr)), _dp(_f64(dqlocdL)), _dp(dxdr), _dp(dxdL))
    return dxdr, dxdL


def eeqbc_get_coulomb_derivs(model, mol, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, qvec):
    n = mol.nat
    dadr = np.zeros((n + 1, n, 3))
    dadL = np.zeros((n + 1, 3, 3))
    atrace = np.zeros((n, 3))
    lib().orc_eeqbc_get_coulomb_derivs(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(_f64(qloc)), _dp(_f64(dcndr)),
                                       _dp(_f64(dcndL)), _dp(_f64(dqlocdr)), _dp(_f64(dqlocdL)), _dp(_f64(qvec)),
                                       _dp(dadr), _dp(dadL), _dp(atrace))
    return dadr, dadL, atrace


def eeqbc_solve(model, mol, cn, qloc, dcndr=None, dcndL=None, dqlocdr=None, dqlocdL=None,
                want=("qvec", "energy", "gradient", "dqdr")):
    out = _outputs(mol.nat, want)
    opt = lambda a: _dp(None if a is None else _f64(a))  # noqa: E731
    out["stat"] = lib().orc_eeqbc_solve(model.cref(), mol.cref(), _dp(_f64(cn)), _dp(_f64(qloc)), opt(dcndr), opt(dcndL),
                                        opt(dqlocdr), opt(dqlocdL), _dp(out.energy), _dp(out.gradient), _dp(out.sigma),
                                        _dp(out.qvec), _dp(out.dqdr), _dp(out.dqdL))
    return out


def eeqbc_singlepoint(model, mol, want=("qvec", "energy", "gradient")):
    out = _outputs(mol.nat, want)
    out["cn"] = np.zeros(mol.nat)
    out["qloc"] = np.zeros(mol.nat)
    out["stat"] = lib().orc_eeqbc_singlepoint(model.cref(), mol.cref(), _dp(out.cn), _dp(out.qloc), _dp(out.qvec),
                                              _dp(out.energy), _dp(out.gradient), _dp(out.sigma), _dp(out.dqdr), _dp(out.dqdL))
    return out
