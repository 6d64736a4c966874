"""Parity of the EEQ-BC device path (molecular and periodic) against the CPU oracle, 1e-10 relative.

The oracle's EEQ-BC branch is "parity unpinned" against reference numbers (tests/test_oracle_eeqbc.py):
the mctc-lib vdW / EN tables are replaced by fixed synthetic stand-ins shared by both sides."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-10


def close(a, b, tol=TOL):
    return np.abs(a - b).max() < tol * max(1.0, np.abs(b).max())


def cases():
    from synth import cluster
    out = []
    g = load_golden("geom_znooh")
    out.append(("znooh", np.array(g["num"]), np.array(g["xyz"]), g["charge"]))
    g = load_golden("kat_eeq_actinides")
    out.append(("actinides", np.array(g["num"]), np.array(g["xyz"]), 0.0))
    for nat, seed, chg in ((2, 31, 1.0), (9, 21, 1.0), (40, 32, 0.0), (129, 33, -1.0), (300, 34, 0.0)):
        num, xyz = cluster(nat, seed)
        out.append((f"cluster{nat}", num, xyz, chg))
    return out


def models(mc, oracle, num, xyz, chg):
    mol = mc.Structure(num, xyz, chg)
    omol = oracle.Mol(num, xyz, chg)
    omodel = oracle.new_eeqbc2025_model(omol)
    # identical parameters on both sides: use the generic constructor with the oracle's arrays
    model = mc.new_eeqbc_model(mol, chi=omodel.chi, rad=omodel.rad, eta=omodel.eta, kcnchi=omodel.kcnchi, kqchi=omodel.kqchi,
                               kqeta=omodel.kqeta, kcnrad=omodel.kcnrad, cap=omodel.cap, avg_cn=omodel.avg_cn, rvdw=omodel.rvdw,
                               rcov=omodel.rcov, en=omodel.en, kbc=omodel.kbc, cutoff=omodel.cutoff, cn_exp=omodel.kcn,
                               norm_exp=omodel.norm_exp)
    return mol, model, omol, omodel


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_eeqbc_pieces(mc, oracle, case):
    _, num, xyz, chg = case
    mol, model, omol, omodel = models(mc, oracle, num, xyz, chg)
    n = mol.nat
    cn, dcndr, dcndL = model.get_coordination_number(mol, grad=True)
    ocn, odcndr, odcndL = oracle.get_cn(omodel, omol, grad=True)
    assert close(cn, ocn) and close(dcndr, odcndr) and close(dcndL, odcndL)
    qloc, dqr, dql = model.local_charge(mol, grad=True)
    oqloc, odqr, odql = oracle.eeqbc_local_charge(omodel, omol, grad=True)
    assert close(qloc, oqloc) and close(dqr, odqr) and close(dql, odql)
    assert close(model.get_xvec(mol, ocn, oqloc), oracle.eeqbc_get_xvec(omodel, omol, ocn, oqloc))
    assert close(model.get_coulomb_matrix(mol, ocn, oqloc), oracle.eeqbc_get_coulomb_matrix(omodel, omol, ocn, oqloc), 1e-12)
    dxdr, dxdL = model.get_xvec_derivs(mol, ocn, odcndr, odcndL, oqloc, odqr, odql)
    odxdr, odxdL = oracle.eeqbc_get_xvec_derivs(omodel, omol, ocn, oqloc, odcndr, odcndL, odqr, odql)
    assert close(dxdr, odxdr) and close(dxdL, odxdL)
    q = np.random.default_rng(n).standard_normal(n)
    dadr, dadL, atrace = model.get_coulomb_derivs(mol, q, ocn, oqloc, odcndr, odcndL, odqr, odql)
    odadr, odadL, oatrace = oracle.eeqbc_get_coulomb_derivs(omodel, omol, ocn, oqloc, odcndr, odcndL, odqr, odql, q)
    assert close(dadr, odadr) and close(dadL, odadL) and close(atrace, oatrace)


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_eeqbc_solve(mc, oracle, case):
    _, num, xyz, chg = case
    mol, model, omol, omodel = models(mc, oracle, num, xyz, chg)
    n = mol.nat
    ocn, odcndr, odcndL = oracle.get_cn(omodel, omol, grad=True)
    oqloc, odqr, odql = oracle.eeqbc_local_charge(omodel, omol, grad=True)
    ref = oracle.eeqbc_solve(omodel, omol, ocn, oqloc, odcndr, odcndL, odqr, odql)
    assert ref.stat == 0
    out = dict(energy=np.zeros(n), gradient=np.zeros((n, 3)), sigma=np.zeros((3, 3)), qvec=np.zeros(n),
               dqdr=np.zeros((n, n, 3)), dqdL=np.zeros((n, 3, 3)))
    model.solve(mol, ocn, oqloc, odcndr, odcndL, odqr, odql, **out)
    for key in ("qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(out[key], ref[key]), key
    # CLI sequence and get_charges
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    o = oracle.eeqbc_singlepoint(omodel, omol, want=("qvec", "energy", "gradient", "dqdr"))
    for key in ("cn", "qloc", "qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(r[key], o[key]), key
    q, dqdr, dqdL = mc.get_charges(model, mol, grad=True)
    assert close(q, o.qvec) and close(dqdr, o.dqdr) and close(dqdL, o.dqdL)
    q2 = np.zeros(n)
    model.solve(mol, ocn, oqloc, qvec=q2)
    assert close(q2, ref.qvec)


def test_eeqbc_requires_qloc(mc, oracle):
    _, num, xyz, chg = cases()[3]
    mol, model, omol, omodel = models(mc, oracle, num, xyz, chg)
    cn = oracle.get_cn(omodel, omol)
    with pytest.raises(mc.MchrgError) as exc:
        model.solve(mol, cn, None, qvec=np.zeros(mol.nat))
    assert "qloc required" in str(exc.value)


def test_eeqbc2025_constructor_matches_generic(mc, oracle):
    from synth import cluster
    num, xyz = cluster(25, 3)
    mol = mc.Structure(num, xyz, 0.0)
    omol = oracle.Mol(num, xyz, 0.0)
    rvdw, en = oracle.synthetic_vdw_en(mol.num)
    model = mc.new_eeqbc2025_model(mol, rvdw, en * 3.98)
    o = oracle.eeqbc_singlepoint(oracle.new_eeqbc2025_model(omol), omol)
    r = model.singlepoint(mol)
    assert close(r.qvec, o.qvec) and close(r.energy, o.energy) and close(r.gradient, o.gradient)


def pbc_cells():
    from synth import periodic_cell
    out = []
    for ncell, seed, shear, chg in ((2, 3, 0.0, 0.0), (2, 4, 0.1, 1.0), (3, 5, 0.1, 0.0), (4, 6, 0.0, -1.0)):
        num, xyz, lat = periodic_cell(ncell, seed, shear=shear)
        out.append((f"cell{ncell}_{'tri' if shear else 'cub'}", num, xyz, lat, chg))
    num, xyz, lat = periodic_cell(3, 8, shear=0.05)
    out.append(("cell3_shifted", num, xyz + 2 * lat[0] - lat[2], lat, 0.0))
    return out


@pytest.mark.parametrize("case", pbc_cells(), ids=lambda c: c[0])
def test_eeqbc_periodic(mc, oracle, case):
    """Periodic EEQ-BC (get_cmat_3d, get_amat_3d, get_damat_3d, get_dcmat_3d, periodic get_xvec / _derivs): device vs oracle."""
    _, num, xyz, lat, chg = case
    mol = mc.Structure(num, xyz, chg, lattice=lat)
    omol = oracle.Mol(num, xyz, chg, lattice=lat)
    omodel = oracle.new_eeqbc2025_model(omol)
    model = mc.new_eeqbc_model(mol, chi=omodel.chi, rad=omodel.rad, eta=omodel.eta, kcnchi=omodel.kcnchi, kqchi=omodel.kqchi,
                               kqeta=omodel.kqeta, kcnrad=omodel.kcnrad, cap=omodel.cap, avg_cn=omodel.avg_cn, rvdw=omodel.rvdw,
                               rcov=omodel.rcov, en=omodel.en, kbc=omodel.kbc, cutoff=omodel.cutoff, cn_exp=omodel.kcn,
                               norm_exp=omodel.norm_exp)
    n = mol.nat
    ocn = oracle.get_cn(omodel, omol)
    oqloc = oracle.eeqbc_local_charge(omodel, omol)
    assert close(model.get_coordination_number(mol), ocn) and close(model.local_charge(mol), oqloc)
    assert close(model.get_xvec(mol, ocn, oqloc), oracle.eeqbc_get_xvec(omodel, omol, ocn, oqloc))
    assert close(model.get_coulomb_matrix(mol, ocn, oqloc), oracle.eeqbc_get_coulomb_matrix(omodel, omol, ocn, oqloc), 1e-12)
    ref = oracle.eeqbc_solve(omodel, omol, ocn, oqloc, want=("qvec", "energy"))
    assert ref.stat == 0
    q, e = np.zeros(n), np.zeros(n)
    model.solve(mol, ocn, oqloc, qvec=q, energy=e)
    assert close(q, ref.qvec) and close(e, ref.energy)
    r = model.singlepoint(mol, want=("qvec", "energy"))
    o = oracle.eeqbc_singlepoint(omodel, omol, want=("qvec", "energy"))
    assert close(r.qvec, o.qvec) and close(r.energy, o.energy)
    assert close(mc.get_charges(model, mol), o.qvec)
    # derivatives: get_damat_3d, get_dcmat_3d, periodic get_xvec_derivs
    odcn = oracle.get_cn(omodel, omol, grad=True)
    odq = oracle.eeqbc_local_charge(omodel, omol, grad=True)
    dxdr, dxdL = model.get_xvec_derivs(mol, odcn[0], odcn[1], odcn[2], odq[0], odq[1], odq[2])
    odxdr, odxdL = oracle.eeqbc_get_xvec_derivs(omodel, omol, odcn[0], odq[0], odcn[1], odcn[2], odq[1], odq[2])
    assert close(dxdr, odxdr) and close(dxdL, odxdL)
    qr = np.random.default_rng(n).standard_normal(n)
    dadr, dadL, atrace = model.get_coulomb_derivs(mol, qr, odcn[0], odq[0], odcn[1], odcn[2], odq[1], odq[2])
    odadr, odadL, oatrace = oracle.eeqbc_get_coulomb_derivs(omodel, omol, odcn[0], odq[0], odcn[1], odcn[2], odq[1], odq[2], qr)
    assert close(dadr, odadr) and close(dadL, odadL) and close(atrace, oatrace)
    ref = oracle.eeqbc_solve(omodel, omol, odcn[0], odq[0], odcn[1], odcn[2], odq[1], odq[2])
    assert ref.stat == 0
    out = dict(energy=np.zeros(n), gradient=np.zeros((n, 3)), sigma=np.zeros((3, 3)), qvec=np.zeros(n),
               dqdr=np.zeros((n, n, 3)), dqdL=np.zeros((n, 3, 3)))
    model.solve(mol, odcn[0], odq[0], odcn[1], odcn[2], odq[1], odq[2], **out)
    for key in ("qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(out[key], ref[key]), key
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    o = oracle.eeqbc_singlepoint(omodel, omol, want=("qvec", "energy", "gradient"))
    for key in ("qvec", "energy", "gradient", "sigma"):
        assert close(r[key], o[key]), key
    q, dqdr, dqdL = mc.get_charges(model, mol, grad=True)
    o = oracle.eeqbc_singlepoint(omodel, omol, want=("qvec", "dqdr"))
    assert close(q, o.qvec) and close(dqdr, o.dqdr) and close(dqdL, o.dqdL)
