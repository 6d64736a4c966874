"""Parity of the periodic (Ewald) EEQ2019 device path against the CPU oracle.

Reference numbers for periodic systems need mstore geometries (not available offline), so the oracle's
periodic branch is pinned by the reference's finite-difference methodology only (tests/test_oracle_golden.py);
here the device path must agree with that oracle to 1e-10."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def cells():
    from synth import periodic_cell
    out = []
    for ncell, seed, shear, chg in ((2, 3, 0.0, 0.0), (2, 4, 0.1, 1.0), (3, 5, 0.1, 0.0), (4, 6, 0.0, -1.0), (5, 7, 0.15, 0.0)):
        num, xyz, lat = periodic_cell(ncell, seed, shear=shear)
        out.append((f"cell{ncell}_{'tri' if shear else 'cub'}", num, xyz, lat, chg))
    # atoms shifted far outside the cell and a strongly anisotropic cell
    num, xyz, lat = periodic_cell(3, 8, shear=0.05)
    xyz = xyz + lat[0] * 2 - lat[2]
    lat2 = lat.copy()
    lat2[2] *= 1.7
    out.append(("cell3_aniso_shifted", num, xyz * np.array([1, 1, 1.7]), lat2, 0.0))
    return out


def close(a, b, tol=TOL):
    return np.abs(a - b).max() < tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("case", cells(), ids=lambda c: c[0])
def test_periodic_pieces(mc, oracle, case):
    _, num, xyz, lat, chg = case
    mol = mc.Structure(num, xyz, chg, lattice=lat)
    omol = oracle.Mol(num, xyz, chg, lattice=lat)
    model, omodel = mc.new_eeq2019_model(mol), oracle.new_eeq2019_model(omol)
    cn, dcndr, dcndL = model.get_coordination_number(mol, grad=True)
    ocn, odcndr, odcndL = oracle.get_cn(omodel, omol, grad=True)
    assert close(cn, ocn) and close(dcndr, odcndr) and close(dcndL, odcndL)
    assert close(model.get_coulomb_matrix(mol), oracle.eeq_get_coulomb_matrix(omodel, omol), 1e-12)
    q = np.random.default_rng(1).standard_normal(mol.nat)
    dadr, dadL, atrace = model.get_coulomb_derivs(mol, q)
    odadr, odadL, oatrace = oracle.eeq_get_coulomb_derivs(omodel, omol, q)
    assert close(dadr, odadr) and close(dadL, odadL) and close(atrace, oatrace)


@pytest.mark.parametrize("case", cells(), ids=lambda c: c[0])
def test_periodic_solve(mc, oracle, case):
    _, num, xyz, lat, chg = case
    mol = mc.Structure(num, xyz, chg, lattice=lat)
    omol = oracle.Mol(num, xyz, chg, lattice=lat)
    model, omodel = mc.new_eeq2019_model(mol), oracle.new_eeq2019_model(omol)
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    o = oracle.eeq_singlepoint(omodel, omol, want=("qvec", "energy", "gradient", "dqdr"))
    assert o.stat == 0
    for key in ("cn", "qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(r[key], o[key]), key
    q, dqdr, dqdL = mc.get_charges(model, mol, grad=True)
    assert close(q, o.qvec) and close(dqdr, o.dqdr) and close(dqdL, o.dqdL)
    # charges + energy only (no gradients requested)
    r2 = model.singlepoint(mol, want=("qvec", "energy"))
    assert close(r2.qvec, o.qvec) and close(r2.energy, o.energy)


def test_periodic_medium(mc, oracle):
    from synth import periodic_cell
    num, xyz, lat = periodic_cell(8, 12, shear=0.1)  # 512 atoms
    mol = mc.Structure(num, xyz, 0.0, lattice=lat)
    omol = oracle.Mol(num, xyz, 0.0, lattice=lat)
    r = mc.new_eeq2019_model(mol).singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient", "dqdr"))
    for key in ("qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(r[key], o[key]), key


@pytest.mark.parametrize("case", cells(), ids=lambda c: c[0])
def test_periodic_gradient_without_dcndr(mc, oracle, case, monkeypatch):
    """Single points without dq/dr contract the CN derivatives pair by pair (no (3, N, N) array): gradient and
    sigma against the oracle, and against the dense-dcndr route (MCHRG_B200_CN_DERIVS=dense)."""
    _, num, xyz, lat, chg = case
    mol = mc.Structure(num, xyz, chg, lattice=lat)
    omol = oracle.Mol(num, xyz, chg, lattice=lat)
    model, omodel = mc.new_eeq2019_model(mol), oracle.new_eeq2019_model(omol)
    o = oracle.eeq_singlepoint(omodel, omol, want=("qvec", "energy", "gradient"))
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    monkeypatch.setenv("MCHRG_B200_CN_DERIVS", "dense")
    d = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    for key in ("qvec", "energy", "gradient", "sigma"):
        assert close(r[key], o[key]), key
        assert close(r[key], d[key], 1e-12), key
