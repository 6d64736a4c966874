"""Parity of the CUDA EEQ2019 path against the CPU oracle and the reference's golden vectors.

Everything goes through the C ABI (multicharge_b200 -> libmchrg_b200.so).  Tolerance: 1e-10 relative
(BASELINE.json north_star) -- the observed agreement is ~1e-13."""
import numpy as np
import pytest

from conftest import GEOM_CASES, GOLDEN_CASES, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def both(mc, oracle, num, xyz, charge, lattice=None):
    mol = mc.Structure(num, xyz, charge, lattice=lattice)
    omol = oracle.Mol(num, xyz, charge, lattice=lattice)
    return mol, mc.new_eeq2019_model(mol), omol, oracle.new_eeq2019_model(omol)


def small_cases():
    from synth import cluster
    cases = []
    for name in GOLDEN_CASES + GEOM_CASES:
        g = load_golden(name)
        cases.append((name, g["num"], np.array(g["xyz"]), g["charge"]))
    for nat, seed, chg in ((1, 1, 0.0), (2, 2, 1.0), (16, 3, 0.0), (43, 4, -1.0), (127, 5, 0.0), (128, 6, 1.0),
                           (129, 7, 0.0), (200, 8, 0.0), (300, 9, 2.0)):
        num, xyz = cluster(nat, seed)
        cases.append((f"cluster{nat}", num, xyz, chg))
    return cases


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_pieces_match_oracle(mc, oracle, case):
    _, num, xyz, charge = case
    mol, model, omol, omodel = both(mc, oracle, num, xyz, charge)
    n = mol.nat
    cn, dcndr, dcndL = model.get_coordination_number(mol, grad=True)
    ocn, odcndr, odcndL = oracle.get_cn(omodel, omol, grad=True)
    assert rel_err(cn, ocn) < TOL
    assert np.abs(dcndr - odcndr).max() < TOL * max(1.0, np.abs(odcndr).max())
    assert np.abs(dcndL - odcndL).max() < TOL * max(1.0, np.abs(odcndL).max())
    assert rel_err(model.get_coordination_number(mol), ocn) < TOL
    assert rel_err(model.get_xvec(mol, ocn), oracle.eeq_get_xvec(omodel, omol, ocn)) < TOL
    assert rel_err(model.get_coulomb_matrix(mol), oracle.eeq_get_coulomb_matrix(omodel, omol)) < 1e-13
    dxdr, dxdL = model.get_xvec_derivs(mol, ocn, odcndr, odcndL)
    odxdr, odxdL = oracle.eeq_get_xvec_derivs(omodel, omol, ocn, odcndr, odcndL)
    assert np.abs(dxdr - odxdr).max() < TOL * max(1.0, np.abs(odxdr).max())
    assert np.abs(dxdL - odxdL).max() < TOL * max(1.0, np.abs(odxdL).max())
    q = np.random.default_rng(n).standard_normal(n)
    dadr, dadL, atrace = model.get_coulomb_derivs(mol, q)
    odadr, odadL, oatrace = oracle.eeq_get_coulomb_derivs(omodel, omol, q)
    for a, b in ((dadr, odadr), (dadL, odadL), (atrace, oatrace)):
        assert np.abs(a - b).max() < TOL * max(1.0, np.abs(b).max())
    qloc = model.local_charge(mol)
    assert np.allclose(qloc, charge / n, rtol=0, atol=1e-15)


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c[0])
def test_solve_matches_oracle(mc, oracle, case):
    _, num, xyz, charge = case
    mol, model, omol, omodel = both(mc, oracle, num, xyz, charge)
    n = mol.nat
    ocn, odcndr, odcndL = oracle.get_cn(omodel, omol, grad=True)
    ref = oracle.eeq_solve(omodel, omol, ocn, odcndr, odcndL)
    assert ref.stat == 0
    # accumulate-into semantics of energy / sigma, overwrite semantics of gradient (type.F90:261,276,279)
    e0 = np.full(n, 0.25)
    s0 = np.full((3, 3), -0.5)
    out = dict(energy=e0.copy(), gradient=np.full((n, 3), 7.0), sigma=s0.copy(), qvec=np.zeros(n),
               dqdr=np.zeros((n, n, 3)), dqdL=np.zeros((n, 3, 3)))
    model.solve(mol, ocn, None, odcndr, odcndL, **out)
    assert rel_err(out["qvec"], ref.qvec) < TOL
    assert rel_err(out["energy"] - e0, ref.energy) < TOL
    assert np.abs(out["gradient"] - ref.gradient).max() < TOL * max(1.0, np.abs(ref.gradient).max())
    assert np.abs(out["sigma"] - s0 - ref.sigma).max() < TOL * max(1.0, np.abs(ref.sigma).max())
    assert np.abs(out["dqdr"] - ref.dqdr).max() < TOL * max(1.0, np.abs(ref.dqdr).max())
    assert np.abs(out["dqdL"] - ref.dqdL).max() < TOL * max(1.0, np.abs(ref.dqdL).max())
    # charges-only call (sytrs branch) and total charge
    q = np.zeros(n)
    model.solve(mol, ocn, qvec=q)
    assert rel_err(q, ref.qvec) < TOL
    assert abs(q.sum() - charge) < 1e-10


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(mc, name):
    """The reference's own numbers (validation JSONs at their official 5e-7; we hold 5e-11)."""
    g = load_golden(name)
    mol = mc.Structure(g["num"], g["xyz"], g["charge"])
    model = mc.new_eeq2019_model(mol)
    r = model.singlepoint(mol)
    assert np.abs(r.qvec - np.array(g["charges"])).max() < 5e-11
    if "cn" in g:
        assert np.abs(r.cn - np.array(g["cn"])).max() < 2e-10
        assert abs(r.energy.sum() - g["energy"]) < 5e-11
        assert np.abs(r.gradient.ravel() - np.array(g["gradient"])).max() < 5e-11
    q, dqdr, dqdL = mc.get_charges(model, mol, grad=True)
    assert np.abs(q - r.qvec).max() < 1e-13
    assert np.abs(mc.get_eeq_charges(mol) - r.qvec).max() < 1e-12  # (charges only: the fused one-launch kernel)


def test_medium_cluster_full_path(mc, oracle):
    from synth import cluster
    num, xyz = cluster(1500, 11)
    mol, model, omol, omodel = both(mc, oracle, num, xyz, 0.0)
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    o = oracle.eeq_singlepoint(omodel, omol, want=("qvec", "energy", "gradient", "dqdr"))
    for key in ("cn", "qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert np.abs(r[key] - o[key]).max() < TOL * max(1.0, np.abs(o[key]).max()), key


@pytest.mark.parametrize("nat", [2, 40, 300])
def test_indefinite_coulomb_block_is_solved_like_dsytrf(mc, oracle, monkeypatch, nat):
    """Two hydrogens 0.5 Bohr apart make the Coulomb block indefinite on the charge-conserving subspace as well
    (SURVEY 7 'hard parts').  The reference's Bunch-Kaufman LDL^T (lapack.F90:165-201) factors the bordered matrix
    anyway; here the shifted Cholesky attempts fail and the solver falls back to the normal equations of the bordered
    system (api.cu, factor_solve).  Charges, energy, gradient and sigma against the oracle (dsytrf / dsytrs); dq/dr is
    refused on that route; with the fallback switched off the failed pivot is reported like fatal_error would."""
    from synth import cluster
    if nat == 2:
        num, xyz = np.array([1, 1]), np.array([[0.0, 0.0, 0.0], [0.5, 0.0, 0.0]])
    else:
        num, xyz = cluster(nat, 50 + nat)
        num[:2] = 1
        xyz[1] = xyz[0] + np.array([0.5, 0.0, 0.0])
    mol, model, omol, omodel = both(mc, oracle, num, xyz, 0.0)
    o = oracle.eeq_singlepoint(omodel, omol, want=("qvec", "energy", "gradient"))
    assert o.stat == 0
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    for key in ("qvec", "energy", "gradient", "sigma"):
        assert np.abs(r[key] - o[key]).max() < 1e-9 * max(1.0, np.abs(o[key]).max()), key
    assert abs(r["qvec"].sum()) < 1e-10
    with pytest.raises(mc.MchrgError):
        model.singlepoint(mol, want=("qvec", "dqdr"))
    monkeypatch.setenv("MCHRG_B200_NO_INDEFINITE", "1")
    with pytest.raises(mc.FactorizationError):
        model.singlepoint(mol)


def test_device_resident_inputs(mc, oracle):
    """Pointers may be device pointers (torch CUDA tensors): same results as host arrays."""
    import torch
    from synth import cluster
    num, xyz = cluster(257, 21)
    mol, model, omol, omodel = both(mc, oracle, num, xyz, 0.0)
    o = oracle.eeq_singlepoint(omodel, omol)
    dev = torch.device("cuda:0")
    txyz = torch.tensor(mol.xyz, device=dev)
    tid = torch.tensor(mol.id, device=dev)
    r = model.singlepoint(mol, xyz=txyz, id=tid)
    for key in ("qvec", "energy", "gradient"):
        assert np.abs(r[key] - o[key]).max() < TOL * max(1.0, np.abs(o[key]).max()), key


def test_gradient_without_dcndr(mc, oracle, monkeypatch):
    """Single points without dq/dr contract the CN derivatives pair by pair (no (3, N, N) array): gradient and
    sigma against the oracle, against the dense-dcndr route (MCHRG_B200_CN_DERIVS=dense), and for a row block."""
    from synth import cluster
    num, xyz = cluster(777, 31)
    mol, model, omol, omodel = both(mc, oracle, num, xyz, -1.0)
    o = oracle.eeq_singlepoint(omodel, omol, want=("qvec", "energy", "gradient"))
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    rows = mc.singlepoint_rows(model, mol, 100, 333, dqdr=False)
    monkeypatch.setenv("MCHRG_B200_CN_DERIVS", "dense")
    d = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    for key in ("qvec", "energy", "gradient", "sigma"):
        scale = max(1.0, np.abs(o[key]).max())
        assert np.abs(r[key] - o[key]).max() < TOL * scale, key
        assert np.abs(r[key] - d[key]).max() < 1e-12 * scale, key
    assert np.abs(rows["gradient"] - o.gradient[100:333]).max() < TOL
    assert np.abs(rows["sigma"] - o.sigma).max() < TOL * max(1.0, np.abs(o.sigma).max())


def test_small_charges_only_route(mc, oracle, monkeypatch):
    """get_charges without derivatives on a molecule of at most 208 atoms runs as ONE launch of the fused batch kernel;
    same charges as the general path (MCHRG_B200_SMALL_FUSED=0) and the oracle, species in order of appearance."""
    from synth import cluster
    for nat, seed, chg in ((1, 41, 0.0), (16, 42, 1.0), (97, 43, -1.0), (208, 44, 0.0), (209, 45, 0.0)):
        num, xyz = cluster(nat, seed)
        mol, model, omol, omodel = both(mc, oracle, num, xyz, chg)
        o = oracle.eeq_singlepoint(omodel, omol, want=("qvec",))
        n0 = mc.default_context().launch_count
        q = mc.get_charges(model, mol)
        fused_launches = mc.default_context().launch_count - n0
        monkeypatch.setenv("MCHRG_B200_SMALL_FUSED", "0")
        n0 = mc.default_context().launch_count
        q_general = mc.get_charges(model, mol)
        general_launches = mc.default_context().launch_count - n0
        monkeypatch.delenv("MCHRG_B200_SMALL_FUSED")
        assert np.abs(q - o.qvec).max() < TOL and np.abs(q - q_general).max() < 1e-12
        assert (fused_launches == 1) == (nat <= 208), (nat, fused_launches, general_launches)
        assert general_launches > 10
