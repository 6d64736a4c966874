"""Pin the CPU oracle against the reference's own golden data and against the
reference's finite-difference consistency methodology (test/unit/test_model.f90:89-797).
CPU only -- the oracle is the checker for the -m gpu parity tests."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

THR2 = np.sqrt(np.finfo(np.float64).eps)  # test_model.f90:32


def make(oracle, g):
    mol = oracle.Mol(g["num"], g["xyz"], g["charge"])
    return mol, oracle.new_eeq2019_model(mol)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_known_answers(oracle, name):
    g = load_golden(name)
    mol, model = make(oracle, g)
    r = oracle.eeq_singlepoint(model, mol)
    assert r.stat == 0
    # charges: the fixtures were written by v0.3.0 (older Bohr constant), agreement is ~1e-12
    assert np.abs(r.qvec - np.array(g["charges"])).max() < 5e-12
    if "cn" in g:
        assert np.abs(r.cn - np.array(g["cn"])).max() < 2e-10
        assert abs(r.energy.sum() - g["energy"]) < 2e-12
        assert np.abs(r.gradient.ravel() - np.array(g["gradient"])).max() < 2e-12
    # get_charges gives the same numbers (test_model.f90:883-908)
    q2 = oracle.eeq_get_charges(model, mol).qvec
    assert np.abs(q2 - r.qvec).max() < 1e-14
    assert abs(r.qvec.sum() - g["charge"]) < 1e-12


def test_lattice_point_counts(oracle):
    """test/unit/test_wignerseitz.f90:49-105: 1 (0-D), 27 (cutoff sqrt(eps)); rep=[2,2,2] gives 125/124."""
    lat = np.array([[9.0, 0, 0], [0.5, 8.0, 0], [0.3, 0.2, 10.0]])
    m0 = oracle.Mol([1], [[0, 0, 0]])
    assert oracle.lattice_points_cutoff(m0, 25.0).shape == (1, 3)
    m3 = oracle.Mol([1], [[0, 0, 0]], lattice=lat)
    assert oracle.lattice_points_cutoff(m3, THR2).shape == (27, 3)
    assert np.all(oracle.lattice_points_cutoff(m3, THR2)[0] == 0.0)
    assert oracle.dir_trans(lat).shape == (125, 3) and oracle.rec_trans(lat).shape == (124, 3)
    # reciprocal vectors: G . a_k is a multiple of 2 pi
    ph = oracle.rec_trans(lat) @ lat.T / (2 * np.pi)
    assert np.abs(ph - np.round(ph)).max() < 1e-12


def _fd_cases(oracle):
    from synth import cluster, periodic_cell
    cases = []
    for name in ("geom_h2plus", "geom_znooh"):
        g = load_golden(name)
        cases.append((name, oracle.Mol(g["num"], g["xyz"], g["charge"])))
    num, xyz = cluster(11, 7)
    cases.append(("cluster11", oracle.Mol(num, xyz, 1.0)))
    num, xyz, lat = periodic_cell(2, 3, shear=0.1)
    cases.append(("pbc8", oracle.Mol(num, xyz, 0.0, lattice=lat)))
    return cases


def test_fd_dadr_and_dbdr(oracle):
    """test_dadr / test_dbdr: 2-point stencil on get_coulomb_matrix.q and get_xvec vs the analytic arrays."""
    step = 1.0e-6
    for name, mol in _fd_cases(oracle):
        model = oracle.new_eeq2019_model(mol)
        n = mol.nat
        cn, dcndr, dcndL = oracle.get_cn(model, mol, grad=True)
        q = oracle.eeq_solve(model, mol, cn, want=("qvec",)).qvec
        qf = np.append(q, 0.0)
        dadr, dadL, atrace = oracle.eeq_get_coulomb_derivs(model, mol, q)
        dxdr, dxdL = oracle.eeq_get_xvec_derivs(model, mol, cn, dcndr, dcndL)
        for j in range(min(n, 4)):
            for c in range(3):
                vals = []
                for sgn in (+1, -1):
                    xyz = mol.xyz.copy()
                    xyz[j, c] += sgn * step
                    m2 = oracle.Mol(mol.num, xyz, mol.charge, lattice=mol.lattice if mol.periodic.any() else None, id=mol.id)
                    a = oracle.eeq_get_coulomb_matrix(model, m2)
                    x = oracle.eeq_get_xvec(model, m2, oracle.get_cn(model, m2))
                    vals.append((a @ qf, x))
                num_da = (vals[0][0] - vals[1][0]) / (2 * step)
                num_dx = (vals[0][1] - vals[1][1]) / (2 * step)
                ana = dadr[:, j, c].copy()
                ana[j] += atrace[j, c]
                assert np.abs(num_da[:n] - ana[:n]).max() < 3 * THR2, (name, j, c)
                assert np.abs(num_dx - dxdr[:, j, c]).max() < THR2, (name, j, c)


def test_fd_gradient_and_dqdr(oracle):
    """test_numgrad / test_numdqdr: energy and charges vs their analytic derivatives."""
    step = 1.0e-6
    for name, mol in _fd_cases(oracle):
        model = oracle.new_eeq2019_model(mol)
        n = mol.nat
        r = oracle.eeq_singlepoint(model, mol, want=("qvec", "energy", "gradient", "dqdr"))
        assert r.stat == 0
        for j in range(min(n, 4)):
            for c in range(3):
                e, q = [], []
                for sgn in (+1, -1):
                    xyz = mol.xyz.copy()
                    xyz[j, c] += sgn * step
                    m2 = oracle.Mol(mol.num, xyz, mol.charge, lattice=mol.lattice if mol.periodic.any() else None, id=mol.id)
                    r2 = oracle.eeq_singlepoint(model, m2, want=("qvec", "energy"))
                    e.append(r2.energy.sum())
                    q.append(r2.qvec)
                assert abs((e[0] - e[1]) / (2 * step) - r.gradient[j, c]) < THR2, (name, j, c)
                assert np.abs((q[0] - q[1]) / (2 * step) - r.dqdr[:, j, c]).max() < THR2, (name, j, c)


def test_fd_sigma_and_dqdL(oracle):
    """test_numsigma / test_numdqdL: strain derivatives (xyz, lattice strained together)."""
    step = 1.0e-6
    for name, mol in _fd_cases(oracle):
        model = oracle.new_eeq2019_model(mol)
        r = oracle.eeq_singlepoint(model, mol, want=("qvec", "energy", "gradient", "dqdr"))
        for a in range(3):
            for b in range(3):
                e, q = [], []
                for sgn in (+1, -1):
                    eps = np.eye(3)
                    eps[b, a] += sgn * step  # Fortran eps(a, b): x_a' = sum_b eps(a,b) x_b
                    xyz = mol.xyz @ eps
                    lat = mol.lattice @ eps if mol.periodic.any() else None
                    m2 = oracle.Mol(mol.num, xyz, mol.charge, lattice=lat, id=mol.id)
                    r2 = oracle.eeq_singlepoint(model, m2, want=("qvec", "energy"))
                    e.append(r2.energy.sum())
                    q.append(r2.qvec)
                assert abs((e[0] - e[1]) / (2 * step) - r.sigma[b, a]) < 3 * THR2, (name, a, b)
                assert np.abs((q[0] - q[1]) / (2 * step) - r.dqdL[:, b, a]).max() < 3 * THR2, (name, a, b)


def test_oracle_invariances(oracle):
    """Size-independent properties of the path (the same ones the full-size GPU property tests use): atom
    permutation, rigid translation and rotation of a molecule; rotation of a periodic cell together with its atoms."""
    from synth import cluster, periodic_cell
    rng = np.random.default_rng(5)
    qm, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    num, xyz = cluster(37, 19)
    mol = oracle.Mol(num, xyz, -1.0)
    r = oracle.eeq_singlepoint(oracle.new_eeq2019_model(mol), mol, want=("qvec", "energy", "gradient", "dqdr"))
    perm = rng.permutation(len(num))
    pm = oracle.Mol(np.asarray(num)[perm], xyz[perm], -1.0)
    rp = oracle.eeq_singlepoint(oracle.new_eeq2019_model(pm), pm, want=("qvec", "energy", "gradient", "dqdr"))
    assert np.abs(rp.qvec - r.qvec[perm]).max() < 1e-12 and np.abs(rp.gradient - r.gradient[perm]).max() < 1e-12
    assert np.abs(rp.dqdr - r.dqdr[np.ix_(perm, perm)]).max() < 1e-11
    tm = oracle.Mol(num, xyz @ qm.T + np.array([3.0, -7.0, 11.0]), -1.0)
    rt = oracle.eeq_singlepoint(oracle.new_eeq2019_model(tm), tm, want=("qvec", "energy", "gradient", "dqdr"))
    assert np.abs(rt.qvec - r.qvec).max() < 1e-12 and abs(rt.energy.sum() - r.energy.sum()) < 1e-12
    assert np.abs(rt.gradient - r.gradient @ qm.T).max() < 1e-12
    assert np.abs(rt.sigma - qm @ r.sigma @ qm.T).max() < 1e-11
    assert np.abs(rt.dqdr - r.dqdr @ qm.T).max() < 1e-11
    assert np.abs(r.gradient.sum(axis=0)).max() < 1e-12 and np.abs(r.dqdr.sum(axis=1)).max() < 1e-12  # translation
    num, xyz, lat = periodic_cell(2, 8, shear=0.1)
    pc = oracle.Mol(num, xyz, 0.0, lattice=lat)
    r3 = oracle.eeq_singlepoint(oracle.new_eeq2019_model(pc), pc, want=("qvec", "energy", "gradient"))
    pr = oracle.Mol(num, xyz @ qm.T, 0.0, lattice=lat @ qm.T)
    r3r = oracle.eeq_singlepoint(oracle.new_eeq2019_model(pr), pr, want=("qvec", "energy", "gradient"))
    assert np.abs(r3r.qvec - r3.qvec).max() < 1e-11 and abs(r3r.energy.sum() - r3.energy.sum()) < 1e-11
    assert np.abs(r3r.gradient - r3.gradient @ qm.T).max() < 1e-11
    assert np.abs(r3r.sigma - qm @ r3.sigma @ qm.T).max() < 1e-10
