"""Large-N parity against the CPU oracle (dsytrf / dsytrs / dsytri + dgemm, the reference's own LAPACK route): the
code paths only large systems take -- 512-wide adaptive Cholesky panels, the two-stream look-ahead at depth, the
strided-tile trailing update of the distributed factorisation, deep TRSM recursion -- compared with LAPACK results
instead of conservation laws.  Tolerance 1e-10 relative (BASELINE.json north_star).  Oracle cost: a few seconds each.
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def close(a, b, tol=TOL):
    return np.abs(np.asarray(a) - np.asarray(b)).max() < tol * max(1.0, np.abs(b).max())


def test_cluster_4096_full_derivatives(mc, oracle):
    """configs[2] path at N = 4096: q, E, gradient, sigma, dq/dr, dq/dL against the oracle."""
    from synth import cluster
    nat = 4096
    num, xyz = cluster(nat, 4096)
    mol = mc.Structure(num, xyz, 1.0)
    r = mc.new_eeq2019_model(mol).singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    omol = oracle.Mol(num, xyz, 1.0)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient", "dqdr"))
    for key in ("cn", "qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(r[key], o[key]), (key, np.abs(r[key] - o[key]).max())


@pytest.mark.parametrize("nat", [8192, 8191])
def test_cluster_8192_charges_energy(mc, oracle, nat):
    """N = 8192: the factorisation starts with 512-wide panels (more than 7168 columns left), narrows to 256 and
    128; N = 8191 adds the augmented right-hand-side rows.  Charges and energy against dsytrf / dsytrs."""
    from synth import cluster
    num, xyz = cluster(nat, 8192)
    mol = mc.Structure(num, xyz, -1.0)
    r = mc.new_eeq2019_model(mol).singlepoint(mol, want=("qvec", "energy"))
    omol = oracle.Mol(num, xyz, -1.0)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy"))
    for key in ("cn", "qvec", "energy"):
        assert close(r[key], o[key]), (key, np.abs(r[key] - o[key]).max())


def test_dpotrf_8192_vs_lapack(mc):
    n = 8192
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    del M
    a = np.ascontiguousarray(S.T)
    assert mc.default_context().dpotrf(a) == 0
    Lref = np.linalg.cholesky(S)
    assert np.abs(np.tril(a.T) - Lref).max() < 1e-12 * np.abs(Lref).max()


def test_dpotrf_8192_two_ranks_vs_lapack(mc):
    """The distributed factorisation at its production settings (column groups of 512, sub-panels of 128, strided-tile
    bulk updates) on two contexts of one GPU tied by the local communicator, against LAPACK."""
    n = 8192
    rng = np.random.default_rng(n + 1)
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    del M
    Lref = np.linalg.cholesky(S)
    ctxs = [mc.Context(0), mc.Context(0)]
    mc.Context.comm_init_local(ctxs)
    out, err = [None, None], [None, None]

    def body(r):
        try:
            a = np.ascontiguousarray(S.T)
            info = ctxs[r].dpotrf(a)
            out[r] = (info, float(np.abs(np.tril(a.T) - Lref).max()))
        except BaseException as exc:  # noqa: BLE001
            err[r] = exc

    th = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    for c in ctxs:
        c.comm_destroy()
        c.close()
    assert err == [None, None], err
    for info, e in out:
        assert info == 0 and e < 1e-12 * np.abs(Lref).max()


def test_periodic_1728_vs_oracle(mc, oracle):
    """configs[3] path at 12^3 = 1728 atoms: Ewald real-space tiles, Phi W Phi^T, rank-one shift; q, E, gradient,
    sigma against the oracle (periodic branch of the oracle: parity unpinned against reference numbers)."""
    from synth import periodic_cell
    num, xyz, lat = periodic_cell(12, 1728)
    mol = mc.Structure(num, xyz, 0.0, lattice=lat)
    r = mc.new_eeq2019_model(mol).singlepoint(mol, want=("qvec", "energy", "gradient"))
    omol = oracle.Mol(num, xyz, 0.0, lattice=lat)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient"))
    for key in ("cn", "qvec", "energy", "gradient", "sigma"):
        assert close(r[key], o[key]), (key, np.abs(r[key] - o[key]).max())


def test_eeqbc_2048_vs_oracle(mc, oracle):
    """configs[4] path at N = 2048: EEQ-BC q, E, gradient, sigma, dq/dr, dq/dL against the oracle (EEQ-BC branch of
    the oracle: parity unpinned against reference numbers; synthetic vdW / EN tables on both sides)."""
    from synth import cluster
    from test_gpu_eeqbc import models
    num, xyz = cluster(2048, 2048)
    mol, model, omol, omodel = models(mc, oracle, num, xyz, 0.0)
    r = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    o = oracle.eeqbc_singlepoint(omodel, omol, want=("qvec", "energy", "gradient", "dqdr"))
    for key in ("qvec", "energy", "gradient", "sigma", "dqdr", "dqdL"):
        assert close(r[key], o[key]), (key, np.abs(r[key] - o[key]).max())


def test_eeqbc_2048_two_ranks_vs_oracle(mc, oracle, monkeypatch):
    """The same N = 2048 EEQ-BC system shared by two ranks (two contexts of one GPU, in-process communicator): distributed
    factorisation, row-sharded pair passes, shared gradient pass, dq/dr rows of each rank -- against the oracle."""
    from synth import cluster
    from multicharge_b200.shard import row_bounds
    monkeypatch.setenv("MCHRG_B200_DIST_MIN", "1024")
    num, xyz = cluster(2048, 2048)
    omol = oracle.Mol(num, xyz, 0.0)
    om = oracle.new_eeqbc2025_model(omol)
    mol = mc.Structure(num, xyz, 0.0)
    ctxs = [mc.Context(0), mc.Context(0)]
    mc.Context.comm_init_local(ctxs)
    bounds = row_bounds(2048, 2)
    out, err = [None, None], [None, None]

    def body(r):
        try:
            model = mc.new_eeqbc_model(mol, chi=om.chi, rad=om.rad, eta=om.eta, kcnchi=om.kcnchi, kqchi=om.kqchi,
                                       kqeta=om.kqeta, kcnrad=om.kcnrad, cap=om.cap, avg_cn=om.avg_cn, rvdw=om.rvdw,
                                       rcov=om.rcov, en=om.en, kbc=om.kbc, cutoff=om.cutoff, cn_exp=om.kcn,
                                       norm_exp=om.norm_exp, ctx=ctxs[r])
            a = mc.singlepoint_rows(model, mol, 0, 2048, dqdr=False)
            b = mc.singlepoint_rows(model, mol, int(bounds[r]), int(bounds[r + 1]))
            out[r] = (a, b)
        except BaseException as exc:  # noqa: BLE001
            err[r] = exc

    th = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    for c in ctxs:
        c.comm_destroy()
        c.close()
    assert err == [None, None], err
    o = oracle.eeqbc_singlepoint(om, omol, want=("qvec", "energy", "gradient", "dqdr"))
    dqdr = np.zeros_like(o["dqdr"])
    for r, (a, b) in enumerate(out):
        for key in ("qvec", "energy", "gradient", "sigma"):
            assert close(a[key], o[key]), (r, key, np.abs(a[key] - o[key]).max())
        jb, je = int(bounds[r]), int(bounds[r + 1])
        dqdr[:, jb:je, :] = b["dqdr"]
        assert close(b["gradient"], o["gradient"][jb:je]) and close(b["dqdL"], o["dqdL"])
    assert close(dqdr, o["dqdr"])
