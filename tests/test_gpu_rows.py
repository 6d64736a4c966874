"""Row-sharded single-system path (multi-GPU building block): the derivative outputs of disjoint atom blocks,
computed by independent calls (one per rank in a multi-GPU run), concatenate to the full single-call result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nat,blocks", [(300, 2), (517, 3), (130, 4)])
def test_row_blocks_concatenate_to_full(mc, oracle, nat, blocks):
    from synth import cluster
    num, xyz = cluster(nat, 40 + nat)
    mol = mc.Structure(num, xyz, 1.0)
    model = mc.new_eeq2019_model(mol)
    full = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    bounds = np.linspace(0, nat, blocks + 1).astype(int)
    grad = np.zeros((nat, 3))
    dqdr = np.zeros((nat, nat, 3))
    for b in range(blocks):
        jb, je = int(bounds[b]), int(bounds[b + 1])
        r = mc.singlepoint_rows(model, mol, jb, je)
        grad[jb:je] = r["gradient"]
        dqdr[:, jb:je, :] = r["dqdr"]
        for key in ("cn", "qvec", "energy", "sigma", "dqdL"):
            assert np.abs(r[key] - full[key]).max() < 1e-12 * max(1.0, np.abs(full[key]).max()), key
    assert np.abs(grad - full["gradient"]).max() < 1e-12 * max(1.0, np.abs(full["gradient"]).max())
    assert np.abs(dqdr - full["dqdr"]).max() < 1e-12 * max(1.0, np.abs(full["dqdr"]).max())
    # and against the oracle
    omol = oracle.Mol(num, xyz, 1.0)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient", "dqdr"))
    assert np.abs(dqdr - o.dqdr).max() < 1e-10 * max(1.0, np.abs(o.dqdr).max())
    assert np.abs(grad - o.gradient).max() < 1e-10 * max(1.0, np.abs(o.gradient).max())


def test_row_block_rejects_bad_ranges(mc):
    from synth import cluster
    num, xyz = cluster(20, 1)
    mol = mc.Structure(num, xyz, 0.0)
    model = mc.new_eeq2019_model(mol)
    for jb, je in ((-1, 5), (5, 5), (3, 21)):
        with pytest.raises(mc.MchrgError):
            mc.singlepoint_rows(model, mol, jb, je)


def test_cluster_full_size_properties(mc):
    """BASELINE.json configs[2] at its full size (N = 16384 cluster) through size-independent properties:
    charge conservation, vanishing net force and virial symmetry of q + E + gradient, and -- on one 512-row block
    of dq/dr (the full tensor is 6.4 GB) -- d(sum_i q_i)/dr_j = 0 plus agreement of the block's gradient rows with
    the full gradient."""
    import torch
    from synth import cluster
    nat = 16384
    num, xyz = cluster(nat, 20250103)
    mol = mc.Structure(num, xyz, -1.0)
    ctx = mc.default_context(0)
    model = mc.new_eeq2019_model(mol, ctx)
    dev = torch.device("cuda", 0)
    f64 = dict(dtype=torch.float64, device=dev)
    txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
    full = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                              gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64))
    mc.singlepoint_rows(model, mol, 0, nat, out=full, xyz=txyz, id=tid)
    torch.cuda.synchronize()
    q, g, sig = full["qvec"].cpu().numpy(), full["gradient"].cpu().numpy(), full["sigma"].cpu().numpy()
    assert abs(q.sum() + 1.0) < 1e-10
    assert np.abs(g.sum(axis=0)).max() < 1e-9 * max(1.0, np.abs(g).max())
    assert np.abs(sig - sig.T).max() < 1e-9 * max(1.0, np.abs(sig).max())
    jb, je = 4096, 4608
    blk = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                             gradient=torch.empty(je - jb, 3, **f64), sigma=torch.zeros(3, 3, **f64),
                             dqdr=torch.empty(nat, je - jb, 3, **f64), dqdL=torch.empty(nat, 3, 3, **f64))
    mc.singlepoint_rows(model, mol, jb, je, out=blk, xyz=txyz, id=tid)
    torch.cuda.synchronize()
    dq = blk["dqdr"]
    assert float(dq.sum(dim=0).abs().max()) < 1e-10 * max(1.0, float(dq.abs().max()))
    assert float(blk["dqdL"].sum(dim=0).abs().max()) < 1e-9 * max(1.0, float(blk["dqdL"].abs().max()))
    assert np.abs(blk["gradient"].cpu().numpy() - g[jb:je]).max() < 1e-11 * max(1.0, np.abs(g).max())
    assert np.abs(blk["qvec"].cpu().numpy() - q).max() < 1e-12


def test_eeqbc_row_blocks_concatenate_to_full(mc, oracle):
    """EEQ-BC through singlepoint_rows on one context: disjoint atom blocks reproduce the full dq/dr and gradient."""
    from synth import cluster
    from test_gpu_eeqbc import models
    nat = 260
    num, xyz = cluster(nat, 12)
    mol, model, omol, omodel = models(mc, oracle, num, xyz, 0.0)
    full = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    o = oracle.eeqbc_singlepoint(omodel, omol, want=("qvec", "energy", "gradient", "dqdr"))
    grad = np.zeros((nat, 3))
    dqdr = np.zeros((nat, nat, 3))
    for jb, je in ((0, 100), (100, 131), (131, nat)):
        r = mc.singlepoint_rows(model, mol, jb, je)
        grad[jb:je] = r["gradient"]
        dqdr[:, jb:je, :] = r["dqdr"]
        for key in ("qvec", "energy", "sigma", "dqdL"):
            assert np.abs(r[key] - full[key]).max() < 1e-12 * max(1.0, np.abs(full[key]).max()), key
    for got, key in ((grad, "gradient"), (dqdr, "dqdr")):
        assert np.abs(got - full[key]).max() < 1e-12 * max(1.0, np.abs(full[key]).max())
        assert np.abs(got - o[key]).max() < 1e-10 * max(1.0, np.abs(o[key]).max())


@pytest.mark.parametrize("ncell,blocks,shear", [(6, 3, 0.0), (7, 2, 0.1)])
def test_periodic_row_blocks_concatenate_to_full(mc, oracle, ncell, blocks, shear):
    """Periodic EEQ2019 (configs[3] path): dq/dr rows of disjoint atom blocks -- the pair tiles of a block in a second
    pass, Psi rows of the block for the reciprocal-space part -- concatenate to the one-call result and match the oracle."""
    from synth import periodic_cell
    num, xyz, lat = periodic_cell(ncell, 77 + ncell, shear=shear)
    nat = len(num)
    mol = mc.Structure(num, xyz, 0.0, lattice=lat)
    model = mc.new_eeq2019_model(mol)
    full = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    bounds = np.linspace(0, nat, blocks + 1).astype(int)
    bounds[1] = bounds[1] // 2 + 7  # unequal blocks, boundaries off the 32-atom tile grid
    grad = np.zeros((nat, 3))
    dqdr = np.zeros((nat, nat, 3))
    for b in range(blocks):
        jb, je = int(bounds[b]), int(bounds[b + 1])
        r = mc.singlepoint_rows(model, mol, jb, je)
        grad[jb:je] = r["gradient"]
        dqdr[:, jb:je, :] = r["dqdr"]
        for key in ("cn", "qvec", "energy", "sigma", "dqdL"):
            assert np.abs(r[key] - full[key]).max() < 1e-11 * max(1.0, np.abs(full[key]).max()), key
    assert np.abs(grad - full["gradient"]).max() < 1e-11 * max(1.0, np.abs(full["gradient"]).max())
    assert np.abs(dqdr - full["dqdr"]).max() < 1e-11 * max(1.0, np.abs(full["dqdr"]).max())
    omol = oracle.Mol(num, xyz, 0.0, lattice=lat)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient", "dqdr"))
    assert np.abs(dqdr - o.dqdr).max() < 1e-10 * max(1.0, np.abs(o.dqdr).max())
    assert np.abs(grad - o.gradient).max() < 1e-10 * max(1.0, np.abs(o.gradient).max())
