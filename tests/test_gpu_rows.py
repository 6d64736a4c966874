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
