"""Parity of the one-CTA-per-molecule batch kernel against the CPU oracle (per molecule) and the
reference's golden vectors.  Tolerance 1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-10


def oracle_batch(oracle, offsets, num, xyz, charge):
    q, e, g = np.zeros(len(num)), np.zeros(len(num)), np.zeros((len(num), 3))
    for m in range(len(charge)):
        s = slice(offsets[m], offsets[m + 1])
        mol = oracle.Mol(num[s], xyz[s], charge[m])
        r = oracle.eeq_singlepoint(oracle.new_eeq2019_model(mol), mol)
        assert r.stat == 0
        q[s], e[s], g[s] = r.qvec, r.energy, r.gradient
    return q, e, g


@pytest.mark.parametrize("mode", ["fused", "split"])
def test_batch_matches_oracle_all_size_classes(mc, oracle, monkeypatch, mode):
    """Every size class through both schedules (the split path pads to the augmented order ceil((n + 2) / 16) * 16:
    sizes around n mod 16 in {13, 14, 15, 0, 1} are the edges)."""
    from synth import cluster
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", mode)
    sizes = [1, 2, 3, 13, 14, 15, 16, 17, 29, 30, 31, 32, 33, 46, 47, 48, 49, 64, 65, 80, 96, 100, 112, 127, 128, 129, 144, 160,
             161, 176, 192, 200, 205, 206, 207, 208]
    nums, xyzs = [], []
    for k, n in enumerate(sizes):
        num, xyz = cluster(n, 100 + k)
        nums.append(num)
        xyzs.append(xyz)
    offsets = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    num, xyz = np.concatenate(nums), np.concatenate(xyzs)
    charge = np.array([(k % 3) - 1.0 for k in range(len(sizes))])
    model = mc.new_eeq2019_batch_model()
    out = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    assert np.all(out["status"] == 0)
    q, e, g = oracle_batch(oracle, offsets, num, xyz, charge)
    assert np.abs(out["qvec"] - q).max() < TOL * max(1.0, np.abs(q).max())
    assert np.abs(out["energy"] - e).max() < TOL * max(1.0, np.abs(e).max())
    assert np.abs(out["gradient"] - g).max() < TOL * max(1.0, np.abs(g).max())
    for m in range(len(sizes)):
        assert abs(out["qvec"][offsets[m]:offsets[m + 1]].sum() - charge[m]) < 1e-10
    # charges + energy only
    out2 = mc.singlepoint_batch(model, offsets, num, xyz, charge, gradient=False)
    assert np.abs(out2["qvec"] - q).max() < TOL and np.abs(out2["energy"] - e).max() < TOL


def test_batch_golden_vectors(mc):
    gs = [load_golden(n) for n in GOLDEN_CASES]
    sizes = [len(g["num"]) for g in gs]
    offsets = np.zeros(len(gs) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    num = np.concatenate([np.array(g["num"], dtype=np.int32) for g in gs])
    xyz = np.concatenate([np.array(g["xyz"]) for g in gs])
    charge = np.array([g["charge"] for g in gs])
    out = mc.singlepoint_batch(mc.new_eeq2019_batch_model(), offsets, num, xyz, charge)
    for m, g in enumerate(gs):
        s = slice(offsets[m], offsets[m + 1])
        assert np.abs(out["qvec"][s] - np.array(g["charges"])).max() < 5e-11
        if "energy" in g:
            assert abs(out["energy"][s].sum() - g["energy"]) < 5e-11
            assert np.abs(out["gradient"][s].ravel() - np.array(g["gradient"])).max() < 5e-11


def test_batch_random_1000_and_bad_molecule(mc, oracle):
    from synth import batch
    offsets, num, xyz, charge = batch(300, 77, nmin=30, nmax=200)
    # molecule 5: two hydrogens on top of each other's charge cloud -> not positive definite
    s = offsets[5]
    num[s] = num[s + 1] = 1
    xyz[s + 1] = xyz[s] + np.array([0.4, 0.0, 0.0])
    out = mc.singlepoint_batch(mc.new_eeq2019_batch_model(), offsets, num, xyz, charge)
    assert out["status"][5] > 0 and np.isnan(out["qvec"][offsets[5]])
    ok = np.ones(len(charge), dtype=bool)
    ok[5] = False
    assert np.all(out["status"][ok] == 0)
    pick = [0, 1, 2, 50, 123, 299]
    for m in pick:
        sl = slice(offsets[m], offsets[m + 1])
        mol = oracle.Mol(num[sl], xyz[sl], charge[m])
        r = oracle.eeq_singlepoint(oracle.new_eeq2019_model(mol), mol)
        assert np.abs(out["qvec"][sl] - r.qvec).max() < TOL
        assert np.abs(out["energy"][sl] - r.energy).max() < TOL
        assert np.abs(out["gradient"][sl] - r.gradient).max() < TOL


def test_batch_device_resident(mc, oracle):
    import torch
    from synth import batch
    offsets, num, xyz, charge = batch(64, 5, nmin=10, nmax=120)
    model = mc.new_eeq2019_batch_model()
    ref = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    dev = torch.device("cuda:0")
    ntot = int(offsets[-1])
    out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
               gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(64, dtype=torch.int32, device=dev))
    mc.singlepoint_batch(model, torch.tensor(offsets, device=dev), torch.tensor(num, device=dev), torch.tensor(xyz, device=dev),
                         torch.tensor(charge, device=dev), out=out)
    torch.cuda.synchronize()
    assert np.array_equal(out["qvec"].cpu().numpy(), ref["qvec"])
    assert np.array_equal(out["gradient"].cpu().numpy(), ref["gradient"])


def test_batch_with_molecules_above_kernel_limit(mc, oracle):
    """Molecules above the 208 atoms one CTA holds go through the single-system path inside the same call
    (host and device buffers); their neighbours are unaffected.  Also: a batch made of large molecules only."""
    import torch
    from synth import cluster
    sizes = [40, 209, 17, 300, 208, 256, 5]
    nums, xyzs = zip(*[cluster(n, 700 + k) for k, n in enumerate(sizes)])
    offsets = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    num, xyz = np.concatenate(nums), np.concatenate(xyzs)
    charge = np.array([(k % 3) - 1.0 for k in range(len(sizes))])
    model = mc.new_eeq2019_batch_model()
    q, e, g = oracle_batch(oracle, offsets, num, xyz, charge)
    out = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    assert np.all(out["status"] == 0)
    for key, ref in (("qvec", q), ("energy", e), ("gradient", g)):
        assert np.abs(out[key] - ref).max() < TOL * max(1.0, np.abs(ref).max()), key
    out2 = mc.singlepoint_batch(model, offsets, num, xyz, charge, gradient=False)
    assert np.abs(out2["qvec"] - q).max() < TOL and np.abs(out2["energy"] - e).max() < TOL
    dev = torch.device("cuda:0")
    ntot = int(offsets[-1])
    dout = dict(qvec=torch.full((ntot,), 7.0, dtype=torch.float64, device=dev), energy=torch.full((ntot,), 7.0, dtype=torch.float64, device=dev),
                gradient=torch.full((ntot, 3), 7.0, dtype=torch.float64, device=dev), status=torch.full((len(sizes),), 7, dtype=torch.int32, device=dev))
    mc.singlepoint_batch(model, torch.tensor(offsets, device=dev), torch.tensor(num, device=dev), torch.tensor(xyz, device=dev),
                         torch.tensor(charge, device=dev), out=dout)
    torch.cuda.synchronize()
    assert np.all(dout["status"].cpu().numpy() == 0)
    for key in ("qvec", "energy", "gradient"):  # (the single-system gradient sums in a run-dependent order)
        assert np.abs(dout[key].cpu().numpy() - out[key]).max() < 1e-12, key
    # large molecules only
    sl = slice(offsets[3], offsets[4])
    only = mc.singlepoint_batch(model, np.array([0, 300], dtype=np.int64), num[sl], xyz[sl], charge[3:4])
    assert only["status"][0] == 0
    assert np.abs(only["qvec"] - out["qvec"][sl]).max() < 1e-12 and np.abs(only["gradient"] - out["gradient"][sl]).max() < 1e-12


def test_batch_pipelined_host_path_equals_device_path(mc):
    """Large host batches are pipelined in chunks (H2D / kernels / D2H overlap): same bits as one device-resident pass."""
    import torch
    from synth import batch
    offsets, num, xyz, charge = batch(3000, 9, nmin=30, nmax=200)
    assert offsets[-1] > (1 << 18)
    model = mc.new_eeq2019_batch_model()
    host = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    dev = torch.device("cuda:0")
    ntot = int(offsets[-1])
    out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
               gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(3000, dtype=torch.int32, device=dev))
    mc.singlepoint_batch(model, torch.tensor(offsets, device=dev), torch.tensor(num, device=dev), torch.tensor(xyz, device=dev),
                         torch.tensor(charge, device=dev), out=out)
    torch.cuda.synchronize()
    assert np.all(host["status"] == 0)
    for key in ("qvec", "energy", "gradient"):
        assert np.array_equal(out[key].cpu().numpy(), host[key]), key


def test_batch_split_path_equals_fused(mc, oracle, monkeypatch):
    """The split path (pair phase A / factorisation / pair phase C as separate kernels on several streams)
    produces the same bits as the fused one-kernel path, on host and on device buffers."""
    import torch
    from synth import batch
    offsets, num, xyz, charge = batch(700, 31, nmin=1, nmax=208)
    model = mc.new_eeq2019_batch_model()
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "fused")
    fused = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "split")
    split = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    assert np.array_equal(fused["status"], split["status"])
    for key in ("qvec", "energy", "gradient"):  # different factorisation order: agreement to rounding, not to the bit
        assert np.abs(fused[key] - split[key]).max() < 1e-12, key
    dev = torch.device("cuda:0")
    ntot = int(offsets[-1])
    out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
               gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(700, dtype=torch.int32, device=dev))
    mc.singlepoint_batch(model, torch.tensor(offsets, device=dev), torch.tensor(num, device=dev), torch.tensor(xyz, device=dev),
                         torch.tensor(charge, device=dev), out=out)
    torch.cuda.synchronize()
    assert np.abs(out["gradient"].cpu().numpy() - fused["gradient"]).max() < 1e-12
    # charges only (no pair phase C) and a failing molecule through the split path
    q_only = mc.singlepoint_batch(model, offsets, num, xyz, charge, gradient=False)
    assert np.abs(q_only["qvec"] - fused["qvec"]).max() < 1e-12 and np.abs(q_only["energy"] - fused["energy"]).max() < 1e-12
    s = offsets[3]
    num2, xyz2 = num.copy(), xyz.copy()
    num2[s] = num2[s + 1] = 1
    xyz2[s + 1] = xyz2[s] + np.array([0.4, 0.0, 0.0])
    bad = mc.singlepoint_batch(model, offsets, num2, xyz2, charge)
    assert bad["status"][3] > 0 and np.isnan(bad["qvec"][offsets[3]]) and np.isnan(bad["gradient"][offsets[3]]).all()
    ok = np.ones(700, dtype=bool)
    ok[3] = False
    assert np.all(bad["status"][ok] == 0)
    m = 10
    sl = slice(offsets[m], offsets[m + 1])
    mol = oracle.Mol(num[sl], xyz[sl], charge[m])
    r = oracle.eeq_singlepoint(oracle.new_eeq2019_model(mol), mol)
    assert np.abs(split["gradient"][sl] - r.gradient).max() < 1e-10


def test_batch_plan_is_reused_and_rebuilt(mc, monkeypatch):
    """The split path keeps what it derives from the offsets (size classes, permutation, scratch layout) on the context:
    the same offsets with new coordinates reuse it, other offsets -- same number of molecules and atoms included --
    rebuild it."""
    from synth import batch
    model = mc.new_eeq2019_batch_model()
    offs_a, num_a, xyz_a, chg_a = batch(300, 5, nmin=3, nmax=120)
    sizes = np.diff(offs_a)[::-1].copy()  # the same sizes in reverse order: same nmol, same ntot, other offsets
    offs_b = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rng = np.random.default_rng(9)
    num_b = num_a[rng.permutation(len(num_a))].copy()
    xyz_b = np.concatenate([xyz_a[offs_a[m]:offs_a[m + 1]] for m in range(299, -1, -1)]) + 0.01
    chg_b = chg_a[::-1].copy()
    xyz_a2 = xyz_a * 1.01
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "fused")
    ref = [mc.singlepoint_batch(model, o, n, x, c) for o, n, x, c in
           ((offs_a, num_a, xyz_a, chg_a), (offs_a, num_a, xyz_a2, chg_a), (offs_b, num_b, xyz_b, chg_b), (offs_a, num_a, xyz_a, chg_a))]
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "split")
    got = [mc.singlepoint_batch(model, o, n, x, c) for o, n, x, c in
           ((offs_a, num_a, xyz_a, chg_a), (offs_a, num_a, xyz_a2, chg_a), (offs_b, num_b, xyz_b, chg_b), (offs_a, num_a, xyz_a, chg_a))]
    for r, g in zip(ref, got):
        assert np.array_equal(r["status"], g["status"])
        ok = np.repeat(r["status"] == 0, np.diff(offs_a) if g is not got[2] else np.diff(offs_b))
        for key in ("qvec", "energy", "gradient"):
            assert np.abs(r[key][ok] - g[key][ok]).max() < 1e-11, key


@pytest.mark.parametrize("group", ["13", "4", "1"])
def test_batch_factor_launch_groups(mc, monkeypatch, group):
    """MCHRG_B200_FACTOR_GROUP: size classes per factorisation launch.  The shared memory of a launch is sized for its
    largest class, and the backward substitution stages its operands in whatever the molecule leaves free of it (or
    falls back to direct loads when two stages do not fit): every grouping gives the same results."""
    from synth import batch
    model = mc.new_eeq2019_batch_model()
    offsets, num, xyz, charge = batch(500, 78, nmin=1, nmax=208)
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "fused")
    ref = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "split")
    monkeypatch.setenv("MCHRG_B200_FACTOR_GROUP", group)
    got = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    assert np.array_equal(ref["status"], got["status"]) and int((ref["status"] != 0).sum()) == 0
    for key in ("qvec", "energy", "gradient"):
        assert np.abs(ref[key] - got[key]).max() < 1e-11, key


def test_batch_tma_fed_factorisation_variant(mc, monkeypatch):
    """MCHRG_B200_FACTOR=tma: the factorisation kernel whose operands are streamed into shared-memory rings with
    cp.async.bulk + mbarriers (measured slower than the default, kept as the documented alternative) gives the same
    results as the default kernel, every size class included."""
    from synth import batch
    model = mc.new_eeq2019_batch_model()
    offsets, num, xyz, charge = batch(400, 77, nmin=1, nmax=208)
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", "split")
    ref = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    monkeypatch.setenv("MCHRG_B200_FACTOR", "tma")
    got = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    assert np.array_equal(ref["status"], got["status"]) and int((ref["status"] != 0).sum()) == 0
    for key in ("qvec", "energy", "gradient"):
        assert np.abs(ref[key] - got[key]).max() < 1e-11, key


def test_batch_full_size_properties(mc):
    """BASELINE.json configs[1] at its full size (100 000 molecules of 30-200 atoms, the bench workload), checked
    through size-independent properties: charge conservation per molecule, vanishing net force (translational
    invariance), invariance of q and E under a rigid rotation + translation of every molecule with the gradient
    rotating along, and independence of the batch order."""
    from synth import batch
    nmol = 100_000
    offsets, num, xyz, charge = batch(nmol, 20250102, 30, 200)
    model = mc.new_eeq2019_batch_model()
    r = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    assert int((r["status"] != 0).sum()) == 0
    seg = np.repeat(np.arange(nmol), np.diff(offsets))
    qsum = np.bincount(seg, weights=r["qvec"], minlength=nmol)
    assert np.abs(qsum - charge).max() < 1e-11
    fnet = np.stack([np.bincount(seg, weights=r["gradient"][:, c], minlength=nmol) for c in range(3)], axis=1)
    assert np.abs(fnet).max() < 1e-10
    # rigid motion (one rotation for all, a different translation per molecule) and reversed molecule order
    rng = np.random.default_rng(5)
    rot, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    shift = rng.uniform(-20.0, 20.0, (nmol, 3))
    xyz2 = xyz @ rot.T + shift[seg]
    order = np.arange(nmol)[::-1]
    sizes = np.diff(offsets)
    off2 = np.concatenate([[0], np.cumsum(sizes[order])])
    idx = np.concatenate([np.arange(offsets[m], offsets[m + 1]) for m in order])
    r2 = mc.singlepoint_batch(model, off2, num[idx], xyz2[idx], charge[order])
    assert int((r2["status"] != 0).sum()) == 0
    assert np.abs(r2["qvec"] - r["qvec"][idx]).max() < 1e-11
    assert np.abs(r2["energy"] - r["energy"][idx]).max() < 1e-11
    assert np.abs(r2["gradient"] - (r["gradient"] @ rot.T)[idx]).max() < 1e-10


@pytest.mark.parametrize("mode", ["fused", "split"])
def test_batch_rejects_atomic_numbers_outside_the_model(mc, monkeypatch, mode):
    """An atomic number the model has no species for is reported per molecule (MCHRG_B200_ERR_ARG = -1, NaN outputs)
    instead of reading a neighbouring parameter slot; the other molecules of the batch are unaffected."""
    from synth import batch
    monkeypatch.setenv("MCHRG_B200_BATCH_MODE", mode)
    offsets, num, xyz, charge = batch(40, 3, 10, 60)
    model = mc.new_eeq2019_batch_model()
    good = mc.singlepoint_batch(model, offsets, num, xyz, charge)
    bad_num = num.copy()
    bad_num[offsets[7] + 2] = 0
    bad_num[offsets[21]] = 104
    out = mc.singlepoint_batch(model, offsets, bad_num, xyz, charge)
    st = out["status"]
    assert st[7] == -1 and st[21] == -1 and int((st != 0).sum()) == 2
    assert np.all(np.isnan(out["qvec"][offsets[7]:offsets[8]]))
    ok = np.ones(len(num), dtype=bool)
    ok[offsets[7]:offsets[8]] = False
    ok[offsets[21]:offsets[22]] = False
    for key in ("qvec", "energy"):
        assert np.array_equal(out[key][ok], good[key][ok])


def test_get_charges_fast_path_agrees_with_general_path_on_failed_pivot(mc, monkeypatch):
    """Two H atoms 0.5 Bohr apart: the Coulomb block is indefinite on the charge-conserving subspace.  The one-launch
    get_charges route must not report anything the general route does not (it falls back to it)."""
    mol = mc.Structure(np.array([1, 1]), np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.5]]), 0.0)
    model = mc.new_eeq2019_model(mol)
    res = []
    for fused in ("1", "0"):
        monkeypatch.setenv("MCHRG_B200_SMALL_FUSED", fused)
        try:
            res.append(("ok", mc.get_charges(model, mol)))
        except mc.FactorizationError as exc:
            res.append(("fail", exc.info))
    assert res[0][0] == res[1][0]
    if res[0][0] == "ok":
        assert np.abs(res[0][1] - res[1][1]).max() < 1e-12
