"""The multi-GPU single-system code on ONE GPU: R contexts driven by R host threads and tied together by the
library's local communicator (mchrg_b200_comm_init_local -- the same schedule, kernels and stream choreography as
the NCCL backend, with stream-ordered copies standing in for the NVLink transfers).

  * collectives (broadcast / all-reduce / all-gather) against numpy;
  * block-cyclic Cholesky (dpotrf_dist) against numpy/LAPACK, every rank starting with ONLY its own column groups;
  * the whole single point (q, E, gradient, sigma; dq/dr rows) with the row-sharded O(N^2) passes against the
    one-context result and the CPU oracle (tolerance 1e-10, BASELINE.json);
  * two contexts used from two threads without a communicator (per-context kernel attributes, re-entrancy).
"""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fort(a):
    return np.ascontiguousarray(a.T)


def run_ranks(world, body):
    """body(rank) on `world` threads; re-raises the first exception."""
    err = [None] * world
    out = [None] * world

    def run(r):
        try:
            out[r] = body(r)
        except BaseException as exc:  # noqa: BLE001
            err[r] = exc

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.fixture()
def dist_env(monkeypatch):
    """Lower the size threshold of the distributed route so that small systems take it (knobs are read per call)."""
    monkeypatch.setenv("MCHRG_B200_DIST_MIN", "512")
    yield


@pytest.fixture(params=[2, 3])
def group(request, mc):
    ctxs = [mc.Context(0) for _ in range(request.param)]
    mc.Context.comm_init_local(ctxs)
    yield ctxs
    for c in ctxs:
        c.comm_destroy()
        c.close()


def test_local_collectives(mc, group):
    import torch
    world = len(group)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    base = rng.standard_normal((world, 1000))

    def body(r):
        ctx = group[r]
        assert ctx.comm_rank == r and ctx.comm_size == world
        t = torch.tensor(base[r], device=dev)
        ctx.comm_bcast(t, world - 1)
        b = t.cpu().numpy().copy()
        t = torch.tensor(base[r], device=dev)
        ctx.comm_allreduce_sum(t)
        s = t.cpu().numpy().copy()
        send = torch.tensor(base[r], device=dev)
        recv = torch.empty(world, 1000, dtype=torch.float64, device=dev)
        ctx.comm_allgather(send, recv)
        return b, s, recv.cpu().numpy()

    for b, s, g in run_ranks(world, body):
        assert np.array_equal(b, base[world - 1])
        assert np.abs(s - base.sum(axis=0)).max() < 1e-14
        assert np.array_equal(g, base)


@pytest.mark.parametrize("fuse", ["0", "1"])
@pytest.mark.parametrize("n", [1024, 1500, 2176])
def test_block_cyclic_cholesky_vs_lapack(mc, group, dist_env, n, fuse, monkeypatch):
    """fuse = 1: the cluster step kernel (leaf + head row solve + next diagonal tile in one launch, the arriving
    sub-panel's last update as its phase 0) instead of the three launches."""
    world = len(group)
    monkeypatch.setenv("MCHRG_B200_DIST_GROUP", "256")
    monkeypatch.setenv("MCHRG_B200_DIST_FUSE", fuse)
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    Lref = np.linalg.cholesky(S)
    from multicharge_b200.shard import group_width
    npad = -(-n // 128) * 128
    gw = group_width(npad, world, 256)

    def body(r):
        a = S.copy()
        for g in range(-(-npad // gw)):  # a rank may only read its own column groups
            if g % world != r:
                a[:, g * gw:(g + 1) * gw] = np.nan
        buf = fort(a)
        info = group[r].dpotrf(buf)
        return info, np.tril(buf.T)

    for info, L in run_ranks(world, body):
        assert info == 0
        assert np.abs(L - Lref).max() < 1e-12 * np.abs(Lref).max()


def test_block_cyclic_cholesky_reports_pivot_on_every_rank(mc, group, dist_env):
    n = 1024
    S = np.eye(n) * 2.0
    S[700, 700] = -1.0

    def body(r):
        return group[r].dpotrf(fort(S.copy()))

    assert run_ranks(len(group), body) == [701] * len(group)


@pytest.mark.parametrize("nat", [1100, 1536])
def test_sharded_singlepoint_vs_oracle(mc, oracle, group, dist_env, nat):
    """q + E + gradient + sigma with the distributed factorisation and the row-sharded pair / CN passes; then dq/dr
    by row blocks (one block per rank), against the one-context result and the oracle."""
    from synth import cluster
    from multicharge_b200.shard import row_bounds
    world = len(group)
    num, xyz = cluster(nat, 70 + nat)
    mol = mc.Structure(num, xyz, -1.0)
    single = mc.new_eeq2019_model(mol).singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
    bounds = row_bounds(nat, world)

    def body(r):
        model = mc.new_eeq2019_model(mol, group[r])
        a = mc.singlepoint_rows(model, mol, 0, nat, dqdr=False)
        b = mc.singlepoint_rows(model, mol, int(bounds[r]), int(bounds[r + 1]))
        return a, b

    res = run_ranks(world, body)
    omol = oracle.Mol(num, xyz, -1.0)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient", "dqdr"))
    grad = np.zeros((nat, 3))
    dqdr = np.zeros((nat, nat, 3))
    for r, (a, b) in enumerate(res):
        for key in ("cn", "qvec", "energy", "gradient", "sigma"):
            scale = max(1.0, np.abs(o[key]).max())
            assert np.abs(a[key] - o[key]).max() < 1e-10 * scale, (r, key)
            assert np.abs(a[key] - single[key]).max() < 1e-11 * scale, (r, key)
        jb, je = int(bounds[r]), int(bounds[r + 1])
        grad[jb:je] = b["gradient"]
        dqdr[:, jb:je, :] = b["dqdr"]
        assert np.abs(b["dqdL"] - o["dqdL"]).max() < 1e-10 * max(1.0, np.abs(o["dqdL"]).max())
    assert np.abs(grad - o["gradient"]).max() < 1e-10 * max(1.0, np.abs(o["gradient"]).max())
    assert np.abs(dqdr - o["dqdr"]).max() < 1e-10 * max(1.0, np.abs(o["dqdr"]).max())


def test_sharded_eeqbc_singlepoint_vs_oracle(mc, oracle, group, dist_env):
    """EEQ-BC (configs[4] path) shared by several ranks: distributed factorisation, row-sharded pair passes joined by one
    all-reduce, the gradient's pair pass shared as well (every rank ends with the whole gradient), dq/dr by row blocks;
    against the one-context result and the oracle."""
    from synth import cluster
    from multicharge_b200.shard import row_bounds
    world = len(group)
    nat = 700
    num, xyz = cluster(nat, 790)
    mol = mc.Structure(num, xyz, 1.0)
    omol = oracle.Mol(num, xyz, 1.0)
    om = oracle.new_eeqbc2025_model(omol)

    def make(ctx):
        return mc.new_eeqbc_model(mol, chi=om.chi, rad=om.rad, eta=om.eta, kcnchi=om.kcnchi, kqchi=om.kqchi, kqeta=om.kqeta,
                                  kcnrad=om.kcnrad, cap=om.cap, avg_cn=om.avg_cn, rvdw=om.rvdw, rcov=om.rcov, en=om.en,
                                  kbc=om.kbc, cutoff=om.cutoff, cn_exp=om.kcn, norm_exp=om.norm_exp, ctx=ctx)

    single = make(None).singlepoint(mol, want=("qvec", "energy", "gradient"))
    bounds = row_bounds(nat, world, multiple=8)
    assert np.all(np.diff(bounds) > 0)

    def body(r):
        model = make(group[r])
        a = mc.singlepoint_rows(model, mol, 0, nat, dqdr=False)
        b = mc.singlepoint_rows(model, mol, int(bounds[r]), int(bounds[r + 1]))
        return a, b

    res = run_ranks(world, body)
    o = oracle.eeqbc_singlepoint(om, omol, want=("qvec", "energy", "gradient", "dqdr"))
    grad = np.zeros((nat, 3))
    dqdr = np.zeros((nat, nat, 3))
    for r, (a, b) in enumerate(res):
        for key in ("qvec", "energy", "gradient", "sigma"):
            scale = max(1.0, np.abs(o[key]).max())
            assert np.abs(a[key] - o[key]).max() < 1e-10 * scale, (r, key)
            assert np.abs(a[key] - single[key]).max() < 1e-11 * scale, (r, key)
        jb, je = int(bounds[r]), int(bounds[r + 1])
        grad[jb:je] = b["gradient"]
        dqdr[:, jb:je, :] = b["dqdr"]
        assert np.abs(b["dqdL"] - o["dqdL"]).max() < 1e-10 * max(1.0, np.abs(o["dqdL"]).max())
    assert np.abs(grad - o["gradient"]).max() < 1e-10 * max(1.0, np.abs(o["gradient"]).max())
    assert np.abs(dqdr - o["dqdr"]).max() < 1e-10 * max(1.0, np.abs(o["dqdr"]).max())


def test_sharded_periodic_singlepoint_vs_oracle(mc, oracle, group, monkeypatch):
    """Periodic EEQ2019 (configs[3] path) shared by several ranks: Ewald pair tiles and rows of Phi W Phi^T per rank, ONE
    all-reduce of the matrix, pair tiles of the derivatives per rank joined with the CN-gradient rows; q, E, gradient,
    sigma against the one-context result and the oracle; dq/dr in full on every rank and by row blocks (one per rank)."""
    from synth import periodic_cell
    monkeypatch.setenv("MCHRG_B200_DIST_MIN_PBC", "128")
    world = len(group)
    num, xyz, lat = periodic_cell(7, 343, shear=0.1)
    mol = mc.Structure(num, xyz, 0.0, lattice=lat)
    single = mc.new_eeq2019_model(mol).singlepoint(mol, want=("qvec", "energy", "gradient"))

    from multicharge_b200.shard import row_bounds
    bounds = row_bounds(mol.nat, world, multiple=8)
    assert np.all(np.diff(bounds) > 0)

    def body(r):
        model = mc.new_eeq2019_model(mol, group[r])
        l0 = group[r].launch_count
        a = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
        launches[r] = group[r].launch_count - l0
        b = model.singlepoint(mol, want=("qvec", "energy", "gradient", "dqdr"))
        rows[r] = mc.singlepoint_rows(model, mol, int(bounds[r]), int(bounds[r + 1]))  # dq/dr rows of this rank
        return a, b

    rows = [None] * world

    launches = [0] * world
    res = run_ranks(world, body)
    # the shared path really ran: only rank 0 launches the diagonal / reciprocal-row kernels
    assert all(launches[r] < launches[0] for r in range(1, world)), launches
    omol = oracle.Mol(num, xyz, 0.0, lattice=lat)
    o = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient", "dqdr"))
    for r, (a, b) in enumerate(res):
        for key in ("cn", "qvec", "energy", "gradient", "sigma"):
            scale = max(1.0, np.abs(o[key]).max())
            assert np.abs(a[key] - o[key]).max() < 1e-10 * scale, (r, key)
            assert np.abs(a[key] - single[key]).max() < 1e-11 * scale, (r, key)
            assert np.abs(b[key] - o[key]).max() < 1e-10 * scale, (r, key, "dqdr call")
        for key in ("dqdr", "dqdL"):
            assert np.abs(b[key] - o[key]).max() < 1e-10 * max(1.0, np.abs(o[key]).max()), (r, key)
    dqdr = np.zeros_like(o["dqdr"])
    grad = np.zeros_like(o["gradient"])
    for r in range(world):
        jb, je = int(bounds[r]), int(bounds[r + 1])
        dqdr[:, jb:je, :] = rows[r]["dqdr"]
        grad[jb:je] = rows[r]["gradient"]
        for key in ("qvec", "energy", "sigma", "dqdL"):
            assert np.abs(rows[r][key] - o[key]).max() < 1e-10 * max(1.0, np.abs(o[key]).max()), (r, key, "rows")
    assert np.abs(dqdr - o["dqdr"]).max() < 1e-10 * max(1.0, np.abs(o["dqdr"]).max())
    assert np.abs(grad - o["gradient"]).max() < 1e-10 * max(1.0, np.abs(o["gradient"]).max())


def test_two_contexts_two_threads(mc, oracle):
    """A second context in the same process (what a Fortran caller driving several GPUs with OpenMP threads does):
    shared-memory attributes and launch state are per context, the calls are re-entrant."""
    from synth import batch, cluster
    ctxs = [mc.Context(0), mc.Context(0)]
    num, xyz = cluster(700, 9)
    mol = mc.Structure(num, xyz, 0.0)
    offsets, bnum, bxyz, charge = batch(300, 5, 20, 150)
    ref_b = oracle.eeq2019_singlepoint_batch(offsets, bnum, bxyz, charge)
    omol = oracle.Mol(num, xyz, 0.0)
    ref_c = oracle.eeq_singlepoint(oracle.new_eeq2019_model(omol), omol, want=("qvec", "energy", "gradient"))

    def body(r):
        ctx = ctxs[r]
        outs = []
        for _ in range(3):
            c = mc.new_eeq2019_model(mol, ctx).singlepoint(mol, want=("qvec", "energy", "gradient"))
            b = mc.singlepoint_batch(mc.new_eeq2019_batch_model(ctx), offsets, bnum, bxyz, charge)
            outs.append((c, b))
        return outs

    for outs in run_ranks(2, body):
        for c, b in outs:
            for key in ("qvec", "energy", "gradient"):
                assert np.abs(c[key] - ref_c[key]).max() < 1e-10 * max(1.0, np.abs(ref_c[key]).max())
                assert np.abs(b[key] - ref_b[key]).max() < 1e-10 * max(1.0, np.abs(ref_b[key]).max())
            assert int((b["status"] != 0).sum()) == 0
    for c in ctxs:
        c.close()


def test_env_knobs_are_read_per_call(mc, monkeypatch):
    """Tuning knobs must take effect between calls of one process (no process-wide caching)."""
    ctx = mc.default_context(0)
    n = 1024
    rng = np.random.default_rng(1)
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    Lref = np.linalg.cholesky(S)
    counts = []
    for tracked in ("0", "1"):  # two-stream look-ahead vs the tracked schedule of the distributed factorisation
        monkeypatch.setenv("MCHRG_B200_POTRF_TRACKED", tracked)
        l0 = ctx.launch_count
        a = fort(S)
        assert ctx.dpotrf(a) == 0
        assert np.abs(np.tril(a.T) - Lref).max() < 1e-12 * np.abs(Lref).max()
        counts.append(ctx.launch_count - l0)
    assert counts[0] != counts[1]
