"""Partitioning of independent molecules over ranks: molecules are independent units,
so every rank owns a disjoint subset and the data path needs no collective.

`partition_by_cost` deals molecules to ranks by estimated cost (n^3/3 factorisation + pair terms), largest
first onto the least loaded rank, so that heterogeneous batches stay balanced."""
import numpy as np

C_PAIR = 230.0  # flop per unique pair, see DESIGN.md "Algorithmic work"


def molecule_cost(sizes):
    n = np.asarray(sizes, dtype=np.float64)
    return n ** 3 / 3.0 + 4.0 * n ** 2 + C_PAIR * n * (n - 1) / 2.0


def partition_by_cost(sizes, world_size):
    """Returns a list of index arrays (one per rank), each sorted ascending; disjoint, covering all molecules."""
    sizes = np.asarray(sizes)
    cost = molecule_cost(sizes)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world_size)
    owner = np.empty(len(sizes), dtype=np.int64)
    # greedy longest-processing-time in blocks of equal size (cheap and within a few per cent of optimal)
    for idx in order:
        r = int(np.argmin(load))
        owner[idx] = r
        load[r] += cost[idx]
    return [np.flatnonzero(owner == r) for r in range(world_size)]


def gather_shard(offsets, num, xyz, charge, idx):
    """Extract the molecules `idx` of a concatenated batch: returns (offsets, num, xyz, charge) of the shard."""
    offsets = np.asarray(offsets)
    sizes = np.diff(offsets)[idx]
    new_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(sizes, out=new_off[1:])
    take = np.concatenate([np.arange(offsets[m], offsets[m + 1]) for m in idx]) if len(idx) else np.zeros(0, dtype=np.int64)
    return new_off, np.ascontiguousarray(num[take]), np.ascontiguousarray(xyz[take]), np.ascontiguousarray(charge[idx])


def scatter_results(nmol_total, offsets_total, idx, shard_offsets, shard_out, out):
    """Write a shard's per-atom results back into batch-ordered arrays (host side gather of outputs)."""
    for k, m in enumerate(idx):
        s = slice(shard_offsets[k], shard_offsets[k + 1])
        d = slice(offsets_total[m], offsets_total[m + 1])
        for key in ("qvec", "energy", "gradient"):
            if shard_out.get(key) is not None and out.get(key) is not None:
                out[key][d] = shard_out[key][s]
        out["status"][m] = shard_out["status"][k]
    return out


def block_cyclic_cholesky_host(w, n, panel, exchange):
    """Host (numpy) mirror of the multi-GPU factorisation schedule in csrc/dense.cu (dpotrf_dist): the matrix is
    replicated, block column b is owned by rank b % world, the owner factors panel b and the exchange broadcasts
    it; the broadcast of panel b+1 is started before the remaining trailing updates with panel b.  `w` is the flat
    column-major n x n matrix (lower triangle referenced), factored in place on every rank.  The device code makes
    exactly this sequence of exchange.start / exchange.wait calls (plus one more slot per panel for the inverted
    diagonal blocks); tests/test_exchange_gloo.py runs it over gloo with two ranks."""
    import numpy as np
    rank, world = exchange.rank, exchange.world
    a = w.reshape(n, n).T  # a[i, j] view of the column-major storage
    nblk = (n + panel - 1) // panel
    c0 = lambda b: b * panel  # noqa: E731
    pw = lambda b: min(panel, n - c0(b))  # noqa: E731

    def factor_panel(b):
        s, e = c0(b), c0(b) + pw(b)
        a[s:e, s:e] = np.linalg.cholesky(a[s:e, s:e])
        if e < n:
            a[e:, s:e] = np.linalg.solve(a[s:e, s:e], a[e:, s:e].T).T  # X L^T = B

    def start_panel(b):
        s = c0(b)
        lo, hi = s + s * n, (s + pw(b) - 1) * n + n  # diagonal element .. end of the last column: contiguous
        exchange.start(w[lo:hi], (hi - lo) * 8, b % world, b)

    def update(b, j):
        s, e, sj, ej = c0(b), c0(b) + pw(b), c0(j), c0(j) + pw(j)
        a[sj:, sj:ej] -= a[sj:, s:e] @ a[sj:ej, s:e].T

    if rank == 0 % world:
        factor_panel(0)
    start_panel(0)
    for b in range(nblk):
        exchange.wait(b)
        if b + 1 < nblk:
            if (b + 1) % world == rank:
                update(b, b + 1)
                factor_panel(b + 1)
            start_panel(b + 1)
        for j in range(b + 2, nblk):
            if j % world == rank:
                update(b, j)
    return w
