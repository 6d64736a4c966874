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


def row_bounds(nat, world, multiple=128):
    """Contiguous atom blocks of the row-sharded dq/dr solve (mchrg_b200_singlepoint_rows): world + 1 bounds,
    interior ones rounded down to a multiple of `multiple` (whole GEMM tiles per rank)."""
    b = np.linspace(0, nat, world + 1).astype(np.int64)
    b = (b // multiple) * multiple
    b[0], b[-1] = 0, nat
    return b


def group_width(n, world, group=512, tile=128):
    """Width of the column groups of the multi-GPU factorisation (dist_group_width in csrc/dense.cu): `group`
    columns, narrowed until every rank owns at least two groups."""
    db = group // tile * tile
    while db > tile and n // db < 2 * world:
        db -= tile
    return max(db, tile)


def block_cyclic_cholesky_host(w, n, group, sub, rank, world, bcast):
    """Host (numpy) mirror of the multi-GPU factorisation schedule in csrc/dense.cu (dpotrf_dist).

    Column group g = columns [g*group, (g+1)*group) is owned by rank g % world and factored in sub-panels of
    `sub` columns.  On entry a rank needs only the groups it owns (`w` may hold anything else elsewhere); the owner
    factors a sub-panel and `bcast(view, root)` sends it -- one contiguous range of the column-major storage, from
    the diagonal element to the end of the sub-panel's last column -- to every rank, in place.  The owner then
    updates the rest of its group, the owner of the NEXT group updates that group sub-panel by sub-panel, and once
    a group is complete every rank applies it to all later groups it owns (one rank-`group` update).  Every rank
    ends with the complete factor.  `w` is the flat column-major n x n matrix (lower triangle referenced), factored
    in place.  tests/test_dist_gloo.py runs this over gloo with two and three ranks."""
    a = w.reshape(n, n).T  # a[i, j] view of the column-major storage
    ngrp = (n + group - 1) // group
    nbcast = 0
    for g in range(ngrp):
        g0, gend = g * group, min(n, (g + 1) * group)
        mine = g % world == rank
        next_mine = g + 1 < ngrp and (g + 1) % world == rank
        h0, hend = gend, min(n, gend + group)
        for c0 in range(g0, gend, sub):
            c1 = min(gend, c0 + sub)
            if mine:
                a[c0:c1, c0:c1] = np.linalg.cholesky(np.tril(a[c0:c1, c0:c1]) + np.tril(a[c0:c1, c0:c1], -1).T)
                if c1 < n:
                    a[c1:, c0:c1] = np.linalg.solve(a[c0:c1, c0:c1], a[c1:, c0:c1].T).T  # X L^T = B
            lo, hi = c0 + c0 * n, (c1 - 1) * n + n  # diagonal element .. end of the last column: contiguous
            bcast(w[lo:hi], g % world)
            nbcast += 1
            if mine and c1 < gend:
                a[c1:, c1:gend] -= a[c1:, c0:c1] @ a[c1:gend, c0:c1].T
            if next_mine:
                a[h0:, h0:hend] -= a[h0:, c0:c1] @ a[h0:hend, c0:c1].T
        for j in range(g + 2, ngrp):
            if j % world == rank:
                j0, jend = j * group, min(n, (j + 1) * group)
                a[j0:, j0:jend] -= a[j0:, g0:gend] @ a[j0:jend, g0:gend].T
    return nbcast


# ---------------------------------------------------------------------------------------------------------------
# Periodic structures on several ranks (pbc_amat / pbc_derivs in csrc/eeq_pbc.cu): what each rank evaluates
# ---------------------------------------------------------------------------------------------------------------
def pair_tiles(nat, tile=32):
    """The 32 x 32 atom tiles of the lower block triangle in the order of the device kernels (tile_coords in
    csrc/eeq_pbc.cu): index b = ti (ti + 1) / 2 + tj, tj <= ti."""
    nt = (nat + tile - 1) // tile
    return [(ti, tj) for ti in range(nt) for tj in range(ti + 1)]


def tile_share(ntiles, rank, world):
    """Tiles of one rank: round robin (tile0 = rank, stride = world), as launched by pbc_amat / pbc_derivs."""
    return range(rank, ntiles, world)


def rowblock_share(nblocks, rank, world):
    """128-row blocks of the reciprocal-space product Phi W Phi^T computed by one rank."""
    return range(nblocks * rank // world, nblocks * (rank + 1) // world)


def periodic_matrix_shared_host(nat, pair, recip, diag, rank, world, allreduce, tile=32, rowblock=128):
    """numpy mirror of the shared assembly of pbc_amat: this rank's pair tiles (both triangles), its row blocks of the
    reciprocal-space matrix and its share of the diagonal into a zero matrix, joined by ONE all-reduce.
    pair(i, j) -> direct part of A(i, j), i > j; recip -> full (nat, nat) reciprocal matrix (only this rank's row
    blocks are read); diag(i) -> self term; allreduce(array) sums in place over the ranks."""
    a = np.zeros((nat, nat))
    tiles = pair_tiles(nat, tile)
    for b in tile_share(len(tiles), rank, world):
        ti, tj = tiles[b]
        for i in range(ti * tile, min(nat, (ti + 1) * tile)):
            for j in range(tj * tile, min(nat, (tj + 1) * tile)):
                if j < i:
                    v = pair(i, j)
                    a[i, j] += v
                    a[j, i] += v
    nblk = (nat + rowblock - 1) // rowblock
    for rb in rowblock_share(nblk, rank, world):
        lo, hi = rb * rowblock, min(nat, (rb + 1) * rowblock)
        a[lo:hi, :] += recip[lo:hi, :]
    r = row_bounds(nat, world, multiple=8)
    for i in range(int(r[rank]), int(r[rank + 1])):
        a[i, i] += diag(i)
    allreduce(a)
    return a
