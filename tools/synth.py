"""Synthetic structures for tests and bench: seeded, deterministic.

Clusters: simple-cubic grid, spacing 2.6 Bohr, uniform jitter +-0.35 Bohr per coordinate (minimum
distance ~1.9 Bohr, so the Coulomb block stays positive definite), species from {H, C, N, O} with
p = {.5, .3, .1, .1}.
"""
import numpy as np

SPACING = 2.6
JITTER = 0.35
PALETTE = np.array([1, 6, 7, 8], dtype=np.int32)
PROBS = np.array([0.5, 0.3, 0.1, 0.1])


def cluster(nat, seed, palette=PALETTE, probs=PROBS):
    """Compact jittered-grid cluster of `nat` atoms: returns (num[nat], xyz[nat,3])."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(nat ** (1.0 / 3.0)))
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
    # keep the sites closest to the centre so the fragment is compact
    centre = (side - 1) / 2.0
    order = np.argsort(((grid - centre) ** 2).sum(axis=1), kind="stable")
    sites = grid[order[:nat]]
    xyz = sites * SPACING + rng.uniform(-JITTER, JITTER, size=(nat, 3))
    num = rng.choice(palette, size=nat, p=probs).astype(np.int32)
    return num, np.ascontiguousarray(xyz)


def periodic_cell(ncell, seed, shear=0.0, palette=PALETTE, probs=PROBS):
    """ncell^3 jittered sites in a cubic cell a = ncell * 2.6 Bohr (optionally sheared):
    returns (num, xyz, lattice[3,3] with lattice[k] = k-th cell vector)."""
    rng = np.random.default_rng(seed)
    grid = np.stack(np.meshgrid(*[np.arange(ncell)] * 3, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
    a = ncell * SPACING
    lattice = np.eye(3) * a
    if shear:
        lattice[1, 0] = shear * a
        lattice[2, 1] = shear * a
    frac = (grid + 0.5) / ncell
    xyz = frac @ lattice + rng.uniform(-JITTER, JITTER, size=(len(grid), 3))
    num = rng.choice(palette, size=len(grid), p=probs).astype(np.int32)
    return num, np.ascontiguousarray(xyz), lattice


def batch(nmol, seed, nmin=30, nmax=200):
    """nmol independent molecules with nat ~ U{nmin..nmax}, charge in {-1, 0, +1}.
    Returns (offsets[nmol+1] int64, num[ntot] int32, xyz[ntot,3], charge[nmol])."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(nmin, nmax + 1, size=nmol)
    offsets = np.zeros(nmol + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    ntot = int(offsets[-1])
    num = rng.choice(PALETTE, size=ntot, p=PROBS).astype(np.int32)
    xyz = np.empty((ntot, 3))
    # grid templates per size are cached: only the jitter differs between molecules of equal size
    cache = {}
    for m in range(nmol):
        n = int(sizes[m])
        if n not in cache:
            side = int(np.ceil(n ** (1.0 / 3.0)))
            grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
            centre = (side - 1) / 2.0
            order = np.argsort(((grid - centre) ** 2).sum(axis=1), kind="stable")
            cache[n] = grid[order[:n]] * SPACING
        xyz[offsets[m]:offsets[m + 1]] = cache[n]
    xyz += rng.uniform(-JITTER, JITTER, size=xyz.shape)
    charge = rng.integers(-1, 2, size=nmol).astype(np.float64)
    return offsets, num, xyz, charge


def synthetic_vdw_en(num):
    """Deterministic stand-ins for the mctc-lib tables that are not available offline (pairwise D3 vdW radii
    x autoaa, Pauling EN / 3.98), per species: NOT the reference's numbers, only fixed inputs (SURVEY.md 8d,
    "rvdw from per-species table (synthetic but fixed)").  Same formula as the oracle's helper of that name."""
    z = np.asarray(num, dtype=np.float64)
    r = 1.1 + 0.35 * np.log1p(z)
    rvdw = 0.5 * (r[:, None] + r[None, :]) * 0.529177210903 * 3.2
    en = (0.8 + 2.6 * np.exp(-((z % 8) - 6.5) ** 2 / 18.0)) / 3.98
    return np.ascontiguousarray(rvdw), en
