#!/usr/bin/env python3
"""Benchmark of the EEQ charge hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload batch|cluster]

Headline workload (N = 1 and every N): BASELINE.json configs[1] -- EEQ2019 single points (charges + energy +
gradient) for a batch of 100 000 synthetic 30-200-atom molecules per GPU; a "step" is one pass of the hot path
over that batch.  Molecules are independent, so ranks get disjoint batches and there is no data-path collective
("scaling": "weak").

  value         whole-job molecules/s with the batch already resident in HBM (CUDA events on the library's
                stream, max over ranks)
  e2e           the same through the public C-ABI call with pinned HOST buffers: H2D of the inputs and D2H of
                charges / energies / gradients / status inside the timed region
  roofline      the dominant kernel (eeq_factor_ll_kernel, the per-molecule factorisation) alone: algorithmic flops
                / its CUDA-event time (MCHRG_B200_TRACE=1 runs the three batch kernels one after the other, each
                timed on its own) against the FP64 peak (cuBLAS DGEMM 8192^3 measured live: MEASURED_PEAKS.json has
                no FP64 entry); roofline_kernels holds the same for the two pair kernels
  cpu_baseline  the CPU oracle (port of the reference's OpenMP+LAPACK path; the Fortran reference cannot be built
                in this image) on a bounded sample of the same batch, all host cores
  extra         everything else BASELINE.json names, at every N:
                  batch_strong   the SAME 100 000 molecules split over the N ranks by cost (multicharge_b200.shard)
                  cluster        configs[2]: one N = 16384 cluster -- q+E+gradient seconds and q+E+gradient+dq/dr
                                 seconds through singlepoint_rows (block-cyclic Cholesky with native NCCL panel
                                 broadcasts, row-sharded O(N^2) passes joined by all-reduce, row-sharded dq/dr),
                                 the factorisation alone, each with its own roofline
                  periodic       configs[3]: periodic N = 4096 cell, q+E+gradient (Ewald pair tiles shared between the ranks)
                                 and + dq/dr, dq/dL (dq/dr rows sharded over the ranks)
                  eeqbc          configs[4] size: EEQ-BC cluster (N = 1: largest single-GPU size; N > 1: shared)

`--impl reference` times the CPU oracle alone on the same workload (bounded sample per step) and prints the same
line shape.  `--workload cluster` prints the configs[2] block as its own line (strong scaling of one system).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "EEQ2019 q+E+gradient single points, batched molecules/s"
UNIT = "molecules/s"
NMOL = 100_000
NMIN, NMAX = 30, 200
SEED = 20250102
CPU_SAMPLE = 100_000  # molecules in the bounded CPU sample (~10 s on the box's 16 cores: one full batch)
CPU_BUDGET_S = 150.0  # --impl reference: the K + W steps together stay below this

# algorithmic FP64 work per molecule of n atoms (DESIGN.md "Algorithmic work"):
#   factorisation kernel: Cholesky n^3/3 + two triangular solves with 2 right-hand sides 4 n^2   (dense contraction)
#   pair kernels: per unique pair, phase A erf + rsqrt x2 + CN ring (C_PAIR_A), phase C erf + exp + 60 flop (C_PAIR_C)
C_PAIR_A = 40 + 30
C_PAIR_C = 40 + 2 * 25 + 60


def factor_flops(sizes):
    n = sizes.astype(np.float64)
    return float((n ** 3 / 3.0 + 4.0 * n ** 2).sum())


def pair_flops(sizes, c_pair):
    n = sizes.astype(np.float64)
    return float((c_pair * n * (n - 1) / 2.0).sum())


def batch_config(world, nmol, ntot_rank0):
    """The `config` object of the headline line -- identical in both arms (ours / reference)."""
    return {"workload": "configs[1]: 100k synthetic 30-200-atom molecules per GPU, EEQ2019 q+E+gradient",
            "molecules_per_gpu": nmol, "atoms_per_gpu": ntot_rank0, "seed": SEED,
            "l2": "inputs+outputs of one step (0.78 GB) exceed the 126 MB L2 (no flush needed)",
            "parallelism": f"independent molecules, {world} rank(s), no collective"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows[lo:hi]:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(ngpu):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads to the CPUs next to its GPU (the PCI device's local_cpulist) BEFORE any pinned
    buffer is allocated, so that first-touch places the staging memory on the GPU's NUMA node -- what `numactl
    --cpunodebind/--membind` does for a one-process-per-GPU launch.  Returns a short description for the bench line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"cpus {spec} (GPU {bus})"
        return f"unchanged ({len(allowed)} cpus allowed, GPU-local {spec or 'unknown'})"
    except Exception as exc:  # noqa: BLE001 -- affinity is an optimisation, never a failure
        return f"unavailable ({type(exc).__name__})"


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3, best of 5 (the FP64 roofline denominator; not in MEASURED_PEAKS.json)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    best = 1e30
    for it in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1) * 1e-3)
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e12


def hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200 (driver-written), GB/s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        return 6500.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def cpu_reference_rate(nsample, seed, threads=None):
    """CPU oracle on the first `nsample` molecules of the bench batch (rank 0's): molecules/s, threads, seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from synth import batch
    offsets, num, xyz, charge = batch(NMOL if nsample <= NMOL else nsample, seed, NMIN, NMAX)
    offsets, charge = offsets[:nsample + 1], charge[:nsample]
    num, xyz = num[:offsets[-1]], xyz[:offsets[-1]]
    orc.eeq2019_singlepoint_batch(offsets[:9], num[:offsets[8]], xyz[:offsets[8]], charge[:8], threads=threads)  # warm
    t0 = time.perf_counter()
    out = orc.eeq2019_singlepoint_batch(offsets, num, xyz, charge, threads=threads)
    dt = time.perf_counter() - t0
    assert int((out["status"] != 0).sum()) == 0
    return nsample / dt, out["threads"], dt


def run_reference(args):
    """The reference arm: the CPU port of the reference path on the SAME workload (same generator, same seed as rank 0
    of the GPU arm); every step is the first `nsample` molecules of that batch, sized so that K + W steps stay within
    CPU_BUDGET_S on this box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from synth import batch
    offsets = batch(args.nmol, SEED, NMIN, NMAX)[0]
    probe, thr, _ = cpu_reference_rate(2000, SEED)
    nsample = int(min(args.nmol, max(1000, probe * CPU_BUDGET_S / max(1, args.steps + args.warmup))))
    rates = []
    for it in range(args.warmup + args.steps):
        rate, thr, _ = cpu_reference_rate(nsample, SEED)
        if it >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    sample = (f"the first {nsample} of the {args.nmol} molecules of the GPU arm's rank-0 batch per step "
              f"(U{{{NMIN}..{NMAX}}} atoms), all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * nsample / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": batch_config(args.gpus, args.nmol, int(offsets[-1])),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# single-system blocks (configs[2..4]); every rank calls them, rank 0 reports
# ---------------------------------------------------------------------------------------------------------
def _timed(torch, ext, fn, reps, world, dev):
    """best-of-`reps` CUDA-event seconds of fn() on the library's stream, max over ranks per repetition"""
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        torch.cuda.synchronize()
        best = min(best, max_over_ranks(e0.elapsed_time(e1) * 1e-3, world, dev))
    return best


def cluster_block(torch, mc, ctx, ext, dev, rank, world, nat, fp64_peak, want_dqdr=True):
    """configs[2]: one N-atom cluster.  q+E+gradient (every rank ends with everything) and q+E+gradient+dq/dr (dq/dr
    rows sharded over the ranks, gradient rows all-gathered through the library's communicator)."""
    from synth import cluster
    from multicharge_b200.shard import row_bounds
    num, xyz = cluster(nat, SEED + 3)
    mol = mc.Structure(num, xyz, 0.0)
    model = mc.new_eeq2019_model(mol, ctx)
    txyz = torch.tensor(mol.xyz, device=dev)
    tid = torch.tensor(mol.id, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    res = {"nat": nat, "n_gpus": world,
           "parallelism": "one GPU" if world == 1 else
           (f"block-cyclic Cholesky over {world} ranks (native NCCL broadcast per sub-panel), row-sharded CN / pair / "
            "CN-gradient passes joined by ncclAllReduce, dq/dr rows sharded (no collective), gradient rows all-gathered")}
    o = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                           gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64))
    l0 = ctx.launch_count
    mc.singlepoint_rows(model, mol, 0, nat, out=o, xyz=txyz, id=tid)
    res["q_e_grad_launches"] = ctx.launch_count - l0
    res["q_e_grad_seconds"] = _timed(torch, ext, lambda: mc.singlepoint_rows(model, mol, 0, nat, out=o, xyz=txyz, id=tid), 3,
                                     world, dev)
    res["charge_sum"] = float(o["qvec"].sum().item())
    del o
    # factorisation alone: blocked Cholesky of a dense SPD matrix of the same order (device resident; shared by the
    # ranks when a communicator is installed)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    m = torch.randn(nat, nat, generator=g, **f64)
    s = m @ m.T + nat * torch.eye(nat, **f64)
    del m
    w = s.clone()

    t_chol = 1e30
    for it in range(3):
        w.copy_(s)
        torch.cuda.synchronize()
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        ctx.dpotrf(w)
        e1.record(ext)
        torch.cuda.synchronize()
        if it >= 1:
            t_chol = min(t_chol, max_over_ranks(e0.elapsed_time(e1) * 1e-3, world, dev))
    res["cholesky_seconds_incl_pad_copies"] = t_chol
    res["roofline_cholesky"] = {"bound": "tensor", "kernel": "dpotrf (dgemm_dmma_kernel trailing updates + potrf_leaf_kernel)",
                                "achieved": nat ** 3 / 3.0 / t_chol / 1e12 / world, "peak": fp64_peak,
                                "unit": "TFLOP/s per GPU", "frac": nat ** 3 / 3.0 / t_chol / 1e12 / world / fp64_peak,
                                "traffic": None, "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"}
    del s, w
    torch.cuda.empty_cache()
    if want_dqdr:
        bounds = row_bounds(nat, world)
        jb, je = int(bounds[rank]), int(bounds[rank + 1])
        nj = je - jb
        out = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                                 gradient=torch.empty(nj, 3, **f64), sigma=torch.zeros(3, 3, **f64),
                                 dqdr=torch.empty(nat, nj, 3, **f64), dqdL=torch.empty(nat, 3, 3, **f64))
        nmax = int(np.diff(bounds).max())
        gsend = torch.zeros(nmax, 3, **f64)
        grecv = torch.empty(world, nmax, 3, **f64)

        def step():
            mc.singlepoint_rows(model, mol, jb, je, out=out, xyz=txyz, id=tid)
            if world > 1:
                gsend[:nj].copy_(out["gradient"])
                ctx.comm_allgather(gsend, grecv)

        t = _timed(torch, ext, step, 2, world, dev)
        res["q_e_grad_dqdr_seconds"] = t
        res["rows_per_rank"] = [int(x) for x in np.diff(bounds)]
        # the dq/dr stage = both right-side triangular sweeps over this rank's rows: 2 * (3 nj + 9) * n^2 flop
        t_dq = max(t - res["q_e_grad_seconds"], 1e-9)
        fl = 2.0 * (3.0 * nmax + 9.0) * float(nat) ** 2
        res["roofline_dqdr"] = {"bound": "tensor", "kernel": "dgemm_dmma_kernel (two right-side TRSM sweeps X = B L^-T L^-1)",
                                "achieved": fl / t_dq / 1e12, "peak": fp64_peak, "unit": "TFLOP/s per GPU",
                                "frac": fl / t_dq / 1e12 / fp64_peak,
                                "traffic": 5.72e9 / 3.90e9 * (2.0 * 8.0 * (3.0 * nmax + 9.0) * nat) * (nat / 512.0 / 2.0 + 1.0),
                                "traffic_source": "ncu --set full of one trailing-update-shaped launch (15360 x 15360 x 512, beta = 1): "
                                                  "5.72 GB of DRAM traffic for 3.90 GB of algorithmic C read+write and operand reads "
                                                  "(profiles/r2_gemm_ncu_key_metrics.txt), applied to the read-modify-write passes of the "
                                                  "recursive right-side sweeps over this rank's rows",
                                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"}
        del out, gsend, grecv
        torch.cuda.empty_cache()
    return res


def _singlepoint_device(torch, mc, model, mol, dev, dq):
    """device-resident single point through mchrg_b200_singlepoint; returns the closure that runs it"""
    import ctypes as C
    from multicharge_b200 import _lib
    nat = mol.nat
    f64 = dict(dtype=torch.float64, device=dev)
    b = dict(cn=torch.empty(nat, **f64), qloc=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
             gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64), qvec=torch.empty(nat, **f64))
    if dq:
        b["dqdr"] = torch.empty(nat, nat, 3, **f64)
        b["dqdL"] = torch.empty(nat, 3, 3, **f64)
    txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
    s = mol._c(xyz=txyz, id=tid)
    p = lambda k: b[k].data_ptr() if k in b else None  # noqa: E731

    def run():
        rc = _lib.load().mchrg_b200_singlepoint(model._h, C.byref(s), p("cn"), p("qloc"), p("energy"), p("gradient"),
                                                p("sigma"), p("qvec"), p("dqdr"), p("dqdL"))
        if rc != 0:
            raise RuntimeError(f"singlepoint failed: {rc}")
    run.keep = (b, txyz, tid, s)
    return run


def periodic_block(torch, mc, ctx, ext, dev, rank, world, fp64_peak):
    """configs[3]: periodic EEQ2019, 16^3 = 4096 jittered sites in a cubic cell, Ewald sums; q+E+gradient+sigma (on
    several GPUs the Ewald pair tiles and the rows of Phi W Phi^T are shared between the ranks, one all-reduce of the
    matrix; every rank ends with everything) and q+dq/dr+dq/dL (the cpq path; on several GPUs the dq/dr rows are
    sharded over the ranks through singlepoint_rows, dq/dL on every rank)."""
    from synth import periodic_cell
    from multicharge_b200.shard import row_bounds
    num, xyz, lat = periodic_cell(16, 20250104)
    mol = mc.Structure(num, xyz, 0.0, lattice=lat)
    model = mc.new_eeq2019_model(mol, ctx)
    nat = mol.nat
    res = {"nat": nat, "n_gpus": world}
    run = _singlepoint_device(torch, mc, model, mol, dev, False)
    res["q_e_grad_seconds"] = _timed(torch, ext, run, 2, world, dev)
    res["charge_sum"] = float(run.keep[0]["qvec"].sum().item())
    del run
    torch.cuda.empty_cache()
    if world == 1:
        run = _singlepoint_device(torch, mc, model, mol, dev, True)
        res["q_e_grad_dqdr_dqdL_seconds"] = _timed(torch, ext, run, 2, world, dev)
        del run
    else:
        bounds = row_bounds(nat, world)
        jb, je = int(bounds[rank]), int(bounds[rank + 1])
        f64 = dict(dtype=torch.float64, device=dev)
        txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
        out = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                                 gradient=torch.empty(je - jb, 3, **f64), sigma=torch.zeros(3, 3, **f64),
                                 dqdr=torch.empty(nat, je - jb, 3, **f64), dqdL=torch.empty(nat, 3, 3, **f64))
        res["q_e_grad_dqdr_dqdL_seconds"] = _timed(
            torch, ext, lambda: mc.singlepoint_rows(model, mol, jb, je, out=out, xyz=txyz, id=tid), 2, world, dev)
        res["rows_per_rank"] = [int(x) for x in np.diff(bounds)]
        del out
    torch.cuda.empty_cache()
    n = float(nat)
    fl = n ** 3 / 3.0 + 2.0 * n * n * 256.0
    res["roofline"] = {"bound": "fp64-alu", "kernel": "pbc_amat_tiles_kernel / pbc_damat_tiles_kernel (Ewald real-space pair tiles)",
                       "note": "pair kernels are FP64-ALU / issue bound; dense part (Cholesky + Phi W Phi^T) for scale",
                       "achieved": fl / res["q_e_grad_seconds"] / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                       "frac": fl / res["q_e_grad_seconds"] / 1e12 / fp64_peak, "traffic": None}
    return res


def eeqbc_block(torch, mc, ctx, ext, dev, rank, world, nat, fp64_peak, dq):
    """configs[4]: EEQ-BC 2025 on a synthetic cluster (vdW radii / EN from the fixed synthetic species tables of
    tools/synth.py: the mctc-lib tables are not available offline).  q+E+gradient on every rank (distributed
    factorisation, row-sharded O(N^2) passes when world > 1); with dq: + dq/dr, rows sharded over the ranks."""
    from synth import cluster, synthetic_vdw_en
    from multicharge_b200.shard import row_bounds
    num, xyz = cluster(nat, 20250105)
    mol = mc.Structure(num, xyz, 0.0)
    rvdw, en = synthetic_vdw_en(mol.num)
    model = mc.new_eeqbc2025_model(mol, rvdw, en * 3.98, ctx)
    txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    res = {"nat": nat, "n_gpus": world}
    n = float(nat)
    if not dq:
        o = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                               gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64))
        res["q_e_grad_seconds"] = _timed(torch, ext, lambda: mc.singlepoint_rows(model, mol, 0, nat, out=o, xyz=txyz, id=tid),
                                         1, world, dev)
        res["charge_sum"] = float(o["qvec"].sum().item())
        res["roofline_q_e_grad"] = {"bound": "tensor", "kernel": "dpotrf (DMMA) + bc_* pair kernels (FP64 ALU)",
                                    "achieved": n ** 3 / 3.0 / res["q_e_grad_seconds"] / 1e12 / world, "peak": fp64_peak,
                                    "unit": "TFLOP/s per GPU (Cholesky flops only)",
                                    "frac": n ** 3 / 3.0 / res["q_e_grad_seconds"] / 1e12 / world / fp64_peak, "traffic": None}
        return res
    bounds = row_bounds(nat, world)
    jb, je = int(bounds[rank]), int(bounds[rank + 1])
    nj = je - jb
    out = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                             gradient=torch.empty(nj, 3, **f64), sigma=torch.zeros(3, 3, **f64),
                             dqdr=torch.empty(nat, nj, 3, **f64), dqdL=torch.empty(nat, 3, 3, **f64))
    res["q_e_grad_dqdr_seconds"] = _timed(torch, ext, lambda: mc.singlepoint_rows(model, mol, jb, je, out=out, xyz=txyz, id=tid),
                                          1, world, dev)
    res["rows_per_rank"] = [int(x) for x in np.diff(bounds)]
    fl = n ** 3 / 3.0 + 12.0 * n ** 3  # Cholesky + dtmpdr C GEMM (6 n^3) + two TRSM sweeps (6 n^3)
    res["roofline_dqdr"] = {"bound": "tensor", "kernel": "dgemm_dmma_kernel (dtmpdr C, TRSM sweeps)",
                            "achieved": fl / res["q_e_grad_dqdr_seconds"] / 1e12 / world, "peak": fp64_peak,
                            "unit": "TFLOP/s per GPU", "frac": fl / res["q_e_grad_dqdr_seconds"] / 1e12 / world / fp64_peak,
                            "traffic": None}
    return res


def batch_strong_block(torch, mc, ctx, ext, dev, rank, world, nmol, steps):
    """The SAME nmol molecules (rank 0's batch of the weak run) split over the ranks by estimated cost
    (multicharge_b200.shard.partition_by_cost); device-resident and end-to-end rates of the whole job."""
    from synth import batch
    from multicharge_b200 import shard
    model = mc.new_eeq2019_batch_model(ctx)
    offsets, num, xyz, charge = batch(nmol, SEED, NMIN, NMAX)
    if world > 1:
        idx = shard.partition_by_cost(np.diff(offsets), world)[rank]
        offsets, num, xyz, charge = shard.gather_shard(offsets, num, xyz, charge, idx)
    nloc, ntot = len(charge), int(offsets[-1])
    d_in = [torch.tensor(a, device=dev) for a in (offsets, num, xyz, charge)]
    d_out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
                 gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(nloc, dtype=torch.int32, device=dev))
    h_in = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in (offsets, num, xyz, charge)]
    h_out = dict(qvec=torch.empty(ntot, dtype=torch.float64).pin_memory().numpy(),
                 energy=torch.empty(ntot, dtype=torch.float64).pin_memory().numpy(),
                 gradient=torch.empty(ntot, 3, dtype=torch.float64).pin_memory().numpy(),
                 status=torch.empty(nloc, dtype=torch.int32).pin_memory().numpy())
    res = {"molecules_total": nmol, "molecules_this_rank": nloc}
    for tag, fn in (("value", lambda: mc.singlepoint_batch(model, *d_in, out=d_out)),
                    ("e2e", lambda: mc.singlepoint_batch(model, *h_in, out=h_out))):
        fn()
        fn()
        torch.cuda.synchronize()
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        torch.cuda.synchronize()
        barrier(world)
        res[tag] = nmol * steps / max_over_ranks(e0.elapsed_time(e1) * 1e-3, world, dev)
    res["unit"] = UNIT
    res["scaling"] = "strong"
    return res


def run_cluster(args):
    """configs[2] as its own line: one N-atom cluster over the ranks (strong scaling of one system)."""
    import torch
    import multicharge_b200 as mc
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = mc.Context(local)
    if world > 1 and not args.replicated_factor:
        mc.init_comm_from_torch(ctx)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    fp64_peak = measure_fp64_peak(torch, dev)
    res = cluster_block(torch, mc, ctx, ext, dev, rank, world, args.nat, fp64_peak)
    ctx.comm_destroy()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    sec = res["q_e_grad_dqdr_seconds"]
    line = {"metric": "EEQ2019 q+E+gradient+dq/dr time, one N-atom cluster", "value": sec, "unit": "s", "n_gpus": world,
            "steps": 2, "warmup": 1, "ms_per_step": 1e3 * sec, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2]: one synthetic cluster, N = {args.nat}, q + E + gradient + full dq/dr",
                       "parallelism": res["parallelism"]},
            "extra": res, "roofline": res.get("roofline_dqdr"), "gpu_launches": ctx.launch_count}
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import multicharge_b200 as mc
    from synth import batch

    rank, world, local = dist_setup(args.gpus)
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(torch, local) if world > 1 else "not applied (one rank)"
    ctx = mc.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    model = mc.new_eeq2019_batch_model(ctx)

    nmol = args.nmol
    offsets, num, xyz, charge = batch(nmol, SEED + 17 * rank, NMIN, NMAX)
    sizes = np.diff(offsets)
    ntot = int(offsets[-1])

    # ---- device-resident arm -------------------------------------------------------------------
    d_in = (torch.tensor(offsets, device=dev), torch.tensor(num, device=dev), torch.tensor(xyz, device=dev),
            torch.tensor(charge, device=dev))
    d_out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
                 gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(nmol, dtype=torch.int32, device=dev))

    def step_dev():
        mc.singlepoint_batch(model, *d_in, out=d_out)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    barrier(world)
    torch.cuda.synchronize()
    mark0 = sampler.mark()
    l0 = ctx.launch_count
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record(ext)
    for k in range(args.steps):
        step_dev()
    e_end.record(ext)
    torch.cuda.synchronize()
    barrier(world)
    launches = ctx.launch_count - l0
    total_s = e_beg.elapsed_time(e_end) * 1e-3
    total_s = max_over_ranks(total_s, world, dev)
    mark1 = sampler.mark()
    nfail = int((d_out["status"] != 0).sum().item())
    q_dev = d_out["qvec"].cpu().numpy().copy()

    # per-kernel times: the three batch kernels one after the other over the whole batch, each timed alone with CUDA
    # events on the launching stream inside the library (MCHRG_B200_TRACE=1)
    os.environ["MCHRG_B200_TRACE"] = "1"
    try:
        step_dev()
        best = None
        for _ in range(3):
            step_dev()
            ms, nl = ctx.batch_trace()
            best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    finally:
        os.environ.pop("MCHRG_B200_TRACE", None)
    t_a, t_f, t_c = [x * 1e-3 for x in best]

    # ---- end-to-end arm: pinned host buffers through the public call -----------------------------
    h_in = [torch.from_numpy(a).pin_memory() for a in (offsets, num, xyz, charge)]
    h_out = dict(qvec=torch.empty(ntot, dtype=torch.float64).pin_memory(), energy=torch.empty(ntot, dtype=torch.float64).pin_memory(),
                 gradient=torch.empty(ntot, 3, dtype=torch.float64).pin_memory(), status=torch.empty(nmol, dtype=torch.int32).pin_memory())
    n_in = [t.numpy() for t in h_in]
    n_out = {k: v.numpy() for k, v in h_out.items()}
    h2d = int(sum(a.nbytes for a in n_in))
    d2h = int(sum(a.nbytes for a in n_out.values()))

    def step_e2e():
        mc.singlepoint_batch(model, *n_in, out=n_out)

    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    torch.cuda.synchronize()
    barrier(world)
    e_steps = max(2, min(args.steps, 5))
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(ext)
    for _ in range(e_steps):
        step_e2e()
    b1.record(ext)
    torch.cuda.synchronize()
    barrier(world)
    e2e_s = max_over_ranks(b0.elapsed_time(b1) * 1e-3, world, dev)
    # parity spot check of what the timed calls produced (device arm vs host arm, same inputs) -- before the copy test
    # below reuses the host buffers
    same = bool(np.array_equal(n_out["qvec"], q_dev))
    # what the host <-> device fabric of this box gives when ALL ranks move one step's bytes at the same time and nothing
    # else runs: H2D of the inputs and D2H of the outputs on two streams, pinned buffers (the ceiling of the e2e arm)
    d_stage = [torch.empty_like(t, device=dev) for t in h_in]
    d_res = {k: torch.empty_like(v, device=dev) for k, v in h_out.items()}
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copies():
        with torch.cuda.stream(s_in):
            for a, b in zip(d_stage, h_in):
                a.copy_(b, non_blocking=True)
        with torch.cuda.stream(s_out):
            for k in h_out:
                h_out[k].copy_(d_res[k], non_blocking=True)

    copies()
    torch.cuda.synchronize()
    barrier(world)
    t0c = time.perf_counter()
    for _ in range(3):
        copies()
    torch.cuda.synchronize()
    copy_s = max_over_ranks((time.perf_counter() - t0c) / 3, world, dev)
    del d_stage, d_res
    clocks = None
    if rank == 0:
        time.sleep(0.3)
        clocks = sampler.stop(mark0, max(mark1, mark0 + 1))
    del d_in, d_out, h_in, h_out, n_in, n_out
    torch.cuda.empty_cache()

    # ---- everything else the metric names, at this N ----------------------------------------------
    fp64_peak = measure_fp64_peak(torch, dev)
    extra = {}
    if not args.no_extra:
        if world > 1:
            mc.init_comm_from_torch(ctx)  # native NCCL communicator of the single-system path
        blocks = [("batch_strong", lambda: batch_strong_block(torch, mc, ctx, ext, dev, rank, world, nmol, 3)),
                  ("cluster", lambda: cluster_block(torch, mc, ctx, ext, dev, rank, world, args.nat, fp64_peak))]
        blocks.append(("periodic", lambda: periodic_block(torch, mc, ctx, ext, dev, rank, world, fp64_peak)))
        # configs[4]: N = 32768; the full dq/dr of that size needs the ranks' memory and flops together (25.8 GB per
        # (3, N, N) array, 4.6e14 flop): one GPU reports it at N = 8192
        nat_dq = args.nat_bc if world >= 2 else min(args.nat_bc, 8192)
        blocks.append(("eeqbc", lambda: eeqbc_block(torch, mc, ctx, ext, dev, rank, world, args.nat_bc, fp64_peak, False)))
        blocks.append(("eeqbc_dqdr", lambda: eeqbc_block(torch, mc, ctx, ext, dev, rank, world, nat_dq, fp64_peak, True)))
        if args.extras:
            blocks = [b for b in blocks if b[0] in args.extras.split(",")]
        for name, fn in blocks:
            # a failure on one rank would leave the others inside a collective: agree on the outcome per block
            try:
                extra[name] = fn()
            except Exception as exc:  # keep the headline line even if a large case fails
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
                if world > 1:
                    raise
            torch.cuda.empty_cache()
        ctx.comm_destroy()

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    cpu_rate, cpu_thr, cpu_dt = cpu_reference_rate(args.cpu_sample, SEED)
    hbm, hbm_src = hbm_peak()
    fl_f = factor_flops(sizes)
    peak_src = "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"
    alg_bytes = 28.0 * ntot + 40.0 * ntot  # xyz + Z in, q + E + gradient out
    line = {
        "metric": METRIC, "value": world * nmol * args.steps / total_s, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": batch_config(world, nmol, ntot),
        "step_note": "every step is a complete single point of all molecules (coordination numbers, Coulomb blocks, factorisations, "
                     "energies, gradients are recomputed; outputs are overwritten); what the library derives from the molecule offsets "
                     "alone -- size classes, launch order, scratch layout -- is kept on the context and reused while the offsets "
                     "stay the same",
        "e2e": {"value": world * nmol * e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e_steps, "host_equals_device_result": same, "host_affinity": affinity,
                "copy_only_ceiling": {"value": world * nmol / copy_s, "unit": UNIT, "aggregate_gb_s": world * (h2d + d2h) / copy_s / 1e9,
                                      "note": "all ranks moving one step's H2D + D2H bytes concurrently from / to pinned buffers, "
                                              "no kernels: the host<->device fabric bound of the e2e arm on this box"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "eeq_factor_ll_kernel (per-molecule Cholesky + substitutions, DMMA)",
                     "achieved": fl_f / t_f / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl_f / t_f / 1e12 / fp64_peak,
                     "traffic": NCU_FACTOR_DRAM_BYTES_PER_MOLECULE * nmol,
                     "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src,
                     "algorithmic_flops_per_step": fl_f, "kernel_seconds_per_step": t_f, "launches_per_step": nl[1],
                     "algorithmic_bytes_per_step": alg_bytes},
        "roofline_kernels": [
            {"kernel": "eeq_batch_kernel<256,1> (pair phase A: Coulomb block, CN, xvec)", "bound": "fp64-alu / issue",
             "seconds_per_step": t_a, "achieved": pair_flops(sizes, C_PAIR_A) / t_a / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
             "frac": pair_flops(sizes, C_PAIR_A) / t_a / 1e12 / fp64_peak,
             "note": f"{C_PAIR_A} flop per unique pair (erf series 40 + distances / rsqrt 30); ncu: issue slots 75 % busy"},
            {"kernel": "eeq_batch_kernel<256,3> (pair phase C: energy, gradient)", "bound": "fp64-alu / issue",
             "seconds_per_step": t_c, "achieved": pair_flops(sizes, C_PAIR_C) / t_c / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
             "frac": pair_flops(sizes, C_PAIR_C) / t_c / 1e12 / fp64_peak,
             "note": f"{C_PAIR_C} flop per unique pair (erf 40 + 2 exp 50 + 60); ncu: issue slots 75 % busy"},
            {"kernel": "whole step vs HBM", "bound": "hbm", "achieved": alg_bytes / (total_s / args.steps) / 1e9, "peak": hbm,
             "unit": "GB/s", "frac": alg_bytes / (total_s / args.steps) / 1e9 / hbm, "peak_source": hbm_src,
             "note": "algorithmic bytes (28 B/atom in + 40 B/atom out) over the step time: the path is compute bound"}],
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cpu_thr, "kind": "port",
                         "sample": f"the first {args.cpu_sample} molecules of rank 0's batch in {cpu_dt:.1f} s"},
        "failed_molecules": nfail,
    }
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    return 0


# DRAM traffic of the factorisation kernel per molecule of the bench distribution, from the ncu --set full capture
# under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum of one launch over 5000 molecules)
NCU_FACTOR_DRAM_BYTES_PER_MOLECULE = (489.03e6 + 333.94e6) / 5000
NCU_TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum of one eeq_factor_ll_kernel launch over 5000 molecules of this "
                      "distribution (profiles/r2_batch_factor_ncu_key_metrics.txt), scaled to the molecules of one step; the "
                      "per-molecule scratch matrices (~60 KB each, 300 MB per chunk) do not stay in the 126 MB L2 between the "
                      "three kernels of a chunk")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nmol", type=int, default=NMOL)
    ap.add_argument("--cpu-sample", type=int, default=CPU_SAMPLE)
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--extras", default="", help="comma-separated subset of the extra blocks to run (default: all)")
    ap.add_argument("--workload", default="batch", choices=["batch", "cluster"])
    ap.add_argument("--nat", type=int, default=16384, help="atoms of the configs[2] cluster")
    ap.add_argument("--nat-bc", type=int, default=32768, help="atoms of the configs[4] EEQ-BC cluster")
    ap.add_argument("--replicated-factor", action="store_true",
                    help="cluster workload: no communicator (every rank factors the whole matrix)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cluster":
        return run_cluster(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
