#!/usr/bin/env python3
"""Host <-> device copy ceilings of a multi-GPU box by NUMA placement of the pinned buffers.

One process, one thread per GPU.  Prints the topology the container sees, then for host buffers bound to every
allowed NUMA node (mmap + mbind + cudaHostRegister): per-GPU H2D / D2H bandwidth alone, and the aggregate with ALL
GPUs copying in both directions at once (what the end-to-end arm of bench.py needs: 781 MB per GPU and step).
Usage: python tools/numa_probe.py [MB per buffer]
"""
import ctypes as C
import glob
import mmap
import os
import subprocess
import sys
import threading
import time

import torch

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NB = MB << 20
libc = C.CDLL("libc.so.6", use_errno=True)
SYS_MBIND = 237
MPOL_BIND = 2


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as exc:  # noqa: BLE001
        return f"<{exc}>"


def node_buffer(node):
    """NB bytes of anonymous memory bound to `node` (None: default policy), touched, registered with CUDA."""
    m = mmap.mmap(-1, NB, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = C.addressof(C.c_char.from_buffer(m))
    if node is not None:
        mask = C.c_ulong(1 << node)
        rc = libc.syscall(SYS_MBIND, C.c_void_p(addr), C.c_ulong(NB), C.c_int(MPOL_BIND), C.byref(mask), C.c_ulong(64), C.c_uint(0))
        if rc != 0:
            raise OSError(C.get_errno(), f"mbind(node {node}) failed: {os.strerror(C.get_errno())}")
    C.memset(addr, 1, NB)
    rc = torch.cuda.cudart().cudaHostRegister(addr, NB, 0)
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister -> {rc}")
    t = torch.frombuffer(m, dtype=torch.uint8)
    return m, t


def main():
    ng = torch.cuda.device_count()
    print("GPUs", ng, "| cpus allowed", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
    print(sh("grep -i 'allowed_list' /proc/self/status"))
    print("nodes online:", sh("cat /sys/devices/system/node/online"), "| has cpu:", sh("cat /sys/devices/system/node/has_cpu"),
          "| has memory:", sh("cat /sys/devices/system/node/has_memory"))
    for g in range(ng):
        pr = torch.cuda.get_device_properties(g)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        print(f"GPU {g} {bus} numa_node", sh(f"cat /sys/bus/pci/devices/{bus}/numa_node"), "local_cpulist",
              sh(f"cat /sys/bus/pci/devices/{bus}/local_cpulist"))
    print(sh("nvidia-smi topo -m | head -24"))
    nodes = []
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        nodes.append(int(p.rsplit("node", 1)[1]))
    placements = [None] + nodes
    bufs = {}
    for node in placements:
        try:
            bufs[node] = [node_buffer(node) for _ in range(2)]  # [0] H2D source, [1] D2H target
        except Exception as exc:  # noqa: BLE001
            print(f"node {node}: {exc}")
    dev_bufs = []
    for g in range(ng):
        with torch.cuda.device(g):
            dev_bufs.append((torch.empty(NB, dtype=torch.uint8, device=f"cuda:{g}"), torch.ones(NB, dtype=torch.uint8, device=f"cuda:{g}"),
                             torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)))

    def copy(g, node, h2d=True, d2h=True, reps=3):
        din, dout, s_in, s_out = dev_bufs[g]
        hin, hout = bufs[node][0][1], bufs[node][1][1]
        with torch.cuda.device(g):
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s_in):
                        din.copy_(hin, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s_out):
                        hout.copy_(dout, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()

    # per GPU, alone
    for node in bufs:
        line = []
        for g in range(ng):
            copy(g, node, reps=1)
            t0 = time.perf_counter(); copy(g, node, True, False); t1 = time.perf_counter(); copy(g, node, False, True); t2 = time.perf_counter()
            copy(g, node, True, True); t3 = time.perf_counter()
            line.append(f"{3 * NB / (t1 - t0) / 1e9:5.1f}/{3 * NB / (t2 - t1) / 1e9:5.1f}/{6 * NB / (t3 - t2) / 1e9:5.1f}")
        print(f"host memory on node {node}: per GPU alone, GB/s h2d/d2h/both:", "  ".join(line), flush=True)

    # all GPUs at once, both directions
    def concurrent(choice, label):
        bar = threading.Barrier(ng + 1)
        def run(g):
            copy(g, choice(g), reps=1)
            bar.wait()
            copy(g, choice(g), reps=4)
            bar.wait()
        th = [threading.Thread(target=run, args=(g,)) for g in range(ng)]
        for t in th: t.start()
        bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
        for t in th: t.join()
        print(f"all {ng} GPUs at once, both directions, {label}: aggregate {ng * 8 * NB / dt / 1e9:6.1f} GB/s", flush=True)

    for node in bufs:
        concurrent(lambda g, node=node: node, f"host memory on node {node}")
    gpu_node = []
    for g in range(ng):
        pr = torch.cuda.get_device_properties(g)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        try:
            nd = int(sh(f"cat /sys/bus/pci/devices/{bus}/numa_node"))
        except ValueError:
            nd = -1
        gpu_node.append(nd if nd in bufs else None)
    concurrent(lambda g: gpu_node[g], f"host memory on each GPU's own node {gpu_node}")
    # concurrency of the bound pairs sharing a buffer is fine for a bandwidth probe (contents are not checked)


if __name__ == "__main__":
    main()
