#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` (SASS view) dump by CUDA source line, using the line table
of the locally built object (`nvdisasm -g -c`).  Usage: ncu_source_by_line.py <source.csv> <object.o> <kernel-substr>"""
import csv
import re
import subprocess
import sys
import tempfile
import os
import collections


def line_table(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    table, cur, on = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    src, obj, kernel = sys.argv[1:4]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    table = line_table(obj, kernel)
    blocks, cur = [], None
    with open(src) as f:
        for row in csv.reader(f):
            if row and row[0] == "Kernel Name":
                cur = {"name": row[1], "rows": []}
                blocks.append(cur)
            elif row and row[0] == "Address":
                cur["hdr"] = row
            elif row and cur is not None:
                cur["rows"].append(row)
    ktag = os.environ.get("NCU_KERNEL_NAME") or kernel.replace("Li", "(int)").rstrip("E")
    blk = [b for b in blocks if ktag in b["name"]][which]
    hdr = blk["hdr"]
    ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(blk["rows"][0][ia], 16) if blk["rows"][0][ia].startswith("0x") else int(blk["rows"][0][ia])
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_i = 0
    for r in blk["rows"]:
        a = (int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])) - base
        line = table.get(a, (None, ""))[0]
        s, n = int(r[isamp] or 0), int(r[iinst] or 0)
        agg[line][0] += s
        agg[line][1] += n
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                agg[line][2][hdr[i]] += v
        tot_s += s
        tot_i += n
    print(f"kernel {blk['name']}  samples={tot_s} warp-instructions={tot_i}")
    for line, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
        print(f"  {str(line):32s} samples {100 * s / tot_s:5.1f}%  inst {100 * n / max(tot_i, 1):5.1f}%   {top}")


if __name__ == "__main__":
    main()
