#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the EEQ hot path into tests/golden/*.json.

Runs in the build container only (reads /root/reference); the JSON files are committed so that
nothing touches /root/reference at test time.  Sources:

  * test/validation/{01_h3+,02_heada,03_etlicl-}/{coord,multicharge_ref.json}
      EEQ2019 CLI results (energy, gradient, charges, coordination numbers; official
      tolerance 5e-7 abs, test/validation/tester.py:43)
  * test/unit/test_model.f90:934-981   EEQ2019 charges for the Fr..Lr "actinides" geometry
      (tolerance 100*eps, test_model.f90:31)
  * test/unit/test_model.f90:1159-1278 inline geometries H2+ and [ZnOOH]- used by the
      finite-difference consistency tests (no reference numbers)
"""
import json
import re
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

SYM = ("h he li be b c n o f ne na mg al si p s cl ar k ca sc ti v cr mn fe co ni cu zn ga ge as se br kr "
       "rb sr y zr nb mo tc ru rh pd ag cd in sn sb te i xe cs ba la ce pr nd pm sm eu gd tb dy ho er tm yb lu "
       "hf ta w re os ir pt au hg tl pb bi po at rn fr ra ac th pa u np pu am cm bk cf es fm md no lr").split()


def read_coord(path):
    num, xyz = [], []
    on = False
    for line in path.read_text().splitlines():
        if line.startswith("$coord"):
            on = True
            continue
        if line.startswith("$"):
            on = False
            continue
        if on and line.strip():
            x, y, z, s = line.split()[:4]
            xyz.append([float(x), float(y), float(z)])
            num.append(SYM.index(s.lower()) + 1)
    return num, xyz


def fnum(tok):
    return float(tok.replace("_wp", ""))


def fortran_reals(block):
    return [fnum(t) for t in re.findall(r"[-+]?\d+\.\d*(?:[eE][-+]?\d+)?_wp", block)]


def main():
    # --- validation fixtures
    for name, charge in (("01_h3+", 1.0), ("02_heada", 0.0), ("03_etlicl-", -1.0)):
        d = REF / "test" / "validation" / name
        num, xyz = read_coord(d / "coord")
        ref = json.loads((d / "multicharge_ref.json").read_text())
        out = {"source": f"test/validation/{name}", "model": "eeq2019", "charge": charge, "num": num, "xyz": xyz,
               "energy": ref["energy"], "gradient": ref["gradient"], "charges": ref["charges"],
               "cn": ref["coordination numbers"], "tolerance_abs": 5.0e-7, "reference_version": ref["version"]}
        (OUT / f"validation_{name.replace('+', 'p').replace('-', 'm')}.json").write_text(json.dumps(out, indent=1))

    src = (REF / "test" / "unit" / "test_model.f90").read_text()

    # --- actinides known-answer test (EEQ2019)
    m = re.search(r"subroutine test_eeq_q_actinides.*?end subroutine test_eeq_q_actinides", src, re.S)
    blk = m.group(0)
    qref = fortran_reals(re.search(r"ref\(17\) = \[(.*?)\]", blk, re.S).group(1))
    num = [int(t) for t in re.findall(r"\d+", re.search(r"mol%num = \[(.*?)\]", blk, re.S).group(1))]
    xyz = fortran_reals(re.search(r"mol%xyz = reshape\(\[(.*?)\],", blk, re.S).group(1))
    assert len(qref) == 17 and len(num) == 17 and len(xyz) == 51
    out = {"source": "test/unit/test_model.f90:934-981", "model": "eeq2019", "charge": 0.0, "num": num,
           "xyz": [xyz[3 * i:3 * i + 3] for i in range(17)], "charges": qref, "tolerance_abs": 100 * 2.220446049250313e-16}
    (OUT / "kat_eeq_actinides.json").write_text(json.dumps(out, indent=1))

    # --- inline FD geometries
    for sub, tag in (("test_g_h2plus", "h2plus"), ("test_eeq_dadr_znooh", "znooh")):
        m = re.search(r"subroutine " + sub + r".*?end subroutine " + sub, src, re.S)
        blk = m.group(0)
        charge = fnum(re.search(r"charge = ([-+0-9.]+_wp)", blk).group(1))
        num = [int(t) for t in re.findall(r"\d+", re.search(r"num\(nat\) = \[(.*?)\]", blk).group(1))]
        xyz = fortran_reals(re.search(r"xyz\(3, nat\) = reshape\(\[(.*?)\],", blk, re.S).group(1))
        out = {"source": f"test/unit/test_model.f90 {sub}", "charge": charge, "num": num,
               "xyz": [xyz[3 * i:3 * i + 3] for i in range(len(num))]}
        (OUT / f"geom_{tag}.json").write_text(json.dumps(out, indent=1))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
