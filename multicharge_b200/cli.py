"""Command line single point with the argument set and result file of the reference's program
(app/main.f90:45-135, 205-270; json_results in output.f90:143-196), on the B200 path:

    python -m multicharge_b200.cli [options] <input>
      -m, --model <eeq2019|eeqbc2025>   charge model (default eeq2019)
      -i, --input <format>              xyz | coord (tmol, turbomole); default: from the file name
      -c, --charge <value>              molecular charge (else `.CHRG` next to the working directory, else the file's)
      -g, --grad                        gradient and virial (and dq/dr, dq/dL as the reference allocates them)
      -j, --json                        write `multicharge.json`: version, energy, gradient, charges,
                                        "coordination numbers" (the keys test/validation/tester.py compares)
      --eeqbc-tables <file.json>        pairwise vdW radii / Pauling EN for eeqbc2025 ({"rvdw": [[..]], "en": [..]}
                                        per species in order of first appearance; they live in mctc-lib)

Readers: xyz (Angstrom) and Turbomole coord ($coord [angs|bohr|frac], $periodic, $lattice, $cell, $eht charge=).
Only what the hot path needs is restated; this is a test harness for the reference's validation cases, not a
re-implementation of mctc-lib's I/O."""
import argparse
import json
import math
import os
import sys

import numpy as np

AATOAU = 1.0 / 0.529177210903  # Angstrom -> Bohr (CODATA 2018 Bohr radius, mctc_io_convert / mctc_io_constants)
VERSION = "0.5.0"            # result-file schema of the reference version this mirrors

SYMBOLS = ("h he li be b c n o f ne na mg al si p s cl ar k ca sc ti v cr mn fe co ni cu zn ga ge as se br kr rb sr y zr "
           "nb mo tc ru rh pd ag cd in sn sb te i xe cs ba la ce pr nd pm sm eu gd tb dy ho er tm yb lu hf ta w re os ir pt "
           "au hg tl pb bi po at rn fr ra ac th pa u np pu am cm bk cf es fm md no lr").split()


def atomic_number(sym):
    s = "".join(ch for ch in sym.strip().lower() if ch.isalpha())
    if s in SYMBOLS:
        return SYMBOLS.index(s) + 1
    if sym.strip().isdigit() and 1 <= int(sym) <= len(SYMBOLS):
        return int(sym)
    raise ValueError(f"unknown element '{sym}'")


def read_xyz(text):
    lines = text.splitlines()
    nat = int(lines[0].split()[0])
    num, xyz = [], []
    for line in lines[2:2 + nat]:
        tok = line.split()
        num.append(atomic_number(tok[0]))
        xyz.append([float(t) * AATOAU for t in tok[1:4]])
    if len(num) != nat:
        raise ValueError("xyz file is shorter than its atom count")
    return dict(num=num, xyz=np.array(xyz), charge=None, lattice=None, periodic=None)


def cell_to_lattice(a, b, c, alpha, beta, gamma):
    """Lattice vectors (rows) from cell lengths and angles in degrees: a along x, b in the xy plane."""
    al, be, ga = (math.radians(v) for v in (alpha, beta, gamma))
    cx = c * math.cos(be)
    cy = c * (math.cos(al) - math.cos(be) * math.cos(ga)) / math.sin(ga)
    cz = math.sqrt(max(c * c - cx * cx - cy * cy, 0.0))
    return np.array([[a, 0.0, 0.0], [b * math.cos(ga), b * math.sin(ga), 0.0], [cx, cy, cz]])


def read_coord(text):
    groups, key = {}, None
    for line in text.splitlines():
        s = line.strip()
        if not s or s.startswith("#"):
            continue
        if s.startswith("$"):
            key = s.split()[0][1:].lower()
            groups.setdefault(key, {"args": s.split()[1:], "lines": []})
            if key == "end":
                break
            continue
        if key:
            groups[key]["lines"].append(s)
    if "coord" not in groups:
        raise ValueError("no $coord group")
    unit = (groups["coord"]["args"] or ["bohr"])[0].lower()
    num, xyz = [], []
    for s in groups["coord"]["lines"]:
        tok = s.split()
        xyz.append([float(t) for t in tok[:3]])
        num.append(atomic_number(tok[3]))
    xyz = np.array(xyz)
    lattice, periodic = None, None
    nper = int(groups["periodic"]["args"][0]) if "periodic" in groups and groups["periodic"]["args"] else 0
    if nper not in (0, 3):
        raise ValueError("only $periodic 0 and 3 are supported by this reader")
    if nper == 3:
        if "lattice" in groups:
            lu = (groups["lattice"]["args"] or ["bohr"])[0].lower()
            lattice = np.array([[float(t) for t in s.split()[:3]] for s in groups["lattice"]["lines"][:3]])
            if lu.startswith("ang"):
                lattice = lattice * AATOAU
        elif "cell" in groups:
            cu = (groups["cell"]["args"] or ["bohr"])[0].lower()
            v = [float(t) for t in " ".join(groups["cell"]["lines"]).split()[:6]]
            scale = AATOAU if cu.startswith("ang") else 1.0
            lattice = cell_to_lattice(v[0] * scale, v[1] * scale, v[2] * scale, v[3], v[4], v[5])
        else:
            raise ValueError("$periodic 3 needs $lattice or $cell")
        periodic = [True] * 3
    if unit.startswith("ang"):
        xyz = xyz * AATOAU
    elif unit.startswith("frac"):
        if lattice is None:
            raise ValueError("fractional coordinates need a lattice")
        xyz = xyz @ lattice
    charge = None
    if "eht" in groups:
        for tok in " ".join(groups["eht"]["args"] + groups["eht"]["lines"]).split():
            if tok.lower().startswith("charge="):
                charge = float(tok.split("=", 1)[1])
    return dict(num=num, xyz=xyz, charge=charge, lattice=lattice, periodic=periodic)


def read_structure(path, fmt=None):
    text = sys.stdin.read() if path == "-" else open(path, encoding="utf8").read()
    if fmt is None:
        base = os.path.basename(path).lower()
        fmt = "xyz" if (path == "-" or base.endswith(".xyz")) else "coord"
    fmt = fmt.lower()
    if fmt == "xyz":
        return read_xyz(text)
    if fmt in ("coord", "tmol", "turbomole"):
        return read_coord(text)
    raise ValueError(f"unsupported input format '{fmt}' (xyz, coord)")


def json_results(path, energy=None, gradient=None, charges=None, cn=None):
    """The reference's result file (output.f90:143-196): same keys, gradient flattened atom by atom."""
    out = {"version": VERSION}
    if energy is not None:
        out["energy"] = float(energy)
    if gradient is not None:
        out["gradient"] = [float(v) for v in np.asarray(gradient).ravel()]
    if charges is not None:
        out["charges"] = [float(v) for v in charges]
    if cn is not None:
        out["coordination numbers"] = [float(v) for v in cn]
    with open(path, "w", encoding="utf8") as f:
        json.dump(out, f, indent=2)
        f.write("\n")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="multicharge", description="Electronegativity equilibration model for atomic charges "
                                 "(B200 path behind the reference's command line)")
    ap.add_argument("input")
    ap.add_argument("-m", "-model", "--model", default="eeq2019")
    ap.add_argument("-i", "-input", "--input", dest="fmt", default=None)
    ap.add_argument("-c", "-charge", "--charge", type=float, default=None)
    ap.add_argument("-g", "-grad", "--grad", action="store_true")
    ap.add_argument("-j", "-json", "--json", action="store_true")
    ap.add_argument("-v", "-version", "--version", action="version", version=f"multicharge version {VERSION} (mchrg-b200)")
    ap.add_argument("--eeqbc-tables", default=None)
    args = ap.parse_args(argv)

    try:
        geo = read_structure(args.input, args.fmt)
    except (OSError, ValueError, IndexError) as exc:
        print(exc, file=sys.stderr)
        return 1
    charge = args.charge
    if charge is None and os.path.exists(".CHRG"):  # app/main.f90:63-80
        try:
            charge = float(open(".CHRG", encoding="utf8").read().split()[0])
            print("[Info] Molecular charge read from '.CHRG'\n")
        except (ValueError, IndexError):
            print("[Warn] Could not read molecular charge read from '.CHRG'\n")
    if charge is None:
        charge = geo["charge"] if geo["charge"] is not None else 0.0

    import multicharge_b200 as mc  # needs the CUDA library: fails loudly without it
    mol = mc.Structure(geo["num"], geo["xyz"], charge, lattice=geo["lattice"], periodic=geo["periodic"])
    model_id = args.model.lower()
    try:
        if model_id == "eeq2019":
            model = mc.new_eeq2019_model(mol)
        elif model_id == "eeqbc2025":
            if not args.eeqbc_tables:
                raise ValueError("eeqbc2025 needs --eeqbc-tables (pairwise vdW radii and Pauling EN live in mctc-lib)")
            tab = json.load(open(args.eeqbc_tables, encoding="utf8"))
            model = mc.new_eeqbc2025_model(mol, np.array(tab["rvdw"]), np.array(tab["en"]))
        else:
            raise ValueError("Invalid model was choosen.")
        want = ("qvec", "energy", "gradient", "dqdr") if args.grad else ("qvec", "energy")
        res = model.singlepoint(mol, want=want)
    except (ValueError, mc.MchrgError) as exc:
        print(exc, file=sys.stderr)
        return 1

    sym = [SYMBOLS[z - 1].capitalize() for z in geo["num"]]
    print("Electrostatic properties (in atomic units)")
    print("-" * 40)
    print(f"{'#':>6} {'Z':>4}      {'CN':>10} {'q':>10}")
    print("-" * 40)
    for i, (z, s) in enumerate(zip(geo["num"], sym)):
        print(f"{i + 1:6d} {z:4d} {s:<4s} {res.cn[i]:10.4f} {res.qvec[i]:10.4f}")
    print("-" * 40)
    print(f"\nElectrostatic energy: {res.energy.sum():22.14e} Eh")
    if args.grad:
        print(f"Gradient norm:        {np.linalg.norm(res.gradient):22.14e} Eh/a0")
        print("Virial:")
        for row in res.sigma:
            print("  " + " ".join(f"{v:16.8e}" for v in row))
    if args.json:
        json_results("multicharge.json", energy=res.energy.sum(), gradient=res.gradient if args.grad else None,
                     charges=res.qvec, cn=res.cn)
        print("[Info] JSON dump of results written to 'multicharge.json'")
    return 0


if __name__ == "__main__":
    sys.exit(main())
