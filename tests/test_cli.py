"""Command line front end (multicharge_b200/cli.py): readers and result file on CPU; on the GPU the reference's three
validation cases are run the way test/validation/tester.py runs them (program + coord + --grad --json, result file
compared key by key at its 5e-7)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THR = 5.0e-7  # tester.py:43

H3P = """$coord
   -0.47073898552969      0.81534384004086      0.00000000000000      h
   -0.47073898552969     -0.81534384004086      0.00000000000000      h
    0.94147797105939      0.00000000000000      0.00000000000000      h
$eht charge=+1
$end
"""


def write_coord(path, g):
    from multicharge_b200.cli import SYMBOLS
    with open(path, "w", encoding="utf8") as f:
        f.write("$coord\n")
        for z, r in zip(g["num"], g["xyz"]):
            f.write(f" {r[0]!r} {r[1]!r} {r[2]!r} {SYMBOLS[z - 1]}\n")
        f.write(f"$eht charge={g['charge']!r}\n$end\n")


def test_readers_and_result_file(tmp_path):
    from multicharge_b200 import cli
    g = cli.read_coord(H3P)
    assert g["num"] == [1, 1, 1] and g["charge"] == 1.0 and g["lattice"] is None
    assert g["xyz"][2, 0] == 0.94147797105939
    x = cli.read_xyz("2\ncomment\nO 0.0 0.0 0.0\nH 0.0 0.0 0.9572\n")
    assert x["num"] == [8, 1] and abs(x["xyz"][1, 2] - 0.9572 / 0.529177210903) < 1e-14
    per = cli.read_coord("$coord frac\n 0.5 0.5 0.5 na\n 0 0 0 cl\n$periodic 3\n$lattice angs\n 4 0 0\n 0 4 0\n 0 0 4\n$end\n")
    assert per["periodic"] == [True] * 3 and np.allclose(per["lattice"], np.eye(3) * 4 * cli.AATOAU)
    assert np.allclose(per["xyz"][0], 2 * cli.AATOAU) and per["num"] == [11, 17]
    lat = cli.cell_to_lattice(3.0, 4.0, 5.0, 90.0, 90.0, 120.0)
    assert np.allclose(np.linalg.norm(lat, axis=1), [3, 4, 5]) and abs(lat[0] @ lat[1] - 12 * np.cos(np.radians(120))) < 1e-12
    assert cli.atomic_number("Lr") == 103 and cli.atomic_number("c1") == 6
    with pytest.raises(ValueError):
        cli.atomic_number("xx")
    out = tmp_path / "multicharge.json"
    cli.json_results(out, energy=1.5, gradient=np.arange(6.0).reshape(2, 3), charges=[0.5, -0.5], cn=[1.0, 1.0])
    r = json.load(open(out))
    assert list(r) == ["version", "energy", "gradient", "charges", "coordination numbers"] and r["gradient"] == [0, 1, 2, 3, 4, 5]


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if n.startswith("validation")])
def test_validation_cases_through_the_command_line(tmp_path, name):
    g = load_golden(name)
    write_coord(tmp_path / "coord", g)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    stat = subprocess.call([sys.executable, "-m", "multicharge_b200.cli", "coord", "--grad", "--json"], cwd=tmp_path, env=env)
    assert stat == 0
    res = json.load(open(tmp_path / "multicharge.json"))
    ref = {"energy": g["energy"], "gradient": np.ravel(g["gradient"]), "charges": g["charges"], "coordination numbers": g["cn"]}
    for key, val in ref.items():
        assert key in res, key
        assert np.abs(np.array(res[key]) - np.array(val)).max() < THR, key


@pytest.mark.gpu
def test_command_line_charge_file_and_errors(tmp_path):
    g = load_golden("validation_01_h3p")
    write_coord(tmp_path / "coord", dict(g, charge=0.0))
    (tmp_path / ".CHRG").write_text("1\n")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    run = lambda *a: subprocess.call([sys.executable, "-m", "multicharge_b200.cli", *a], cwd=tmp_path, env=env)  # noqa: E731
    assert run("coord", "--json") == 0
    res = json.load(open(tmp_path / "multicharge.json"))
    assert abs(sum(res["charges"]) - 1.0) < 1e-12 and "gradient" not in res
    assert run("coord", "-c", "0", "--json") == 0
    assert abs(sum(json.load(open(tmp_path / "multicharge.json"))["charges"])) < 1e-12
    assert run("coord", "-m", "nonsense") != 0
    assert run("coord", "-m", "eeqbc2025") != 0  # tables not supplied
    assert run("missing_file") != 0
