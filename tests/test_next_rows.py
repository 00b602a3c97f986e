"""SURVEY.md §8(f) "next" rows: the serde-layout JSON of a state and the CLI option surface.
The JSON layout cannot be checked against the Rust binary here (it cannot be built); these tests
pin the structure derived from the reference's type definitions, and the round trip."""
import json
import math
import os
import subprocess
import sys

import pytest

import crystal_packing_b200 as cp
from crystal_packing_b200 import cli
from crystal_packing_b200.serialise import format_f64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("v,text", [
    (8.0, "8.0"), (0.1, "0.1"), (1e-7, "1e-7"), (1e-5, "0.00001"), (1e16, "1e16"), (1e15, "1000000000000000.0"),
    (-0.25, "-0.25"), (1.5707963267948966, "1.5707963267948966"), (5e-324, "5e-324"), (-0.0, "-0.0"),
    (6.123233995736766e-17, "6.123233995736766e-17"), (123456.789, "123456.789"),
])
def test_float_format_matches_serde_json(v, text):  # ryu, as serde_json prints f64
    assert format_f64(v) == text and float(text) == v


def test_state_json_layout():
    state = cp.PackedState.from_group(cp.LineShape.polygon(4), "p1g1")
    d = json.loads(state.to_json())
    assert list(d) == ["wallpaper", "shape", "cell", "occupied_sites"]          # packed.rs:24-33
    assert d["wallpaper"] == {"name": "p1m1", "family": "Orthorhombic"}         # the p1g1 label quirk
    assert list(d["cell"]) == ["length", "ratio", "angle", "family"]            # cell.rs:48-54
    assert d["cell"]["length"] == 8.0 and d["cell"]["angle"] == math.pi / 2
    site = d["occupied_sites"][0]
    assert list(site) == ["wyckoff", "x", "y", "angle"]                          # site.rs:13-19
    assert list(site["wyckoff"]) == ["letter", "symmetries", "num_rotations", "mirror_primary", "mirror_secondary"]
    # "-x,y+1/2" as a column-major 3x3 with an all-zero last row (transform.rs:139)
    assert site["wyckoff"]["symmetries"][1] == [-1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.5, 0.0]
    assert d["shape"]["name"] == "Polygon" and list(d["shape"]["items"][0]) == ["start", "end"]
    lj = cp.PotentialState.from_group(cp.LJShape2.circle(), "p1")
    item = json.loads(lj.to_json())["shape"]["items"][0]
    assert item == {"position": [0.0, 0.0], "sigma": 1.0, "epsilon": 1.0, "cutoff": None}
    discs = cp.PackedState.from_group(cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p2")
    assert list(json.loads(discs.to_json())["shape"]["items"][1]) == ["position", "radius"]


@pytest.mark.parametrize("make", [
    lambda: cp.PackedState.from_group(cp.LineShape.polygon(7), "p2gg"),
    lambda: cp.PackedState.from_group(cp.LineShape.from_radial("star", [1.0, 0.5] * 3), "p2mg"),
    lambda: cp.PackedState.from_group(cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p1m1"),
    lambda: cp.PotentialState.from_group(cp.LJShape2.from_trimer(0.63, 120.0, 1.0), "p2"),
])
def test_state_json_round_trip(make):
    state = make()
    state.dof = [v * 0.9 if i < 2 else v for i, v in enumerate(state.dof)]
    text = state.to_json()
    back = type(state).from_json(text)
    assert back.to_json() == text and back.dof == state.dof
    assert back.shape.area() == state.shape.area()
    assert back.shape.enclosing_radius() == state.shape.enclosing_radius()
    assert back.wallpaper.key == state.wallpaper.key


def test_cli_option_surface():  # main.rs:27-64,154-213
    ap = cli.build_parser()
    a = ap.parse_args(["p2", "--outfile", "o", "polygon"])
    assert (a.replications, a.steps, a.kt_start, a.kt_finish, a.kt_ratio, a.max_step_size, a.inner_steps,
            a.convergence, a.potential, a.sides) == (100, 100, 0.1, None, None, 0.01, 1000, None, "Hard", 4)
    a = ap.parse_args(["p2gg", "-p", "LJ", "--outfile", "o", "-s", "500", "--kt-finish", "0.001", "trimer", "-r", "0.7"])
    assert (a.shape, a.radius, a.angle, a.distance, a.steps) == ("trimer", 0.7, 120.0, 1.0, 500)
    opt = cli.make_optimiser(a)
    assert opt.pipeline()[1].kt_ratio == math.pow(0.001 / 0.1, 1 / 500)
    assert isinstance(cli.make_state(a), cp.PotentialState)
    with pytest.raises(SystemExit):  # main.rs:330-332
        cli.make_state(ap.parse_args(["p2", "-p", "LJ", "--outfile", "o", "polygon"]))
    with pytest.raises(cp.CrystalPackingError):  # "p2m1" is advertised but not parsable (wallpaper.rs:100-132)
        cli.make_state(ap.parse_args(["p2m1", "--outfile", "o", "circle"]))


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    """conda-recipe/meta.yaml:17-20 smoke (`packing p2 -s 10 --outfile test.out trimer`) plus a run
    whose result must equal the oracle's best replica."""
    import orc

    env = dict(os.environ, PYTHONPATH=ROOT)
    out = tmp_path / "test.out"
    r = subprocess.run([sys.executable, "-m", "crystal_packing_b200.cli", "p2", "-s", "10", "--outfile", str(out),
                        "trimer"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert json.loads(open(str(out) + ".json").read())["shape"]["name"] == "Trimer"
    svg = open(str(out) + ".svg").read()  # main.rs:266
    assert svg.startswith("<svg ") and svg.count("<use ") == 9 + 9 * 2 and svg.count("<circle ") == 3
    out2 = tmp_path / "pent"
    r = subprocess.run([sys.executable, "-m", "crystal_packing_b200.cli", "p2", "-s", "600", "--inner-steps", "200",
                        "--replications", "16", "--outfile", str(out2), "polygon", "--sides", "5"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    d = json.loads(open(str(out2) + ".json").read())
    ostate = orc.state_from_group(orc.polygon(5), "p2", orc.MATH_CP)
    idx, best, _, _ = orc.analyse_state(ostate, orc.build_optimiser(steps=600, inner_steps=200), 0, 16, 4)
    got = [d["cell"]["length"], d["cell"]["ratio"], d["cell"]["angle"], d["occupied_sites"][0]["x"],
           d["occupied_sites"][0]["y"], d["occupied_sites"][0]["angle"]]
    assert got == list(best.dof)
    assert f"Final score: {best.score}" in r.stderr
    # --start-config: continue from the written state
    r = subprocess.run([sys.executable, "-m", "crystal_packing_b200.cli", "p2", "-s", "100", "--replications", "4",
                        "--start-config", str(out2) + ".json", "--outfile", str(tmp_path / "cont"), "polygon",
                        "--sides", "5"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    cont = json.loads(open(str(tmp_path / "cont") + ".json").read())
    assert cont["cell"]["length"] <= d["cell"]["length"]
