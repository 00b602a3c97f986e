"""SURVEY.md §8(f) rows 1 (as_positions) and 3 (SVG): host-side formatting of a state.
The text formats cannot be compared with the Rust binary here (it cannot be built); these tests pin
the numbers against the oracle's matrices (libm mode, what Rust executes) bit for bit, and the
structure against the reference's to_svg.rs."""
import ctypes as C
import math
import re
import xml.etree.ElementTree as ET

import pytest

import crystal_packing_b200 as cp
import orc
from common import GROUPS, make_states, random_dofs
from crystal_packing_b200 import render


@pytest.mark.parametrize("v,disp,dbg", [
    (8.0, "8", "8.0"), (0.1, "0.1", "0.1"), (1e-7, "0.0000001", "1e-7"), (1e16, "10000000000000000", "1e16"),
    (-0.25, "-0.25", "-0.25"), (1.5707963267948966, "1.5707963267948966", "1.5707963267948966"),
    (6.123233995736766e-17, "0.00000000000000006123233995736766", "6.123233995736766e-17"), (-0.0, "-0", "-0.0"),
])
def test_rust_float_formats(v, disp, dbg):  # `{}` and `{:?}` of an f64
    assert render.display_f64(v) == disp and float(disp) == v
    assert render.debug_f64(v) == dbg and float(dbg) == v


def test_f32_path_parameters():  # svg 0.10: path parameters are f32
    assert render.display_f32(0.1) == "0.1"
    assert render.display_f32(1.0 / 3.0) == "0.33333334"
    assert render.display_f32(-4.000000000000001) == "-4"
    assert render.display_f32(6.123233995736766e-17) == "0.000000000000000061232343" or \
        float(render.display_f32(6.123233995736766e-17)) == pytest.approx(6.1232e-17, rel=1e-4)


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("kind", [("polygon", 5), ("lj_trimer",)])
def test_positions_equal_the_oracle(kind, group):
    """relative_positions / cartesian_positions (site.rs:33-45, cell.rs:110-112) evaluated by
    render.py are the oracle's matrices (libm) to the last bit, on random states."""
    cstate, ostate = make_states(kind, group, orc.MATH_LIBM)
    rel = (orc.Transform * orc.MAX_SYMS)()
    cart = (orc.Transform * orc.MAX_SYMS)()
    s = orc.clone(ostate)
    for d in [orc.get_dof(ostate)] + random_dofs(ostate, 100, seed=5):
        orc.set_dof(s, d)
        cstate.dof = d
        n = orc.lib().orc_state_relative_positions(C.byref(s), rel)
        assert orc.lib().orc_state_cartesian_positions(C.byref(s), cart) == n == cstate.total_shapes()
        assert render.relative_positions(cstate) == [rel[k].rows() for k in range(n)]
        assert cstate.cartesian_positions() == [cart[k].rows() for k in range(n)]
        images = (orc.Transform * 9)()
        cell = render._Cell(d)
        for k in range(n):
            m = orc.lib().orc_cell_periodic_images(C.byref(s), C.byref(rel[k]), 1, 0, images)
            assert [images[q].rows() for q in range(m)] == list(cell.periodic_images(rel[k].rows(), 1, False))


def test_as_positions_text():
    """packed.rs:97-106: `{}` of the cell (cell.rs:67-77), `Positions`, `{:?}` of each transform."""
    state = cp.PackedState.from_group(cp.LineShape.polygon(4), "p2")
    text = state.as_positions()
    lines = text.splitlines()
    assert text.endswith("\n") and len(lines) == 2 + state.total_shapes()
    assert lines[0] == "Cell2 { a: 8, b: 8, angle: 1.5707963267948966 }"
    assert lines[1] == "Positions"
    # the template state: site (x, y, theta) = (0.25 a, 0.25 b, 0) wrapped, identity rotation, zero third row
    pos = state.cartesian_positions()
    pat = re.compile(r"^Transform2\(Transform \{ matrix: Matrix \{ data: ArrayStorage\(\[\[(.*)\], \[(.*)\], \[(.*)\]\]\) \}, "
                     r"_phantom: PhantomData \}\)$")
    for line, m in zip(lines[2:], pos):
        cols = [[float(v) for v in c.split(", ")] for c in pat.match(line).groups()]
        assert cols == [[m[r][c] for r in range(3)] for c in range(3)]  # column-major, exact round trip
        assert cols[0][2] == 0.0 and cols[2][2] == 0.0                   # transform.rs:139
    lj = cp.PotentialState.from_group(cp.LJShape2.circle(), "p2gg")
    assert len(lj.as_positions().splitlines()) == 6


@pytest.mark.parametrize("make,n_shape_nodes", [
    (lambda: cp.PackedState.from_group(cp.LineShape.polygon(5), "p2gg"), 1),
    (lambda: cp.PackedState.from_group(cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p2"), 3),
    (lambda: cp.PotentialState.from_group(cp.LJShape2.from_trimer(0.637556, 120.0, 1.0), "p1"), 3),
])
def test_svg_structure(make, n_shape_nodes):
    """to_svg.rs:127-275: defs (cell, mol), 9 cells, then per shape one blue use and 8 green ones."""
    state = make()
    state.dof = [4.0, 0.8, 1.3 if state.wallpaper.family == cp.capi.CP_MONOCLINIC else math.pi / 2, 0.1, -0.2, 0.7]
    text = state.as_svg()
    ns = {"s": "http://www.w3.org/2000/svg"}
    root = ET.fromstring(text)
    assert root.tag == "{http://www.w3.org/2000/svg}svg"
    defs = root.find("s:defs", ns)
    cell_g, mol_g = defs.findall("s:g", ns)
    assert cell_g.get("id") == "cell" and mol_g.get("id") == "mol"
    path = cell_g.find("s:path", ns)
    assert (path.get("fill"), path.get("stroke"), path.get("stroke-width")) == ("None", "grey", "0.1")
    assert re.fullmatch(r"M\S+,\S+ L\S+,\S+ L\S+,\S+ L\S+,\S+ z", path.get("d"))
    assert len(list(mol_g)) == n_shape_nodes
    uses = root.findall("s:use", ns)
    n = state.total_shapes()
    assert len(uses) == 9 + 9 * n
    assert [u.get("href") for u in uses[:9]] == ["#cell"] * 9
    fills = [u.get("fill") for u in uses[9:]]
    assert fills == (["blue"] + ["green"] * 8) * n
    # matrix(m00 m10 m01 m11 m02 m12) of the Cartesian transforms, exact round trip of the numbers
    cart = state.cartesian_positions()
    for k in range(n):
        vals = [float(v) for v in re.fullmatch(r"matrix\((.*)\)", uses[9 + 9 * k].get("transform")).group(1).split()]
        m = cart[k]
        assert vals == [m[0][0], m[1][0], m[0][1], m[1][1], m[0][2], m[1][2]]
    # view box: fold of (min, min, max, max) over 3 x corners with the enclosing radius as padding
    box = [float(v) for v in root.get("viewBox").split()]
    pad = state.shape.enclosing_radius()
    corners = render._Cell(state.dof).get_corners()
    assert box[0] == min(0.0, min(3.0 * x - pad for x, _ in corners))
    assert box[3] == max(0.0, max(2.0 * (3.0 * y + pad) for _, y in corners))
    # attributes sorted by name, as the svg crate writes them
    first_use = re.search(r"<use [^>]*>", text).group(0)
    assert first_use.index("href=") < first_use.index("transform=")


def test_svg_shapes():
    sq = cp.LineShape.polygon(4)
    d = render.shape_svg(sq).children[0].attrs["d"]
    assert d.startswith("M0,1 L1,") and d.endswith(" z") and d.count("L") == 4   # line_shape.rs:128-150 vertices
    lj = render.shape_svg(cp.LJShape2.circle()).children[0].attrs
    assert lj == {"r": "0.5", "cx": "0", "cy": "0"}                                # sigma / 2 (to_svg.rs:27-30)
    at = render.shape_svg(cp.MolecularShape2.circle()).children[0].attrs
    assert at["r"] == "1"
