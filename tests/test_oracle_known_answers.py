"""Pins the CPU oracle (oracle/cp_oracle.c) to every known-answer test the reference holds for the
Monte-Carlo hot path, to the upstream rand_pcg vectors and to the survey anchors.

Each test names the reference test it restates (paths relative to /root/reference).
"""
import ctypes as C
import math

import pytest

import orc

L = orc.lib()
PI = math.pi


def T_new(rot, tx, ty, mode=orc.MATH_LIBM):
    t = orc.Transform()
    L.orc_transform_new(mode, rot, tx, ty, C.byref(t))
    return t


def T_ops(s):
    t = orc.Transform()
    rc = L.orc_transform_from_operations(s.encode(), C.byref(t))
    return rc, t


def mul_point(t, x, y):
    ox, oy = C.c_double(), C.c_double()
    L.orc_transform_mul_point(C.byref(t), x, y, C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def item(*v):
    return orc.Item(*(tuple(v) + (0.0,) * (5 - len(v))))


def line_hits(a, b):
    return bool(L.orc_line_intersects(C.byref(item(*a)), C.byref(item(*b))))


def shape_hits(shape, t):
    kind, items = shape
    n = len(items)
    out = (orc.Item * n)()
    L.orc_shape_transform(kind, items, n, C.byref(t), out)
    return bool(L.orc_shape_intersects(kind, items, out, n))


# ------------------------------------------------------------------ RNG
def test_pcg64mcg_upstream_vectors():
    r = orc.Pcg()
    L.orc_pcg_new(C.byref(r), 42, 0)
    got = [L.orc_pcg_next_u64(C.byref(r)) for _ in range(6)]
    assert got == [0x63B4A3A813CE700A, 0x382954200617AB24, 0xA7FD85AE3FE950CE,
                   0xD715286AA2887737, 0x60C92FEE2E59F32C, 0x84C4E96BEFF30017]
    L.orc_pcg_from_seed(C.byref(r), bytes(range(1, 17)))
    assert L.orc_pcg_next_u64(C.byref(r)) == 7071994460355047496
    L.orc_pcg_seed_from_u64(C.byref(r), 0)
    assert L.orc_pcg_next_u64(C.byref(r)) == 6198063878555692194


def test_rand_distributions_survey_anchor():
    # seed_from_u64(0), D=6 -> first three (index, step, threshold) triples
    r = orc.Pcg()
    L.orc_pcg_seed_from_u64(C.byref(r), 0)
    got = [(L.orc_uniform_usize(C.byref(r), 6), L.orc_uniform_step(C.byref(r)),
            L.orc_gen_f64(C.byref(r))) for _ in range(3)]
    assert got == [(2, 0.33795735005140437, 0.352094216140564),
                   (4, -0.3487267040552302, 0.39516991633859344),
                   (0, -0.22602508532114252, 0.014925128909664798)]


def test_rand_ranges():
    r = orc.Pcg()
    L.orc_pcg_seed_from_u64(C.byref(r), 7)
    for _ in range(2000):
        assert 0 <= L.orc_uniform_usize(C.byref(r), 5) < 5
        assert -0.5 <= L.orc_uniform_step(C.byref(r)) < 0.5
        assert 0.0 <= L.orc_gen_f64(C.byref(r)) < 1.0


# ------------------------------------------------------------------ Transform2 (transform.rs tests)
def test_transform_new_zeros_is_identity():  # transform.rs:222-225
    ident = orc.Transform()
    L.orc_transform_identity(C.byref(ident))
    assert T_new(0.0, 0.0, 0.0).rows() == ident.rows()


def test_transform_point_exact():  # transform.rs:401-407
    assert mul_point(T_new(PI / 2, 1.0, 1.0), 0.2, 0.2) == (0.8, 1.2)


@pytest.mark.parametrize("ops,expect", [
    ("(x, y)", (0.1, 0.2)),           # transform.rs:409-415
    ("(-x, x+y)", (-0.1, 0.3)),       # :417-423
    ("(x+1/2, -y)", (0.6, -0.2)),     # :425-431
    ("(x-1/2, -y)", (-0.4, -0.2)),    # :433-439
    ("(-y, 0)", (-0.2, 0.0)),         # :441-447
])
def test_parse_operations(ops, expect):
    rc, t = T_ops(ops)
    assert rc == 0
    got = mul_point(t, 0.1, 0.2)
    assert got == pytest.approx(expect, abs=2.3e-16)
    assert t.rows()[2] == [0.0, 0.0, 0.0]  # Matrix3::zeros(): row 2 never written (transform.rs:139)


@pytest.mark.parametrize("ops", ["(z, y)", "(x, y, z)", "(x)"])  # transform.rs:449-468
def test_parse_operations_errors(ops):
    rc, _ = T_ops(ops)
    assert rc != 0


def test_transform_mul_matches_isometry_composition():  # transform.rs:383-398 (vs nalgebra Isometry2)
    import random
    rnd = random.Random(3)
    for _ in range(200):
        r1, r2 = rnd.uniform(-100, 100), rnd.uniform(-100, 100)
        t1, t2 = (rnd.uniform(-100, 100), rnd.uniform(-100, 100)), (rnd.uniform(-100, 100), rnd.uniform(-100, 100))
        a, b, out = T_new(r1, *t1), T_new(r2, *t2), orc.Transform()
        L.orc_transform_mul(C.byref(a), C.byref(b), C.byref(out))
        c, s = math.cos(r1 + r2), math.sin(r1 + r2)
        px = math.cos(r1) * t2[0] - math.sin(r1) * t2[1] + t1[0]
        py = math.sin(r1) * t2[0] + math.cos(r1) * t2[1] + t1[1]
        exp = [[c, -s, px], [s, c, py], [0, 0, 1]]
        for i in range(3):
            assert out.rows()[i] == pytest.approx(exp[i], abs=1e-10)
        assert out.rows()[2][2] == 1.0  # :335-341


def test_periodic_wrap():  # transform.rs:107-112
    for v, e in [(0.7, -0.3), (-0.7, 0.3), (0.5, -0.5), (-0.5, -0.5), (0.25, 0.25), (1.75, -0.25)]:
        out = orc.Transform()
        t = T_new(0.0, v, v)
        L.orc_transform_periodic(C.byref(t), 1.0, -0.5, C.byref(out))
        assert out.rows()[0][2] == pytest.approx(e, abs=1e-15)
        assert -0.5 <= out.rows()[0][2] < 0.5


# ------------------------------------------------------------------ Line2 / LineShape
def test_line_intersects_radial():  # line2.rs:118-139
    pts = [(-1.0, -1.0), (0.0, 0.0), (1.0, 1.0)]
    for a in pts:
        for b in pts:
            assert not line_hits(a + (0.0, 0.0), b + (0.0, 0.0))


def test_line_intersects():  # line2.rs:169-181
    l1, l2, l3 = (-1, 0, 0, -1), (-1, -1, 0, 0), (-2, -1, 1, 0)
    for a, b in [(l1, l2), (l2, l1), (l2, l3), (l3, l2), (l1, l3), (l3, l1)]:
        assert line_hits(tuple(map(float, a)), tuple(map(float, b)))


def test_line_transform():  # line2.rs:141-149
    ident = orc.Transform()
    L.orc_transform_identity(C.byref(ident))
    items = orc.items_array([(1.0, 1.0, 0.0, 0.0)])
    out = (orc.Item * 1)()
    L.orc_shape_transform(orc.POLYGON, items, 1, C.byref(ident), out)
    assert out[0].tup()[:4] == (1.0, 1.0, 0.0, 0.0)
    t = T_new(0.0, 1.0, 1.0)
    L.orc_shape_transform(orc.POLYGON, items, 1, C.byref(t), out)
    assert out[0].tup()[:4] == (2.0, 2.0, 1.0, 1.0)


def test_square_area_radius():  # line_shape.rs:169-180
    kind, sq = orc.polygon(4)
    assert L.orc_shape_area(kind, sq, 4) == pytest.approx(2.0, abs=2.3e-16)
    kind, sh = orc.polygon(4, [1.0, 2.0, 3.0, 4.0])
    assert L.orc_shape_enclosing_radius(kind, sh, 4) == pytest.approx(4.0, abs=2.3e-16)


@pytest.mark.parametrize("t,hit", [((1.0, 1.0), True), ((2.0, 2.0), False), ((0.0, 0.0), True),
                                   ((2.01, 2.01), False)])  # line_shape.rs:182-208
def test_square_intersections(t, hit):
    assert shape_hits(orc.polygon(4), T_new(0.0, *t)) is hit


def test_from_radial_too_few():  # line_shape.rs:129-131
    with pytest.raises(ValueError):
        orc.polygon(2)


# ------------------------------------------------------------------ Cell2 (cell.rs tests)
def default_cell(length=1.0, ratio=1.0, angle=PI / 2):
    s = orc.State()
    s.family, s.length, s.ratio, s.angle = orc.MONOCLINIC, length, ratio, angle
    s.math_mode = orc.MATH_LIBM
    return s


def cart(s, x, y):
    ox, oy = C.c_double(), C.c_double()
    L.orc_cell_to_cartesian(C.byref(s), x, y, C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def test_to_cartesian_doctest():  # cell.rs:249-256, :118-124
    s = default_cell(8.0)
    assert cart(s, 0.25, 0.25) == (2.0, 2.0)
    assert cart(s, 0.0, 0.0) == (0.0, 0.0)
    assert cart(s, 1.0, 1.0) == (8.0, 8.0)
    assert cart(s, 0.5, 0.5) == (4.0, 4.0)


def test_to_cartesian_angle():  # cell.rs:280-293
    s = default_cell()
    assert cart(s, 0.5, 0.5) == pytest.approx((0.5, 0.5), abs=2.3e-16)
    s.angle = PI / 4
    assert cart(s, 0.5, 0.5) == pytest.approx((0.5 + 0.5 / math.sqrt(2), 0.5 / math.sqrt(2)), abs=2.3e-16)


def images(s, t, shells, zero):
    out = (orc.Transform * 49)()
    n = L.orc_cell_periodic_images(C.byref(s), C.byref(t), shells, int(zero), out)
    return [out[i] for i in range(n)]


def test_periodic_images_order():  # cell.rs:334-373
    ident = orc.Transform()
    L.orc_transform_identity(C.byref(ident))
    s = default_cell()
    exp = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    got = [(t.rows()[0][2], t.rows()[1][2]) for t in images(s, ident, 1, False)]
    assert len(got) == 8
    for g, e in zip(got, exp):
        assert g == pytest.approx(e, abs=2.3e-16)
    got9 = images(s, ident, 1, True)
    assert len(got9) == 9
    assert (got9[4].rows()[0][2], got9[4].rows()[1][2]) == pytest.approx((0, 0), abs=2.3e-16)
    assert len(images(s, ident, 3, False)) == 48


@pytest.mark.parametrize("radius,cell,hit", [
    (1.0, (1.0, 1.0, PI / 2), True),       # cell.rs:295-306
    (0.5, (1.0, 1.0, PI / 2), True),       # :308-319
    (0.49, (1.0, 1.0, PI / 2), False),     # :321-332
    (1.0, (1.59, 0.83, 1.21), True),       # :375-392
])
def test_periodic_intersection(radius, cell, hit):
    shape = orc.polygon(4, [radius] * 4)
    s = default_cell(*cell)
    t = T_new(0.0, 0.0, 0.0)
    assert any(shape_hits(shape, im) for im in images(s, t, 1, False)) is hit


# ------------------------------------------------------------------ Atom2 / MolecularShape2
def atom_hits(a, b):
    return bool(L.orc_atom_intersects(C.byref(item(*a)), C.byref(item(*b))))


def test_atom_intersections():  # atom2.rs:66-98
    assert atom_hits((0, 0, 1), (0.5, 0, 1))
    a0, a1, a2, a3 = (0, 0, 1), (1, 0, 1), (0.5, 0.5, 1), (1.5, 1.5, 1)
    assert atom_hits(a0, a1) and atom_hits(a1, a2) and atom_hits(a3, a2)
    assert not atom_hits(a0, a3)
    eps = 2.220446049250313e-16
    r = math.sqrt(2) / 2
    b0, b1, b2 = (0, 0, r), (1, 1, r), (1, 1, r - 2 * eps)
    assert atom_hits(b0, b1) and atom_hits(b1, b2)
    assert not atom_hits(b0, b2)


def test_circle_overlap():  # molecular_shape2.rs:182-206
    a1 = item(0, 0, 1)
    assert L.orc_circle_overlap(C.byref(a1), C.byref(item(2.0, 0, 1))) == 0.0
    for i in range(10):
        d = (i + 1) / 10.0 * 2.0
        exp = 2.0 * (math.acos(d / 2) - d / 2 * math.sqrt(1 - (d / 2) ** 2))
        assert L.orc_circle_overlap(C.byref(a1), C.byref(item(d, 0, 1))) == pytest.approx(exp, abs=1e-7)


def test_trimer_geometry_and_area():  # molecular_shape2.rs:208-239
    kind, t = orc.trimer(orc.DISCS, 1.0, 180.0, 1.0)
    assert (t[0].a, t[0].b) == pytest.approx((0, 0), abs=2.3e-16)
    assert (t[1].a, t[1].b) == pytest.approx((-1, 0), abs=2.3e-16)
    assert (t[2].a, t[2].b) == pytest.approx((1, 0), abs=2.3e-16)
    kind, t = orc.trimer(orc.DISCS, 0.637556, 120.0, 1.0)
    assert (t[0].a, t[0].b) == pytest.approx((0, -1 / 3), abs=2.3e-16)
    assert (t[1].a, t[1].b) == pytest.approx((-0.866, 1 / 6), abs=1e-3)
    assert (t[2].a, t[2].b) == pytest.approx((0.866, 1 / 6), abs=1e-3)
    assert L.orc_shape_area(kind, t, 3) > 0
    kind, t = orc.trimer(orc.DISCS, 1.0, 180.0, 2.0)
    assert L.orc_shape_area(kind, t, 3) == pytest.approx(3 * PI, abs=2.3e-16)


@pytest.mark.parametrize("t,hit", [((1.0, 1.0), True), ((2.0, 0.0), False), ((0.0, 0.0), True),
                                   ((2.01, 0.0), False), ((2.0, 2.0), False)])  # molecular_shape2.rs:241-274
def test_circle_intersections(t, hit):
    assert shape_hits(orc.circle(orc.DISCS), T_new(0.0, *t)) is hit


# ------------------------------------------------------------------ LJ2 / LJShape2
def lj(x, y=0.0, sigma=1.0, eps=1.0, cutoff=float("nan")):
    return item(x, y, sigma, eps, cutoff)


def lj_e(a, b):
    return L.orc_lj_energy(C.byref(a), C.byref(b))


def test_lj_known_values():  # lj2.rs:101-150
    assert lj_e(lj(0), lj(1.0)) == 0.0
    assert lj_e(lj(0), lj(2 ** (1 / 6))) == pytest.approx(-1.0, abs=2.3e-16)
    for i in range(1, 100):
        assert lj_e(lj(0), lj(i / 100.0)) > 0
    for i in range(1, 500):
        e = lj_e(lj(0), lj(1 + i / 100.0))
        assert -1 < e < 0


def test_lj_cutoff():  # lj2.rs:152-185
    a = lj(0, cutoff=3.5)
    assert lj_e(a, lj(3.5, cutoff=3.5)) == 0.0
    for i in range(1, 250):
        e = lj_e(a, lj(1 + i / 100.0, cutoff=3.5))
        assert -1 < e < 0


def test_lj_trimer():  # lj_shape.rs:160-198
    kind, t = orc.trimer(orc.LJ, 2.0, 180.0, 0.5)
    assert [t[i].c for i in range(3)] == [2.0, 4.0, 4.0]
    assert [t[i].e for i in range(3)] == [3.5, 3.5, 3.5]
    kind, t = orc.trimer(orc.LJ, 0.637556, 120.0, 1.0)
    assert (t[0].a, t[0].b) == pytest.approx((0, -1 / 3), abs=2.3e-16)
    assert (t[1].a, t[1].b) == pytest.approx((-0.866, 1 / 6), abs=1e-3)


# ------------------------------------------------------------------ states
def custom_state(shape, family, ops, lj_state=False):
    """PackedState::initialise with a hand-built Wyckoff site (packed.rs:204-249)."""
    kind, items = shape
    s = orc.state_from_group(shape, "p1")
    s.family = family
    s.n_syms = len(ops)
    for i, o in enumerate(ops):
        rc = L.orc_transform_from_operations(o.encode(), C.byref(s.syms[i]))
        assert rc == 0
    r = L.orc_shape_enclosing_radius(kind, items, len(items))
    s.length = (2.0 if kind == orc.LJ else 4.0) * r * len(ops)
    pos = -0.5 + 0.5 / len(ops)
    s.x = s.y = pos
    return s


def test_packing_fraction_p1():  # packed.rs:257-267
    s = custom_state(orc.polygon(4), orc.MONOCLINIC, ["x,y"])
    assert L.orc_state_total_shapes(C.byref(s)) == 1
    assert orc.score(s) == pytest.approx(1 / 8, abs=2.3e-16)


def test_packing_fraction_p2mg_monoclinic():  # packed.rs:269-279
    s = custom_state(orc.polygon(4), orc.MONOCLINIC, ["x,y", "-x,-y", "-x+1/2,y", "x+1/2,-y"])
    assert L.orc_state_total_shapes(C.byref(s)) == 4
    assert orc.score(s) == pytest.approx(1 / 32, abs=2.3e-16)


def test_from_group_matches_custom():
    a = orc.state_from_group(orc.polygon(4), "p1")
    assert orc.score(a) == pytest.approx(1 / 8, abs=2.3e-16)
    assert L.orc_state_n_dof(C.byref(a)) == 6
    b = orc.state_from_group(orc.polygon(4), "p2mg")
    assert L.orc_state_n_dof(C.byref(b)) == 5
    assert orc.score(b) == pytest.approx(1 / 32, abs=2.3e-16)


def test_lj_total_shapes():  # potential.rs:231-241
    s = custom_state(orc.circle(orc.LJ), orc.MONOCLINIC, ["x,y"])
    assert L.orc_state_total_shapes(C.byref(s)) == 1
    s = custom_state(orc.circle(orc.LJ), orc.MONOCLINIC, ["x,y", "-x,-y", "-x+1/2,y", "x+1/2,-y"])
    assert L.orc_state_total_shapes(C.byref(s)) == 4


def test_group_table():  # wallpaper.rs:100-110,134-176
    fam, n = C.c_int32(), C.c_int32()
    syms = (orc.Transform * 8)()
    exp = {"p1": (0, 1), "p2": (0, 2), "p1m1": (1, 2), "p1g1": (1, 2), "pg": (1, 2), "p2mm": (1, 4),
           "p2mg": (1, 4), "p2gg": (1, 4)}
    for g, (f, m) in exp.items():
        assert L.orc_group(g.encode(), C.byref(fam), C.byref(n), syms) == 0
        assert (fam.value, n.value) == (f, m)
    assert L.orc_group(b"p2m1", C.byref(fam), C.byref(n), syms) != 0  # advertised, not parsable
    assert L.orc_group(b"p2gg", C.byref(fam), C.byref(n), syms) == 0
    assert syms[3].rows() == [[1.0, 0.0, 0.5], [0.0, -1.0, 0.5], [0.0, 0.0, 0.0]]


# ------------------------------------------------------------------ optimiser
def test_energy_surface():  # optimisation.rs:199-221
    import random
    rnd = random.Random(5)
    for mode in (orc.MATH_LIBM, orc.MATH_CP):
        for _ in range(500):
            new, old = rnd.uniform(-1e3, 1e3), rnd.uniform(-1e3, 1e3)
            r0 = L.orc_energy_surface(mode, new, old, 0.0)
            assert r0 == (0.0 if new < old else 1.0)
            r5 = L.orc_energy_surface(mode, new, old, 0.5)
            assert (0.0 <= r5 <= 1.0) if new < old else r5 == 1.0
        assert L.orc_energy_surface(mode, 1.0, 1.0, 0.0) == 1.0  # exp(NaN) -> min(NaN, 1) = 1


def test_build_optimiser():  # main.rs:122-143
    o = orc.Optimiser()
    b = orc.build_optimiser(steps=100, kt_start=0.1, kt_finish=0.001, inner_steps=1000)
    L.orc_build(C.byref(b), 7, C.byref(o))
    assert o.kt_ratio == math.pow(0.001 / 0.1, 1 / 100) and o.inner_steps == 100 and o.seed == 7
    b = orc.build_optimiser(kt_ratio=0.25)
    L.orc_build(C.byref(b), 0, C.byref(o))
    assert o.kt_ratio == 0.75
    b = orc.build_optimiser()
    L.orc_build(C.byref(b), 0, C.byref(o))
    assert o.kt_ratio == 0.1
    b = orc.build_optimiser(kt_start=0.0, kt_finish=0.001)
    L.orc_build(C.byref(b), 0, C.byref(o))
    assert o.kt_ratio == math.inf


def test_packing_improves():  # tests/packing.rs:14-52
    s = orc.state_from_group(orc.polygon(4), "p2")
    init = orc.score(s)
    rc, _, rej = orc.optimise(s, 0.0, 0.0, 0.001, 1000, 100, 0)
    assert rc == 0 and init < orc.score(s)
    # survey anchor (independent restatement)
    assert (init, orc.score(s), rej) == (0.0625, 0.6424637854786934, 525)


def test_potential_improves():  # tests/potential.rs:14-52
    s = orc.state_from_group(orc.trimer(orc.LJ, 0.63, 120.0, 1.0), "p2")
    init = orc.score(s)
    rc, _, rej = orc.optimise(s, 0.0, 0.0, 0.001, 1000, 100, 0)
    assert rc == 0 and init < orc.score(s)
    assert rej == 691
    assert orc.score(s) == pytest.approx(2.0102873836561734, rel=1e-12)


@pytest.mark.parametrize("sides,group,init,final,rej", [
    (5, "p1", 0.14860258067111773, 0.4938406351801332, 495),
    (5, "p2", 0.07430129033555886, 0.4907041612462589, 512),
    (5, "p2mg", 0.03715064516777943, 0.3557624755297206, 513),
    (5, "p2gg", 0.03715064516777943, 0.5661803539816451, 497),
    (12, "p2gg", 0.04687499999999999, 0.5848404578359193, 486),
])
def test_survey_anchors_polygons(sides, group, init, final, rej):
    s = orc.state_from_group(orc.polygon(sides), group)
    assert orc.score(s) == init
    rc, _, r = orc.optimise(s, 0.0, 0.0, 0.001, 1000, 100, 0)
    assert (rc, orc.score(s), r) == (0, final, rej)


def test_survey_anchor_trimer():
    s = orc.state_from_group(orc.trimer(orc.DISCS), "p2gg")
    assert orc.score(s) == pytest.approx(0.031084646689336408, rel=1e-14)
    rc, _, r = orc.optimise(s, 0.0, 0.0, 0.001, 1000, 100, 0)
    assert rc == 0 and r == 524
    assert orc.score(s) == pytest.approx(0.5730301134843927, rel=1e-14)
    c = orc.circle(orc.LJ)
    assert orc.score(orc.state_from_group(c, "p1")) == pytest.approx(2.352422182328277, rel=1e-14)


def test_invalid_initial_state():  # optimisation.rs:81-84 (panic -> rc 1)
    s = orc.state_from_group(orc.polygon(4), "p2")
    s.length = 0.5
    rc, _, _ = orc.optimise(s, 0.0, 0.0, 0.001, 100, 100, 0)
    assert rc == 1


def test_move_records_consistent():
    s = orc.state_from_group(orc.polygon(5), "p2")
    rc, recs, rej = orc.optimise(s, 0.1, 0.9, 0.01, 2000, 500, 3, record=True)
    assert rc == 0 and len(recs) == 2000
    assert sum(1 for r in recs if not r.accepted) == rej
    for r in recs:
        if not r.in_bounds:
            assert math.isnan(r.threshold) and not r.accepted
        if r.accepted:
            assert r.score_current == r.new_score


def test_three_stage_pipeline_and_max():  # main.rs:221-253
    s = orc.state_from_group(orc.polygon(4), "p2")
    b = orc.build_optimiser(steps=400, inner_steps=100)
    idx, best, results, _ = orc.analyse_state(s, b, 0, 6, n_threads=3)
    scores = [r.score for r in results]
    assert best.score == max(scores) and idx == scores.index(max(scores))
    assert all(r.moves == 100 + 400 + 400 for r in results)
    idx1, best1, results1, _ = orc.analyse_state(s, b, 0, 6, n_threads=1)
    assert [r.score for r in results1] == scores and idx1 == idx


def test_math_modes_agree_closely():
    # the deterministic cp_math functions (shared with the CUDA kernels) vs glibc
    import random
    rnd = random.Random(11)
    s1, c1, s2, c2 = (C.c_double() for _ in range(4))
    worst = 0.0
    for _ in range(20000):
        x = rnd.uniform(0, 2 * PI)
        L.orc_sincos(orc.MATH_LIBM, x, C.byref(s1), C.byref(c1))
        L.orc_sincos(orc.MATH_CP, x, C.byref(s2), C.byref(c2))
        for a, b in ((s1.value, s2.value), (c1.value, c2.value)):
            worst = max(worst, abs(a - b) / max(math.ulp(a), 5e-324))
    assert worst <= 1.0
    for _ in range(20000):
        x = -rnd.expovariate(0.2)
        a, b = L.orc_exp(orc.MATH_LIBM, x), L.orc_exp(orc.MATH_CP, x)
        assert abs(a - b) <= math.ulp(a)
    assert L.orc_exp(orc.MATH_CP, float("-inf")) == 0.0
    assert L.orc_exp(orc.MATH_CP, float("inf")) == math.inf
    assert math.isnan(L.orc_exp(orc.MATH_CP, float("nan")))


def test_golden_fixture_matches_oracle():
    """tests/golden/trajectories.json is what tests/golden/make_golden.py writes today: the oracle
    in libm mode reproduces every committed final state bit for bit."""
    import json
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden

    doc = json.load(open(os.path.join(here, "golden", "trajectories.json")))
    assert len(doc["cases"]) == len(make_golden.CASES)
    for case in doc["cases"]:
        state = orc.state_from_group(make_golden.oracle_shape(case["spec"]), case["group"], orc.MATH_LIBM)
        bopt = orc.build_optimiser(**case["optimiser"])
        _, _, results, _ = orc.analyse_state(state, bopt, 0, len(case["replicas"]), n_threads=4)
        for got, exp in zip(results, case["replicas"]):
            assert [v.hex() for v in got.dof] == exp["dof"] and got.score.hex() == exp["score"]
            assert (got.rejections, got.moves) == (exp["rejections"], exp["moves"])
