"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, fails loudly without a GPU, and its host-side constructors (shapes, plane groups,
initial states, schedule derivation) equal the oracle's bit for bit."""
import ctypes as C
import math
import os
import re

import pytest

import crystal_packing_b200 as cp
import orc
from common import GROUPS, make_shapes, make_states
from crystal_packing_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cp_gpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cp_[a-z_0-9]+)\s*\(", header))
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    lib = cp.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.cp_version()


def test_struct_layouts_match_header():
    assert C.sizeof(capi.ReplicaResult) == 72 and C.sizeof(capi.Best) == 80
    assert C.sizeof(capi.Schedule) == 48 and C.sizeof(capi.Move) == 32 and C.sizeof(capi.Decision) == 24
    assert C.sizeof(capi.Problem) == 8 + 16 * 40 + 16 + 8 + 4 * 48 + 48


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cp.CrystalPackingError) as e:
        cp.Engine(0)
    assert e.value.status in (capi.CP_ERR_NO_DEVICE, capi.CP_ERR_CUDA)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "crystal-packing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "cp_math.h":
                text = open(os.path.join(dirpath, f)).read()
                assert "cp_oracle" not in text and "import orc" not in text, f


SHAPES = [("polygon", 3), ("polygon", 4), ("polygon", 5), ("polygon", 12), ("radial", (1.0, 2.0, 3.0, 4.0)),
          ("trimer",), ("circle",), ("lj_trimer",), ("lj_trimer", 0.63), ("lj_circle",)]


@pytest.mark.parametrize("spec", SHAPES)
def test_shape_constructors_match_oracle(spec):
    cshape, (okind, oitems) = make_shapes(spec)
    assert cshape.kind == okind and len(cshape.items) == len(oitems)
    width = {orc.POLYGON: 4, orc.DISCS: 3, orc.LJ: 5}[okind]
    for mine, ref in zip(cshape.items, oitems):
        for a, b in zip(mine[:width], ref.tup()[:width]):
            assert (math.isnan(a) and math.isnan(b)) or a == b
    L = orc.lib()
    assert cshape.enclosing_radius() == L.orc_shape_enclosing_radius(okind, oitems, len(oitems))
    assert cshape.area() == L.orc_shape_area(okind, oitems, len(oitems))


def test_shape_errors():
    with pytest.raises(cp.CrystalPackingError):  # line_shape.rs:129-131
        cp.LineShape.from_radial("x", [1.0, 1.0])
    with pytest.raises(cp.CrystalPackingError):
        cp.LineShape.polygon(capi.MAX_ITEMS + 1)


@pytest.mark.parametrize("group", GROUPS + ["pg"])
def test_groups_and_initial_state_match_oracle(group):
    for spec in [("polygon", 5), ("trimer",), ("lj_trimer",)]:
        cstate, ostate = make_states(spec, group)
        assert cstate.dof == orc.get_dof(ostate)
        assert cstate.total_shapes() == ostate.n_syms
        assert cstate._problem.family == ostate.family
        for i in range(ostate.n_syms):
            rows = ostate.syms[i].rows()
            assert list(cstate._problem.syms[i]) == rows[0] + rows[1]
            assert rows[2] == [0.0, 0.0, 0.0]
        lo, hi = (C.c_double * 6)(), (C.c_double * 6)()
        n = orc.lib().orc_state_bounds(C.byref(ostate), lo, hi)
        basis = cstate.generate_basis()
        assert len(basis) == n
        assert [(b[1], b[2]) for b in basis] == [(lo[i], hi[i]) for i in range(n)]


def test_group_parsing_quirks():  # wallpaper.rs:100-110,128-132
    with pytest.raises(cp.CrystalPackingError):
        cp.WallpaperGroups.from_str("p2m1")
    assert "p2m1" in cp.WallpaperGroups.variants() and "p1m1" not in cp.WallpaperGroups.variants()
    assert cp.WallpaperGroups.from_str("pg").name == "p1m1"
    assert cp.WallpaperGroups.from_str("p1g1").name == "p1m1"


@pytest.mark.parametrize("ops", ["(x, y)", "(-x, x+y)", "(x+1/2, -y)", "(x-1/2, -y)", "(-y, 0)", "x*2,y"])
def test_parse_operations_match_oracle(ops):  # transform.rs:409-447
    t = cp.Transform2.from_operations(ops)
    o = orc.Transform()
    assert orc.lib().orc_transform_from_operations(ops.encode(), C.byref(o)) == 0
    assert t.rows == o.rows()


@pytest.mark.parametrize("ops", ["(z, y)", "(x, y, z)", "(x)", ""])
def test_parse_operations_errors(ops):  # transform.rs:449-468
    with pytest.raises(cp.CrystalPackingError):
        cp.Transform2.from_operations(ops)


@pytest.mark.parametrize("kw", [
    dict(steps=100, kt_start=0.1, kt_finish=0.001, inner_steps=1000),
    dict(steps=10000, kt_start=0.1, kt_finish=None, kt_ratio=0.25, inner_steps=1000),
    dict(steps=5000, kt_start=0.1, kt_finish=None, kt_ratio=None, inner_steps=300, convergence=1e-6),
    dict(steps=1000, kt_start=0.0, kt_finish=0.001, inner_steps=100),
])
def test_schedules_match_oracle(kw):  # main.rs:122-143, 224-252
    full = dict(kt_finish=None, kt_ratio=None, max_step_size=0.01, convergence=None)
    full.update(kw)
    b = cp.BuildOptimiser(**full)
    ob = orc.build_optimiser(**full)
    stages = b.pipeline()
    variants = [dict(steps=100, kt_start=0.0, convergence=None), {}, dict(kt_start=0.0)]
    for st, v in zip(stages, variants):
        kw2 = dict(full)
        kw2.update(v)
        o = orc.Optimiser()
        orc.lib().orc_build(C.byref(orc.build_optimiser(**kw2)), 0, C.byref(o))
        for f in ("kt_start", "kt_ratio", "max_step_size", "convergence"):
            a, e = getattr(st, f), getattr(o, f)
            assert (math.isnan(a) and math.isnan(e)) or a == e, f
        assert (st.steps, st.inner_steps) == (o.steps, o.inner_steps)
    b.seed = 5
    m = b.build()
    assert m.seed == 5 and m.schedule.kt_ratio == stages[1].kt_ratio


def test_cuda_sources_target_sm100a_only():
    mk = open(os.path.join(ROOT, "crystal-packing_b200", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk and "-fmad=false" in mk and "-lineinfo" in mk
