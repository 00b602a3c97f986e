"""Shared helpers for the parity tests: the same configuration built on both sides — the product
(crystal_packing_b200, through the C ABI) and the CPU oracle (orc)."""
import math
import random

import crystal_packing_b200 as cp
import orc

GROUPS = ["p1", "p2", "p1m1", "p1g1", "p2mm", "p2mg", "p2gg"]


def make_shapes(spec):
    """spec: ("polygon", sides) | ("trimer",) | ("circle",) | ("lj_trimer", radius) | ("lj_circle",)"""
    kind = spec[0]
    if kind == "polygon":
        return cp.LineShape.polygon(spec[1]), orc.polygon(spec[1])
    if kind == "radial":
        return cp.LineShape.from_radial("r", list(spec[1])), orc.polygon(len(spec[1]), list(spec[1]))
    if kind == "trimer":
        return cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), orc.trimer(orc.DISCS)
    if kind == "circle":
        return cp.MolecularShape2.circle(), orc.circle(orc.DISCS)
    if kind == "lj_trimer":
        r = spec[1] if len(spec) > 1 else 0.637556
        return cp.LJShape2.from_trimer(r, 120.0, 1.0), orc.trimer(orc.LJ, r)
    if kind == "lj_circle":
        return cp.LJShape2.circle(), orc.circle(orc.LJ)
    if kind == "discs":  # hand-built molecule: ((x, y, radius), ...) -> the run-time-count kernels
        atoms = spec[1]
        shape = cp.MolecularShape2.circle()
        shape._problem.n_items = len(atoms)
        for k, a in enumerate(atoms):
            for c, v in enumerate(a):
                shape._problem.items[k][c] = v
        cp.api.refresh_shape(shape)
        return shape, (orc.DISCS, orc.items_array(atoms))
    if kind == "lj_sites":  # ((x, y, sigma, epsilon, cutoff), ...)
        sites = spec[1]
        shape = cp.LJShape2.circle()
        shape._problem.n_items = len(sites)
        for k, a in enumerate(sites):
            for c, v in enumerate(a):
                shape._problem.items[k][c] = v
        cp.api.refresh_shape(shape)
        return shape, (orc.LJ, orc.items_array(sites))
    raise ValueError(spec)


def make_states(spec, group, math_mode=orc.MATH_CP):
    cshape, oshape = make_shapes(spec)
    cls = cp.PotentialState if spec[0].startswith("lj") else cp.PackedState
    cstate, ostate = cls.from_group(cshape, group), orc.state_from_group(oshape, group, math_mode)
    assert cstate.dof == orc.get_dof(ostate)  # same enclosing radius -> same initial cell
    return cstate, ostate


def random_dofs(ostate, n, seed, spread=1.0):
    """Random states around a moderately dense cell so that both overlapping and valid states occur."""
    rnd = random.Random(seed)
    base = orc.get_dof(ostate)
    out = []
    for _ in range(n):
        scale = rnd.uniform(0.18, 1.0) if rnd.random() < 0.8 else rnd.uniform(0.05, 0.2)
        out.append([
            base[0] * scale,
            rnd.uniform(0.25, 1.0) if rnd.random() < 0.7 else rnd.uniform(0.1, 0.3),
            rnd.uniform(math.pi / 4, math.pi / 2) if ostate.family == orc.MONOCLINIC else math.pi / 2,
            rnd.uniform(-0.5, 0.5), rnd.uniform(-0.5, 0.5), rnd.uniform(0, math.tau),
        ])
    return out


def oracle_scores(ostate, dofs):
    s = orc.clone(ostate)
    res = []
    for d in dofs:
        orc.set_dof(s, d)
        res.append(orc.score(s))
    return res


def same_float(a, b):
    """bit-equality with NaN == NaN (None == None)"""
    return (math.isnan(a) and math.isnan(b)) or a == b


def c_build(bopt):
    """orc.BuildOptimiser -> cp.BuildOptimiser with the same fields"""
    f = lambda v: None if math.isnan(v) else v
    return cp.BuildOptimiser(steps=bopt.steps, kt_start=bopt.kt_start, kt_finish=f(bopt.kt_finish),
                             kt_ratio=f(bopt.kt_ratio), max_step_size=bopt.max_step_size,
                             inner_steps=bopt.inner_steps, convergence=f(bopt.convergence))
