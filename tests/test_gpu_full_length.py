"""GPU parity at the schedule bench.py times, and of the pieces a score cannot see.

* whole three-stage pipelines of BASELINE.json's length (100 + 10,000 + 10,000 steps, inner_steps
  1,000) with the reference's PCG stream on hundreds of replicas: final DOFs, score, rejection and
  move counts bit-identical to the oracle, oracle on the shared sin/cos/exp and on glibc's;
* replay of a dump taken from a LATE inner loop of stage 2 and of stage 3 (dense cells, two or three
  image shells, every proposal close to a contact), from the state it was taken at;
* the symmetry images themselves (wrapped fractional and Cartesian positions, rotation blocks) bit
  for bit against OccupiedSite::positions / Cell2::to_cartesian_isometry of the oracle.

Oracle = CPU restatement of the reference (oracle/cp_oracle.c); see its header for what is pinned.
"""
import ctypes as C
import math
import os
import random

import pytest

import crystal_packing_b200 as cp
import orc
from common import GROUPS, c_build, make_states, random_dofs, same_float

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = cp.Engine(0)
    yield e
    e.close()


FULL_CASES = [
    (("polygon", 5), "p2", {}, 320),
    (("polygon", 5), "p2gg", {}, 256),
    (("trimer",), "p2gg", dict(kt_start=0.1, kt_finish=0.001), 256),
]


@pytest.mark.parametrize("math_mode", [orc.MATH_CP, orc.MATH_LIBM])
@pytest.mark.parametrize("spec,group,kw,n", FULL_CASES)
def test_bench_schedule_bit_exact(eng, spec, group, kw, n, math_mode):
    """The schedule bench.py times, PCG stream, default kernels.  With the oracle on glibc's libm
    (what the Rust reference executes) the DOF trajectories must still be identical: sin/cos/exp
    only enter the score, never the DOF update, so only a knife-edge decision could differ."""
    bopt = orc.build_optimiser(steps=10000, inner_steps=1000, **kw)
    cstate, ostate = make_states(spec, group, math_mode)
    best, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=eng, return_all=True)
    idx, obest, ores, _ = orc.analyse_state(ostate, bopt, 0, n, n_threads=os.cpu_count() or 1)
    for k in range(n):
        assert list(res[k].dof) == list(ores[k].dof), k
        assert (res[k].rejections, res[k].moves) == (ores[k].rejections, ores[k].moves), k
        if math_mode == orc.MATH_CP:
            assert same_float(res[k].score, ores[k].score), k
        else:
            assert res[k].score == pytest.approx(ores[k].score, rel=1e-14), k
    if math_mode == orc.MATH_CP:
        assert best.best_replica == idx and best.best_score == obest.score


def test_bench_schedule_lj_bit_exact(eng):
    """LJ trimer in p2gg (BASELINE.json configs[3]) on a 100 + 1,000 + 1,000 schedule."""
    bopt = orc.build_optimiser(steps=1000, inner_steps=1000)
    cstate, ostate = make_states(("lj_trimer",), "p2gg")
    n = 128
    best, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=eng, return_all=True)
    idx, obest, ores, _ = orc.analyse_state(ostate, bopt, 0, n, n_threads=os.cpu_count() or 1)
    for k in range(n):
        assert list(res[k].dof) == list(ores[k].dof) and res[k].score == ores[k].score, k
        assert res[k].rejections == ores[k].rejections
    assert best.best_replica == idx


def _late_dump(spec, group, seed, stage, skip, take):
    """Runs stages 1..stage of the pipeline on the oracle for one replica, recording the last one;
    returns (problem state, DOFs when the stage began, DOFs after `skip` moves, the next `take` records)."""
    cstate, ostate = make_states(spec, group)
    quick = orc.build_optimiser(steps=100, kt_start=0.0, inner_steps=100)
    o = orc.Optimiser()
    orc.lib().orc_build(C.byref(quick), seed, C.byref(o))
    rc, _, _ = orc.optimise(ostate, o.kt_start, o.kt_ratio, o.max_step_size, o.steps, o.inner_steps, seed)
    assert rc == 0
    main = orc.build_optimiser(steps=10000, inner_steps=1000)
    orc.lib().orc_build(C.byref(main), seed, C.byref(o))
    if stage == 3:
        rc, _, _ = orc.optimise(ostate, o.kt_start, o.kt_ratio, o.max_step_size, o.steps, o.inner_steps, seed)
        assert rc == 0
        o.kt_start = 0.0
    at_stage_start = orc.get_dof(ostate)
    rc, recs, _ = orc.optimise(ostate, o.kt_start, o.kt_ratio, o.max_step_size, o.steps, o.inner_steps, seed,
                               record=True)
    assert rc == 0 and len(recs) == 10000
    n_cell = 3 if ostate.family == orc.MONOCLINIC else 2
    dof = list(at_stage_start)
    for r in recs[:skip]:
        if r.accepted:
            slot = r.basis_index if r.basis_index < n_cell else r.basis_index - n_cell + 3
            dof[slot] = r.new_value
    return cstate, at_stage_start, dof, recs[skip:skip + take]


@pytest.mark.parametrize("stage", [2, 3])
@pytest.mark.parametrize("spec,group", [(("polygon", 5), "p2"), (("polygon", 5), "p2gg"), (("trimer",), "p2gg")])
def test_replay_late_dense_loop(eng, spec, group, stage):
    """Decisions and Option<score> of moves 8,000..9,500 of stage 2 / stage 3 (step_ratio >> 1, the
    cell nearly closed), replayed from the state the dump starts at with the stage's bounds."""
    inits, bounds, seqs, dumps = [], [], [], []
    for seed in range(4):
        cstate, start, dof, recs = _late_dump(spec, group, seed, stage, 8000, 1500)
        inits.append(dof)
        bounds.append(start)
        seqs.append([(r.basis_index, r.new_value, 0.0 if math.isnan(r.threshold) else r.threshold, r.kt)
                     for r in recs])
        dumps.append(recs)
    assert sum(r.in_bounds for recs in dumps for r in recs) > 1000
    out = eng.replay(cstate._problem, inits, seqs, bound_dofs=bounds)
    for recs, decs in zip(dumps, out):
        for r, d in zip(recs, decs):
            assert (d.in_bounds, d.accepted) == (r.in_bounds, r.accepted)
            assert same_float(d.new_score, r.new_score) and d.score_current == r.score_current


@pytest.mark.parametrize("group", GROUPS)
def test_symmetry_images_bit_exact(eng, group):
    """OccupiedSite::positions (site.rs:33-45) and Cell2::to_cartesian_isometry (cell.rs:110-112):
    wrapped fractional position, Cartesian position and rotation block of every symmetry image."""
    cstate, ostate = make_states(("polygon", 5), group)
    rnd = random.Random(3)
    dofs = [orc.get_dof(ostate)] + random_dofs(ostate, 300, seed=11)
    # positions on and next to the wrap boundaries, angles at the ends of their ranges
    for x, y, th in ((-0.5, 0.5, 0.0), (0.5, -0.5, math.tau), (0.0, 0.0, math.pi), (0.4999999999999999, -0.25, 1e-300)):
        d = list(dofs[rnd.randrange(1, 300)])
        d[3], d[4], d[5] = x, y, th
        dofs.append(d)
    got = eng.images(cstate._problem, dofs)
    s = orc.clone(ostate)
    rel = (orc.Transform * orc.MAX_SYMS)()
    cart = (orc.Transform * orc.MAX_SYMS)()
    for d, images in zip(dofs, got):
        orc.set_dof(s, d)
        n = orc.lib().orc_state_relative_positions(C.byref(s), rel)
        assert orc.lib().orc_state_cartesian_positions(C.byref(s), cart) == n == len(images)
        for k, (fx, fy, cx, cy, m00, m01, m10, m11) in enumerate(images):
            r, c = rel[k].rows(), cart[k].rows()
            assert (fx, fy) == (r[0][2], r[1][2]), (group, d, k)
            assert (cx, cy) == (c[0][2], c[1][2]), (group, d, k)
            # the rotation block is the same in both transforms (cell.rs:110-112 keeps it)
            assert (m00, m01, m10, m11) == (r[0][0], r[0][1], r[1][0], r[1][1]) == (c[0][0], c[0][1], c[1][0], c[1][1])
