"""GPU parity tests (B200): every call goes through the C ABI of lib/libcpgpu.so and is compared
with the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): overlap predicate, symmetry images and accept/reject decisions
bit-exact; packing fraction and LJ energy within 1e-6 relative (asserted much tighter where the
arithmetic is identical); best-packing distributions statistically equivalent for the Philox stream.
"""
import ctypes as C
import json
import math
import os

import pytest

import crystal_packing_b200 as cp
import orc
from common import GROUPS, c_build, make_states, oracle_scores, random_dofs, same_float

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LJ_RTOL = 1e-6  # north_star tolerance for energies; asserted bit-exact below because the kernels keep
                # the reference's accumulation order


@pytest.fixture(scope="module")
def eng():
    e = cp.Engine(0)
    yield e
    e.close()


DIMER = ("discs", ((-0.6, 0.0, 0.8), (0.6, 0.1, 0.7)))
TETRAMER = ("discs", ((0.0, 0.0, 1.0), (1.1, 0.0, 0.5), (-1.1, 0.0, 0.5), (0.0, 1.2, 0.4)))
LJ_DIMER = ("lj_sites", ((-0.5, 0.0, 1.0, 1.0, 3.0), (0.5, 0.0, 0.8, 0.7, 2.5)))
LJ_PAIR_NOCUT = ("lj_sites", ((-0.4, 0.0, 1.0, 1.0, float("nan")), (0.4, 0.1, 0.9, 1.2, float("nan"))))
# every build of the kernels: specialised component counts and the run-time-count ones (13-, 16-gons,
# 2- and 4-disc molecules)
HARD_SPECS = [("polygon", 3), ("polygon", 4), ("polygon", 5), ("polygon", 8), ("polygon", 12),
              ("polygon", 13), ("polygon", 16), ("radial", (1.0, 0.6, 1.0, 0.6, 1.0, 0.6)), ("trimer",), ("circle",),
              DIMER, TETRAMER]


# ------------------------------------------------------------------ State::score (overlap predicate)
@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("spec", HARD_SPECS)
def test_score_bit_exact_hard(eng, spec, group):
    cstate, ostate = make_states(spec, group)
    dofs = [orc.get_dof(ostate)] + random_dofs(ostate, 400, seed=hash((str(spec), group)) & 0xFFFF)
    ref = oracle_scores(ostate, dofs)
    assert any(math.isnan(r) for r in ref) and any(not math.isnan(r) for r in ref)
    for lanes in (0, 1, 4, 8, 16, 32):
        eng.set_lanes(lanes)
        got = eng.score(cstate._problem, dofs)
        bad = [i for i, (a, b) in enumerate(zip(got, ref)) if not same_float(a, b)]
        assert not bad, (lanes, bad[:5], [(got[i], ref[i], dofs[i]) for i in bad[:2]])
    eng.set_lanes(0)


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("spec", [("lj_trimer",), ("lj_trimer", 0.63), ("lj_circle",), LJ_DIMER, LJ_PAIR_NOCUT])
def test_score_lj_bit_exact(eng, spec, group):
    cstate, ostate = make_states(spec, group)
    dofs = [orc.get_dof(ostate)] + random_dofs(ostate, 200, seed=17)
    ref = oracle_scores(ostate, dofs)
    for lanes in (0, 1, 4, 8, 16, 32):
        eng.set_lanes(lanes)
        got = eng.score(cstate._problem, dofs)
        for a, b in zip(got, ref):
            assert a == pytest.approx(b, rel=LJ_RTOL, abs=1e-12)  # the contract
            assert same_float(a, b)                                # what the kernels deliver
    eng.set_lanes(0)


def test_reference_known_answers_on_gpu(eng):
    # packed.rs:263-267 (square p1 = 1/8), :275-279 analogue through from_group (p2mg = 1/32)
    sq = cp.LineShape.from_radial("Square", [1.0] * 4)
    assert cp.PackedState.from_group(sq, "p1").score(eng) == pytest.approx(1 / 8, abs=2.3e-16)
    assert cp.PackedState.from_group(sq, "p2mg").score(eng) == pytest.approx(1 / 32, abs=2.3e-16)
    assert cp.PackedState.from_group(sq, "p1").total_shapes() == 1
    # cell.rs:295-332: unit cell, square of radius 1 / 0.5 overlaps its images, 0.49 does not
    for r, overlaps in ((1.0, True), (0.5, True), (0.49, False)):
        s = cp.PackedState.from_group(cp.LineShape.from_radial("Square", [r] * 4), "p1")
        s.dof = [1.0, 1.0, math.pi / 2, 0.0, 0.0, 0.0]
        assert (s.score(eng) is None) is overlaps
    # cell.rs:375-392: skewed cell (1.59, 0.83, 1.21) overlaps
    s = cp.PackedState.from_group(sq, "p1")
    s.dof = [1.59, 0.83, 1.21, 0.0, 0.0, 0.0]
    assert s.score(eng) is None
    # molecular_shape2.rs:241-274: touching discs do not intersect (strict <)
    c = cp.PackedState.from_group(cp.MolecularShape2.circle(), "p1")
    c.dof = [2.0, 1.0, math.pi / 2, 0.0, 0.0, 0.0]
    assert c.score(eng) == pytest.approx(math.pi / 4, rel=1e-15)
    c.dof = [1.99, 1.0, math.pi / 2, 0.0, 0.0, 0.0]
    assert c.score(eng) is None
    # survey anchor: LJ circle in p1 initial score
    lj = cp.PotentialState.from_group(cp.LJShape2.circle(), "p1")
    assert lj.score(eng) == pytest.approx(2.352422182328277, rel=1e-12)


# ------------------------------------------------------------------ replayed move sequences
def dump_moves(spec, group, math_mode, seed, steps=600, inner=200, kt_start=0.1, kt_ratio=0.5):
    cstate, ostate = make_states(spec, group, math_mode)
    init = orc.get_dof(ostate)
    rc, recs, _ = orc.optimise(ostate, kt_start, kt_ratio, 0.01, steps, inner, seed, record=True)
    assert rc == 0
    return cstate, init, recs


@pytest.mark.parametrize("math_mode", [orc.MATH_CP, orc.MATH_LIBM])
@pytest.mark.parametrize("spec,group", [(("polygon", 4), "p2"), (("polygon", 5), "p1"),
                                        (("polygon", 5), "p2"), (("polygon", 5), "p2mg"),
                                        (("polygon", 5), "p2gg"), (("polygon", 7), "p1g1"),
                                        (("trimer",), "p2gg"), (("circle",), "p2mm")])
def test_replay_decisions_bit_exact(eng, spec, group, math_mode):
    """Accept/reject decisions and Option<score> replayed from oracle-dumped move sequences.  With
    the oracle on glibc's sin/cos/exp (what the Rust reference executes) the decisions must still
    match; scores may then differ in the last place (different libm), hence rel 1e-14 there."""
    seqs, inits, dumps = [], [], []
    for seed in range(4):
        cstate, init, recs = dump_moves(spec, group, math_mode, seed)
        seqs.append([(r.basis_index, r.new_value, 0.0 if math.isnan(r.threshold) else r.threshold, r.kt)
                     for r in recs])
        inits.append(init)
        dumps.append(recs)
    out = eng.replay(cstate._problem, inits, seqs)
    for recs, decs in zip(dumps, out):
        for r, d in zip(recs, decs):
            assert (d.in_bounds, d.accepted) == (r.in_bounds, r.accepted)
            assert math.isnan(d.new_score) == math.isnan(r.new_score)  # the overlap predicate
            if math_mode == orc.MATH_CP:
                assert same_float(d.new_score, r.new_score) and d.score_current == r.score_current
            else:
                assert d.score_current == pytest.approx(r.score_current, rel=1e-14)


def test_replay_lj_decisions(eng):
    seqs, inits, dumps = [], [], []
    for seed in range(3):
        cstate, init, recs = dump_moves(("lj_trimer", 0.63), "p2", orc.MATH_CP, seed, steps=300, inner=100)
        seqs.append([(r.basis_index, r.new_value, 0.0 if math.isnan(r.threshold) else r.threshold, r.kt)
                     for r in recs])
        inits.append(init)
        dumps.append(recs)
    out = eng.replay(cstate._problem, inits, seqs)
    for recs, decs in zip(dumps, out):
        for r, d in zip(recs, decs):
            assert (d.in_bounds, d.accepted) == (r.in_bounds, r.accepted)
            assert same_float(d.new_score, r.new_score) and d.score_current == r.score_current


# ------------------------------------------------------------------ free-running trajectories (PCG)
def run_both(eng, spec, group, bopt, n, begin=0, math_mode=orc.MATH_CP, lanes=0):
    cstate, ostate = make_states(spec, group, math_mode)
    eng.set_lanes(lanes)
    best, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=eng,
                                            replica_begin=begin, return_all=True)
    eng.set_lanes(0)
    idx, obest, ores, _ = orc.analyse_state(ostate, bopt, begin, n, n_threads=os.cpu_count() or 1)
    return best, res, idx, obest, ores


TRAJ_CASES = [
    # BASELINE.json configs[0]: square in p2
    (("polygon", 4), "p2", dict(steps=2000, inner_steps=500)),
    # configs[1]: pentagon across p1/p2/p2mg/p2gg
    (("polygon", 5), "p1", dict(steps=2000, inner_steps=500)),
    (("polygon", 5), "p2", dict(steps=2000, inner_steps=500)),
    (("polygon", 5), "p2mg", dict(steps=1500, inner_steps=500)),
    (("polygon", 5), "p2gg", dict(steps=1500, inner_steps=500)),
    # configs[2]: trimer in p2gg with the annealing schedule kt 0.1 -> 0.001
    (("trimer",), "p2gg", dict(steps=1500, inner_steps=500, kt_start=0.1, kt_finish=0.001)),
    # configs[4] corners: 3- and 12-gons, remaining groups
    (("polygon", 3), "p1m1", dict(steps=1000, inner_steps=250)),
    (("polygon", 12), "p2gg", dict(steps=600, inner_steps=200)),
    (("polygon", 6), "p2mm", dict(steps=1000, inner_steps=250)),
    (("polygon", 9), "p1g1", dict(steps=1000, inner_steps=250)),
    (("circle",), "p2", dict(steps=1000, inner_steps=250, kt_ratio=0.2)),
    # quirks: steps < inner_steps clamps; NaN temperature in stage 3
    (("polygon", 5), "p2", dict(steps=100, inner_steps=1000)),
    (("polygon", 4), "p2", dict(steps=900, inner_steps=300, kt_start=0.1, kt_finish=0.01)),
    # convergence early exit (optimisation.rs:142-159)
    (("polygon", 5), "p2", dict(steps=6000, inner_steps=200, convergence=1e-3)),
    # steps not a multiple of inner_steps: the remainder is dropped (optimisation.rs:98)
    (("polygon", 4), "p1", dict(steps=1150, inner_steps=300)),
    # the run-time-count builds of the kernels
    (("polygon", 13), "p2", dict(steps=600, inner_steps=200)),
    (("polygon", 16), "p2gg", dict(steps=300, inner_steps=100)),
    (DIMER, "p2mg", dict(steps=900, inner_steps=300)),
    (TETRAMER, "p2", dict(steps=900, inner_steps=300)),
]


@pytest.mark.parametrize("spec,group,kw", TRAJ_CASES)
def test_trajectories_bit_exact(eng, spec, group, kw):
    """Whole three-stage pipeline per replica with the reference's PCG stream: final DOFs, score,
    rejection and move counts identical to the oracle for every replica, and the same argmax."""
    bopt = orc.build_optimiser(**kw)
    best, res, idx, obest, ores = run_both(eng, spec, group, bopt, 24)
    for k in range(24):
        assert list(res[k].dof) == list(ores[k].dof), k
        assert same_float(res[k].score, ores[k].score), k
        assert (res[k].rejections, res[k].moves) == (ores[k].rejections, ores[k].moves), k
    assert best.best_replica == idx and best.best_score == obest.score


@pytest.mark.parametrize("lanes", [1, 4, 8, 16, 32])
def test_lane_width_does_not_change_results(eng, lanes):
    bopt = orc.build_optimiser(steps=800, inner_steps=200)
    for spec, group in ((("polygon", 5), "p2gg"), (("trimer",), "p2")):
        best, res, idx, obest, ores = run_both(eng, spec, group, bopt, 16, lanes=lanes)
        assert [list(r.dof) for r in res] == [list(r.dof) for r in ores]
        assert [r.score for r in res] == [r.score for r in ores]


@pytest.mark.parametrize("spec,group,kw", TRAJ_CASES)
def test_trajectories_bit_exact_thread_kernel(eng, spec, group, kw):
    """The one-thread-per-replica kernels (lanes = 1) against the oracle, same cases."""
    bopt = orc.build_optimiser(**kw)
    best, res, idx, obest, ores = run_both(eng, spec, group, bopt, 70, lanes=1)  # 70: a ragged last block
    for k in range(70):
        assert list(res[k].dof) == list(ores[k].dof), k
        assert same_float(res[k].score, ores[k].score), k
        assert (res[k].rejections, res[k].moves) == (ores[k].rejections, ores[k].moves), k
    assert best.best_replica == idx and best.best_score == obest.score


def test_philox_stream_is_kernel_independent(eng):
    """Every kernel variant consumes the Philox stream identically."""
    state = cp.PackedState.from_group(cp.LineShape.polygon(5), "p2")
    opts = cp.BuildOptimiser.cli_defaults(steps=600, inner_steps=200)
    runs = []
    for lanes in (1, 8, 32):
        eng.set_lanes(lanes)
        _, res = cp.monte_carlo_best_packing(state, opts, 200, engine=eng, rng=cp.RNG_PHILOX, return_all=True)
        runs.append([(list(r.dof), r.score, r.rejections) for r in res])
    eng.set_lanes(0)
    assert runs[0] == runs[1] == runs[2]


def test_replay_thread_kernel(eng):
    eng.set_lanes(1)
    try:
        for spec, group in ((("polygon", 5), "p2"), (("polygon", 7), "p2gg"), (("trimer",), "p2gg")):
            seqs, inits, dumps = [], [], []
            for seed in range(3):
                cstate, init, recs = dump_moves(spec, group, orc.MATH_CP, seed)
                seqs.append([(r.basis_index, r.new_value, 0.0 if math.isnan(r.threshold) else r.threshold, r.kt)
                             for r in recs])
                inits.append(init)
                dumps.append(recs)
            out = eng.replay(cstate._problem, inits, seqs)
            for recs, decs in zip(dumps, out):
                for r, d in zip(recs, decs):
                    assert (d.in_bounds, d.accepted) == (r.in_bounds, r.accepted)
                    assert same_float(d.new_score, r.new_score) and d.score_current == r.score_current
    finally:
        eng.set_lanes(0)


def test_trajectories_against_glibc_oracle(eng):
    """Oracle in libm mode = what the Rust reference executes.  DOF trajectories depend only on the
    RNG stream and the decisions, so they must be bit-identical; the score differs at most in the
    last place (sin of the cell angle)."""
    bopt = orc.build_optimiser(steps=2000, inner_steps=500)
    for spec, group in ((("polygon", 4), "p2"), (("polygon", 5), "p2"), (("trimer",), "p2gg")):
        best, res, idx, obest, ores = run_both(eng, spec, group, bopt, 32, math_mode=orc.MATH_LIBM)
        for k in range(32):
            assert list(res[k].dof) == list(ores[k].dof), (spec, group, k)
            assert res[k].score == pytest.approx(ores[k].score, rel=1e-14)
            assert res[k].rejections == ores[k].rejections


@pytest.mark.parametrize("lanes", [1, 8, 32])
def test_lj_trajectories(eng, lanes):
    """BASELINE.json configs[3] (LJ trimer / circle): the energy sum keeps the reference's order, so
    whole trajectories are bit-identical as well (the contract only asks for 1e-6 relative)."""
    bopt = orc.build_optimiser(steps=600, inner_steps=200)
    for spec, group in ((("lj_trimer",), "p2"), (("lj_circle",), "p1"), (("lj_trimer", 0.63), "p2gg"),
                        (("lj_circle",), "p2mg"), (LJ_DIMER, "p2"), (LJ_PAIR_NOCUT, "p1g1")):
        best, res, idx, obest, ores = run_both(eng, spec, group, bopt, 16, lanes=lanes)
        for k in range(16):
            assert res[k].score == pytest.approx(ores[k].score, rel=LJ_RTOL)
            assert list(res[k].dof) == list(ores[k].dof) and res[k].score == ores[k].score, (spec, group, k)
            assert res[k].rejections == ores[k].rejections
        assert best.best_replica == idx and best.best_score == obest.score


@pytest.mark.parametrize("group", GROUPS)
def test_c5_sweep_polygons_all_groups(eng, group):
    """BASELINE.json configs[4]: 3..12-sided polygons x all seven plane groups, default kernel
    selection, whole pipeline bit-identical to the oracle (few replicas, short schedule)."""
    bopt = orc.build_optimiser(steps=300, inner_steps=100)
    for sides in range(3, 13):
        best, res, idx, obest, ores = run_both(eng, ("polygon", sides), group, bopt, 6)
        for k in range(6):
            assert list(res[k].dof) == list(ores[k].dof), (sides, group, k)
            assert res[k].score == ores[k].score and res[k].rejections == ores[k].rejections
        assert best.best_replica == idx


def test_reference_integration_tests(eng):
    # crystal-packing/tests/packing.rs:14-52 and tests/potential.rs:14-52 through the mirrored API
    state = cp.PackedState.from_group(cp.LineShape.from_radial("Square", [1.0] * 4), "p2")
    init = state.score(eng)
    final_state = cp.MCOptimiser(0.0, 0.0, 0.001, 1000, 100, 0, None).optimise_state(state, eng)
    assert init < final_state.score(eng)
    assert (init, final_state.score(eng), final_state.last_rejections) == (0.0625, 0.6424637854786934, 525)
    lj = cp.PotentialState.from_group(cp.LJShape2.from_trimer(0.63, 120.0, 1.0), "p2")
    init = lj.score(eng)
    final_lj = cp.MCOptimiser(0.0, 0.0, 0.001, 1000, 100, 0, None).optimise_state(lj, eng)
    assert init < final_lj.score(eng)
    assert final_lj.score(eng) == pytest.approx(2.0102873836561734, rel=1e-12)
    assert final_lj.last_rejections == 691


def test_per_replica_initial_states_and_split_calls(eng):
    """cp_run with init_dof (each replica continues from its own state, as a resumed run would),
    and cp_prepare + cp_launch + cp_fetch giving the same answer as cp_run."""
    import random

    rnd = random.Random(9)
    cstate, ostate = make_states(("polygon", 5), "p2gg")
    base = orc.get_dof(ostate)
    inits = [[base[0] * rnd.uniform(0.7, 1.0), rnd.uniform(0.8, 1.0), base[2], rnd.uniform(-0.3, -0.2),
              rnd.uniform(-0.3, -0.2), rnd.uniform(0.0, 0.3)] for _ in range(40)]
    assert all(not math.isnan(v) for v in oracle_scores(ostate, inits))
    sched = cp.MCOptimiser(0.05, 0.5, 0.01, 900, 300, 0, None).schedule
    for lanes in (0, 32):
        eng.set_lanes(lanes)
        best, res = eng.run(cstate._problem, [sched], 40, replica_begin=100, rng=cp.RNG_PCG64MCG, init_dof=inits)
        for k, init in enumerate(inits):
            s = orc.clone(ostate)
            orc.set_dof(s, init)
            rc, _, rej = orc.optimise(s, 0.05, 0.5, 0.01, 900, 300, 100 + k)
            assert rc == 0 and list(res[k].dof) == orc.get_dof(s) and res[k].rejections == rej
            assert res[k].score == orc.score(s)
        scores = [r.score for r in res]
        assert best.replica == 100 + scores.index(max(scores)) and best.result.score == max(scores)
    eng.set_lanes(0)
    flat = (C.c_double * (6 * 40))(*[v for d in inits for v in d])
    eng.prepare(cstate._problem, [sched], 40, replica_begin=100, rng=cp.RNG_PCG64MCG, init_dof_buffer=flat)
    eng.launch()
    out = (cp.capi.ReplicaResult * 40)()
    best2 = eng.fetch(out)
    assert [list(r.dof) for r in out] == [list(r.dof) for r in res] and best2.replica == best.replica
    t = eng.timing()
    assert t.launches == 3 and t.kernel_ms > 0 and t.lanes_per_replica == 1


def test_million_replicas(eng):
    """BASELINE.json configs[3]/[4] sizes: 1,048,576 replicas in one launch (short schedule)."""
    state = cp.PackedState.from_group(cp.LineShape.polygon(4), "p1")
    opts = cp.BuildOptimiser.cli_defaults(steps=50, inner_steps=50)
    n = 1 << 20
    eng.prepare(state._problem, opts.pipeline(), n, rng=cp.RNG_PHILOX)
    eng.launch()
    out = (cp.capi.ReplicaResult * n)()
    best = eng.fetch(out)
    assert all(out[i].moves == 200 for i in range(0, n, 4099)) and out[n - 1].moves == 200
    assert 0.125 < best.result.score <= 1.0 and best.replica < n
    assert out[best.replica].score == best.result.score


def test_invalid_initial_state_is_an_error(eng):
    # optimisation.rs:81-84: panic!("Invalid configuration passed to function, exiting.")
    state = cp.PackedState.from_group(cp.LineShape.polygon(4), "p2")
    state.dof = [0.5, 1.0, math.pi / 2, -0.25, -0.25, 0.0]
    with pytest.raises(cp.CrystalPackingError) as e:
        cp.MCOptimiser(0.0, 0.0, 0.001, 100, 100, 0, None).optimise_state(state, eng)
    assert e.value.status == cp.capi.CP_ERR_INVALID_STATE


def test_zero_steps_returns_initial_state(eng):
    state = cp.PackedState.from_group(cp.LineShape.polygon(5), "p2")
    out = cp.MCOptimiser(0.1, 0.5, 0.01, 0, 100, 0, None).optimise_state(state, eng)
    assert out.dof == state.dof


# ------------------------------------------------------------------ golden fixtures
def test_golden_trajectories(eng):
    """tests/golden/trajectories.json: final states written by tests/golden/make_golden.py from the
    oracle in libm mode (the Rust reference cannot run here: 'parity unpinned')."""
    cases = json.load(open(os.path.join(GOLDEN, "trajectories.json")))["cases"]
    for case in cases:
        spec = tuple(tuple(x) if isinstance(x, list) else x for x in case["spec"])
        cstate, _ = make_states(spec, case["group"])
        b = cp.BuildOptimiser(**case["optimiser"])
        _, res = cp.monte_carlo_best_packing(cstate, b, len(case["replicas"]), engine=eng, return_all=True)
        for got, exp in zip(res, case["replicas"]):
            if case["kind"] == "lj":
                assert got.score == pytest.approx(float.fromhex(exp["score"]), rel=LJ_RTOL)
            else:
                assert [v.hex() for v in got.dof] == exp["dof"]
                assert got.score == pytest.approx(float.fromhex(exp["score"]), rel=1e-14)
                assert got.rejections == exp["rejections"]


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_properties(eng):
    """BASELINE.json configs[1] size: 65,536 pentagon/p2 replicas (Philox stream)."""
    n = 65536
    state = cp.PackedState.from_group(cp.LineShape.polygon(5), "p2")
    opts = cp.BuildOptimiser.cli_defaults(steps=400, inner_steps=100)
    best, res = cp.monte_carlo_best_packing(state, opts, n, engine=eng, rng=cp.RNG_PHILOX, return_all=True)
    scores = [r.score for r in res]
    assert all(0.0 < s <= 1.0 for s in scores)                    # packing fractions, none invalid
    last_max = n - 1 - scores[::-1].index(max(scores))          # ties: the last maximal replica (rayon's max)
    assert best.best_score == max(scores) and best.best_replica == last_max
    assert all(r.moves == 100 + 400 + 400 for r in res)
    # idempotence: re-scoring the returned states reproduces the reported score bit for bit
    sample = list(range(0, n, 37))
    again = eng.score(state._problem, [list(res[i].dof) for i in sample])
    assert again == [scores[i] for i in sample]
    # determinism and shard independence: any sub-range reproduces its slice of the full run
    _, part = cp.monte_carlo_best_packing(state, opts, 1000, engine=eng, rng=cp.RNG_PHILOX,
                                          replica_begin=40000, return_all=True)
    assert [list(r.dof) for r in part] == [list(r.dof) for r in res[40000:41000]]
    # replicas are distinct streams
    assert len({r.dof[0] for r in res[:2000]}) > 1900


@pytest.mark.parametrize("spec,group,n,kw", [
    (("trimer",), "p2gg", 512, dict(steps=900, inner_steps=300, kt_start=0.1, kt_finish=0.001)),
    (("lj_circle",), "p1", 512, dict(steps=600, inner_steps=200)),
    (("polygon", 4), "p2mm", 384, dict(steps=900, inner_steps=300)),
])
def test_philox_equivalent_other_configs(eng, spec, group, n, kw):
    """Same statistical check for the disc molecule with annealing (configs[2]), the LJ circle
    (configs[3]) and an N = 4 polygon group."""
    from scipy.stats import ks_2samp

    bopt = orc.build_optimiser(**kw)
    cstate, ostate = make_states(spec, group)
    _, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=eng, rng=cp.RNG_PHILOX, return_all=True)
    _, _, ores, _ = orc.analyse_state(ostate, bopt, 0, n, n_threads=os.cpu_count() or 1)
    a, b = [r.score for r in res], [r.score for r in ores]
    assert ks_2samp(a, b).pvalue > 1e-3

    def mean_var(v):
        m = sum(v) / len(v)
        return m, sum((x - m) ** 2 for x in v) / (len(v) - 1)

    (ma, va), (mb, vb) = mean_var(a), mean_var(b)
    assert abs(ma - mb) < 4.0 * math.sqrt(va / n + vb / n)  # means agree within 4 standard errors
    ra, rb = [float(r.rejections) for r in res], [float(r.rejections) for r in ores]
    (ma, va), (mb, vb) = mean_var(ra), mean_var(rb)
    assert abs(ma - mb) < 4.0 * math.sqrt(va / n + vb / n)


def test_custom_group_three_operators_hexagonal(eng):
    """A plane group the reference's table does not hold (three-fold rotation, Hexagonal family:
    the only cell DOF is the length, cell.rs:136-170, angle fixed at pi/3): three symmetry
    operators exercise the generic N and the 4-entry basis on both sides."""
    ops = ["x,y", "-y,x-y", "-x+y,-x"]
    for spec in (("polygon", 5), ("trimer",)):
        cshape, oshape = __import__("common").make_shapes(spec)
        ostate = orc.state_from_group(oshape, "p1", orc.MATH_CP)
        ostate.family, ostate.n_syms = orc.HEXAGONAL, 3
        mats = []
        for i, o in enumerate(ops):
            assert orc.lib().orc_transform_from_operations(o.encode(), C.byref(ostate.syms[i])) == 0
            r = ostate.syms[i].rows()
            mats.append([r[0][0], r[1][0], r[2][0], r[0][1], r[1][1], r[2][1], r[0][2], r[1][2], r[2][2]])
        radius = orc.lib().orc_shape_enclosing_radius(oshape[0], oshape[1], len(oshape[1]))
        dof = [4.0 * radius * 3, 1.0, math.pi / 3, -0.5 + 0.5 / 3, -0.5 + 0.5 / 3, 0.0]
        orc.set_dof(ostate, dof)
        group = cp.api.group_from_symmetries("p3", "Hexagonal", mats)
        assert group.key is None
        cstate = cp.PackedState.from_group(cshape, group)
        cstate.dof = dof
        assert len(cstate.generate_basis()) == 4 == orc.lib().orc_state_n_dof(C.byref(ostate))
        dofs = [dof] + [[dof[0] * f, 1.0, math.pi / 3, x, y, th] for f, x, y, th in
                        [(0.9, -0.3, -0.2, 0.4), (0.5, 0.1, -0.4, 2.0), (0.3, 0.25, 0.25, 5.0), (0.2, -0.1, 0.3, 1.0)]]
        ref = oracle_scores(ostate, dofs)
        for lanes in (1, 8, 32):
            eng.set_lanes(lanes)
            got = eng.score(cstate._problem, dofs)
            assert all(same_float(a, b) for a, b in zip(got, ref)), (spec, lanes, got, ref)
            sched = cp.MCOptimiser(0.05, 0.5, 0.01, 600, 200, 0, None).schedule
            _, res = eng.run(cstate._problem, [sched], 12, rng=cp.RNG_PCG64MCG)
            for k in range(12):
                s = orc.clone(ostate)
                rc, _, rej = orc.optimise(s, 0.05, 0.5, 0.01, 600, 200, k)
                assert rc == 0 and list(res[k].dof) == orc.get_dof(s) and res[k].rejections == rej, (spec, lanes, k)
        eng.set_lanes(0)


@pytest.mark.parametrize("ops,family", [
    (["-x+1/2,y+1/2", "x,y", "x+1/2,-y+1/2", "-x,-y"], "Orthorhombic"),   # p2gg, reordered: operator 0 is a glide
    (["-x,-y", "x,y"], "Monoclinic"),                                      # p2 with the two-fold axis first
    (["x,y", "-x,-y", "y,-x", "-y,x"], "Orthorhombic"),                    # a four-fold axis: NOT row-sign related
])
def test_custom_groups_signed_and_unsigned_images(eng, ops, family):
    """The multiplicity-4 builds and the LJ kernel store the rotated points of image 0 only when every
    operator's rotation part is operator 0's with the signs of its rows flipped (KProblem::
    signed_images), whatever operator 0 is; a group with a four-fold axis is not of that kind and
    keeps the points of every image.  Scores and trajectories against the oracle, hard shapes and LJ."""
    fam = {"Monoclinic": orc.MONOCLINIC, "Orthorhombic": orc.ORTHORHOMBIC}[family]
    for spec in (("polygon", 5), ("trimer",), ("lj_trimer",)):
        cshape, oshape = __import__("common").make_shapes(spec)
        ostate = orc.state_from_group(oshape, "p1", orc.MATH_CP)
        ostate.family, ostate.n_syms = fam, len(ops)
        mats = []
        for i, o in enumerate(ops):
            assert orc.lib().orc_transform_from_operations(o.encode(), C.byref(ostate.syms[i])) == 0
            r = ostate.syms[i].rows()
            mats.append([r[0][0], r[1][0], r[2][0], r[0][1], r[1][1], r[2][1], r[0][2], r[1][2], r[2][2]])
        radius = orc.lib().orc_shape_enclosing_radius(oshape[0], oshape[1], len(oshape[1]))
        angle = math.pi / 2
        dof = [4.0 * radius * len(ops), 1.0, angle, -0.5 + 0.5 / len(ops), -0.5 + 0.5 / len(ops), 0.0]
        orc.set_dof(ostate, dof)
        group = cp.api.group_from_symmetries("custom", family, mats)
        cls = cp.PotentialState if spec[0].startswith("lj") else cp.PackedState
        cstate = cls.from_group(cshape, group)
        cstate.dof = dof
        dofs = [dof] + random_dofs(ostate, 150, seed=17)
        ref = oracle_scores(ostate, dofs)
        got = eng.score(cstate._problem, dofs)
        assert all(same_float(a, b) for a, b in zip(got, ref)), (spec, ops)
        assert any(math.isnan(v) for v in ref) or spec[0].startswith("lj")  # overlapping states are in the sample
        sched = cp.MCOptimiser(0.05, 0.5, 0.01, 900, 300, 0, None).schedule
        _, res = eng.run(cstate._problem, [sched], 16, rng=cp.RNG_PCG64MCG)
        for k in range(16):
            s = orc.clone(ostate)
            rc, _, rej = orc.optimise(s, 0.05, 0.5, 0.01, 900, 300, k)
            assert rc == 0 and list(res[k].dof) == orc.get_dof(s) and res[k].rejections == rej, (spec, ops, k)
            assert res[k].score == orc.score(s)


def test_large_replica_ids_and_seed_base(eng):
    """Seeds are 64-bit: replica ids beyond 2^32 and a seed_base offset reproduce the oracle's
    seed_from_u64(seed_base + id) streams."""
    cstate, ostate = make_states(("polygon", 5), "p2")
    bopt = orc.build_optimiser(steps=300, inner_steps=100)
    begin = (1 << 40) + 12345
    best, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), 40, engine=eng, replica_begin=begin, return_all=True)
    idx, obest, ores, _ = orc.analyse_state(ostate, bopt, begin, 40, n_threads=4)
    assert [list(r.dof) for r in res] == [list(r.dof) for r in ores]
    assert best.best_replica == begin + idx
    # seed_base shifts the seed, not the replica id
    b2, r2 = eng.run(cstate._problem, c_build(bopt).pipeline(), 40, replica_begin=345, seed_base=(1 << 40) + 12000,
                     rng=cp.RNG_PCG64MCG)
    assert [list(r.dof) for r in r2] == [list(r.dof) for r in ores] and b2.replica == 345 + idx


def test_external_stream_and_null_outputs(eng):
    """cp_set_stream with a caller-owned CUDA stream (torch's), and cp_run without out_best."""
    import torch

    state = cp.PackedState.from_group(cp.LineShape.polygon(4), "p2")
    sched = cp.BuildOptimiser.cli_defaults(steps=200, inner_steps=100).pipeline()
    ref_best, ref = eng.run(state._problem, sched, 100, rng=cp.RNG_PHILOX)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    try:
        eng.prepare(state._problem, sched, 100, rng=cp.RNG_PHILOX)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        eng.launch()
        stop.record(stream)
        stream.synchronize()
        assert start.elapsed_time(stop) > 0
        out = (cp.capi.ReplicaResult * 100)()
        best = eng.fetch(out)
        assert [list(r.dof) for r in out] == [list(r.dof) for r in ref] and best.replica == ref_best.replica
    finally:
        eng.set_stream(0)
    arr = (cp.capi.Schedule * 3)(*sched)
    out = (cp.capi.ReplicaResult * 100)()
    assert eng._lib.cp_run(eng._ctx, C.byref(state._problem), arr, 3, cp.RNG_PHILOX, 0, 0, 100, None, None, out) == 0
    assert [list(r.dof) for r in out] == [list(r.dof) for r in ref]


def test_argument_errors(eng):
    """Error behaviour at the C ABI: bad arguments are refused with a status and a message, never a
    crash or a silent fallback."""
    state = cp.PackedState.from_group(cp.LineShape.polygon(5), "p2")
    sched = cp.BuildOptimiser.cli_defaults(steps=10).pipeline()
    lib, ctx = eng._lib, eng._ctx
    arr = (cp.capi.Schedule * 3)(*sched)
    best = cp.capi.Best()
    ok = cp.capi.CP_OK
    assert lib.cp_run(ctx, C.byref(state._problem), arr, 0, 0, 0, 0, 8, None, C.byref(best), None) == cp.capi.CP_ERR_INVALID_ARG
    assert lib.cp_run(ctx, C.byref(state._problem), arr, 5, 0, 0, 0, 8, None, C.byref(best), None) == cp.capi.CP_ERR_INVALID_ARG
    assert lib.cp_run(ctx, C.byref(state._problem), arr, 3, 7, 0, 0, 8, None, C.byref(best), None) == cp.capi.CP_ERR_INVALID_ARG
    assert lib.cp_run(ctx, C.byref(state._problem), arr, 3, 0, 0, 0, 0, None, C.byref(best), None) == cp.capi.CP_ERR_INVALID_ARG
    assert b"n_replicas" in lib.cp_last_error()
    bad = cp.capi.Problem()
    C.memmove(C.byref(bad), C.byref(state._problem), C.sizeof(bad))
    bad.n_items = 99
    assert lib.cp_run(ctx, C.byref(bad), arr, 3, 0, 0, 0, 8, None, C.byref(best), None) == cp.capi.CP_ERR_INVALID_ARG
    assert lib.cp_set_lanes(ctx, 3) == cp.capi.CP_ERR_INVALID_ARG
    assert lib.cp_launch(C.c_void_p()) == cp.capi.CP_ERR_INVALID_ARG
    fresh = cp.Engine(0)
    assert fresh._lib.cp_launch(fresh._ctx) == cp.capi.CP_ERR_NOT_PREPARED
    fresh.close()
    assert lib.cp_run(ctx, C.byref(state._problem), arr, 3, 0, 0, 0, 8, None, C.byref(best), None) == ok
    dc = C.c_int()
    assert lib.cp_device_count(C.byref(dc)) == ok and dc.value >= 1


def test_philox_statistically_equivalent(eng):
    """Free-running production stream vs the oracle's PCG stream: same distribution of final packing
    fractions over replicas (two-sample KS), and the same best packing to 2%."""
    from scipy.stats import ks_2samp

    n = 768
    bopt = orc.build_optimiser(steps=1500, inner_steps=500)
    cstate, ostate = make_states(("polygon", 5), "p2")
    _, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=eng, rng=cp.RNG_PHILOX, return_all=True)
    _, obest, ores, _ = orc.analyse_state(ostate, bopt, 0, n, n_threads=os.cpu_count() or 1)
    a, b = [r.score for r in res], [r.score for r in ores]
    assert ks_2samp(a, b).pvalue > 1e-3
    qa, qb = sorted(a)[int(0.9 * n)], sorted(b)[int(0.9 * n)]  # upper decile: where the best packings live
    assert qa == pytest.approx(qb, rel=0.02)
    assert sum(a) / n == pytest.approx(sum(b) / n, rel=0.01)
    ra, rb = sum(r.rejections for r in res) / n, sum(r.rejections for r in ores) / n
    assert ra == pytest.approx(rb, rel=0.02)


# ------------------------------------------------------------------ several GPUs behind one handle
def test_multi_engine_matches_single_context(eng):
    """cp_init_multi / cp_multi_run: the replicas of analyse_state's loop (main.rs:221-253) sharded
    over a device list.  Seeds are global replica ids, so the result equals the one-context run and
    the oracle's, whatever the list - here every visible GPU, or the same GPU three times."""
    import torch

    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0, 0]
    multi = cp.MultiEngine(devices)
    try:
        for spec, group in ((("polygon", 5), "p2"), (("trimer",), "p2gg"), (("lj_trimer",), "p2")):
            cstate, ostate = make_states(spec, group)
            bopt = orc.build_optimiser(steps=200, inner_steps=100)
            n, begin = 101, 7
            best, res = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=multi, replica_begin=begin,
                                                    return_all=True)
            best1, res1 = cp.monte_carlo_best_packing(cstate, c_build(bopt), n, engine=eng, replica_begin=begin,
                                                      return_all=True)
            assert [(list(r.dof), r.score, r.rejections, r.moves) for r in res] == \
                   [(list(r.dof), r.score, r.rejections, r.moves) for r in res1]
            assert (best.best_replica, best.best_score, best.dof) == (best1.best_replica, best1.best_score, best1.dof)
            idx, obest, ores, _ = orc.analyse_state(ostate, bopt, begin, n, n_threads=4)
            assert best.best_replica == begin + idx and best.best_score == obest.score
            shards = multi.shards()
            assert [s[0] for s in shards][0] == begin and sum(s[1] for s in shards) == n
            assert max(s[1] for s in shards) - min(s[1] for s in shards) <= 1
        # fewer replicas than devices: the empty shards are skipped
        cstate, _ = make_states(("polygon", 4), "p1")
        best, res = cp.monte_carlo_best_packing(cstate, cp.BuildOptimiser.cli_defaults(steps=100, inner_steps=100),
                                                len(devices) - 1, engine=multi, return_all=True)
        assert len(res) == len(devices) - 1 and best.best_score > 0
    finally:
        multi.close()


def test_multi_engine_rejects_bad_lists():
    with pytest.raises(cp.CrystalPackingError):
        cp.MultiEngine([])
    with pytest.raises(cp.CrystalPackingError):
        cp.MultiEngine([0, 4096])
