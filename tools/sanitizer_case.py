"""Small workload for compute-sanitizer: every kernel family once (thread and group kernels; MC,
score, replay), a few replicas and moves."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crystal_packing_b200 as cp  # noqa: E402

eng = cp.Engine(0)
cases = [(cp.PackedState, cp.LineShape.polygon(5), "p2"), (cp.PackedState, cp.LineShape.polygon(7), "p2gg"),
         (cp.PackedState, cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p2mg"),
         (cp.PotentialState, cp.LJShape2.from_trimer(0.637556, 120.0, 1.0), "p2")]
for cls, shape, group in cases:
    state = cls.from_group(shape, group)
    opts = cp.BuildOptimiser.cli_defaults(steps=60, inner_steps=20)
    for lanes in (1, 8, 32):
        eng.set_lanes(lanes)
        for rng in (cp.RNG_PCG64MCG, cp.RNG_PHILOX):
            best, res = cp.monte_carlo_best_packing(state, opts, 70, engine=eng, rng=rng, return_all=True)
        scores = eng.score(state._problem, [list(r.dof) for r in res[:40]])
        assert scores == [r.score for r in res[:40]]
        moves = [[(k % 5, state.dof[min(k % 5, 5)] * 0.999, 0.5, 0.1) for k in range(10)] for _ in range(5)]
        eng.replay(state._problem, [state.dof] * 5, moves)
print("sanitizer case done")
