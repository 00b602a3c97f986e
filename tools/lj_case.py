"""One LJ workload for profiling: BASELINE.json configs[3], LJShape2::from_trimer in p2gg,
131,072 replicas, 3-stage pipeline with 100 + steps + steps moves (default 200; the packings become
dense - and a move expensive - after ~2,000).  usage: python tools/lj_case.py [steps] [replicas]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crystal_packing_b200 as cp  # noqa: E402

eng = cp.Engine(0)
state = cp.PotentialState.from_group(cp.LJShape2.from_trimer(0.637556, 120.0, 1.0), "p2gg")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sched = cp.BuildOptimiser.cli_defaults(steps=steps, inner_steps=min(1000, steps)).pipeline()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
for _ in range(2):
    eng.prepare(state._problem, sched, n, rng=cp.RNG_PHILOX)
    eng.launch()
    eng.synchronize()
t = eng.timing()
moves = n * sum((s.steps // s.inner_steps) * s.inner_steps for s in sched)
print(f"LJ trimer p2gg: {moves / (t.kernel_ms * 1e-3):.4g} moves/s, kernel {t.kernel_ms:.1f} ms")
