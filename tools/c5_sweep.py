"""BASELINE.json configs[4]: 3..12-sided polygons x all seven plane groups.  Kernel time per cell on
one GPU (device-resident, Philox).  usage: python tools/c5_sweep.py [replicas per cell] [mc_steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crystal_packing_b200 as cp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119837  # 8,388,608 replicas / 70 cells
mc_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = cp.Engine(0)
groups = ["p1", "p2", "p1m1", "p1g1", "p2mm", "p2mg", "p2gg"]
total_ms, total_moves = 0.0, 0
print("sides " + " ".join(f"{g:>9s}" for g in groups) + "   (1e9 moves/s)")
for sides in range(3, 13):
    row = []
    for g in groups:
        state = cp.PackedState.from_group(cp.LineShape.polygon(sides), g)
        sched = cp.BuildOptimiser.cli_defaults(steps=mc_steps, inner_steps=min(1000, mc_steps)).pipeline()
        moves = n * sum((s.steps // s.inner_steps) * s.inner_steps for s in sched)
        eng.prepare(state._problem, sched, 64, rng=cp.RNG_PHILOX)  # loads this build of the kernel
        eng.launch()
        eng.synchronize()
        eng.prepare(state._problem, sched, n, rng=cp.RNG_PHILOX)
        eng.launch()
        eng.synchronize()
        t = eng.timing()
        total_ms += t.kernel_ms
        total_moves += moves
        row.append(moves / (t.kernel_ms * 1e-3) / 1e9)
    print(f"{sides:5d} " + " ".join(f"{v:9.3f}" for v in row), flush=True)
print(f"whole sweep: {70 * n} replicas, {total_moves:.4g} moves in {total_ms / 1e3:.2f} s = {total_moves / (total_ms * 1e-3):.4g} moves/s")
