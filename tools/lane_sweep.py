"""Kernel variant comparison per configuration.  usage: python tools/lane_sweep.py [mc_steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crystal_packing_b200 as cp  # noqa: E402

mc_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eng = cp.Engine(0)
cases = [
    ("pentagon p2", cp.PackedState, cp.LineShape.polygon(5), "p2", 65536),
    ("pentagon p2gg", cp.PackedState, cp.LineShape.polygon(5), "p2gg", 65536),
    ("8-gon p2mm", cp.PackedState, cp.LineShape.polygon(8), "p2mm", 65536),
    ("12-gon p2", cp.PackedState, cp.LineShape.polygon(12), "p2", 65536),
    ("12-gon p2gg", cp.PackedState, cp.LineShape.polygon(12), "p2gg", 65536),
    ("trimer p2gg", cp.PackedState, cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p2gg", 65536),
    ("circle p2", cp.PackedState, cp.MolecularShape2.circle(), "p2", 65536),
]
for name, cls, shape, group, n in cases:
    state = cls.from_group(shape, group)
    opts = cp.BuildOptimiser.cli_defaults(steps=mc_steps, inner_steps=min(1000, mc_steps))
    sched = opts.pipeline()
    moves = n * sum((s.steps // s.inner_steps) * s.inner_steps for s in sched)
    row = []
    for lanes in (1, 4, 8, 32):
        eng.set_lanes(lanes)
        eng.prepare(state._problem, sched, n, rng=cp.RNG_PHILOX)
        eng.launch()
        eng.synchronize()
        eng.launch()
        eng.synchronize()
        t = eng.timing()
        row.append(f"L{t.lanes_per_replica}: {moves / (t.kernel_ms * 1e-3):9.3g}")
    print(f"{name:16s} " + "  ".join(row), flush=True)
