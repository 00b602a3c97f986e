"""Throughput of the BASELINE.json configurations other than the bench workload (device-resident,
Philox stream, CUDA-event kernel time).  usage: python tools/config_sweep.py [mc_steps]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crystal_packing_b200 as cp  # noqa: E402

mc_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
eng = cp.Engine(0)
cases = [
    ("C1 square p2", cp.PackedState, cp.LineShape.polygon(4), "p2", 65536, {}),
    ("C2 pentagon p1", cp.PackedState, cp.LineShape.polygon(5), "p1", 65536, {}),
    ("C2 pentagon p2", cp.PackedState, cp.LineShape.polygon(5), "p2", 65536, {}),
    ("C2 pentagon p2mg", cp.PackedState, cp.LineShape.polygon(5), "p2mg", 65536, {}),
    ("C2 pentagon p2gg", cp.PackedState, cp.LineShape.polygon(5), "p2gg", 65536, {}),
    ("C3 trimer p2gg anneal", cp.PackedState, cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p2gg", 262144,
     dict(kt_finish=0.001)),
    ("C4 LJ trimer p2gg", cp.PotentialState, cp.LJShape2.from_trimer(0.637556, 120.0, 1.0), "p2gg", 131072, {}),
    ("C4 LJ circle p1", cp.PotentialState, cp.LJShape2.circle(), "p1", 131072, {}),
    ("C5 triangle p1m1", cp.PackedState, cp.LineShape.polygon(3), "p1m1", 65536, {}),
    ("C5 12-gon p2gg", cp.PackedState, cp.LineShape.polygon(12), "p2gg", 65536, {}),
    ("C5 8-gon p2mm", cp.PackedState, cp.LineShape.polygon(8), "p2mm", 65536, {}),
]
for name, cls, shape, group, n, kw in cases:
    state = cls.from_group(shape, group)
    steps = mc_steps if "LJ" not in name else max(200, mc_steps // 10)
    opts = cp.BuildOptimiser.cli_defaults(steps=steps, inner_steps=min(1000, steps), **kw)
    sched = opts.pipeline()
    moves = n * sum((s.steps // s.inner_steps) * s.inner_steps for s in sched)
    eng.prepare(state._problem, sched, n, rng=cp.RNG_PHILOX)
    eng.launch()
    eng.synchronize()
    eng.launch()
    eng.synchronize()
    t = eng.timing()
    best = eng.fetch()
    print(f"{name:24s} replicas {n:7d} lanes {t.lanes_per_replica:2d}  {moves / (t.kernel_ms * 1e-3):10.4g} moves/s  "
          f"kernel {t.kernel_ms:9.1f} ms  best score {best.result.score:.6f}", flush=True)
