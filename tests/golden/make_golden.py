"""Writes tests/golden/trajectories.json: final replica states of the three-stage pipeline
(crystal-packing-cli/src/main.rs:224-252) computed by the CPU oracle in libm mode — the arithmetic
the Rust reference executes (glibc sin/cos/exp).  The Rust crate itself cannot be built in this
environment (no rustc/cargo, crates not vendored), so these vectors pin the oracle, not the Rust
binary: "parity unpinned" in the sense of DESIGN.md.  Floats are stored as hex for bit-exactness.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402

CASES = [
    # BASELINE.json configs[0..4], shortened so the oracle finishes in seconds
    dict(spec=["polygon", 4], group="p2", kind="hard", n=8,
         optimiser=dict(steps=1000, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["polygon", 5], group="p1", kind="hard", n=8,
         optimiser=dict(steps=1000, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["polygon", 5], group="p2", kind="hard", n=8,
         optimiser=dict(steps=1000, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["polygon", 5], group="p2mg", kind="hard", n=8,
         optimiser=dict(steps=1000, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["polygon", 5], group="p2gg", kind="hard", n=8,
         optimiser=dict(steps=1000, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["trimer"], group="p2gg", kind="hard", n=8,
         optimiser=dict(steps=1000, kt_start=0.1, kt_finish=0.001, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["polygon", 12], group="p2gg", kind="hard", n=4,
         optimiser=dict(steps=500, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["polygon", 3], group="p1g1", kind="hard", n=4,
         optimiser=dict(steps=500, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=250, convergence=None)),
    dict(spec=["lj_trimer"], group="p2", kind="lj", n=4,
         optimiser=dict(steps=300, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=100, convergence=None)),
    dict(spec=["lj_circle"], group="p1", kind="lj", n=4,
         optimiser=dict(steps=300, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                        inner_steps=100, convergence=None)),
]


def oracle_shape(spec):
    if spec[0] == "polygon":
        return orc.polygon(spec[1])
    if spec[0] == "trimer":
        return orc.trimer(orc.DISCS)
    if spec[0] == "lj_trimer":
        return orc.trimer(orc.LJ)
    if spec[0] == "lj_circle":
        return orc.circle(orc.LJ)
    raise ValueError(spec)


def main():
    out = {"generator": "tests/golden/make_golden.py", "math": "glibc libm (ORC_MATH_LIBM)",
           "pinned_to": "CPU oracle oracle/cp_oracle.c (Rust reference not buildable here)", "cases": []}
    for case in CASES:
        state = orc.state_from_group(oracle_shape(case["spec"]), case["group"], orc.MATH_LIBM)
        bopt = orc.build_optimiser(**case["optimiser"])
        _, _, results, _ = orc.analyse_state(state, bopt, 0, case["n"], n_threads=4)
        entry = {k: case[k] for k in ("spec", "group", "kind", "optimiser")}
        entry["replicas"] = [
            {"dof": [v.hex() for v in r.dof], "score": r.score.hex(), "rejections": r.rejections,
             "moves": r.moves} for r in results]
        out["cases"].append(entry)
    with open(os.path.join(HERE, "trajectories.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
