"""CPU soundness check of the single-precision certificates (crystal-packing_b200/csrc/cp_filter.h).

tests/filter_check.cpp compiles the SAME header the kernels compile and, along real three-stage
trajectories of the oracle, checks at every scored proposal that
  * the single-precision image sweep lists every image the reference's f64 filter keeps
    (packed.rs:152-157), itself or as its mirror image;
  * every component pair outside a certificate's ambiguity mask is false with the reference's f64
    expressions, on both mirror sides (line2.rs:29-54, atom2.rs:21-26);
  * every certified hit is a true shape test of a candidate the reference does test;
  * the verdict assembled the way the kernel assembles it equals check_intersection
    (packed.rs:127-167);
  * the certificate evaluated from image 0's points alone (row signs, shape i's frame - what the
    multiplicity-4 builds do) has the same cmin, tau and ambiguous set, bit for bit.
No GPU involved: this is the proof obligation of the certificates exercised on the host.
"""
import json
import os
import subprocess

import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")


@pytest.fixture(scope="module")
def checker():
    orc.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "filter_check")
    src = os.path.join(HERE, "filter_check.cpp")
    deps = [src, os.path.join(ROOT, "crystal-packing_b200", "csrc", "cp_filter.h"),
            os.path.join(ROOT, "oracle", "libcp_oracle.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, src,
                               "-L" + os.path.join(ROOT, "oracle"), "-lcp_oracle",
                               "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lm"])
    return exe


CASES = [
    # shape, sides, group, replicas, steps, extra (kt_start, kt_finish)
    ("polygon", 5, "p2", 3, 10000, ()),          # the bench workload, full schedule
    ("polygon", 5, "p2gg", 2, 3000, ()),
    ("polygon", 3, "p2mm", 2, 3000, ()),         # mirror images: aligned edges at theta = 0
    ("polygon", 12, "p1g1", 1, 2000, ()),
    ("polygon", 4, "p1", 2, 4000, ()),
    ("star", 8, "p2", 1, 3000, ()),              # non-convex chain
    ("soup", 5, "p2mg", 1, 2000, ()),            # run-time-count build: unrelated segments
    ("trimer", 3, "p2gg", 2, 3000, ("0.1", "0.001")),
    ("circle", 1, "p2mm", 2, 3000, ()),          # touching discs sit exactly on the 2R filter
]


@pytest.mark.parametrize("shape,sides,group,replicas,steps,extra", CASES)
def test_certificates_are_sound(checker, shape, sides, group, replicas, steps, extra):
    out = subprocess.run([checker, shape, str(sides), group, str(replicas), str(steps), *extra],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    stats = json.loads(out.stdout.splitlines()[0])
    assert stats["errors"] == 0
    assert stats["scored"] > 1000 and stats["clear"] + stats["hit"] > 0
    # every built-in group is a signed-image group: each certificate was also evaluated the way the
    # multiplicity-4 builds do (image 0's points, row signs, shape i's frame) and gave the same bits
    assert stats["signed_checked"] >= stats["clear"] + stats["hit"] + stats["ambiguous"]
    # the certificates decide most shape pairs; the f64 path is the exception
    assert stats["ambiguous"] < 0.5 * (stats["clear"] + stats["hit"] + stats["ambiguous"])


def test_reach_limited_sweep_equals_straight_line_sweep(checker):
    """cpf_sweep_pair_reach (multiplicity-4 builds) returns the mask of cpf_sweep_pair on two million
    random cells, degenerate ones included."""
    out = subprocess.run([checker, "sweep", "2000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    stats = json.loads(out.stdout.splitlines()[0])
    assert stats["errors"] == 0 and stats["kept"] > stats["cases"]
