"""The C++ host layer (include/crystal_packing.hpp): compiles against the C ABI, its host-side
constructors agree with the oracle, and — on a GPU — the reference's integration test runs through it."""
import os
import subprocess

import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "crystal-packing_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "host_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_demo.cpp"), "-L", LIBDIR, "-lcpgpu",
                           f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def test_cpp_host_constructors(tmp_path):
    out = subprocess.run([build(tmp_path), "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    first, second, third = out.stdout.splitlines()[:3]
    tok = first.split()
    kind, items = orc.polygon(4)
    assert float.fromhex(tok[1]) == orc.lib().orc_shape_area(kind, items, 4)
    assert float.fromhex(tok[3]) == orc.lib().orc_shape_enclosing_radius(kind, items, 4)
    assert tok[5] == "2"
    ostate = orc.state_from_group((kind, items), "p2")
    assert [float.fromhex(v) for v in tok[7:13]] == orc.get_dof(ostate)
    assert tok[14] == "6"
    assert float.fromhex(second.split()[1]) == pow(0.001 / 0.1, 1.0 / 100) and second.split()[3] == "100"
    assert third.startswith("error 1 The number of points provided is too few")


@pytest.mark.gpu
def test_cpp_host_packing_improves(tmp_path):
    out = subprocess.run([build(tmp_path), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("init")][0].split()
    assert (float(line[1]), float(line[3])) == (0.0625, 0.6424637854786934)  # oracle / survey anchor
    ostate = orc.state_from_group(orc.polygon(4), "p2", orc.MATH_CP)
    idx, best, _, _ = orc.analyse_state(ostate, orc.build_optimiser(steps=300, inner_steps=100, kt_finish=0.001), 0, 32, 4)
    bl = [l for l in out.stdout.splitlines() if l.startswith("best")][0].split()
    assert float(bl[1]) == best.score and int(bl[3]) == idx


@pytest.mark.gpu
def test_cpp_multi_gpu_same_answer(tmp_path):
    """cp_init_multi / cp_multi_run (the replica loop of main.rs:221-253 sharded over a device list)
    gives the single-context answer: best state, replica index and every replica's result, for all
    three state types.  On a one-GPU box the list is (0, 0): two contexts, two streams; with more
    GPUs every device takes a shard."""
    import torch

    n = torch.cuda.device_count()
    ids = [str(k) for k in range(n)] if n > 1 else ["0", "0"]
    out = subprocess.run([build(tmp_path), "multi", *ids], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("multi")]
    assert len(lines) == 3 and all(l.endswith("same") for l in lines), out.stdout
    last = [l for l in out.stdout.splitlines() if l.startswith("last shard")][0].split()
    assert int(last[2]) + int(last[3].lstrip("+")) == 5 + 97
