"""world_size-2 `gloo` test of the multi-GPU host logic (crystal-packing_b200/distributed.py):
replica sharding and the single all-gather + argmax of per-rank best states.  The per-rank results
come from the CPU oracle here (there is no GPU in this tier); on the GPU box the same functions run
over NCCL inside bench.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import crystal_packing_b200  # noqa: F401  (registers the package)
    from crystal_packing_b200 import distributed as cpd
    import orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, count = cpd.shard_range(total, rank, world)
    state = orc.state_from_group(orc.polygon(4), "p2")
    bopt = orc.build_optimiser(steps=300, inner_steps=100)
    idx, best, _, _ = orc.analyse_state(state, bopt, begin, count, n_threads=1)
    local = cpd.pack_best(list(best.dof), best.score, best.rejections, best.moves, begin + idx)
    gathered = cpd.all_gather_best(torch.frombuffer(bytearray(local), dtype=torch.uint8))
    g = cpd.global_best(gathered)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{g['replica']} {g['score'].hex()} {' '.join(v.hex() for v in g['dof'])}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    import crystal_packing_b200  # noqa: F401
    from crystal_packing_b200 import distributed as cpd

    for total in (1, 7, 8, 65536, 8388608 + 3):
        for world in (1, 2, 4, 8):
            spans = [cpd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert cpd.weak_range(65536, 3) == (196608, 65536)


def test_pick_best_rules():
    sys.path.insert(0, ROOT)
    import crystal_packing_b200  # noqa: F401
    from crystal_packing_b200 import distributed as cpd

    nan = float("nan")
    recs = [dict(score=nan, replica=0), dict(score=0.5, replica=9), dict(score=0.7, replica=5),
            dict(score=0.7, replica=3)]
    assert cpd.pick_best(recs)["replica"] == 5          # ties -> highest replica id (rayon's max)
    assert cpd.pick_best([dict(score=nan, replica=1)]) is None
    blob = cpd.pack_best([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], 0.25, 7, 8, 9)
    assert len(blob) == cpd.BEST_BYTES
    assert cpd.unpack_best(blob) == dict(dof=[1.0, 2.0, 3.0, 4.0, 5.0, 6.0], score=0.25, rejections=7,
                                         moves=8, replica=9)


@pytest.mark.timeout(300)
def test_two_ranks_agree_with_single_process(tmp_path):
    import orc

    total, world = 10, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    lines = [open(tmp_path / f"rank{r}.txt").read() for r in range(world)]
    assert lines[0] == lines[1]                          # every rank ends with the same best state
    state = orc.state_from_group(orc.polygon(4), "p2")
    bopt = orc.build_optimiser(steps=300, inner_steps=100)
    idx, best, _, _ = orc.analyse_state(state, bopt, 0, total, n_threads=2)
    replica, score = lines[0].split()[:2]
    assert int(replica) == idx and float.fromhex(score) == best.score
