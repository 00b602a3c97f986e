"""Multi-GPU plumbing: replicas are independent, so they shard across ranks with no data-path
collective (the reference's only cross-replica operation is the final `.max()`,
crystal-packing-cli/src/main.rs:253).  One process per GPU; the single exchange is an all-gather of
each rank's best replica (80 bytes per rank) followed by the same argmax on every rank.

Works with backend "nccl" on GPUs and "gloo" on CPU tensors (used by the world_size-2 tests).
"""
import struct

BEST_BYTES = 80  # sizeof(cp_best): dof[6] f64, score f64, rejections u64, moves u64, replica u64
_FMT = "<7d3Q"


def shard_range(total, rank, world):
    """Contiguous global replica-id range [begin, begin+count) of `rank`; seeds are global ids, so
    results do not depend on the number of ranks."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def weak_range(per_rank, rank):
    return rank * per_rank, per_rank


# ---------------------------------------------------------------------------------- sweeps of cells
# BASELINE.json configs[4]: every (polygon, plane group) cell is its own batch of replicas, and a
# replica of a 12-gon in a multiplicity-4 group costs ~18x one of a triangle in p1.  A fixed batch
# is therefore split over the ranks by COST, not by count: the cells are cut into chunks of
# replicas and the chunks dealt out longest-first.
SWEEP_GROUPS = ("p1", "p2", "p1m1", "p1g1", "p2mm", "p2mg", "p2gg")


def sweep_cells(sides=range(3, 13), groups=SWEEP_GROUPS):
    return [(n, g) for n in sides for g in groups]


def split_chunks(n_cells, replicas_per_cell, chunks_per_cell):
    """(cell, replica_begin, count) pieces of every cell; seeds stay the replica's index within its
    cell, so the results do not depend on how the cells were cut."""
    out = []
    for c in range(n_cells):
        for k in range(chunks_per_cell):
            begin, count = shard_range(replicas_per_cell, k, chunks_per_cell)
            if count:
                out.append((c, begin, count))
    return out


def lpt_assign(costs, world):
    """Longest-processing-time-first: chunk indices per rank and the estimated load of each rank.
    Deterministic (ties by index), so every rank computes the same assignment from the same costs."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    mine = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        mine[r].append(i)
        loads[r] += costs[i]
    return mine, loads


def estimated_cell_cost(sides, n_syms):
    """Relative cost of one replica of a cell, from the one-GPU table (profiles/r2_c5_sweep.txt):
    only decides into how many chunks a cell is cut; the assignment itself uses measured times."""
    return {1: 0.05, 2: 0.075}.get(n_syms, 0.24) * (1.0 + 0.2 * (sides - 3))


def sweep_chunks(cell_costs, replicas_per_cell, world, granularity=6, min_chunk=16384):
    """Cuts the cells into chunks (cell, replica_begin, count): a cell whose estimated cost exceeds
    1/granularity of one rank's share of the whole sweep is cut into pieces below that size (never
    below `min_chunk` replicas: a launch that does not fill the GPU wastes it).  world = 1 cuts
    nothing.  Seeds are the replica's index within its cell, so the cut does not change any result."""
    total = sum(cell_costs)
    target = total / (world * granularity)
    out = []
    for c, cost in enumerate(cell_costs):
        k = 1
        if world > 1 and cost > target:
            k = min(-(-cost // target), max(1, replicas_per_cell // min_chunk))
        k = int(k)
        for j in range(k):
            begin, count = shard_range(replicas_per_cell, j, k)
            if count:
                out.append((c, begin, count))
    return out


def round_robin_assign(n_items, world):
    """Calibration pass: chunk i is measured by rank i % world."""
    return [list(range(r, n_items, world)) for r in range(world)]


def pack_best(dof, score, rejections, moves, replica):
    return struct.pack(_FMT, *dof, score, rejections, moves, replica)


def unpack_best(buf):
    v = struct.unpack(_FMT, bytes(buf))
    return {"dof": list(v[:6]), "score": v[6], "rejections": v[7], "moves": v[8], "replica": v[9]}


def pick_best(records):
    """argmax by score, NaN (None) never wins, ties -> HIGHEST replica id: rayon's max() is
    reduce_with(Ord::max) in index order and Ord::max returns its second argument on Equal
    (main.rs:253), so the reference keeps the last of several maximal replicas.  Independent of the
    sharding."""
    best = None
    for r in records:
        if r["score"] != r["score"]:
            continue
        if best is None or r["score"] > best["score"] or (
                r["score"] == best["score"] and r["replica"] > best["replica"]):
            best = r
    return best


def all_gather_best(local_best_u8, group=None):
    """local_best_u8: uint8 tensor of BEST_BYTES (on the GPU for nccl, CPU for gloo).  Returns the
    gathered uint8 tensor [world * BEST_BYTES] on the same device."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty(world * BEST_BYTES, dtype=torch.uint8, device=local_best_u8.device)
    dist.all_gather_into_tensor(out, local_best_u8.contiguous(), group=group)
    return out


def global_best(gathered_u8):
    raw = gathered_u8.cpu().numpy().tobytes()
    return pick_best([unpack_best(raw[i:i + BEST_BYTES]) for i in range(0, len(raw), BEST_BYTES)])
