#!/usr/bin/env python
"""bench.py — MC trial moves/s of the packing optimisation on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU path on the box's host cores

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): regular pentagon
(LineShape::polygon(5)) in plane group p2, 65,536 replicas per GPU, each through the three-stage
pipeline of the CLI's analyse_state (100 quick steps at kT = 0, `--mc-steps` annealing steps, the
same again at kT = 0; inner_steps 1000, max_step_size 0.01).  One step = one pass over that batch.
Replicas shard across ranks with no data-path collective ("scaling": "weak", 65,536 per GPU); the
one exchange is the 80-byte all-gather of each rank's best state.

  value     trial moves/s, whole job, template state and schedules already resident in HBM; timed
            on the device with CUDA events on the stream the kernels are launched on
  e2e       the same through cp_run with HOST buffers: per-replica initial DOFs copied from pinned
            memory each step and every replica's result copied back
  roofline  the MC kernel against the FP64 pipe (this path does no HBM streaming and no GEMM):
            achieved = reference-formula FLOPs per move (oracle counters, SURVEY.md §8d unit costs)
            x moves per launch / kernel time; peak = FP64 FMA throughput measured here
  cpu_baseline  the C restatement of the reference's rayon path on all host cores, bounded sample
                (oracle/libcp_oracle_fast.so: -O3 -march=native, operation counters compiled out)
  configs   the other BASELINE.json configurations, device-resident and device-timed like `value`:
            C3 (disc trimer in p2gg with annealing, 262,144 replicas/GPU, weak), C4 (LJ trimer in
            p2gg, 131,072 replicas/GPU = 1,048,576 on 8 GPUs, weak) and C5 (3..12-gons x 7 plane
            groups, 8,388,590 replicas in total, STRONG: the 70 cells are cut into chunks and dealt
            to the ranks by measured cost; per-rank kernel times are printed)
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "mc_trial_moves_per_s"
UNIT = "moves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=65536, help="replicas per GPU")
    ap.add_argument("--mc-steps", type=int, default=10000, help="stage-2/3 steps per replica")
    ap.add_argument("--inner-steps", type=int, default=1000)
    ap.add_argument("--sides", type=int, default=5)
    ap.add_argument("--group", default="p2")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="c3,c4,c5", help="extra BASELINE configurations to time (comma list, '' = none)")
    ap.add_argument("--config-steps", type=int, default=2, help="timed passes of each extra configuration")
    ap.add_argument("--c5-total", type=int, default=8388608, help="replicas of the C5 sweep (rounded down to 70 equal cells)")
    ap.add_argument("--c5-mc-steps", type=int, default=1000)
    ap.add_argument("--c5-granularity", type=int, default=6,
                    help="C5 cells costing more than 1/granularity of a rank's share are cut into chunks below that")
    ap.add_argument("--c5-streams", type=int, default=2, choices=[1, 2],
                    help="contexts per GPU the C5 chunks alternate over (2: cell tails overlap)")
    return ap.parse_args()


def moves_per_replica(args):
    inner = min(args.inner_steps, args.mc_steps)
    quick_inner = min(args.inner_steps, 100)
    return (100 // quick_inner) * quick_inner + 2 * (args.mc_steps // inner) * inner


def workload_config(args, world):
    return {
        "workload": f"polygon({args.sides}) in {args.group}, {args.replicas} replicas/GPU, 3-stage analyse_state "
                    f"pipeline (100 + {args.mc_steps} + {args.mc_steps} steps, inner_steps {args.inner_steps}, "
                    f"max_step_size 0.01, kt 0.1 x0.1/loop)",
        "replicas_per_gpu": args.replicas,
        "global_replicas": args.replicas * world,
        "moves_per_replica": moves_per_replica(args),
        "rng": "philox4x32-10 (production stream; PCG64-MCG replay is the parity mode)",
        "parallelism": f"replica-sharded x{world}, no data-path collective",
        "l2": "256 MiB write between timed steps (the kernel streams nothing: its working set is KBs of shared memory "
              "and registers, so L2 state does not matter; kept because the timing rules ask for it)",
    }


# ----------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_sample(args, seconds, threads):
    """Times the oracle (C restatement of the reference's rayon path) on `threads` host threads for
    about `seconds`; returns (moves/s, moves, wall, replicas).  The timed build is
    oracle/libcp_oracle_fast.so (-O3 -march=native, no operation counters), made on this machine."""
    import orc

    state = orc.state_from_group(orc.polygon(args.sides), args.group, orc.MATH_LIBM)
    bopt = orc.build_optimiser(steps=args.mc_steps, inner_steps=args.inner_steps)
    orc.lib_fast()  # build + load outside the timed region
    batch = threads
    total_moves, total_wall, done = 0, 0.0, 0
    cpu_sample.best = float("nan")
    while True:
        t0 = time.perf_counter()
        _, best, res = orc.analyse_state_fast(state, bopt, done, batch, n_threads=threads)
        dt = time.perf_counter() - t0
        if not (cpu_sample.best >= best.score):
            cpu_sample.best = best.score
        total_wall += dt
        total_moves += sum(r.moves for r in res)
        done += batch
        rate = done / total_wall
        if total_wall >= seconds * 0.6 or done >= 8192:
            break
        batch = max(threads, int(min(rate * (seconds - total_wall), 4 * done)) // threads * threads)
    return total_moves / total_wall, total_moves, total_wall, done


def counted_flops_per_move(args, threads, replicas=64):
    """Reference-formula operations per trial move (SURVEY.md 8d unit costs) on a small PCG sample,
    from the counting build of the oracle; untimed."""
    import orc

    state = orc.state_from_group(orc.polygon(args.sides), args.group, orc.MATH_LIBM)
    bopt = orc.build_optimiser(steps=args.mc_steps, inner_steps=args.inner_steps)
    _, _, res, cnt = orc.analyse_state(state, bopt, 0, replicas, n_threads=threads, want_counters=True)
    return orc.lib().orc_counters_flops(ctypes.byref(cnt)) / sum(r.moves for r in res)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(4.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(args, min(1.0, per_step), threads)
    walls, moves, reps = [], 0, 0
    for _ in range(args.steps):
        v, m, w, n = cpu_sample(args, per_step, threads)
        walls.append(w)
        moves += m
        reps += n
    value = moves / sum(walls)
    sample = f"{reps // max(1, args.steps)} replicas/step of the same workload (seeds from 0), full 3-stage schedule"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "best_packing_fraction": cpu_sample.best,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference's rayon path (Rust toolchain absent: "
                                 "the crate cannot be compiled here); gcc -O3 -march=native -ffp-contract=off, "
                                 "operation counters compiled out, pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name not in ("gpu_idle",):
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------- the B200 arm
EXECUTED_PROFILE = os.path.join(ROOT, "profiles", "r2_executed_flops.json")


class Bench:
    """One rank of the B200 arm: engine, launch stream, process group."""

    def __init__(self, args, rank, local_rank, world):
        import torch

        import crystal_packing_b200 as cp
        from crystal_packing_b200 import distributed as cpd

        self.torch, self.cp, self.cpd = torch, cp, cpd
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference)")
        torch.cuda.set_device(local_rank)
        self.dev = f"cuda:{local_rank}"
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
        self.eng = cp.Engine(local_rank, lanes=args.lanes)
        self.stream = torch.cuda.Stream(device=local_rank)
        self.eng.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        _ = self.best_view()

    def best_view(self):
        """The ctx's device cp_best (allocated once by cp_init) as a uint8 tensor."""
        cpd = self.cpd

        class BestView:
            def __init__(self, ptr):
                self.__cuda_array_interface__ = {"shape": (cpd.BEST_BYTES,), "typestr": "|u1", "data": (ptr, False),
                                                 "version": 2}

        if not hasattr(self, "_best_dev"):
            state = self.cp.PackedState.from_group(self.cp.LineShape.polygon(4), "p1")
            self.eng.prepare(state._problem, self.cp.BuildOptimiser.cli_defaults().pipeline(), 32)
            _, ptr = self.eng.device_buffers()
            self._best_dev = self.torch.as_tensor(BestView(ptr), device=self.dev)
        return self._best_dev

    def second_lane(self):
        """(engine, stream, device cp_best view) of a second context on this GPU"""
        if not hasattr(self, "_lane2"):
            torch, cp, cpd = self.torch, self.cp, self.cpd
            eng = cp.Engine(self.local_rank, lanes=self.args.lanes)
            stream = torch.cuda.Stream(device=self.local_rank)
            eng.set_stream(stream.cuda_stream)
            state = cp.PackedState.from_group(cp.LineShape.polygon(4), "p1")
            eng.prepare(state._problem, cp.BuildOptimiser.cli_defaults().pipeline(), 32)
            _, ptr = eng.device_buffers()

            class BestView:
                def __init__(self, ptr):
                    self.__cuda_array_interface__ = {"shape": (cpd.BEST_BYTES,), "typestr": "|u1",
                                                     "data": (ptr, False), "version": 2}

            self._lane2 = (eng, stream, torch.as_tensor(BestView(ptr), device=self.dev))
        return self._lane2

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(self, xs):
        """every rank's list of floats (same length on all ranks)"""
        if not self.dist:
            return [list(xs)]
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device=self.dev)
        out = self.torch.empty(self.world * len(xs), dtype=self.torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        v = out.cpu().tolist()
        return [v[r * len(xs):(r + 1) * len(xs)] for r in range(self.world)]

    def gather_best(self, local_u8):
        """all-gather of cp_best records on the launch stream: the stream waits for the collective, so
        the next launch (which overwrites the device cp_best) and the closing event are ordered after it"""
        if not self.dist:
            return local_u8
        torch = self.torch
        with torch.cuda.stream(self.stream):
            out = torch.empty(self.world * local_u8.numel(), dtype=torch.uint8, device=self.dev)
            work = self.dist.all_gather_into_tensor(out, local_u8, async_op=True)
            work.wait()  # the CURRENT stream (the launch stream) waits; the host does not
        return out

    def timed_passes(self, one_pass, warmup, steps, sample_clocks=False):
        """warmup untimed passes, then `steps` passes each bracketed by CUDA events on the launch
        stream (256 MiB L2 flush before the opening event), barrier + synchronize on both sides.
        Returns (per-pass ms on this rank, max-over-ranks total ms, wall s, clocks)."""
        torch = self.torch
        for _ in range(warmup):
            one_pass()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(physical_gpu_index(self.local_rank))
            sampler.start()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall0 = time.perf_counter()
        for a, b in ev:
            with torch.cuda.stream(self.stream):
                self.flush.fill_(1)
                a.record(self.stream)
            one_pass()
            b.record(self.stream)
            self.stream.synchronize()
        self.barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if sampler else None
        ms = [a.elapsed_time(b) for a, b in ev]
        return ms, self.max_over_ranks(sum(ms)), wall, clocks


def schedule_moves(schedules):
    return sum((s.steps // s.inner_steps) * s.inner_steps for s in schedules)


def run_weak_config(B, name, state, opts, per_gpu, note):
    """C3 / C4: `per_gpu` replicas on every rank, device-resident, timed like the headline."""
    cp, args = B.cp, B.args
    schedules = opts.pipeline()
    begin, _ = B.cpd.weak_range(per_gpu, B.rank)
    B.eng.prepare(state._problem, schedules, per_gpu, replica_begin=begin, rng=cp.RNG_PHILOX)
    best_dev = B.best_view()
    kernel_ms = []

    def one_pass():
        B.eng.launch()
        one_pass.gathered = B.gather_best(best_dev)

    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        one_pass()
    B.torch.cuda.synchronize()

    def timed():
        one_pass()
        B.stream.synchronize()
        kernel_ms.append(B.eng.timing_kernel_ms())

    ms, total_ms, _, _ = B.timed_passes(timed, 0, args.config_steps)
    moves = per_gpu * schedule_moves(schedules)
    gbest = B.cpd.global_best(one_pass.gathered)
    t = B.eng.timing()
    return {
        "workload": note, "scaling": "weak", "replicas_per_gpu": per_gpu, "global_replicas": per_gpu * B.world,
        "moves_per_replica": schedule_moves(schedules), "value": B.world * moves * args.config_steps / (total_ms * 1e-3),
        "unit": UNIT, "ms_per_step": total_ms / args.config_steps, "kernel_ms_rank0": sum(kernel_ms) / len(kernel_ms),
        "steps": args.config_steps, "warmup": warm, "gpu_launches": 3 * args.config_steps,
        "lanes_per_replica": t.lanes_per_replica, "best_score": gbest["score"], "best_replica": gbest["replica"],
    }


def run_c5(B):
    """BASELINE.json configs[4]: 3..12-gons x the seven plane groups, a FIXED batch split over the
    ranks (strong scaling).  Cells differ ~18x in cost per replica, so the cells are cut into chunks
    and the chunks are dealt longest-first by MEASURED time: the first warm-up pass is the
    calibration (chunk i timed by rank i mod N, times all-gathered), every later pass uses the
    balanced assignment.  No data-path collective; one all-gather of the per-chunk best states."""
    cp, cpd, args, torch = B.cp, B.cpd, B.args, B.torch
    cells = cpd.sweep_cells()
    per_cell = args.c5_total // len(cells)
    opts = cp.BuildOptimiser.cli_defaults(steps=args.c5_mc_steps, inner_steps=min(1000, args.c5_mc_steps))
    schedules = opts.pipeline()
    mpr = schedule_moves(schedules)
    states = [cp.PackedState.from_group(cp.LineShape.polygon(n), g) for n, g in cells]
    est = [cpd.estimated_cell_cost(n, st.total_shapes()) for (n, g), st in zip(cells, states)]
    chunks = cpd.sweep_chunks(est, per_cell, B.world, granularity=args.c5_granularity)
    best_dev = B.best_view()
    slots = len(chunks)  # a rank never holds more chunks than there are
    nan_rec = torch.frombuffer(bytearray(cpd.pack_best([0.0] * 6, float("nan"), 0, 0, 2 ** 64 - 1)), dtype=torch.uint8)
    template = nan_rec.to(B.dev).repeat(slots)
    torch.cuda.synchronize()

    # Two contexts, two streams on this GPU: the chunks alternate between them, so the tail of one
    # cell's kernel (its last, partly filled wave of blocks) overlaps the start of the next cell's.
    lanes = [(B.eng, B.stream, best_dev)] + ([B.second_lane()] if args.c5_streams > 1 else [])

    def run_chunks(mine, timed_each):
        """launches this rank's chunks, alternating over the lanes (one lane, back to back, when each
        chunk is timed on its own); keeps each chunk's cp_best in its slot"""
        with torch.cuda.stream(B.stream):
            bests = template.clone()
            ready = torch.cuda.Event()
            ready.record(B.stream)
        times = {}
        use = lanes[:1] if timed_each else lanes
        for _, st, _ in use[1:]:
            st.wait_event(ready)
        for k, ci in enumerate(mine):
            c, begin, count = chunks[ci]
            eng, st, bdev = use[k % len(use)]
            eng.prepare(states[c]._problem, schedules, count, replica_begin=begin, rng=cp.RNG_PHILOX)
            eng.launch()
            with torch.cuda.stream(st):
                bests[k * cpd.BEST_BYTES:(k + 1) * cpd.BEST_BYTES].copy_(bdev)
            if timed_each:
                st.synchronize()
                times[ci] = eng.timing_kernel_ms()
        for _, st, _ in use[1:]:  # the launch stream closes the pass: it waits for the other lane
            done = torch.cuda.Event()
            done.record(st)
            B.stream.wait_event(done)
        return bests, times

    # calibration = warm-up pass 1 (also loads every kernel build this rank will launch)
    cal = cpd.round_robin_assign(len(chunks), B.world)
    _, times = run_chunks(cal[B.rank], True)
    mine_t = [times.get(i, 0.0) for i in range(len(chunks))]
    measured = [max(col) for col in zip(*B.gather_floats(mine_t))]  # each chunk was timed by exactly one rank
    assign, est_loads = cpd.lpt_assign(measured, B.world)

    def load_builds(mine):  # the kernel builds of an assignment, outside any timed region
        for ci in mine:
            c, begin, _ = chunks[ci]
            for eng, _, _ in lanes:
                eng.prepare(states[c]._problem, schedules, 64, replica_begin=begin, rng=cp.RNG_PHILOX)
                eng.launch()
        B.torch.cuda.synchronize()

    def rank_times(mine):
        """one untimed-for-the-bench pass of an assignment as it will run (both lanes): every rank's time"""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        B.barrier()
        a.record(B.stream)
        run_chunks(mine, False)
        b.record(B.stream)
        B.stream.synchronize()
        return [v[0] for v in B.gather_floats([a.elapsed_time(b)])]

    load_builds(assign[B.rank])
    refinements = []
    if B.world > 1:
        # The chunk times were measured one kernel at a time; with two lanes the kernels of a rank
        # overlap, and how much that helps depends on which cells meet.  Refinement: run the
        # assignment as it will run, scale every chunk's cost by its rank's measured / planned time,
        # deal again (four rounds); keep the assignment with the smallest measured maximum.  Every rank computes the
        # same assignment from the same all-gathered numbers.
        best_assign, best_t = assign, rank_times(assign[B.rank])
        refinements.append(max(best_t))
        cur_assign, cur_t, eff = assign, best_t, list(measured)
        for _ in range(4):
            planned = [sum(eff[i] for i in cur_assign[r]) for r in range(B.world)]
            for r in range(B.world):
                for i in cur_assign[r]:
                    eff[i] *= cur_t[r] / planned[r] if planned[r] > 0 else 1.0
            cur_assign, est = cpd.lpt_assign(eff, B.world)
            load_builds(cur_assign[B.rank])
            cur_t = rank_times(cur_assign[B.rank])
            refinements.append(max(cur_t))
            if max(cur_t) < max(best_t):
                best_assign, best_t, est_loads = cur_assign, cur_t, est
        assign = best_assign
    mine = assign[B.rank]
    warm = max(1, min(args.warmup, 3) - 1)
    state = {}

    marks = []

    def one_pass():
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record(B.stream)
        bests, _ = run_chunks(mine, False)
        b.record(B.stream)  # this rank's own work ends here; the all-gather below waits for the slowest rank
        marks.append((a, b))
        state["gathered"] = B.gather_best(bests)

    ms, total_ms, _, _ = B.timed_passes(one_pass, warm, args.config_steps)
    own = [a.elapsed_time(b) for a, b in marks[-args.config_steps:]]
    per_rank = [v[0] for v in B.gather_floats([sum(own) / len(own)])]
    # per-cell best over all chunks of all ranks
    raw = state["gathered"].cpu().numpy().tobytes()
    recs = [cpd.unpack_best(raw[i:i + cpd.BEST_BYTES]) for i in range(0, len(raw), cpd.BEST_BYTES)]
    cell_best = {}
    for r in range(B.world):
        for k, ci in enumerate(assign[r]):
            rec = recs[r * slots + k]
            c = chunks[ci][0]
            if rec["score"] == rec["score"] and (c not in cell_best or rec["score"] > cell_best[c]["score"] or (
                    rec["score"] == cell_best[c]["score"] and rec["replica"] > cell_best[c]["replica"])):
                cell_best[c] = rec
    total_moves = per_cell * len(cells) * mpr
    heavy = sorted(range(len(cells)), key=lambda c: -sum(measured[i] for i, ch in enumerate(chunks) if ch[0] == c))[:3]
    cell_ms = {c: sum(measured[i] for i, ch in enumerate(chunks) if ch[0] == c) for c in range(len(cells))}
    return {
        "workload": f"polygon(3..12) x {{{','.join(cpd.SWEEP_GROUPS)}}}: 70 cells x {per_cell} replicas, 3-stage pipeline "
                    f"(100 + {args.c5_mc_steps} + {args.c5_mc_steps} steps)",
        "scaling": "strong", "global_replicas": per_cell * len(cells), "moves_per_replica": mpr,
        "value": total_moves * args.config_steps / (total_ms * 1e-3), "unit": UNIT,
        "wall_ms_per_pass": total_ms / args.config_steps, "per_rank_ms": per_rank,
        "per_rank_ms_note": "each rank's own chunks, first launch to last kernel, before the all-gather",
        "rank_spread": (max(per_rank) - min(per_rank)) / max(per_rank),
        "chunks": len(chunks), "chunks_per_rank": [len(a) for a in assign],
        "balance": "longest-first by measured chunk time (calibration = first warm-up pass), refined up to four times by the "
                   "ranks' measured times with both lanes running",
        "refinement_max_rank_ms": refinements,
        "streams_per_gpu": len(lanes),
        "planned_loads_ms": est_loads, "steps": args.config_steps, "warmup": warm + 1,
        "gpu_launches": 3 * len(mine) * args.config_steps,
        "limit": "heaviest cells: " + ", ".join(
            f"{cells[c][0]}-gon/{cells[c][1]} {cell_ms[c]:.0f} ms" for c in heavy) +
            f" of {sum(cell_ms.values()):.0f} ms summed over all cells",
        "cells_with_a_packing": len(cell_best),
        "best_of_sweep": max((v["score"] for v in cell_best.values()), default=None),
    }


def run_b200(args, rank, local_rank, world):
    import torch

    B = Bench(args, rank, local_rank, world)
    cp, cpd, capi, eng, stream = B.cp, B.cpd, B.cp.capi, B.eng, B.stream

    state = cp.PackedState.from_group(cp.LineShape.polygon(args.sides), args.group)
    schedules = cp.BuildOptimiser.cli_defaults(steps=args.mc_steps, inner_steps=args.inner_steps).pipeline()
    n = args.replicas
    begin, _ = cpd.weak_range(n, rank)
    mpr = moves_per_replica(args)
    moves_per_step_rank = n * mpr
    fp64_peak, fp32_peak = eng.measure_peaks()

    # ---- device-resident metric ("value")
    eng.prepare(state._problem, schedules, n, replica_begin=begin, rng=cp.RNG_PHILOX)
    best_dev = B.best_view()
    hold = {}
    kernel_ms = []

    def one_step_resident():
        eng.launch()
        hold["gathered"] = B.gather_best(best_dev)

    for _ in range(args.warmup):
        one_step_resident()

    def timed():
        one_step_resident()
        stream.synchronize()
        kernel_ms.append(eng.timing_kernel_ms())

    step_ms, total_ms, wall, clocks = B.timed_passes(timed, 0, args.steps, sample_clocks=True)
    best_local = eng.fetch()
    lanes = eng.timing().lanes_per_replica
    if world > 1:
        gbest = cpd.global_best(hold["gathered"])
    else:
        gbest = {"score": best_local.result.score, "replica": best_local.replica,
                 "dof": list(best_local.result.dof)}
    value = world * moves_per_step_rank * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through cp_run with host buffers ("e2e")
    e2e = None
    if not args.no_e2e:
        init = torch.empty((n, 6), dtype=torch.float64).pin_memory()
        init[:] = torch.tensor(state.dof, dtype=torch.float64)
        out = torch.empty(n * ctypes.sizeof(capi.ReplicaResult), dtype=torch.uint8).pin_memory()
        h2d, d2h = init.numel() * 8, out.numel() + cpd.BEST_BYTES + 4

        def one_step_e2e():
            b = eng.run_host_buffers(state._problem, schedules, n, init.data_ptr(), out.data_ptr(),
                                     replica_begin=begin, rng=cp.RNG_PHILOX)
            if world > 1:
                mine = torch.frombuffer(bytearray(bytes(b)), dtype=torch.uint8).to(B.dev)
                cpd.global_best(cpd.all_gather_best(mine))
            return b

        for _ in range(min(args.warmup, 1)):
            one_step_e2e()
        B.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_step_e2e()
        torch.cuda.synchronize()
        e2e_s = B.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * moves_per_step_rank * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "api": "cp_run(ctx, problem, stages, 3, PHILOX, ..., init_dof=pinned host, out_all=pinned host)"}

    # ---- the other BASELINE configurations
    configs = {}
    wanted = [c.strip().lower() for c in args.configs.split(",") if c.strip()]
    if "c3" in wanted:
        st3 = cp.PackedState.from_group(cp.MolecularShape2.from_trimer(0.637556, 120.0, 1.0), "p2gg")
        o3 = cp.BuildOptimiser.cli_defaults(steps=args.mc_steps, inner_steps=args.inner_steps, kt_finish=0.001)
        configs["C3"] = run_weak_config(B, "C3", st3, o3, 262144,
                                        f"MolecularShape2 trimer(0.637556, 120, 1) in p2gg, 262144 replicas/GPU, annealed "
                                        f"(kt 0.1 -> 0.001 over {args.mc_steps} steps), 3-stage pipeline")
    if "c4" in wanted:
        st4 = cp.PotentialState.from_group(cp.LJShape2.from_trimer(0.637556, 120.0, 1.0), "p2gg")
        o4 = cp.BuildOptimiser.cli_defaults(steps=args.mc_steps, inner_steps=args.inner_steps)
        configs["C4"] = run_weak_config(B, "C4", st4, o4, 131072,
                                        f"LJShape2 trimer in p2gg (PotentialState, energy minimisation), 131072 replicas/GPU "
                                        f"(1,048,576 on 8 GPUs), 3-stage pipeline 100 + {args.mc_steps} + {args.mc_steps}")
    if "c5" in wanted:
        configs["C5"] = run_c5(B)

    # ---- CPU baseline + algorithmic FLOPs per move (rank 0, N=1 only for the baseline figure)
    cpu, flops_per_move = None, None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        flops_per_move = counted_flops_per_move(args, threads)
        if world == 1:
            v, m, w, reps = cpu_sample(args, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{reps} replicas (seeds 0..{reps - 1}) of the same workload, full 3-stage schedule, "
                             f"{w:.1f} s wall",
                   "best_packing_fraction": cpu_sample.best,
                   "note": "C restatement of the reference's rayon path (gcc -O3 -march=native, -ffp-contract=off, no "
                           "operation counters); the Rust crate cannot be built here; its best packing is over the "
                           "sample's replicas only (PCG stream)"}
    roofline = None
    if rank == 0:
        k_ms = sum(kernel_ms) / len(kernel_ms)
        if flops_per_move is not None:
            achieved = flops_per_move * moves_per_step_rank / (k_ms * 1e-3) / 1e12
            roofline = {
                "bound": "fp64",
                "kernel": ("mc_thread_kernel<POLYGON,%d,PHILOX>" % (args.sides if 3 <= args.sides <= 12 else 0))
                if lanes == 1 else "mc_kernel<POLYGON,%d,PHILOX>" % lanes,
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                "peak_source": "measured here: FP64 FMA chains, 2 flops/FMA (MEASURED_PEAKS.json has no FP64 entry)",
                "fp32_fma_peak_tflops": fp32_peak, "frac_of_fp32_peak": achieved / fp32_peak if fp32_peak else None,
                "flops_per_move": flops_per_move,
                "flops_definition": "REFERENCE-EQUIVALENT work: operations the reference's formulas execute per trial "
                                    "move, counted by the oracle (SURVEY.md 8d unit costs) - not the instructions this "
                                    "kernel issues (it prunes the image sweep, decides most shape pairs with f32 "
                                    "certificates and skips check_intersection for moves the Metropolis rule rejects)",
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (sum(step_ms) / len(step_ms)),
                "moves_per_launch": moves_per_step_rank,
                "hbm_note": "no HBM streaming on this path: the algorithmic bytes of a launch are its results "
                            f"({n * ctypes.sizeof(capi.ReplicaResult)} B written once) and, in e2e, the initial DOFs",
            }
            # what the kernel actually executes, from the committed ncu capture of this same command
            traffic = None
            try:
                prof = json.load(open(EXECUTED_PROFILE))
                key = f"polygon({args.sides})/{args.group}/{args.replicas}/{mpr}"
                ex = prof.get(key)
                if ex and lanes == 1:
                    f64, f32 = ex["fp64_flops_per_move"], ex["fp32_flops_per_move"]
                    moves_s = moves_per_step_rank / (k_ms * 1e-3)
                    roofline.update({
                        "executed_fp64_flops_per_move": f64, "executed_fp32_flops_per_move": f32,
                        "frac_executed": f64 * moves_s / 1e12 / fp64_peak if fp64_peak else None,
                        "frac_executed_fp32": f32 * moves_s / 1e12 / fp32_peak if fp32_peak else None,
                        "issue_slots_busy": ex["issue_active_pct"] / 100.0,
                        "warp_instructions_per_move": ex["warp_instructions_per_move"],
                        "active_lanes_per_instruction": ex["thread_inst_per_inst"],
                        "from_profile": ex["from_profile"],
                        "executed_note": "per-move instruction counts of the ncu capture x this run's moves/s; the kernel "
                                         "is bound by issue slots and fixed-latency dependencies at 4 warps per "
                                         "scheduler, not by an FP pipe",
                    })
                    traffic = ex["dram_bytes_per_launch"]
            except (OSError, ValueError, KeyError):
                pass
            roofline["traffic"] = traffic
            roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload, "
                                        "from the profile named in from_profile" if traffic is not None else
                                        "no ncu capture committed for this workload")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": 3 * args.steps, "lanes_per_replica": lanes,
            "roofline": roofline, "cpu_baseline": cpu, "configs": configs,
            "best_packing_fraction": gbest["score"], "best_replica": gbest["replica"],
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line), flush=True)
    if B.dist:
        B.dist.barrier()
        B.dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
