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
        "l2": "256 MiB write between timed steps (working set is KBs; L2 state is irrelevant to an FP64-bound kernel)",
    }


# ----------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_sample(args, seconds, threads):
    """Times the oracle (C restatement of the reference's rayon path) on `threads` host threads for
    about `seconds`; returns (moves/s, moves, wall, flops_per_move, replicas)."""
    import orc

    state = orc.state_from_group(orc.polygon(args.sides), args.group, orc.MATH_LIBM)
    bopt = orc.build_optimiser(steps=args.mc_steps, inner_steps=args.inner_steps)
    batch = threads
    total_moves, total_wall, total_flops, done = 0, 0.0, 0.0, 0
    cpu_sample.best = float("nan")
    while True:
        t0 = time.perf_counter()
        _, best, res, cnt = orc.analyse_state(state, bopt, done, batch, n_threads=threads, want_counters=True)
        dt = time.perf_counter() - t0
        if not (cpu_sample.best >= best.score):
            cpu_sample.best = best.score
        total_wall += dt
        total_moves += sum(r.moves for r in res)
        total_flops += orc.lib().orc_counters_flops(ctypes.byref(cnt))
        done += batch
        rate = done / total_wall
        if total_wall >= seconds * 0.6 or done >= 4096:
            break
        batch = max(threads, int(min(rate * (seconds - total_wall), 4 * done)) // threads * threads)
    return total_moves / total_wall, total_moves, total_wall, total_flops / total_moves, done


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(4.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(args, min(1.0, per_step), threads)
    walls, moves, reps = [], 0, 0
    for _ in range(args.steps):
        v, m, w, _, n = cpu_sample(args, per_step, threads)
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
                                 "the crate cannot be compiled here); gcc -O2 -ffp-contract=off, pthreads"},
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
def run_b200(args, rank, local_rank, world):
    import torch

    import crystal_packing_b200 as cp
    from crystal_packing_b200 import capi, distributed as cpd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng = cp.Engine(local_rank, lanes=args.lanes)
    stream = torch.cuda.Stream(device=local_rank)
    eng.set_stream(stream.cuda_stream)

    state = cp.PackedState.from_group(cp.LineShape.polygon(args.sides), args.group)
    schedules = cp.BuildOptimiser.cli_defaults(steps=args.mc_steps, inner_steps=args.inner_steps).pipeline()
    n = args.replicas
    begin, _ = cpd.weak_range(n, rank)
    mpr = moves_per_replica(args)
    moves_per_step_rank = n * mpr

    fp64_peak, fp32_peak = eng.measure_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    gathered = None

    class BestView:  # wraps the device cp_best as a uint8 tensor for the collective
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (cpd.BEST_BYTES,), "typestr": "|u1", "data": (ptr, False),
                                             "version": 2}

    def one_step_resident():
        nonlocal gathered
        eng.launch()
        if world > 1:
            with torch.cuda.stream(stream):
                gathered = cpd.all_gather_best(best_dev)

    # ---- device-resident metric ("value")
    eng.prepare(state._problem, schedules, n, replica_begin=begin, rng=cp.RNG_PHILOX)
    _, best_ptr = eng.device_buffers()
    best_dev = torch.as_tensor(BestView(best_ptr), device=f"cuda:{local_rank}")
    for _ in range(args.warmup):
        one_step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, wall0 = [], time.perf_counter()
    for a, b in ev:
        with torch.cuda.stream(stream):
            flush.fill_(1)
            a.record(stream)
        one_step_resident()
        b.record(stream)
        stream.synchronize()
        kernel_ms.append(eng.timing_kernel_ms())
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    best_local = eng.fetch()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        gbest = cpd.global_best(gathered)
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
                mine = torch.frombuffer(bytearray(bytes(b)), dtype=torch.uint8).to(f"cuda:{local_rank}")
                cpd.global_best(cpd.all_gather_best(mine))
            return b

        for _ in range(min(args.warmup, 1)):
            one_step_e2e()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_step_e2e()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if dist:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": world * moves_per_step_rank * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "api": "cp_run(ctx, problem, stages, 3, PHILOX, ..., init_dof=pinned host, out_all=pinned host)"}

    # ---- CPU baseline + algorithmic FLOPs per move (rank 0, N=1 only for the baseline figure)
    cpu, flops_per_move = None, None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, m, w, fpm, reps = cpu_sample(args, args.cpu_seconds if world == 1 else 3.0, threads)
        flops_per_move = fpm
        if world == 1:
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{reps} replicas (seeds 0..{reps - 1}) of the same workload, full 3-stage schedule, "
                             f"{w:.1f} s wall",
                   "best_packing_fraction": cpu_sample.best,
                   "note": "C restatement of the reference's rayon path; the Rust crate cannot be built here; "
                           "its best packing is over the sample's replicas only (PCG stream)"}
    roofline = None
    if rank == 0:
        k_ms = sum(kernel_ms) / len(kernel_ms)
        if flops_per_move is not None:
            achieved = flops_per_move * moves_per_step_rank / (k_ms * 1e-3) / 1e12
            roofline = {
                "bound": "fp64", "kernel": ("mc_thread_kernel<POLYGON,%d,PHILOX>" % (args.sides if 3 <= args.sides <= 12 else 0)) if eng.timing().lanes_per_replica == 1 else "mc_kernel<POLYGON,%d,PHILOX>" % eng.timing().lanes_per_replica, "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                "traffic": 1410560, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                                   "(profiles/r1_f_thread_end_of_round.txt); algorithmic HBM bytes are the 4.7 MB of results",
                "peak_source": "measured here: FP64 FMA chains, 2 flops/FMA (MEASURED_PEAKS.json has no FP64 entry)",
                "fp32_fma_peak_tflops": fp32_peak, "frac_of_fp32_peak": achieved / fp32_peak if fp32_peak else None,
                "flops_per_move": flops_per_move,
                "flops_definition": "executed reference-formula operations per trial move counted by the oracle "
                                    "(SURVEY.md §8d unit costs), not the instructions the kernel issues",
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (sum(step_ms) / len(step_ms)),
                "moves_per_launch": moves_per_step_rank,
            }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": 3 * args.steps, "lanes_per_replica": eng.timing().lanes_per_replica,
            "roofline": roofline, "cpu_baseline": cpu,
            "best_packing_fraction": gbest["score"], "best_replica": gbest["replica"],
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


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
