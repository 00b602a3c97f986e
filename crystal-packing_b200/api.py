"""Host-side mirror of the reference crate's interface for the Monte-Carlo packing path.

Names, argument order and error behaviour follow the Rust crate so that tests read like the
reference's own (citations are relative to the reference tree):

  LineShape::polygon / from_radial            crystal-packing/src/shape/line_shape.rs:128-150
  MolecularShape2::from_trimer / circle       crystal-packing/src/shape/molecular_shape2.rs:141-172
  LJShape2::from_trimer / circle              crystal-packing/src/shape/lj_shape.rs:111-151
  WallpaperGroups (FromStr) / WallpaperGroup  crystal-packing/src/wallpaper.rs:87-176
  PackedState / PotentialState ::from_group   crystal-packing/src/state/{packed,potential}.rs
  State::{score,total_shapes,generate_basis}  crystal-packing/src/traits.rs:71-87
  MCOptimiser::new(..).optimise_state(state)  crystal-packing/src/optimisation.rs:25-33,80-180
  BuildOptimiser (+ build)                    crystal-packing-cli/src/main.rs:27-143
  monte_carlo_best_packing                    GPU drop-in for analyse_state, main.rs:215-269

All arithmetic on the Monte-Carlo path runs in the CUDA kernels behind the C ABI (capi.py); the
host constructors (shape geometry, group table, schedule derivation) are the C++ ones exported by
the same library.
"""
import ctypes as C
import math

from . import capi
from .capi import (CP_DISCS, CP_LJ, CP_POLYGON, library_path, load_library)  # noqa: F401

RNG_PCG64MCG, RNG_PHILOX = capi.CP_RNG_PCG64MCG, capi.CP_RNG_PHILOX
NAN = float("nan")


class CrystalPackingError(RuntimeError):
    """anyhow::Error at the reference's API edges; carries the cp_status code."""

    def __init__(self, status, message):
        super().__init__(f"[cp_status {status}] {message}")
        self.status = status


def _check(status):
    if status != capi.CP_OK:
        raise CrystalPackingError(status, load_library().cp_last_error().decode())


def _opt(v):
    return NAN if v is None else float(v)


# ---------------------------------------------------------------------------------------- shapes
class _Shape:
    kind = None

    def __init__(self, name, problem):
        self.name = name
        self._problem = problem

    @property
    def items(self):
        p = self._problem
        return [tuple(p.items[i][j] for j in range(5)) for i in range(p.n_items)]

    def area(self):
        return self._problem.area

    def enclosing_radius(self):
        return self._problem.enclosing_radius


class LineShape(_Shape):
    kind = CP_POLYGON

    @classmethod
    def from_radial(cls, name, points):
        p = capi.Problem()
        arr = (C.c_double * len(points))(*points)
        _check(load_library().cp_shape_polygon(arr, len(points), C.byref(p)))
        return cls(name, p)

    @classmethod
    def polygon(cls, sides):
        return cls.from_radial("Polygon", [1.0] * sides)


class MolecularShape2(_Shape):
    kind = CP_DISCS

    @classmethod
    def from_trimer(cls, radius, angle, distance):
        p = capi.Problem()
        _check(load_library().cp_shape_trimer(CP_DISCS, radius, angle, distance, C.byref(p)))
        return cls("Trimer", p)

    @classmethod
    def circle(cls):
        p = capi.Problem()
        _check(load_library().cp_shape_circle(CP_DISCS, C.byref(p)))
        return cls("circle", p)


class LJShape2(_Shape):
    kind = CP_LJ

    @classmethod
    def from_trimer(cls, radius, angle, distance):
        p = capi.Problem()
        _check(load_library().cp_shape_trimer(CP_LJ, radius, angle, distance, C.byref(p)))
        return cls("Trimer", p)

    @classmethod
    def circle(cls):
        p = capi.Problem()
        _check(load_library().cp_shape_circle(CP_LJ, C.byref(p)))
        return cls("circle", p)


def refresh_shape(shape):
    """Recompute area() and enclosing_radius() after the items were edited by hand."""
    _check(load_library().cp_shape_finish(C.byref(shape._problem)))
    return shape


# ---------------------------------------------------------------------------------------- symmetry
class Transform2:
    """Rows 0-1 of a symmetry operator (Transform2::from_operations, transform.rs:124-183)."""

    def __init__(self, rows):
        self.rows = rows

    @classmethod
    def from_operations(cls, sym_ops):
        out = (C.c_double * 6)()
        _check(load_library().cp_parse_operations(sym_ops.encode(), out))
        return cls([list(out[0:3]), list(out[3:6]), [0.0, 0.0, 0.0]])


class WallpaperGroup:
    def __init__(self, key, family, symmetries):
        self.key = key  # the string accepted by WallpaperGroups::from_str
        self.family = family
        self.symmetries = symmetries

    @property
    def name(self):
        # the reference labels p1g1 as "p1m1" (wallpaper.rs:154-158)
        if self.key is None:
            return getattr(self, "custom_name", "custom")
        return "p1m1" if self.key in ("p1g1", "pg") else self.key


_FAMILY_NAMES = {"Monoclinic": capi.CP_MONOCLINIC, "Orthorhombic": capi.CP_ORTHORHOMBIC,
                 "Hexagonal": capi.CP_HEXAGONAL, "Tetragonal": capi.CP_TETRAGONAL}


def group_from_symmetries(name, family, matrices):
    """WallpaperGroup from serialised data: `matrices` are 3x3 column-major lists (the serde form of
    Transform2).  Matches the built-in table when it can (so `key` round-trips), else keeps the
    operators as given."""
    rows = [[m[0], m[3], m[6], m[1], m[4], m[7]] for m in matrices]
    for key in ("p1", "p2", "p1m1", "p1g1", "p2mm", "p2mg", "p2gg"):
        g = WallpaperGroups.from_str(key)
        if g.symmetries == rows and g.name == name:
            return g
    g = WallpaperGroup(None, _FAMILY_NAMES[family], rows)
    g.custom_name = name
    return g


class WallpaperGroups:
    """The enum + FromStr (wallpaper.rs:87-132)."""

    @staticmethod
    def variants():
        return ["p1", "p2", "p2m1", "p1g1", "p2mm", "p2mg", "p2gg"]  # as advertised, quirk included

    @staticmethod
    def from_str(s):
        fam, n = C.c_int32(), C.c_int32()
        syms = ((C.c_double * 6) * capi.MAX_SYMS)()
        _check(load_library().cp_group_lookup(s.encode(), C.byref(fam), C.byref(n), C.byref(syms)))
        return WallpaperGroup(s, fam.value, [[syms[i][j] for j in range(6)] for i in range(n.value)])


# ---------------------------------------------------------------------------------------- engine
class Engine:
    """One cp_ctx: one GPU, one stream.  `device=None` picks LOCAL_RANK (one process per GPU)."""

    def __init__(self, device=None, lanes=0):
        import os

        self._lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self._ctx = C.c_void_p()
        _check(self._lib.cp_init(device, C.byref(self._ctx)))
        if lanes:
            self.set_lanes(lanes)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.cp_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_lanes(self, lanes):
        _check(self._lib.cp_set_lanes(self._ctx, lanes))

    def set_stream(self, cuda_stream_handle):
        _check(self._lib.cp_set_stream(self._ctx, C.c_void_p(cuda_stream_handle)))

    # -- State::score for many states at once
    def score(self, problem, dofs):
        n = len(dofs)
        flat = (C.c_double * (6 * n))(*[v for d in dofs for v in d])
        out = (C.c_double * n)()
        _check(self._lib.cp_score(self._ctx, C.byref(problem), flat, n, out))
        return list(out)

    def run(self, problem, schedules, n_replicas, replica_begin=0, seed_base=0, rng=RNG_PCG64MCG,
            init_dof=None, want_all=True):
        sched = (capi.Schedule * len(schedules))(*schedules)
        init = None
        if init_dof is not None:
            init = (C.c_double * (6 * n_replicas))(*[v for d in init_dof for v in d])
        best = capi.Best()
        out_all = (capi.ReplicaResult * n_replicas)() if want_all else None
        _check(self._lib.cp_run(self._ctx, C.byref(problem), sched, len(schedules), rng, seed_base,
                                replica_begin, n_replicas, init, C.byref(best), out_all))
        return best, out_all

    def prepare(self, problem, schedules, n_replicas, replica_begin=0, seed_base=0, rng=RNG_PHILOX,
                init_dof_buffer=None):
        sched = (capi.Schedule * len(schedules))(*schedules)
        _check(self._lib.cp_prepare(self._ctx, C.byref(problem), sched, len(schedules), rng,
                                    seed_base, replica_begin, n_replicas, init_dof_buffer))

    def launch(self):
        _check(self._lib.cp_launch(self._ctx))

    def synchronize(self):
        _check(self._lib.cp_synchronize(self._ctx))

    def fetch(self, out_all_buffer=None):
        best = capi.Best()
        _check(self._lib.cp_fetch(self._ctx, C.byref(best), out_all_buffer))
        return best

    def timing(self):
        t = capi.Timing()
        _check(self._lib.cp_last_timing(self._ctx, C.byref(t)))
        return t

    def device_buffers(self):
        """(results_ptr, best_ptr): device addresses of the prepared job's outputs."""
        a, b = C.c_void_p(), C.c_void_p()
        _check(self._lib.cp_device_buffers(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_peaks(self):
        """(fp64, fp32) FMA TFLOP/s measured on this GPU."""
        a, b = C.c_double(), C.c_double()
        _check(self._lib.cp_measure_peaks(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def run_host_buffers(self, problem, schedules, n_replicas, init_ptr, out_ptr, replica_begin=0,
                         seed_base=0, rng=RNG_PHILOX):
        """cp_run on caller-owned host buffers given as raw addresses (pinned memory)."""
        sched = (capi.Schedule * len(schedules))(*schedules)
        best = capi.Best()
        _check(self._lib.cp_run(self._ctx, C.byref(problem), sched, len(schedules), rng, seed_base,
                                replica_begin, n_replicas,
                                C.cast(init_ptr, C.POINTER(C.c_double)) if init_ptr else None,
                                C.byref(best),
                                C.cast(out_ptr, C.POINTER(capi.ReplicaResult)) if out_ptr else None))
        return best

    def timing_kernel_ms(self):
        return self.timing().kernel_ms

    def images(self, problem, dofs):
        """Per state and symmetry image: (fx, fy, cx, cy, m00, m01, m10, m11) — OccupiedSite::positions
        and Cell2::to_cartesian_isometry as the kernels evaluate them."""
        n, ns = len(dofs), problem.n_syms
        flat = (C.c_double * (6 * n))(*[v for d in dofs for v in d])
        out = (C.c_double * (8 * ns * n))()
        _check(self._lib.cp_images(self._ctx, C.byref(problem), flat, n, out))
        return [[tuple(out[(r * ns + s) * 8:(r * ns + s) * 8 + 8]) for s in range(ns)] for r in range(n)]

    def replay(self, problem, init_dofs, moves, bound_dofs=None):
        """moves: per replica a list of (basis_index, new_value, threshold, kt).  bound_dofs: the
        states the bounds were captured from (default: init_dofs)."""
        n, m = len(init_dofs), len(moves[0])
        flat = (C.c_double * (6 * n))(*[v for d in init_dofs for v in d])
        if bound_dofs is not None:
            bflat = (C.c_double * (6 * n))(*[v for d in bound_dofs for v in d])
            mv = (capi.Move * (n * m))()
            for r, seq in enumerate(moves):
                for k, (bi, nv, th, kt) in enumerate(seq):
                    mv[r * m + k] = capi.Move(bi, 0, nv, th, kt)
            out = (capi.Decision * (n * m))()
            _check(self._lib.cp_replay_bounded(self._ctx, C.byref(problem), flat, bflat, mv, m, n, out))
            return [[out[r * m + k] for k in range(m)] for r in range(n)]
        mv = (capi.Move * (n * m))()
        for r, seq in enumerate(moves):
            for k, (bi, nv, th, kt) in enumerate(seq):
                mv[r * m + k] = capi.Move(bi, 0, nv, th, kt)
        out = (capi.Decision * (n * m))()
        _check(self._lib.cp_replay(self._ctx, C.byref(problem), flat, mv, m, n, out))
        return [[out[r * m + k] for k in range(m)] for r in range(n)]


class MultiEngine:
    """One cp_multi: a list of GPUs behind one handle (cp_init_multi).  `run` has the signature of
    Engine.run; the replicas are sharded contiguously over the devices and the best state is the
    one a single device would return."""

    def __init__(self, devices):
        self._lib = load_library()
        self.devices = list(devices)
        ids = (C.c_int * len(self.devices))(*self.devices)
        self._m = C.c_void_p()
        _check(self._lib.cp_init_multi(ids, len(self.devices), C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None):
            self._lib.cp_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, problem, schedules, n_replicas, replica_begin=0, seed_base=0, rng=RNG_PCG64MCG,
            init_dof=None, want_all=True):
        sched = (capi.Schedule * len(schedules))(*schedules)
        init = None
        if init_dof is not None:
            init = (C.c_double * (6 * n_replicas))(*[v for d in init_dof for v in d])
        best = capi.Best()
        out_all = (capi.ReplicaResult * n_replicas)() if want_all else None
        _check(self._lib.cp_multi_run(self._m, C.byref(problem), sched, len(schedules), rng, seed_base,
                                      replica_begin, n_replicas, init, C.byref(best), out_all))
        return best, out_all

    def shards(self):
        """(replica_begin, n_replicas, kernel_ms) of every device in the last run."""
        out = []
        for k in range(len(self.devices)):
            b, n = C.c_uint64(), C.c_uint64()
            _check(self._lib.cp_multi_shard(self._m, k, C.byref(b), C.byref(n)))
            ctx, t = C.c_void_p(), capi.Timing()
            _check(self._lib.cp_multi_context(self._m, k, C.byref(ctx)))
            _check(self._lib.cp_last_timing(ctx, C.byref(t)))
            out.append((b.value, n.value, t.kernel_ms))
        return out


_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


# ---------------------------------------------------------------------------------------- states
class _State:
    """Common part of PackedState / PotentialState (traits.rs:71-87)."""

    shape_types = ()

    def __init__(self, shape, group, problem):
        self.shape = shape
        self.wallpaper = group
        self._problem = problem

    @classmethod
    def from_group(cls, shape, group):
        if not isinstance(shape, cls.shape_types):
            raise TypeError(f"{cls.__name__} needs one of {[t.__name__ for t in cls.shape_types]}")
        if isinstance(group, str):
            group = WallpaperGroups.from_str(group)
        p = capi.Problem()
        C.memmove(C.byref(p), C.byref(shape._problem), C.sizeof(capi.Problem))
        if group.key is None:  # operators given explicitly (deserialised state); DOFs are set by the caller
            p.family, p.n_syms = group.family, len(group.symmetries)
            for s_, row in enumerate(group.symmetries):
                for c_, v in enumerate(row):
                    p.syms[s_][c_] = v
        else:
            _check(load_library().cp_problem_from_group(C.byref(p), group.key.encode()))
        return cls(shape, group, p)

    def clone(self):
        p = capi.Problem()
        C.memmove(C.byref(p), C.byref(self._problem), C.sizeof(capi.Problem))
        return type(self)(self.shape, self.wallpaper, p)

    # cell + site DOFs: length, ratio, angle, x, y, theta
    @property
    def dof(self):
        return list(self._problem.dof)

    @dof.setter
    def dof(self, values):
        for i, v in enumerate(values):
            self._problem.dof[i] = v

    def total_shapes(self):
        return self._problem.n_syms

    def score(self, engine=None):
        """State::score(): Some(f64) -> float, None -> None.  Evaluated on the GPU."""
        s = (engine or default_engine()).score(self._problem, [self.dof])[0]
        return None if math.isnan(s) else s

    def generate_basis(self):
        """(name, min, max) of each degree of freedom in the reference's basis order
        (cell.rs:136-170 then site.rs:65-91); bounds are captured from the current state."""
        p = self._problem
        basis = [("length", 0.01, p.dof[0])]
        if p.family in (capi.CP_MONOCLINIC, capi.CP_ORTHORHOMBIC):
            basis.append(("ratio", 0.1, p.dof[1]))
        if p.family == capi.CP_MONOCLINIC:
            basis.append(("angle", math.pi / 4, math.pi / 2))
        basis += [("x", -0.5, 0.5), ("y", -0.5, 0.5), ("site_angle", 0.0, math.tau)]
        return basis

    def __lt__(self, other):  # Ord by score (packed.rs:49-68)
        return self.score() < other.score()

    def to_json(self):
        """serde_json::to_string(&state) (main.rs:263); see serialise.py for the layout."""
        from .serialise import state_to_json

        return state_to_json(self)

    @classmethod
    def from_json(cls, text):
        from .serialise import state_from_json

        return state_from_json(text, cls, cls.shape_types)

    def as_positions(self):
        """State::as_positions (packed.rs:97-106, potential.rs:123-132); see render.py."""
        from .render import as_positions

        return as_positions(self)

    def cartesian_positions(self):
        """3x3 Cartesian transforms of the shapes in the cell (packed.rs:112-115)."""
        from .render import cartesian_positions

        return cartesian_positions(self)

    def as_svg(self):
        """ToSVG::as_svg (to_svg.rs:127-275) as the text svg::save writes (main.rs:266)."""
        from .render import as_svg

        return as_svg(self)


class PackedState(_State):
    shape_types = (LineShape, MolecularShape2)


class PotentialState(_State):
    shape_types = (LJShape2,)


PackedState2, PotentialState2 = PackedState, PotentialState


# ---------------------------------------------------------------------------------------- optimiser
class MCOptimiser:
    """MCOptimiser::new(kt_start, kt_ratio, max_step_size, steps, inner_steps, seed, convergence)."""

    def __init__(self, kt_start, kt_ratio, max_step_size, steps, inner_steps, seed, convergence=None):
        self.schedule = capi.Schedule(kt_start, kt_ratio, max_step_size, steps, inner_steps,
                                      _opt(convergence))
        self.seed = seed

    def optimise_state(self, state, engine=None, rng=RNG_PCG64MCG):
        """Runs the MC loop for this one state on the GPU and returns the optimised state.
        Raises (the reference panics) when the initial state is invalid."""
        eng = engine or default_engine()
        best, _ = eng.run(state._problem, [self.schedule], 1, replica_begin=0, seed_base=self.seed,
                          rng=rng, want_all=False)
        out = state.clone()
        out.dof = list(best.result.dof)
        out.last_rejections = best.result.rejections
        return out


class BuildOptimiser:
    """The CLI's optimiser options (main.rs:27-64); defaults are BuildOptimiser::default() (:66-79)."""

    def __init__(self, steps=1000, kt_start=0.1, kt_finish=0.001, kt_ratio=None, max_step_size=0.01,
                 inner_steps=1000, seed=None, convergence=None):
        self.steps, self.kt_start, self.kt_finish, self.kt_ratio = steps, kt_start, kt_finish, kt_ratio
        self.max_step_size, self.inner_steps, self.seed, self.convergence = (
            max_step_size, inner_steps, seed, convergence)

    @classmethod
    def cli_defaults(cls, **kw):
        """The structopt defaults of the `packing` binary (main.rs:29-63)."""
        base = dict(steps=100, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                    inner_steps=1000, convergence=None)
        base.update(kw)
        return cls(**base)

    def _c(self):
        return capi.BuildOptimiser(self.steps, self.kt_start, _opt(self.kt_finish), _opt(self.kt_ratio),
                                   self.max_step_size, self.inner_steps, _opt(self.convergence))

    def build(self):
        s = capi.Schedule()
        _check(load_library().cp_build_schedule(C.byref(self._c()), C.byref(s)))
        if self.seed is None:
            raise ValueError("BuildOptimiser.build: set a seed (the reference draws one from entropy)")
        return MCOptimiser(s.kt_start, s.kt_ratio, s.max_step_size, s.steps, s.inner_steps, self.seed,
                           None if math.isnan(s.convergence) else s.convergence)

    def pipeline(self):
        """The three schedules analyse_state chains per replica (main.rs:224-252)."""
        out = (capi.Schedule * 3)()
        _check(load_library().cp_pipeline_schedules(C.byref(self._c()), out))
        return list(out)


def monte_carlo_best_packing(state, optimiser, replications, engine=None, rng=RNG_PCG64MCG,
                             replica_begin=0, return_all=False):
    """GPU drop-in for the CLI's analyse_state (main.rs:215-269): `replications` independent
    replicas (seed = replica index) through the three-stage pipeline, then the max by score.
    Returns the best state (and, optionally, every replica's result)."""
    eng = engine or default_engine()  # an Engine (one GPU) or a MultiEngine (a list of GPUs)
    best, out_all = eng.run(state._problem, optimiser.pipeline(), replications,
                            replica_begin=replica_begin, seed_base=0, rng=rng, want_all=return_all)
    if math.isnan(best.result.score):
        raise CrystalPackingError(capi.CP_ERR_INVALID_STATE, "Error in running optimisation.")
    out = state.clone()
    out.dof = list(best.result.dof)
    out.best_replica = best.replica
    out.best_score = best.result.score
    return (out, out_all) if return_all else out
