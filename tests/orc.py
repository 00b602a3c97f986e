"""ctypes binding of the CPU oracle (oracle/libcp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke().  Never imported by the product package.
"""
import ctypes as C
import math
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "libcp_oracle.so")

MAX_ITEMS, MAX_SYMS, MAX_DOF = 64, 8, 6
POLYGON, DISCS, LJ = 0, 1, 2
MONOCLINIC, ORTHORHOMBIC, HEXAGONAL, TETRAGONAL = 0, 1, 2, 3
MATH_LIBM, MATH_CP = 0, 1
NAN = float("nan")


class Transform(C.Structure):
    _fields_ = [("m", (C.c_double * 3) * 3)]

    def rows(self):
        return [[self.m[i][j] for j in range(3)] for i in range(3)]


class Item(C.Structure):
    _fields_ = [(k, C.c_double) for k in "abcde"]

    def tup(self):
        return (self.a, self.b, self.c, self.d, self.e)


class State(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_items", C.c_int32),
        ("items", Item * MAX_ITEMS),
        ("family", C.c_int32),
        ("n_syms", C.c_int32),
        ("syms", Transform * MAX_SYMS),
        ("length", C.c_double),
        ("ratio", C.c_double),
        ("angle", C.c_double),
        ("x", C.c_double),
        ("y", C.c_double),
        ("theta", C.c_double),
        ("math_mode", C.c_int32),
        ("pad_", C.c_int32),
    ]


class Optimiser(C.Structure):
    _fields_ = [
        ("kt_start", C.c_double),
        ("kt_ratio", C.c_double),
        ("max_step_size", C.c_double),
        ("steps", C.c_uint64),
        ("inner_steps", C.c_uint64),
        ("seed", C.c_uint64),
        ("convergence", C.c_double),
    ]


class BuildOptimiser(C.Structure):
    _fields_ = [
        ("steps", C.c_uint64),
        ("kt_start", C.c_double),
        ("kt_finish", C.c_double),
        ("kt_ratio", C.c_double),
        ("max_step_size", C.c_double),
        ("inner_steps", C.c_uint64),
        ("convergence", C.c_double),
    ]


class MoveRecord(C.Structure):
    _fields_ = [
        ("basis_index", C.c_uint32),
        ("in_bounds", C.c_uint32),
        ("accepted", C.c_uint32),
        ("pad_", C.c_uint32),
        ("old_value", C.c_double),
        ("new_value", C.c_double),
        ("threshold", C.c_double),
        ("new_score", C.c_double),
        ("score_current", C.c_double),
        ("kt", C.c_double),
        ("step_ratio", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("score", C.c_double),
        ("dof", C.c_double * MAX_DOF),
        ("rejections", C.c_uint64),
        ("moves", C.c_uint64),
    ]


class Counters(C.Structure):
    _fields_ = [
        (k, C.c_uint64)
        for k in (
            "moves scored segment_tests disc_tests lj_in lj_out image_cands image_pass "
            "point_xforms site_images trig"
        ).split()
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Pcg(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_ORACLE_DIR, f) for f in ("cp_oracle.c", "cp_oracle.h")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "crystal-packing_b200", "csrc", "cp_math.h"))
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s)
    )
    if stale:
        subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])
    return _SO


def build_fast():
    """The timed CPU baseline: -O3 -march=native, counters compiled out (oracle/Makefile `fast`).
    Always rebuilt on the machine that runs it (-march=native must not travel)."""
    subprocess.check_call(["make", "-s", "-B", "-C", _ORACLE_DIR, "fast"])
    return os.path.join(_ORACLE_DIR, "libcp_oracle_fast.so")


_lib = None
_fast = None


def lib_fast():
    """ctypes handle of the fast build; only orc_analyse_state is typed (bench.py's timed legs)."""
    global _fast
    if _fast is None:
        try:
            L = C.CDLL(build_fast())
        except (subprocess.CalledProcessError, OSError):  # no compiler on this machine: the checker build
            L = C.CDLL(build())
        L.orc_analyse_state.restype = C.c_int64
        L.orc_analyse_state.argtypes = [C.POINTER(State), C.POINTER(BuildOptimiser), C.c_uint64, C.c_uint64,
                                        C.c_int, C.POINTER(Result), C.POINTER(Result)]
        _fast = L
    return _fast


def analyse_state_fast(state, bopt, begin, n, n_threads=1):
    """orc_analyse_state of the fast build: (best index, best, results)."""
    results = (Result * n)()
    best = Result()
    idx = lib_fast().orc_analyse_state(C.byref(state), C.byref(bopt), begin, n, n_threads, results, C.byref(best))
    return idx, best, results


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    d, u64, i32, i64 = C.c_double, C.c_uint64, C.c_int, C.c_int64
    P = C.POINTER
    sig = {
        "orc_counters_reset": (None, []),
        "orc_counters_get": (None, [P(Counters)]),
        "orc_counters_flops": (d, [P(Counters)]),
        "orc_pcg_new": (None, [P(Pcg), u64, u64]),
        "orc_pcg_from_seed": (None, [P(Pcg), C.c_char_p]),
        "orc_pcg_seed_from_u64": (None, [P(Pcg), u64]),
        "orc_pcg_next_u64": (u64, [P(Pcg)]),
        "orc_uniform_usize": (u64, [P(Pcg), u64]),
        "orc_uniform_step": (d, [P(Pcg)]),
        "orc_gen_f64": (d, [P(Pcg)]),
        "orc_sincos": (None, [i32, d, P(d), P(d)]),
        "orc_exp": (d, [i32, d]),
        "orc_cp_sincos": (None, [d, P(d), P(d)]),
        "orc_cp_exp": (d, [d]),
        "orc_transform_new": (None, [i32, d, d, d, P(Transform)]),
        "orc_transform_identity": (None, [P(Transform)]),
        "orc_transform_mul": (None, [P(Transform), P(Transform), P(Transform)]),
        "orc_transform_mul_point": (None, [P(Transform), d, d, P(d), P(d)]),
        "orc_transform_position": (None, [P(Transform), P(d), P(d)]),
        "orc_transform_periodic": (None, [P(Transform), d, d, P(Transform)]),
        "orc_transform_from_operations": (i32, [C.c_char_p, P(Transform)]),
        "orc_cell_to_cartesian": (None, [P(State), d, d, P(d), P(d)]),
        "orc_cell_area": (d, [P(State)]),
        "orc_cell_periodic_images": (i32, [P(State), P(Transform), i32, i32, P(Transform)]),
        "orc_line_intersects": (i32, [P(Item), P(Item)]),
        "orc_atom_intersects": (i32, [P(Item), P(Item)]),
        "orc_lj_energy": (d, [P(Item), P(Item)]),
        "orc_shape_from_radial": (i32, [P(d), i32, P(Item)]),
        "orc_shape_trimer": (None, [i32, d, d, d, P(Item)]),
        "orc_shape_circle": (None, [i32, P(Item)]),
        "orc_shape_area": (d, [i32, P(Item), i32]),
        "orc_shape_enclosing_radius": (d, [i32, P(Item), i32]),
        "orc_shape_transform": (None, [i32, P(Item), i32, P(Transform), P(Item)]),
        "orc_shape_intersects": (i32, [i32, P(Item), P(Item), i32]),
        "orc_shape_energy": (d, [P(Item), P(Item), i32]),
        "orc_circle_overlap": (d, [P(Item), P(Item)]),
        "orc_group": (i32, [C.c_char_p, P(C.c_int32), P(C.c_int32), P(Transform)]),
        "orc_state_from_group": (i32, [P(State), i32, P(Item), i32, C.c_char_p, i32]),
        "orc_state_relative_positions": (i32, [P(State), P(Transform)]),
        "orc_state_cartesian_positions": (i32, [P(State), P(Transform)]),
        "orc_state_check_intersection": (i32, [P(State)]),
        "orc_state_score": (d, [P(State)]),
        "orc_state_total_shapes": (i32, [P(State)]),
        "orc_state_n_dof": (i32, [P(State)]),
        "orc_state_set_dof": (None, [P(State), P(d)]),
        "orc_state_get_dof": (None, [P(State), P(d)]),
        "orc_state_bounds": (i32, [P(State), P(d), P(d)]),
        "orc_energy_surface": (d, [i32, d, d, d]),
        "orc_optimise_state": (i32, [P(Optimiser), P(State), P(MoveRecord), u64, P(u64), P(u64)]),
        "orc_replay_move": (i32, [P(State), i32, d, d, d, d, d, P(d), P(C.c_int), P(d)]),
        "orc_build": (None, [P(BuildOptimiser), u64, P(Optimiser)]),
        "orc_run_replica": (i32, [P(State), P(BuildOptimiser), u64, P(State), P(Result)]),
        "orc_analyse_state": (i64, [P(State), P(BuildOptimiser), u64, u64, i32, P(Result), P(Result)]),
        "orc_analyse_state_counted": (
            i64,
            [P(State), P(BuildOptimiser), u64, u64, i32, P(Result), P(Result), P(Counters)],
        ),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# ---------------------------------------------------------------- convenience helpers
def items_array(tuples):
    arr = (Item * len(tuples))()
    for i, t in enumerate(tuples):
        t = tuple(t) + (0.0,) * (5 - len(t))
        arr[i] = Item(*t)
    return arr


def polygon(sides, radii=None):
    radii = radii if radii is not None else [1.0] * sides
    r = (C.c_double * len(radii))(*radii)
    items = (Item * len(radii))()
    rc = lib().orc_shape_from_radial(r, len(radii), items)
    if rc:
        raise ValueError("The number of points provided is too few to create a 2D shape.")
    return POLYGON, items


def trimer(kind, radius=0.637556, angle=120.0, distance=1.0):
    items = (Item * 3)()
    lib().orc_shape_trimer(kind, radius, angle, distance, items)
    return kind, items


def circle(kind):
    items = (Item * 1)()
    lib().orc_shape_circle(kind, items)
    return kind, items


def state_from_group(shape, group, math_mode=MATH_LIBM):
    kind, items = shape
    s = State()
    rc = lib().orc_state_from_group(C.byref(s), kind, items, len(items), group.encode(), math_mode)
    if rc:
        raise ValueError(f"orc_state_from_group({group}) -> {rc}")
    return s


def clone(state):
    s = State()
    C.memmove(C.byref(s), C.byref(state), C.sizeof(State))
    return s


def score(state):
    return lib().orc_state_score(C.byref(state))


def get_dof(state):
    d = (C.c_double * MAX_DOF)()
    lib().orc_state_get_dof(C.byref(state), d)
    return list(d)


def set_dof(state, dof):
    d = (C.c_double * MAX_DOF)(*dof)
    lib().orc_state_set_dof(C.byref(state), d)


def optimise(state, kt_start, kt_ratio, max_step_size, steps, inner_steps, seed, convergence=None,
             record=False):
    """MCOptimiser::new(...).optimise_state(state); mutates `state`.  Returns (rc, records, rejections)."""
    opt = Optimiser(kt_start, kt_ratio, max_step_size, steps, inner_steps, seed,
                    NAN if convergence is None else convergence)
    n_max = (steps // inner_steps) * inner_steps if (record and inner_steps) else 0
    recs = (MoveRecord * max(n_max, 1))()
    n = C.c_uint64(0)
    rej = C.c_uint64(0)
    rc = lib().orc_optimise_state(C.byref(opt), C.byref(state), recs if record else None, n_max,
                                  C.byref(n), C.byref(rej))
    return rc, (recs[: n.value] if record else None), rej.value


def build_optimiser(steps=100, kt_start=0.1, kt_finish=None, kt_ratio=None, max_step_size=0.01,
                    inner_steps=1000, convergence=None):
    f = lambda v: NAN if v is None else v
    return BuildOptimiser(steps, kt_start, f(kt_finish), f(kt_ratio), max_step_size, inner_steps,
                          f(convergence))


def analyse_state(state, bopt, begin, n, n_threads=1, want_results=True, want_counters=False):
    results = (Result * n)() if want_results else None
    best = Result()
    cnt = Counters()
    idx = lib().orc_analyse_state_counted(C.byref(state), C.byref(bopt), begin, n, n_threads,
                                          results, C.byref(best), C.byref(cnt))
    return idx, best, results, (cnt if want_counters else None)


def isnan(x):
    return math.isnan(x)
