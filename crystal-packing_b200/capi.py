"""ctypes declarations of include/cp_gpu.h (the C ABI of lib/libcpgpu.so)."""
import ctypes as C
import os

MAX_ITEMS, MAX_SYMS, MAX_DOF, MAX_STAGES = 16, 4, 6, 4

CP_OK, CP_ERR_INVALID_ARG, CP_ERR_INVALID_STATE, CP_ERR_CUDA, CP_ERR_NO_DEVICE, CP_ERR_PARSE, \
    CP_ERR_NOT_PREPARED = range(7)
CP_POLYGON, CP_DISCS, CP_LJ = 0, 1, 2
CP_MONOCLINIC, CP_ORTHORHOMBIC, CP_HEXAGONAL, CP_TETRAGONAL = 0, 1, 2, 3
CP_RNG_PCG64MCG, CP_RNG_PHILOX = 0, 1


class Problem(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_items", C.c_int32),
        ("items", (C.c_double * 5) * MAX_ITEMS),
        ("area", C.c_double),
        ("enclosing_radius", C.c_double),
        ("family", C.c_int32),
        ("n_syms", C.c_int32),
        ("syms", (C.c_double * 6) * MAX_SYMS),
        ("dof", C.c_double * MAX_DOF),
    ]


class Schedule(C.Structure):
    _fields_ = [
        ("kt_start", C.c_double),
        ("kt_ratio", C.c_double),
        ("max_step_size", C.c_double),
        ("steps", C.c_uint64),
        ("inner_steps", C.c_uint64),
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


class ReplicaResult(C.Structure):
    _fields_ = [
        ("dof", C.c_double * MAX_DOF),
        ("score", C.c_double),
        ("rejections", C.c_uint64),
        ("moves", C.c_uint64),
    ]


class Best(C.Structure):
    _fields_ = [("result", ReplicaResult), ("replica", C.c_uint64)]


class Move(C.Structure):
    _fields_ = [
        ("basis_index", C.c_uint32),
        ("pad_", C.c_uint32),
        ("new_value", C.c_double),
        ("threshold", C.c_double),
        ("kt", C.c_double),
    ]


class Decision(C.Structure):
    _fields_ = [
        ("in_bounds", C.c_uint32),
        ("accepted", C.c_uint32),
        ("new_score", C.c_double),
        ("score_current", C.c_double),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_float),
        ("reduce_ms", C.c_float),
        ("h2d_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("launches", C.c_uint32),
        ("lanes_per_replica", C.c_uint32),
    ]


# every symbol include/cp_gpu.h declares: name -> (restype, argtypes)
_P = C.POINTER
_ctx = C.c_void_p
_d, _u64, _i = C.c_double, C.c_uint64, C.c_int
SIGNATURES = {
    "cp_device_count": (_i, [_P(_i)]),
    "cp_init": (_i, [_i, _P(_ctx)]),
    "cp_destroy": (None, [_ctx]),
    "cp_last_error": (C.c_char_p, []),
    "cp_version": (C.c_char_p, []),
    "cp_set_stream": (_i, [_ctx, C.c_void_p]),
    "cp_set_lanes": (_i, [_ctx, _i]),
    "cp_shape_polygon": (_i, [_P(_d), _i, _P(Problem)]),
    "cp_shape_trimer": (_i, [_i, _d, _d, _d, _P(Problem)]),
    "cp_shape_circle": (_i, [_i, _P(Problem)]),
    "cp_shape_finish": (_i, [_P(Problem)]),
    "cp_parse_operations": (_i, [C.c_char_p, _P(_d)]),
    "cp_group_lookup": (_i, [C.c_char_p, _P(C.c_int32), _P(C.c_int32), _P((_d * 6) * MAX_SYMS)]),
    "cp_problem_from_group": (_i, [_P(Problem), C.c_char_p]),
    "cp_build_schedule": (_i, [_P(BuildOptimiser), _P(Schedule)]),
    "cp_pipeline_schedules": (_i, [_P(BuildOptimiser), _P(Schedule)]),
    "cp_run": (_i, [_ctx, _P(Problem), _P(Schedule), _i, _i, _u64, _u64, _u64, _P(_d), _P(Best),
                    _P(ReplicaResult)]),
    "cp_prepare": (_i, [_ctx, _P(Problem), _P(Schedule), _i, _i, _u64, _u64, _u64, _P(_d)]),
    "cp_launch": (_i, [_ctx]),
    "cp_fetch": (_i, [_ctx, _P(Best), _P(ReplicaResult)]),
    "cp_synchronize": (_i, [_ctx]),
    "cp_device_buffers": (_i, [_ctx, _P(C.c_void_p), _P(C.c_void_p)]),
    "cp_last_timing": (_i, [_ctx, _P(Timing)]),
    "cp_measure_peaks": (_i, [_ctx, _P(_d), _P(_d)]),
    "cp_score": (_i, [_ctx, _P(Problem), _P(_d), _u64, _P(_d)]),
    "cp_replay": (_i, [_ctx, _P(Problem), _P(_d), _P(Move), _u64, _u64, _P(Decision)]),
    "cp_replay_bounded": (_i, [_ctx, _P(Problem), _P(_d), _P(_d), _P(Move), _u64, _u64, _P(Decision)]),
    "cp_images": (_i, [_ctx, _P(Problem), _P(_d), _u64, _P(_d)]),
    "cp_init_multi": (_i, [_P(_i), _i, _P(_ctx)]),
    "cp_multi_destroy": (None, [_ctx]),
    "cp_multi_device_count": (_i, [_ctx, _P(_i)]),
    "cp_multi_context": (_i, [_ctx, _i, _P(_ctx)]),
    "cp_multi_shard": (_i, [_ctx, _i, _P(_u64), _P(_u64)]),
    "cp_multi_run": (_i, [_ctx, _P(Problem), _P(Schedule), _i, _i, _u64, _u64, _u64, _P(_d), _P(Best),
                          _P(ReplicaResult)]),
}


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libcpgpu.so")


_lib = None


def load_library():
    """dlopen lib/libcpgpu.so and type every exported entry point.  Raises if it is missing: the
    product has no other implementation to fall back to."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise OSError(
                f"{path} is missing: build it with `make -C crystal-packing_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback"
            )
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
