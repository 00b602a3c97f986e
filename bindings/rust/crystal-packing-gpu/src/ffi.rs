//! Raw bindings of include/cp_gpu.h.  Layouts are `#[repr(C)]` mirrors of the C structs; sizes are
//! checked against the library by `tests/test_host_boundary.py::test_struct_layouts_match_header`
//! on the Python side (Problem 912 B, Schedule 48 B, ReplicaResult 72 B, Best 80 B).
use std::os::raw::{c_char, c_int, c_void};

pub const CP_MAX_ITEMS: usize = 16;
pub const CP_MAX_SYMS: usize = 4;
pub const CP_MAX_DOF: usize = 6;

pub const CP_OK: c_int = 0;
pub const CP_ERR_INVALID_STATE: c_int = 2;
pub const CP_ERR_NO_DEVICE: c_int = 4;

pub const CP_POLYGON: i32 = 0;
pub const CP_DISCS: i32 = 1;
pub const CP_LJ: i32 = 2;
pub const CP_RNG_PCG64MCG: c_int = 0;
pub const CP_RNG_PHILOX: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct CpProblem {
    pub kind: i32,
    pub n_items: i32,
    pub items: [[f64; 5]; CP_MAX_ITEMS],
    pub area: f64,
    pub enclosing_radius: f64,
    pub family: i32,
    pub n_syms: i32,
    pub syms: [[f64; 6]; CP_MAX_SYMS],
    pub dof: [f64; CP_MAX_DOF],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct CpSchedule {
    pub kt_start: f64,
    pub kt_ratio: f64,
    pub max_step_size: f64,
    pub steps: u64,
    pub inner_steps: u64,
    pub convergence: f64, // NaN = None
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct CpBuildOptimiser {
    pub steps: u64,
    pub kt_start: f64,
    pub kt_finish: f64, // NaN = None
    pub kt_ratio: f64,  // NaN = None
    pub max_step_size: f64,
    pub inner_steps: u64,
    pub convergence: f64, // NaN = None
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct CpReplicaResult {
    pub dof: [f64; CP_MAX_DOF],
    pub score: f64, // NaN = None
    pub rejections: u64,
    pub moves: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct CpBest {
    pub result: CpReplicaResult,
    pub replica: u64,
}

#[repr(C)]
pub struct CpCtx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct CpMulti {
    _private: [u8; 0],
}

extern "C" {
    pub fn cp_device_count(out: *mut c_int) -> c_int;
    pub fn cp_init(device: c_int, out: *mut *mut CpCtx) -> c_int;
    pub fn cp_destroy(ctx: *mut CpCtx);
    pub fn cp_last_error() -> *const c_char;
    pub fn cp_set_stream(ctx: *mut CpCtx, cuda_stream: *mut c_void) -> c_int;
    pub fn cp_set_lanes(ctx: *mut CpCtx, lanes: c_int) -> c_int;
    pub fn cp_build_schedule(opts: *const CpBuildOptimiser, out: *mut CpSchedule) -> c_int;
    pub fn cp_pipeline_schedules(opts: *const CpBuildOptimiser, out: *mut CpSchedule) -> c_int;
    pub fn cp_run(
        ctx: *mut CpCtx,
        problem: *const CpProblem,
        stages: *const CpSchedule,
        n_stages: c_int,
        rng_mode: c_int,
        seed_base: u64,
        replica_begin: u64,
        n_replicas: u64,
        init_dof: *const f64,
        out_best: *mut CpBest,
        out_all: *mut CpReplicaResult,
    ) -> c_int;
    pub fn cp_score(ctx: *mut CpCtx, problem: *const CpProblem, dof: *const f64, n: u64, out_scores: *mut f64) -> c_int;
    pub fn cp_init_multi(device_ids: *const c_int, n_devices: c_int, out: *mut *mut CpMulti) -> c_int;
    pub fn cp_multi_destroy(m: *mut CpMulti);
    pub fn cp_multi_device_count(m: *const CpMulti, out: *mut c_int) -> c_int;
    pub fn cp_multi_context(m: *mut CpMulti, k: c_int, out: *mut *mut CpCtx) -> c_int;
    pub fn cp_multi_shard(m: *const CpMulti, k: c_int, replica_begin: *mut u64, n_replicas: *mut u64) -> c_int;
    pub fn cp_multi_run(
        m: *mut CpMulti,
        problem: *const CpProblem,
        stages: *const CpSchedule,
        n_stages: c_int,
        rng_mode: c_int,
        seed_base: u64,
        replica_begin: u64,
        n_replicas: u64,
        init_dof: *const f64,
        out_best: *mut CpBest,
        out_all: *mut CpReplicaResult,
    ) -> c_int;
}
