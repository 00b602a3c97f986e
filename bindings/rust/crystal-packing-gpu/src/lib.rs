//! `monte_carlo_best_packing`: the GPU drop-in for the replica loop of the `packing` binary
//! (`analyse_state`, crystal-packing-cli/src/main.rs:215-269).
//!
//! The function flattens a `PackedState<LineShape>` into the C ABI's `cp_problem`, runs the
//! three-stage pipeline for `replications` replicas (seed = replica index) on one GPU and writes the
//! best replica's degrees of freedom back through the crate's own `Basis` handles.
//!
//! The existing crate needs one addition for this to compile: `PackedState::occupied_sites` is
//! private (state/packed.rs:32), so `PackedState` gains
//!     pub fn site_dofs(&self) -> [f64; 3]           // x, y, angle of the single occupied site
//!     pub fn symmetries(&self) -> Vec<Transform2>   // the Wyckoff site's operators
//! Source only — never compiled (no Rust toolchain in the build image).
pub mod ffi;

use std::ffi::CStr;

use anyhow::{anyhow, Error};
use crystal_packing::traits::{Intersect, Shape, State};
use crystal_packing::{CrystalFamily, LineShape, PackedState};

use ffi::*;

/// Mirrors the CLI's `BuildOptimiser` (main.rs:27-64), which lives in the binary crate.
#[derive(Clone, Copy, Debug)]
pub struct GpuOptimiser {
    pub steps: u64,
    pub kt_start: f64,
    pub kt_finish: Option<f64>,
    pub kt_ratio: Option<f64>,
    pub max_step_size: f64,
    pub inner_steps: u64,
    pub convergence: Option<f64>,
}

fn check(status: i32) -> Result<(), Error> {
    if status == CP_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(cp_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow!("libcpgpu status {}: {}", status, msg))
}

fn family_code(family: CrystalFamily) -> i32 {
    match family {
        CrystalFamily::Monoclinic => 0,
        CrystalFamily::Orthorhombic => 1,
        CrystalFamily::Hexagonal => 2,
        CrystalFamily::Tetragonal => 3,
    }
}

/// The kernels always carry six values (length, ratio, angle, x, y, theta); `generate_basis()`
/// omits the cell DOFs a family does not have (cell.rs:136-170).
fn dof_in_basis_order(dof: &[f64; 6], family: CrystalFamily) -> Vec<f64> {
    match family {
        CrystalFamily::Monoclinic => dof.to_vec(),
        CrystalFamily::Orthorhombic => vec![dof[0], dof[1], dof[3], dof[4], dof[5]],
        _ => vec![dof[0], dof[3], dof[4], dof[5]],
    }
}

pub fn monte_carlo_best_packing(
    state: PackedState<LineShape>,
    opts: &GpuOptimiser,
    replications: u64,
    device: i32,
) -> Result<PackedState<LineShape>, Error> {
    let mut p: CpProblem = unsafe { std::mem::zeroed() };
    if state.shape.items.len() > CP_MAX_ITEMS {
        return Err(anyhow!("shapes above {} components are not supported by the GPU engine", CP_MAX_ITEMS));
    }
    p.kind = CP_POLYGON;
    p.n_items = state.shape.items.len() as i32;
    for (k, l) in state.shape.items.iter().enumerate() {
        p.items[k] = [l.start.x, l.start.y, l.end.x, l.end.y, 0.0];
    }
    p.area = state.shape.area();
    p.enclosing_radius = state.shape.enclosing_radius();
    p.family = family_code(state.wallpaper.family);
    let syms = state.symmetries();
    if syms.len() > CP_MAX_SYMS {
        return Err(anyhow!("more than {} symmetry operators", CP_MAX_SYMS));
    }
    p.n_syms = syms.len() as i32;
    for (s, t) in syms.iter().enumerate() {
        let m: nalgebra::Matrix3<f64> = (*t).into();
        p.syms[s] = [m[(0, 0)], m[(0, 1)], m[(0, 2)], m[(1, 0)], m[(1, 1)], m[(1, 2)]];
    }
    let [x, y, theta] = state.site_dofs();
    // ratio = b / a reproduces Cell2::ratio only to rounding; a `Cell2::ratio()` accessor is the
    // exact route (cell.rs keeps length / ratio / angle private)
    p.dof = [state.cell.a(), state.cell.b() / state.cell.a(), state.cell.angle(), x, y, theta];

    let b = CpBuildOptimiser {
        steps: opts.steps,
        kt_start: opts.kt_start,
        kt_finish: opts.kt_finish.unwrap_or(f64::NAN),
        kt_ratio: opts.kt_ratio.unwrap_or(f64::NAN),
        max_step_size: opts.max_step_size,
        inner_steps: opts.inner_steps,
        convergence: opts.convergence.unwrap_or(f64::NAN),
    };
    let mut stages = [unsafe { std::mem::zeroed::<CpSchedule>() }; 3];
    let mut best = CpBest::default();
    unsafe {
        let mut ctx = std::ptr::null_mut();
        check(cp_init(device, &mut ctx))?; // CP_ERR_NO_DEVICE: there is no CPU fallback in the library
        let rc = cp_pipeline_schedules(&b, stages.as_mut_ptr());
        let rc = if rc == CP_OK {
            cp_run(ctx, &p, stages.as_ptr(), 3, CP_RNG_PCG64MCG, 0, 0, replications, std::ptr::null(), &mut best,
                   std::ptr::null_mut())
        } else {
            rc
        };
        cp_destroy(ctx);
        check(rc)?; // CP_ERR_INVALID_STATE is the reference's panic on an overlapping initial state
    }
    if best.result.score.is_nan() {
        return Err(anyhow!("Error in running optimisation."));
    }
    let out = state.clone();
    for (basis, v) in out.generate_basis().iter().zip(dof_in_basis_order(&best.result.dof, state.wallpaper.family)) {
        basis.set_value(v)?;
    }
    Ok(out)
}
