//! `monte_carlo_best_packing`: the GPU drop-in for the replica loop of the `packing` binary
//! (`analyse_state`, crystal-packing-cli/src/main.rs:215-269), for the three state types the CLI
//! builds (main.rs:160-213):
//!
//!   PackedState<LineShape>          polygons          (state/packed.rs, shape/line_shape.rs)
//!   PackedState<MolecularShape2>    disc molecules    (state/packed.rs, shape/molecular_shape2.rs)
//!   PotentialState<LJShape2>        Lennard-Jones     (state/potential.rs, shape/lj_shape.rs)
//!
//! A state is flattened into the C ABI's `cp_problem` (`GpuState::to_problem`), the three-stage
//! pipeline runs for `replications` replicas (seed = replica index) on the listed GPUs behind one
//! `cp_multi` handle (replicas sharded contiguously over the devices, per-device maxima reduced on the
//! host by the rule of rayon's `max()`), and the best replica's degrees of freedom are written back
//! through the crate's own `Basis` handles.
//!
//! The existing crate needs one addition for this to compile: `occupied_sites` is private in both
//! state types (state/packed.rs:32, state/potential.rs:31), so each gains
//!     pub fn site_dofs(&self) -> [f64; 3]           // x, y, angle of the single occupied site
//!     pub fn symmetries(&self) -> Vec<Transform2>   // the Wyckoff site's operators
//! Source only — never compiled (no Rust toolchain in the build image).
pub mod ffi;

use std::ffi::CStr;

use anyhow::{anyhow, Error};
use crystal_packing::traits::{Shape, State};
use crystal_packing::{Cell2, CrystalFamily, LJShape2, LineShape, MolecularShape2, PackedState, PotentialState, Transform2};

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

/// A shape the engine has kernels for: its components as rows of `cp_problem::items`.
pub trait GpuShape: Shape {
    const KIND: i32;
    fn item_rows(&self) -> Vec<[f64; 5]>;
    /// `Shape::area` where the state uses it (packed.rs:84); 0 for potentials
    fn packed_area(&self) -> f64 {
        self.area()
    }
}

impl GpuShape for LineShape {
    const KIND: i32 = CP_POLYGON;
    fn item_rows(&self) -> Vec<[f64; 5]> {
        // Line2 { start, end } (shape/components/line2.rs:16-20)
        self.items.iter().map(|l| [l.start.x, l.start.y, l.end.x, l.end.y, 0.0]).collect()
    }
}

impl GpuShape for MolecularShape2 {
    const KIND: i32 = CP_DISCS;
    fn item_rows(&self) -> Vec<[f64; 5]> {
        // Atom2 { position, radius } (shape/components/atom2.rs:14-18)
        self.items.iter().map(|a| [a.position.x, a.position.y, a.radius, 0.0, 0.0]).collect()
    }
}

impl GpuShape for LJShape2 {
    const KIND: i32 = CP_LJ;
    fn item_rows(&self) -> Vec<[f64; 5]> {
        // LJ2 { position, sigma, epsilon, cutoff: Option<f64> } (shape/components/lj2.rs:18-29); None = NaN
        self.items
            .iter()
            .map(|s| [s.position.x, s.position.y, s.sigma, s.epsilon, s.cutoff.unwrap_or(f64::NAN)])
            .collect()
    }
    fn packed_area(&self) -> f64 {
        0.0
    }
}

/// A state the engine can optimise: one shape, one plane group, one occupied general-position site.
pub trait GpuState: State + Clone {
    type S: GpuShape;
    fn shape(&self) -> &Self::S;
    fn cell(&self) -> &Cell2;
    fn family(&self) -> CrystalFamily;
    fn symmetries(&self) -> Vec<Transform2>;
    fn site_dofs(&self) -> [f64; 3];

    fn to_problem(&self) -> Result<CpProblem, Error> {
        let mut p: CpProblem = unsafe { std::mem::zeroed() };
        let rows = self.shape().item_rows();
        if rows.is_empty() || rows.len() > CP_MAX_ITEMS {
            return Err(anyhow!("the GPU engine takes shapes of 1..{} components", CP_MAX_ITEMS));
        }
        p.kind = <Self::S as GpuShape>::KIND;
        p.n_items = rows.len() as i32;
        for (k, r) in rows.iter().enumerate() {
            p.items[k] = *r;
        }
        p.area = self.shape().packed_area();
        p.enclosing_radius = self.shape().enclosing_radius();
        p.family = family_code(self.family());
        let syms = self.symmetries();
        if syms.is_empty() || syms.len() > CP_MAX_SYMS {
            return Err(anyhow!("the GPU engine takes 1..{} symmetry operators", CP_MAX_SYMS));
        }
        p.n_syms = syms.len() as i32;
        for (s, t) in syms.iter().enumerate() {
            let m: nalgebra::Matrix3<f64> = (*t).into();
            p.syms[s] = [m[(0, 0)], m[(0, 1)], m[(0, 2)], m[(1, 0)], m[(1, 1)], m[(1, 2)]];
        }
        let [x, y, theta] = self.site_dofs();
        // ratio = b / a reproduces Cell2::ratio only to rounding; a `Cell2::ratio()` accessor is the
        // exact route (cell.rs keeps length / ratio / angle private)
        let cell = self.cell();
        p.dof = [cell.a(), cell.b() / cell.a(), cell.angle(), x, y, theta];
        Ok(p)
    }
}

impl<T: GpuShape + crystal_packing::traits::Intersect> GpuState for PackedState<T> {
    type S = T;
    fn shape(&self) -> &T {
        &self.shape
    }
    fn cell(&self) -> &Cell2 {
        &self.cell
    }
    fn family(&self) -> CrystalFamily {
        self.wallpaper.family
    }
    fn symmetries(&self) -> Vec<Transform2> {
        PackedState::symmetries(self)
    }
    fn site_dofs(&self) -> [f64; 3] {
        PackedState::site_dofs(self)
    }
}

impl<T: GpuShape + crystal_packing::traits::Potential> GpuState for PotentialState<T> {
    type S = T;
    fn shape(&self) -> &T {
        &self.shape
    }
    fn cell(&self) -> &Cell2 {
        &self.cell
    }
    fn family(&self) -> CrystalFamily {
        self.wallpaper.family
    }
    fn symmetries(&self) -> Vec<Transform2> {
        PotentialState::symmetries(self)
    }
    fn site_dofs(&self) -> [f64; 3] {
        PotentialState::site_dofs(self)
    }
}

/// Owns a `cp_multi`: one context and one stream per listed device.
pub struct Devices(*mut CpMulti);

impl Devices {
    /// `device_ids` as CUDA sees them; an id may repeat (two streams on that GPU).
    pub fn new(device_ids: &[i32]) -> Result<Self, Error> {
        let mut m = std::ptr::null_mut();
        // CP_ERR_NO_DEVICE: there is no CPU fallback in the library
        check(unsafe { cp_init_multi(device_ids.as_ptr(), device_ids.len() as i32, &mut m) })?;
        Ok(Devices(m))
    }
    /// every GPU the process can see
    pub fn all() -> Result<Self, Error> {
        let mut n = 0;
        check(unsafe { cp_device_count(&mut n) })?;
        Self::new(&(0..n).collect::<Vec<i32>>())
    }
}

impl Drop for Devices {
    fn drop(&mut self) {
        unsafe { cp_multi_destroy(self.0) }
    }
}

/// What `analyse_state` computes: `(0..replications).into_par_iter().map(three stages).max()`
/// (main.rs:221-253), on the GPUs of `devices`.
pub fn monte_carlo_best_packing<St: GpuState>(
    state: St,
    opts: &GpuOptimiser,
    replications: u64,
    devices: &Devices,
) -> Result<St, Error> {
    let p = state.to_problem()?;
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
        check(cp_pipeline_schedules(&b, stages.as_mut_ptr()))?;
        // CP_ERR_INVALID_STATE is the reference's panic on an overlapping initial state
        check(cp_multi_run(
            devices.0,
            &p,
            stages.as_ptr(),
            3,
            CP_RNG_PCG64MCG,
            0,
            0,
            replications,
            std::ptr::null(),
            &mut best,
            std::ptr::null_mut(),
        ))?;
    }
    if best.result.score.is_nan() {
        return Err(anyhow!("Error in running optimisation."));
    }
    let out = state.clone();
    for (basis, v) in out.generate_basis().iter().zip(dof_in_basis_order(&best.result.dof, state.family())) {
        basis.set_value(v)?;
    }
    Ok(out)
}
