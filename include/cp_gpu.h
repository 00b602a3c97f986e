/*
 * cp_gpu.h — C ABI of libcpgpu.so, the B200 (sm_100a) engine for crystal-packing's Monte-Carlo
 * packing optimisation over many independent replicas.
 *
 * The reference (malramsay64/crystal-packing, Rust) has NO FFI for this path: the boundary it
 * offers is the Rust trait surface (`State::score`, `MCOptimiser::optimise_state`) driven by the
 * CLI's rayon loop `analyse_state`.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  INTEGRATION.md shows the `extern "C"` block
 * and the `monte_carlo_best_packing` wrapper a maintainer adds on the Rust side.
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, every function returns a
 * cp_status (0 = ok) and never throws; f64 everywhere (the reference is f64 end to end);
 * Option<f64> is encoded as NaN = None.  A cp_ctx is bound to ONE GPU and is not thread-safe: use
 * one ctx (and one host thread or process) per GPU, or a cp_multi (below), which owns one ctx per
 * device and drives them from the calling thread.  There is no CPU fallback: without a usable
 * CUDA device cp_init fails with CP_ERR_NO_DEVICE.
 */
#ifndef CP_GPU_H
#define CP_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CP_MAX_ITEMS 16 /* components per shape (polygon sides / atoms / LJ sites) */
#define CP_MAX_SYMS 4   /* general-position operators per plane group (wallpaper.rs:134-176) */
#define CP_MAX_DOF 6    /* length, ratio, angle, x, y, theta */
#define CP_MAX_STAGES 4 /* analyse_state chains 3 optimise_state calls (main.rs:224-252) */

typedef enum {
    CP_OK = 0,
    CP_ERR_INVALID_ARG = 1,
    CP_ERR_INVALID_STATE = 2, /* initial state overlaps: the reference panics (optimisation.rs:81-84) */
    CP_ERR_CUDA = 3,
    CP_ERR_NO_DEVICE = 4,
    CP_ERR_PARSE = 5, /* unknown group / malformed symmetry string */
    CP_ERR_NOT_PREPARED = 6
} cp_status;

/* which State/Shape pair (traits.rs:36-43,71-87) */
typedef enum {
    CP_POLYGON = 0, /* PackedState<LineShape>        (state/packed.rs, shape/line_shape.rs) */
    CP_DISCS = 1,   /* PackedState<MolecularShape2>  (state/packed.rs, shape/molecular_shape2.rs) */
    CP_LJ = 2       /* PotentialState<LJShape2>      (state/potential.rs, shape/lj_shape.rs) */
} cp_kind;

/* CrystalFamily (cell.rs:19-25) */
typedef enum { CP_MONOCLINIC = 0, CP_ORTHORHOMBIC = 1, CP_HEXAGONAL = 2, CP_TETRAGONAL = 3 } cp_family;

typedef enum {
    CP_RNG_PCG64MCG = 0, /* rand_pcg::Pcg64Mcg::seed_from_u64(replica) + rand 0.8 distributions:
                            reproduces the reference's random stream (optimisation.rs:86-93) */
    CP_RNG_PHILOX = 1    /* Philox4x32-10 keyed by the replica id: production stream,
                            statistically equivalent only */
} cp_rng;

/* Template state shared by all replicas: shape + plane group + initial DOFs.
 * Replaces PackedState::from_group / PotentialState::from_group (packed.rs:191-195,
 * potential.rs:148-152) as the thing handed to the optimiser. */
typedef struct {
    int32_t kind;    /* cp_kind */
    int32_t n_items; /* 1..CP_MAX_ITEMS */
    /* per component: POLYGON x0,y0,x1,y1,-  (Line2, line2.rs:16-20)
     *                DISCS   x,y,radius,-,- (Atom2, atom2.rs:14-18)
     *                LJ      x,y,sigma,epsilon,cutoff (NaN = None) (LJ2, lj2.rs:18-29) */
    double items[CP_MAX_ITEMS][5];
    double area;             /* shape.area() (line_shape.rs:60-72, molecular_shape2.rs:39-52); 0 for LJ */
    double enclosing_radius; /* shape.enclosing_radius() */
    int32_t family;          /* cp_family of the plane group */
    int32_t n_syms;          /* multiplicity of the general Wyckoff site */
    /* rows 0 and 1 of each 3x3 operator; row 2 is all zero as Transform2::from_operations leaves
     * it (transform.rs:139) */
    double syms[CP_MAX_SYMS][6];
    double dof[CP_MAX_DOF]; /* initial length, ratio, angle, x, y, theta */
} cp_problem;

/* MCOptimiser (optimisation.rs:14-22) minus the seed: the seed of replica k is
 * `seed_base + k` for every stage, as analyse_state seeds every stage with the index
 * (main.rs:229,239,249). */
typedef struct {
    double kt_start;
    double kt_ratio;
    double max_step_size;
    uint64_t steps;
    uint64_t inner_steps;
    double convergence; /* NaN = None */
} cp_schedule;

/* BuildOptimiser (crystal-packing-cli/src/main.rs:27-64); NaN = None */
typedef struct {
    uint64_t steps;
    double kt_start;
    double kt_finish;
    double kt_ratio;
    double max_step_size;
    uint64_t inner_steps;
    double convergence;
} cp_build_optimiser;

typedef struct {
    double dof[CP_MAX_DOF]; /* final length, ratio, angle, x, y, theta */
    double score;           /* State::score() of the final state; NaN = None */
    uint64_t rejections;    /* summed over stages (optimisation.rs:138) */
    uint64_t moves;         /* trial moves executed (early convergence exit shortens it) */
} cp_replica_result;

typedef struct {
    cp_replica_result result;
    uint64_t replica; /* global replica index (= seed) of the best state; of several replicas with the
                         maximal score the highest index, as rayon's max() returns (main.rs:253) */
} cp_best;

/* one replayed trial move (inputs): what optimisation.rs:104-119,62 drew for this move */
typedef struct {
    uint32_t basis_index;
    uint32_t pad_;
    double new_value; /* proposed value (optimisation.rs:115-119) */
    double threshold; /* rng.gen::<f64>() (optimisation.rs:62); ignored when out of bounds */
    double kt;
} cp_move;

/* one replayed trial move (outputs) */
typedef struct {
    uint32_t in_bounds; /* Basis::set_value Ok? (basis/mod.rs:25-37) */
    uint32_t accepted;
    double new_score;     /* state.score() after the proposal; NaN = None or not evaluated */
    double score_current; /* after the move */
} cp_decision;

/* device-side timing of the last launch (CUDA events on the ctx stream) */
typedef struct {
    float kernel_ms; /* MC kernel(s) only */
    float reduce_ms; /* best-of-replicas reduction */
    float h2d_ms;
    float d2h_ms;
    uint32_t launches; /* kernels launched by the last cp_launch/cp_run */
    uint32_t lanes_per_replica;
} cp_timing;

typedef struct cp_ctx cp_ctx;

/* ---- lifetime ---- */
cp_status cp_device_count(int *out);
cp_status cp_init(int device, cp_ctx **out);
void cp_destroy(cp_ctx *ctx);
const char *cp_last_error(void); /* thread-local message of the last failing call */
const char *cp_version(void);
/* launch on a caller-provided cudaStream_t (e.g. torch's current stream); NULL = ctx's own */
cp_status cp_set_stream(cp_ctx *ctx, void *cuda_stream);
/* lanes of a warp that cooperate on one replica: 0 = choose automatically; 1 = one thread per
 * replica (hard shapes; falls back to 8 when the geometry does not fit); 4/8/16/32 = lane groups */
cp_status cp_set_lanes(cp_ctx *ctx, int lanes);

/* ---- host-side constructors (no GPU work) ----
 * LineShape::polygon / from_radial (line_shape.rs:128-150) */
cp_status cp_shape_polygon(const double *radii, int n, cp_problem *out);
/* MolecularShape2::from_trimer / LJShape2::from_trimer (molecular_shape2.rs:141-162,
 * lj_shape.rs:111-131); kind = CP_DISCS or CP_LJ */
cp_status cp_shape_trimer(int kind, double radius, double angle, double distance, cp_problem *out);
/* MolecularShape2::circle / LJShape2::circle (molecular_shape2.rs:167-172, lj_shape.rs:146-151) */
cp_status cp_shape_circle(int kind, cp_problem *out);
/* Shape::area / Shape::enclosing_radius (line_shape.rs:60-72,86-93, molecular_shape2.rs:39-52,66-74,
 * lj_shape.rs:52-60) recomputed from `items` — for shapes filled in by hand or read from JSON */
cp_status cp_shape_finish(cp_problem *shape);
/* Transform2::from_operations (transform.rs:124-183): rows 0-1 of the operator */
cp_status cp_parse_operations(const char *sym_ops, double out[6]);
/* WallpaperGroups::from_str + TryFrom (wallpaper.rs:100-110,134-176) */
cp_status cp_group_lookup(const char *name, int32_t *family, int32_t *n_syms,
                          double syms[CP_MAX_SYMS][6]);
/* PackedState::from_group / PotentialState::from_group on a problem whose shape fields are set
 * (packed.rs:169-195, potential.rs:148-174, cell.rs:200-214, site.rs:51-63) */
cp_status cp_problem_from_group(cp_problem *problem, const char *group);
/* BuildOptimiser::build (main.rs:122-143) */
cp_status cp_build_schedule(const cp_build_optimiser *opts, cp_schedule *out);
/* the three stages analyse_state chains per replica (main.rs:224-252) */
cp_status cp_pipeline_schedules(const cp_build_optimiser *opts, cp_schedule out[3]);

/* ---- the hot path ----
 * cp_run replaces the body of analyse_state (main.rs:221-253): for replicas
 * [replica_begin, replica_begin + n_replicas) run every stage of `stages` through
 * MCOptimiser::optimise_state (optimisation.rs:80-180) on the GPU and take the max by score.
 * init_dof: NULL -> every replica starts from problem->dof; else n_replicas x 6 host doubles.
 * out_all: NULL or n_replicas host entries.  Host buffers; copies are part of the call. */
cp_status cp_run(cp_ctx *ctx, const cp_problem *problem, const cp_schedule *stages, int n_stages,
                 int rng_mode, uint64_t seed_base, uint64_t replica_begin, uint64_t n_replicas,
                 const double *init_dof, cp_best *out_best, cp_replica_result *out_all);

/* the same split in three so a caller can keep everything resident in HBM:
 * prepare = allocate + upload, launch = kernels only (asynchronous on the ctx stream),
 * fetch = synchronise + copy back.
 * Buffer lifetimes the split leaves to the caller:
 *  - cp_prepare ENQUEUES the copy of init_dof on the ctx stream and returns: a pinned init_dof must
 *    stay valid and unchanged until the stream has passed the copy (cp_synchronize, cp_fetch, or
 *    any later synchronisation of that stream); pageable memory is staged before the call returns.
 *  - every cp_launch overwrites the device results and the device cp_best of the previous launch;
 *    work on OTHER streams that reads them (a collective on the communicator's stream, a peer copy)
 *    must be ordered before the next cp_launch by the caller (event or stream wait).
 *  - the addresses returned by cp_device_buffers go stale when a cp_prepare with more replicas than
 *    any before it reallocates the result buffer; ask again after every cp_prepare.
 * Tuning knobs read from the environment by cp_prepare (measurement only; results never depend on
 * them): CP_WARP_FILL = replicas per warp of the one-thread-per-replica kernels (default: chosen at
 * the launch from the occupancy), CP_DRAW_CAP = proposals a lane may draw per pass (default 3). */
cp_status cp_prepare(cp_ctx *ctx, const cp_problem *problem, const cp_schedule *stages,
                     int n_stages, int rng_mode, uint64_t seed_base, uint64_t replica_begin,
                     uint64_t n_replicas, const double *init_dof);
cp_status cp_launch(cp_ctx *ctx);
cp_status cp_fetch(cp_ctx *ctx, cp_best *out_best, cp_replica_result *out_all);
cp_status cp_synchronize(cp_ctx *ctx);
/* device addresses of the prepared job's outputs, valid until the next cp_prepare / cp_destroy:
 * d_results = n_replicas x cp_replica_result, d_best = one cp_best written by cp_launch.  Lets a
 * multi-GPU host hand the best state to its collective (NCCL) without a host round trip. */
cp_status cp_device_buffers(cp_ctx *ctx, void **d_results, void **d_best);
cp_status cp_last_timing(cp_ctx *ctx, cp_timing *out);

/* ---- several GPUs behind one handle ----
 * The replica loop of analyse_state (main.rs:221-253: `(0..replications).into_par_iter()...max()`)
 * over a list of CUDA devices: [replica_begin, replica_begin + n_replicas) is cut into contiguous
 * shards, one per entry of device_ids (an id may repeat: two contexts, two streams on that GPU);
 * all devices run at the same time, driven by the calling thread; the per-device maxima (80 bytes
 * each, already on the host) are reduced on the host by the rule of cp_best.  Seeds are global
 * replica indices, so the answer is the one a single device gives.  init_dof / out_all cover the
 * whole range, as in cp_run.  Not thread-safe, like cp_ctx. */
typedef struct cp_multi cp_multi;
cp_status cp_init_multi(const int *device_ids, int n_devices, cp_multi **out);
void cp_multi_destroy(cp_multi *m);
cp_status cp_multi_device_count(const cp_multi *m, int *out);
/* the k-th device's context, owned by the handle (cp_last_timing, cp_set_lanes, cp_score ...) */
cp_status cp_multi_context(cp_multi *m, int k, cp_ctx **out);
/* the shard the k-th device ran in the last cp_multi_run */
cp_status cp_multi_shard(const cp_multi *m, int k, uint64_t *replica_begin, uint64_t *n_replicas);
cp_status cp_multi_run(cp_multi *m, const cp_problem *problem, const cp_schedule *stages, int n_stages,
                       int rng_mode, uint64_t seed_base, uint64_t replica_begin, uint64_t n_replicas,
                       const double *init_dof, cp_best *out_best, cp_replica_result *out_all);

/* Measured FP64 / FP32 FMA throughput of this GPU (2 flops per FMA, register-resident chains):
 * the roofline denominators bench.py reports against (no reference counterpart). */
cp_status cp_measure_peaks(cp_ctx *ctx, double *fp64_tflops, double *fp32_tflops);

/* State::score for a batch of states (packed.rs:80-86, potential.rs:82-115): dof = n x 6 host
 * doubles, out_scores = n host doubles (NaN = None). */
cp_status cp_score(cp_ctx *ctx, const cp_problem *problem, const double *dof, uint64_t n,
                   double *out_scores);

/* Replay dumped move sequences: replica r starts from init_dof[r] (bounds captured from it as
 * generate_basis does, cell.rs:136-170), applies moves[r*n_moves + m] in order and reports each
 * decision of the inner-loop body optimisation.rs:104-136. */
cp_status cp_replay(cp_ctx *ctx, const cp_problem *problem, const double *init_dof,
                    const cp_move *moves, uint64_t n_moves, uint64_t n_replicas,
                    cp_decision *out);

/* The same with the bounds taken from another state: bound_dof (n_replicas x 6, NULL = init_dof)
 * holds the DOFs at the time generate_basis() captured the bounds, i.e. at the start of the stage
 * the moves were dumped from (cell.rs:139-152: length and ratio are bounded by their values then).
 * Lets a dump from the middle of a stage be replayed from the state it was taken at. */
cp_status cp_replay_bounded(cp_ctx *ctx, const cp_problem *problem, const double *init_dof,
                            const double *bound_dof, const cp_move *moves, uint64_t n_moves,
                            uint64_t n_replicas, cp_decision *out);

/* OccupiedSite::positions and Cell2::to_cartesian_isometry for a batch of states (site.rs:33-45,
 * cell.rs:110-112,258-263): dof = n x 6 host doubles; out = n x n_syms x 8 host doubles, per
 * symmetry image the wrapped fractional position (fx, fy), the Cartesian position (cx, cy) and
 * rows 0-1 of the rotation block (m00, m01, m10, m11) of sym * Transform2::new(theta, (x, y)),
 * computed by the device functions the Monte-Carlo kernels use. */
cp_status cp_images(cp_ctx *ctx, const cp_problem *problem, const double *dof, uint64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif /* CP_GPU_H */
