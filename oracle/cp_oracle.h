/*
 * cp_oracle.h — CPU oracle for the crystal-packing Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference algorithm
 * (malramsay64/crystal-packing, Rust) used as the checker for the CUDA path and as the timed CPU
 * baseline in bench.py.  Nothing in the shipped library (crystal-packing_b200/) links, loads or
 * calls it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * PARITY STATUS: "parity unpinned" at trajectory level — the Rust reference cannot be built in
 * this environment (no rustc/cargo, crates not vendored) and its own tests hold no RNG / trajectory
 * / final-score golden values.  What IS pinned (tests/test_oracle_*.py): every known-answer test
 * of the reference for this path, the upstream rand_pcg known-answer vectors, and
 * the anchors of an independent scratch restatement.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef CP_ORACLE_H
#define CP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ITEMS 64
#define ORC_MAX_SYMS 8
#define ORC_MAX_DOF 6

enum { ORC_POLYGON = 0, ORC_DISCS = 1, ORC_LJ = 2 };
enum { ORC_MONOCLINIC = 0, ORC_ORTHORHOMBIC = 1, ORC_HEXAGONAL = 2, ORC_TETRAGONAL = 3 };
/* which sin/cos/exp the oracle calls: the platform libm (what the Rust reference does) or the
 * deterministic cp_math.h versions that the CUDA kernels execute (bit-exact GPU comparison) */
enum { ORC_MATH_LIBM = 0, ORC_MATH_CP = 1 };

typedef struct {
    double m[3][3];
} orc_transform;

/* one shape component; which fields are meaningful depends on orc_state.kind:
 *   POLYGON: a=x0 b=y0 c=x1 d=y1            (Line2, shape/components/line2.rs:16-20)
 *   DISCS:   a=x  b=y  c=radius              (Atom2, shape/components/atom2.rs:14-18)
 *   LJ:      a=x  b=y  c=sigma d=epsilon e=cutoff (NaN = None)  (LJ2, components/lj2.rs:18-29) */
typedef struct {
    double a, b, c, d, e;
} orc_item;

typedef struct {
    int32_t kind;
    int32_t n_items;
    orc_item items[ORC_MAX_ITEMS];
    int32_t family;
    int32_t n_syms;
    orc_transform syms[ORC_MAX_SYMS];
    double length, ratio, angle; /* Cell2 (cell.rs:48-54) */
    double x, y, theta;          /* OccupiedSite (site.rs:13-19) */
    int32_t math_mode;
    int32_t pad_;
} orc_state;

/* MCOptimiser (optimisation.rs:14-22); convergence NaN = None */
typedef struct {
    double kt_start;
    double kt_ratio;
    double max_step_size;
    uint64_t steps;
    uint64_t inner_steps;
    uint64_t seed;
    double convergence;
} orc_optimiser;

/* BuildOptimiser (crystal-packing-cli/src/main.rs:27-64); NaN = None for the Option<f64> fields */
typedef struct {
    uint64_t steps;
    double kt_start;
    double kt_finish;
    double kt_ratio;
    double max_step_size;
    uint64_t inner_steps;
    double convergence;
} orc_build_optimiser;

/* one trial move as dumped by orc_optimise_state (one per inner-loop iteration) */
typedef struct {
    uint32_t basis_index;
    uint32_t in_bounds;
    uint32_t accepted;
    uint32_t pad_;
    double old_value;
    double new_value;
    double threshold;     /* NaN when the proposal was out of bounds (no draw) */
    double new_score;     /* NaN = None (overlap) or not evaluated */
    double score_current; /* after the move */
    double kt;
    double step_ratio;
} orc_move_record;

typedef struct {
    double score;       /* NaN = None */
    double dof[ORC_MAX_DOF]; /* length, ratio, angle, x, y, theta (always all six) */
    uint64_t rejections;
    uint64_t moves;     /* trial moves executed */
} orc_result;

/* executed reference-formula operation counts (SURVEY.md §8d unit costs), per calling thread */
typedef struct {
    uint64_t moves;          /* trial moves (inner-loop iterations) */
    uint64_t scored;         /* moves that reached state.score() */
    uint64_t segment_tests;  /* Line2::intersects calls */
    uint64_t disc_tests;     /* Atom2::intersects calls */
    uint64_t lj_in, lj_out;  /* LJ2::energy inside / outside the cutoff (no cutoff counts as in) */
    uint64_t image_cands;    /* periodic image candidates generated */
    uint64_t image_pass;     /* candidates passing the distance filter */
    uint64_t point_xforms;   /* Transform2 * Point2 evaluations */
    uint64_t site_images;    /* (sym * transform).periodic() evaluations */
    uint64_t trig;           /* sin / cos / exp evaluations */
} orc_counters;
void orc_counters_reset(void);
void orc_counters_get(orc_counters *out);
/* FLOPs under the unit costs of SURVEY.md §8d */
double orc_counters_flops(const orc_counters *c);

/* ---- RNG (rand_pcg 0.3.1 / rand 0.8.4 / rand_core 0.6.3) ---- */
typedef struct {
    uint64_t lo, hi;
} orc_pcg;
void orc_pcg_new(orc_pcg *r, uint64_t state_lo, uint64_t state_hi);
void orc_pcg_from_seed(orc_pcg *r, const uint8_t seed[16]);
void orc_pcg_seed_from_u64(orc_pcg *r, uint64_t seed);
uint64_t orc_pcg_next_u64(orc_pcg *r);
uint64_t orc_uniform_usize(orc_pcg *r, uint64_t range);
double orc_uniform_step(orc_pcg *r);
double orc_gen_f64(orc_pcg *r);

/* ---- math wrappers ---- */
void orc_sincos(int math_mode, double x, double *s, double *c);
double orc_exp(int math_mode, double x);
void orc_cp_sincos(double x, double *s, double *c);
double orc_cp_exp(double x);

/* ---- Transform2 (transform.rs) ---- */
void orc_transform_new(int math_mode, double rotation, double tx, double ty, orc_transform *out);
void orc_transform_identity(orc_transform *out);
void orc_transform_mul(const orc_transform *a, const orc_transform *b, orc_transform *out);
void orc_transform_mul_point(const orc_transform *t, double px, double py, double *ox, double *oy);
void orc_transform_position(const orc_transform *t, double *ox, double *oy);
void orc_transform_periodic(const orc_transform *t, double period, double offset, orc_transform *out);
int orc_transform_from_operations(const char *ops, orc_transform *out);

/* ---- Cell2 (cell.rs) ---- */
void orc_cell_to_cartesian(const orc_state *s, double x, double y, double *ox, double *oy);
double orc_cell_area(const orc_state *s);
/* periodic_images(transform, shells, zero) -> out[(2*shells+1)^2 (-1)]; returns count */
int orc_cell_periodic_images(const orc_state *s, const orc_transform *t, int shells, int zero,
                             orc_transform *out);

/* ---- shapes ---- */
int orc_line_intersects(const orc_item *s, const orc_item *o);
int orc_atom_intersects(const orc_item *s, const orc_item *o);
double orc_lj_energy(const orc_item *s, const orc_item *o);
int orc_shape_from_radial(const double *radii, int n, orc_item *items);
void orc_shape_trimer(int kind, double radius, double angle, double distance, orc_item *items);
void orc_shape_circle(int kind, orc_item *items);
double orc_shape_area(int kind, const orc_item *items, int n);
double orc_shape_enclosing_radius(int kind, const orc_item *items, int n);
void orc_shape_transform(int kind, const orc_item *items, int n, const orc_transform *t,
                         orc_item *out);
int orc_shape_intersects(int kind, const orc_item *a, const orc_item *b, int n);
double orc_shape_energy(const orc_item *a, const orc_item *b, int n);
double orc_circle_overlap(const orc_item *a1, const orc_item *a2);

/* ---- wallpaper groups (wallpaper.rs:134-176) ---- */
/* name in {"p1","p2","p1m1","p1g1","pg","p2mm","p2mg","p2gg"}; returns 0 ok */
int orc_group(const char *name, int32_t *family, int32_t *n_syms, orc_transform *syms);

/* ---- states ---- */
/* from_group + initialise (packed.rs:169-195 / potential.rs:148-174) */
int orc_state_from_group(orc_state *s, int kind, const orc_item *items, int n_items,
                         const char *group, int math_mode);
int orc_state_relative_positions(const orc_state *s, orc_transform *out);
int orc_state_cartesian_positions(const orc_state *s, orc_transform *out);
int orc_state_check_intersection(const orc_state *s);
double orc_state_score(const orc_state *s); /* NaN = None */
int orc_state_total_shapes(const orc_state *s);
int orc_state_n_dof(const orc_state *s);

/* ---- optimiser ---- */
double orc_energy_surface(int math_mode, double new_score, double old_score, double kt);
/* optimise_state (optimisation.rs:80-180).  records may be NULL; at most max_records written,
 * *n_records = number of trial moves executed.  Returns 0 ok, 1 = invalid initial state (the
 * reference panics), 2 = final state invalid (reference assert). */
int orc_optimise_state(const orc_optimiser *opt, orc_state *s, orc_move_record *records,
                       uint64_t max_records, uint64_t *n_records, uint64_t *rejections);
/* BuildOptimiser::build (main.rs:122-143) */
void orc_build(const orc_build_optimiser *b, uint64_t seed, orc_optimiser *out);
/* the three-stage pipeline of analyse_state for ONE replica (main.rs:224-252) */
int orc_run_replica(const orc_state *tmpl, const orc_build_optimiser *b, uint64_t index,
                    orc_state *out_state, orc_result *out);
/* all replicas [begin, begin+n) on n_threads host threads + .max() (main.rs:221-253);
 * results (nullable) gets one entry per replica; returns index (relative) of the best */
int64_t orc_analyse_state(const orc_state *tmpl, const orc_build_optimiser *b, uint64_t begin,
                          uint64_t n, int n_threads, orc_result *results, orc_result *best);
/* same, also returning the operation counters summed over all worker threads (nullable) */
int64_t orc_analyse_state_counted(const orc_state *tmpl, const orc_build_optimiser *b,
                                  uint64_t begin, uint64_t n, int n_threads, orc_result *results,
                                  orc_result *best, orc_counters *counters);
/* set the six DOFs / read them back (length, ratio, angle, x, y, theta) */
void orc_state_set_dof(orc_state *s, const double dof[ORC_MAX_DOF]);
void orc_state_get_dof(const orc_state *s, double dof[ORC_MAX_DOF]);
/* one replayed move: apply `new_value` to basis `basis_index` of `s` (bounds from `lo/hi`),
 * score, decide with `threshold`; mirrors the inner-loop body optimisation.rs:104-136.
 * Returns accepted flag; *in_bounds, *new_score (NaN=None), *score_current updated. */
int orc_replay_move(orc_state *s, int basis_index, double new_value, double lo, double hi,
                    double threshold, double kt, double *score_current, int *in_bounds,
                    double *new_score);
/* bounds as generate_basis() would capture them now (cell.rs:136-170, site.rs:65-91) */
int orc_state_bounds(const orc_state *s, double lo[ORC_MAX_DOF], double hi[ORC_MAX_DOF]);

#ifdef __cplusplus
}
#endif
#endif
