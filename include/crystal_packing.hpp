// crystal_packing.hpp — header-only C++ host layer over the C ABI of libcpgpu.so (cp_gpu.h).
//
// The reference is compiled code (Rust) whose toolchain is absent from the build image, so the
// host side above the C ABI is mirrored here in C++ with the crate's own vocabulary: the same type
// and function names, argument meaning and error behaviour (Result::Err -> cp::Error, panic ->
// cp::Error with CP_ERR_INVALID_STATE).  Citations are relative to the reference tree.
//
//   cp::LineShape::polygon(5)                       crystal-packing/src/shape/line_shape.rs:148-150
//   cp::MolecularShape2::from_trimer(r, a, d)       crystal-packing/src/shape/molecular_shape2.rs:141-162
//   cp::LJShape2::circle()                          crystal-packing/src/shape/lj_shape.rs:146-151
//   cp::PackedState::from_group(shape, "p2")        crystal-packing/src/state/packed.rs:191-195
//   state.score(engine)                             crystal-packing/src/traits.rs:72
//   cp::MCOptimiser(...).optimise_state(state, e)   crystal-packing/src/optimisation.rs:25-33,80-180
//   cp::BuildOptimiser{...}.build(seed)             crystal-packing-cli/src/main.rs:122-143
//   cp::monte_carlo_best_packing(state, opts, n, e) GPU drop-in for analyse_state, main.rs:215-269
//
// All Monte-Carlo arithmetic runs in the CUDA kernels behind the C ABI; nothing here computes a
// score on the CPU.
#ifndef CRYSTAL_PACKING_HPP
#define CRYSTAL_PACKING_HPP

#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "cp_gpu.h"

namespace cp {

struct Error : std::runtime_error {
    cp_status status;
    Error(cp_status st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};

inline void check(cp_status st) {
    if (st != CP_OK) throw Error(st, cp_last_error());
}

inline double opt(const std::optional<double> &v) { return v ? *v : std::nan(""); }

// ---- shapes (Shape + Intersect / Potential, traits.rs:36-65)
struct Shape {
    std::string name;
    cp_problem problem{};  // only the shape fields are meaningful
    double area() const { return problem.area; }
    double enclosing_radius() const { return problem.enclosing_radius; }
    int len() const { return problem.n_items; }
};

struct LineShape : Shape {
    static LineShape from_radial(const std::string &name, const std::vector<double> &points) {
        LineShape s;
        s.name = name;
        check(cp_shape_polygon(points.data(), static_cast<int>(points.size()), &s.problem));
        return s;
    }
    static LineShape polygon(std::size_t sides) { return from_radial("Polygon", std::vector<double>(sides, 1.0)); }
};

struct MolecularShape2 : Shape {
    static MolecularShape2 from_trimer(double radius, double angle, double distance) {
        MolecularShape2 s;
        s.name = "Trimer";
        check(cp_shape_trimer(CP_DISCS, radius, angle, distance, &s.problem));
        return s;
    }
    static MolecularShape2 circle() {
        MolecularShape2 s;
        s.name = "circle";
        check(cp_shape_circle(CP_DISCS, &s.problem));
        return s;
    }
};

struct LJShape2 : Shape {
    static LJShape2 from_trimer(double radius, double angle, double distance) {
        LJShape2 s;
        s.name = "Trimer";
        check(cp_shape_trimer(CP_LJ, radius, angle, distance, &s.problem));
        return s;
    }
    static LJShape2 circle() {
        LJShape2 s;
        s.name = "circle";
        check(cp_shape_circle(CP_LJ, &s.problem));
        return s;
    }
};

// ---- one GPU (cp_ctx)
class Engine {
  public:
    explicit Engine(int device = 0) { check(cp_init(device, &ctx_)); }
    ~Engine() { cp_destroy(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    cp_ctx *raw() const { return ctx_; }

  private:
    cp_ctx *ctx_ = nullptr;
};

// ---- a list of GPUs behind one handle (cp_multi): the replica loop of analyse_state sharded over them
class MultiEngine {
  public:
    explicit MultiEngine(const std::vector<int> &devices) {
        check(cp_init_multi(devices.data(), static_cast<int>(devices.size()), &m_));
    }
    ~MultiEngine() { cp_multi_destroy(m_); }
    MultiEngine(const MultiEngine &) = delete;
    MultiEngine &operator=(const MultiEngine &) = delete;
    cp_multi *raw() const { return m_; }
    int devices() const {
        int n = 0;
        check(cp_multi_device_count(m_, &n));
        return n;
    }

  private:
    cp_multi *m_ = nullptr;
};

// ---- State (traits.rs:71-87): PackedState<S> and PotentialState<S> share this shape
struct State {
    cp_problem problem{};
    std::string wallpaper;

    int total_shapes() const { return problem.n_syms; }
    // State::score(): Option<f64>
    std::optional<double> score(const Engine &e) const {
        double s = 0.;
        check(cp_score(e.raw(), &problem, problem.dof, 1, &s));
        return std::isnan(s) ? std::nullopt : std::optional<double>(s);
    }
    // (min, max) per degree of freedom in the reference's basis order (cell.rs:136-170, site.rs:65-91)
    std::vector<std::pair<double, double>> generate_basis() const {
        const double pi = 3.14159265358979323846;
        std::vector<std::pair<double, double>> b{{0.01, problem.dof[0]}};
        if (problem.family == CP_MONOCLINIC || problem.family == CP_ORTHORHOMBIC) b.push_back({0.1, problem.dof[1]});
        if (problem.family == CP_MONOCLINIC) b.push_back({pi / 4., pi / 2.});
        b.push_back({-0.5, 0.5});
        b.push_back({-0.5, 0.5});
        b.push_back({0., 2. * pi});
        return b;
    }

  protected:
    static State build(const Shape &shape, const std::string &group) {
        State s;
        s.problem = shape.problem;
        s.wallpaper = group;
        check(cp_problem_from_group(&s.problem, group.c_str()));
        return s;
    }
};

struct PackedState : State {
    static PackedState from_group(const LineShape &shape, const std::string &group) { return {build(shape, group)}; }
    static PackedState from_group(const MolecularShape2 &shape, const std::string &group) {
        return {build(shape, group)};
    }
};
struct PotentialState : State {
    static PotentialState from_group(const LJShape2 &shape, const std::string &group) { return {build(shape, group)}; }
};

// ---- optimiser
struct MCOptimiser {
    cp_schedule schedule{};
    std::uint64_t seed = 0;
    MCOptimiser(double kt_start, double kt_ratio, double max_step_size, std::uint64_t steps,
                std::uint64_t inner_steps, std::uint64_t seed_, std::optional<double> convergence)
        : schedule{kt_start, kt_ratio, max_step_size, steps, inner_steps, opt(convergence)}, seed(seed_) {}

    // runs the Monte-Carlo loop for this one state on the GPU with the reference's random stream
    template <class S>
    S optimise_state(S state, const Engine &e) const {
        cp_best best{};
        check(cp_run(e.raw(), &state.problem, &schedule, 1, CP_RNG_PCG64MCG, seed, 0, 1, nullptr, &best, nullptr));
        for (int i = 0; i < CP_MAX_DOF; ++i) state.problem.dof[i] = best.result.dof[i];
        return state;
    }
};

struct BuildOptimiser {  // defaults of BuildOptimiser::default() (main.rs:66-79)
    std::uint64_t steps = 1000;
    double kt_start = 0.1;
    std::optional<double> kt_finish = 0.001;
    std::optional<double> kt_ratio;
    double max_step_size = 0.01;
    std::uint64_t inner_steps = 1000;
    std::optional<double> convergence;

    cp_build_optimiser raw() const {
        return {steps, kt_start, opt(kt_finish), opt(kt_ratio), max_step_size, inner_steps, opt(convergence)};
    }
    MCOptimiser build(std::uint64_t seed) const {
        cp_build_optimiser b = raw();
        cp_schedule s{};
        check(cp_build_schedule(&b, &s));
        return MCOptimiser(s.kt_start, s.kt_ratio, s.max_step_size, s.steps, s.inner_steps, seed,
                           std::isnan(s.convergence) ? std::nullopt : std::optional<double>(s.convergence));
    }
};

struct BestPacking {
    cp_best best{};
    std::vector<cp_replica_result> replicas;  // filled when requested
};

// the replica-parallel driver: `replications` independent replicas (seed = replica index) through
// the three-stage pipeline of analyse_state, then the max by score.  Returns the optimised state.
template <class S>
S monte_carlo_best_packing(S state, const BuildOptimiser &opts, std::uint64_t replications, const Engine &e,
                           int rng_mode = CP_RNG_PCG64MCG, BestPacking *details = nullptr,
                           std::uint64_t replica_begin = 0) {
    cp_build_optimiser b = opts.raw();
    cp_schedule stages[3];
    check(cp_pipeline_schedules(&b, stages));
    BestPacking local;
    BestPacking &d = details ? *details : local;
    if (details) d.replicas.resize(replications);
    check(cp_run(e.raw(), &state.problem, stages, 3, rng_mode, 0, replica_begin, replications, nullptr, &d.best,
                 details ? d.replicas.data() : nullptr));
    if (std::isnan(d.best.result.score)) throw Error(CP_ERR_INVALID_STATE, "Error in running optimisation.");
    for (int i = 0; i < CP_MAX_DOF; ++i) state.problem.dof[i] = d.best.result.dof[i];
    return state;
}

// the same over several GPUs: replicas are sharded contiguously, the answer is the single-GPU one
template <class S>
S monte_carlo_best_packing(S state, const BuildOptimiser &opts, std::uint64_t replications, const MultiEngine &e,
                           int rng_mode = CP_RNG_PCG64MCG, BestPacking *details = nullptr,
                           std::uint64_t replica_begin = 0) {
    cp_build_optimiser b = opts.raw();
    cp_schedule stages[3];
    check(cp_pipeline_schedules(&b, stages));
    BestPacking local;
    BestPacking &d = details ? *details : local;
    if (details) d.replicas.resize(replications);
    check(cp_multi_run(e.raw(), &state.problem, stages, 3, rng_mode, 0, replica_begin, replications, nullptr, &d.best,
                       details ? d.replicas.data() : nullptr));
    if (std::isnan(d.best.result.score)) throw Error(CP_ERR_INVALID_STATE, "Error in running optimisation.");
    for (int i = 0; i < CP_MAX_DOF; ++i) state.problem.dof[i] = d.best.result.dof[i];
    return state;
}

}  // namespace cp
#endif  // CRYSTAL_PACKING_HPP
