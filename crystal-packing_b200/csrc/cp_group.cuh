// cp_group.cuh — kernels with one group of L lanes (4, 8, 16 or a whole warp of 32) per replica.
// Device functions are in cp_device.cuh; this file holds the shared-memory layout and the three
// kernels (Monte-Carlo run, batch score, replay).  Instantiated per shape kind in cp_group_*.cu.
#ifndef CP_GROUP_CUH
#define CP_GROUP_CUH

#include <cuda_runtime.h>

#include "cp_device.cuh"

constexpr int kBlock = 128;

struct SharedLayout {
    int items_off, shift_off, group_off, group_stride, img_doubles;
    size_t bytes;
};

__host__ __device__ inline SharedLayout shared_layout(int kind, int n_items, int n_syms, int lanes) {
    SharedLayout s;
    const int cw = kind == CP_POLYGON ? 8 : 4;
    s.items_off = 0;
    s.shift_off = n_items * 5;
    s.group_off = s.shift_off + n_items;
    s.img_doubles = n_syms * 8;
    s.group_stride = s.img_doubles + n_syms * n_items * cw;
    s.bytes = sizeof(double) * (size_t)(s.group_off + (kBlock / lanes) * s.group_stride);
    return s;
}

template <int KIND, int L>
__device__ __forceinline__ Geo stage_shared(const KProblem &P, double *smem, const double *&s_items,
                                            const double *&s_shift) {
    const SharedLayout lay = shared_layout(KIND, P.n_items, P.n_syms, L);
    for (int i = threadIdx.x; i < P.n_items * 5; i += kBlock) smem[lay.items_off + i] = P.items[i / 5][i % 5];
    for (int i = threadIdx.x; i < P.n_items; i += kBlock) smem[lay.shift_off + i] = P.lj_shift[i];
    __syncthreads();
    s_items = smem + lay.items_off;
    s_shift = smem + lay.shift_off;
    Geo g;
    g.img = smem + lay.group_off + (threadIdx.x / L) * lay.group_stride;
    g.comp = g.img + lay.img_doubles;
    return g;
}

// The Monte-Carlo optimisation of one replica per group: every stage of MCOptimiser::optimise_state
// (optimisation.rs:80-180) chained as analyse_state does (main.rs:224-252), annealing schedule,
// step-size adaptation and convergence counter on the device.
template <int KIND, int L, int RNG>
__global__ void __launch_bounds__(kBlock) mc_kernel(const __grid_constant__ KParams K) {
    extern __shared__ double smem[];
    const KProblem &P = K.prob;
    const double *s_items, *s_shift;
    Geo g = stage_shared<KIND, L>(P, smem, s_items, s_shift);
    const int lane = threadIdx.x % L;
    const unsigned gmask = group_mask<L>();
    const unsigned long long rep = (unsigned long long)blockIdx.x * (kBlock / L) + threadIdx.x / L;
    if (rep >= K.n_replicas) return;
    const unsigned long long replica = K.replica_begin + rep;

    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = K.init_dof ? K.init_dof[rep * CP_MAX_DOF + i] : P.dof[i];

    const int n_cell = n_cell_dof(P.family);
    const int n_basis = n_cell + 3;
    unsigned long long rejections = 0, moves = 0;
    Trig trig = trig_of(d);
    double score_current = state_score<KIND, L>(P, s_items, s_shift, g, d, trig, gmask, lane);
    bool valid = score_current == score_current;
    if (!valid && lane == 0) atomicExch(K.error_flag, 1);  // panic!("Invalid configuration ...")

    Rng rng;
    for (int stage = 0; valid && stage < K.n_stages; ++stage) {
        const KStage S = K.stages[stage];
        // generate_basis(): the upper bounds of length and ratio are their values now (cell.rs:139-152)
        const double len0 = d.v[0], ratio0 = d.v[1];
        rng_begin_stage<RNG>(rng, K.seed_base, replica, n_basis);
        double kt = S.kt_start, step_ratio = 1.0;
        int convergence_count = 0;
        const unsigned long long outer = S.inner ? S.steps / S.inner : 0ull;
        bool converged = false;
        for (unsigned long long loop = 1; loop <= outer && !converged; ++loop) {
            const double score_start = score_current;
            unsigned long long loop_rejections = 0;
            for (unsigned long long it = 0; it < S.inner; ++it) {
                int basis_index;
                double sample;
                rng_draw_move<RNG>(rng, basis_index, sample);
                const int slot = basis_to_slot(basis_index, n_cell);
                const double val = d.get(slot);
                const double new_val = val + S.max_step * step_ratio * sample;  // optimisation.rs:115-119
                double lo, hi;
                slot_bounds(slot, len0, ratio0, lo, hi);
                if (!(lo <= new_val && new_val <= hi)) {  // Basis::set_value Err (basis/mod.rs:31-35)
                    ++loop_rejections;
                    continue;
                }
                // For hard shapes Some(new_score) depends on the cell DOFs only, so the Metropolis
                // rule can be evaluated before the overlap test; when it would reject even a valid
                // state the answer is "rejected" whatever check_intersection returns, and the test
                // is skipped.  Draw order and every decision are unchanged (see cp_thread.cuh).
                double threshold = 0., s_ = 0., c_ = 0., known_score = score_current;
                bool prechecked = false;
                if (KIND != CP_LJ) {
                    threshold = rng_draw_threshold<RNG>(rng);
                    prechecked = true;
                    if (slot <= 2) {
                        double sa_new = trig.sa;
                        if (slot == 2) {
                            cp_sincos(new_val, &s_, &c_);
                            sa_new = s_;
                        }
                        const double len = slot == 0 ? new_val : d.v[0], ratio = slot == 1 ? new_val : d.v[1];
                        known_score = (P.area * (double)P.n_syms) / (sa_new * len * (len * ratio));
                        if (!accept_move(known_score, score_current, kt, threshold)) {
                            ++loop_rejections;
                            continue;
                        }
                    }
                }
                d.set(slot, new_val);
                double old_s = 0., old_c = 0.;
                if (slot == 5) {  // the only DOFs that enter sin/cos are the site and cell angles
                    old_s = trig.st; old_c = trig.ct;
                    cp_sincos(new_val, &trig.st, &trig.ct);
                }
                if (slot == 2) {
                    old_s = trig.sa; old_c = trig.ca;
                    if (KIND == CP_LJ) cp_sincos(new_val, &s_, &c_);
                    trig.sa = s_; trig.ca = c_;
                }
                const double new_score = state_score<KIND, L>(P, s_items, s_shift, g, d, trig, gmask, lane);
                bool accepted;
                if (prechecked) {
                    accepted = new_score == new_score;  // Some(_): the rule itself already said yes
                } else {
                    threshold = rng_draw_threshold<RNG>(rng);
                    accepted = accept_move(new_score, score_current, kt, threshold);
                }
                if (accepted) {
                    score_current = new_score;
                } else {
                    d.set(slot, val);
                    if (slot == 5) { trig.st = old_s; trig.ct = old_c; }
                    if (slot == 2) { trig.sa = old_s; trig.ca = old_c; }
                    ++loop_rejections;
                }
            }
            moves += S.inner;
            rejections += loop_rejections;
            kt *= S.kt_ratio;
            if (S.convergence == S.convergence) {  // Some(precision) (optimisation.rs:142-159)
                if (score_current - score_start < S.convergence) {
                    if (++convergence_count > 5) converged = true;
                } else {
                    convergence_count = 0;
                }
            }
            if (!converged && step_ratio > 1e-4)  // optimisation.rs:165-167
                step_ratio *= (double)S.inner / ((double)loop_rejections + 1.);
        }
    }
    if (lane == 0) {
        cp_replica_result r;
#pragma unroll
        for (int i = 0; i < CP_MAX_DOF; ++i) r.dof[i] = d.v[i];
        r.score = score_current;
        r.rejections = rejections;
        r.moves = moves;
        K.out[rep] = r;
    }
}

template <int KIND, int L>
__global__ void __launch_bounds__(kBlock)
score_kernel(const __grid_constant__ KProblem P, const double *dof, unsigned long long n, double *out) {
    extern __shared__ double smem[];
    const double *s_items, *s_shift;
    Geo g = stage_shared<KIND, L>(P, smem, s_items, s_shift);
    const int lane = threadIdx.x % L;
    const unsigned gmask = group_mask<L>();
    const unsigned long long rep = (unsigned long long)blockIdx.x * (kBlock / L) + threadIdx.x / L;
    if (rep >= n) return;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = dof[rep * CP_MAX_DOF + i];
    double s = state_score<KIND, L>(P, s_items, s_shift, g, d, trig_of(d), gmask, lane);
    if (lane == 0) out[rep] = s;
}

// the inner-loop body optimisation.rs:104-136 driven by recorded draws
template <int KIND, int L>
__global__ void __launch_bounds__(kBlock)
replay_kernel(const __grid_constant__ KProblem P, const double *init_dof, const cp_move *mv,
              unsigned long long n_moves, unsigned long long n, cp_decision *out) {
    extern __shared__ double smem[];
    const double *s_items, *s_shift;
    Geo g = stage_shared<KIND, L>(P, smem, s_items, s_shift);
    const int lane = threadIdx.x % L;
    const unsigned gmask = group_mask<L>();
    const unsigned long long rep = (unsigned long long)blockIdx.x * (kBlock / L) + threadIdx.x / L;
    if (rep >= n) return;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = init_dof[rep * CP_MAX_DOF + i];
    const int n_cell = n_cell_dof(P.family);
    const double *bd = P.replay_has_bounds ? init_dof + n * CP_MAX_DOF : init_dof;
    const double len0 = bd[rep * CP_MAX_DOF + 0], ratio0 = bd[rep * CP_MAX_DOF + 1];
    double score_current = state_score<KIND, L>(P, s_items, s_shift, g, d, trig_of(d), gmask, lane);
    for (unsigned long long m = 0; m < n_moves; ++m) {
        const cp_move M = mv[rep * n_moves + m];
        const int slot = basis_to_slot((int)M.basis_index, n_cell);
        const double val = d.get(slot);
        double lo, hi;
        slot_bounds(slot, len0, ratio0, lo, hi);
        cp_decision D;
        D.in_bounds = 0;
        D.accepted = 0;
        D.new_score = __longlong_as_double(0x7FF8000000000000ll);
        if (lo <= M.new_value && M.new_value <= hi) {
            D.in_bounds = 1;
            d.set(slot, M.new_value);
            D.new_score = state_score<KIND, L>(P, s_items, s_shift, g, d, trig_of(d), gmask, lane);
            if (accept_move(D.new_score, score_current, M.kt, M.threshold)) {
                score_current = D.new_score;
                D.accepted = 1;
            } else {
                d.set(slot, val);
            }
        }
        D.score_current = score_current;
        if (lane == 0) out[rep * n_moves + m] = D;
    }
}


#endif  // CP_GROUP_CUH
