// cp_thread.cuh — one THREAD per replica (lanes_per_replica = 1).
//
// The group kernels of cp_api.cu spread one replica over L lanes; profiling them on B200
// (profiles/r1_base_*.txt) showed ~75 % of the issued instructions were replicated scalar work
// (RNG, trigonometry, index arithmetic) and votes.  Here every lane owns a whole replica, so all of
// that is useful work, and the sweeps of one replica run as straight-line loops with the
// instruction-level parallelism of 25+ independent predicate evaluations.  What SIMT needs in
// return is warp-uniform control flow, which the move loop below provides:
//   * proposals are drawn in a short divergent loop until each lane holds an in-bounds one (the
//     37 % of out-of-bounds proposals cost a draw, not a score slot), then all lanes score together;
//   * periodic-image candidates that pass the distance filter are queued per lane and drained
//     afterwards, so the expensive component tests run once per queued candidate instead of once
//     per candidate index that any lane of the warp happens to pass.
// Per-thread geometry lives in shared memory as [element][thread] (bank-conflict free).
//
// Arithmetic is the same as cp_device.cuh: reference expression order, f64, no contraction.
#ifndef CP_THREAD_CUH
#define CP_THREAD_CUH

#include "cp_device.cuh"

constexpr int kThreadBlock = 64;

// doubles of shared memory each thread needs
__host__ __device__ inline int thread_geo_doubles(int kind, int n_items, int n_syms) {
    return n_syms * 8 + n_syms * n_items * (kind == CP_POLYGON ? 8 : 4);
}

struct TGeo {
    double *base;  // element e of this thread is base[e * kThreadBlock]
    int comp0;     // first component element (= N * 8)
    __device__ __forceinline__ double ld(int e) const { return base[e * kThreadBlock]; }
    __device__ __forceinline__ void st(int e, double v) const { base[e * kThreadBlock] = v; }
};

template <int KIND, int NI>
__device__ __forceinline__ void t_build_geometry(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
    const int N = P.n_syms, n = NI > 0 ? NI : P.n_items;
    constexpr int CW = KIND == CP_POLYGON ? 8 : 4;
    const double a = d.v[0], b = d.v[0] * d.v[1];
    for (int s = 0; s < N; ++s) {
        const double *S = P.syms[s];
        // sym * Transform2::new(theta, (x, y)), rows 0-1 (site.rs:40-45)
        double m00 = ((S[0] * t.ct) + S[1] * t.st) + S[2] * 0.0;
        double m01 = ((S[0] * -t.st) + S[1] * t.ct) + S[2] * 0.0;
        double tx = ((S[0] * d.v[3]) + S[1] * d.v[4]) + S[2] * 1.0;
        double m10 = ((S[3] * t.ct) + S[4] * t.st) + S[5] * 0.0;
        double m11 = ((S[3] * -t.st) + S[4] * t.ct) + S[5] * 0.0;
        double ty = ((S[3] * d.v[3]) + S[4] * d.v[4]) + S[5] * 1.0;
        double fx = wrap_unit(tx), fy = wrap_unit(ty);
        double yb = fy * b;
        double cx = fx * a + yb * t.ca, cy = yb * t.sa;  // cell.rs:258-263
        g.st(s * 8 + 4, fx);
        g.st(s * 8 + 5, fy);
        g.st(s * 8 + 6, cx);
        g.st(s * 8 + 7, cy);
#pragma unroll
        for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
            if (NI == 0 && k >= n) break;
            const double *it = P.items[k];
            const int e = g.comp0 + (s * n + k) * CW;
            double rx = m00 * it[0] + m01 * it[1];
            double ry = m10 * it[0] + m11 * it[1];
            if (KIND == CP_POLYGON) {
                double ex = m00 * it[2] + m01 * it[3];
                double ey = m10 * it[2] + m11 * it[3];
                double sx = rx + cx, sy = ry + cy;
                g.st(e + 0, rx); g.st(e + 1, ry); g.st(e + 2, ex); g.st(e + 3, ey);
                g.st(e + 4, sx); g.st(e + 5, sy);
                g.st(e + 6, (ex + cx) - sx);
                g.st(e + 7, (ey + cy) - sy);
            } else {
                g.st(e + 0, rx); g.st(e + 1, ry);
                g.st(e + 2, rx + cx);
                g.st(e + 3, ry + cy);
            }
        }
    }
}

// shape i (placed) against shape j displaced to (X, Y) (rotated points + image position), or, when
// IN_CELL, against shape j's own placement
template <int KIND, int NI, bool IN_CELL>
__device__ __forceinline__ bool t_shapes_overlap(const KProblem &P, TGeo g, int i, int j, double X, double Y) {
    const int n = NI > 0 ? NI : P.n_items;
    constexpr int CW = KIND == CP_POLYGON ? 8 : 4;
    bool hit = false;
    const int ei = g.comp0 + i * n * CW, ej = g.comp0 + j * n * CW;
#pragma unroll
    for (int l = 0; l < (NI > 0 ? NI : CP_MAX_ITEMS); ++l) {
        if (NI == 0 && l >= n) break;
        const int eb = ej + l * CW;
        double ox, oy, odx = 0., ody = 0., orad = 0.;
        if (KIND == CP_POLYGON) {
            if (IN_CELL) {
                ox = g.ld(eb + 4); oy = g.ld(eb + 5); odx = g.ld(eb + 6); ody = g.ld(eb + 7);
            } else {
                ox = g.ld(eb + 0) + X; oy = g.ld(eb + 1) + Y;
                double ex = g.ld(eb + 2) + X, ey = g.ld(eb + 3) + Y;
                odx = ex - ox; ody = ey - oy;
            }
        } else {
            if (IN_CELL) { ox = g.ld(eb + 2); oy = g.ld(eb + 3); }
            else { ox = g.ld(eb + 0) + X; oy = g.ld(eb + 1) + Y; }
            orad = P.items[l][2];
        }
#pragma unroll
        for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
            if (NI == 0 && k >= n) break;
            const int ea = ei + k * CW;
            if (KIND == CP_POLYGON) {
                hit |= segments_cross(g.ld(ea + 4), g.ld(ea + 5), g.ld(ea + 6), g.ld(ea + 7), ox, oy, odx, ody);
            } else {
                double rs = P.items[k][2] + orad;  // Atom2::intersects (atom2.rs:21-26)
                double dx = g.ld(ea + 2) - ox, dy = g.ld(ea + 3) - oy;
                hit |= (dx * dx + dy * dy) < rs * rs;
            }
        }
    }
    return hit;
}

// PackedState::check_intersection (state/packed.rs:127-167) for one replica in one thread
template <int KIND, int NI>
__device__ __forceinline__ bool t_hard_overlap(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
    const int N = P.n_syms;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            if (t_shapes_overlap<KIND, NI, true>(P, g, i, j, 0., 0.)) return true;

    const double a = d.v[0], b = d.v[0] * d.v[1];
    const int shells = periodic_range(a, b, d.v[2]);
    const int side = 2 * shells + 1, n_img = side * side, centre = n_img / 2;
    const double r2x = P.radius * 2.;
    const double radius_sq = r2x * r2x;
    // queue of candidates that passed the distance filter: 6 x 10-bit codes ((i*N + j) * n_img + q)
    unsigned long long queue = 0;
    int queued = 0;
    bool hit = false;
    auto drain = [&]() {
        while (queued > 0 && !hit) {
            --queued;
            const int code = (int)(queue & 1023ull);
            queue >>= 10;
            const int ij = code / n_img, q = code - ij * n_img;
            const int i = ij / N, j = ij - i * N;
            const int qx = q / side;
            const double fx = g.ld(j * 8 + 4) + (double)(qx - shells);
            const double fy = g.ld(j * 8 + 5) + (double)(q - qx * side - shells);
            const double yb = fy * b;
            hit = t_shapes_overlap<KIND, NI, false>(P, g, i, j, fx * a + yb * t.ca, yb * t.sa);
        }
    };
    for (int i = 0; i < N && !hit; ++i) {
        const double cx = g.ld(i * 8 + 6), cy = g.ld(i * 8 + 7);
        for (int j = 0; j < N && !hit; ++j) {
            const double fx0 = g.ld(j * 8 + 4), fy0 = g.ld(j * 8 + 5);
            const int code0 = (i * N + j) * n_img;
            int q = 0;
            for (int ix = -shells; ix <= shells; ++ix) {
                const double fx = fx0 + (double)ix;  // to_cartesian_translate (cell.rs:241-245)
                const double xa = fx * a;
                for (int iy = -shells; iy <= shells; ++iy, ++q) {
                    const double fy = fy0 + (double)iy;
                    const double yb = fy * b;
                    const double ddx = cx - (xa + yb * t.ca), ddy = cy - yb * t.sa;
                    const bool pass = (ddx * ddx + ddy * ddy) <= radius_sq && q != centre;  // packed.rs:156-157
                    if (pass) {
                        if (queued == 6) drain();  // rare: more than six neighbours in range
                        queue |= (unsigned long long)(code0 + q) << (10 * queued);
                        ++queued;
                    }
                }
            }
        }
    }
    drain();
    return hit;
}

template <int KIND, int NI>
__device__ __forceinline__ double t_state_score(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
    t_build_geometry<KIND, NI>(P, g, d, t);
    const bool overlap = t_hard_overlap<KIND, NI>(P, g, d, t);
    const double cell_area = t.sa * d.v[0] * (d.v[0] * d.v[1]);
    return overlap ? __longlong_as_double(0x7FF8000000000000ll) : (P.area * (double)P.n_syms) / cell_area;
}

// Monte-Carlo optimisation, one replica per thread (see the header comment for the control flow).
// Semantics are those of mc_kernel in cp_api.cu: MCOptimiser::optimise_state (optimisation.rs:80-180)
// for every stage, seeds and bounds per stage as analyse_state / generate_basis set them.
template <int KIND, int NI, int RNG>
__global__ void __launch_bounds__(kThreadBlock) mc_thread_kernel(const __grid_constant__ KParams K) {
    extern __shared__ double smem[];
    const KProblem &P = K.prob;
    const TGeo g{smem + threadIdx.x, P.n_syms * 8};
    const unsigned long long rep = (unsigned long long)blockIdx.x * kThreadBlock + threadIdx.x;
    const bool active = rep < K.n_replicas;  // tail lanes stay in the warp-wide votes, doing nothing
    const unsigned long long replica = K.replica_begin + (active ? rep : 0ull);

    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i)
        d.v[i] = (K.init_dof && active) ? K.init_dof[rep * CP_MAX_DOF + i] : P.dof[i];
    Trig t;
    cp_sincos(d.v[5], &t.st, &t.ct);
    cp_sincos(d.v[2], &t.sa, &t.ca);

    const int n_cell = n_cell_dof(P.family);
    const int n_basis = n_cell + 3;
    unsigned long long rejections = 0, moves = 0;
    double score_current = t_state_score<KIND, NI>(P, g, d, t);
    const bool valid = score_current == score_current;
    if (active && !valid) atomicExch(K.error_flag, 1);
    const bool alive = active && valid;

    Rng rng;
    for (int stage = 0; stage < K.n_stages; ++stage) {
        const KStage S = K.stages[stage];
        const double len0 = d.v[0], ratio0 = d.v[1];
        rng_begin_stage<RNG>(rng, K.seed_base + replica, replica, n_basis);
        double kt = S.kt_start, step_ratio = 1.0;
        int convergence_count = 0;
        bool converged = false;
        const unsigned long long outer = S.inner ? S.steps / S.inner : 0ull;
        for (unsigned long long loop = 1; loop <= outer; ++loop) {
            const bool live = alive && !converged;
            const double score_start = score_current;
            unsigned long long loop_rejections = 0, it = 0;
            for (;;) {
                bool have = false;
                int slot = 0;
                double val = 0., new_val = 0.;
                while (live && it < S.inner && !have) {  // draw until a proposal is in bounds
                    ++it;
                    int basis_index;
                    double sample;
                    rng_draw_move<RNG>(rng, basis_index, sample);
                    slot = basis_to_slot(basis_index, n_cell);
                    val = d.get(slot);
                    new_val = val + S.max_step * step_ratio * sample;
                    double lo, hi;
                    slot_bounds(slot, len0, ratio0, lo, hi);
                    have = lo <= new_val && new_val <= hi;
                    if (!have) ++loop_rejections;
                }
                if (!__any_sync(0xffffffffu, have)) break;
                if (have) {
                    d.set(slot, new_val);
                    double old_s = 0., old_c = 0.;
                    if (slot == 5 || slot == 2) {  // only these two DOFs enter sin/cos
                        double s_, c_;
                        cp_sincos(new_val, &s_, &c_);
                        if (slot == 5) { old_s = t.st; old_c = t.ct; t.st = s_; t.ct = c_; }
                        else { old_s = t.sa; old_c = t.ca; t.sa = s_; t.ca = c_; }
                    }
                    const double new_score = t_state_score<KIND, NI>(P, g, d, t);
                    const double threshold = rng_draw_threshold<RNG>(rng);
                    if (accept_move(new_score, score_current, kt, threshold)) {
                        score_current = new_score;
                    } else {
                        d.set(slot, val);
                        if (slot == 5) { t.st = old_s; t.ct = old_c; }
                        if (slot == 2) { t.sa = old_s; t.ca = old_c; }
                        ++loop_rejections;
                    }
                }
            }
            if (live) {
                moves += S.inner;
                rejections += loop_rejections;
                kt *= S.kt_ratio;
                if (S.convergence == S.convergence) {
                    if (score_current - score_start < S.convergence) {
                        if (++convergence_count > 5) converged = true;
                    } else {
                        convergence_count = 0;
                    }
                }
                if (!converged && step_ratio > 1e-4)
                    step_ratio *= (double)S.inner / ((double)loop_rejections + 1.);
            }
        }
    }
    if (active) {
        cp_replica_result r;
#pragma unroll
        for (int i = 0; i < CP_MAX_DOF; ++i) r.dof[i] = d.v[i];
        r.score = score_current;
        r.rejections = rejections;
        r.moves = moves;
        K.out[rep] = r;
    }
}

template <int KIND, int NI>
__global__ void __launch_bounds__(kThreadBlock)
score_thread_kernel(const __grid_constant__ KProblem P, const double *dof, unsigned long long n, double *out) {
    extern __shared__ double smem[];
    const TGeo g{smem + threadIdx.x, P.n_syms * 8};
    const unsigned long long rep = (unsigned long long)blockIdx.x * kThreadBlock + threadIdx.x;
    if (rep >= n) return;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = dof[rep * CP_MAX_DOF + i];
    Trig t;
    cp_sincos(d.v[5], &t.st, &t.ct);
    cp_sincos(d.v[2], &t.sa, &t.ca);
    out[rep] = t_state_score<KIND, NI>(P, g, d, t);
}

template <int KIND, int NI>
__global__ void __launch_bounds__(kThreadBlock)
replay_thread_kernel(const __grid_constant__ KProblem P, const double *init_dof, const cp_move *mv,
                     unsigned long long n_moves, unsigned long long n, cp_decision *out) {
    extern __shared__ double smem[];
    const TGeo g{smem + threadIdx.x, P.n_syms * 8};
    const unsigned long long rep = (unsigned long long)blockIdx.x * kThreadBlock + threadIdx.x;
    if (rep >= n) return;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = init_dof[rep * CP_MAX_DOF + i];
    Trig t;
    const int n_cell = n_cell_dof(P.family);
    const double len0 = d.v[0], ratio0 = d.v[1];
    cp_sincos(d.v[5], &t.st, &t.ct);
    cp_sincos(d.v[2], &t.sa, &t.ca);
    double score_current = t_state_score<KIND, NI>(P, g, d, t);
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
            cp_sincos(d.v[5], &t.st, &t.ct);
            cp_sincos(d.v[2], &t.sa, &t.ca);
            D.new_score = t_state_score<KIND, NI>(P, g, d, t);
            if (accept_move(D.new_score, score_current, M.kt, M.threshold)) {
                score_current = D.new_score;
                D.accepted = 1;
            } else {
                d.set(slot, val);
                cp_sincos(d.v[5], &t.st, &t.ct);
                cp_sincos(d.v[2], &t.sa, &t.ca);
            }
        }
        D.score_current = score_current;
        out[rep * n_moves + m] = D;
    }
}

#endif  // CP_THREAD_CUH
