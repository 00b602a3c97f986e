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
    // per image: fx fy cx cy; per component: the rotated points (polygon: start and end; disc: centre).
    // Placed coordinates (rotated + position) are re-derived where they are used: two additions
    // are cheaper than the occupancy a wider per-thread footprint costs.
    // plus four cell values (a, b, cos, sin of the cell angle) so that ANY lane of the warp can
    // evaluate a queued candidate of this replica (cooperative drain).
    return 4 + n_syms * 4 + n_syms * n_items * (kind == CP_POLYGON ? 4 : 2);
}

struct TGeo {
    double *base;  // element e of this thread is base[e * kThreadBlock]
    int comp0;     // first component element (= 4 + N * 4); images start at element 4
    __device__ __forceinline__ double ld(int e) const { return base[e * kThreadBlock]; }
    __device__ __forceinline__ void st(int e, double v) const { base[e * kThreadBlock] = v; }
};

template <int KIND, int NI>
__device__ __forceinline__ void t_build_geometry(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
    const int N = P.n_syms, n = NI > 0 ? NI : P.n_items;
    constexpr int CW = KIND == CP_POLYGON ? 4 : 2;
    const double a = d.v[0], b = d.v[0] * d.v[1];
    g.st(0, a);
    g.st(1, b);
    g.st(2, t.ca);
    g.st(3, t.sa);
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
        g.st(4 + s * 4 + 0, fx);
        g.st(4 + s * 4 + 1, fy);
        g.st(4 + s * 4 + 2, fx * a + yb * t.ca);  // cell.rs:258-263
        g.st(4 + s * 4 + 3, yb * t.sa);
#pragma unroll
        for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
            if (NI == 0 && k >= n) break;
            const double *it = P.items[k];
            const int e = g.comp0 + (s * n + k) * CW;
            g.st(e + 0, m00 * it[0] + m01 * it[1]);  // Transform * Point without its translation
            g.st(e + 1, m10 * it[0] + m11 * it[1]);
            if (KIND == CP_POLYGON) {
                g.st(e + 2, m00 * it[2] + m01 * it[3]);
                g.st(e + 3, m10 * it[2] + m11 * it[3]);
            }
        }
    }
}

// shape i in its cell placement against shape j displaced to (X, Y): the image position, or j's own
// Cartesian position for the in-cell pairs.  Both are shape.transform(&t) = rotated point + t's
// translation (line_shape.rs:103-108), so one routine serves both.
template <int KIND, int NI>
__device__ __forceinline__ bool t_shapes_overlap(const KProblem &P, TGeo g, int i, int j, double X, double Y) {
    const int n = NI > 0 ? NI : P.n_items;
    constexpr int CW = KIND == CP_POLYGON ? 4 : 2;
    bool hit = false;
    const int ei = g.comp0 + i * n * CW, ej = g.comp0 + j * n * CW;
    const double cx = g.ld(4 + i * 4 + 2), cy = g.ld(4 + i * 4 + 3);
#pragma unroll
    for (int l = 0; l < (NI > 0 ? NI : CP_MAX_ITEMS); ++l) {
        if (NI == 0 && l >= n) break;
        const int eb = ej + l * CW;
        const double ox = g.ld(eb + 0) + X, oy = g.ld(eb + 1) + Y;
        double odx = 0., ody = 0., orad = 0.;
        if (KIND == CP_POLYGON) {
            odx = (g.ld(eb + 2) + X) - ox;  // Line2::dx / dy of the placed segment (line2.rs:94-101)
            ody = (g.ld(eb + 3) + Y) - oy;
        } else {
            orad = P.items[l][2];
        }
#pragma unroll
        for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
            if (NI == 0 && k >= n) break;
            const int ea = ei + k * CW;
            const double sx = g.ld(ea + 0) + cx, sy = g.ld(ea + 1) + cy;
            if (KIND == CP_POLYGON) {
                const double sdx = (g.ld(ea + 2) + cx) - sx, sdy = (g.ld(ea + 3) + cy) - sy;
                hit |= segments_cross(sx, sy, sdx, sdy, ox, oy, odx, ody);
            } else {
                const double rs = P.items[k][2] + orad;  // Atom2::intersects (atom2.rs:21-26)
                const double dx = sx - ox, dy = sy - oy;
                hit |= (dx * dx + dy * dy) < rs * rs;
            }
        }
    }
    return hit;
}

// A queued periodic-image candidate: shape i against image (ix, iy) of shape j, 10 bits
__device__ __forceinline__ unsigned cand_code(int i, int j, int ix, int iy) {
    return (unsigned)((i << 8) | (j << 6) | ((ix + 3) << 3) | (iy + 3));
}

// shape.transform(&transform2) + intersects for one queued candidate of the replica whose geometry
// is `g` (packed.rs:158-161).  Reads only shared memory, so any lane can evaluate it.
template <int KIND, int NI>
__device__ __forceinline__ bool t_test_candidate(const KProblem &P, TGeo g, unsigned code) {
    const int i = (code >> 8) & 3, j = (code >> 6) & 3;
    const int ix = (int)((code >> 3) & 7) - 3, iy = (int)(code & 7) - 3;
    const double a = g.ld(0), b = g.ld(1), ca = g.ld(2), sa = g.ld(3);
    const double fx = g.ld(4 + j * 4 + 0) + (double)ix;  // to_cartesian_translate (cell.rs:241-245)
    const double fy = g.ld(4 + j * 4 + 1) + (double)iy;
    const double yb = fy * b;
    return t_shapes_overlap<KIND, NI>(P, g, i, j, fx * a + yb * ca, yb * sa);
}

struct CandQueue {
    unsigned long long codes = 0;  // up to six 10-bit codes
    int count = 0;
};

// PackedState::check_intersection (state/packed.rs:127-167), first part: the in-cell pairs and the
// distance filter over the periodic images.  Returns true when an overlap is already certain;
// otherwise the candidates inside 2R are left in `q` for the caller to test (t_drain_local, or the
// warp-cooperative drain of the MC kernel).
template <int KIND, int NI>
__device__ __forceinline__ bool t_sweep(const KProblem &P, TGeo g, const Dof &d, const Trig &t, int shells,
                                        CandQueue &q) {
    const int N = P.n_syms;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            if (t_shapes_overlap<KIND, NI>(P, g, i, j, g.ld(4 + j * 4 + 2), g.ld(4 + j * 4 + 3))) return true;

    const double a = d.v[0], b = d.v[0] * d.v[1];
    const double r2x = P.radius * 2.;
    const double radius_sq = r2x * r2x;
    // A row of images (fixed iy) whose |dy| alone exceeds 2R cannot pass the filter: dy*dy is then
    // above radius_sq by a relative 2e-12, far more than the rounding of the reference's
    // dx*dx + dy*dy, so skipping the row leaves every reference decision unchanged.
    const double row_reach = r2x * (1.0 + 1e-12);
    bool hit = false;
    for (int i = 0; i < N && !hit; ++i) {
        const double cx = g.ld(4 + i * 4 + 2), cy = g.ld(4 + i * 4 + 3);
        for (int j = 0; j < N && !hit; ++j) {
            const double fx0 = g.ld(4 + j * 4 + 0), fy0 = g.ld(4 + j * 4 + 1);
            for (int iy = -shells; iy <= shells; ++iy) {
                const double fy = fy0 + (double)iy;  // to_cartesian_translate (cell.rs:241-245)
                const double yb = fy * b;
                const double ddy = cy - yb * t.sa;
                if (fabs(ddy) > row_reach) continue;
                const double ybca = yb * t.ca, ddy2 = ddy * ddy;
                for (int ix = -shells; ix <= shells; ++ix) {
                    const double fx = fx0 + (double)ix;
                    const double ddx = cx - (fx * a + ybca);
                    const bool pass = (ddx * ddx + ddy2) <= radius_sq && (ix | iy) != 0;  // packed.rs:156-157
                    if (pass) {
                        if (q.count == 6) {  // rare: more than six neighbours in range; test them now
                            while (q.count > 0 && !hit) {
                                --q.count;
                                hit = t_test_candidate<KIND, NI>(P, g, (unsigned)(q.codes & 1023ull));
                                q.codes >>= 10;
                            }
                            q.codes = 0;
                            q.count = 0;
                        }
                        q.codes |= (unsigned long long)cand_code(i, j, ix, iy) << (10 * q.count);
                        ++q.count;
                    }
                }
            }
        }
    }
    return hit;
}

template <int KIND, int NI>
__device__ __forceinline__ bool t_drain_local(const KProblem &P, TGeo g, CandQueue &q) {
    bool hit = false;
    while (q.count > 0 && !hit) {
        --q.count;
        hit = t_test_candidate<KIND, NI>(P, g, (unsigned)(q.codes & 1023ull));
        q.codes >>= 10;
    }
    return hit;
}

// Warp-cooperative drain: the queued candidates of all 32 replicas of the warp are pooled and dealt
// out evenly, one per lane per round, because their number varies a lot between replicas (0..6+)
// and a lane-local drain would make every lane wait for the longest queue.  Must be called by all
// 32 lanes converged.  wq: 192 words + 1 flag word of shared memory per warp.
template <int KIND, int NI>
__device__ __forceinline__ bool t_drain_warp(const KProblem &P, double *smem_geo, unsigned *wq, const CandQueue &q) {
    const unsigned lane = threadIdx.x & 31u, warp_first = threadIdx.x & ~31u;
    int incl = q.count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return false;
    const int base = incl - q.count;
    for (int k = 0; k < q.count; ++k) wq[base + k] = (lane << 10) | (unsigned)((q.codes >> (10 * k)) & 1023ull);
    if (lane == 0) wq[192] = 0u;
    __syncwarp();
    for (int e = (int)lane; e < total; e += 32) {
        const unsigned entry = wq[e];
        const unsigned owner = entry >> 10;
        const TGeo og{smem_geo + warp_first + owner, 0};
        TGeo gg = og;
        gg.comp0 = 4 + P.n_syms * 4;
        if (t_test_candidate<KIND, NI>(P, gg, entry & 1023u)) atomicOr(&wq[192], 1u << owner);
    }
    __syncwarp();
    const bool hit = (wq[192] >> lane) & 1u;
    __syncwarp();  // wq is rewritten by the next move
    return hit;
}

// State::score for one replica in one thread, everything local (batch score, replay, initial score)
template <int KIND, int NI>
__device__ __forceinline__ double t_state_score(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
    t_build_geometry<KIND, NI>(P, g, d, t);
    CandQueue q;
    const int shells = periodic_range(d.v[0], d.v[0] * d.v[1], d.v[2]);
    bool overlap = t_sweep<KIND, NI>(P, g, d, t, shells, q);
    if (!overlap) overlap = t_drain_local<KIND, NI>(P, g, q);
    const double cell_area = t.sa * d.v[0] * (d.v[0] * d.v[1]);
    return overlap ? __longlong_as_double(0x7FF8000000000000ll) : (P.area * (double)P.n_syms) / cell_area;
}

// Monte-Carlo optimisation, one replica per thread (see the header comment for the control flow).
// Semantics are those of mc_kernel in cp_api.cu: MCOptimiser::optimise_state (optimisation.rs:80-180)
// for every stage, seeds and bounds per stage as analyse_state / generate_basis set them.
template <int KIND, int NI, int RNG>
__global__ void __launch_bounds__(kThreadBlock, 7) mc_thread_kernel(const __grid_constant__ KParams K) {
    extern __shared__ double smem[];
    const KProblem &P = K.prob;
    const TGeo g{smem + threadIdx.x, 4 + P.n_syms * 4};
    // after the per-thread geometry: one candidate pool (192 words + flag) per warp
    unsigned *wq = reinterpret_cast<unsigned *>(smem + kThreadBlock * thread_geo_doubles(KIND, P.n_items, P.n_syms)) +
                   (threadIdx.x >> 5) * 200;
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
    int shells = periodic_range(d.v[0], d.v[0] * d.v[1], d.v[2]);  // depends on the cell DOFs only
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
                // phase A (lanes holding a proposal): geometry, in-cell pairs, distance filter
                CandQueue q;
                bool overlap = false;
                double old_s = 0., old_c = 0.;
                int old_shells = shells;
                if (have) {
                    d.set(slot, new_val);
                    if (slot == 5 || slot == 2) {  // only these two DOFs enter sin/cos
                        double s_, c_;
                        cp_sincos(new_val, &s_, &c_);
                        if (slot == 5) { old_s = t.st; old_c = t.ct; t.st = s_; t.ct = c_; }
                        else { old_s = t.sa; old_c = t.ca; t.sa = s_; t.ca = c_; }
                    }
                    if (slot <= 2) shells = periodic_range(d.v[0], d.v[0] * d.v[1], d.v[2]);
                    t_build_geometry<KIND, NI>(P, g, d, t);
                    overlap = t_sweep<KIND, NI>(P, g, d, t, shells, q);
                    if (overlap) q.count = 0;
                }
                // phase B (whole warp): the queued image candidates, pooled and shared out
                overlap |= t_drain_warp<KIND, NI>(P, smem, wq, q);
                // phase C: score and Metropolis decision
                if (have) {
                    const double cell_area = t.sa * d.v[0] * (d.v[0] * d.v[1]);
                    const double new_score = overlap ? __longlong_as_double(0x7FF8000000000000ll)
                                                     : (P.area * (double)P.n_syms) / cell_area;
                    const double threshold = rng_draw_threshold<RNG>(rng);
                    if (accept_move(new_score, score_current, kt, threshold)) {
                        score_current = new_score;
                    } else {
                        d.set(slot, val);
                        if (slot == 5) { t.st = old_s; t.ct = old_c; }
                        if (slot == 2) { t.sa = old_s; t.ca = old_c; }
                        shells = old_shells;
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
    const TGeo g{smem + threadIdx.x, 4 + P.n_syms * 4};
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
    const TGeo g{smem + threadIdx.x, 4 + P.n_syms * 4};
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
