// cp_thread_lj.cuh — PotentialState::score for one replica in one thread (LJShape2): geometry
// layout and energy sum of the one-thread-per-replica LJ kernels.  Included by cp_thread.cuh.
#ifndef CP_THREAD_LJ_CUH
#define CP_THREAD_LJ_CUH

#include "cp_device.cuh"

// doubles of shared memory each LJ thread needs: four cell values, per image fx fy cx cy, per site
// the rotated position
__host__ __device__ inline int lj_geo_doubles(int n_items, int n_syms) { return 4 + n_syms * 4 + n_syms * n_items * 2; }

struct TGeo {
    double *base;  // element e of this thread is base[e * 32]
    int comp0;     // first component element (= 4 + N * 4); images start at element 4
    __device__ __forceinline__ double ld(int e) const { return base[e * 32]; }
    __device__ __forceinline__ void st(int e, double v) const { base[e * 32] = v; }
};

// per-replica values that depend on the cell DOFs only, kept across moves
struct CellCache {
    int shells;    // periodic_range (packed.rs:128-132)
    double inv_a;  // 1 / a, used only to bound the image columns worth testing (never in a decision)
};
__device__ __forceinline__ CellCache cell_cache_of(const Dof &d) {
    CellCache c;
    c.shells = periodic_range(d.v[0], d.v[0] * d.v[1], d.v[2]);
    c.inv_a = 1.0 / d.v[0];
    return c;
}

template <int NI>
__device__ __forceinline__ void lj_build_geometry(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
    const int N = P.n_syms, n = NI > 0 ? NI : P.n_items;
    constexpr int CW = 2;
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
        }
    }
}

// LJShape2::energy (shape/lj_shape.rs:35-39): iproduct!(self.items, other.items) summed in order,
// `other` = shape j displaced to (X, Y)
template <int NI>
__device__ __forceinline__ double t_shape_energy(const KProblem &P, TGeo g, int i, int j, double X, double Y) {
    const int n = NI > 0 ? NI : P.n_items;
    const int ei = g.comp0 + i * n * 2, ej = g.comp0 + j * n * 2;
    const double cx = g.ld(4 + i * 4 + 2), cy = g.ld(4 + i * 4 + 3);
    double e = 0.;
    // the run-time-count build (NI = 0) keeps these loops rolled: the function is inlined at nine
    // call sites of t_lj_score and 16 x 16 unrolled pair evaluations each would not compile in
    // reasonable time
#pragma unroll(NI > 0 ? NI : 1)
    for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
        if (NI == 0 && k >= n) break;
        const double sx = g.ld(ei + k * 2 + 0) + cx, sy = g.ld(ei + k * 2 + 1) + cy;
        const double *it = P.items[k];
#pragma unroll(NI > 0 ? NI : 1)
        for (int l = 0; l < (NI > 0 ? NI : CP_MAX_ITEMS); ++l) {
            if (NI == 0 && l >= n) break;
            const double ox = g.ld(ej + l * 2 + 0) + X, oy = g.ld(ej + l * 2 + 1) + Y;
            e += lj_pair(sx - ox, sy - oy, it[2], it[3], it[4], P.lj_shift[k]);
        }
    }
    return e;
}

// PotentialState::score (state/potential.rs:82-115) for one replica in one thread: the reference's
// accumulation order is simply the loop order here.  Exact pruning: when every site has a cutoff,
// a shape pair whose centres are further apart than max cutoff + 2 max|site| (P.lj_reach, widened
// by 1e-9) has all its site distances beyond the cutoff, each LJ2::energy is exactly 0.0 and adding
// it cannot change the sum - so the pair, its row or its columns are skipped outright.
template <int NI>
__device__ __forceinline__ double t_lj_score(const KProblem &P, TGeo g, const Dof &d, const Trig &t,
                                             const CellCache &cc) {
    lj_build_geometry<NI>(P, g, d, t);
    const int N = P.n_syms;
    const double a = d.v[0], b = d.v[0] * d.v[1];
    const double reach = P.lj_reach * (1.0 + 1e-9);  // +inf: no pruning
    const double reach_sq = reach * reach;
    double sum = 0.;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            sum += t_shape_energy<NI>(P, g, i, j, g.ld(4 + j * 4 + 2), g.ld(4 + j * 4 + 3));
    for (int i = 0; i < N; ++i) {
        const double cx = g.ld(4 + i * 4 + 2), cy = g.ld(4 + i * 4 + 3);
        for (int j = 0; j < N; ++j) {
            const double fx0 = g.ld(4 + j * 4 + 0), fy0 = g.ld(4 + j * 4 + 1);
            if (!(reach < 1e300)) {  // a site without cutoff: every image contributes, nothing to prune
                for (int ix = -3; ix <= 3; ++ix) {  // periodic_images(position, 3, false): x outer, y inner
                    const double fx = fx0 + (double)ix;
                    const double xa = fx * a;
                    for (int iy = -3; iy <= 3; ++iy) {
                        if ((ix | iy) == 0) continue;
                        const double fy = fy0 + (double)iy;  // to_cartesian_translate (cell.rs:241-245)
                        const double yb = fy * b;
                        sum += t_shape_energy<NI>(P, g, i, j, xa + yb * t.ca, yb * t.sa);
                    }
                }
                continue;
            }
            // the seven image rows of this pair: Y and the skew term depend on iy only, and a row
            // whose |dy| alone exceeds the reach holds no pair within any cutoff
            double rowY[7], rowC[7];
            unsigned rows = 0;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const double fy = fy0 + (double)(r - 3);  // to_cartesian_translate (cell.rs:241-245)
                const double yb = fy * b;
                rowY[r] = yb * t.sa;
                rowC[r] = yb * t.ca;
                if (!(fabs(cy - rowY[r]) > reach)) rows |= 1u << r;
            }
            double dix = -3.0;
            for (int ix = -3; ix <= 3; ++ix, dix += 1.0) {  // periodic_images(position, 3, false): x outer, y inner
                const double fx = fx0 + dix;
                const double xa = fx * a;
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    if (!((rows >> r) & 1u) || (ix == 0 && r == 3)) continue;
                    const double X = xa + rowC[r], Y = rowY[r];
                    const double ddx = cx - X, ddy = cy - Y;
                    if (ddx * ddx + ddy * ddy > reach_sq) continue;  // every site pair beyond its cutoff
                    sum += t_shape_energy<NI>(P, g, i, j, X, Y);
                }
            }
        }
    }
    return -sum / (double)N;
}

#endif  // CP_THREAD_LJ_CUH
