// cp_thread.cuh — one THREAD per replica (lanes_per_replica = 1), the default kernels.
//
// The group kernels (cp_group.cuh) spread one replica over L lanes; profiling them on B200
// (profiles/r1_a_group32_first.txt) showed ~75 % of the issued instructions were replicated scalar
// work (RNG, trigonometry, index arithmetic) and votes.  Here every lane owns a whole replica, so
// all of that is useful work, and the sweeps of one replica run as straight-line loops with the
// instruction-level parallelism of 25+ independent predicate evaluations.  What SIMT needs in
// return is warp-uniform control flow around the expensive parts, which the move loop provides:
//   * every lane walks its own (stage, outer loop, inner step) schedule and draws proposals in a
//     short divergent loop until it holds one that needs scoring: out-of-bounds proposals cost a
//     draw, and for hard shapes a cell move that the Metropolis rule rejects even if valid is
//     rejected there too, before any overlap test (the packing fraction depends on the cell only);
//   * phase A, per lane: geometry into shared memory ([element][thread], conflict-free) and the
//     exactly pruned sweep over the periodic images; candidates become bits of one 64-bit word per
//     ordered shape pair;
//   * phase B, whole warp: the candidates of all 32 replicas are tested in rounds from a shared
//     pool of 32 entries, one entry per lane, replicas dropping out at their first hit;
//   * phase C, per lane: the decision.
// Arithmetic is the same as cp_device.cuh: reference expression order, f64, no contraction.
#ifndef CP_THREAD_CUH
#define CP_THREAD_CUH

#include "cp_device.cuh"

constexpr int kThreadBlock = 32;

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

// A shape-pair candidate of one replica: shape i against image (ix, iy) of shape j, |ix|, |iy| <= 3.
// Bit q = (ix + 3) * 7 + (iy + 3) of the 64-bit word of the ordered pair (i, j); the centre bit
// (ix = iy = 0) stands for the in-cell pair i < j, whose displacement to_cartesian(f_j + (0, 0)) is
// bit for bit j's own Cartesian position (x + 0.0 == x), so one routine tests both kinds.
constexpr int kCentreBit = 24;

// shape.transform(&transform2) + intersects for candidate bit q of pair word `ij` of the replica
// whose geometry is `g` (packed.rs:134-148,158-161).  Reads only shared memory, so any lane of the
// warp can evaluate it.
template <int KIND, int NI>
__device__ __forceinline__ bool t_test_candidate(const KProblem &P, TGeo g, int i, int j, int q) {
    const int qx = (q * 37) >> 8;  // q / 7 for q < 49
    const int ix = qx - 3, iy = q - qx * 7 - 3;
    const double a = g.ld(0), b = g.ld(1), ca = g.ld(2), sa = g.ld(3);
    const double fx = g.ld(4 + j * 4 + 0) + (double)ix;  // to_cartesian_translate (cell.rs:241-245)
    const double fy = g.ld(4 + j * 4 + 1) + (double)iy;
    const double yb = fy * b;
    return t_shapes_overlap<KIND, NI>(P, g, i, j, fx * a + yb * ca, yb * sa);
}

// Candidate masks of one thread: N*N image words in shared memory, [word][thread]; the in-cell
// pairs (which word's centre bit is meant) and the scan cursor stay in registers.
struct CandMasks {
    unsigned long long *base;
    unsigned incell = 0;  // bit w set: in-cell pair of word w still to test
    int cursor = 0;       // words below it are exhausted
    __device__ __forceinline__ unsigned long long ld(int w) const { return base[w * kThreadBlock]; }
    __device__ __forceinline__ void st(int w, unsigned long long v) const { base[w * kThreadBlock] = v; }
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

// PackedState::check_intersection (state/packed.rs:127-167), candidate generation: every in-cell
// pair i < j (no distance filter there, packed.rs:134-148) and every periodic image that passes
// the 2R distance filter (packed.rs:150-157) is recorded as a bit; the caller tests them
// (t_drain_local, or the warp-cooperative drain of the MC kernel).  Returns the number of bits set.
template <int KIND, int NI>
__device__ __forceinline__ int t_sweep(const KProblem &P, TGeo g, const Dof &d, const Trig &t, const CellCache &cc,
                                       CandMasks &m) {
    const int N = P.n_syms, shells = cc.shells;
    const double a = d.v[0], b = d.v[0] * d.v[1];
    const double r2x = P.radius * 2.;
    const double radius_sq = r2x * r2x;
    // Exact pruning of the (2*shells+1)^2 image sweep.  A candidate passes only if dx*dx + dy*dy <=
    // (2R)^2, so (a) a row (fixed iy) with |dy| > 2R(1 + 1e-12) cannot pass, and (b) within a row
    // only the columns with |dx| <= 2R can: ix in [(u - 2R)/a, (u + 2R)/a] with u the row's offset.
    // The column bounds are computed approximately and widened by 1e-6 columns (a distance of
    // 1e-6 * a, against rounding errors of ~1e-15), so every candidate left out fails the
    // reference's filter by a wide margin; the ones kept are evaluated with the reference's
    // expressions, bit for bit.
    const double row_reach = r2x * (1.0 + 1e-12);
    int count = 0;
    m.incell = 0;
    m.cursor = 0;
    const double first_row = (double)(-shells);
    for (int i = 0; i < N; ++i) {
        const double cx = g.ld(4 + i * 4 + 2), cy = g.ld(4 + i * 4 + 3);
        for (int j = 0; j < N; ++j) {
            const double fx0 = g.ld(4 + j * 4 + 0), fy0 = g.ld(4 + j * 4 + 1);
            unsigned long long bits = 0ull;
            if (j > i) {
                m.incell |= 1u << (i * N + j);
                ++count;
            }
            // the image offsets as doubles, stepped by 1.0 (exact for these small integers) rather
            // than converted from the loop counters: int -> double conversions are slow
            double diy = first_row;
            for (int iy = -shells; iy <= shells; ++iy, diy += 1.0) {
                const double fy = fy0 + diy;  // to_cartesian_translate (cell.rs:241-245)
                const double yb = fy * b;
                const double ddy = cy - yb * t.sa;
                if (fabs(ddy) > row_reach) continue;
                const double ybca = yb * t.ca, ddy2 = ddy * ddy;
                const double u = (cx - ybca) * cc.inv_a - fx0;  // column where dx would vanish
                const double w = r2x * cc.inv_a + 1e-6;
                int lo = __double2int_ru(u - w), hi = __double2int_rd(u + w);
                lo = lo < -shells ? -shells : lo;
                hi = hi > shells ? shells : hi;
                double dix = (double)lo;
                for (int ix = lo; ix <= hi; ++ix, dix += 1.0) {
                    const double fx = fx0 + dix;
                    const double ddx = cx - (fx * a + ybca);
                    const bool pass = (ddx * ddx + ddy2) <= radius_sq && (ix | iy) != 0;  // packed.rs:156-157
                    if (pass) bits |= 1ull << ((ix + 3) * 7 + (iy + 3));
                }
            }
            m.st(i * N + j, bits);
            count += __popcll(bits);
        }
    }
    return count;
}

// pops the next candidate of this thread: in-cell pairs first (they decide most moves), then the
// image candidates in pair / bit order.  Returns false when none is left.
__device__ __forceinline__ bool t_pop_candidate(CandMasks &m, int NN, int &ij, int &q) {
    if (m.incell) {
        ij = __ffs((int)m.incell) - 1;
        m.incell &= m.incell - 1;
        q = kCentreBit;
        return true;
    }
    while (m.cursor < NN) {
        const unsigned long long bits = m.ld(m.cursor);
        if (bits) {
            q = __ffsll((long long)bits) - 1;
            m.st(m.cursor, bits & (bits - 1));
            ij = m.cursor;
            return true;
        }
        ++m.cursor;
    }
    return false;
}

template <int KIND, int NI>
__device__ __forceinline__ bool t_drain_local(const KProblem &P, TGeo g, CandMasks &m) {
    const int N = P.n_syms, NN = N * N;
    int ij, q;
    while (t_pop_candidate(m, NN, ij, q)) {
        const int i = ij / N;
        if (t_test_candidate<KIND, NI>(P, g, i, ij - i * N, q)) return true;
    }
    return false;
}

// Warp-cooperative drain.  The number of candidates varies a lot between replicas (1 .. dozens in
// dense, elongated cells), and a lane-local loop would make every lane wait for the longest list
// while most replicas are decided by their first hit.  So the warp works in rounds: every replica
// that is still undecided contributes its next few candidates to a pool of at most 32, the lanes
// test one pool entry each (reading the owner's geometry from shared memory), owners with a hit
// drop out.  `remaining` = this lane's candidate count (0 for lanes without a proposal).  Must be
// called by all 32 lanes converged; pool = 33 words of shared memory per warp.
template <int KIND, int NI>
__device__ __forceinline__ bool t_drain_warp(const KProblem &P, double *smem_geo, unsigned *pool, CandMasks &m,
                                             int remaining) {
    const unsigned lane = threadIdx.x & 31u, warp_first = threadIdx.x & ~31u;
    const int N = P.n_syms, NN = N * N, comp0 = 4 + N * 4;
    bool hit = false;
    for (;;) {
        const unsigned wanting = __ballot_sync(0xffffffffu, remaining > 0);
        if (wanting == 0u) break;
        const int quota = 32 / __popc(wanting);  // >= 1; pool never exceeds 32 entries
        const int take = remaining < quota ? remaining : quota;
        int incl = take;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int slot = incl - take;
        for (int k = 0; k < take; ++k) {
            int ij, q;
            t_pop_candidate(m, NN, ij, q);
            const int i = N == 1 ? 0 : (N == 2 ? ij >> 1 : (N == 4 ? ij >> 2 : (ij * 11) >> 5));  // ij / N, N <= 4
            pool[slot++] = (lane << 10) | (unsigned)(i << 8) | (unsigned)((ij - i * N) << 6) | (unsigned)q;
        }
        remaining -= take;
        if (lane == 0) pool[32] = 0u;
        __syncwarp();
        if ((int)lane < total) {
            const unsigned entry = pool[lane];
            const unsigned owner = entry >> 10;
            const TGeo og{smem_geo + warp_first + owner, comp0};
            if (t_test_candidate<KIND, NI>(P, og, (int)((entry >> 8) & 3u), (int)((entry >> 6) & 3u),
                                           (int)(entry & 63u)))
                atomicOr(&pool[32], 1u << owner);
        }
        __syncwarp();
        if ((pool[32] >> lane) & 1u) {
            hit = true;
            remaining = 0;  // decided: the rest of this replica's candidates need no test
        }
        __syncwarp();  // pool is rewritten by the next round
    }
    return hit;
}

template <int NI>
__device__ __forceinline__ double t_lj_score(const KProblem &P, TGeo g, const Dof &d, const Trig &t,
                                             const CellCache &cc);

// State::score for one replica in one thread, everything local (batch score, replay, initial score)
template <int KIND, int NI>
__device__ __forceinline__ double t_state_score(const KProblem &P, TGeo g, CandMasks m, const Dof &d,
                                                const Trig &t) {
    if constexpr (KIND == CP_LJ) return t_lj_score<NI>(P, g, d, t, cell_cache_of(d));
    t_build_geometry<KIND, NI>(P, g, d, t);
    t_sweep<KIND, NI>(P, g, d, t, cell_cache_of(d), m);
    const bool overlap = t_drain_local<KIND, NI>(P, g, m);
    const double cell_area = t.sa * d.v[0] * (d.v[0] * d.v[1]);
    return overlap ? __longlong_as_double(0x7FF8000000000000ll) : (P.area * (double)P.n_syms) / cell_area;
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
    t_build_geometry<CP_LJ, NI>(P, g, d, t);
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

// shared memory of a block: per-thread geometry, per-thread candidate masks, one pool per warp
__host__ __device__ inline size_t thread_block_shared_bytes(int kind, int n_items, int n_syms) {
    return sizeof(double) * kThreadBlock * (size_t)(thread_geo_doubles(kind, n_items, n_syms) + n_syms * n_syms) +
           sizeof(unsigned) * 40 * (kThreadBlock / 32);
}
struct ThreadShared {
    TGeo g;
    CandMasks m;
    unsigned long long *masks_all;
    unsigned *pool;
};
__device__ __forceinline__ ThreadShared thread_shared(double *smem, int kind, int n_items, int n_syms) {
    ThreadShared s;
    const int geo = thread_geo_doubles(kind, n_items, n_syms);
    s.g = TGeo{smem + threadIdx.x, 4 + n_syms * 4};
    s.masks_all = reinterpret_cast<unsigned long long *>(smem + kThreadBlock * geo);
    s.m.base = s.masks_all + threadIdx.x;
    s.pool = reinterpret_cast<unsigned *>(smem + kThreadBlock * (geo + n_syms * n_syms)) + (threadIdx.x >> 5) * 40;
    return s;
}

// Monte-Carlo optimisation, one replica per thread (see the header comment for the control flow).
// Semantics are those of mc_kernel in cp_api.cu: MCOptimiser::optimise_state (optimisation.rs:80-180)
// for every stage, seeds and bounds per stage as analyse_state / generate_basis set them.
// Minimum resident blocks per SM, i.e. the register budget: 14 warps per SM hold all 2,048 warps of
// a 65,536-replica batch at once (<= 146 registers); many-sided polygons are limited by shared
// memory to fewer warps anyway and get the registers their unrolled n x n tests want.
// LJ batches are several waves anyway (configs[3]: 131,072 replicas per GPU) and the row-hoisted
// image loop wants registers.
constexpr int thread_min_blocks(int kind, int ni) {
    return kind == CP_LJ ? 10 : (ni == 0 ? 10 : (ni <= 6 ? 14 : (ni <= 9 ? 10 : 8)));
}

// WIDE: the build for problems whose shared-memory footprint caps the occupancy below 12 warps per
// SM anyway (small polygons in the multiplicity-4 groups): same code, register budget of 8 blocks.
template <int KIND, int NI, int RNG, bool WIDE = false>
__global__ void __launch_bounds__(kThreadBlock, WIDE ? 8 : thread_min_blocks(KIND, NI))
mc_thread_kernel(const __grid_constant__ KParams K) {
    extern __shared__ double smem[];
    const KProblem &P = K.prob;
    const ThreadShared sh = thread_shared(smem, KIND, P.n_items, P.n_syms);
    const TGeo g = sh.g;
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
    double score_current = t_state_score<KIND, NI>(P, g, sh.m, d, t);
    const bool valid = score_current == score_current;
    if (active && !valid) atomicExch(K.error_flag, 1);
    const bool alive = active && valid;

    Rng rng;
    CellCache cc = cell_cache_of(d);
    CandMasks masks = sh.m;
    // Every lane walks its own (stage, outer loop, inner step) schedule: replicas of one warp need
    // different numbers of scored proposals per inner loop (the in-bounds fraction depends on the
    // replica's cell), and synchronising at loop ends would leave the faster lanes idle there.  Only
    // the score phases below are warp-wide.
    // Per-lane schedule state is kept small (the unrolled test block needs the registers): the
    // stage's constants are re-read from the kernel parameters at loop ends, and only what every
    // draw needs lives in registers (the inner-loop length, the step scale, the temperature).
    int stage = 0;
    bool done = !alive;
    double len0 = d.v[0], ratio0 = d.v[1], kt = 0., step_ratio = 1.0, score_start = score_current;
    double step_scale = 0.;  // max_step_size * step_ratio, the loop-invariant factor of optimisation.rs:115-119
    int convergence_count = 0;
    unsigned long long inner = 0, loops_left = 0, it = 0, loop_rejections = 0;
    auto begin_stage = [&]() {
        const KStage &S = K.stages[stage];
        len0 = d.v[0];  // generate_basis(): bounds captured per optimise_state call (cell.rs:139-152)
        ratio0 = d.v[1];
        rng_begin_stage<RNG>(rng, K.seed_base + replica, replica, n_basis);
        kt = S.kt_start;
        step_ratio = 1.0;
        step_scale = S.max_step * step_ratio;
        convergence_count = 0;
        inner = S.inner;
        loops_left = S.inner ? S.steps / S.inner : 0ull;
        it = 0;
        loop_rejections = 0;
        score_start = score_current;
    };
    auto next_stage = [&]() {  // skips stages that run no move (steps < inner_steps cannot happen after build())
        for (;;) {
            ++stage;
            if (stage >= K.n_stages) {
                done = true;
                return;
            }
            begin_stage();
            if (loops_left > 0) return;
        }
    };
    if (!done) {
        stage = -1;
        next_stage();
    }
    for (;;) {
        bool have = false;
        int slot = 0;
        double val = 0., new_val = 0., threshold = 0., new_score = 0., new_s = 0., new_c = 0.;
        // Draw until a proposal needs scoring.  Out-of-bounds proposals are rejected by
        // Basis::set_value without a score (optimisation.rs:120-123).  For hard shapes the
        // packing fraction depends on the cell DOFs only, so `Some(new_score)` is known
        // before any overlap test: when the Metropolis rule would reject even a valid
        // state (a cell move that lowers the score, lost against the threshold) the
        // reference's answer is "rejected" whether check_intersection says None or Some,
        // and the overlap test is skipped.  Draw order (index, step, threshold) and every
        // decision are unchanged.
        while (!done && !have) {
            if (it == inner) {  // end of an inner loop (optimisation.rs:138-167)
                const KStage &S = K.stages[stage];
                moves += inner;
                rejections += loop_rejections;
                kt *= S.kt_ratio;
                bool converged = false;
                if (S.convergence == S.convergence) {
                    if (score_current - score_start < S.convergence) {
                        if (++convergence_count > 5) converged = true;
                    } else {
                        convergence_count = 0;
                    }
                }
                if (!converged && step_ratio > 1e-4)
                    step_ratio *= (double)inner / ((double)loop_rejections + 1.);
                step_scale = S.max_step * step_ratio;
                --loops_left;
                it = 0;
                loop_rejections = 0;
                score_start = score_current;
                if (converged || loops_left == 0) next_stage();
                continue;
            }
            ++it;
            int basis_index;
            double sample;
            rng_draw_move<RNG>(rng, basis_index, sample);
            slot = basis_to_slot(basis_index, n_cell);
            val = d.get(slot);
            new_val = val + step_scale * sample;
            double lo, hi;
            slot_bounds(slot, len0, ratio0, lo, hi);
            have = lo <= new_val && new_val <= hi;
            if (!have) {
                ++loop_rejections;
                continue;
            }
            if constexpr (KIND != CP_LJ) {
                threshold = rng_draw_threshold<RNG>(rng);
                new_score = score_current;  // site moves leave the packing fraction as it is
                if (slot <= 2) {
                    double sa_new = t.sa;
                    if (slot == 2) {
                        cp_sincos(new_val, &new_s, &new_c);
                        sa_new = new_s;
                    }
                    const double len = slot == 0 ? new_val : d.v[0], ratio = slot == 1 ? new_val : d.v[1];
                    const double cell_area = sa_new * len * (len * ratio);  // cell.rs:190-192
                    new_score = (P.area * (double)P.n_syms) / cell_area;    // packed.rs:84
                    have = accept_move(new_score, score_current, kt, threshold);
                    if (!have) ++loop_rejections;
                }
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // phase A (lanes holding a proposal): geometry and candidate generation
        int candidates = 0;
        double old_s = 0., old_c = 0.;
        const CellCache old_cc = cc;
        if (have) {
            d.set(slot, new_val);
            if (slot == 5) {  // the site angle and the cell angle are the only DOFs inside sin/cos
                old_s = t.st; old_c = t.ct;
                cp_sincos(new_val, &t.st, &t.ct);
            }
            if (slot == 2) {
                old_s = t.sa; old_c = t.ca;
                if constexpr (KIND == CP_LJ) cp_sincos(new_val, &new_s, &new_c);
                t.sa = new_s; t.ca = new_c;
            }
            if (slot <= 2) cc = cell_cache_of(d);
            if constexpr (KIND == CP_LJ) {
                new_score = t_lj_score<NI>(P, g, d, t, cc);
            } else {
                t_build_geometry<KIND, NI>(P, g, d, t);
                candidates = t_sweep<KIND, NI>(P, g, d, t, cc, masks);
            }
        }
        // phase B (whole warp): shape-pair tests, pooled over the warp's replicas
        bool overlap = false;
        if constexpr (KIND != CP_LJ) overlap = t_drain_warp<KIND, NI>(P, smem, sh.pool, masks, candidates);
        // phase C: Metropolis decision
        if (have) {
            bool accepted;
            if constexpr (KIND == CP_LJ) {
                threshold = rng_draw_threshold<RNG>(rng);
                accepted = accept_move(new_score, score_current, kt, threshold);
            } else {
                accepted = !overlap;  // the rule itself was evaluated before the overlap test
            }
            if (accepted) {
                score_current = new_score;
            } else {
                d.set(slot, val);
                if (slot == 5) { t.st = old_s; t.ct = old_c; }
                if (slot == 2) { t.sa = old_s; t.ca = old_c; }
                cc = old_cc;
                ++loop_rejections;
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
    const ThreadShared sh = thread_shared(smem, KIND, P.n_items, P.n_syms);
    const unsigned long long rep = (unsigned long long)blockIdx.x * kThreadBlock + threadIdx.x;
    if (rep >= n) return;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = dof[rep * CP_MAX_DOF + i];
    out[rep] = t_state_score<KIND, NI>(P, sh.g, sh.m, d, trig_of(d));
}

template <int KIND, int NI>
__global__ void __launch_bounds__(kThreadBlock)
replay_thread_kernel(const __grid_constant__ KProblem P, const double *init_dof, const cp_move *mv,
                     unsigned long long n_moves, unsigned long long n, cp_decision *out) {
    extern __shared__ double smem[];
    const ThreadShared sh = thread_shared(smem, KIND, P.n_items, P.n_syms);
    const TGeo g = sh.g;
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
    double score_current = t_state_score<KIND, NI>(P, g, sh.m, d, t);
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
            D.new_score = t_state_score<KIND, NI>(P, g, sh.m, d, t);
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
