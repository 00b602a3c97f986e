// cp_device.cuh — device code of the Monte-Carlo packing engine (sm_100a).
//
// Shared by both kernel families: RNG streams, DOF bookkeeping, the f64 predicates, the Metropolis
// rule.  The default kernels give every replica one thread (cp_thread.cuh).  The lane-group
// kernels (cp_group.cuh; on request through cp_set_lanes) give one replica a group of L lanes
// (L = 4, 8, 16 or 32): the scalar replica state is replicated in the registers of every lane of
// the group, the per-move geometry is staged in shared memory by the group, and the sweeps over
// in-cell pairs and periodic-image candidates are strided over the lanes and combined with warp
// votes (any-overlap, early exit) or shuffle reductions (energy sum); their device functions
// (build_geometry, hard_overlap, lj_score, state_score) are the second half of this file.
//
// Arithmetic contract: every value that feeds a decision is computed with the reference's own
// expression order in IEEE f64 with no contraction (the file is compiled with -fmad=false), so
// scores, overlap predicates and accept/reject decisions are bit-identical to the CPU restatement
// of the reference.  Where an expression is rearranged, the comment proves the rearrangement
// yields the same decision.  Citations are relative to the reference tree.
#ifndef CP_DEVICE_CUH
#define CP_DEVICE_CUH

#include <stdint.h>

#include "../../include/cp_gpu.h"
#include "cp_math.h"

#define CP_PI 3.14159265358979323846264338327950288
#define CP_TAU 6.28318530717958647692528676655900577

struct KProblem {
    int kind, n_items, family, n_syms;
    double items[CP_MAX_ITEMS][5];
    double syms[CP_MAX_SYMS][6];
    double area, radius;
    double dof[CP_MAX_DOF];
    double lj_shift[CP_MAX_ITEMS];  // 4*eps*((sigma/x)^12 - (sigma/x)^6) per site (lj2.rs:50-51)
    double lj_reach;  // max cutoff + 2 * max |site position|; +inf when a site has no cutoff
    // single-precision copies for the certificates of the one-thread-per-replica kernels (cp_filter.h)
    float symsf[CP_MAX_SYMS][4];        // rotation part of the operators: S[0], S[1], S[3], S[4]
    float sgnf[CP_MAX_SYMS][2];         // signed_images: row signs of operator s relative to operator 0
    int signed_images;                  // every rotation part = operator 0's with the signs of its rows
                                        // flipped (+-1, exact): one set of rotated points serves all images
    float ptsf[2 * CP_MAX_ITEMS][2];    // polygon chain: the n start points; other polygons: start, end
                                        // of every segment; discs: the centres
    float discrf[CP_MAX_ITEMS];         // disc radii
    float radf;                         // enclosing radius
    int chain;                          // polygon whose segment k ends where segment k + 1 starts
    int replay_has_bounds;              // replay kernels: a second n x 6 block after init_dof holds the DOFs
                                        // generate_basis captured its bounds from (cp_replay_bounded)
};

struct KStage {
    double kt_start, kt_ratio, max_step;
    unsigned long long steps, inner;
    double convergence;
};

struct KParams {
    KProblem prob;
    KStage stages[CP_MAX_STAGES];
    int n_stages;
    int rng_mode;
    unsigned long long seed_base;
    unsigned long long replica_begin;
    unsigned long long n_replicas;
    const double *init_dof;  // nullable: n_replicas x 6
    cp_replica_result *out;
    int *error_flag;
    int draw_cap;  // thread kernels: proposals a lane may draw per pass (cp_thread.cuh)
    int warp_fill; // thread kernels: replicas per warp (<= 32; cp_thread_inst.cuh: thread_mc_launch)
};

// ------------------------------------------------------------------------------------------ RNG
struct Rng {
    // PCG: 128-bit MCG state.  Philox: (ctr) in lo, replica id in hi.
    unsigned long long lo, hi;
    unsigned long long key;
    unsigned long long zone;  // Uniform<usize> rejection zone (rand 0.8.4 uniform.rs)
    unsigned long long range;
    unsigned long long thr;   // Philox: second output word of the move's block (threshold bits)
};

__device__ __forceinline__ unsigned long long rotr64(unsigned long long v, unsigned r) {
    return (v >> r) | (v << ((64u - r) & 63u));
}

// rand_pcg 0.3.1 Mcg128Xsl64::next_u64
__device__ __forceinline__ unsigned long long pcg_next(Rng &r) {
    const unsigned long long MH = 0x2360ED051FC65DA4ull, ML = 0x4385DF649FCCF645ull;
    unsigned long long nlo = r.lo * ML;
    unsigned long long nhi = __umul64hi(r.lo, ML) + r.lo * MH + r.hi * ML;
    r.lo = nlo;
    r.hi = nhi;
    unsigned rot = (unsigned)(nhi >> 58);
    return rotr64(nhi ^ nlo, rot);
}

// rand_core 0.6.3 seed_from_u64 -> Pcg64Mcg::from_seed
__device__ __forceinline__ void pcg_seed(Rng &r, unsigned long long state) {
    const unsigned long long MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        state = state * MUL + INC;
        unsigned xs = (unsigned)(((state >> 18) ^ state) >> 27);
        unsigned rot = (unsigned)(state >> 59);
        w[i] = (xs >> rot) | (xs << ((32u - rot) & 31u));
    }
    r.lo = ((unsigned long long)w[0] | ((unsigned long long)w[1] << 32)) | 1ull;
    r.hi = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1) -> two u64
__device__ __forceinline__ void philox(unsigned long long ctr, unsigned long long id,
                                       unsigned long long key, unsigned long long &o0,
                                       unsigned long long &o1) {
    unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = (unsigned)id, c3 = (unsigned)(id >> 32);
    unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o0 = (unsigned long long)c0 | ((unsigned long long)c1 << 32);
    o1 = (unsigned long long)c2 | ((unsigned long long)c3 << 32);
}

template <int RNG>
__device__ __forceinline__ void rng_begin_stage(Rng &r, unsigned long long seed_base,
                                                unsigned long long replica, int n_basis) {
    r.range = (unsigned long long)n_basis;
    r.zone = ~0ull - ((0ull - r.range) % r.range);
    if (RNG == CP_RNG_PCG64MCG) {
        pcg_seed(r, seed_base + replica);  // seed = index for every stage (main.rs:229,239,249)
    } else {
        // Philox: the key is the job's seed (the same for every replica, so the ten round keys are
        // warp-uniform), the stream is selected by the global replica id in the counter's upper half
        r.lo = 0;         // draw-block counter; restarts every stage like the reference's re-seeding
        r.hi = replica;   // stream id
        r.key = seed_base;
    }
}

// rand 0.8.4 conversions
__device__ __forceinline__ double u64_to_step(unsigned long long v) {
    double value1_2 = __longlong_as_double((long long)((v >> 12) | 0x3FF0000000000000ull));
    return (value1_2 - 1.0) * 1.0 + -0.5;
}
__device__ __forceinline__ double u64_to_unit(unsigned long long v) {
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// draws #1 and #2 of a move: basis index and step sample (optimisation.rs:104,119)
template <int RNG>
__device__ __forceinline__ void rng_draw_move(Rng &r, int &basis_index, double &sample) {
    if (RNG == CP_RNG_PCG64MCG) {
        for (;;) {
            unsigned long long v = pcg_next(r);
            unsigned long long lo = v * r.range;
            if (lo <= r.zone) {
                basis_index = (int)__umul64hi(v, r.range);
                break;
            }
        }
        sample = u64_to_step(pcg_next(r));
    } else {
        // one Philox block per move: word a -> step (its top 52 bits), word b -> threshold (its top
        // 53 bits), the 12 + 11 bits those conversions drop -> basis index (bias <= 6 * 2^-23)
        unsigned long long a, b;
        philox(r.lo, r.hi, r.key, a, b);
        r.lo += 1;
        r.thr = b;
        const unsigned spare = (unsigned)(((a & 0xFFFull) << 11) | (b & 0x7FFull));
        basis_index = (int)((spare * (unsigned)r.range) >> 23);
        sample = u64_to_step(a);
    }
}

// draw #3 as raw bits: the conversion to [0, 1) is needed only where the Metropolis rule gets as
// far as comparing against the threshold
template <int RNG>
__device__ __forceinline__ unsigned long long rng_draw_threshold_bits(Rng &r) {
    if (RNG == CP_RNG_PCG64MCG) {
        return pcg_next(r);
    } else {
        return r.thr;
    }
}

// draw #3: the acceptance threshold (optimisation.rs:62), only for in-bounds proposals
template <int RNG>
__device__ __forceinline__ double rng_draw_threshold(Rng &r) {
    if (RNG == CP_RNG_PCG64MCG) {
        return u64_to_unit(pcg_next(r));
    } else {
        return u64_to_unit(r.thr);
    }
}

// ------------------------------------------------------------------------------------------ state
struct Dof {
    double v[CP_MAX_DOF];  // length, ratio, angle, x, y, theta
    __device__ __forceinline__ double get(int slot) const {
        double r = v[0];
        r = slot == 1 ? v[1] : r;
        r = slot == 2 ? v[2] : r;
        r = slot == 3 ? v[3] : r;
        r = slot == 4 ? v[4] : r;
        r = slot == 5 ? v[5] : r;
        return r;
    }
    __device__ __forceinline__ void set(int slot, double x) {
#pragma unroll
        for (int i = 0; i < CP_MAX_DOF; ++i) v[i] = slot == i ? x : v[i];
    }
};

// number of cell DOFs in the basis (cell.rs:136-170); site DOFs follow (site.rs:65-91)
__device__ __forceinline__ int n_cell_dof(int family) {
    return family == CP_MONOCLINIC ? 3 : (family == CP_ORTHORHOMBIC ? 2 : 1);
}
__device__ __forceinline__ int basis_to_slot(int basis_index, int n_cell) {
    return basis_index < n_cell ? basis_index : basis_index - n_cell + 3;
}
// bounds of a slot as generate_basis captures them (cell.rs:139-160, site.rs:69-88)
__device__ __forceinline__ void slot_bounds(int slot, double len0, double ratio0, double &lo, double &hi) {
    lo = 0.01;
    hi = len0;
    if (slot == 1) { lo = 0.1; hi = ratio0; }
    if (slot == 2) { lo = CP_PI / 4.; hi = CP_PI / 2.; }
    if (slot == 3 || slot == 4) { lo = -0.5; hi = 0.5; }
    if (slot == 5) { lo = 0.; hi = CP_TAU; }
}

// ------------------------------------------------------------------------------------------ geometry
struct Trig {
    double st, ct, sa, ca;  // sin/cos of the site angle (theta) and of the cell angle
};

// shared-memory view of one group's per-move geometry
struct Geo {
    double *img;   // [N][8]: m00 m01 m10 m11 | fx fy (fractional, wrapped) | cx cy (Cartesian)
    double *comp;  // POLYGON [N][n][8]: rsx rsy rex rey (rotated) | sx sy (placed start) | dx dy
                   // DISCS/LJ [N][n][4]: rx ry (rotated) | px py (placed)
};

template <int L>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (L == 32) {
        return 0xffffffffu;
    } else {
        unsigned lane = threadIdx.x & 31u;
        return ((1u << L) - 1u) << (lane / L * L);
    }
}
template <int L>
__device__ __forceinline__ bool group_any(unsigned gmask, bool p) {
    if (L == 32) return __any_sync(0xffffffffu, p);
    return (__ballot_sync(gmask, p) & gmask) != 0u;
}
template <int L>
__device__ __forceinline__ unsigned group_ballot(unsigned gmask, bool p) {
    unsigned b = __ballot_sync(gmask, p);
    if constexpr (L == 32) {
        return b;
    } else {
        unsigned lane = threadIdx.x & 31u;
        return (b >> (lane / L * L)) & ((1u << L) - 1u);
    }
}

// Transform2::periodic(1., -0.5) on one coordinate (transform.rs:107-112):
//   (((v - (-0.5)) % 1.) + 1.) % 1. + (-0.5)
// fmod(t, 1) == t - trunc(t) exactly for finite t (the result is representable), so no libm call.
__device__ __forceinline__ double wrap_unit(double v) {
    double t = v - (-0.5);
    t = t - trunc(t);
    t = t + 1.0;
    t = t - trunc(t);
    return t + (-0.5);
}

// Line2::intersects (shape/components/line2.rs:29-54) on (start, delta) pairs, division-free.
// The reference divides: ua = ua_t / u_b and tests 0 <= ua <= 1 (same for ub).  For finite
// operands with u_b != 0 and a correctly rounded quotient q = RN(t / d):
//   q >= 0  <=>  t == 0 or sign(t) == sign(d)          (a quotient of opposite signs rounds to a
//                                                        negative number or, only on underflow, -0)
//   q <= 1  <=>  |t| <= |d|                             (|t| > |d| means |t| >= nextafter(|d|), whose
//                                                        quotient exceeds 1 + 2^-53 and cannot round to 1)
// Underflow of t/d needs |t| < 2^-1022 |d|, unreachable for differences of products of O(1)
// coordinates, and the DOF bounds keep every operand finite.
// Written without short-circuit operators so that a run of independent tests is one basic block
// (instruction-level parallelism instead of a branch per test).  With tt = t carrying d's sign
// flipped onto it (the bit pattern of t * sign(d), one integer XOR):
//   t == 0 or sign(t) == sign(d)   <=>  tt >= 0      (-0 >= 0 holds in IEEE comparison)
//   and then  |t| <= |d|           <=>  tt <= |d|
// so each quotient test is two chained FP64 compares.  d != 0 is the caller's u_b != 0 factor;
// operands are finite by the DOF bounds.
__device__ __forceinline__ bool unit_range(double t, double d) {
    const double tt = __hiloint2double(__double2hiint(t) ^ (__double2hiint(d) & (int)0x80000000), __double2loint(t));
    return (tt >= 0.0) & (tt <= fabs(d));
}
__device__ __forceinline__ bool segments_cross(double sx, double sy, double sdx, double sdy,
                                               double ox, double oy, double odx, double ody) {
    double u_b = ody * sdx - odx * sdy;
    double wy = sy - oy, wx = sx - ox;
    double ua_t = odx * wy - ody * wx;
    double ub_t = sdx * wy - sdy * wx;
    return (u_b != 0.0) & unit_range(ua_t, u_b) & unit_range(ub_t, u_b);
}

// LJ2::energy (shape/components/lj2.rs:43-60); parameters of `self` only
__device__ __forceinline__ double lj_pair(double dx, double dy, double sigma, double eps,
                                          double cutoff, double shift) {
    double r2 = dx * dx + dy * dy;
    if (cutoff == cutoff) {
        if (!(r2 < cutoff * cutoff)) return 0.0;  // the quotient below is dead on this branch
        double q = (sigma * sigma) / r2;
        double c3 = (q * q) * q;
        return 4. * eps * (c3 * c3 - c3) - shift;
    }
    double q = (sigma * sigma) / r2;
    double c3 = (q * q) * q;
    return 4. * eps * (c3 * c3 - c3);
}

// Build the per-move geometry of the group's replica in shared memory:
// OccupiedSite::positions (site.rs:33-45), Cell2::to_cartesian_isometry (cell.rs:110-112,258-263)
// and Shape::transform (line_shape.rs:103-108 etc.).
template <int KIND, int L>
__device__ __forceinline__ void build_geometry(const KProblem &P, const double *s_items, Geo g,
                                               const Dof &d, double st, double ct, double sa,
                                               double ca, unsigned gmask, int lane) {
    const int N = P.n_syms, n = P.n_items;
    const double a = d.v[0], b = d.v[0] * d.v[1];  // Cell2::a / b (cell.rs:91-97)
    for (int s = lane; s < N; s += L) {
        // sym * Transform2::new(theta, (x, y)) with nalgebra's Matrix3 product order;
        // the operator's row 2 is zero, so only rows 0-1 of the product are ever used.
        const double *S = P.syms[s];
        double m00 = ((S[0] * ct) + S[1] * st) + S[2] * 0.0;
        double m01 = ((S[0] * -st) + S[1] * ct) + S[2] * 0.0;
        double tx = ((S[0] * d.v[3]) + S[1] * d.v[4]) + S[2] * 1.0;
        double m10 = ((S[3] * ct) + S[4] * st) + S[5] * 0.0;
        double m11 = ((S[3] * -st) + S[4] * ct) + S[5] * 0.0;
        double ty = ((S[3] * d.v[3]) + S[4] * d.v[4]) + S[5] * 1.0;
        double fx = wrap_unit(tx), fy = wrap_unit(ty);
        double yb = fy * b;
        double *o = g.img + s * 8;
        o[0] = m00; o[1] = m01; o[2] = m10; o[3] = m11;
        o[4] = fx;  o[5] = fy;
        o[6] = fx * a + yb * ca;  // to_cartesian (cell.rs:258-263)
        o[7] = yb * sa;
    }
    __syncwarp(gmask);
    for (int t = lane; t < N * n; t += L) {
        int s = t / n, k = t - s * n;
        const double *im = g.img + s * 8;
        const double *it = s_items + k * 5;
        // Transform * Point = (m00*px + m01*py) + tx; the normaliser row is zero
        double rx = im[0] * it[0] + im[1] * it[1];
        double ry = im[2] * it[0] + im[3] * it[1];
        if (KIND == CP_POLYGON) {
            double ex = im[0] * it[2] + im[1] * it[3];
            double ey = im[2] * it[2] + im[3] * it[3];
            double sx = rx + im[6], sy = ry + im[7];
            double px = ex + im[6], py = ey + im[7];
            double *o = g.comp + t * 8;
            o[0] = rx; o[1] = ry; o[2] = ex; o[3] = ey;
            o[4] = sx; o[5] = sy;
            o[6] = px - sx;  // Line2::dx / dy of the placed segment (line2.rs:94-101)
            o[7] = py - sy;
        } else {
            double *o = g.comp + t * 4;
            o[0] = rx; o[1] = ry;
            o[2] = rx + im[6];
            o[3] = ry + im[7];
        }
    }
    __syncwarp(gmask);
}

// periodic_range of PackedState::check_intersection (state/packed.rs:128-132)
__device__ __forceinline__ int periodic_range(double a, double b, double angle) {
    double p = a / b;
    double off = fabs(angle - CP_PI / 2.);
    if (0.5 < p && p < 2. && off < 0.2) return 1;
    if (0.3 < p && p < 3. && off < 0.5) return 2;
    return 3;
}

__device__ __forceinline__ void pair_from_index(int q, int N, int &i, int &j) {
    // q-th pair (i < j) in the order (0,1),(0,2),...,(N-2,N-1); N <= 4
    i = 0;
    int row = N - 1;
    while (q >= row) { q -= row; ++i; --row; }
    j = i + 1 + q;
}

// PackedState::check_intersection (state/packed.rs:127-167): true when any two shapes overlap,
// inside the cell or across the periodic images.  The boolean is an OR over all tests, so the
// order in which lanes evaluate them does not change it.
template <int KIND, int L>
__device__ __forceinline__ bool hard_overlap(const KProblem &P, const double *s_items, Geo g,
                                             const Dof &d, double sa, double ca, unsigned gmask,
                                             int lane) {
    const int N = P.n_syms, n = P.n_items, nn = n * n;
    constexpr int CW = KIND == CP_POLYGON ? 8 : 4;
    // --- within the cell: every pair i < j, no distance filter (packed.rs:134-148)
    const int n_pairs = N * (N - 1) / 2;
    {
        bool hit = false;
        for (int t = lane; t < n_pairs * nn; t += L) {
            int q = t / nn, r = t - q * nn;
            int k = r / n, l = r - k * n;
            int i, j;
            pair_from_index(q, N, i, j);
            const double *A = g.comp + (i * n + k) * CW, *B = g.comp + (j * n + l) * CW;
            if (KIND == CP_POLYGON) {
                hit |= segments_cross(A[4], A[5], A[6], A[7], B[4], B[5], B[6], B[7]);
            } else {
                // Atom2::intersects (atom2.rs:21-26): strict <
                double rs = s_items[k * 5 + 2] + s_items[l * 5 + 2];
                double dx = A[2] - B[2], dy = A[3] - B[3];
                hit |= (dx * dx + dy * dy) < rs * rs;
            }
        }
        if (group_any<L>(gmask, hit)) return true;
    }
    // --- periodic images (packed.rs:150-165)
    const double a = d.v[0], b = d.v[0] * d.v[1];
    const int shells = periodic_range(a, b, d.v[2]);
    const int side = 2 * shells + 1, n_img = side * side;
    const double r2x = P.radius * 2.;
    const double radius_sq = r2x * r2x;
    const int total = N * N * n_img;
    for (int base = 0; base < total; base += L) {
        int c = base + lane;
        bool pass = false;
        int ci = 0, cj = 0;
        double X = 0., Y = 0.;
        if (c < total) {
            int ij = c / n_img, q = c - ij * n_img;
            ci = ij / N;
            cj = ij - ci * N;
            int ix = q / side - shells, iy = q - (q / side) * side - shells;
            if (ix != 0 || iy != 0) {
                // to_cartesian_translate (cell.rs:241-245): Translation2(ix,iy) * position
                double fx = g.img[cj * 8 + 4] + (double)ix;
                double fy = g.img[cj * 8 + 5] + (double)iy;
                double yb = fy * b;
                X = fx * a + yb * ca;
                Y = yb * sa;
                double ddx = g.img[ci * 8 + 6] - X, ddy = g.img[ci * 8 + 7] - Y;
                pass = (ddx * ddx + ddy * ddy) <= radius_sq;  // packed.rs:156-157
            }
        }
        unsigned todo = group_ballot<L>(gmask, pass);
        while (todo) {
            int src = __ffs(todo) - 1;
            todo &= todo - 1;
            int i = __shfl_sync(gmask, ci, src, L), j = __shfl_sync(gmask, cj, src, L);
            double IX = __shfl_sync(gmask, X, src, L), IY = __shfl_sync(gmask, Y, src, L);
            bool hit = false;
            for (int t = lane; t < nn; t += L) {
                int k = t / n, l = t - k * n;
                const double *A = g.comp + (i * n + k) * CW, *B = g.comp + (j * n + l) * CW;
                if (KIND == CP_POLYGON) {
                    // shape.transform(&transform2): rotated point + image position
                    double osx = B[0] + IX, osy = B[1] + IY;
                    double oex = B[2] + IX, oey = B[3] + IY;
                    hit |= segments_cross(A[4], A[5], A[6], A[7], osx, osy, oex - osx, oey - osy);
                } else {
                    double rs = s_items[k * 5 + 2] + s_items[l * 5 + 2];
                    double dx = A[2] - (B[0] + IX), dy = A[3] - (B[1] + IY);
                    hit |= (dx * dx + dy * dy) < rs * rs;
                }
            }
            if (group_any<L>(gmask, hit)) return true;
        }
    }
    return false;
}

// LJShape2::energy (shape/lj_shape.rs:35-39): iproduct!(self.items, other.items) summed in order;
// `other` is shape j displaced by (X, Y) (the rotated points of j plus the image position, or its
// in-cell placement when in_cell is set).
__device__ __forceinline__ double shape_energy(const double *s_items, const double *s_shift, Geo g,
                                               int n, int i, int j, bool in_cell, double X, double Y) {
    double e = 0.;
    for (int k = 0; k < n; ++k) {
        const double *A = g.comp + (i * n + k) * 4;
        const double *it = s_items + k * 5;
        for (int l = 0; l < n; ++l) {
            const double *B = g.comp + (j * n + l) * 4;
            double ox = in_cell ? B[2] : B[0] + X, oy = in_cell ? B[3] : B[1] + Y;
            e += lj_pair(A[2] - ox, A[3] - oy, it[2], it[3], it[4], s_shift[k]);
        }
    }
    return e;
}

// sum += e over the lanes of the group in lane order.  Adding +-0.0 never changes a sum that is not
// -0.0 (and this sum starts at +0.0, so it never is), hence zero terms are skipped; the non-zero
// ones are added one by one in the reference's order, in every lane alike.
template <int L>
__device__ __forceinline__ double ordered_add(double sum, double e, unsigned gmask) {
    unsigned nz = group_ballot<L>(gmask, e != 0.0);
    while (nz) {
        int src = __ffs(nz) - 1;
        nz &= nz - 1;
        sum += __shfl_sync(gmask, e, src, L);
    }
    return sum;
}

// PotentialState::score (state/potential.rs:82-115): -(sum of shape-pair energies)/N over the
// in-cell pairs i < j and then, for every i, j, the 48 images of three shells in iproduct order.
// Lanes evaluate one shape pair each (the expensive part); the accumulation keeps the reference's
// sequential order (ordered_add), so the energy is bit-identical to the CPU restatement.
template <int L>
__device__ __forceinline__ double lj_score(const KProblem &P, const double *s_items,
                                           const double *s_shift, Geo g, const Dof &d, double sa,
                                           double ca, unsigned gmask, int lane) {
    const int N = P.n_syms, n = P.n_items;
    double sum = 0.;
    const int n_pairs = N * (N - 1) / 2;
    for (int base = 0; base < n_pairs; base += L) {
        int q = base + lane;
        double e = 0.;
        if (q < n_pairs) {
            int i, j;
            pair_from_index(q, N, i, j);
            e = shape_energy(s_items, s_shift, g, n, i, j, true, 0., 0.);
        }
        sum = ordered_add<L>(sum, e, gmask);
    }
    const double a = d.v[0], b = d.v[0] * d.v[1];
    const int side = 7, n_img = 49;  // always 3 shells (potential.rs:105)
    const int total = N * N * n_img;
    for (int base = 0; base < total; base += L) {
        int c = base + lane;
        double e = 0.;
        if (c < total) {
            int ij = c / n_img, q = c - ij * n_img;
            int i = ij / N, j = ij - i * N;
            int ix = q / side - 3, iy = q - (q / side) * side - 3;
            if (ix != 0 || iy != 0) {
                double fx = g.img[j * 8 + 4] + (double)ix;
                double fy = g.img[j * 8 + 5] + (double)iy;
                double yb = fy * b;
                e = shape_energy(s_items, s_shift, g, n, i, j, false, fx * a + yb * ca, yb * sa);
            }
        }
        sum = ordered_add<L>(sum, e, gmask);
    }
    return -sum / (double)N;
}

// State::score (packed.rs:80-86 / potential.rs:82-115); NaN = None
template <int KIND, int L>
__device__ __forceinline__ double state_score(const KProblem &P, const double *s_items,
                                              const double *s_shift, Geo g, const Dof &d,
                                              const Trig &trig, unsigned gmask, int lane) {
    // trig.st/ct: Transform2::new(theta, ..) (transform.rs:81-87); trig.sa/ca: the cell angle
    // (cell.rs:190-192,258-263).  The caller keeps them in step with the DOFs: sin/cos are pure
    // functions of one DOF each, so caching them per DOF is arithmetically identical.
    const double st = trig.st, ct = trig.ct, sa = trig.sa, ca = trig.ca;
    build_geometry<KIND, L>(P, s_items, g, d, st, ct, sa, ca, gmask, lane);
    double score;
    if (KIND == CP_LJ) {
        score = lj_score<L>(P, s_items, s_shift, g, d, sa, ca, gmask, lane);
    } else {
        bool overlap = hard_overlap<KIND, L>(P, s_items, g, d, sa, ca, gmask, lane);
        double cell_area = sa * d.v[0] * (d.v[0] * d.v[1]);  // sin(angle) * a * b (cell.rs:190-192)
        score = overlap ? __longlong_as_double(0x7FF8000000000000ll)
                        : (P.area * (double)P.n_syms) / cell_area;
    }
    __syncwarp(gmask);  // geometry is rewritten by the next move
    return score;
}

__device__ __forceinline__ Trig trig_of(const Dof &d) {
    Trig t;
    cp_sincos(d.v[5], &t.st, &t.ct);
    cp_sincos(d.v[2], &t.sa, &t.ca);
    return t;
}

// one out-of-line copy of the transcendental functions for the thread kernels (instruction cache:
// the kernels call them from several places, a few lanes at a time)
struct SinCos {
    double s, c;
};
static __device__ __noinline__ SinCos dev_sincos(double x) {
    SinCos r;
    cp_sincos(x, &r.s, &r.c);
    return r;
}
static __device__ __noinline__ double dev_exp(double x) { return cp_exp(x); }

// accept_move with the cheap cases decided first; the result is the same in every case:
//   new > old                      -> accepted (optimisation.rs:66)
//   kt == 0 (stages 1 and 3)       -> (new - old) / 0 is -inf or NaN: exp gives 0 (rejected: the
//                                     threshold is never negative) or NaN (new == old; f64::min
//                                     ignores it: surface 1, accepted since the threshold is < 1)
//   (new - old) / kt < -745.13..   -> cp_exp returns exactly 0.0 there: rejected
__device__ __forceinline__ bool accept_move_fast(double new_score, double old, double kt, unsigned long long threshold_bits) {
    if (!(new_score == new_score)) return false;  // None
    if (new_score > old) return true;
    if (__double_as_longlong(kt) == 0ll) return new_score == old;  // +0.0 only: -0.0 flips the sign of the quotient
    const double x = (new_score - old) / kt;
    if (x < -7.45133219101941108420e+02) return false;
    const double e = dev_exp(x);
    const double surface = (e != e) ? 1.0 : (e < 1.0 ? e : 1.0);
    return u64_to_unit(threshold_bits) < surface;
}

// MCOptimiser::accept_score (optimisation.rs:55-78) with the threshold already drawn
__device__ __forceinline__ bool accept_move(double new_score, double old, double kt, double threshold) {
    if (!(new_score == new_score)) return false;  // None
    if (new_score > old) return true;
    // energy_surface: f64::min(exp((new - old) / kt), 1.) -- Rust's min ignores a NaN operand
    double e = cp_exp((new_score - old) / kt);
    double surface = (e != e) ? 1.0 : (e < 1.0 ? e : 1.0);
    return threshold < surface;
}

#endif  // CP_DEVICE_CUH
