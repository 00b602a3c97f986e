// cp_thread_lj.cuh — PotentialState::score for one replica in one thread (LJShape2): geometry
// layout and energy sum of the one-thread-per-replica LJ kernels.  Included by cp_thread.cuh.
#ifndef CP_THREAD_LJ_CUH
#define CP_THREAD_LJ_CUH

#include "cp_device.cuh"

// the filtered evaluation (t_lj_score_pooled) serves shapes of up to this many sites; larger shapes
// and sites without cutoff take the plain loops
constexpr int kLjQueueSites = 4;

// doubles of shared memory each LJ thread needs: four cell values, per image fx fy cx cy, per stored
// site the rotated position (f64), then the same positions in single precision (one double = x, y of
// a site), the image masks of the shape pairs i < j and of a shape with itself (64 bits each) and two
// elements of single-precision sweep parameters (af, bxf | byf, ec).  The warp's pool (32 entries,
// 32 energies) follows the 32 columns.
// Stored sites: those of image 0 only when every operator's rotation part is operator 0's with the
// signs of its rows flipped (KProblem::signed_images - all seven plane groups of the reference): the
// rotated sites of image s are then image 0's with those signs, bit for bit up to the sign of a zero,
// which no energy can see (a rotated coordinate only enters a sum with the image's position and the
// difference of two such sums is squared).  Otherwise the sites of every image.
__host__ __device__ inline int lj_mask_words(int n_syms) { return n_syms * (n_syms - 1) / 2 + 1; }
__host__ __device__ inline int lj_stored_images(const KProblem &P) { return P.signed_images ? 1 : P.n_syms; }
__host__ __device__ inline int lj_geo_doubles(const KProblem &P) {
    return 4 + P.n_syms * 4 + lj_stored_images(P) * P.n_items * 3 + lj_mask_words(P.n_syms) + 2;
}
__host__ __device__ inline size_t lj_block_bytes(const KProblem &P) {
    return sizeof(double) * 32 * (size_t)(lj_geo_doubles(P) + 1) + 32 * sizeof(unsigned);
}

// sign flips as integer operations (the FP64 pipe is what the dense LJ regime is short of)
__device__ __forceinline__ double lj_flip(double v, unsigned sign) {
    return __hiloint2double(__double2hiint(v) ^ (int)sign, __double2loint(v));
}
__device__ __forceinline__ float lj_flipf(float v, unsigned sign) { return __uint_as_float(__float_as_uint(v) ^ sign); }

// a thread's column of the [element][thread] planes (STRIDE = 32)
template <int STRIDE>
struct TGeoT {
    double *base;  // element e is base[e * STRIDE]
    int comp0;     // first site element (= 4 + N * 4); images start at element 4
    int stored;    // rotated sites stored: (images stored) * n
    int words;     // mask words
    int geo;       // doubles per column
    __device__ __forceinline__ double ld(int e) const { return base[e * STRIDE]; }
    __device__ __forceinline__ void st(int e, double v) const { base[e * STRIDE] = v; }
    // single-precision copy of the stored site c: element comp0 + 2 * stored + c
    __device__ __forceinline__ float2 ldf(int c) const {
        return *reinterpret_cast<const float2 *>(base + (comp0 + 2 * stored + c) * STRIDE);
    }
    __device__ __forceinline__ void stf(int c, float x, float y) const {
        *reinterpret_cast<float2 *>(base + (comp0 + 2 * stored + c) * STRIDE) = make_float2(x, y);
    }
    // image mask m: element comp0 + 3 * stored + m
    __device__ __forceinline__ unsigned long long *mask(int m) const {
        return reinterpret_cast<unsigned long long *>(base + (comp0 + 3 * stored + m) * STRIDE);
    }
    // sweep parameters: two elements behind the masks
    __device__ __forceinline__ float2 *fpar(int e) const {
        return reinterpret_cast<float2 *>(base + (comp0 + 3 * stored + words + e) * STRIDE);
    }
    // the same column of another lane of the warp
    __device__ __forceinline__ TGeoT of(unsigned lane) const {
        TGeoT o = *this;
        o.base = base - threadIdx.x + lane;
        return o;
    }
    // the warp's pool behind the 32 columns (call on the lane's OWN column): 32 energies, 32 entries
    __device__ __forceinline__ double *pool_energies() const { return base - threadIdx.x + 32 * geo; }
    __device__ __forceinline__ unsigned *pool_entries() const { return reinterpret_cast<unsigned *>(pool_energies() + 32); }
};
using TGeo = TGeoT<32>;
// this thread's column
__device__ __forceinline__ TGeo lj_geo(double *smem, const KProblem &P) {
    return TGeo{smem + threadIdx.x, 4 + P.n_syms * 4, lj_stored_images(P) * P.n_items, lj_mask_words(P.n_syms), lj_geo_doubles(P)};
}
// Where the rotated sites of image s are stored and the signs to apply to them: element of site 0 (x;
// y follows), sign masks of x and y.
struct LjImage {
    int e0;
    unsigned sx, sy;
};
__device__ __forceinline__ LjImage lj_image(const KProblem &P, const TGeo &g, int n, int s) {
    if (g.stored == n && P.n_syms > 1)  // image 0's sites with this image's row signs
        return LjImage{g.comp0, __float_as_uint(P.sgnf[s][0]) & 0x80000000u, __float_as_uint(P.sgnf[s][1]) & 0x80000000u};
    return LjImage{g.comp0 + s * n * 2, 0u, 0u};
}

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
        if (s * n >= g.stored) continue;  // signed images: the sites of image 0 serve every image
#pragma unroll
        for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
            if (NI == 0 && k >= n) break;
            const double *it = P.items[k];
            const int e = g.comp0 + (s * n + k) * CW;
            const double rx = m00 * it[0] + m01 * it[1], ry = m10 * it[0] + m11 * it[1];
            g.st(e + 0, rx);  // Transform * Point without its translation
            g.st(e + 1, ry);
            g.stf(s * n + k, (float)rx, (float)ry);  // the pre-filter's copy (t_lj_score)
        }
    }
}

// LJShape2::energy (shape/lj_shape.rs:35-39): iproduct!(self.items, other.items) summed in order,
// `other` = shape j displaced to (X, Y)
template <int NI>
__device__ __forceinline__ double t_shape_energy(const KProblem &P, TGeo g, int i, int j, double X, double Y) {
    const int n = NI > 0 ? NI : P.n_items;
    const LjImage Ii = lj_image(P, g, n, i), Ij = lj_image(P, g, n, j);
    const int ei = Ii.e0, ej = Ij.e0;
    const double cx = g.ld(4 + i * 4 + 2), cy = g.ld(4 + i * 4 + 3);
    double e = 0.;
    // the run-time-count build (NI = 0) keeps these loops rolled: the function is inlined at nine
    // call sites of t_lj_score and 16 x 16 unrolled pair evaluations each would not compile in
    // reasonable time
#pragma unroll(NI > 0 ? NI : 1)
    for (int k = 0; k < (NI > 0 ? NI : CP_MAX_ITEMS); ++k) {
        if (NI == 0 && k >= n) break;
        const double sx = lj_flip(g.ld(ei + k * 2 + 0), Ii.sx) + cx, sy = lj_flip(g.ld(ei + k * 2 + 1), Ii.sy) + cy;
        const double *it = P.items[k];
#pragma unroll(NI > 0 ? NI : 1)
        for (int l = 0; l < (NI > 0 ? NI : CP_MAX_ITEMS); ++l) {
            if (NI == 0 && l >= n) break;
            const double ox = lj_flip(g.ld(ej + l * 2 + 0), Ij.sx) + X, oy = lj_flip(g.ld(ej + l * 2 + 1), Ij.sy) + Y;
            e += lj_pair(sx - ox, sy - oy, it[2], it[3], it[4], P.lj_shift[k]);
        }
    }
    return e;
}

// PotentialState::score (state/potential.rs:82-115) for one replica in one thread, plain loops (shapes
// of more than kLjQueueSites sites): the reference's accumulation order is simply the loop order here.  Exact pruning: when every site has a cutoff,
// a shape pair whose centres are further apart than max cutoff + 2 max|site| (P.lj_reach, widened
// by 1e-9) has all its site distances beyond the cutoff, each LJ2::energy is exactly 0.0 and adding
// it cannot change the sum - so the pair, its row or its columns are skipped outright.
template <int NI>
__device__ __noinline__ double t_lj_score_plain(const KProblem &P, TGeo g, const Dof &d, const Trig &t) {
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

// ---------------------------------------------------------------------------------- filtered, pooled evaluation
// The same sum for shapes of up to kLjQueueSites sites, organised for SIMT.
//   1. every lane sweeps the 48 images of its shape pairs in single precision into bit masks of the
//      images whose centre can be within reach (a superset: an image outside the mask has every site
//      distance beyond every cutoff, so each LJ2::energy is exactly 0.0 and the sum is unchanged;
//      an extra image inside the mask only costs its tests).  One sweep per pair i < j - the mask of
//      (j, i) is its mirror image - and one for a shape with itself;
//   2. the candidate images of all 32 replicas are drawn, in rounds, into a pool of 32 entries, each
//      replica's in the reference's order (potential.rs:88-110: in-cell pairs i < j, then i, j, x
//      outer, y inner).  Each lane takes ONE entry, whoever owns it: it pre-filters the n x n site
//      pairs in single precision against (cutoff + margin)^2 - again a superset of the pairs inside
//      their cutoff - and evaluates the survivors (a few dozen per move out of ~7,000) with the
//      reference's f64 expressions, cutoff test included (lj2.rs:43-60), adding them up in iproduct
//      order: that is the energy of one shape pair (lj_shape.rs:35-39);
//   3. the owner adds the shape-pair energies of its entries to its sum, in order.
// Why a pool: the replicas of a warp differ enormously in the work one score needs.  Most sit in
// dilute cells with a handful of images in reach, a few have collapsed into dense, elongated cells
// with dozens (measured after 2,000 steps of the trimer in p2gg: 2 images for 85 % of the replicas,
// 12 to 42 for the rest), and a loop per thread runs at the pace of the busiest lane
// (profiles/r2_e_lj_queued.txt: 9 of 32 lanes active).
// The order of the additions is the reference's: site-pair energies of one shape pair in iproduct
// order, shape-pair energies in sequence order, each by one lane, one after the other.  Terms that
// are skipped are exactly +0.0 and neither a shape-pair energy nor the sum can be -0.0 (both start
// at +0.0, and x + (-x) = +0.0), so skipping them - or adding a +0.0 - leaves every bit as it is.
struct LjSweep {
    float af, bxf, byf;  // lattice vectors a = (af, 0), b = (bxf, byf)
    float wide, wide2;   // reach + 2 E and its square, E = bound of the single-precision displacement error
    float ec;            // bound of the error of a single-precision site-pair displacement component
};
__device__ __forceinline__ LjSweep lj_sweep_setup(const KProblem &P, const Dof &d, const Trig &t) {
    LjSweep s;
    const double b = d.v[0] * d.v[1];
    s.af = (float)d.v[0];
    s.bxf = (float)(b * t.ca);
    s.byf = (float)(b * t.sa);
    // |v| <= |d0| + 3 |a| + 3 |b| <= 4 S with S = |a|_1 + |b|_1 (fractional positions are wrapped into
    // [-0.5, 0.5)); the conversions of d0, a, b and the two chained fmas of a component add up to
    // less than 2^-20 S; E = 2^-18 S leaves a factor four
    const float S = s.af + fabsf(s.bxf) + fabsf(s.byf);
    const float E = 3.814697265625e-06f * S;
    s.wide = (float)P.lj_reach * (1.0f + 1e-6f) + 2.0f * E;
    s.wide2 = s.wide * s.wide * (1.0f + 1e-6f);
    // site pair: the two converted site positions (2^-24 R each), the sum B + D and the difference
    // (2^-24 (4 S + 2 R) each) on top of E: below 2^-17 (S + R)
    s.ec = 7.62939453125e-06f * (S + (float)P.lj_reach);
    return s;
}
// bit (ix + 3) * 7 + (iy + 3) = image (ix, iy) of three shells, the zero offset left out
// (cell.rs:216-225 with zero = false), in the order periodic_images yields them.  A row of images
// (fixed iy) whose |dy| alone exceeds the reach is skipped as a whole.
constexpr int kLjCentre = 24;
template <bool SELF>
__device__ __forceinline__ unsigned long long lj_sweep_pair(const LjSweep &s, float d0x, float d0y) {
    unsigned long long m = 0ull;
#pragma unroll
    for (int iy = SELF ? 0 : -3; iy <= 3; ++iy) {
        const float vy = fmaf((float)iy, s.byf, d0y);
        if (fabsf(vy) <= s.wide) {
            const float x0 = fmaf((float)iy, s.bxf, d0x), vy2 = vy * vy;
#pragma unroll
            for (int ix = (SELF && iy == 0) ? 1 : -3; ix <= 3; ++ix) {
                if (ix == 0 && iy == 0) continue;
                const float vx = fmaf((float)ix, s.af, x0);
                const float d2 = fmaf(vx, vx, vy2);
                if (d2 <= s.wide2) m |= 1ull << ((ix + 3) * 7 + (iy + 3));
            }
        }
    }
    if (SELF) m |= __brevll(m) >> 15;  // image -t of a shape around itself is as far as image t
    return m;
}

// the single-precision pre-filter of one candidate image: bit k * n + l set when site k of shape i
// may be within its cutoff of site l of image qb of shape j; g = the OWNER's column
template <int NI, class Geo>
__device__ __forceinline__ unsigned lj_prefilter(const KProblem &P, Geo g, int N, int i, int j, int qb) {
    constexpr int NN = NI > 0 ? NI : kLjQueueSites;
    const int n = NI > 0 ? NI : P.n_items;
    const LjImage Ii = lj_image(P, g, n, i), Ij = lj_image(P, g, n, j);
    const int ci = (Ii.e0 - g.comp0) >> 1, cj = (Ij.e0 - g.comp0) >> 1;  // first stored site of either image
    const float2 p0 = *g.fpar(0), p1 = *g.fpar(1);  // (af, bxf), (byf, ec)
    const float d0x = (float)(g.ld(4 + j * 4 + 2) - g.ld(4 + i * 4 + 2));
    const float d0y = (float)(g.ld(4 + j * 4 + 3) - g.ld(4 + i * 4 + 3));
    const int cix = (qb * 37) >> 8;
    const float fix = (float)(cix - 3), fiy = (float)(qb - cix * 7 - 3);
    // Evaluated in shape i's frame with its row signs undone (a sign flip of one coordinate of
    // everything leaves every squared distance as it is, bit for bit): A's sites as stored, B's with
    // the product of the two sign pairs, the displacement with i's.
    const float Dx = lj_flipf(fmaf(fix, p0.x, fmaf(fiy, p0.y, d0x)), Ii.sx), Dy = lj_flipf(fmaf(fiy, p1.x, d0y), Ii.sy);
    const unsigned bsx = Ii.sx ^ Ij.sx, bsy = Ii.sy ^ Ij.sy;
    float bx[NN], by[NN];
#pragma unroll
    for (int l = 0; l < NN; ++l) {
        const float2 B = g.ldf(cj + (l < n ? l : 0));
        bx[l] = lj_flipf(B.x, bsx) + Dx;
        by[l] = lj_flipf(B.y, bsy) + Dy;
    }
    unsigned m = 0u;
#pragma unroll
    for (int k = 0; k < NN; ++k) {
        // true |d| < cutoff implies |d_f| < cutoff + 1.5 ec and r2_f <= (cutoff + 2 ec)^2
        const double cut = P.items[k < n ? k : 0][4];
        const float c = (float)cut * (1.0f + 1e-6f) + 2.0f * p1.y;
        const float lim = cut == cut ? c * c * (1.0f + 1e-6f) : 3.0e38f;
        const float2 A = g.ldf(ci + (k < n ? k : 0));
#pragma unroll
        for (int l = 0; l < NN; ++l) {
            const float dx = A.x - bx[l], dy = A.y - by[l];
            const float r2 = fmaf(dx, dx, dy * dy);
            if ((NI > 0 || (k < n && l < n)) && r2 <= lim) m |= 1u << (k * n + l);
        }
    }
    return m;
}

// (i << 2) | j of in-cell pair p: pairs in the order (0,1),(0,2),..,(N-2,N-1) (potential.rs:88-96)
__device__ __forceinline__ int lj_pair_of(int p, int N) {
    const unsigned lut = N == 4 ? 0xB76321u : (N == 3 ? 0x621u : 0x1u);
    return (int)((lut >> (4 * p)) & 15u);
}
// the image mask of the ordered pair (i, j): stored for i < j, mirrored for i > j, one for all i == j
template <class Geo>
__device__ __forceinline__ unsigned long long lj_image_mask(Geo g, int N, int i, int j) {
    const int PW = N * (N - 1) / 2;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const unsigned long long m = *g.mask(i == j ? PW : lo * N - lo * (lo + 1) / 2 + (hi - lo - 1));
    return i > j ? __brevll(m) >> 15 : m;  // bit q <-> bit 48 - q
}

// each lane that wants to contribute `want` entries gets its share of the 32 pool slots; returns how
// many it may write (take), where (slot) and the number of entries of the whole warp (total)
__device__ __forceinline__ void lj_pool_quota(int want, int &take, int &slot, int &total) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wanting = __ballot_sync(0xffffffffu, want > 0);
    const int nw = __popc(wanting);
    const int quota = (int)(__fdividef(32.0f, (float)(nw ? nw : 1)) + 0.01f);  // 32 / nw (exact where it is an integer)
    const int rank = __popc(wanting & ((1u << lane) - 1u));
    const int share = quota + (rank < 32 - quota * nw ? 1 : 0);
    take = want < share ? want : share;
    int incl = take;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += up;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    slot = incl - take;
}

// Warp-collective: all 32 lanes call it converged; lanes without a state to score pass live = false
// and work on the others' entries.
template <int NI>
__device__ __forceinline__ double t_lj_score_pooled(const KProblem &P, TGeo g, const Dof &d, const Trig &t, bool live) {
    const unsigned lane = threadIdx.x & 31u;
    const int N = P.n_syms, n = NI > 0 ? NI : P.n_items;
    const int PW = N * (N - 1) / 2;
    unsigned *const pool = g.pool_entries();
    double *const E = g.pool_energies();
    // 1. image masks.  The in-cell pairs i < j (potential.rs:88-96) pass the same reach test: beyond it
    // every site pair is outside its cutoff and the shape-pair energy is exactly +0.0.
    int cands = 0;          // candidates this lane still has to hand out
    unsigned incell = 0u;   // in-cell pairs within reach, bit = pair index
    unsigned nonempty = 0u; // ordered pairs (i, j) with an image in reach, bit = i * N + j
    if (live) {
        const LjSweep S = lj_sweep_setup(P, d, t);
        *g.fpar(0) = make_float2(S.af, S.bxf);
        *g.fpar(1) = make_float2(S.byf, S.ec);
        int m = 0;
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j, ++m) {
                const float d0x = (float)(g.ld(4 + j * 4 + 2) - g.ld(4 + i * 4 + 2));
                const float d0y = (float)(g.ld(4 + j * 4 + 3) - g.ld(4 + i * 4 + 3));
                const unsigned long long bits = lj_sweep_pair<false>(S, d0x, d0y);
                *g.mask(m) = bits;
                cands += 2 * __popcll(bits);  // (i, j) and its mirror (j, i)
                if (bits) nonempty |= (1u << (i * N + j)) | (1u << (j * N + i));
                if (fmaf(d0x, d0x, d0y * d0y) <= S.wide2) {
                    incell |= 1u << m;
                    ++cands;
                }
            }
        const unsigned long long self = lj_sweep_pair<true>(S, 0.0f, 0.0f);
        *g.mask(PW) = self;
        cands += N * __popcll(self);
        if (self)
            for (int i = 0; i < N; ++i) nonempty |= 1u << (i * N + i);
    }
    int i = 0, j = 0;
    unsigned long long bits = 0ull;
    double sum = 0.;
    __syncwarp();
    for (;;) {
        int take, slot, entries;
        lj_pool_quota(cands, take, slot, entries);
        if (entries == 0) break;
        // 2. this lane's next `take` candidates, in sequence order
        for (int c = 0; c < take; ++c) {
            int qb;
            if (incell) {
                const int p = __ffs((int)incell) - 1;
                incell &= incell - 1u;
                const int ij = lj_pair_of(p, N);
                i = ij >> 2;
                j = ij & 3;
                qb = kLjCentre;
            } else {
                if (bits == 0ull) {
                    const int w = __ffs((int)nonempty) - 1;
                    nonempty &= nonempty - 1u;
                    i = N == 4 ? w >> 2 : (N == 2 ? w >> 1 : (N == 1 ? w : w / N));
                    j = w - i * N;
                    bits = lj_image_mask(g, N, i, j);
                }
                qb = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
            }
            pool[slot + c] = (lane << 16) | (unsigned)(i << 12) | (unsigned)(j << 10) | (unsigned)(qb << 4);
        }
        cands -= take;
        __syncwarp();
        if ((int)lane < entries) {  // one entry per lane: the energy of that shape pair
            const unsigned en = pool[lane];
            const TGeo G = g.of(en >> 16);
            const int ei = (en >> 12) & 3, ej = (en >> 10) & 3, qb = (en >> 4) & 63;
            unsigned m = lj_prefilter<NI>(P, G, N, ei, ej, qb);
            double e = 0.;
            if (m) {
                // the displacement of the image (to_cartesian_translate, cell.rs:241-245), once per entry;
                // the zero offset of an in-cell pair gives shape j's own position bit for bit (x + 0.0 == x)
                const int cix = (qb * 37) >> 8;  // qb / 7 for qb < 49
                const double dix = (double)(cix - 3), diy = (double)(qb - cix * 7 - 3);
                const double fx = G.ld(4 + ej * 4 + 0) + dix, fy = G.ld(4 + ej * 4 + 1) + diy;
                const double yb = fy * G.ld(1);
                const LjImage Ii = lj_image(P, G, n, ei), Ij = lj_image(P, G, n, ej);
                // in shape i's frame with its row signs undone: the differences below come out with
                // i's signs, exactly (negation commutes with rounding), and LJ2::energy squares them
                const double X = lj_flip(fx * G.ld(0) + yb * G.ld(2), Ii.sx), Y = lj_flip(yb * G.ld(3), Ii.sy);
                const double cxi = lj_flip(G.ld(4 + ei * 4 + 2), Ii.sx), cyi = lj_flip(G.ld(4 + ei * 4 + 3), Ii.sy);
                const unsigned bsx = Ii.sx ^ Ij.sx, bsy = Ii.sy ^ Ij.sy;
                const int bi = Ii.e0, bj = Ij.e0;
                do {  // the surviving site pairs in iproduct order (lj_shape.rs:35-39), LJ2::energy in f64
                    const int p = __ffs((int)m) - 1;
                    m &= m - 1;
                    const int k = (p * (n == 3 ? 86 : (n == 4 ? 64 : (n == 2 ? 128 : 256)))) >> 8;  // p / n
                    const int l = p - k * n;
                    const double sx = G.ld(bi + k * 2 + 0) + cxi, sy = G.ld(bi + k * 2 + 1) + cyi;
                    const double ox = lj_flip(G.ld(bj + l * 2 + 0), bsx) + X, oy = lj_flip(G.ld(bj + l * 2 + 1), bsy) + Y;
                    const double *pk = P.items[k];
                    e += lj_pair(sx - ox, sy - oy, pk[2], pk[3], pk[4], P.lj_shift[k]);
                } while (m);
            }
            E[lane] = e;
        }
        __syncwarp();
        // 3. the owner's sum, one shape pair after the other
        for (int c = 0; c < take; ++c) sum += E[slot + c];
        __syncwarp();
    }
    return -sum / (double)N;
}

// PotentialState::score (state/potential.rs:82-115) for one replica in one thread.  Warp-collective:
// called by all 32 lanes converged; the result of a lane with live = false is meaningless.
template <int NI>
__device__ __forceinline__ double t_lj_score(const KProblem &P, TGeo g, const Dof &d, const Trig &t,
                                             const CellCache &cc, bool live) {
    if (live) lj_build_geometry<NI>(P, g, d, t);
    // a site without cutoff: every image and every site pair contributes, the plain loops are dense
    const bool plain = !(P.lj_reach < 1e300) || (NI > kLjQueueSites) || (NI == 0 && P.n_items > kLjQueueSites);
    if (plain) return live ? t_lj_score_plain<NI>(P, g, d, t) : 0.;
    if constexpr (NI > 0 && NI <= kLjQueueSites)
        return t_lj_score_pooled<NI>(P, g, d, t, live);
    else if constexpr (NI == 0)
        return t_lj_score_pooled<0>(P, g, d, t, live);
    else
        return 0.;
}

#endif  // CP_THREAD_LJ_CUH
