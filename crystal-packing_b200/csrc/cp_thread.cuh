// cp_thread.cuh — one THREAD per replica (lanes_per_replica = 1), the default kernels.
//
// Every lane owns a whole replica: RNG, schedule, DOFs and the Metropolis rule are scalar work that
// a warp per replica would replicate 32 times (profiles/r1_a_group32_first.txt).  What SIMT needs in
// return is warp-uniform control flow around the expensive parts, which the move loop provides:
//   * every lane walks its own (stage, outer loop, inner step) schedule and draws proposals in a
//     short divergent loop until it holds one that needs scoring: out-of-bounds proposals cost a
//     draw, and for hard shapes a cell move that the Metropolis rule rejects even if valid is
//     rejected there too, before any overlap test (the packing fraction depends on the cell only);
//   * phase A, per lane: positions of the symmetry images in f64 (reference expressions), the
//     rotated shape points in f32 (of image 0 only where the others are its row-sign flips, see
//     TLayout), and the single-precision sweep over the periodic images of HALF of the ordered shape
//     pairs - a proven superset of the images the reference's f64 filter keeps (the reference tests
//     (i, j, t) and its mirror image (j, i, -t); one certificate covers both, cp_filter.h);
//     candidates become bits of one 64-bit word per pair;
//   * phase B, whole warp: the candidates of all 32 replicas are drawn in rounds into a shared pool
//     of 32 entries, one entry per lane.  A lane evaluates the f32 CERTIFICATE of its entry
//     (cp_filter.h): "every component test is false" / "some component test is true", proven about
//     the reference's f64 result.  Component pairs it cannot certify go to a warp-wide queue and are
//     evaluated with the reference's own f64 expressions (both mirror sides, each behind the
//     reference's own distance filter), one pair per lane.  Replicas drop out at their first hit;
//   * phase C, per lane: the decision.
// A warp holds up to 32 replicas; how many is chosen at the launch (KParams::warp_fill,
// cp_thread_inst.cuh).
// So every decision is the reference's: certificates only ever replace f64 evaluations whose result
// they prove.  Arithmetic of everything that reaches a decision: reference expression order, f64,
// no contraction (-fmad=false).
#ifndef CP_THREAD_CUH
#define CP_THREAD_CUH

#include <type_traits>

#include "cp_device.cuh"
#include "cp_filter.h"
#include "cp_thread_lj.cuh"

constexpr int kThreadBlock = 32;
constexpr int kQueueCap = 96;  // component pairs waiting for their f64 evaluation, per warp

// A shape-pair candidate of one replica: shape i against image (ix, iy) of shape j, |ix|, |iy| <= 3.
// Bit q = (iy + 3) * 7 + (ix + 3) of the 64-bit word of the pair; the centre bit (ix = iy = 0)
// stands for the in-cell pair i < j, whose displacement to_cartesian(f_j + (0, 0)) is bit for bit
// j's own Cartesian position (x + 0.0 == x), so one routine serves both kinds.
constexpr int kCentreBit = 24;
// the replica's state between moves in the f64 plane, relative to dof0()
enum { kStUpper = 6, kStTrig = 8, kStSched = 12, kStElems = 18 };
// slow schedule variables, relative to dof0() + kStSched (raw bits for the integers)
enum { kSchedScoreStart = 0, kSchedStepRatio, kSchedRejections, kSchedMoves, kSchedLoopsLeft, kSchedStage };
// decodes a candidate bit into its image offsets
__device__ __forceinline__ void image_of_bit(int q, int &ix, int &iy) {
    const int row = (q * 37) >> 8;  // q / 7 for q < 49
    iy = row - 3;
    ix = q - row * 7 - 3;
}

// ---------------------------------------------------------------------------------- shared memory
// One block = one warp.  Planes are [element][lane] so that a lane's own accesses are conflict-free
// and any lane can read any replica's column (cooperative tests).
//   f64 plane : 0 a, 1 b, 2 cos, 3 sin of the cell angle, 4 cos, 5 sin of the site angle, then per
//               symmetry image s: 6 + 4s + {0 fx, 1 fy (wrapped fractional), 2 cx, 3 cy (Cartesian)},
//               then (at dof0()) the replica's state between moves, so that the registers hold only
//               what every draw needs: + 0..5 the six DOFs of the current (accepted) state (the draw
//               loop reads "the DOF the RNG picked" with one indexed load), + kStUpper.. the upper
//               bounds of length and ratio for this stage (the other four are constants), + kStTrig..
//               sin, cos of the site angle and of the cell angle of that state, + kStSched.. the
//               schedule's slow variables (kSched*)
//   f32 plane : per point p: x, y of the rotated point (relative to the image's centre) - of image 0
//               only when every operator's rotation part is that of operator 0 up to the signs of
//               its rows (KProblem::signed_images: all seven plane groups of the reference; the
//               other images' points are those with their signs flipped, bit for bit), else of
//               every image; last element: the squared distance below which an image is a candidate
//               for certain
//   masks     : N (N - 1) / 2 pair words + 1 self word (u64); in registers for multiplicities 1, 2
//   constants : the shape components, the symmetry operators, the lower bounds of the DOFs and the
//               upper bounds that do not depend on the stage (f64)
template <int KIND, int NI, int NS>
struct TLayout {
    // NF: images whose rotated points are stored.  Multiplicities 1 and 2 keep all of theirs (their
    // footprint does not bound the occupancy, and the certificates save the sign multiplications).
    int n, N, NP, NF;
    __host__ __device__ TLayout(int n_items, int n_syms, int signed_images)
        : n(NI > 0 ? NI : n_items), N(NS > 0 ? NS : n_syms), NP((KIND == CP_POLYGON && NI == 0) ? 2 * n : n),
          NF((signed_images && NS != 1 && NS != 2) ? 1 : N) {}
    __host__ __device__ TLayout(const KProblem &P) : TLayout(P.n_items, P.n_syms, P.signed_images) {}
    __host__ __device__ int dof0() const { return 6 + 4 * N; }   // element of DOF 0 (kSt* follow)
    __host__ __device__ int f64_elems() const { return 6 + 4 * N + kStElems; }
    __host__ __device__ int f32_elems() const { return 2 * NF * NP; }
    __host__ __device__ int pairs() const { return N * (N - 1) / 2; }
    __host__ __device__ int const_doubles() const { return n * 4 + N * 6 + 12; }
    __host__ __device__ size_t bytes() const {
        return 8 * (size_t)(32 * f64_elems() + 32 * (pairs() + 1) + const_doubles()) + 4 * (size_t)(32 * (f32_elems() + 1)) +
               4 * (size_t)(72 + kQueueCap);
    }
};

struct TView {
    double *f64;                // plane base (lane 0)
    unsigned long long *masks;  // plane base
    double *c_items, *c_syms, *c_lo;  // c_lo[0..5] lower bounds, c_lo[6..11] stage-independent upper bounds
    float *f32;       // plane base
    unsigned *pool;   // [0..31] entries, [32..63] hit flag of each lane's replica, [64] queue fill
    unsigned *queue;  // kQueueCap items
};

template <int KIND, int NI, int NS>
__device__ __forceinline__ TView thread_view(double *smem, const KProblem &P) {
    const TLayout<KIND, NI, NS> L(P);
    TView v;
    v.f64 = smem;
    v.masks = reinterpret_cast<unsigned long long *>(smem + 32 * L.f64_elems());
    v.c_items = smem + 32 * (L.f64_elems() + L.pairs() + 1);
    v.c_syms = v.c_items + L.n * 4;
    v.c_lo = v.c_syms + L.N * 6;
    v.f32 = reinterpret_cast<float *>(v.c_lo + 12);
    v.pool = reinterpret_cast<unsigned *>(v.f32 + 32 * (L.f32_elems() + 1));
    v.queue = v.pool + 72;
    return v;
}

// block constants for the f64 evaluations (any lane, any component index)
template <int KIND, int NI, int NS>
__device__ __forceinline__ void thread_stage_constants(const TView &v, const KProblem &P) {
    const TLayout<KIND, NI, NS> L(P);
    for (int e = threadIdx.x; e < L.n * 4; e += kThreadBlock) v.c_items[e] = P.items[e >> 2][e & 3];
    for (int e = threadIdx.x; e < L.N * 6; e += kThreadBlock) v.c_syms[e] = P.syms[e / 6][e % 6];
    if (threadIdx.x < 6) {  // lower bounds per slot (cell.rs:139-160, site.rs:69-88)
        const int e = threadIdx.x;
        v.c_lo[e] = e == 0 ? 0.01 : (e == 1 ? 0.1 : (e == 2 ? CP_PI / 4. : (e == 5 ? 0. : -0.5)));
        // upper bounds; those of length and ratio are captured per stage and live in the lane's column
        v.c_lo[6 + e] = e == 2 ? CP_PI / 2. : (e == 5 ? CP_TAU : 0.5);
    }
    __syncwarp();
}

// the rotated points of image s of replica `lane` as the certificates read them
struct TPts {
    const float *p;  // element 0 of the image, lane column
    const float *rad;
    __device__ __forceinline__ float x(int k) const { return p[(2 * k) * 32]; }
    __device__ __forceinline__ float y(int k) const { return p[(2 * k + 1) * 32]; }
    __device__ __forceinline__ float r(int k) const { return rad[k]; }
};
// the same with the signs of the image's rows applied (sx, sy = +-1: exact)
struct TPtsSigned {
    const float *p;
    const float *rad;
    float sx, sy;
    __device__ __forceinline__ float x(int k) const { return sx * p[(2 * k) * 32]; }
    __device__ __forceinline__ float y(int k) const { return sy * p[(2 * k + 1) * 32]; }
    __device__ __forceinline__ float r(int k) const { return rad[k]; }
};

// One symmetry image of the occupied site: the translation column of sym * Transform2::new(theta,
// (x, y)) wrapped into [-0.5, 0.5) (site.rs:40-45, transform.rs:107-112) and its
// Cartesian position (cell.rs:110-112,258-263).  a = cell length, b = a * ratio.
__device__ __forceinline__ void image_position(const double *S, const Dof &d, const Trig &t, double a, double b,
                                               double &fx, double &fy, double &cx, double &cy) {
    const double tx = ((S[0] * d.v[3]) + S[1] * d.v[4]) + S[2] * 1.0;
    const double ty = ((S[3] * d.v[3]) + S[4] * d.v[4]) + S[5] * 1.0;
    fx = wrap_unit(tx);
    fy = wrap_unit(ty);
    const double yb = fy * b;
    cx = fx * a + yb * t.ca;
    cy = yb * t.sa;
}
// rotation block of the same product, rows 0-1 (nalgebra's Matrix3 product order)
__device__ __forceinline__ void image_rotation(const double *S, double ct, double st, double &m00, double &m01,
                                               double &m10, double &m11) {
    m00 = ((S[0] * ct) + S[1] * st) + S[2] * 0.0;
    m01 = ((S[0] * -st) + S[1] * ct) + S[2] * 0.0;
    m10 = ((S[3] * ct) + S[4] * st) + S[5] * 0.0;
    m11 = ((S[3] * -st) + S[4] * ct) + S[5] * 0.0;
}

// ---------------------------------------------------------------------------------- phase A
// OccupiedSite::positions (site.rs:33-45) and Cell2::to_cartesian_isometry (cell.rs:110-112,258-263)
// in f64 with the reference's expressions; the rotated shape points in f32 for the certificates.
template <int KIND, int NI, int NS>
__device__ __forceinline__ void t_place(const KProblem &P, const TView &v, const Dof &d, const Trig &t) {
    const TLayout<KIND, NI, NS> L(P);
    double *g = v.f64 + threadIdx.x;
    float *h = v.f32 + threadIdx.x;
    const double a = d.v[0], b = d.v[0] * d.v[1];
    g[0 * 32] = a;
    g[1 * 32] = b;
    g[2 * 32] = t.ca;
    g[3 * 32] = t.sa;
    g[4 * 32] = t.ct;
    g[5 * 32] = t.st;
    const float ctf = (float)t.ct, stf = (float)t.st;
#pragma unroll
    for (int s = 0; s < (NS > 0 ? NS : CP_MAX_SYMS); ++s) {
        if (NS == 0 && s >= L.N) break;
        double fx, fy, cx, cy;
        image_position(P.syms[s], d, t, a, b, fx, fy, cx, cy);
        g[(6 + 4 * s + 0) * 32] = fx;
        g[(6 + 4 * s + 1) * 32] = fy;
        g[(6 + 4 * s + 2) * 32] = cx;
        g[(6 + 4 * s + 3) * 32] = cy;
        // rotation block of the same product, single precision (certificates only)
        if (s >= L.NF) continue;  // signed images: the points of image 0 serve every image
        const float *F = P.symsf[s];
        const float m00 = fmaf(F[0], ctf, F[1] * stf), m01 = fmaf(F[1], ctf, -(F[0] * stf));
        const float m10 = fmaf(F[2], ctf, F[3] * stf), m11 = fmaf(F[3], ctf, -(F[2] * stf));
        if (NI > 0) {
#pragma unroll
            for (int p = 0; p < (NI > 0 ? NI : 1); ++p) {
                const float px = P.ptsf[p][0], py = P.ptsf[p][1];
                h[((s * NI + p) * 2 + 0) * 32] = fmaf(m00, px, m01 * py);
                h[((s * NI + p) * 2 + 1) * 32] = fmaf(m10, px, m11 * py);
            }
        } else {
            for (int p = 0; p < L.NP; ++p) {
                const float px = P.ptsf[p][0], py = P.ptsf[p][1];
                h[((s * L.NP + p) * 2 + 0) * 32] = fmaf(m00, px, m01 * py);
                h[((s * L.NP + p) * 2 + 1) * 32] = fmaf(m10, px, m11 * py);
            }
        }
    }
}

// candidates of one lane: the masks (registers for multiplicities 1 and 2, shared memory otherwise)
// and the cursor that draws them
template <int NS>
struct TCands {
    static constexpr bool kRegs = NS == 1 || NS == 2;
    unsigned long long r0 = 0, r1 = 0;  // NS = 1: r0 = self word; NS = 2: r0 = pair (0, 1), r1 = self word
    unsigned long long *sh = nullptr;   // this lane's column of the mask plane
    unsigned long long cur = 0;         // bits of the current word still to draw
    int phase = -1;                     // words: 0 .. P - 1 the pairs i < j, P .. P + N - 1 the self word once per shape
    int remaining = 0;
    int first = -1;                     // (word << 6) | bit of the candidate to draw first; -1: none
    unsigned long long skip = 0;        // bit of `first` when it is a self image: shape 0's pass over the self word leaves it out
    __device__ __forceinline__ void st(int w, unsigned long long x) {
        if constexpr (kRegs) {
            if (w == 0) r0 = x; else r1 = x;
        } else {
            sh[w * 32] = x;
        }
    }
    __device__ __forceinline__ unsigned long long ld(int w) const {
        if constexpr (kRegs) return w == 0 ? r0 : r1;
        else return sh[w * 32];
    }
};

// (i << 2) | j of pair word w: pairs in the order (0,1),(0,2),..,(N-2,N-1)
__device__ __forceinline__ int pair_of_word(int w, int N) {
    const unsigned lut = N == 4 ? 0xB76321u : (N == 3 ? 0x621u : 0x1u);
    return (int)((lut >> (4 * w)) & 15u);
}

// The image sweep (cp_filter.h: cpf_sweep_setup, cpf_sweep_pair) on this lane's proposal.
__device__ __forceinline__ CpfSweep t_sweep_setup(const KProblem &P, const Dof &d, const Trig &t, const CellCache &cc) {
    const double b = d.v[0] * d.v[1];
    return cpf_sweep_setup((float)d.v[0], (float)(b * t.ca), (float)(b * t.sa), P.radf, cc.shells);
}

// Candidate generation for PackedState::check_intersection (state/packed.rs:127-167), half of the
// ordered pairs: every in-cell pair i < j (no distance filter there, packed.rs:134-148), every
// image of j against i for i < j, and the upper half of the self images (once for all shapes: the
// displacement is the lattice vector whichever the shape).
template <int KIND, int NI, int NS>
__device__ __forceinline__ void t_sweep(const KProblem &P, const TView &v, const Dof &d, const Trig &t,
                                        const CellCache &cc, TCands<NS> &c) {
    const TLayout<KIND, NI, NS> L(P);
    const double *g = v.f64 + threadIdx.x;
    const CpfSweep s = t_sweep_setup(P, d, t, cc);
    v.f32[L.f32_elems() * 32 + threadIdx.x] = s.narrow2;
    const int PW = L.pairs();
    int count = 0;
    // The candidate nearest to its partner is drawn first: when the proposal overlaps at all it
    // almost always overlaps there (first hit at rank 1.1 on average instead of 2.3 for a pentagon
    // in p2, 1.05 instead of 6.5 in p2gg), and a replica leaves the pool at its first hit.
    // Worth its instructions where a replica has many candidates (multiplicity 4: ten on average);
    // with two shapes per cell the plain order is as fast.
    constexpr bool kNearestFirst = NS == 0 || NS > 2;
    if constexpr (kNearestFirst) {
        // six pair words and the self word, a handful of images in reach out of 48 each: loops over
        // the rows and columns in reach (cpf_sweep_pair_reach returns the straight-line sweep's mask)
        float best = 3.0e38f;
        int first = -1;
#pragma unroll 1
        for (int w = 0; w < PW; ++w) {
            const int ij = pair_of_word(w, L.N), i = ij >> 2, j = ij & 3;
            const float d0x = (float)(g[(6 + 4 * j + 2) * 32] - g[(6 + 4 * i + 2) * 32]);
            const float d0y = (float)(g[(6 + 4 * j + 3) * 32] - g[(6 + 4 * i + 3) * 32]);
            int q = -1;
            const unsigned long long bits = cpf_sweep_pair_reach<false>(s, d0x, d0y, best, q) | (1ull << kCentreBit);
            // the in-cell pair competes with its own distance (the zero offset)
            const float dc = fmaf(d0x, d0x, d0y * d0y);
            if (dc < best) {
                best = dc;
                q = kCentreBit;
            }
            if (q >= 0) first = (w << 6) | q;
            c.st(w, bits);
            count += __popcll(bits);
        }
        int q = -1;
        const unsigned long long self = cpf_sweep_pair_reach<true>(s, 0.0f, 0.0f, best, q);
        if (q >= 0) first = (PW << 6) | q;
        c.st(PW, self);
        count += __popcll(self) * L.N;
        // the nearest candidate is drawn first and must not be drawn again: a pair word gives its bit
        // up here; the self word keeps it (shapes 1 .. N - 1 need it) and shape 0's pass skips it
        c.first = first;
        c.skip = 0ull;
        if (first >= 0) {
            const int w = first >> 6;
            const unsigned long long bit = 1ull << (first & 63);
            if (w < PW) c.st(w, c.ld(w) & ~bit);
            else c.skip = bit;
        }
    } else {
        // a rolled loop over the pair words: a single copy of the 48-offset sweep in the kernel
#pragma unroll 1
        for (int w = 0; w < PW; ++w) {
            const int ij = pair_of_word(w, L.N), i = ij >> 2, j = ij & 3;
            const float d0x = (float)(g[(6 + 4 * j + 2) * 32] - g[(6 + 4 * i + 2) * 32]);
            const float d0y = (float)(g[(6 + 4 * j + 3) * 32] - g[(6 + 4 * i + 3) * 32]);
            const unsigned long long bits = cpf_sweep_pair<false>(s, d0x, d0y) | (1ull << kCentreBit);
            c.st(w, bits);
            count += __popcll(bits);
        }
        const unsigned long long self = cpf_sweep_pair<true>(s, 0.0f, 0.0f);
        c.st(PW, self);
        count += __popcll(self) * L.N;
        c.first = -1;
        c.skip = 0ull;
    }
    c.cur = 0;
    c.phase = -1;
    c.remaining = count;
}

// draws the next candidate of this lane (remaining > 0 guarantees there is one): the nearest one
// first, then the words in order
template <int NS>
__device__ __forceinline__ void t_pop(TCands<NS> &c, int P, int N, int &ij, int &q) {
    constexpr bool kNearestFirst = NS == 0 || NS > 2;
    if (kNearestFirst && c.first >= 0) {
        const int w = c.first >> 6;
        q = c.first & 63;
        ij = w < P ? pair_of_word(w, N) : 0;  // a self image: the one of shape 0
        c.first = -1;
        return;
    }
    while (c.cur == 0ull) {
        ++c.phase;
        c.cur = c.ld(c.phase < P ? c.phase : P);
        if (kNearestFirst && c.phase == P) c.cur &= ~c.skip;
    }
    q = __ffsll((long long)c.cur) - 1;
    c.cur &= c.cur - 1;
    ij = c.phase < P ? pair_of_word(c.phase, N) : (c.phase - P) * 5;
}

// ---------------------------------------------------------------------------------- f64 evaluation
// One component pair of one candidate with the reference's expressions, both mirror sides:
//   side A : component k of shape i in its cell placement (self) against component l of image
//            (ix, iy) of shape j (other)           -- packed.rs:134-148 (in-cell) / :152-165
//   side A': component l of shape j (self) against component k of image (-ix, -iy) of shape i
// each behind the reference's own distance filter (packed.rs:156-157; none for the in-cell pair,
// which has no mirror).  Reads the owner's column of the f64 plane, so any lane can evaluate it.
// item: owner << 18 | i << 16 | j << 14 | q << 8 | l << 4 | k
template <int KIND, int NI, int NS>
__device__ __noinline__ bool t_exact_item(const KProblem &P, const TView &v, unsigned item) {
    const unsigned owner = item >> 18;
    const int i = (item >> 16) & 3, j = (item >> 14) & 3, q = (item >> 8) & 63, l = (item >> 4) & 15, k = item & 15;
    const double *o = v.f64 + owner;
    const double a = o[0 * 32], b = o[1 * 32], ca = o[2 * 32], sa = o[3 * 32], ct = o[4 * 32], st = o[5 * 32];
    const double fxi = o[(6 + 4 * i + 0) * 32], fyi = o[(6 + 4 * i + 1) * 32], cxi = o[(6 + 4 * i + 2) * 32],
                 cyi = o[(6 + 4 * i + 3) * 32];
    const double fxj = o[(6 + 4 * j + 0) * 32], fyj = o[(6 + 4 * j + 1) * 32], cxj = o[(6 + 4 * j + 2) * 32],
                 cyj = o[(6 + 4 * j + 3) * 32];
    int ix, iy;
    image_of_bit(q, ix, iy);
    const bool incell = q == kCentreBit && i != j;
    const double r2x = P.radius * 2.;
    const double radius_sq = r2x * r2x;
    // rotation blocks of sym * Transform2::new(theta, ..), rows 0-1 (site.rs:40-45)
    const double *Si = v.c_syms + i * 6, *Sj = v.c_syms + j * 6;
    double i00, i01, i10, i11, j00, j01, j10, j11;
    image_rotation(Si, ct, st, i00, i01, i10, i11);
    image_rotation(Sj, ct, st, j00, j01, j10, j11);
    const double *ik = v.c_items + k * 4, *jl = v.c_items + l * 4;
    // Transform * Point without its translation (line_shape.rs:103-108)
    const double ksx = i00 * ik[0] + i01 * ik[1], ksy = i10 * ik[0] + i11 * ik[1];
    const double lsx = j00 * jl[0] + j01 * jl[1], lsy = j10 * jl[0] + j11 * jl[1];
    double kex = 0., key = 0., lex = 0., ley = 0.;
    if (KIND == CP_POLYGON) {
        kex = i00 * ik[2] + i01 * ik[3];
        key = i10 * ik[2] + i11 * ik[3];
        lex = j00 * jl[2] + j01 * jl[3];
        ley = j10 * jl[2] + j11 * jl[3];
    }
    bool hit = false;
    {  // side A
        const double fx = fxj + (double)ix, fy = fyj + (double)iy;  // to_cartesian_translate (cell.rs:241-245)
        const double yb = fy * b;
        const double X = fx * a + yb * ca, Y = yb * sa;
        const double ddx = cxi - X, ddy = cyi - Y;
        const bool cand = incell || (ddx * ddx + ddy * ddy) <= radius_sq;  // packed.rs:156-157
        const double sx = ksx + cxi, sy = ksy + cyi, ox = lsx + X, oy = lsy + Y;
        bool t;
        if (KIND == CP_POLYGON) {
            const double sdx = (kex + cxi) - sx, sdy = (key + cyi) - sy;  // Line2::dx / dy (line2.rs:94-101)
            const double odx = (lex + X) - ox, ody = (ley + Y) - oy;
            t = segments_cross(sx, sy, sdx, sdy, ox, oy, odx, ody);
        } else {
            const double rs = ik[2] + jl[2];  // Atom2::intersects (atom2.rs:21-26)
            const double dx = sx - ox, dy = sy - oy;
            t = (dx * dx + dy * dy) < rs * rs;
        }
        hit = cand & t;
    }
    if (!incell) {  // side A'
        const double fx = fxi + (double)(-ix), fy = fyi + (double)(-iy);
        const double yb = fy * b;
        const double X = fx * a + yb * ca, Y = yb * sa;
        const double ddx = cxj - X, ddy = cyj - Y;
        const bool cand = (ddx * ddx + ddy * ddy) <= radius_sq;
        const double sx = lsx + cxj, sy = lsy + cyj, ox = ksx + X, oy = ksy + Y;
        bool t;
        if (KIND == CP_POLYGON) {
            const double sdx = (lex + cxj) - sx, sdy = (ley + cyj) - sy;
            const double odx = (kex + X) - ox, ody = (key + Y) - oy;
            t = segments_cross(sx, sy, sdx, sdy, ox, oy, odx, ody);
        } else {
            const double rs = jl[2] + ik[2];
            const double dx = sx - ox, dy = sy - oy;
            t = (dx * dx + dy * dy) < rs * rs;
        }
        hit |= cand & t;
    }
    return hit;
}

// evaluates the queued component pairs, one per lane; owners that already have a hit are skipped
template <int KIND, int NI, int NS>
__device__ __noinline__ void t_flush(const KProblem &P, const TView &v, int qcount) {
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp();
#pragma unroll 1
    for (int base = 0; base < qcount; base += 32) {
        const int idx = base + (int)lane;
        if (idx < qcount) {
            const unsigned item = v.queue[idx];
            const unsigned owner = item >> 18;
            volatile unsigned *flag = v.pool + 32 + owner;
            if (!*flag)
                if (t_exact_item<KIND, NI, NS>(P, v, item)) *flag = 1u;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------- phase B
// PackedState::check_intersection for the 32 replicas of the warp at once (see the header comment).
// Must be called by all 32 lanes converged; lanes without a proposal pass remaining = 0.
template <int KIND, int NI, int NS>
__device__ __forceinline__ bool t_overlap_warp(const KProblem &P, const TView &v, TCands<NS> &c) {
    const TLayout<KIND, NI, NS> L(P);
    constexpr int W = NI > 0 ? cpf_words(NI) : CPF_MAX_WORDS;
    const unsigned lane = threadIdx.x & 31u;
    const int NP2 = L.NP * 2, PW = L.pairs(), nn = L.n * L.n;
    int qcount = 0;  // warp-uniform
    v.pool[32 + lane] = 0u;  // hit flag of this lane's replica: any lane may raise it (plain stores of 1)
    if (lane == 0) v.pool[64] = 0u;
    __syncwarp();
    for (;;) {
        const unsigned wanting = __ballot_sync(0xffffffffu, c.remaining > 0);
        if (wanting == 0u) break;
        // every wanting lane contributes 32 / (lanes wanting) candidates, the first 32 mod (lanes wanting)
        // of them one more: the pool holds at most 32 entries and is full whenever the lanes have
        // that many left (the quotient of floats is exact where it is an integer and at least 1/32
        // away from one otherwise)
        const int nw = __popc(wanting);
        const int quota = (int)(__fdividef(32.0f, (float)nw) + 0.01f);
        const int rank = __popc(wanting & ((1u << lane) - 1u));
        const int share = quota + (rank < 32 - quota * nw ? 1 : 0);
        const int take = c.remaining < share ? c.remaining : share;
        int incl = take;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int slot = incl - take;
        for (int k = 0; k < take; ++k) {
            int ij, q;
            t_pop<NS>(c, PW, L.N, ij, q);
            v.pool[slot++] = (lane << 11) | (unsigned)(ij << 6) | (unsigned)q;
        }
        c.remaining -= take;
        __syncwarp();
        // one entry per lane: the certificate
        unsigned amb[W];
        int count = 0;
        unsigned entry = 0u;
        if ((int)lane < total) {
            entry = v.pool[lane];
            const unsigned owner = entry >> 11;
            const int i = (entry >> 8) & 3, j = (entry >> 6) & 3, q = entry & 63;
            const double *o = v.f64 + owner;
            int ix, iy;
            image_of_bit(q, ix, iy);
            // image position of shape j relative to shape i: to_cartesian_translate (cell.rs:241-245)
            const double a = o[0 * 32], b = o[1 * 32], ca = o[2 * 32], sa = o[3 * 32];
            const double fx = o[(6 + 4 * j + 0) * 32] + (double)ix, fy = o[(6 + 4 * j + 1) * 32] + (double)iy;
            const double yb = fy * b;
            float dX = (float)((fx * a + yb * ca) - o[(6 + 4 * i + 2) * 32]);
            float dY = (float)(yb * sa - o[(6 + 4 * i + 3) * 32]);
            // Signed images: the stored points are image 0's; image s has them with the signs
            // (sgnf[s]) of its rows, exactly.  The certificate is evaluated in shape i's frame with its
            // signs undone: A's points as stored, B's with the product of the two sign pairs, delta
            // with i's.  Flipping one coordinate of everything multiplies every cross product by -1
            // exactly and leaves the squared distances as they are, and the certificate depends on the
            // H pair and the G pair only through min |.| and the XOR of their signs (cp_filter.h):
            // cmin, tau and the ambiguous set are bit for bit those of the flipped evaluation.
            constexpr bool kAllStored = NS == 1 || NS == 2;
            const int iA = L.NF == 1 ? 0 : i, iB = L.NF == 1 ? 0 : j;
            const float sAx = kAllStored ? 1.0f : P.sgnf[i][0], sAy = kAllStored ? 1.0f : P.sgnf[i][1];
            const TPts A{v.f32 + owner + (iA * NP2) * 32, P.discrf};
            using PtsB = typename std::conditional<kAllStored, TPts, TPtsSigned>::type;
            PtsB B;
            if constexpr (kAllStored) {
                B = TPts{v.f32 + owner + (iB * NP2) * 32, P.discrf};
            } else {
                const bool one = L.NF == 1;  // else every image is stored with its own signs
                B = TPtsSigned{v.f32 + owner + (iB * NP2) * 32, P.discrf, one ? sAx * P.sgnf[j][0] : 1.0f,
                               one ? sAy * P.sgnf[j][1] : 1.0f};
                if (one) {
                    dX *= sAx;
                    dY *= sAy;
                }
            }
            // an image within rounding of the reference's 2R filter (or beyond it): candidacy itself is
            // in doubt, so no certificate is used and the f64 path applies the reference's own filter
            const bool force = q != kCentreBit && !(fmaf(dX, dX, dY * dY) <= v.f32[L.f32_elems() * 32 + owner]);
            float cmin, tau;
            if constexpr (KIND == CP_POLYGON) {
                tau = cpf_tau(P.radf, dX, dY);
                if constexpr (NI > 0)
                    cmin = cpf_chain<NI>(A, B, dX, dY, tau, amb);
                else
                    cmin = cpf_segments(L.n, A, B, dX, dY, tau, amb);
            } else {
                tau = cpf_tau_discs(P.radf, dX, dY);
                cmin = cpf_discs<NI>(L.n, A, B, dX, dY, tau, amb);
            }
            if (force) {  // candidacy itself is in doubt: every component pair goes through the f64 path
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const int left = nn - 32 * w;
                    amb[w] = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
                }
                count = nn;
            } else if (cmin < -tau) {
                v.pool[32 + owner] = 1u;  // certified: some component test is true
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w) count += __popc(amb[w]);
            }
        }
        // component pairs without a certificate: queue them for the f64 evaluation
        const int fresh = __reduce_add_sync(0xffffffffu, count);
        if (fresh > 0) {
            if (qcount + fresh > kQueueCap) {
                t_flush<KIND, NI, NS>(P, v, qcount);
                qcount = 0;
                if (lane == 0) v.pool[64] = 0u;
                __syncwarp();
            }
            const unsigned head = (entry >> 11) << 18 | ((entry >> 6) & 15u) << 14 | (entry & 63u) << 8;
            if (fresh > kQueueCap) {
                // degenerate amounts (aligned mirror images at theta = 0): each lane evaluates its own
                bool h = false;
                if (count > 0) {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        unsigned bits = amb[w];
#pragma unroll 1
                        while (bits && !h) {
                            const int p = 32 * w + __ffs((int)bits) - 1;
                            bits &= bits - 1;
                            const int l = p / L.n, k = p - l * L.n;
                            h = t_exact_item<KIND, NI, NS>(P, v, head | (unsigned)(l << 4) | (unsigned)k);
                        }
                    }
                    if (h) v.pool[32 + (entry >> 11)] = 1u;
                }
            } else {
                if (count > 0) {
                    int at = (int)atomicAdd(&v.pool[64], (unsigned)count);  // few lanes have any: no scan
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        unsigned bits = amb[w];
#pragma unroll 1
                        while (bits) {
                            const int p = 32 * w + __ffs((int)bits) - 1;
                            bits &= bits - 1;
                            const int l = p / L.n, k = p - l * L.n;
                            v.queue[at++] = head | (unsigned)(l << 4) | (unsigned)k;
                        }
                    }
                }
                qcount += fresh;
            }
        }
        __syncwarp();
        if (v.pool[32 + lane]) c.remaining = 0;  // decided: the rest of this replica's candidates need no test
        __syncwarp();                                    // the pool is rewritten by the next round
    }
    if (qcount > 0) t_flush<KIND, NI, NS>(P, v, qcount);
    return v.pool[32 + lane] != 0u;
}

// State::score for the warp's 32 replicas (batch score, replay); warp-collective.  LJ: lane-local.
template <int KIND, int NI, int NS>
__device__ __forceinline__ double t_state_score(const KProblem &P, const TView &v, double *smem, const Dof &d,
                                                const Trig &t, bool live) {
    if constexpr (KIND == CP_LJ) {
        const TGeo g = lj_geo(smem, P);
        return t_lj_score<NI>(P, g, d, t, cell_cache_of(d), live);
    } else {
        TCands<NS> c;
        c.sh = v.masks + threadIdx.x;
        if (live) {
            t_place<KIND, NI, NS>(P, v, d, t);
            t_sweep<KIND, NI, NS>(P, v, d, t, cell_cache_of(d), c);
        }
        const bool overlap = t_overlap_warp<KIND, NI, NS>(P, v, c);
        const double cell_area = t.sa * d.v[0] * (d.v[0] * d.v[1]);
        return overlap ? __longlong_as_double(0x7FF8000000000000ll) : (P.area * (double)P.n_syms) / cell_area;
    }
}

// shared memory of a block
template <int KIND, int NI, int NS>
__host__ __device__ inline size_t thread_block_shared_bytes(const KProblem &P) {
    if (KIND == CP_LJ) return lj_block_bytes(P);
    return TLayout<KIND, NI, NS>(P).bytes();
}

// Monte-Carlo optimisation of the LJ potential, one replica per thread: MCOptimiser::optimise_state
// (optimisation.rs:80-180) for every stage, seeds and bounds per stage as analyse_state /
// generate_basis set them.  The replica's state lives in registers; the warp meets in t_lj_score.
// Register budget = minimum resident blocks per SM (hard shapes: 14 warps per SM hold all 2,048
// warps of a 65,536-replica batch at once; the budgets per component count and multiplicity are the
// measured optima of the C5 sweep, tools/c5_sweep.py).
constexpr int thread_min_blocks(int kind, int ni, int ns = 0) {
    return kind == CP_LJ ? 14 : (ni == 0 ? 8 : (ni <= 6 ? 14 : (ns == 4 ? (ni <= 9 ? 14 : (ni <= 11 ? 12 : 8)) : (ni <= 11 ? 10 : 8))));
}
// proposals a lane may draw per pass before the warp goes on to score what it has: the draw loop
// is divergent (lanes leave it as they find a proposal worth scoring), so its tail is cut
constexpr int kDrawCap = 3;  // default; KParams::draw_cap

template <int KIND, int NI, int NS, int RNG>
__global__ void __launch_bounds__(kThreadBlock, thread_min_blocks(KIND, NI))
mc_thread_lj_kernel(const __grid_constant__ KParams K) {
    static_assert(KIND == CP_LJ, "LJ only: hard shapes run mc_thread_kernel");
    extern __shared__ double smem[];
    const KProblem &P = K.prob;
    const unsigned long long rep = (unsigned long long)blockIdx.x * K.warp_fill + threadIdx.x;
    const bool active = (int)threadIdx.x < K.warp_fill && rep < K.n_replicas;  // idle lanes stay in the warp-wide votes, doing nothing
    const unsigned long long replica = K.replica_begin + (active ? rep : 0ull);

    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i)
        d.v[i] = (K.init_dof && active) ? K.init_dof[rep * CP_MAX_DOF + i] : P.dof[i];
    Trig t;
    {
        const SinCos th = dev_sincos(d.v[5]), an = dev_sincos(d.v[2]);
        t.st = th.s; t.ct = th.c; t.sa = an.s; t.ca = an.c;
    }
    const int n_cell = n_cell_dof(P.family);
    const int n_basis = n_cell + 3;
    unsigned long long rejections = 0, moves = 0;
    double score_current = 0.;

    Rng rng;
    CellCache cc = cell_cache_of(d);
    // Every lane walks its own (stage, outer loop, inner step) schedule: replicas of one warp need
    // different numbers of scored proposals per inner loop (the in-bounds fraction depends on the
    // replica's cell), and synchronising at loop ends would leave the faster lanes idle there.  Only
    // the score phases below are warp-wide.
    int stage = -1;
    bool done = !active;
    double kt = 0., step_ratio = 1.0, score_start = 0.;
    double len0 = d.v[0], ratio0 = d.v[1];  // LJ keeps the captured bounds in registers
    double step_scale = 0.;  // max_step_size * step_ratio, the loop-invariant factor of optimisation.rs:115-119
    int convergence_count = 0;
    unsigned long long inner = 0, loops_left = 0, it = 0, loop_rejections = 0;
    auto begin_stage = [&]() {
        const KStage &S = K.stages[stage];
        // generate_basis(): bounds captured per optimise_state call (cell.rs:139-152)
        len0 = d.v[0];
        ratio0 = d.v[1];
        rng_begin_stage<RNG>(rng, K.seed_base, replica, n_basis);
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
    // The first pass scores the initial state through the same code as every proposal
    // (optimisation.rs:81-84: an invalid initial state is an error); `first` is warp-uniform.
    bool first = true;
    for (;;) {
        bool have = false;
        int slot = 6;  // 6 = no DOF changes (the first pass)
        double val = 0., new_val = 0., new_score = 0.;
        if (first) {
            have = true;  // tail lanes score the template state too: their result is never written
        } else {
            // Draw until a proposal needs scoring, at most K.draw_cap draws per pass.  Out-of-bounds
            // proposals are rejected by Basis::set_value without a score (optimisation.rs:120-123).
            for (int draws = 0; draws < K.draw_cap && !done && !have; ++draws) {
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
                double lo, hi;
                val = d.get(slot);
                slot_bounds(slot, len0, ratio0, lo, hi);
                new_val = val + step_scale * sample;
                have = lo <= new_val && new_val <= hi;
                if (!have) ++loop_rejections;
            }
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (__all_sync(0xffffffffu, done)) break;
            continue;  // every lane used up its draws without a proposal to score
        }
        // the proposed state (lanes holding a proposal)
        double old_s = 0., old_c = 0.;
        const CellCache old_cc = cc;
        if (have) {
            d.set(slot, new_val);
            if (slot == 2 || slot == 5) {  // the cell angle and the site angle are the only DOFs inside sin/cos
                const SinCos sc = dev_sincos(new_val);
                if (slot == 5) {
                    old_s = t.st; old_c = t.ct;
                    t.st = sc.s; t.ct = sc.c;
                } else {
                    old_s = t.sa; old_c = t.ca;
                    t.sa = sc.s; t.ca = sc.c;
                }
            }
            if (slot <= 2) cc = cell_cache_of(d);
        }
        {  // PotentialState::score, whole warp: lanes without a proposal work on the others' entries
            const TGeo g = lj_geo(smem, P);
            const double s = t_lj_score<NI>(P, g, d, t, cc, have);
            if (have) new_score = s;
        }
        // the Metropolis decision
        if (first) {
            score_current = new_score;
            const bool valid = score_current == score_current;
            if (active && !valid) atomicExch(K.error_flag, 1);
            if (!valid) done = true;
            if (!done) next_stage();
            first = false;
        } else if (have) {
            const unsigned long long threshold = rng_draw_threshold_bits<RNG>(rng);  // raw bits of draw #3
            if (accept_move_fast(new_score, score_current, kt, threshold)) {
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

// ---------------------------------------------------------------------------------- hard shapes
// Registers of the hard-shape kernel between moves: what every draw needs.  Everything else lives in
// the lane's column of the f64 plane (TLayout) and is touched at inner-loop ends only.
struct THot {
    double score_current, kt, step_scale;
    unsigned long long left;             // trial moves left in this inner loop
    unsigned long long loop_rejections;  // optimisation.rs:100,122,134
    Rng rng;
    bool done;
};

// End of an inner loop (optimisation.rs:138-167), and the start of the next stage when the loop was
// the stage's last or the stage converged (analyse_state, main.rs:224-252).  Runs once per
// inner_steps trial moves, so it is kept out of line and works on the plane.
template <int RNG>
__device__ __noinline__ void t_loop_end(const KParams &K, double *mine, int e_dof, unsigned long long replica,
                                        int n_basis, THot &h) {
    double *sch = mine + (e_dof + kStSched) * 32;
    int packed = (int)__double_as_longlong(sch[kSchedStage * 32]);
    int stage = (packed & 0xff) - 1, convergence_count = packed >> 8;  // low byte: stage + 1 (0: none opened yet)
    unsigned long long loops_left = (unsigned long long)__double_as_longlong(sch[kSchedLoopsLeft * 32]);
    if (stage >= 0 && stage < K.n_stages && loops_left > 0) {  // not the call that opens the first stage
        const KStage &S = K.stages[stage];
        sch[kSchedMoves * 32] = __longlong_as_double(__double_as_longlong(sch[kSchedMoves * 32]) + (long long)S.inner);
        sch[kSchedRejections * 32] =
            __longlong_as_double(__double_as_longlong(sch[kSchedRejections * 32]) + (long long)h.loop_rejections);
        h.kt *= S.kt_ratio;
        bool converged = false;
        if (S.convergence == S.convergence) {
            if (h.score_current - sch[kSchedScoreStart * 32] < S.convergence) {
                if (++convergence_count > 5) converged = true;
            } else {
                convergence_count = 0;
            }
        }
        double step_ratio = sch[kSchedStepRatio * 32];
        if (!converged && step_ratio > 1e-4) step_ratio *= (double)S.inner / ((double)h.loop_rejections + 1.);
        sch[kSchedStepRatio * 32] = step_ratio;
        h.step_scale = S.max_step * step_ratio;
        --loops_left;
        h.left = S.inner;
        h.loop_rejections = 0;
        sch[kSchedScoreStart * 32] = h.score_current;
        if (converged) loops_left = 0;
    }
    while (loops_left == 0) {  // next stage; stages that run no move are skipped
        ++stage;
        if (stage >= K.n_stages) {
            h.done = true;
            break;
        }
        const KStage &S = K.stages[stage];
        // generate_basis(): bounds captured per optimise_state call (cell.rs:139-152)
        mine[(e_dof + kStUpper + 0) * 32] = mine[(e_dof + 0) * 32];
        mine[(e_dof + kStUpper + 1) * 32] = mine[(e_dof + 1) * 32];
        rng_begin_stage<RNG>(h.rng, K.seed_base, replica, n_basis);
        h.kt = S.kt_start;
        sch[kSchedStepRatio * 32] = 1.0;
        h.step_scale = S.max_step * 1.0;
        convergence_count = 0;
        h.left = S.inner;
        loops_left = S.inner ? S.steps / S.inner : 0ull;
        h.loop_rejections = 0;
        sch[kSchedScoreStart * 32] = h.score_current;
    }
    sch[kSchedLoopsLeft * 32] = __longlong_as_double((long long)loops_left);
    sch[kSchedStage * 32] = __longlong_as_double((long long)((stage + 1) | (convergence_count << 8)));
}

// Monte-Carlo optimisation of hard shapes, one replica per thread (see the header comment for the
// control flow): MCOptimiser::optimise_state (optimisation.rs:80-180) for every stage, seeds and
// bounds per stage as analyse_state / generate_basis set them.
// Register budget = minimum resident blocks per SM: 14 warps per SM hold all 2,048 warps of a
// 65,536-replica batch at once.
template <int KIND, int NI, int NS, int RNG>
__global__ void __launch_bounds__(kThreadBlock, thread_min_blocks(KIND, NI, NS))
mc_thread_kernel(const __grid_constant__ KParams K) {
    static_assert(KIND != CP_LJ, "hard shapes");
    extern __shared__ double smem[];
    const KProblem &P = K.prob;
    const TView v = thread_view<KIND, NI, NS>(smem, P);
    thread_stage_constants<KIND, NI, NS>(v, P);
    const int e_dof = TLayout<KIND, NI, NS>(P).dof0();
    double *const mine = v.f64 + threadIdx.x;  // this lane's column of the f64 plane
    const unsigned long long rep = (unsigned long long)blockIdx.x * K.warp_fill + threadIdx.x;
    const bool active = (int)threadIdx.x < K.warp_fill && rep < K.n_replicas;  // idle lanes stay in the warp-wide votes, doing nothing
    const unsigned long long replica = K.replica_begin + (active ? rep : 0ull);
    const int n_cell = n_cell_dof(P.family);
    const int n_basis = n_cell + 3;
    const double area_n = P.area * (double)P.n_syms;  // numerator of packed.rs:84

    // the replica's state between moves, in the plane
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i)
        mine[(e_dof + i) * 32] = (K.init_dof && active) ? K.init_dof[rep * CP_MAX_DOF + i] : P.dof[i];
    {
        const SinCos th = dev_sincos(mine[(e_dof + 5) * 32]), an = dev_sincos(mine[(e_dof + 2) * 32]);
        mine[(e_dof + kStTrig + 0) * 32] = th.s;
        mine[(e_dof + kStTrig + 1) * 32] = th.c;
        mine[(e_dof + kStTrig + 2) * 32] = an.s;
        mine[(e_dof + kStTrig + 3) * 32] = an.c;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) mine[(e_dof + kStSched + i) * 32] = 0.;  // counters 0, no stage opened yet

    THot h;
    h.score_current = 0.;
    h.kt = 0.;
    h.step_scale = 0.;
    h.left = 0;
    h.loop_rejections = 0;
    h.done = !active;
    TCands<NS> cands;
    cands.sh = v.masks + threadIdx.x;
    // Every lane walks its own (stage, outer loop, inner step) schedule: replicas of one warp need
    // different numbers of scored proposals per inner loop, and synchronising at loop ends would
    // leave the faster lanes idle there.  Only the score phases below are warp-wide.
    // The first pass scores the initial state through the same code as every proposal
    // (optimisation.rs:81-84: an invalid initial state is an error); `first` is warp-uniform.
    bool first = true;
    for (;;) {
        bool have = false;
        int slot = 6;  // 6 = no DOF changes (the first pass)
        double new_val = 0., new_score = 0.;
        unsigned long long threshold = 0ull;  // raw bits of draw #3
        if (first) {
            have = true;  // tail lanes score the template state too: their result is never written
        } else {
            // Draw until a proposal needs scoring, at most K.draw_cap draws per pass.  Out-of-bounds
            // proposals are rejected by Basis::set_value without a score (optimisation.rs:120-123).
            // The packing fraction depends on the cell DOFs only, so `Some(new_score)` is known
            // before any overlap test: when the Metropolis rule would reject even a valid state (a
            // cell move that lowers the score, lost against the threshold) the reference's answer is
            // "rejected" whether check_intersection says None or Some, and the overlap test is
            // skipped.  Draw order (index, step, threshold) and every decision are unchanged.
            for (int draws = 0; draws < K.draw_cap && !h.done && !have; ++draws) {
                if (h.left == 0) {
                    THot at_loop_end = h;  // a copy, so that `h` itself never leaves the registers
                    t_loop_end<RNG>(K, mine, e_dof, replica, n_basis, at_loop_end);
                    h = at_loop_end;
                    continue;
                }
                --h.left;
                int basis_index;
                double sample;
                rng_draw_move<RNG>(h.rng, basis_index, sample);
                slot = basis_to_slot(basis_index, n_cell);
                const double val = mine[(e_dof + slot) * 32];
                const double *const hp = slot < 2 ? mine + (e_dof + kStUpper + slot) * 32 : v.c_lo + 6 + slot;
                const double hi = *hp, lo = v.c_lo[slot];
                new_val = val + h.step_scale * sample;
                have = lo <= new_val && new_val <= hi;
                if (!have) {
                    ++h.loop_rejections;
                    continue;
                }
                threshold = rng_draw_threshold_bits<RNG>(h.rng);
                new_score = h.score_current;  // site moves leave the packing fraction as it is
                if (slot <= 2) {
                    double sa_new = mine[(e_dof + kStTrig + 2) * 32];
                    bool settled = false;
                    if (slot == 2) {
                        // Cell-angle move.  A single-precision sine (absolute error < 2^-21 on this
                        // range; margin 4e-6) settles the common outcomes without the f64 sine: the
                        // area shrinks by a relative 2e-6 or more, so new > old and the move is
                        // accepted whatever the temperature (the exact sine and score are taken in
                        // phase A); or it grows that much while kT is +0 (stages 1 and 3), where
                        // exp(-inf) = 0 rejects, or while kT is so small that the exponent
                        // underflows.  Everything else takes the exact path.
                        const float s_old = (float)sa_new, s_new = __sinf((float)new_val);
                        if (s_new < s_old * (1.0f - 4.0e-6f)) {
                            settled = true;
                        } else if (s_new > s_old * (1.0f + 4.0e-6f) && __double_as_longlong(h.kt) == 0ll) {
                            settled = true;
                            have = false;
                        } else if (s_new > s_old * (1.0f + 4.0e-6f) && h.kt > 0.0 &&
                                   (float)h.score_current * ((s_new - s_old) / s_new - 2.0e-6f) > 746.75f * (float)h.kt) {
                            // new / old = sin(old) / sin(new) exactly as real numbers, so the score falls by
                            // at least old * rel with rel = (s_new - s_old) / s_new - 2e-6 (the f32 sines
                            // are within 6e-7 of the exact ones).  When that is more than 746.75 kT the
                            // exponent is below -745.14, where cp_exp returns 0: rejected.
                            settled = true;
                            have = false;
                        } else {
                            sa_new = dev_sincos(new_val).s;
                        }
                    }
                    if (!settled) {
                        const double d0 = mine[(e_dof + 0) * 32], d1 = mine[(e_dof + 1) * 32];
                        const double len = slot == 0 ? new_val : d0, ratio = slot == 1 ? new_val : d1;
                        const double cell_area = sa_new * len * (len * ratio);  // cell.rs:190-192
                        new_score = area_n / cell_area;                         // packed.rs:84
                        have = accept_move_fast(new_score, h.score_current, h.kt, threshold);
                    }
                    if (!have) ++h.loop_rejections;
                }
            }
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (__all_sync(0xffffffffu, h.done)) break;
            continue;  // every lane used up its draws without a proposal to score
        }
        // phase A (lanes holding a proposal): the proposed state, its geometry and its candidates
        cands.remaining = 0;
        double new_s = 0., new_c = 0.;
        if (have) {
            Dof d;
#pragma unroll
            for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = mine[(e_dof + i) * 32];
            d.set(slot, new_val);
            Trig t;
            t.st = mine[(e_dof + kStTrig + 0) * 32];
            t.ct = mine[(e_dof + kStTrig + 1) * 32];
            t.sa = mine[(e_dof + kStTrig + 2) * 32];
            t.ca = mine[(e_dof + kStTrig + 3) * 32];
            if (slot == 2 || slot == 5) {  // the cell angle and the site angle are the only DOFs inside sin/cos
                const SinCos sc = dev_sincos(new_val);
                new_s = sc.s;
                new_c = sc.c;
                if (slot == 5) {
                    t.st = sc.s;
                    t.ct = sc.c;
                } else {
                    t.sa = sc.s;
                    t.ca = sc.c;
                    new_score = area_n / (t.sa * d.v[0] * (d.v[0] * d.v[1]));  // cell.rs:190-192, packed.rs:84
                }
            }
            if (first) new_score = area_n / (t.sa * d.v[0] * (d.v[0] * d.v[1]));
            t_place<KIND, NI, NS>(P, v, d, t);
            t_sweep<KIND, NI, NS>(P, v, d, t, cell_cache_of(d), cands);
        }
        // phase B (whole warp): shape-pair tests, pooled over the warp's replicas
        const bool overlap = t_overlap_warp<KIND, NI, NS>(P, v, cands);
        // phase C: the decision (the Metropolis rule itself was evaluated before the overlap test)
        if (first) {
            h.score_current = overlap ? __longlong_as_double(0x7FF8000000000000ll) : new_score;
            if (overlap) {
                if (active) atomicExch(K.error_flag, 1);
                h.done = true;
            }
            if (!h.done) {  // opens the first stage
                THot at_loop_end = h;
                t_loop_end<RNG>(K, mine, e_dof, replica, n_basis, at_loop_end);
                h = at_loop_end;
            }
            first = false;
        } else if (have) {
            if (!overlap) {
                h.score_current = new_score;
                mine[(e_dof + slot) * 32] = new_val;
                if (slot == 5) {
                    mine[(e_dof + kStTrig + 0) * 32] = new_s;
                    mine[(e_dof + kStTrig + 1) * 32] = new_c;
                }
                if (slot == 2) {
                    mine[(e_dof + kStTrig + 2) * 32] = new_s;
                    mine[(e_dof + kStTrig + 3) * 32] = new_c;
                }
            } else {
                ++h.loop_rejections;
            }
        }
    }
    if (active) {
        cp_replica_result r;
#pragma unroll
        for (int i = 0; i < CP_MAX_DOF; ++i) r.dof[i] = mine[(e_dof + i) * 32];
        r.score = h.score_current;
        r.rejections = (unsigned long long)__double_as_longlong(mine[(e_dof + kStSched + kSchedRejections) * 32]);
        r.moves = (unsigned long long)__double_as_longlong(mine[(e_dof + kStSched + kSchedMoves) * 32]);
        K.out[rep] = r;
    }
}

template <int KIND, int NI, int NS>
__global__ void __launch_bounds__(kThreadBlock)
score_thread_kernel(const __grid_constant__ KProblem P, const double *dof, unsigned long long n, double *out) {
    extern __shared__ double smem[];
    TView v{};
    if constexpr (KIND != CP_LJ) {
        v = thread_view<KIND, NI, NS>(smem, P);
        thread_stage_constants<KIND, NI, NS>(v, P);
    }
    const unsigned long long rep = (unsigned long long)blockIdx.x * kThreadBlock + threadIdx.x;
    const bool active = rep < n;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = active ? dof[rep * CP_MAX_DOF + i] : P.dof[i];
    const double s = t_state_score<KIND, NI, NS>(P, v, smem, d, trig_of(d), active);
    if (active) out[rep] = s;
}

template <int KIND, int NI, int NS>
__global__ void __launch_bounds__(kThreadBlock)
replay_thread_kernel(const __grid_constant__ KProblem P, const double *init_dof, const cp_move *mv,
                     unsigned long long n_moves, unsigned long long n, cp_decision *out) {
    extern __shared__ double smem[];
    TView v{};
    if constexpr (KIND != CP_LJ) {
        v = thread_view<KIND, NI, NS>(smem, P);
        thread_stage_constants<KIND, NI, NS>(v, P);
    }
    const unsigned long long rep = (unsigned long long)blockIdx.x * kThreadBlock + threadIdx.x;
    const bool active = rep < n;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = active ? init_dof[rep * CP_MAX_DOF + i] : P.dof[i];
    Trig t;
    const int n_cell = n_cell_dof(P.family);
    const double *bd = P.replay_has_bounds ? init_dof + n * CP_MAX_DOF : init_dof;
    const double len0 = active ? bd[rep * CP_MAX_DOF + 0] : d.v[0], ratio0 = active ? bd[rep * CP_MAX_DOF + 1] : d.v[1];
    cp_sincos(d.v[5], &t.st, &t.ct);
    cp_sincos(d.v[2], &t.sa, &t.ca);
    double score_current = t_state_score<KIND, NI, NS>(P, v, smem, d, t, active);
    for (unsigned long long m = 0; m < n_moves; ++m) {
        cp_move M{};
        if (active) M = mv[rep * n_moves + m];
        const int slot = basis_to_slot((int)M.basis_index, n_cell);
        const double val = d.get(slot);
        double lo, hi;
        slot_bounds(slot, len0, ratio0, lo, hi);
        cp_decision D;
        D.in_bounds = 0;
        D.accepted = 0;
        D.new_score = __longlong_as_double(0x7FF8000000000000ll);
        const bool in_bounds = active && lo <= M.new_value && M.new_value <= hi;
        if (in_bounds) {
            d.set(slot, M.new_value);
            cp_sincos(d.v[5], &t.st, &t.ct);
            cp_sincos(d.v[2], &t.sa, &t.ca);
        }
        // warp-collective: lanes without an in-bounds proposal take part with no candidates
        const double s = t_state_score<KIND, NI, NS>(P, v, smem, d, t, in_bounds);
        if (in_bounds) {
            D.in_bounds = 1;
            D.new_score = s;
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
        if (active) out[rep * n_moves + m] = D;
    }
}

// OccupiedSite::positions + Cell2::to_cartesian_isometry for a batch of states (site.rs:33-45,
// cell.rs:110-112): per symmetry image fx, fy (wrapped fractional position), cx, cy (Cartesian
// position) and the rotation block m00, m01, m10, m11, from the device functions the kernels use.
static __global__ void __launch_bounds__(128)
images_kernel(const __grid_constant__ KProblem P, const double *dof, unsigned long long n, double *out) {
    const unsigned long long rep = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= n) return;
    Dof d;
#pragma unroll
    for (int i = 0; i < CP_MAX_DOF; ++i) d.v[i] = dof[rep * CP_MAX_DOF + i];
    const Trig t = trig_of(d);
    const double a = d.v[0], b = d.v[0] * d.v[1];
    for (int s = 0; s < P.n_syms; ++s) {
        double *o = out + (rep * P.n_syms + s) * 8;
        image_position(P.syms[s], d, t, a, b, o[0], o[1], o[2], o[3]);
        image_rotation(P.syms[s], t.ct, t.st, o[4], o[5], o[6], o[7]);
    }
}

#endif  // CP_THREAD_CUH
