// cp_filter.h — single-precision certificates for the hard-shape overlap predicates.
//
// PackedState::check_intersection (state/packed.rs:127-167) is an OR over shape pairs of an OR over
// component pairs of Line2::intersects (shape/components/line2.rs:29-54) or Atom2::intersects
// (components/atom2.rs:21-26), all in f64.  Almost every one of those component tests is decided by
// a wide margin.  This header computes, in f32 on the FMA/ALU pipes, a CERTIFICATE for a whole
// shape pair: a number `cmin` with
//     cmin >  tau   =>  every component test of the pair, evaluated with the reference's f64
//                       expressions, is false                                  ("clear")
//     cmin < -tau   =>  at least one of them is true                           ("hit")
//     otherwise         nothing is claimed: the caller evaluates the reference expressions in f64.
// The certificate is about the reference's COMPUTED result, not about geometry: see the proofs
// below.  Shared by the kernels (cp_thread.cuh) and by the CPU soundness test (tests/filter_check.cpp),
// so both execute the same f32 operations (fmaf is one correctly rounded operation on both sides).
//
// ---- segments ------------------------------------------------------------------------------------
// For a segment s = (s0, s0 + sd) of shape A and o = (o0, o0 + od) of shape B the reference computes
//     u  = od.y*sd.x - od.x*sd.y                 = cross(sd, od)
//     ta = od.x*(s0.y-o0.y) - od.y*(s0.x-o0.x)   = cross(od, s0 - o0)
//     tb = sd.x*(s0.y-o0.y) - sd.y*(s0.x-o0.x)   = cross(sd, s0 - o0)
// and answers  u != 0  &&  0 <= ta/u <= 1  &&  0 <= tb/u <= 1.  Write, over the reals,
//     H0 = cross(od, s0 - o0),  H1 = cross(od, s0 + sd - o0)      (s's end points against o's line)
//     G0 = cross(sd, o0 - s0),  G1 = cross(sd, o0 + od - s0)      (o's end points against s's line)
// so that ta = H0, u = H0 - H1, tb = -G0, u = G1 - G0.  The f64 evaluation of ta, tb, u differs from
// these reals by at most mu = 16 * 2^-53 * |d||w| (|d| <= 2R an edge, |w| <= 2R + |delta| a
// difference of placed points), i.e. mu < 1e-14 * R * (R + V).
//  (F) If H0 and H1 have the same sign and min(|H0|,|H1|) >= m > 2 mu, the computed test is false:
//      say both >= m.  If H0 > H1 then u > 0 and ta - u = H1 >= m, so computed ta > computed |u|
//      and the quotient test fails (or computed u <= 0: zero, or a sign opposite to ta's).  If
//      H0 <= H1 then u <= 0 while ta >= m: computed u is below mu, so it is zero, of the opposite
//      sign, or smaller than ta.  Same for G0, G1 through tb.  Both signs by symmetry.
//  (T) If H0, H1 have opposite signs with min(|H0|,|H1|) >= m and so have G0, G1, the computed test
//      is true: |u| = |H0| + |H1| >= 2m, computed u keeps u's sign, ta = H0 has u's sign and
//      |ta| <= |u| - m, likewise tb.
// With sH = min(|H0|,|H1|) carrying the XOR of the signs (negative = the end points straddle the
// line) and sG likewise, c = max(sH, sG) is > m exactly when (F) applies and < -m exactly when (T)
// does; cmin = min over the component pairs.
//
// The f32 evaluation works in coordinates relative to A's centre: A's points a_k (|a_k| <= R), B's
// points b_l + delta, delta = B's centre - A's centre rounded from f64.  With eps = 2^-24, points
// carrying an absolute error <= 8 eps R (rotation evaluated in f32 from the f32 sine and cosine) and
// |delta| <= |dx| + |dy|: an edge vector (difference of two unshifted points) is within 18 eps R,
// a shifted point within 9 eps R + 2 eps |delta|, and the three roundings of a cross product add
// 8 eps R (2R + |delta|); every H and G is therefore within E = 30 eps R (3R + |delta|) of its real
// value.  tau = 2^-18 R (3R + |dx| + |dy|) = 64 eps R (3R + |delta|_1) leaves a factor two on top of
// these (generous) bounds; tests/filter_check.cpp checks the certificates against the f64
// evaluation along real trajectories.
//
// Mirror images.  The reference tests shape i against image t of shape j AND shape j against image
// -t of shape i: the same configuration translated by -t, with the roles of the two shapes (and so
// of H and G) swapped.  The certificates are symmetric under that swap and hold with the margin
// above, while the two evaluations differ only by f64 rounding of the placed coordinates
// (< 1e-14 R), so one certificate covers both.  The caller enumerates half of the ordered pairs.
//
// ---- discs ---------------------------------------------------------------------------------------
// Atom2::intersects is  dx*dx + dy*dy < (r1 + r2)^2.  c = |d|^2 - (r1 + r2)^2 in f32 is within
// 30 eps (3R + |delta|)^2 of the real value, the f64 evaluation within 1e-15 of it; tau =
// 2^-18 (3R + |delta|_1)^2; cmin = min c.
#ifndef CP_FILTER_H
#define CP_FILTER_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CPF_HD __host__ __device__ __forceinline__
#define CPF_CX __host__ __device__ constexpr
#else
#define CPF_HD static inline
#define CPF_CX constexpr
#endif

// min(|x|, |y|) with the XOR of the two signs: positive = same side, negative = straddling
CPF_HD float cpf_clear(float x, float y) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y));
    return r;
#else
    uint32_t a, b, m;
    const float mf = fminf(fabsf(x), fabsf(y));
    memcpy(&a, &x, 4);
    memcpy(&b, &y, 4);
    memcpy(&m, &mf, 4);
    m |= (a ^ b) & 0x80000000u;
    float r;
    memcpy(&r, &m, 4);
    return r;
#endif
}

// 32-bit words that hold one bit per component pair (bit l * n + k: component k of A, l of B)
CPF_CX int cpf_words(int n) { return (n * n + 31) / 32; }
#define CPF_MAX_N 16
#define CPF_MAX_WORDS 8

// the margin of one shape pair: R = enclosing radius, (dx, dy) = B's centre relative to A's
CPF_HD float cpf_tau(float R, float dx, float dy) {
    return 3.814697265625e-06f * R * fmaf(3.0f, R, fabsf(dx) + fabsf(dy));  // 2^-18 R (3R + |delta|_1)
}

// Closed chain of N points per shape (segment k runs from point k to point (k + 1) % N; the host
// checks that the shape is one).  A and B expose x(k), y(k): the rotated points relative to the
// shape's own centre.  Everything is unrolled; the H row and the previous G row live in registers.
// Returns cmin; amb[] receives one bit per component pair whose c is <= tau (not certified clear).
template <int N, class PtsA, class PtsB>
CPF_HD float cpf_chain(const PtsA &A, const PtsB &B, float dX, float dY, float tau, unsigned (&amb)[cpf_words(N)]) {
    float ax[N], ay[N], sx[N], sy[N], gk[N], G0[N], Gf[N];
#pragma unroll
    for (int w = 0; w < cpf_words(N); ++w) amb[w] = 0u;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        ax[k] = A.x(k);
        ay[k] = A.y(k);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int k1 = k + 1 == N ? 0 : k + 1;
        sx[k] = ax[k1] - ax[k];
        sy[k] = ay[k1] - ay[k];
        gk[k] = fmaf(sx[k], ay[k], -(sy[k] * ax[k]));  // cross(sd_k, a_k)
    }
    float r0x = B.x(0), r0y = B.y(0);  // B's points before the shift: edge vectors come from these
    float b0x = r0x + dX, b0y = r0y + dY;
    const float rfx = r0x, rfy = r0y;
#pragma unroll
    for (int k = 0; k < N; ++k) G0[k] = Gf[k] = fmaf(sx[k], b0y, fmaf(-sy[k], b0x, -gk[k]));  // cross(sd_k, b_0 - a_k)
    float cmin = 3.0e38f;
#pragma unroll
    for (int l = 0; l < N; ++l) {
        const bool last = l + 1 == N;
        const float r1x = last ? rfx : B.x(last ? 0 : l + 1), r1y = last ? rfy : B.y(last ? 0 : l + 1);
        const float b1x = r1x + dX, b1y = r1y + dY;
        const float dx = r1x - r0x, dy = r1y - r0y;
        const float e = fmaf(dx, b0y, -(dy * b0x));  // cross(d_l, b_l)
        float H[N];
#pragma unroll
        for (int k = 0; k < N; ++k) H[k] = fmaf(dx, ay[k], fmaf(-dy, ax[k], -e));  // cross(d_l, a_k - b_l)
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int k1 = k + 1 == N ? 0 : k + 1;
            const float G1 = last ? Gf[k] : fmaf(sx[k], b1y, fmaf(-sy[k], b1x, -gk[k]));
            const float c = fmaxf(cpf_clear(H[k], H[k1]), cpf_clear(G0[k], G1));
            cmin = fminf(cmin, c);
            if (c <= tau) amb[(l * N + k) >> 5] |= 1u << ((l * N + k) & 31);
            G0[k] = G1;
        }
        r0x = r1x;
        r0y = r1y;
        b0x = b1x;
        b0y = b1y;
    }
    return cmin;
}

// Run-time component count, and shapes that are not closed chains: segment k has its own two points
// 2k and 2k + 1.  Nothing is carried between iterations.
template <class PtsA, class PtsB>
CPF_HD float cpf_segments(int n, const PtsA &A, const PtsB &B, float dX, float dY, float tau,
                          unsigned (&amb)[CPF_MAX_WORDS]) {
    float cmin = 3.0e38f;
    for (int w = 0; w < CPF_MAX_WORDS; ++w) amb[w] = 0u;
    for (int l = 0; l < n; ++l) {
        const float r0x = B.x(2 * l), r0y = B.y(2 * l), r1x = B.x(2 * l + 1), r1y = B.y(2 * l + 1);
        const float b0x = r0x + dX, b0y = r0y + dY, b1x = r1x + dX, b1y = r1y + dY;
        const float dx = r1x - r0x, dy = r1y - r0y;
        const float e = fmaf(dx, b0y, -(dy * b0x));
        for (int k = 0; k < n; ++k) {
            const float a0x = A.x(2 * k), a0y = A.y(2 * k), a1x = A.x(2 * k + 1), a1y = A.y(2 * k + 1);
            const float sx = a1x - a0x, sy = a1y - a0y;
            const float g = fmaf(sx, a0y, -(sy * a0x));
            const float H0 = fmaf(dx, a0y, fmaf(-dy, a0x, -e)), H1 = fmaf(dx, a1y, fmaf(-dy, a1x, -e));
            const float G0 = fmaf(sx, b0y, fmaf(-sy, b0x, -g)), G1 = fmaf(sx, b1y, fmaf(-sy, b1x, -g));
            const float c = fmaxf(cpf_clear(H0, H1), cpf_clear(G0, G1));
            cmin = fminf(cmin, c);
            if (c <= tau) amb[(l * n + k) >> 5] |= 1u << ((l * n + k) & 31);
        }
    }
    return cmin;
}

// Discs: A and B expose x(k), y(k) (rotated centres) and r(k).  c = |a_k - b_l - delta|^2 - (r_k + r_l)^2
template <int N, class PtsA, class PtsB>
CPF_HD float cpf_discs(int n_rt, const PtsA &A, const PtsB &B, float dX, float dY, float tau,
                       unsigned (&amb)[N > 0 ? cpf_words(N) : CPF_MAX_WORDS]) {
    const int n = N > 0 ? N : n_rt;
    float cmin = 3.0e38f;
#pragma unroll
    for (int w = 0; w < (N > 0 ? cpf_words(N) : CPF_MAX_WORDS); ++w) amb[w] = 0u;
#pragma unroll
    for (int l = 0; l < (N > 0 ? N : CPF_MAX_N); ++l) {
        if (N == 0 && l >= n) break;
        const float bx = B.x(l) + dX, by = B.y(l) + dY, br = B.r(l);
#pragma unroll
        for (int k = 0; k < (N > 0 ? N : CPF_MAX_N); ++k) {
            if (N == 0 && k >= n) break;
            const float ex = A.x(k) - bx, ey = A.y(k) - by, rs = A.r(k) + br;
            const float c = fmaf(ex, ex, fmaf(ey, ey, -(rs * rs)));
            cmin = fminf(cmin, c);
            if (c <= tau) amb[(l * n + k) >> 5] |= 1u << ((l * n + k) & 31);
        }
    }
    return cmin;
}
// margin for discs: 2^-18 (3R + |delta|_1)^2
CPF_HD float cpf_tau_discs(float R, float dx, float dy) {
    const float w = fmaf(3.0f, R, fabsf(dx) + fabsf(dy));
    return 3.814697265625e-06f * w * w;
}

// ---- the image sweep -----------------------------------------------------------------------------
// PackedState::check_intersection keeps an image (ix, iy) of a shape when dx*dx + dy*dy <= (2R)^2 in
// f64 (packed.rs:150-157).  Here the displacement of the image is d0 + ix a + iy b from f32 copies of
// the lattice vectors a = (af, 0), b = (bxf, byf); its error is below 2^-24 (|d0|_1 + 6 (|a| + |b|)),
// so with delta = 2^-19 (|a| + |b|) / R + 2^-20 every image the reference keeps has
// d2 <= (2R)^2 (1 + delta) here: those are the candidates.  Whether a candidate is one the
// reference keeps FOR CERTAIN is decided where its certificate is evaluated, from the f64
// displacement rounded to f32, against narrow2 = (2R)^2 (1 - delta).
// Bit (iy + 3) * 7 + (ix + 3) of the result stands for image (ix, iy), |ix|, |iy| <= shells <= 3.
// All 48 offsets of three shells are evaluated in straight-line code (four instructions each, no
// branches): in the packings the optimiser converges to, the cells are skewed enough that
// periodic_range is 3 for nine moves out of ten (packed.rs:128-132), and loops over the few offsets
// in reach cost more in branch latency than they save.
// SELF: the images of a shape against itself come in mirror pairs t, -t; only iy > 0, and ix > 0 on
// the row iy = 0, are produced.
struct CpfSweep {
    float af, bxf, byf;
    float wide2, narrow2;
    int shells;
    float wide, inv_a, inv_by;  // cpf_sweep_pair_reach: sqrt(wide2) rounded up, 1 / af, 1 / byf
};
CPF_HD CpfSweep cpf_sweep_setup(float af, float bxf, float byf, float R, int shells) {
    CpfSweep s;
    s.af = af;
    s.bxf = bxf;
    s.byf = byf;
    const float r2x = 2.0f * R;
    const float delta = fmaf(1.9073486328125e-06f, (af + fabsf(bxf) + fabsf(byf)) / R, 9.5367431640625e-07f);
    s.wide2 = r2x * r2x * (1.0f + delta);
    s.narrow2 = r2x * r2x * (1.0f - delta);
    s.shells = shells;
    s.wide = sqrtf(s.wide2) * (1.0f + 1.0e-6f);
    s.inv_a = 1.0f / af;
    s.inv_by = 1.0f / byf;
    return s;
}
template <bool SELF>
CPF_HD unsigned long long cpf_sweep_pair(const CpfSweep &s, float d0x, float d0y) {
    const bool ring2 = s.shells >= 2, ring3 = s.shells >= 3;
    unsigned lo = 0u, hi = 0u;
#pragma unroll
    for (int iy = SELF ? 0 : -3; iy <= 3; ++iy) {
        const float vy = fmaf((float)iy, s.byf, d0y), x0 = fmaf((float)iy, s.bxf, d0x);
        const float vy2 = vy * vy;
        unsigned row = 0u;
#pragma unroll
        for (int ix = (SELF && iy == 0) ? 1 : -3; ix <= 3; ++ix) {
            if (ix == 0 && iy == 0) continue;  // the zero offset is not an image (cell.rs:216-225)
            const int ring = (ix < 0 ? -ix : ix) > (iy < 0 ? -iy : iy) ? (ix < 0 ? -ix : ix) : (iy < 0 ? -iy : iy);
            const float vx = fmaf((float)ix, s.af, x0);
            const float d2 = fmaf(vx, vx, vy2);
            bool in = d2 <= s.wide2;
            if (ring == 2) in = in && ring2;
            if (ring == 3) in = in && ring3;
            row |= in ? 1u << (ix + 3) : 0u;
        }
        const int sh = (iy + 3) * 7;  // row iy occupies bits sh .. sh + 6
        if (sh < 32) lo |= row << sh;
        if (sh + 7 > 32) hi |= sh >= 32 ? row << (sh - 32) : row >> (32 - sh);
    }
    return ((unsigned long long)hi << 32) | lo;
}

// The same mask from loops over the rows and columns in reach only, for multiplicity-4 groups (six
// pair words and the self word per proposal, a handful of images in reach out of 48 each).  An image
// with d2 <= wide2 has |vy| <= wide and |vx| <= wide, so its row lies in [(-wide - d0y) / byf,
// (wide - d0y) / byf] and its column in [(-wide - x0) / af, (wide - x0) / af]; the bounds are
// evaluated in single precision (relative error below 3e-7 of |d0y| / byf + wide / byf resp.
// (|x0| + wide) / af) and widened by 0.01 + 1e-6 of those magnitudes before they are rounded
// inwards and clamped to the shells, so every offset the straight-line sweep keeps is visited, with
// the straight-line sweep's own expressions: the two return the same mask
// (tests/filter_check.cpp compares them at every scored proposal).  Also reports the nearest image:
// best / best_q are updated when a kept image is strictly nearer (first in bit order among equals).
template <bool SELF>
CPF_HD unsigned long long cpf_sweep_pair_reach(const CpfSweep &s, float d0x, float d0y, float &best, int &best_q) {
    const float sh = (float)s.shells;
    const float ymag = fmaf(fabsf(d0y) + s.wide, s.inv_by * 1.0e-6f, 0.01f);
    const float ylo = fmaxf(fmaf(-s.wide - d0y, s.inv_by, -ymag), SELF ? -0.5f : -sh);
    const float yhi = fminf(fmaf(s.wide - d0y, s.inv_by, ymag), sh);
    const int iy0 = (int)ceilf(ylo), iy1 = (int)floorf(yhi);
    unsigned long long m = 0ull;
    for (int iy = iy0; iy <= iy1; ++iy) {
        const float fiy = (float)iy;
        const float vy = fmaf(fiy, s.byf, d0y), x0 = fmaf(fiy, s.bxf, d0x);
        const float vy2 = vy * vy;
        const float xmag = fmaf(fabsf(x0) + s.wide, s.inv_a * 1.0e-6f, 0.01f);
        const float xlo = fmaxf(fmaf(-s.wide - x0, s.inv_a, -xmag), (SELF && iy == 0) ? 0.5f : -sh);
        const float xhi = fminf(fmaf(s.wide - x0, s.inv_a, xmag), sh);
        const int ix1 = (int)floorf(xhi);
        for (int ix = (int)ceilf(xlo); ix <= ix1; ++ix) {
            if (ix == 0 && iy == 0) continue;  // the zero offset is not an image (cell.rs:216-225)
            const float vx = fmaf((float)ix, s.af, x0);
            const float d2 = fmaf(vx, vx, vy2);
            if (d2 <= s.wide2) {
                const int q = (iy + 3) * 7 + (ix + 3);
                m |= 1ull << q;
                if (d2 < best) {
                    best = d2;
                    best_q = q;
                }
            }
        }
    }
    return m;
}

#endif  // CP_FILTER_H
