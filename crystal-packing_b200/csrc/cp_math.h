/*
 * cp_math.h — deterministic f64 sin/cos/exp shared by the CUDA kernels and the CPU oracle.
 *
 * Why this exists: the reference calls the platform libm through Rust's f64::sin/cos/exp
 * (crystal-packing/src/transform.rs:81-87 via nalgebra Rotation2::new, cell.rs:190-192,258-263,
 * optimisation.rs:46-48).  libm results are platform dependent in the last bit (glibc selects
 * FMA / non-FMA variants at run time) and CUDA's libdevice is a third implementation, so a
 * bit-exact GPU<->CPU comparison needs ONE definition that both sides execute.  Everything here
 * is built from IEEE-754 +,-,*,/ , fma(), rint() only — each of those is correctly rounded on
 * x86-64 and on sm_100a, so host and device produce identical bits provided the compilers do not
 * contract or reassociate (host: -ffp-contract=off, device: -fmad=false).
 *
 * Accuracy (tests/test_oracle_known_answers.py::test_math_modes_agree_closely, against glibc):
 * within 1 ulp, and equal to glibc for the overwhelming majority of arguments in the DOF ranges
 * ([0, 2pi], [pi/4, pi/2]).
 *
 * The polynomial coefficients and the split constants of pi/2 and ln2 are the numeric constants of
 * Sun's fdlibm (k_sin.c, k_cos.c, e_exp.c, e_rem_pio2.c); the code arrangement is ours.  fdlibm's
 * notice, which covers those constants:
 *
 *   Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
 *   Developed at SunSoft, a Sun Microsystems, Inc. business.
 *   Permission to use, copy, modify, and distribute this software is freely granted,
 *   provided that this notice is preserved.
 */
#ifndef CP_MATH_H
#define CP_MATH_H

#include <stdint.h>
#ifndef __CUDACC__
#include <math.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define CP_HD __host__ __device__ __forceinline__
#else
#define CP_HD static inline
#endif

CP_HD double cp_bits_to_double(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
#endif
}

/* sin and cos of x for |x| < 2^20 (the DOFs that reach here are bounded by 2*pi). */
CP_HD void cp_sincos(double x, double *sn, double *cs)
{
    /* pi/2 split in 33+33+53 bits so that fn*PIO2_1 and fn*PIO2_2 are exact for |fn| < 2^20 */
    const double INVPIO2 = 6.36619772367581382433e-01; /* 0x3FE45F306DC9C883 */
    const double PIO2_1 = 1.57079632673412561417e+00;  /* 0x3FF921FB54400000 */
    const double PIO2_2 = 6.07710050630396597660e-11;  /* 0x3DD0B4611A600000 */
    const double PIO2_2T = 2.02226624879595063154e-21; /* 0x3BA3198A2E037073 */

    const double S1 = -1.66666666666666324348e-01;
    const double S2 = 8.33333333332248946124e-03;
    const double S3 = -1.98412698298579493134e-04;
    const double S4 = 2.75573137070700676789e-06;
    const double S5 = -2.50507602534068634195e-08;
    const double S6 = 1.58969099521155010221e-10;

    const double C1 = 4.16666666666666019037e-02;
    const double C2 = -1.38888888888741095749e-03;
    const double C3 = 2.48015872894767294178e-05;
    const double C4 = -2.75573143513906633035e-07;
    const double C5 = 2.08757232129817482790e-09;
    const double C6 = -1.13596475577881948265e-11;

    double fn = rint(x * INVPIO2);
    int n = (int)fn;

    /* two-level Cody-Waite reduction: (y0, y1) = x - fn*pi/2 as an unevaluated hi+lo pair */
    double t = x - fn * PIO2_1;
    double w = fn * PIO2_2;
    double r = t - w;
    w = fn * PIO2_2T - ((t - r) - w);
    double y0 = r - w;
    double y1 = (r - y0) - w;

    double z = y0 * y0;

    /* sin(y0 + y1) on [-pi/4, pi/4] */
    double v = z * y0;
    double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    double sv = y0 - ((z * (0.5 * y1 - v * ps) - y1) - v * S1);

    /* cos(y0 + y1) on [-pi/4, pi/4] */
    double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double hz = 0.5 * z;
    double wc = 1.0 - hz;
    double cv = wc + (((1.0 - wc) - hz) + (z * pc - y0 * y1));

    switch (n & 3) {
    case 0:
        *sn = sv;
        *cs = cv;
        break;
    case 1:
        *sn = cv;
        *cs = -sv;
        break;
    case 2:
        *sn = -sv;
        *cs = -cv;
        break;
    default:
        *sn = -cv;
        *cs = sv;
        break;
    }
}

/* exp(x) for any x (the hot path only needs x <= 0, NaN and -inf; see optimisation.rs:46-48). */
CP_HD double cp_exp(double x)
{
    const double LN2HI = 6.93147180369123816490e-01; /* 0x3FE62E42FEE00000 */
    const double LN2LO = 1.90821492927058770002e-10; /* 0x3DEA39EF35793C76 */
    const double INVLN2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;

    if (x != x)
        return x;
    if (x > 7.09782712893383973096e+02)
        return cp_bits_to_double(0x7FF0000000000000ull);
    if (x < -7.45133219101941108420e+02)
        return 0.0;

    double fk = rint(x * INVLN2);
    int k = (int)fk;
    double hi = x - fk * LN2HI; /* fk*LN2HI exact: LN2HI has 32 significant bits, |fk| < 2^11 */
    double lo = fk * LN2LO;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);

    if (k >= -1021) {
        return y * cp_bits_to_double((uint64_t)(1023 + k) << 52);
    }
    /* subnormal result: scale in two steps so the single rounding happens in the last product */
    y = y * cp_bits_to_double((uint64_t)(1023 + k + 1000) << 52);
    return y * cp_bits_to_double((uint64_t)(1023 - 1000) << 52);
}

#endif /* CP_MATH_H */
