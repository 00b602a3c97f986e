/*
 * cp_oracle.c — CPU oracle for the crystal-packing Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see cp_oracle.h).  Plain-C restatement of the reference algorithm;
 * every function cites the reference file:line (relative to /root/reference) it follows.
 * PARITY STATUS: "parity unpinned" at trajectory level (the Rust reference cannot be built here);
 * pinned against every known-answer test the reference holds for this path, the upstream
 * rand_pcg vectors and the independent survey restatement (tests/test_oracle_*.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  No FMA contraction is
 * allowed: Rust never contracts, and the CUDA side is built with -fmad=false.
 */
#include "cp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../crystal-packing_b200/csrc/cp_math.h"

#define ORC_PI 3.14159265358979323846264338327950288 /* std::f64::consts::PI */
#define ORC_TAU 6.28318530717958647692528676655900577 /* std::f64::consts::TAU */

/* ------------------------------------------------------------------ counters */
static __thread orc_counters g_cnt;
/* The operation counters feed bench.py's FLOPs-per-move figure.  The TIMED build of the oracle
 * (libcp_oracle_fast.so: -O3 -march=native -DORC_NO_COUNTERS, see the Makefile) compiles them out:
 * a thread-local read-modify-write inside every segment / disc / LJ test would handicap the CPU
 * baseline against a release build of the reference, which has none. */
#ifdef ORC_NO_COUNTERS
#define ORC_COUNT(stmt) ((void)0)
#else
#define ORC_COUNT(stmt) (g_cnt.stmt)
#endif

void orc_counters_reset(void) { memset(&g_cnt, 0, sizeof g_cnt); }
void orc_counters_get(orc_counters *out) { *out = g_cnt; }

static void counters_add(orc_counters *a, const orc_counters *b)
{
    a->moves += b->moves;
    a->scored += b->scored;
    a->segment_tests += b->segment_tests;
    a->disc_tests += b->disc_tests;
    a->lj_in += b->lj_in;
    a->lj_out += b->lj_out;
    a->image_cands += b->image_cands;
    a->image_pass += b->image_pass;
    a->point_xforms += b->point_xforms;
    a->site_images += b->site_images;
    a->trig += b->trig;
}

/* unit costs from SURVEY.md §8d: segment test 22, disc test 7, LJ in 28 / out 11, image candidate
 * 14, point transform 8, site image 61, proposal+accept 12 per move, transcendental 1 */
double orc_counters_flops(const orc_counters *c)
{
    return 12.0 * (double)c->moves + 22.0 * (double)c->segment_tests + 7.0 * (double)c->disc_tests +
           28.0 * (double)c->lj_in + 11.0 * (double)c->lj_out + 14.0 * (double)c->image_cands +
           8.0 * (double)c->point_xforms + 61.0 * (double)c->site_images + 1.0 * (double)c->trig;
}

/* ------------------------------------------------------------------ RNG */
/* rand_pcg 0.3.1 Pcg64Mcg (pcg128.rs): MCG, 128-bit state, XSL-RR 128/64 output */
static const uint64_t PCG_MUL_HI = 0x2360ED051FC65DA4ull;
static const uint64_t PCG_MUL_LO = 0x4385DF649FCCF645ull;

void orc_pcg_new(orc_pcg *r, uint64_t state_lo, uint64_t state_hi)
{
    /* Mcg128Xsl64::new: state | 1 */
    r->lo = state_lo | 1ull;
    r->hi = state_hi;
}

void orc_pcg_from_seed(orc_pcg *r, const uint8_t seed[16])
{
    /* from_seed: read two little-endian u64, state = lo | hi << 64, then | 1 */
    uint64_t lo = 0, hi = 0;
    for (int i = 7; i >= 0; --i) {
        lo = (lo << 8) | seed[i];
        hi = (hi << 8) | seed[8 + i];
    }
    orc_pcg_new(r, lo, hi);
}

void orc_pcg_seed_from_u64(orc_pcg *r, uint64_t state)
{
    /* rand_core 0.6.3 SeedableRng::seed_from_u64: PCG32 stream fills the seed 4 bytes at a time */
    const uint64_t MUL = 6364136223846793005ull;
    const uint64_t INC = 11634580027462260723ull;
    uint8_t seed[16];
    for (int chunk = 0; chunk < 4; ++chunk) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        seed[4 * chunk + 0] = (uint8_t)(x);
        seed[4 * chunk + 1] = (uint8_t)(x >> 8);
        seed[4 * chunk + 2] = (uint8_t)(x >> 16);
        seed[4 * chunk + 3] = (uint8_t)(x >> 24);
    }
    orc_pcg_from_seed(r, seed);
}

uint64_t orc_pcg_next_u64(orc_pcg *r)
{
    unsigned __int128 s = ((unsigned __int128)r->hi << 64) | r->lo;
    unsigned __int128 m = ((unsigned __int128)PCG_MUL_HI << 64) | PCG_MUL_LO;
    s *= m;
    r->lo = (uint64_t)s;
    r->hi = (uint64_t)(s >> 64);
    uint32_t rot = (uint32_t)(r->hi >> 58); /* state >> 122 */
    uint64_t xsl = r->hi ^ r->lo;
    return (xsl >> rot) | (xsl << ((64 - rot) & 63));
}

uint64_t orc_uniform_usize(orc_pcg *r, uint64_t range)
{
    /* rand 0.8.4 UniformInt<usize>::sample (uniform.rs): widening multiply with rejection zone */
    uint64_t ints_to_reject = (UINT64_MAX - range + 1) % range;
    uint64_t zone = UINT64_MAX - ints_to_reject;
    for (;;) {
        uint64_t v = orc_pcg_next_u64(r);
        unsigned __int128 w = (unsigned __int128)v * range;
        uint64_t hi = (uint64_t)(w >> 64), lo = (uint64_t)w;
        if (lo <= zone)
            return hi;
    }
}

double orc_uniform_step(orc_pcg *r)
{
    /* rand 0.8.4 UniformFloat<f64>::sample for Uniform::from(-0.5..0.5): scale=1, low=-0.5 */
    uint64_t bits = (orc_pcg_next_u64(r) >> 12) | 0x3FF0000000000000ull;
    double value1_2 = cp_bits_to_double(bits);
    double value0_1 = value1_2 - 1.0;
    return value0_1 * 1.0 + -0.5;
}

double orc_gen_f64(orc_pcg *r)
{
    /* rand 0.8.4 Standard for f64: 53 random bits scaled by 2^-53 */
    return (double)(orc_pcg_next_u64(r) >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------ math */
void orc_cp_sincos(double x, double *s, double *c) { cp_sincos(x, s, c); }
double orc_cp_exp(double x) { return cp_exp(x); }

void orc_sincos(int math_mode, double x, double *s, double *c)
{
    ORC_COUNT(trig += 2);
    if (math_mode == ORC_MATH_CP) {
        cp_sincos(x, s, c);
    } else {
        *s = sin(x);
        *c = cos(x);
    }
}

double orc_exp(int math_mode, double x)
{
    ORC_COUNT(trig += 1);
    return math_mode == ORC_MATH_CP ? cp_exp(x) : exp(x);
}

/* ------------------------------------------------------------------ Transform2 */
void orc_transform_identity(orc_transform *out)
{
    memset(out, 0, sizeof *out);
    out->m[0][0] = out->m[1][1] = out->m[2][2] = 1.0;
}

void orc_transform_new(int math_mode, double rotation, double tx, double ty, orc_transform *out)
{
    /* transform.rs:81-87: IsometryMatrix2(translation, Rotation2::new(angle)).to_homogeneous() */
    double s, c;
    orc_sincos(math_mode, rotation, &s, &c);
    out->m[0][0] = c;
    out->m[0][1] = -s;
    out->m[0][2] = tx;
    out->m[1][0] = s;
    out->m[1][1] = c;
    out->m[1][2] = ty;
    out->m[2][0] = 0.0;
    out->m[2][1] = 0.0;
    out->m[2][2] = 1.0;
}

void orc_transform_mul(const orc_transform *a, const orc_transform *b, orc_transform *out)
{
    /* transform.rs:72-78 -> nalgebra 0.27 Matrix3 * Matrix3 (column gemv, no FMA):
     * C[i][j] = ((A[i][0]*B[0][j]) + A[i][1]*B[1][j]) + A[i][2]*B[2][j]    */
    orc_transform r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.m[i][j] = ((a->m[i][0] * b->m[0][j]) + a->m[i][1] * b->m[1][j]) + a->m[i][2] * b->m[2][j];
    *out = r;
}

void orc_transform_mul_point(const orc_transform *t, double px, double py, double *ox, double *oy)
{
    /* transform.rs:64-70 -> nalgebra Transform<TGeneral> * Point: (M2x2*p + translation), divided
     * by the normaliser n = m20*px + m21*py + m22 only when n != 0 */
    ORC_COUNT(point_xforms += 1);
    double vx = (t->m[0][0] * px + t->m[0][1] * py) + t->m[0][2];
    double vy = (t->m[1][0] * px + t->m[1][1] * py) + t->m[1][2];
    double n = (t->m[2][0] * px + t->m[2][1] * py) + t->m[2][2];
    if (n != 0.0) {
        vx = vx / n;
        vy = vy / n;
    }
    *ox = vx;
    *oy = vy;
}

void orc_transform_position(const orc_transform *t, double *ox, double *oy)
{
    /* transform.rs:93-95: self.0 * Point2::origin() */
    orc_transform_mul_point(t, 0.0, 0.0, ox, oy);
    ORC_COUNT(point_xforms -= 1); /* bookkeeping: position() is a column read, not counted as work */
}

static double wrap_coord(double v, double period, double offset)
{
    /* transform.rs:109-110 */
    return fmod(fmod(v - offset, period) + period, period) + offset;
}

void orc_transform_periodic(const orc_transform *t, double period, double offset, orc_transform *out)
{
    /* transform.rs:107-112 */
    double px, py;
    orc_transform_position(t, &px, &py);
    px = wrap_coord(px, period, offset);
    py = wrap_coord(py, period, offset);
    *out = *t;
    out->m[0][2] = px; /* set_position, transform.rs:101-105 */
    out->m[1][2] = py;
}

int orc_transform_from_operations(const char *ops, orc_transform *out)
{
    /* transform.rs:124-183.  Returns 0 ok, 1 wrong dimension count, 2 invalid character */
    size_t len = strlen(ops);
    size_t b = 0, e = len;
    while (b < e && (ops[b] == '(' || ops[b] == ')'))
        ++b; /* trim_matches(braces) */
    while (e > b && (ops[e - 1] == '(' || ops[e - 1] == ')'))
        --e;
    /* split_terminator(','): split at every comma; a trailing empty piece is dropped */
    int n_ops = 0;
    size_t starts[4], ends[4];
    size_t cur = b;
    for (size_t i = b; i <= e; ++i) {
        if (i < e && ops[i] != ',')
            continue;
        if (i == e && cur == e)
            break; /* nothing after the last comma (or empty input) */
        if (n_ops < 4) {
            starts[n_ops] = cur;
            ends[n_ops] = i;
        }
        ++n_ops;
        cur = i + 1;
    }
    if (n_ops != 2)
        return 1;
    memset(out, 0, sizeof *out); /* Matrix3::zeros(): row 2 stays all zero */
    for (int index = 0; index < 2; ++index) {
        double sign = 1.0, constant = 0.0;
        char op = 0;
        for (size_t i = starts[index]; i < ends[index]; ++i) {
            char c = ops[i];
            if (c == 'x') {
                out->m[index][0] = sign;
                sign = 1.0;
            } else if (c == 'y') {
                out->m[index][1] = sign;
                sign = 1.0;
            } else if (c == '*' || c == '/') {
                op = c;
            } else if (c == '-') {
                sign = -1.0;
            } else if (c >= '0' && c <= '9') {
                double val = (double)(c - '0');
                if (op == '/' || op == '*')
                    constant = sign * constant / val;
                else
                    constant = sign * val;
                op = 0;
                sign = 1.0;
            } else if (c == ' ' || c == '+') {
            } else {
                return 2;
            }
        }
        out->m[index][2] = constant;
    }
    return 0;
}

/* ------------------------------------------------------------------ Cell2 */
static double cell_a(const orc_state *s) { return s->length; }            /* cell.rs:91-93 */
static double cell_b(const orc_state *s) { return s->length * s->ratio; } /* cell.rs:95-97 */

void orc_cell_to_cartesian(const orc_state *s, double x, double y, double *ox, double *oy)
{
    /* cell.rs:258-263: (x*a + y*b*cos(angle), y*b*sin(angle)); the reference evaluates cos and
     * sin of the cell angle on every call */
    double sn, cs;
    orc_sincos(s->math_mode, s->angle, &sn, &cs);
    ORC_COUNT(trig -= 2); /* the reference re-evaluates cos/sin per call; §8d counts the cell trig once per score */
    *ox = x * cell_a(s) + y * cell_b(s) * cs;
    *oy = y * cell_b(s) * sn;
}

double orc_cell_area(const orc_state *s)
{
    /* cell.rs:190-192 */
    double sn, cs;
    orc_sincos(s->math_mode, s->angle, &sn, &cs);
    ORC_COUNT(trig -= 2);
    return sn * cell_a(s) * cell_b(s);
}

static void cell_to_cartesian_isometry(const orc_state *s, const orc_transform *t, orc_transform *out)
{
    /* cell.rs:110-112 */
    double px, py, cx, cy;
    orc_transform_position(t, &px, &py);
    orc_cell_to_cartesian(s, px, py, &cx, &cy);
    *out = *t;
    out->m[0][2] = cx;
    out->m[1][2] = cy;
}

static void cell_to_cartesian_translate(const orc_state *s, const orc_transform *t, int64_t x,
                                        int64_t y, orc_transform *out)
{
    /* cell.rs:241-245: Translation2(x, y) * position, then to_cartesian_point, keep rotation */
    double px, py, cx, cy;
    orc_transform_position(t, &px, &py);
    orc_cell_to_cartesian(s, px + (double)x, py + (double)y, &cx, &cy);
    *out = *t;
    out->m[0][2] = cx;
    out->m[1][2] = cy;
}

int orc_cell_periodic_images(const orc_state *s, const orc_transform *t, int shells, int zero,
                             orc_transform *out)
{
    /* cell.rs:216-225: iproduct!(-shells..=shells, -shells..=shells), x outer, y inner */
    int n = 0;
    for (int64_t x = -shells; x <= shells; ++x)
        for (int64_t y = -shells; y <= shells; ++y) {
            if (!zero && x == 0 && y == 0)
                continue;
            cell_to_cartesian_translate(s, t, x, y, &out[n++]);
        }
    return n;
}

/* ------------------------------------------------------------------ components */
int orc_line_intersects(const orc_item *s, const orc_item *o)
{
    /* line2.rs:29-54 with dx/dy from :94-101; item = (x0,y0,x1,y1) */
    ORC_COUNT(segment_tests += 1);
    double s_dx = s->c - s->a, s_dy = s->d - s->b;
    double o_dx = o->c - o->a, o_dy = o->d - o->b;
    double u_b = o_dy * s_dx - o_dx * s_dy;
    if (u_b == 0.0)
        return 0;
    double ua_t = o_dx * (s->b - o->b) - o_dy * (s->a - o->a);
    double ub_t = s_dx * (s->b - o->b) - s_dy * (s->a - o->a);
    double ua = ua_t / u_b;
    double ub = ub_t / u_b;
    if (0.0 <= ua && ua <= 1.0 && 0.0 <= ub && ub <= 1.0)
        return 1;
    return 0;
}

int orc_atom_intersects(const orc_item *s, const orc_item *o)
{
    /* atom2.rs:21-26; item = (x, y, radius) */
    ORC_COUNT(disc_tests += 1);
    double rs = s->c + o->c;
    double r_squared = rs * rs;
    double dx = s->a - o->a, dy = s->b - o->b;
    return (dx * dx + dy * dy) < r_squared;
}

static double powi6(double x)
{ /* compiler-rt __powidf2 / LLVM powi expansion: x^2 * x^4 */
    double x2 = x * x;
    double x4 = x2 * x2;
    return x2 * x4;
}
static double powi12(double x)
{ /* x^4 * x^8 */
    double x2 = x * x;
    double x4 = x2 * x2;
    double x8 = x4 * x4;
    return x4 * x8;
}

double orc_lj_energy(const orc_item *s, const orc_item *o)
{
    /* lj2.rs:43-60; item = (x, y, sigma, epsilon, cutoff|NaN); uses only self's parameters */
    double sigma_squared = s->c * s->c;
    double dx = s->a - o->a, dy = s->b - o->b;
    double r_squared = dx * dx + dy * dy;
    double q = sigma_squared / r_squared;
    double c3 = (q * q) * q; /* powi(3) */
    if (s->e == s->e) {      /* Some(cutoff) */
        double x = s->e;
        if (r_squared < x * x) {
            ORC_COUNT(lj_in += 1);
            double sx = s->c / x;
            double shift = 4.0 * s->d * (powi12(sx) - powi6(s->c / x));
            return 4.0 * s->d * (c3 * c3 - c3) - shift;
        }
        ORC_COUNT(lj_out += 1);
        return 0.0;
    }
    ORC_COUNT(lj_in += 1);
    return 4.0 * s->d * (c3 * c3 - c3);
}

/* ------------------------------------------------------------------ shapes */
int orc_shape_from_radial(const double *radii, int n, orc_item *items)
{
    /* line_shape.rs:128-146: the end point is computed from angle + dtheta, not copied */
    if (n < 3 || n > ORC_MAX_ITEMS)
        return 1;
    double dtheta = 2.0 * ORC_PI / (double)n;
    for (int index = 0; index < n; ++index) {
        double r1 = radii[index], r2 = radii[(index + 1) % n];
        double angle = (double)index * dtheta;
        items[index].a = r1 * sin(angle);
        items[index].b = r1 * cos(angle);
        items[index].c = r2 * sin(angle + dtheta);
        items[index].d = r2 * cos(angle + dtheta);
        items[index].e = 0.0;
    }
    return 0;
}

static double to_radians(double deg)
{ /* f64::to_radians: deg * (PI / 180) */
    return deg * (ORC_PI / 180.0);
}

void orc_shape_trimer(int kind, double radius, double angle, double distance, orc_item *items)
{
    memset(items, 0, 3 * sizeof *items);
    if (kind == ORC_LJ) {
        /* lj_shape.rs:111-131 */
        double x_base = distance * sin(to_radians(angle) / 2.0);
        double y_base = 1.0 / 3.0 * distance * cos(to_radians(angle) / 2.0);
        double rr[3] = {1.0, radius, radius};
        double px[3] = {0.0, -x_base, x_base};
        double py[3] = {-2.0 * y_base, y_base, y_base};
        for (int i = 0; i < 3; ++i) {
            items[i].a = px[i];
            items[i].b = py[i];
            items[i].c = 2.0 * rr[i]; /* sigma */
            items[i].d = 1.0;         /* epsilon (Default) */
            items[i].e = 3.5;         /* cutoff Some(3.5) */
        }
    } else {
        /* molecular_shape2.rs:141-162 */
        double half = to_radians(angle) / 2.0;
        items[0].a = 0.0;
        items[0].b = -2.0 / 3.0 * distance * cos(half);
        items[0].c = 1.0;
        items[1].a = -distance * sin(half);
        items[1].b = 1.0 / 3.0 * distance * cos(half);
        items[1].c = radius;
        items[2].a = distance * sin(half);
        items[2].b = 1.0 / 3.0 * distance * cos(half);
        items[2].c = radius;
    }
}

void orc_shape_circle(int kind, orc_item *items)
{
    memset(items, 0, sizeof *items);
    if (kind == ORC_LJ) {
        /* lj_shape.rs:146-151: LJ2::new(0,0,1): sigma 1, epsilon 1, no cutoff */
        items[0].c = 1.0;
        items[0].d = 1.0;
        items[0].e = NAN;
    } else {
        /* molecular_shape2.rs:167-172 */
        items[0].c = 1.0;
    }
}

static double dist_origin(double x, double y) { return sqrt(x * x + y * y); }

static double overlap_area(double r, double d)
{
    /* molecular_shape2.rs:103-105 */
    return (r * r) * acos(d / r) - d * sqrt((r * r) - (d * d));
}

double orc_circle_overlap(const orc_item *a1, const orc_item *a2)
{
    /* molecular_shape2.rs:107-117 */
    double dx = a1->a - a2->a, dy = a1->b - a2->b;
    double distance = sqrt(dx * dx + dy * dy);
    if (distance < a1->c + a2->c) {
        double d1 = (distance * distance + a1->c * a1->c - a2->c * a2->c) / (2.0 * distance);
        double d2 = (distance * distance + a2->c * a2->c - a1->c * a1->c) / (2.0 * distance);
        return overlap_area(a1->c, d1) + overlap_area(a2->c, d2);
    }
    return 0.0;
}

double orc_shape_area(int kind, const orc_item *items, int n)
{
    if (kind == ORC_POLYGON) {
        /* line_shape.rs:60-72 */
        double angle_term = sin(2.0 * ORC_PI / (double)n);
        double sum = 0.0;
        for (int i = 0; i < n; ++i)
            sum += 0.5 * angle_term * dist_origin(items[i].a, items[i].b) *
                   dist_origin(items[i].c, items[i].d);
        return sum;
    }
    if (kind == ORC_DISCS) {
        /* molecular_shape2.rs:39-52 */
        double total = 0.0, overlap = 0.0;
        for (int i = 0; i < n; ++i)
            total += ORC_PI * (items[i].c * items[i].c);
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j)
                overlap += orc_circle_overlap(&items[i], &items[j]);
        return total - overlap;
    }
    return 0.0; /* LJShape2 has no area */
}

double orc_shape_enclosing_radius(int kind, const orc_item *items, int n)
{
    /* line_shape.rs:86-93, molecular_shape2.rs:66-74, lj_shape.rs:52-60 */
    double r = -1.7976931348623157e308; /* f64::MIN */
    for (int i = 0; i < n; ++i) {
        double v = dist_origin(items[i].a, items[i].b);
        if (kind == ORC_DISCS)
            v = v + items[i].c;
        else if (kind == ORC_LJ)
            v = v + items[i].c / 2.0;
        r = fmax(r, v);
    }
    return r;
}

void orc_shape_transform(int kind, const orc_item *items, int n, const orc_transform *t,
                         orc_item *out)
{
    /* line_shape.rs:103-108 + line2_ops.rs:24-33; molecular_shape2.rs:84-89 + atom2_ops.rs;
     * lj_shape.rs:70-75 + lj2_ops.rs */
    for (int i = 0; i < n; ++i) {
        out[i] = items[i];
        orc_transform_mul_point(t, items[i].a, items[i].b, &out[i].a, &out[i].b);
        if (kind == ORC_POLYGON)
            orc_transform_mul_point(t, items[i].c, items[i].d, &out[i].c, &out[i].d);
    }
}

int orc_shape_intersects(int kind, const orc_item *a, const orc_item *b, int n)
{
    /* line_shape.rs:55-58 / molecular_shape2.rs:36-38: iproduct!(self, other).any(...) */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            int hit = kind == ORC_POLYGON ? orc_line_intersects(&a[i], &b[j])
                                          : orc_atom_intersects(&a[i], &b[j]);
            if (hit)
                return 1;
        }
    return 0;
}

double orc_shape_energy(const orc_item *a, const orc_item *b, int n)
{
    /* lj_shape.rs:35-39: iproduct!(self, other).map(energy).sum() */
    double sum = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            sum += orc_lj_energy(&a[i], &b[j]);
    return sum;
}

/* ------------------------------------------------------------------ wallpaper groups */
int orc_group(const char *name, int32_t *family, int32_t *n_syms, orc_transform *syms)
{
    /* wallpaper.rs:100-110 (name parsing, "pg" alias) and :134-176 (table) */
    static const struct {
        const char *name;
        int family;
        int n;
        const char *ops[4];
    } table[] = {
        {"p1", ORC_MONOCLINIC, 1, {"x,y"}},
        {"p2", ORC_MONOCLINIC, 2, {"x,y", "-x,-y"}},
        {"p1m1", ORC_ORTHORHOMBIC, 2, {"x,y", "-x,y"}},
        {"p1g1", ORC_ORTHORHOMBIC, 2, {"x,y", "-x,y+1/2"}},
        {"pg", ORC_ORTHORHOMBIC, 2, {"x,y", "-x,y+1/2"}},
        {"p2mm", ORC_ORTHORHOMBIC, 4, {"x,y", "-x,-y", "-x,y", "x,-y"}},
        {"p2mg", ORC_ORTHORHOMBIC, 4, {"x,y", "-x, -y", "-x+1/2, y", "x+1/2, -y"}},
        {"p2gg", ORC_ORTHORHOMBIC, 4, {"x,y", "-x, -y", "-x+1/2, y+1/2", "x+1/2, -y+1/2"}},
    };
    for (size_t g = 0; g < sizeof table / sizeof table[0]; ++g) {
        if (strcmp(name, table[g].name) != 0)
            continue;
        *family = table[g].family;
        *n_syms = table[g].n;
        for (int i = 0; i < table[g].n; ++i)
            if (orc_transform_from_operations(table[g].ops[i], &syms[i]) != 0)
                return 2;
        return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------ states */
int orc_state_total_shapes(const orc_state *s) { return s->n_syms; } /* packed.rs:74-78 */

int orc_state_n_dof(const orc_state *s)
{
    /* cell.rs:136-170 + site.rs:65-91 */
    int cell = s->family == ORC_MONOCLINIC ? 3 : s->family == ORC_ORTHORHOMBIC ? 2 : 1;
    return cell + 3;
}

int orc_state_from_group(orc_state *s, int kind, const orc_item *items, int n_items,
                         const char *group, int math_mode)
{
    /* packed.rs:169-195 / potential.rs:148-174; cell.rs:200-214; site.rs:51-63 */
    if (n_items < 1 || n_items > ORC_MAX_ITEMS)
        return 3;
    memset(s, 0, sizeof *s);
    s->kind = kind;
    s->n_items = n_items;
    memcpy(s->items, items, (size_t)n_items * sizeof *items);
    s->math_mode = math_mode;
    int rc = orc_group(group, &s->family, &s->n_syms, s->syms);
    if (rc)
        return rc;
    double radius = orc_shape_enclosing_radius(kind, items, n_items);
    double num_shapes = (double)s->n_syms;
    double max_cell_size = (kind == ORC_LJ ? 2.0 : 4.0) * radius * num_shapes;
    s->length = max_cell_size;
    s->ratio = 1.0;
    s->angle = s->family == ORC_HEXAGONAL ? ORC_PI / 3.0 : ORC_PI / 2.0;
    double position = -0.5 + 0.5 / (double)s->n_syms;
    s->x = position;
    s->y = position;
    s->theta = 0.0;
    return 0;
}

int orc_state_relative_positions(const orc_state *s, orc_transform *out)
{
    /* packed.rs:117-119 -> site.rs:33-45: (sym * Transform2::new(angle,(x,y))).periodic(1,-0.5) */
    orc_transform t, p;
    orc_transform_new(s->math_mode, s->theta, s->x, s->y, &t);
    for (int i = 0; i < s->n_syms; ++i) {
        ORC_COUNT(site_images += 1);
        orc_transform_mul(&s->syms[i], &t, &p);
        orc_transform_periodic(&p, 1.0, -0.5, &out[i]);
    }
    return s->n_syms;
}

int orc_state_cartesian_positions(const orc_state *s, orc_transform *out)
{
    /* packed.rs:112-115 */
    orc_transform rel[ORC_MAX_SYMS];
    int n = orc_state_relative_positions(s, rel);
    for (int i = 0; i < n; ++i)
        cell_to_cartesian_isometry(s, &rel[i], &out[i]);
    return n;
}

int orc_state_check_intersection(const orc_state *s)
{
    /* packed.rs:127-167.  The reference's lazy iterators re-derive positions and transformed
     * shapes many times; the values are pure functions of the state, so evaluating them once is
     * arithmetically identical.  (Counters count each distinct evaluation once.) */
    double p = cell_a(s) / cell_b(s), a = s->angle;
    int periodic_range;
    if (0.5 < p && p < 2.0 && fabs(a - ORC_PI / 2.0) < 0.2)
        periodic_range = 1;
    else if (0.3 < p && p < 3.0 && fabs(a - ORC_PI / 2.0) < 0.5)
        periodic_range = 2;
    else
        periodic_range = 3;

    orc_transform rel[ORC_MAX_SYMS], cart[ORC_MAX_SYMS];
    static __thread orc_item shapes[ORC_MAX_SYMS][ORC_MAX_ITEMS];
    int n = orc_state_relative_positions(s, rel);
    for (int i = 0; i < n; ++i) {
        cell_to_cartesian_isometry(s, &rel[i], &cart[i]);
        orc_shape_transform(s->kind, s->items, s->n_items, &cart[i], shapes[i]);
    }
    /* within the cell: i < j, no distance filter (packed.rs:134-148) */
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (orc_shape_intersects(s->kind, shapes[i], shapes[j], s->n_items))
                return 1;

    double er2 = orc_shape_enclosing_radius(s->kind, s->items, s->n_items) * 2.0;
    double radius_sq = er2 * er2; /* packed.rs:150 */
    orc_transform images[49];
    orc_item shape2[ORC_MAX_ITEMS];
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            int ni = orc_cell_periodic_images(s, &rel[j], periodic_range, 0, images);
            for (int k = 0; k < ni; ++k) {
                ORC_COUNT(image_cands += 1);
                double dx = cart[i].m[0][2] - images[k].m[0][2];
                double dy = cart[i].m[1][2] - images[k].m[1][2];
                double distance = dx * dx + dy * dy;
                if (distance <= radius_sq) {
                    ORC_COUNT(image_pass += 1);
                    orc_shape_transform(s->kind, s->items, s->n_items, &images[k], shape2);
                    if (orc_shape_intersects(s->kind, shapes[i], shape2, s->n_items))
                        return 1;
                }
            }
        }
    }
    return 0;
}

static double potential_score(const orc_state *s)
{
    /* potential.rs:82-115 */
    orc_transform rel[ORC_MAX_SYMS], cart[ORC_MAX_SYMS];
    static __thread orc_item shapes[ORC_MAX_SYMS][ORC_MAX_ITEMS];
    int n = orc_state_relative_positions(s, rel);
    for (int i = 0; i < n; ++i) {
        cell_to_cartesian_isometry(s, &rel[i], &cart[i]);
        orc_shape_transform(s->kind, s->items, s->n_items, &cart[i], shapes[i]);
    }
    double sum = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            sum += orc_shape_energy(shapes[i], shapes[j], s->n_items);
    orc_transform images[49];
    orc_item shape2[ORC_MAX_ITEMS];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            int ni = orc_cell_periodic_images(s, &rel[j], 3, 0, images);
            for (int k = 0; k < ni; ++k) {
                ORC_COUNT(image_cands += 1);
                orc_shape_transform(s->kind, s->items, s->n_items, &images[k], shape2);
                sum += orc_shape_energy(shapes[i], shape2, s->n_items);
            }
        }
    return -sum / (double)orc_state_total_shapes(s);
}

double orc_state_score(const orc_state *s)
{
    ORC_COUNT(scored += 1);
    ORC_COUNT(trig += 2); /* cos + sin of the cell angle ("cell trig = 2") */
    if (s->kind == ORC_LJ)
        return potential_score(s);
    /* packed.rs:80-86 */
    if (orc_state_check_intersection(s))
        return NAN;
    double area = orc_shape_area(s->kind, s->items, s->n_items);
    return (area * (double)orc_state_total_shapes(s)) / orc_cell_area(s);
}

static double *dof_ptr(orc_state *s, int basis_index)
{
    /* basis order: packed.rs:88-95 = cell DOFs (cell.rs:136-170) then site DOFs (site.rs:65-91) */
    int n_cell = s->family == ORC_MONOCLINIC ? 3 : s->family == ORC_ORTHORHOMBIC ? 2 : 1;
    if (basis_index < n_cell) {
        if (basis_index == 0)
            return &s->length;
        if (basis_index == 1)
            return &s->ratio;
        return &s->angle;
    }
    switch (basis_index - n_cell) {
    case 0:
        return &s->x;
    case 1:
        return &s->y;
    default:
        return &s->theta;
    }
}

int orc_state_bounds(const orc_state *s, double lo[ORC_MAX_DOF], double hi[ORC_MAX_DOF])
{
    int k = 0;
    lo[k] = 0.01;
    hi[k++] = s->length;
    if (s->family == ORC_MONOCLINIC || s->family == ORC_ORTHORHOMBIC) {
        lo[k] = 0.1;
        hi[k++] = s->ratio;
    }
    if (s->family == ORC_MONOCLINIC) {
        lo[k] = ORC_PI / 4.0;
        hi[k++] = ORC_PI / 2.0;
    }
    lo[k] = -0.5;
    hi[k++] = 0.5;
    lo[k] = -0.5;
    hi[k++] = 0.5;
    lo[k] = 0.0;
    hi[k++] = ORC_TAU;
    return k;
}

void orc_state_set_dof(orc_state *s, const double dof[ORC_MAX_DOF])
{
    s->length = dof[0];
    s->ratio = dof[1];
    s->angle = dof[2];
    s->x = dof[3];
    s->y = dof[4];
    s->theta = dof[5];
}

void orc_state_get_dof(const orc_state *s, double dof[ORC_MAX_DOF])
{
    dof[0] = s->length;
    dof[1] = s->ratio;
    dof[2] = s->angle;
    dof[3] = s->x;
    dof[4] = s->y;
    dof[5] = s->theta;
}

/* ------------------------------------------------------------------ optimiser */
double orc_energy_surface(int math_mode, double new_score, double old_score, double kt)
{
    /* optimisation.rs:46-48: f64::min(f64::exp((new - old) / kt), 1.); Rust's f64::min returns
     * the non-NaN operand, so exp(NaN) -> 1 */
    double e = orc_exp(math_mode, (new_score - old_score) / kt);
    if (e != e)
        return 1.0;
    return e < 1.0 ? e : 1.0;
}

static int accept_score(int math_mode, double new_score, double old, double kt, double threshold,
                        double *accepted)
{
    /* optimisation.rs:55-78 (threshold is drawn by the caller, unconditionally) */
    if (new_score == new_score) {
        if (new_score > old) {
            *accepted = new_score;
            return 1;
        }
        if (threshold < orc_energy_surface(math_mode, new_score, old, kt)) {
            *accepted = new_score;
            return 1;
        }
    }
    return 0;
}

int orc_replay_move(orc_state *s, int basis_index, double new_value, double lo, double hi,
                    double threshold, double kt, double *score_current, int *in_bounds,
                    double *new_score)
{
    double *v = dof_ptr(s, basis_index);
    double val = *v;
    *new_score = NAN;
    if (!(lo <= new_value && new_value <= hi)) { /* basis/mod.rs:25-37 */
        *in_bounds = 0;
        return 0;
    }
    *in_bounds = 1;
    *v = new_value;
    double sc = orc_state_score(s);
    *new_score = sc;
    double acc;
    if (accept_score(s->math_mode, sc, *score_current, kt, threshold, &acc)) {
        *score_current = acc;
        return 1;
    }
    *v = val;
    return 0;
}

int orc_optimise_state(const orc_optimiser *opt, orc_state *s, orc_move_record *records,
                       uint64_t max_records, uint64_t *n_records, uint64_t *rejections_out)
{
    /* optimisation.rs:80-180 */
    double score_current = orc_state_score(s);
    if (n_records)
        *n_records = 0;
    if (rejections_out)
        *rejections_out = 0;
    if (!(score_current == score_current))
        return 1; /* panic!("Invalid configuration passed to function, exiting.") */

    orc_pcg rng;
    orc_pcg_seed_from_u64(&rng, opt->seed);
    uint64_t rejections = 0;
    double kt = opt->kt_start;

    double lo[ORC_MAX_DOF], hi[ORC_MAX_DOF];
    int n_basis = orc_state_bounds(s, lo, hi); /* generate_basis(): bounds captured here */

    double step_ratio = 1.0;
    int convergence_count = 0;
    uint64_t n_moves = 0;
    uint64_t outer = opt->inner_steps ? opt->steps / opt->inner_steps : 0;

    for (uint64_t loop_counter = 1; loop_counter <= outer; ++loop_counter) {
        double score_start = score_current;
        uint64_t loop_rejections = 0;
        for (uint64_t it = 0; it < opt->inner_steps; ++it) {
            ORC_COUNT(moves += 1);
            int basis_index = (int)orc_uniform_usize(&rng, (uint64_t)n_basis);
            double *b = dof_ptr(s, basis_index);
            double val = *b;
            double new_val = val + opt->max_step_size * step_ratio * 1.0 * orc_uniform_step(&rng);
            orc_move_record rec;
            memset(&rec, 0, sizeof rec);
            rec.basis_index = (uint32_t)basis_index;
            rec.old_value = val;
            rec.new_value = new_val;
            rec.kt = kt;
            rec.step_ratio = step_ratio;
            rec.threshold = NAN;
            rec.new_score = NAN;
            if (!(lo[basis_index] <= new_val && new_val <= hi[basis_index])) {
                loop_rejections += 1;
                rec.in_bounds = 0;
            } else {
                rec.in_bounds = 1;
                *b = new_val;
                double new_score = orc_state_score(s);
                double threshold = orc_gen_f64(&rng);
                rec.threshold = threshold;
                rec.new_score = new_score;
                double acc;
                if (accept_score(s->math_mode, new_score, score_current, kt, threshold, &acc)) {
                    score_current = acc;
                    rec.accepted = 1;
                } else {
                    *b = val;
                    loop_rejections += 1;
                }
            }
            rec.score_current = score_current;
            if (records && n_moves < max_records)
                records[n_moves] = rec;
            n_moves += 1;
        }
        rejections += loop_rejections;
        kt *= opt->kt_ratio;

        if (opt->convergence == opt->convergence) { /* Some(precision) */
            if (score_current - score_start < opt->convergence) {
                convergence_count += 1;
                if (convergence_count > 5)
                    goto done; /* early return: no final assert on this path */
            } else {
                convergence_count = 0;
            }
        }
        if (step_ratio > 1e-4)
            step_ratio *= (double)opt->inner_steps / ((double)loop_rejections + 1.0);
    }
    {
        /* assert!(state.score().is_some()) */
        double final_score = orc_state_score(s);
        ORC_COUNT(scored -= 1);
        if (!(final_score == final_score)) {
            if (n_records)
                *n_records = n_moves;
            if (rejections_out)
                *rejections_out = rejections;
            return 2;
        }
    }
done:
    if (n_records)
        *n_records = n_moves;
    if (rejections_out)
        *rejections_out = rejections;
    return 0;
}

void orc_build(const orc_build_optimiser *b, uint64_t seed, orc_optimiser *out)
{
    /* main.rs:122-143 */
    double kt_ratio;
    if (b->kt_ratio == b->kt_ratio)
        kt_ratio = 1.0 - b->kt_ratio;
    else if (b->kt_finish == b->kt_finish)
        kt_ratio = pow(b->kt_finish / b->kt_start, 1.0 / (double)b->steps);
    else
        kt_ratio = 0.1;
    out->kt_start = b->kt_start;
    out->kt_ratio = kt_ratio;
    out->max_step_size = b->max_step_size;
    out->steps = b->steps;
    out->inner_steps = b->inner_steps < b->steps ? b->inner_steps : b->steps;
    out->seed = seed;
    out->convergence = b->convergence;
}

int orc_run_replica(const orc_state *tmpl, const orc_build_optimiser *b, uint64_t index,
                    orc_state *out_state, orc_result *out)
{
    /* main.rs:224-252: three stages, every stage seeded with the replica index */
    orc_state s = *tmpl;
    orc_optimiser opt;
    uint64_t rej = 0, n = 0, total_rej = 0, total_moves = 0;
    int rc;

    orc_build_optimiser b1 = *b; /* .steps(100).kt_start(0.).seed(index).convergence(None) */
    b1.steps = 100;
    b1.kt_start = 0.0;
    b1.convergence = NAN;
    orc_build(&b1, index, &opt);
    rc = orc_optimise_state(&opt, &s, NULL, 0, &n, &rej);
    total_rej += rej;
    total_moves += n;
    if (rc)
        goto fail;

    orc_build(b, index, &opt); /* .seed(index) */
    rc = orc_optimise_state(&opt, &s, NULL, 0, &n, &rej);
    total_rej += rej;
    total_moves += n;
    if (rc)
        goto fail;

    orc_build_optimiser b3 = *b; /* .kt_start(0.).seed(index) */
    b3.kt_start = 0.0;
    orc_build(&b3, index, &opt);
    rc = orc_optimise_state(&opt, &s, NULL, 0, &n, &rej);
    total_rej += rej;
    total_moves += n;
fail:
    if (out_state)
        *out_state = s;
    if (out) {
        orc_counters saved = g_cnt; /* the reporting score() below is not MC work */
        out->score = rc ? NAN : orc_state_score(&s);
        g_cnt = saved;
        orc_state_get_dof(&s, out->dof);
        out->rejections = total_rej;
        out->moves = total_moves;
    }
    return rc;
}

typedef struct {
    const orc_state *tmpl;
    const orc_build_optimiser *b;
    uint64_t begin, n;
    int tid, n_threads;
    orc_result *results;
    orc_result best;
    int64_t best_index;
    orc_counters counters;
} worker_arg;

static void *worker(void *p)
{
    worker_arg *w = (worker_arg *)p;
    orc_counters_reset();
    w->best_index = -1;
    for (uint64_t i = (uint64_t)w->tid; i < w->n; i += (uint64_t)w->n_threads) {
        orc_result r;
        orc_run_replica(w->tmpl, w->b, w->begin + i, NULL, &r);
        if (w->results)
            w->results[i] = r;
        /* .max() over states ordered by score (packed.rs:49-68).  rayon's max() is
         * reduce_with(Ord::max) in index order and Ord::max returns its second argument when the two
         * compare equal, so of several replicas with the maximal score the LAST one wins. */
        if (r.score == r.score && (w->best_index < 0 || r.score >= w->best.score)) {
            w->best = r;
            w->best_index = (int64_t)i;
        }
    }
    orc_counters_get(&w->counters);
    return NULL;
}

int64_t orc_analyse_state_counted(const orc_state *tmpl, const orc_build_optimiser *b,
                                  uint64_t begin, uint64_t n, int n_threads, orc_result *results,
                                  orc_result *best, orc_counters *counters)
{
    /* main.rs:221-253: into_par_iter over replica indices, three-stage map, .max() */
    if (n_threads < 1)
        n_threads = 1;
    if ((uint64_t)n_threads > n && n > 0)
        n_threads = (int)n;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    worker_arg *args = (worker_arg *)calloc((size_t)n_threads, sizeof *args);
    for (int t = 0; t < n_threads; ++t) {
        args[t].tmpl = tmpl;
        args[t].b = b;
        args[t].begin = begin;
        args[t].n = n;
        args[t].tid = t;
        args[t].n_threads = n_threads;
        args[t].results = results;
        pthread_create(&th[t], NULL, worker, &args[t]);
    }
    int64_t best_index = -1;
    orc_result best_r;
    memset(&best_r, 0, sizeof best_r);
    best_r.score = NAN;
    orc_counters total;
    memset(&total, 0, sizeof total);
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        counters_add(&total, &args[t].counters);
        if (args[t].best_index >= 0 &&
            (best_index < 0 || args[t].best.score > best_r.score ||
             (args[t].best.score == best_r.score && args[t].best_index > best_index))) {
            best_r = args[t].best;
            best_index = args[t].best_index;
        }
    }
    free(th);
    free(args);
    if (best)
        *best = best_r;
    if (counters)
        *counters = total;
    return best_index;
}

int64_t orc_analyse_state(const orc_state *tmpl, const orc_build_optimiser *b, uint64_t begin,
                          uint64_t n, int n_threads, orc_result *results, orc_result *best)
{
    return orc_analyse_state_counted(tmpl, b, begin, n, n_threads, results, best, NULL);
}
