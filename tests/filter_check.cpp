// filter_check.cpp — CPU soundness check and statistics for the f32 certificates of cp_filter.h.
//
// TEST INFRASTRUCTURE.  Links the CPU oracle (oracle/libcp_oracle.so) and runs real three-stage
// trajectories; at every scored proposal it rebuilds the candidate list the thread kernel uses
// (canonical half of the ordered shape pairs, cp_thread.cuh), evaluates the f32 certificate of each
// candidate with the SAME header the kernel compiles, and checks against the oracle's f64 evaluation:
//   * "clear" candidates: every component test of the candidate and of its mirror image is false;
//   * "hit" candidates: the shape test is true for the candidate and for its mirror image;
//   * the verdict assembled from certificates + f64 evaluation of the ambiguous candidates equals
//     orc_state_check_intersection.
// Prints one JSON line with the statistics the kernel design rests on.
//
// usage: filter_check <polygon|star|soup|trimer|circle> <sides> <group> <replicas> <steps> [kt_start kt_finish]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../crystal-packing_b200/csrc/cp_filter.h"
#include "../oracle/cp_oracle.h"

struct Pts {
    const float *px, *py, *pr;
    float x(int k) const { return px[k]; }
    float y(int k) const { return py[k]; }
    float r(int k) const { return pr[k]; }
};

// the same points with the row signs of an image applied (the kernel's TPtsSigned)
struct PtsSigned {
    const float *px, *py, *pr;
    float sx, sy;
    float x(int k) const { return sx * px[k]; }
    float y(int k) const { return sy * py[k]; }
    float r(int k) const { return pr[k]; }
};

static double wrap_unit(double v) {
    double t = v - (-0.5);
    t = t - std::trunc(t);
    t = t + 1.0;
    t = t - std::trunc(t);
    return t + (-0.5);
}

struct Stats {
    unsigned long long scored = 0, cands = 0, evals = 0, clear = 0, hit = 0, ambiguous = 0, band = 0;
    unsigned long long exact_moves = 0, overlap_moves = 0, exact_cands = 0, max_cands = 0;
    unsigned long long hist[16] = {};
    unsigned long long errors = 0;
    unsigned long long signed_checked = 0;
};

struct Cand {
    int i, j, ix, iy;
    bool incell;
    bool band;        // candidacy itself is within rounding of the 2R filter
    double X, Y;      // image position of j (f64, reference expressions)
};

template <class PA, class PB>
static float run_filter(int kind, int n, bool chain, const PA &A, const PB &B, float dX, float dY, float tau,
                        unsigned (&amb)[CPF_MAX_WORDS]) {
    for (unsigned &w : amb) w = 0u;
#define CHAIN_CASE(NN)                                                   \
    case NN: {                                                           \
        unsigned m[cpf_words(NN)];                                       \
        const float c = cpf_chain<NN>(A, B, dX, dY, tau, m);             \
        for (int w = 0; w < cpf_words(NN); ++w) amb[w] = m[w];           \
        return c;                                                        \
    }
    if (kind == ORC_POLYGON) {
        if (chain) {
            switch (n) {
                CHAIN_CASE(3) CHAIN_CASE(4) CHAIN_CASE(5) CHAIN_CASE(6) CHAIN_CASE(7) CHAIN_CASE(8)
                CHAIN_CASE(9) CHAIN_CASE(10) CHAIN_CASE(11) CHAIN_CASE(12)
            default: break;
            }
        }
        return cpf_segments(n, A, B, dX, dY, tau, amb);
    }
    return cpf_discs<0>(n, A, B, dX, dY, tau, amb);
}

// the verdict for state s the way the kernel assembles it; checks every certificate on the way
static int filtered_check(const orc_state *s, bool chain, Stats &st) {
    const int N = s->n_syms, n = s->n_items, kind = s->kind;
    const double R = orc_shape_enclosing_radius(kind, s->items, n);
    const double r2x = R * 2.0, radius_sq = r2x * r2x;
    const double a = s->length, b = s->length * s->ratio;
    double sa, ca, stt, ctt;
    orc_sincos(s->math_mode, s->angle, &sa, &ca);
    orc_sincos(s->math_mode, s->theta, &stt, &ctt);
    int shells;
    {
        double p = a / b, off = std::fabs(s->angle - 3.14159265358979323846 / 2.0);
        shells = (0.5 < p && p < 2.0 && off < 0.2) ? 1 : ((0.3 < p && p < 3.0 && off < 0.5) ? 2 : 3);
    }
    orc_transform rel[ORC_MAX_SYMS], cart[ORC_MAX_SYMS];
    orc_state_relative_positions(s, rel);
    orc_state_cartesian_positions(s, cart);
    static thread_local orc_item shapes[ORC_MAX_SYMS][ORC_MAX_ITEMS];
    for (int i = 0; i < N; ++i) orc_shape_transform(kind, s->items, n, &cart[i], shapes[i]);
    // kernel-side geometry: positions in f64 (reference expressions), points in f32
    double fx[8], fy[8], cx[8], cy[8];
    float vx[8][2 * ORC_MAX_ITEMS], vy[8][2 * ORC_MAX_ITEMS], vr[8][2 * ORC_MAX_ITEMS];
    const float ctf = (float)ctt, stf = (float)stt;
    for (int k = 0; k < N; ++k) {
        const double (*S)[3] = s->syms[k].m;
        double tx = ((S[0][0] * s->x) + S[0][1] * s->y) + S[0][2] * 1.0;
        double ty = ((S[1][0] * s->x) + S[1][1] * s->y) + S[1][2] * 1.0;
        fx[k] = wrap_unit(tx);
        fy[k] = wrap_unit(ty);
        double yb = fy[k] * b;
        cx[k] = fx[k] * a + yb * ca;
        cy[k] = yb * sa;
        if (fx[k] != rel[k].m[0][2] || fy[k] != rel[k].m[1][2] || cx[k] != cart[k].m[0][2] || cy[k] != cart[k].m[1][2]) {
            st.errors++;
            std::fprintf(stderr, "position mismatch sym %d\n", k);
        }
        const float s00 = (float)S[0][0], s01 = (float)S[0][1], s10 = (float)S[1][0], s11 = (float)S[1][1];
        const float m00 = fmaf(s00, ctf, s01 * stf), m01 = fmaf(s01, ctf, -(s00 * stf));
        const float m10 = fmaf(s10, ctf, s11 * stf), m11 = fmaf(s11, ctf, -(s10 * stf));
        for (int q = 0; q < n; ++q) {
            const orc_item &it = s->items[q];
            if (kind == ORC_POLYGON) {
                if (chain) {
                    vx[k][q] = fmaf(m00, (float)it.a, m01 * (float)it.b);
                    vy[k][q] = fmaf(m10, (float)it.a, m11 * (float)it.b);
                } else {
                    vx[k][2 * q] = fmaf(m00, (float)it.a, m01 * (float)it.b);
                    vy[k][2 * q] = fmaf(m10, (float)it.a, m11 * (float)it.b);
                    vx[k][2 * q + 1] = fmaf(m00, (float)it.c, m01 * (float)it.d);
                    vy[k][2 * q + 1] = fmaf(m10, (float)it.c, m11 * (float)it.d);
                }
            } else {
                vx[k][q] = fmaf(m00, (float)it.a, m01 * (float)it.b);
                vy[k][q] = fmaf(m10, (float)it.a, m11 * (float)it.b);
                vr[k][q] = (float)it.c;
            }
        }
    }
    // signed images (cp_api.cu: make_kproblem): every rotation part = operator 0's with row signs
    bool signed_images = true;
    float sgn[8][2];
    for (int k = 0; k < N; ++k) {
        const double (*S)[3] = s->syms[k].m, (*S0)[3] = s->syms[0].m;
        bool found = false;
        for (int q = 0; q < 4 && !found; ++q) {
            const float ax = (q & 1) ? -1.0f : 1.0f, ay = (q & 2) ? -1.0f : 1.0f;
            if ((float)S[0][0] == ax * (float)S0[0][0] && (float)S[0][1] == ax * (float)S0[0][1] &&
                (float)S[1][0] == ay * (float)S0[1][0] && (float)S[1][1] == ay * (float)S0[1][1]) {
                sgn[k][0] = ax;
                sgn[k][1] = ay;
                found = true;
            }
        }
        if (!found) signed_images = false;
    }
    // candidate list exactly as the kernel builds it (cp_thread.cuh: t_sweep): the single-precision
    // sweep over half of the ordered pairs; `band` = the certificate's evaluation site cannot be sure
    // the reference keeps the image (the f64 displacement, rounded, against narrow2)
    std::vector<Cand> cl;
    const CpfSweep sw = cpf_sweep_setup((float)a, (float)(b * ca), (float)(b * sa), (float)R, shells);
    auto push_word = [&](unsigned long long bits, int i, int j) {
        while (bits) {
            const int q = __builtin_ctzll(bits);
            bits &= bits - 1;
            const int row = q / 7, iy = row - 3, ix = q - row * 7 - 3;
            const double ffx = fx[j] + (double)ix, ffy = fy[j] + (double)iy;
            const double yb = ffy * b, X = ffx * a + yb * ca, Y = yb * sa;
            const bool incell = ix == 0 && iy == 0;
            const float dX = (float)(X - cx[i]), dY = (float)(Y - cy[i]);
            const bool band = !incell && !(fmaf(dX, dX, dY * dY) <= sw.narrow2);
            cl.push_back(Cand{i, j, ix, iy, incell, band, X, Y});
        }
    };
    std::vector<unsigned long long> words;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            const float d0x = (float)(cx[j] - cx[i]), d0y = (float)(cy[j] - cy[i]);
            const unsigned long long bits = cpf_sweep_pair<false>(sw, d0x, d0y) | (1ull << 24);
            words.push_back(bits);
            push_word(bits, i, j);
            // the reach-limited loops (multiplicity-4 builds) must return the same mask
            float best = 3.0e38f;
            int bq = -1;
            if ((cpf_sweep_pair_reach<false>(sw, d0x, d0y, best, bq) | (1ull << 24)) != bits) {
                st.errors++;
                std::fprintf(stderr, "reach sweep differs: pair (%d, %d), d0 = (%a, %a), a %a bx %a by %a shells %d\n", i, j,
                             d0x, d0y, sw.af, sw.bxf, sw.byf, sw.shells);
            }
        }
    const unsigned long long self = cpf_sweep_pair<true>(sw, 0.0f, 0.0f);
    {
        float best = 3.0e38f;
        int bq = -1;
        if (cpf_sweep_pair_reach<true>(sw, 0.0f, 0.0f, best, bq) != self) {
            st.errors++;
            std::fprintf(stderr, "reach sweep differs: self word, a %a bx %a by %a shells %d\n", sw.af, sw.bxf, sw.byf,
                         sw.shells);
        }
    }
    for (int i = 0; i < N; ++i) push_word(self, i, i);
    // the sweep must not miss anything the reference tests: every ordered pair, every image within
    // `shells` that passes the f64 filter (packed.rs:152-157) is in the list, itself or as its mirror
    {
        int w = 0;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                for (int ix = -shells; ix <= shells; ++ix)
                    for (int iy = -shells; iy <= shells; ++iy) {
                        if (ix == 0 && iy == 0) continue;
                        const double ffx = fx[j] + (double)ix, ffy = fy[j] + (double)iy;
                        const double yb = ffy * b, X = ffx * a + yb * ca, Y = yb * sa;
                        const double ddx = cx[i] - X, ddy = cy[i] - Y;
                        if (!(ddx * ddx + ddy * ddy <= radius_sq)) continue;
                        bool listed;
                        if (i < j) {
                            int wi = 0;
                            for (int p = 0; p < i; ++p) wi += N - 1 - p;
                            wi += j - i - 1;
                            listed = (words[wi] >> ((iy + 3) * 7 + (ix + 3))) & 1ull;
                        } else if (i > j) {
                            int wi = 0;
                            for (int p = 0; p < j; ++p) wi += N - 1 - p;
                            wi += i - j - 1;
                            listed = (words[wi] >> ((-iy + 3) * 7 + (-ix + 3))) & 1ull;
                        } else {
                            const bool upper = iy > 0 || (iy == 0 && ix > 0);
                            const int qx = upper ? ix : -ix, qy = upper ? iy : -iy;
                            listed = (self >> ((qy + 3) * 7 + (qx + 3))) & 1ull;
                        }
                        if (!listed) {
                            st.errors++;
                            std::fprintf(stderr, "SWEEP MISSED a reference candidate: (%d,%d) image (%d,%d)\n", i, j, ix, iy);
                        }
                    }
        (void)w;
    }
    st.scored++;
    st.cands += cl.size();
    if (cl.size() > st.max_cands) st.max_cands = cl.size();
    st.hist[cl.size() < 15 ? cl.size() : 15]++;

    // f64 truth table of one side: component k of shape `i` (self) against component l of image
    // (ix, iy) of shape `j` (other), reference expressions; *is_cand = the side passes the 2R filter
    auto exact_side = [&](int i, int j, int ix, int iy, bool incell, bool *is_cand, unsigned char *table) {
        orc_transform img = rel[j];
        double X, Y;
        orc_cell_to_cartesian(s, rel[j].m[0][2] + (double)ix, rel[j].m[1][2] + (double)iy, &X, &Y);
        img.m[0][2] = X;
        img.m[1][2] = Y;
        double dx = cart[i].m[0][2] - X, dy = cart[i].m[1][2] - Y;
        *is_cand = incell || dx * dx + dy * dy <= radius_sq;
        orc_item sh2[ORC_MAX_ITEMS];
        orc_shape_transform(kind, s->items, n, &img, sh2);
        for (int k = 0; k < n; ++k)
            for (int l = 0; l < n; ++l)
                table[k * n + l] = (unsigned char)(kind == ORC_POLYGON ? orc_line_intersects(&shapes[i][k], &sh2[l])
                                                                       : orc_atom_intersects(&shapes[i][k], &sh2[l]));
    };

    bool overlap = false, need_exact = false, exact_hit = false;
    const float Rf = (float)R;
    static thread_local unsigned char tA[ORC_MAX_ITEMS * ORC_MAX_ITEMS], tB[ORC_MAX_ITEMS * ORC_MAX_ITEMS];
    for (size_t c = 0; c < cl.size(); ++c) {
        const Cand &C = cl[c];
        const float dX = (float)(C.X - cx[C.i]), dY = (float)(C.Y - cy[C.i]);
        Pts A{vx[C.i], vy[C.i], vr[C.i]}, B{vx[C.j], vy[C.j], vr[C.j]};
        const float tau = kind == ORC_POLYGON ? cpf_tau(Rf, dX, dY) : cpf_tau_discs(Rf, dX, dY);
        unsigned amb[CPF_MAX_WORDS];
        const float cmin = run_filter(kind, n, chain, A, B, dX, dY, tau, amb);
        st.evals++;
        if (signed_images) {
            // the multiplicity-4 builds evaluate the certificate from image 0's points alone, in shape
            // i's frame with its row signs undone (cp_thread.cuh: t_overlap_warp): same cmin, same
            // tau, same ambiguous set, bit for bit
            const float sAx = sgn[C.i][0], sAy = sgn[C.i][1];
            Pts A0{vx[0], vy[0], vr[0]};
            PtsSigned B0{vx[0], vy[0], vr[0], sAx * sgn[C.j][0], sAy * sgn[C.j][1]};
            const float dX2 = dX * sAx, dY2 = dY * sAy;
            const float tau2 = kind == ORC_POLYGON ? cpf_tau(Rf, dX2, dY2) : cpf_tau_discs(Rf, dX2, dY2);
            unsigned amb2[CPF_MAX_WORDS];
            const float cmin2 = run_filter(kind, n, chain, A0, B0, dX2, dY2, tau2, amb2);
            bool same = std::memcmp(&cmin, &cmin2, sizeof cmin) == 0 && std::memcmp(&tau, &tau2, sizeof tau) == 0;
            for (int w = 0; w < CPF_MAX_WORDS; ++w) same = same && amb[w] == amb2[w];
            if (!same) {
                st.errors++;
                std::fprintf(stderr, "signed-image certificate differs: pair (%d,%d) cmin %a vs %a tau %a vs %a\n", C.i, C.j,
                             cmin, cmin2, tau, tau2);
            }
            st.signed_checked++;
        }
        bool candA, candB = false;
        exact_side(C.i, C.j, C.ix, C.iy, C.incell, &candA, tA);
        if (!C.incell) exact_side(C.j, C.i, -C.ix, -C.iy, false, &candB, tB);
        // pairs outside the ambiguity mask are certified false on both sides
        int n_amb = 0;
        bool anyA = false, anyB = false;
        for (int l = 0; l < n; ++l)
            for (int k = 0; k < n; ++k) {
                const int p = l * n + k;
                const bool in_mask = (amb[p >> 5] >> (p & 31)) & 1u;
                const bool eA = tA[k * n + l], eB = !C.incell && tB[l * n + k];
                anyA |= eA;
                anyB |= eB;
                if (in_mask) ++n_amb;
                if (!in_mask && (eA || eB)) {
                    st.errors++;
                    std::fprintf(stderr, "UNSOUND clear pair (%d,%d): cmin %g tau %g eA %d eB %d\n", k, l, cmin, tau, eA, eB);
                }
            }
        if (cmin < -tau && !C.band) {
            st.hit++;
            if (!anyA || (!C.incell && !anyB) || !candA) {
                st.errors++;
                std::fprintf(stderr, "UNSOUND hit: cmin %g tau %g anyA %d anyB %d cand %d %d\n", cmin, tau, anyA, anyB, candA, candB);
            }
            overlap = true;
            break;  // the kernel stops at the first certified hit
        }
        if (n_amb == 0 && !C.band) {
            st.clear++;
            continue;
        }
        st.ambiguous++;
        if (C.band) st.band++;
        need_exact = true;
        // the exact stage: the masked pairs (all pairs for a band candidate), each side only if it
        // passes the reference's distance filter
        for (int l = 0; l < n; ++l)
            for (int k = 0; k < n; ++k) {
                const int p = l * n + k;
                if (!C.band && !((amb[p >> 5] >> (p & 31)) & 1u)) continue;
                st.exact_cands++;
                if ((candA && tA[k * n + l]) || (candB && tB[l * n + k])) exact_hit = true;
            }
    }
    if (exact_hit) overlap = true;
    if (need_exact) st.exact_moves++;
    if (overlap) st.overlap_moves++;
    return overlap ? 1 : 0;
}

// `filter_check sweep <cases>`: the reach-limited image sweep against the straight-line one on random
// cells, extreme ones included (lengths down to the lower bound 0.01, ratios from 0.1 to 100, every
// cell angle, displacements anywhere in the cell, enclosing radii from 0.05 to 5).
static int sweep_stress(unsigned long long cases) {
    unsigned long long x = 0x9E3779B97F4A7C15ull, errors = 0, kept = 0;
    auto rnd = [&]() {  // xorshift64*, uniform in [0, 1)
        x ^= x >> 12;
        x ^= x << 25;
        x ^= x >> 27;
        return (double)((x * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0);
    };
    for (unsigned long long c = 0; c < cases; ++c) {
        const double a = rnd() < 0.2 ? 0.01 + rnd() * 0.1 : std::exp(std::log(0.01) + rnd() * std::log(2000.0));
        const double ratio = rnd() < 0.8 ? 0.1 + 0.9 * rnd() : std::exp(std::log(0.1) + rnd() * std::log(1000.0));
        const double angle = M_PI / 4 + rnd() * M_PI / 4, b = a * ratio;
        const double R = rnd() < 0.5 ? 1.0 : 0.05 + rnd() * 4.95;
        const int shells = 1 + (int)(rnd() * 3.0);
        const CpfSweep sw = cpf_sweep_setup((float)a, (float)(b * std::cos(angle)), (float)(b * std::sin(angle)), (float)R, shells);
        const double fx = rnd() * 2.0 - 1.0, fy = rnd() * 2.0 - 1.0;  // difference of two wrapped positions
        const float d0x = (float)(fx * a + fy * b * std::cos(angle)), d0y = (float)(fy * b * std::sin(angle));
        float best = 3.0e38f;
        int q = -1;
        const unsigned long long m0 = cpf_sweep_pair<false>(sw, d0x, d0y), m1 = cpf_sweep_pair_reach<false>(sw, d0x, d0y, best, q);
        best = 3.0e38f;
        const unsigned long long s0 = cpf_sweep_pair<true>(sw, 0.0f, 0.0f), s1 = cpf_sweep_pair_reach<true>(sw, 0.0f, 0.0f, best, q);
        kept += __builtin_popcountll(m0);
        if (m0 != m1 || s0 != s1) {
            if (errors++ < 5)
                std::fprintf(stderr, "case %llu: a %g ratio %g angle %g R %g shells %d d0 (%g, %g): %llx vs %llx, self %llx vs %llx\n",
                             c, a, ratio, angle, R, shells, d0x, d0y, m0, m1, s0, s1);
        }
    }
    std::printf("{\"cases\": %llu, \"errors\": %llu, \"kept\": %llu}\n", cases, errors, kept);
    return errors ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc == 3 && !std::strcmp(argv[1], "sweep")) return sweep_stress(std::strtoull(argv[2], nullptr, 10));
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s polygon|trimer|circle|star sides group replicas steps [kt_start kt_finish]\n", argv[0]);
        return 2;
    }
    const char *shape = argv[1];
    const int sides = std::atoi(argv[2]);
    const char *group = argv[3];
    const int replicas = std::atoi(argv[4]);
    const unsigned long long steps = std::strtoull(argv[5], nullptr, 10);
    const double kt_start = argc > 6 ? std::atof(argv[6]) : 0.1;
    const double kt_finish = argc > 7 ? std::atof(argv[7]) : NAN;
    orc_item items[ORC_MAX_ITEMS];
    int kind = ORC_POLYGON, n = sides;
    bool chain = true;
    if (!std::strcmp(shape, "polygon")) {
        double radii[ORC_MAX_ITEMS];
        for (int i = 0; i < sides; ++i) radii[i] = 1.0;
        orc_shape_from_radial(radii, sides, items);
    } else if (!std::strcmp(shape, "star")) {
        double radii[ORC_MAX_ITEMS];
        for (int i = 0; i < sides; ++i) radii[i] = (i & 1) ? 0.5 : 1.0;
        orc_shape_from_radial(radii, sides, items);
    } else if (!std::strcmp(shape, "soup")) {  // same polygon, treated as unrelated segments
        double radii[ORC_MAX_ITEMS];
        for (int i = 0; i < sides; ++i) radii[i] = 1.0;
        orc_shape_from_radial(radii, sides, items);
        chain = false;
    } else if (!std::strcmp(shape, "trimer")) {
        kind = ORC_DISCS;
        n = 3;
        orc_shape_trimer(ORC_DISCS, 0.637556, 120.0, 1.0, items);
    } else {
        kind = ORC_DISCS;
        n = 1;
        orc_shape_circle(ORC_DISCS, items);
    }
    orc_state tmpl;
    if (orc_state_from_group(&tmpl, kind, items, n, group, ORC_MATH_CP)) {
        std::fprintf(stderr, "bad group\n");
        return 2;
    }
    orc_build_optimiser bo{steps, kt_start, kt_finish, NAN, 0.01, 1000, NAN};
    Stats st;
    std::vector<orc_move_record> rec(steps + 1000);
    unsigned long long moves = 0, inb = 0;
    for (int r = 0; r < replicas; ++r) {
        orc_state s = tmpl;
        for (int stage = 0; stage < 3; ++stage) {
            orc_build_optimiser b2 = bo;
            if (stage == 0) { b2.steps = 100; b2.kt_start = 0.0; }
            if (stage == 2) b2.kt_start = 0.0;
            orc_optimiser opt;
            orc_build(&b2, (uint64_t)r, &opt);
            if (stage == 0) opt.convergence = NAN;
            orc_state before = s;
            uint64_t nrec = 0, rej = 0;
            if (orc_optimise_state(&opt, &s, rec.data(), rec.size(), &nrec, &rej)) {
                std::fprintf(stderr, "optimise failed\n");
                return 1;
            }
            // walk the recorded moves again, checking the filtered verdict at every scored proposal
            orc_state w = before;
            double dof[6];
            const int n_cell = w.family == ORC_MONOCLINIC ? 3 : (w.family == ORC_ORTHORHOMBIC ? 2 : 1);
            for (uint64_t m = 0; m < nrec; ++m) {
                const orc_move_record &M = rec[m];
                ++moves;
                if (!M.in_bounds) continue;
                ++inb;
                orc_state_get_dof(&w, dof);
                const int slot = (int)M.basis_index < n_cell ? (int)M.basis_index : (int)M.basis_index - n_cell + 3;
                const double old = dof[slot];
                dof[slot] = M.new_value;
                orc_state_set_dof(&w, dof);
                const int verdict = filtered_check(&w, chain, st);
                const int truth = !(M.new_score == M.new_score);
                if (verdict != truth) {
                    st.errors++;
                    std::fprintf(stderr, "VERDICT MISMATCH replica %d stage %d move %llu: filtered %d oracle %d\n", r, stage,
                                 (unsigned long long)m, verdict, truth);
                }
                if (!M.accepted) {
                    dof[slot] = old;
                    orc_state_set_dof(&w, dof);
                }
            }
        }
    }
    std::printf("{\"shape\": \"%s\", \"sides\": %d, \"group\": \"%s\", \"replicas\": %d, \"moves\": %llu, \"in_bounds\": %llu, "
                "\"scored\": %llu, \"cands_per_scored\": %.3f, \"max_cands\": %llu, \"evals_per_scored\": %.3f, "
                "\"clear\": %llu, \"hit\": %llu, \"ambiguous\": %llu, \"band\": %llu, \"exact_moves_frac\": %.5f, "
                "\"exact_cands_per_scored\": %.5f, \"overlap_frac\": %.4f, \"signed_checked\": %llu, \"errors\": %llu}\n",
                shape, sides, group, replicas, moves, inb, st.scored, (double)st.cands / st.scored, st.max_cands,
                (double)st.evals / st.scored, st.clear, st.hit, st.ambiguous, st.band,
                (double)st.exact_moves / st.scored, (double)st.exact_cands / st.scored,
                (double)st.overlap_moves / st.scored, st.signed_checked, st.errors);
    std::printf("cands histogram:");
    for (int i = 0; i < 16; ++i) std::printf(" %llu", st.hist[i]);
    std::printf("\n");
    return st.errors ? 1 : 0;
}
