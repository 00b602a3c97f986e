// cp_launch.cu — routes a launch to the translation unit that holds the kernels of its shape kind.
#include "cp_launch.h"
#include "cp_thread.cuh"

// component count the one-thread-per-replica build that serves this problem is compiled for
// (0 = the run-time-count build).  Polygon builds with a compile-time count read the shape as a
// closed chain of n points and exist for the multiplicities 1, 2 and 4.
int cp_thread_route_ni(const KProblem &P) {
    const bool ns_ok = P.n_syms == 1 || P.n_syms == 2 || P.n_syms == 4;
    switch (P.kind) {
    case CP_POLYGON: return (P.chain && ns_ok && P.n_items >= 3 && P.n_items <= 12) ? P.n_items : 0;
    case CP_DISCS: return (ns_ok && (P.n_items == 1 || P.n_items == 3)) ? P.n_items : 0;
    default: return (P.n_items == 1 || P.n_items == 3) ? P.n_items : 0;
    }
}

size_t cp_thread_shared_bytes(const KProblem &P) {
    if (P.kind == CP_LJ) return thread_block_shared_bytes<CP_LJ, 0, 0>(P);
    if (P.kind == CP_POLYGON && cp_thread_route_ni(P) == 0)
        return thread_block_shared_bytes<CP_POLYGON, 0, 0>(P);
    return thread_block_shared_bytes<CP_DISCS, 0, 0>(P);  // one point per component
}

bool cp_thread_kernel_fits(const KProblem &P) { return cp_thread_shared_bytes(P) <= 200 * 1024; }

// one thread per replica: pick the build specialised for this component count, if there is one
#define CP_THREAD_ROUTE(WHAT, n_items, ...)                                                      \
    switch (kind) {                                                                              \
    case CP_POLYGON:                                                                             \
        switch (n_items) {                                                                       \
        case 3: return cp_thread_polygon_##WHAT##_3(__VA_ARGS__);                                \
        case 4: return cp_thread_polygon_##WHAT##_4(__VA_ARGS__);                                \
        case 5: return cp_thread_polygon_##WHAT##_5(__VA_ARGS__);                                \
        case 6: return cp_thread_polygon_##WHAT##_6(__VA_ARGS__);                                \
        case 7: return cp_thread_polygon_##WHAT##_7(__VA_ARGS__);                                \
        case 8: return cp_thread_polygon_##WHAT##_8(__VA_ARGS__);                                \
        case 9: return cp_thread_polygon_##WHAT##_9(__VA_ARGS__);                                \
        case 10: return cp_thread_polygon_##WHAT##_10(__VA_ARGS__);                              \
        case 11: return cp_thread_polygon_##WHAT##_11(__VA_ARGS__);                              \
        case 12: return cp_thread_polygon_##WHAT##_12(__VA_ARGS__);                              \
        default: return cp_thread_polygon_##WHAT##_0(__VA_ARGS__);                               \
        }                                                                                        \
    case CP_DISCS:                                                                               \
        switch (n_items) {                                                                       \
        case 1: return cp_thread_discs_##WHAT##_1(__VA_ARGS__);                                  \
        case 3: return cp_thread_discs_##WHAT##_3(__VA_ARGS__);                                  \
        default: return cp_thread_discs_##WHAT##_0(__VA_ARGS__);                                 \
        }                                                                                        \
    case CP_LJ:                                                                                  \
        switch (n_items) {                                                                       \
        case 1: return cp_thread_lj_##WHAT##_1(__VA_ARGS__);                                     \
        case 3: return cp_thread_lj_##WHAT##_3(__VA_ARGS__);                                     \
        default: return cp_thread_lj_##WHAT##_0(__VA_ARGS__);                                    \
        }                                                                                        \
    default: return cudaErrorInvalidValue;                                                       \
    }

static cudaError_t thread_route_mc(const KParams &K, cudaStream_t st) {
    const int kind = K.prob.kind;
    CP_THREAD_ROUTE(mc, cp_thread_route_ni(K.prob), K, st)
}
static cudaError_t thread_route_score(const KProblem &P, const double *dof, unsigned long long n, double *out,
                                      cudaStream_t st) {
    const int kind = P.kind;
    CP_THREAD_ROUTE(score, cp_thread_route_ni(P), P, dof, n, out, st)
}
static cudaError_t thread_route_replay(const KProblem &P, const double *init, const cp_move *mv,
                                       unsigned long long n_moves, unsigned long long n, cp_decision *out,
                                       cudaStream_t st) {
    const int kind = P.kind;
    CP_THREAD_ROUTE(replay, cp_thread_route_ni(P), P, init, mv, n_moves, n, out, st)
}

cudaError_t cp_launch_mc(const KParams &K, int lanes, cudaStream_t st) {
    if (lanes == 1) return thread_route_mc(K, st);
    switch (K.prob.kind) {
    case CP_POLYGON: return cp_group_polygon_mc(K, lanes, st);
    case CP_DISCS: return cp_group_discs_mc(K, lanes, st);
    case CP_LJ: return cp_group_lj_mc(K, lanes, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t cp_launch_score(const KProblem &P, int lanes, const double *dof, unsigned long long n, double *out,
                            cudaStream_t st) {
    if (lanes == 1) return thread_route_score(P, dof, n, out, st);
    switch (P.kind) {
    case CP_POLYGON: return cp_group_polygon_score(P, lanes, dof, n, out, st);
    case CP_DISCS: return cp_group_discs_score(P, lanes, dof, n, out, st);
    case CP_LJ: return cp_group_lj_score(P, lanes, dof, n, out, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t cp_launch_replay(const KProblem &P, int lanes, const double *init, const cp_move *mv,
                             unsigned long long n_moves, unsigned long long n, cp_decision *out, cudaStream_t st) {
    if (lanes == 1) return thread_route_replay(P, init, mv, n_moves, n, out, st);
    switch (P.kind) {
    case CP_POLYGON: return cp_group_polygon_replay(P, lanes, init, mv, n_moves, n, out, st);
    case CP_DISCS: return cp_group_discs_replay(P, lanes, init, mv, n_moves, n, out, st);
    case CP_LJ: return cp_group_lj_replay(P, lanes, init, mv, n_moves, n, out, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t cp_launch_images(const KProblem &P, const double *dof, unsigned long long n, double *out, cudaStream_t st) {
    images_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, dof, n, out);
    return cudaGetLastError();
}
