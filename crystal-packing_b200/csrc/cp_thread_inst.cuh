// cp_thread_inst.cuh — instantiates the one-thread-per-replica kernels of one shape kind for a
// list of compile-time component counts (0 = runtime count) and defines its launch entry points.
#include "cp_launch.h"
#include "cp_thread.cuh"

static inline size_t thread_shared_bytes(const KProblem &P) {
    return thread_block_shared_bytes(P.kind, P.n_items, P.n_syms);
}

template <int KIND, int NI, int RNG, bool WIDE>
static cudaError_t thread_mc_launch(const KParams &K, size_t bytes, cudaStream_t st) {
    const unsigned grid = (unsigned)((K.n_replicas + kThreadBlock - 1) / kThreadBlock);
    cudaError_t e = cp_allow_shared(mc_thread_kernel<KIND, NI, RNG, WIDE>, bytes);
    if (e != cudaSuccess) return e;
    mc_thread_kernel<KIND, NI, RNG, WIDE><<<grid, kThreadBlock, bytes, st>>>(K);
    return cudaGetLastError();
}

template <int KIND, int NI>
static cudaError_t thread_mc(const KParams &K, cudaStream_t st) {
    const size_t bytes = thread_shared_bytes(K.prob);
    // hexagons and smaller have a second build with a wider register budget for the cases where
    // shared memory (not registers) bounds the occupancy: fewer than 12 blocks of this size per SM
    constexpr bool kHasWide = KIND == CP_POLYGON && NI >= 3 && NI <= 6;
    const bool wide = kHasWide && bytes * 12 > 227 * 1024;
    if (K.rng_mode == CP_RNG_PCG64MCG) {
        if constexpr (kHasWide)
            if (wide) return thread_mc_launch<KIND, NI, CP_RNG_PCG64MCG, true>(K, bytes, st);
        return thread_mc_launch<KIND, NI, CP_RNG_PCG64MCG, false>(K, bytes, st);
    }
    if constexpr (kHasWide)
        if (wide) return thread_mc_launch<KIND, NI, CP_RNG_PHILOX, true>(K, bytes, st);
    return thread_mc_launch<KIND, NI, CP_RNG_PHILOX, false>(K, bytes, st);
}

template <int KIND, int NI>
static cudaError_t thread_score(const KProblem &P, const double *dof, unsigned long long n, double *out,
                                cudaStream_t st) {
    const size_t bytes = thread_shared_bytes(P);
    cudaError_t e = cp_allow_shared(score_thread_kernel<KIND, NI>, bytes);
    if (e != cudaSuccess) return e;
    score_thread_kernel<KIND, NI><<<(unsigned)((n + kThreadBlock - 1) / kThreadBlock), kThreadBlock, bytes, st>>>(
        P, dof, n, out);
    return cudaGetLastError();
}

template <int KIND, int NI>
static cudaError_t thread_replay(const KProblem &P, const double *init, const cp_move *mv, unsigned long long n_moves,
                                 unsigned long long n, cp_decision *out, cudaStream_t st) {
    const size_t bytes = thread_shared_bytes(P);
    cudaError_t e = cp_allow_shared(replay_thread_kernel<KIND, NI>, bytes);
    if (e != cudaSuccess) return e;
    replay_thread_kernel<KIND, NI><<<(unsigned)((n + kThreadBlock - 1) / kThreadBlock), kThreadBlock, bytes, st>>>(
        P, init, mv, n_moves, n, out);
    return cudaGetLastError();
}

// Defines the three entry points of one (kind, component count) pair.  Suffix = the count, or 0
// for the run-time-count build.
#define CP_DEFINE_THREAD_NI(KIND, name, NI)                                                                    \
    cudaError_t cp_##name##_mc_##NI(const KParams &K, cudaStream_t st) { return thread_mc<KIND, NI>(K, st); }   \
    cudaError_t cp_##name##_score_##NI(const KProblem &P, const double *dof, unsigned long long n, double *out, \
                                       cudaStream_t st) {                                                      \
        return thread_score<KIND, NI>(P, dof, n, out, st);                                                     \
    }                                                                                                          \
    cudaError_t cp_##name##_replay_##NI(const KProblem &P, const double *init, const cp_move *mv,               \
                                        unsigned long long n_moves, unsigned long long n, cp_decision *out,    \
                                        cudaStream_t st) {                                                     \
        return thread_replay<KIND, NI>(P, init, mv, n_moves, n, out, st);                                      \
    }
