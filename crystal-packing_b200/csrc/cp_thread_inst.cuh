// cp_thread_inst.cuh — instantiates the one-thread-per-replica kernels of one shape kind for a
// list of compile-time component counts (0 = runtime count) and defines its launch entry points.
#include "cp_launch.h"
#include "cp_thread.cuh"

template <int KIND, int NI, int NS, int RNG>
static cudaError_t thread_mc_launch(const KParams &K, cudaStream_t st) {
    const size_t bytes = thread_block_shared_bytes<KIND, NI, NS>(K.prob);
    const unsigned grid = (unsigned)((K.n_replicas + kThreadBlock - 1) / kThreadBlock);
    if constexpr (KIND == CP_LJ) {
        cudaError_t e = cp_allow_shared(mc_thread_lj_kernel<KIND, NI, NS, RNG>, bytes);
        if (e != cudaSuccess) return e;
        mc_thread_lj_kernel<KIND, NI, NS, RNG><<<grid, kThreadBlock, bytes, st>>>(K);
    } else {
        cudaError_t e = cp_allow_shared(mc_thread_kernel<KIND, NI, NS, RNG>, bytes);
        if (e != cudaSuccess) return e;
        mc_thread_kernel<KIND, NI, NS, RNG><<<grid, kThreadBlock, bytes, st>>>(K);
    }
    return cudaGetLastError();
}

// Builds with a compile-time component count exist for multiplicities 1, 2 and 4 (every plane group
// the reference defines); the run-time-count build (NI = 0) takes the multiplicity at run time too.
// LJ kernels take the multiplicity at run time.
#define CP_NS_SWITCH(N, CALL)                                         \
    if constexpr (KIND == CP_LJ || NI == 0) {                         \
        constexpr int NS = 0;                                         \
        return CALL;                                                  \
    } else {                                                          \
        switch (N) {                                                  \
        case 1: { constexpr int NS = 1; return CALL; }                \
        case 2: { constexpr int NS = 2; return CALL; }                \
        case 4: { constexpr int NS = 4; return CALL; }                \
        default: return cudaErrorInvalidValue;                        \
        }                                                             \
    }

template <int KIND, int NI>
static cudaError_t thread_mc(const KParams &K, cudaStream_t st) {
    if (K.rng_mode == CP_RNG_PCG64MCG) {
        CP_NS_SWITCH(K.prob.n_syms, (thread_mc_launch<KIND, NI, NS, CP_RNG_PCG64MCG>(K, st)))
    }
    CP_NS_SWITCH(K.prob.n_syms, (thread_mc_launch<KIND, NI, NS, CP_RNG_PHILOX>(K, st)))
}

template <int KIND, int NI, int NS>
static cudaError_t thread_score_launch(const KProblem &P, const double *dof, unsigned long long n, double *out,
                                       cudaStream_t st) {
    const size_t bytes = thread_block_shared_bytes<KIND, NI, NS>(P);
    cudaError_t e = cp_allow_shared(score_thread_kernel<KIND, NI, NS>, bytes);
    if (e != cudaSuccess) return e;
    score_thread_kernel<KIND, NI, NS><<<(unsigned)((n + kThreadBlock - 1) / kThreadBlock), kThreadBlock, bytes, st>>>(
        P, dof, n, out);
    return cudaGetLastError();
}
template <int KIND, int NI>
static cudaError_t thread_score(const KProblem &P, const double *dof, unsigned long long n, double *out,
                                cudaStream_t st) {
    CP_NS_SWITCH(P.n_syms, (thread_score_launch<KIND, NI, NS>(P, dof, n, out, st)))
}

template <int KIND, int NI, int NS>
static cudaError_t thread_replay_launch(const KProblem &P, const double *init, const cp_move *mv,
                                        unsigned long long n_moves, unsigned long long n, cp_decision *out,
                                        cudaStream_t st) {
    const size_t bytes = thread_block_shared_bytes<KIND, NI, NS>(P);
    cudaError_t e = cp_allow_shared(replay_thread_kernel<KIND, NI, NS>, bytes);
    if (e != cudaSuccess) return e;
    replay_thread_kernel<KIND, NI, NS><<<(unsigned)((n + kThreadBlock - 1) / kThreadBlock), kThreadBlock, bytes, st>>>(
        P, init, mv, n_moves, n, out);
    return cudaGetLastError();
}
template <int KIND, int NI>
static cudaError_t thread_replay(const KProblem &P, const double *init, const cp_move *mv, unsigned long long n_moves,
                                 unsigned long long n, cp_decision *out, cudaStream_t st) {
    CP_NS_SWITCH(P.n_syms, (thread_replay_launch<KIND, NI, NS>(P, init, mv, n_moves, n, out, st)))
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
