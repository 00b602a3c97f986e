// cp_thread_inst.cuh — instantiates the one-thread-per-replica kernels of one shape kind for a
// list of compile-time component counts (0 = runtime count) and defines its launch entry points.
#include "cp_launch.h"
#include "cp_thread.cuh"

static inline size_t thread_shared_bytes(const KProblem &P) {
    return thread_block_shared_bytes(P.kind, P.n_items, P.n_syms);
}

template <int KIND, int NI>
static cudaError_t thread_mc(const KParams &K, cudaStream_t st) {
    const size_t bytes = thread_shared_bytes(K.prob);
    const unsigned grid = (unsigned)((K.n_replicas + kThreadBlock - 1) / kThreadBlock);
    cudaError_t e;
    if (K.rng_mode == CP_RNG_PCG64MCG) {
        if ((e = cp_allow_shared(mc_thread_kernel<KIND, NI, CP_RNG_PCG64MCG>, bytes)) != cudaSuccess) return e;
        mc_thread_kernel<KIND, NI, CP_RNG_PCG64MCG><<<grid, kThreadBlock, bytes, st>>>(K);
    } else {
        if ((e = cp_allow_shared(mc_thread_kernel<KIND, NI, CP_RNG_PHILOX>, bytes)) != cudaSuccess) return e;
        mc_thread_kernel<KIND, NI, CP_RNG_PHILOX><<<grid, kThreadBlock, bytes, st>>>(K);
    }
    return cudaGetLastError();
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
