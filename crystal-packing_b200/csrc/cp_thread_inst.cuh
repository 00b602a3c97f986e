// cp_thread_inst.cuh — instantiates the one-thread-per-replica kernels of one shape kind for a
// list of compile-time component counts (0 = runtime count) and defines its launch entry points.
#include "cp_launch.h"
#include "cp_thread.cuh"

template <int KIND, int NI, int NS, int RNG>
static cudaError_t thread_mc_launch(const KParams &K0, cudaStream_t st) {
    KParams K = K0;
    const size_t bytes = thread_block_shared_bytes<KIND, NI, NS>(K.prob);
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cp_allow_shared(kernel, bytes);
        if (e != cudaSuccess) return e;
        if (K.warp_fill < 1 || K.warp_fill > kThreadBlock) {
            // Replicas per warp.  A batch that fits the GPU's warp slots at once is spread over a whole
            // number of warps per scheduler (4 per SM): every scheduler carries the same load - 65,536
            // replicas are 3.46 warps of 32 per scheduler, or 4 warps of 28 on each - and a small batch
            // runs in more, shorter warps (the pool rounds of a warp scale with its replicas).  Never
            // below 8 replicas per warp, where the pooled tests run out of lanes; batches larger than
            // the slots run full warps in several waves.
            int per_sm = 0, dev = 0, sms = 0;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreadBlock, bytes)) != cudaSuccess) return e;
            if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
            if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
            const unsigned long long slots = (unsigned long long)(per_sm > 0 ? per_sm : 1) * (unsigned long long)sms;
            const unsigned long long full = (K.n_replicas + kThreadBlock - 1) / kThreadBlock, sched = 4ull * (unsigned long long)sms;
            K.warp_fill = kThreadBlock;
            if (full <= slots) {
                unsigned long long warps = (full + sched - 1) / sched * sched;
                if (warps > slots) warps = slots;
                const unsigned long long fill = (K.n_replicas + warps - 1) / warps;
                K.warp_fill = fill >= (unsigned long long)kThreadBlock ? kThreadBlock : (fill < 8 ? 8 : (int)fill);
            }
        }
        const unsigned grid = (unsigned)((K.n_replicas + K.warp_fill - 1) / K.warp_fill);
        kernel<<<grid, kThreadBlock, bytes, st>>>(K);
        return cudaGetLastError();
    };
    if constexpr (KIND == CP_LJ)
        return launch(mc_thread_lj_kernel<KIND, NI, NS, RNG>);
    else
        return launch(mc_thread_kernel<KIND, NI, NS, RNG>);
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
