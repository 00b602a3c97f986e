// cp_group_inst.cuh — instantiates the lane-group kernels of one shape kind and defines its launch
// entry points.  Included by cp_group_{polygon,discs,lj}.cu.
#include "cp_group.cuh"
#include "cp_launch.h"

#define CP_LANE_SWITCH(lanes, CALL)               \
    switch (lanes) {                              \
    case 4: { constexpr int L = 4; CALL; }        \
    case 8: { constexpr int L = 8; CALL; }        \
    case 16: { constexpr int L = 16; CALL; }      \
    case 32: { constexpr int L = 32; CALL; }      \
    default: return cudaErrorInvalidValue;        \
    }

template <int KIND, int L>
static cudaError_t group_mc(const KParams &K, cudaStream_t st) {
    const SharedLayout lay = shared_layout(KIND, K.prob.n_items, K.prob.n_syms, L);
    const unsigned long long gpb = kBlock / L;
    const unsigned grid = (unsigned)((K.n_replicas + gpb - 1) / gpb);
    cudaError_t e;
    if (K.rng_mode == CP_RNG_PCG64MCG) {
        if ((e = cp_allow_shared(mc_kernel<KIND, L, CP_RNG_PCG64MCG>, lay.bytes)) != cudaSuccess) return e;
        mc_kernel<KIND, L, CP_RNG_PCG64MCG><<<grid, kBlock, lay.bytes, st>>>(K);
    } else {
        if ((e = cp_allow_shared(mc_kernel<KIND, L, CP_RNG_PHILOX>, lay.bytes)) != cudaSuccess) return e;
        mc_kernel<KIND, L, CP_RNG_PHILOX><<<grid, kBlock, lay.bytes, st>>>(K);
    }
    return cudaGetLastError();
}

template <int KIND, int L>
static cudaError_t group_score(const KProblem &P, const double *dof, unsigned long long n, double *out,
                               cudaStream_t st) {
    const SharedLayout lay = shared_layout(KIND, P.n_items, P.n_syms, L);
    const unsigned long long gpb = kBlock / L;
    cudaError_t e = cp_allow_shared(score_kernel<KIND, L>, lay.bytes);
    if (e != cudaSuccess) return e;
    score_kernel<KIND, L><<<(unsigned)((n + gpb - 1) / gpb), kBlock, lay.bytes, st>>>(P, dof, n, out);
    return cudaGetLastError();
}

template <int KIND, int L>
static cudaError_t group_replay(const KProblem &P, const double *init, const cp_move *mv, unsigned long long n_moves,
                                unsigned long long n, cp_decision *out, cudaStream_t st) {
    const SharedLayout lay = shared_layout(KIND, P.n_items, P.n_syms, L);
    const unsigned long long gpb = kBlock / L;
    cudaError_t e = cp_allow_shared(replay_kernel<KIND, L>, lay.bytes);
    if (e != cudaSuccess) return e;
    replay_kernel<KIND, L><<<(unsigned)((n + gpb - 1) / gpb), kBlock, lay.bytes, st>>>(P, init, mv, n_moves, n, out);
    return cudaGetLastError();
}

#define CP_DEFINE_GROUP_KIND(KIND, name)                                                                       \
    cudaError_t cp_##name##_mc(const KParams &K, int lanes, cudaStream_t st) {                                 \
        CP_LANE_SWITCH(lanes, return (group_mc<KIND, L>(K, st)))                                               \
    }                                                                                                          \
    cudaError_t cp_##name##_score(const KProblem &P, int lanes, const double *dof, unsigned long long n,       \
                                  double *out, cudaStream_t st) {                                              \
        CP_LANE_SWITCH(lanes, return (group_score<KIND, L>(P, dof, n, out, st)))                               \
    }                                                                                                          \
    cudaError_t cp_##name##_replay(const KProblem &P, int lanes, const double *init, const cp_move *mv,        \
                                   unsigned long long n_moves, unsigned long long n, cp_decision *out,         \
                                   cudaStream_t st) {                                                          \
        CP_LANE_SWITCH(lanes, return (group_replay<KIND, L>(P, init, mv, n_moves, n, out, st)))                \
    }
