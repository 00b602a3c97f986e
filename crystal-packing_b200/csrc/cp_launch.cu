// cp_launch.cu — routes a launch to the translation unit that holds the kernels of its shape kind.
#include "cp_launch.h"
#include "cp_thread.cuh"

bool cp_thread_kernel_fits(const KProblem &P) {
    return thread_block_shared_bytes(P.kind, P.n_items, P.n_syms) <= 200 * 1024;
}

size_t cp_thread_shared_bytes(const KProblem &P) { return thread_block_shared_bytes(P.kind, P.n_items, P.n_syms); }

cudaError_t cp_launch_mc(const KParams &K, int lanes, cudaStream_t st) {
    switch (K.prob.kind) {
    case CP_POLYGON: return lanes == 1 ? cp_thread_polygon_mc(K, lanes, st) : cp_group_polygon_mc(K, lanes, st);
    case CP_DISCS: return lanes == 1 ? cp_thread_discs_mc(K, lanes, st) : cp_group_discs_mc(K, lanes, st);
    case CP_LJ: return lanes == 1 ? cp_thread_lj_mc(K, lanes, st) : cp_group_lj_mc(K, lanes, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t cp_launch_score(const KProblem &P, int lanes, const double *dof, unsigned long long n, double *out,
                            cudaStream_t st) {
    switch (P.kind) {
    case CP_POLYGON:
        return lanes == 1 ? cp_thread_polygon_score(P, lanes, dof, n, out, st)
                          : cp_group_polygon_score(P, lanes, dof, n, out, st);
    case CP_DISCS:
        return lanes == 1 ? cp_thread_discs_score(P, lanes, dof, n, out, st)
                          : cp_group_discs_score(P, lanes, dof, n, out, st);
    case CP_LJ:
        return lanes == 1 ? cp_thread_lj_score(P, lanes, dof, n, out, st) : cp_group_lj_score(P, lanes, dof, n, out, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t cp_launch_replay(const KProblem &P, int lanes, const double *init, const cp_move *mv,
                             unsigned long long n_moves, unsigned long long n, cp_decision *out, cudaStream_t st) {
    switch (P.kind) {
    case CP_POLYGON:
        return lanes == 1 ? cp_thread_polygon_replay(P, lanes, init, mv, n_moves, n, out, st)
                          : cp_group_polygon_replay(P, lanes, init, mv, n_moves, n, out, st);
    case CP_DISCS:
        return lanes == 1 ? cp_thread_discs_replay(P, lanes, init, mv, n_moves, n, out, st)
                          : cp_group_discs_replay(P, lanes, init, mv, n_moves, n, out, st);
    case CP_LJ:
        return lanes == 1 ? cp_thread_lj_replay(P, lanes, init, mv, n_moves, n, out, st)
                          : cp_group_lj_replay(P, lanes, init, mv, n_moves, n, out, st);
    default: return cudaErrorInvalidValue;
    }
}
