// cp_launch.h — host-side launch entry points of the kernels, one translation unit per shape kind
// so that the library builds in parallel (make -j).
#ifndef CP_LAUNCH_H
#define CP_LAUNCH_H

#include <cuda_runtime.h>

#include "cp_device.cuh"

// lanes: 1 = one thread per replica (cp_thread.cuh), 4/8/16/32 = lane groups (cp_group.cuh)
cudaError_t cp_launch_mc(const KParams &K, int lanes, cudaStream_t st);
cudaError_t cp_launch_score(const KProblem &P, int lanes, const double *dof, unsigned long long n, double *out,
                            cudaStream_t st);
cudaError_t cp_launch_replay(const KProblem &P, int lanes, const double *init, const cp_move *mv,
                             unsigned long long n_moves, unsigned long long n, cp_decision *out, cudaStream_t st);
cudaError_t cp_launch_images(const KProblem &P, const double *dof, unsigned long long n, double *out, cudaStream_t st);
// does the per-thread geometry of this problem fit the one-thread-per-replica kernels?
bool cp_thread_kernel_fits(const KProblem &P);
size_t cp_thread_shared_bytes(const KProblem &P);
// component count of the build that serves this problem (0 = run-time count); decides how
// KProblem::ptsf is filled (cp_api.cu)
int cp_thread_route_ni(const KProblem &P);

// per-kind implementations (cp_group_<kind>.cu, cp_thread_<kind>.cu)
#define CP_DECLARE_KIND(name)                                                                                  \
    cudaError_t cp_##name##_mc(const KParams &K, int lanes, cudaStream_t st);                                  \
    cudaError_t cp_##name##_score(const KProblem &P, int lanes, const double *dof, unsigned long long n,       \
                                  double *out, cudaStream_t st);                                               \
    cudaError_t cp_##name##_replay(const KProblem &P, int lanes, const double *init, const cp_move *mv,        \
                                   unsigned long long n_moves, unsigned long long n, cp_decision *out,         \
                                   cudaStream_t st);
CP_DECLARE_KIND(group_polygon)
CP_DECLARE_KIND(group_discs)
CP_DECLARE_KIND(group_lj)
#define CP_DECLARE_THREAD_NI(name, NI)                                                                          \
    cudaError_t cp_##name##_mc_##NI(const KParams &K, cudaStream_t st);                                         \
    cudaError_t cp_##name##_score_##NI(const KProblem &P, const double *dof, unsigned long long n, double *out, \
                                       cudaStream_t st);                                                       \
    cudaError_t cp_##name##_replay_##NI(const KProblem &P, const double *init, const cp_move *mv,               \
                                        unsigned long long n_moves, unsigned long long n, cp_decision *out,    \
                                        cudaStream_t st);
CP_DECLARE_THREAD_NI(thread_polygon, 0)
CP_DECLARE_THREAD_NI(thread_polygon, 3)
CP_DECLARE_THREAD_NI(thread_polygon, 4)
CP_DECLARE_THREAD_NI(thread_polygon, 5)
CP_DECLARE_THREAD_NI(thread_polygon, 6)
CP_DECLARE_THREAD_NI(thread_polygon, 7)
CP_DECLARE_THREAD_NI(thread_polygon, 8)
CP_DECLARE_THREAD_NI(thread_polygon, 9)
CP_DECLARE_THREAD_NI(thread_polygon, 10)
CP_DECLARE_THREAD_NI(thread_polygon, 11)
CP_DECLARE_THREAD_NI(thread_polygon, 12)
CP_DECLARE_THREAD_NI(thread_discs, 0)
CP_DECLARE_THREAD_NI(thread_discs, 1)
CP_DECLARE_THREAD_NI(thread_discs, 3)
CP_DECLARE_THREAD_NI(thread_lj, 0)
CP_DECLARE_THREAD_NI(thread_lj, 1)
CP_DECLARE_THREAD_NI(thread_lj, 3)

// dynamic shared memory above the 48 KB default needs an explicit opt-in per kernel
template <typename Kernel>
static inline cudaError_t cp_allow_shared(Kernel kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#endif  // CP_LAUNCH_H
