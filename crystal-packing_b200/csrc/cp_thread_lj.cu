// one-thread-per-replica kernels, LJShape2: trimer, single site, run-time count
#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_NI(CP_LJ, thread_lj, 0)
CP_DEFINE_THREAD_NI(CP_LJ, thread_lj, 1)
CP_DEFINE_THREAD_NI(CP_LJ, thread_lj, 3)
