// one-thread-per-replica kernels, MolecularShape2: trimer, single disc, run-time count
#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_NI(CP_DISCS, thread_discs, 0)
CP_DEFINE_THREAD_NI(CP_DISCS, thread_discs, 1)
CP_DEFINE_THREAD_NI(CP_DISCS, thread_discs, 3)
