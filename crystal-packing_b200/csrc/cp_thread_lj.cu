#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_KIND(CP_LJ, thread_lj, 3, 1)
