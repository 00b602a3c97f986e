#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_KIND(CP_DISCS, thread_discs, 3, 1)
