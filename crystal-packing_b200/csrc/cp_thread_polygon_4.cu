// one-thread-per-replica kernels, LineShape with 4 components (0 = run-time count)
#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 4)
