// one-thread-per-replica kernels, LineShape with 0 components (0 = run-time count)
#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 0)
