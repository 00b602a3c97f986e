// one-thread-per-replica kernels, LineShape: triangle and hexagon
#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 3)
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 6)
