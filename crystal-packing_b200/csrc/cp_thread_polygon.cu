// one-thread-per-replica kernels, LineShape: pentagon, square and the run-time-count build
#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 0)
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 4)
CP_DEFINE_THREAD_NI(CP_POLYGON, thread_polygon, 5)
