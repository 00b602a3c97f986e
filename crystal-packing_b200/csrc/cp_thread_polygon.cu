#include "cp_thread_inst.cuh"
CP_DEFINE_THREAD_KIND(CP_POLYGON, thread_polygon, 5, 4)
