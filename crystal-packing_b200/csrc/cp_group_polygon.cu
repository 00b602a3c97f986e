#include "cp_group_inst.cuh"
CP_DEFINE_GROUP_KIND(CP_POLYGON, group_polygon)
