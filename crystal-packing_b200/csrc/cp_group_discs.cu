#include "cp_group_inst.cuh"
CP_DEFINE_GROUP_KIND(CP_DISCS, group_discs)
