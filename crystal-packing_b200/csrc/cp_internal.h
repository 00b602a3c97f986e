// cp_internal.h — declarations shared by the translation units of libcpgpu.so.
#ifndef CP_INTERNAL_H
#define CP_INTERNAL_H

#include "../../include/cp_gpu.h"

// records a thread-local message (returned by cp_last_error) and passes the status through
cp_status cp_fail(cp_status st, const char *msg);

#endif
