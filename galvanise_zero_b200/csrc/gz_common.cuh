// Shared host-side helpers of libgz_b200: error reporting and CUDA call checking.
#ifndef GZ_COMMON_CUH
#define GZ_COMMON_CUH

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/gz_b200.h"

void gz_set_error(const char* fmt, ...);

#define GZ_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            gz_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                         cudaGetErrorString(_e));                                            \
            return GZ_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#endif
