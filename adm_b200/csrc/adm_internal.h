// Internal helpers shared by the .cu translation units of libadm_b200.so.
#pragma once
#include <stdarg.h>
#include <cuda_runtime.h>
#include "../../include/adm_b200.h"

namespace adm {
void set_error(const char* fmt, ...);
const char* get_error();
int num_sms();
void count_launch();

#define ADM_CHECK_LAUNCH(name)                                              \
    do {                                                                    \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) {                                           \
            adm::set_error("%s launch: %s", name, cudaGetErrorString(e__)); \
            return ADM_ERR_CUDA;                                            \
        }                                                                   \
        adm::count_launch();                                                \
    } while (0)
}  // namespace adm
