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
bool pdl_enabled();  // ADM_PDL=1: programmatic dependent launch of the hot kernels (off by default, see tc_gemm.cu)

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+).  A kernel launched with the attribute may be scheduled while its predecessor in
// the stream is still draining: it runs its prologue (barrier init, TMEM allocation, index math) on the SMs that are
// already free and blocks in pdl_wait() until the predecessor grid has completed and its memory is visible.  Every kernel
// calls pdl_trigger() first thing, which lets ITS successor be scheduled as soon as all of its own CTAs have started.
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// <<<>>> replacement for the hot kernels: optional thread-block cluster + the PDL attribute.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                   int cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (cluster_x > 0) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster_x;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

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
