// Shared host/device helpers for the D_VINS B200 perception engine.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace dv {

// Sticky error text of the calling thread's last failure (surfaced through dv_last_error()).
void set_error(const std::string& msg);
const char* get_error();
// levelled logger behind dv_log_set_level / dv_log_set_sink (1 error .. 4 debug); printf-style
void log_msg(int level, const char* fmt, ...);
bool log_enabled(int level);

#define DV_CUDA_OK(expr)                                                                     \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      char _b[512];                                                                          \
      snprintf(_b, sizeof(_b), "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,         \
               cudaGetErrorString(_e));                                                      \
      dv::set_error(_b);                                                                     \
      return DV_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// A kernel launched through launch_pdl() may become resident while its predecessor in the stream is still draining:
// its prologue (barrier init, TMEM allocation, tensor-map prefetch, weight staging) then overlaps the predecessor's
// tail.  It MUST call pdl_wait() before touching anything an earlier kernel wrote (or still reads), and every kernel
// calls pdl_trigger() first thing so that its successor can be scheduled as soon as SMs free up.  Without the launch
// attribute both instructions are no-ops.  DV_PDL=0 disables the attribute (A/B switch).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace dv
