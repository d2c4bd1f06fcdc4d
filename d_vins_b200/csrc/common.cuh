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

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace dv
