#include <algorithm>
#include <stdlib.h>

#include <nvtx3/nvToolsExt.h>

#include "engine.h"

namespace dv {

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DV_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

__global__ void k_f32_to_f16(const float* __restrict__ s, __half* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = __float2half_rn(s[i]);
}
__global__ void k_f16_to_f32(const __half* __restrict__ s, float* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = __half2float(s[i]);
}
void f32_to_f16(const float* src, __half* dst, int64_t n, cudaStream_t st) {
  if (n <= 0) return;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  k_f32_to_f16<<<blocks, 256, 0, st>>>(src, dst, n);
}
void f16_to_f32(const __half* src, float* dst, int64_t n, cudaStream_t st) {
  if (n <= 0) return;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  k_f16_to_f32<<<blocks, 256, 0, st>>>(src, dst, n);
}

int Engine::upload_f16(const std::vector<float>& v, __half** out) {
  std::vector<__half> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(v[i]);
  int rc = alloc(out, v.size());
  if (rc) return rc;
  DV_CUDA_OK(cudaMemcpyAsync(*out, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice, st));
  DV_CUDA_OK(cudaStreamSynchronize(st));
  return DV_OK;
}
int Engine::upload_f32(const std::vector<float>& v, float** out) {
  int rc = alloc(out, v.size());
  if (rc) return rc;
  DV_CUDA_OK(cudaMemcpyAsync(*out, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  DV_CUDA_OK(cudaStreamSynchronize(st));
  return DV_OK;
}

// NVTX range per pipeline stage (SURVEY §5 tracing): header-only nvtx3, a no-op unless a profiler injected itself
static const char* const kStageName[ST_COUNT] = {"dv.sp_convs", "dv.sp_post", "dv.mixvpr", "dv.knn", "dv.lightglue", "dv.copies"};

StageScope::StageScope(Engine* e_, int s) : e(e_), stage(s) {
  nvtxRangePushA(kStageName[s]);
  if (!e->stats_on) return;
  auto get = [&]() {
    cudaEvent_t ev;
    if (!e->ev_pool.empty()) { ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
    else cudaEventCreate(&ev);
    return ev;
  };
  a = get(); b = get();
  cudaEventRecord(a, e->st);
}
StageScope::~StageScope() {
  nvtxRangePop();
  if (!a) return;
  cudaEventRecord(b, e->st);
  e->pending.push_back({stage, a, b});
}

ProbeScope::ProbeScope(Engine* e_, int which) : e(e_) {
  if (!e->probe_on || e->probe_sel != which) return;
  auto get = [&]() {
    cudaEvent_t ev;
    if (!e->ev_pool.empty()) { ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
    else cudaEventCreate(&ev);
    return ev;
  };
  a = get(); b = get();
  cudaEventRecord(a, e->st);
}
ProbeScope::~ProbeScope() {
  if (!a) return;
  cudaEventRecord(b, e->st);
  e->probe_pending.push_back({0, a, b});
}

}  // namespace dv
