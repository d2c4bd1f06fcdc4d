// LightGlue internal interface (shared by lg.cu and store.cu).
#pragma once
#include <stdint.h>

#include <functional>

namespace dv {

struct Engine;

// One image's tokens: device pointers to pixel keypoints [n,2] and descriptors [n,256] (fp32).
struct LgSeg {
  const float* kpts;
  const float* desc;
  int n;        // keypoints
  int w, h;     // image size used for the caller-side normalisation (deep_net.cpp:839-841)
  int off;      // first row in the packed token buffers (filled by lg_run)
};

// tcgen05 attention (lg_attn.cu): operands are windows of the packed qkv buffer [T,768]
struct AttnJobU { int q_row, nq, k_row, nk, q_col, k_col, v_col, pad; };

// segs [2P]: (query, old) per pair; results stay on device.  `after_load` (optional) is invoked right after the kernel
// that reads the callers' keypoint / descriptor pointers has been queued (store.cu verifies remote reads there).
// `out_base`: first output slot (matches / mscores / kcount of pair p land at slot out_base + p), so a batch processed
// in several chunks is fetched with ONE lg_fetch_batch at the end.
int lg_run(Engine* e, int P, const LgSeg* segs, const std::function<int()>* after_load = nullptr, int out_base = 0);
int lg_fetch_batch(Engine* e, int P, int cap, const int* slot, int32_t* matches, float* mscores, int32_t* k_out);
int lg_fetch_batch_begin(Engine* e, int P, int cap);      // queue the result copies behind the match kernels
int lg_fetch_batch_end(Engine* e, int P, int cap, const int* slot, int32_t* matches, float* mscores, int32_t* k_out);
int lg_fetch(Engine* e, int p, int cap, int32_t* matches, float* mscores, float* mk0, float* mk1, int32_t* k_out);

}  // namespace dv

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
namespace dv {
int lg_attn_init();
int plan_lg_attn(CUtensorMap* tm, const __half* qkv, int T_cap);
// ctx: output base, row stride ldo halves (256 = the ctx buffer; 512 = straight into the msg half of X2)
int launch_lg_attn(const CUtensorMap& tm, const AttnJobU* jobs, int n_jobs, int max_nq, __half* ctx, int ldo, float scale,
                   cudaStream_t st);
// persistent variant: two resident CTAs per SM walk a cost-sorted list of complete (job, head, query tile) descriptors
int lg_attn_item_bytes();
int lg_attn_items(const AttnJobU* jobs, int n_jobs, void* items_out);     // host: builds the list, returns its length
int launch_lg_attn_persist(const CUtensorMap& tm, const void* items, int n_items, __half* ctx, int ldo, float scale,
                           cudaStream_t st);
}  // namespace dv
