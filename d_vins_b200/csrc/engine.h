// Internal engine state (not part of the C ABI).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dvins_perception.h"
#include "common.cuh"
#include "gemm.h"

namespace dv {

struct HostTensor {
  std::vector<int> dims;
  std::vector<float> data;
  int64_t numel() const { int64_t n = 1; for (int d : dims) n *= d; return n; }
};
using WeightMap = std::map<std::string, HostTensor>;
int load_weight_file(const char* path, WeightMap* out);   // weights.cpp

enum Stage { ST_SP_CONV = 0, ST_SP_POST = 1, ST_MIX = 2, ST_KNN = 3, ST_LG = 4, ST_COPY = 5, ST_COUNT = 6 };

struct SpNet;   // sp.cu
struct MixNet;  // mix.cu
struct LgNet;   // lg.cu
struct Bank;    // knn.cu
struct Store;   // engine.cpp (device-resident keyframe features)
struct Comm;    // comm.cpp

struct Engine {
  dv_config cfg{};
  std::string weights_path;
  cudaStream_t st = nullptr;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  int B = 1, H = 0, W = 0;
  int h8 = 0, w8 = 0;          // encoder output grid = floor(H/8), floor(W/8)
  WeightMap weights;
  std::vector<void*> allocs;   // every device allocation, freed in dv_destroy
  std::vector<void*> pinned;

  // raw frames of the current batch
  // Frames are double-buffered and uploaded on their own stream: dv_batch_upload / dv_frame_upload return as soon as
  // the copy is queued, so the upload of keyframe round R+1 overlaps the matching of round R.  `d_img` is the buffer
  // of the most recent upload; consumers call image_acquire() first and image_release() after their last read.
  uint8_t* d_img = nullptr;    // [B, H, W, ch] u8 as uploaded (= d_img_buf[img_idx])
  uint8_t* d_img_buf[2] = {nullptr, nullptr};
  int img_idx = 0;
  bool img_pending = false;    // an upload has been queued that the compute stream has not waited for yet
  cudaStream_t st_copy = nullptr;
  cudaEvent_t ev_img_ready = nullptr, ev_img_free[2] = {nullptr, nullptr};
  uint8_t* h_img = nullptr;    // pinned staging
  // upload side: pick the other buffer, make the copy stream wait until the compute stream has released it
  uint8_t* image_begin_upload() {
    const int nb = img_idx ^ 1;
    cudaStreamWaitEvent(st_copy, ev_img_free[nb], 0);
    return d_img_buf[nb];
  }
  void image_end_upload() {
    cudaEventRecord(ev_img_ready, st_copy);
    img_idx ^= 1;
    d_img = d_img_buf[img_idx];
    img_pending = true;
  }
  // compute side
  void image_acquire() {
    if (img_pending) { cudaStreamWaitEvent(st, ev_img_ready, 0); img_pending = false; }
  }
  void image_release() { cudaEventRecord(ev_img_free[img_idx], st); }
  int img_ch = 1;
  int cur_b = 0;               // frames in the current (being / last extracted) batch
  // dv_batch_upload may run one round ahead: the uploaded batch becomes current only when something consumes it
  // (dv_batch_extract, or the lazily evaluated per-frame entry points), so commit / search / match of the previous
  // round stay valid while its successor's frames are already travelling.
  int next_b = 0;
  bool next_pending = false;
  void adopt_upload() {
    if (next_pending) { cur_b = next_b; enc_done = det_done = mix_done = false; next_pending = false; }
  }
  bool enc_done = false, det_done = false, mix_done = false;
  bool match_pending = false;   // dv_batch_match_begin .. _end: LightGlue's result staging buffers hold uncollected matches

  SpNet* sp = nullptr;
  MixNet* mix = nullptr;
  LgNet* lg = nullptr;
  Bank* bank = nullptr;
  Store* store = nullptr;
  Comm* comm = nullptr;

  // CUDA graphs for the launch-bound B = 1 (per-keyframe latency) paths; DV_GRAPHS=0 disables (A/B)
  bool graphs_on = true;

  // stats
  bool stats_on = false;
  double stage_ms[ST_COUNT] = {0, 0, 0, 0, 0, 0};
  int64_t launches = 0;
  struct Pending { int stage; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> ev_pool;
  bool probe_on = false;
  int probe_sel = 0;            // which kernel the probe brackets: 0 conv1a+conv1b implicit GEMM, 1 kNN bank scan
  double probe_ms = 0; int64_t probe_n = 0;
  std::vector<Pending> probe_pending;

  // debug tensors by name: device pointer + element count + dtype (0 f32, 1 f16, 2 i32)
  struct Dbg { const void* p; int64_t n; int dtype; };
  std::map<std::string, Dbg> dbg;

  template <class T> int alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); return DV_ERR_CUDA; }
    cudaMemsetAsync(q, 0, count * sizeof(T) + 256, st);
    allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return DV_OK;
  }
  template <class T> int alloc_pinned(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMallocHost(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) { set_error(std::string("cudaMallocHost failed: ") + cudaGetErrorString(e)); return DV_ERR_CUDA; }
    pinned.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return DV_OK;
  }
  const HostTensor* weight(const std::string& name) const {
    auto it = weights.find(name);
    return it == weights.end() ? nullptr : &it->second;
  }
  int upload_f16(const std::vector<float>& v, __half** out);
  int upload_f32(const std::vector<float>& v, float** out);
};

// One captured-and-instantiated launch sequence.  The B = 1 paths are launch-bound (~130 launches of a few
// microseconds each per LightGlue match): the sequence is captured ONCE per shape into a CUDA graph - PDL edges
// included - and replayed with a single cudaGraphLaunch.  Per-call host tables live in pinned memory at fixed
// addresses and are read by the graph's memcpy nodes at replay time.
struct GraphCache {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;        // kernels inside (added to Engine::launches on every replay)
  bool failed = false;         // capture / instantiation was refused once: stay on the eager path
  ~GraphCache() { if (exec) cudaGraphExecDestroy(exec); }
  GraphCache() = default;
  GraphCache(const GraphCache&) = delete;
  GraphCache& operator=(const GraphCache&) = delete;
};
// enqueue(): queues the whole sequence on e->st and returns a dv_status.  Event-timed modes (stage stats, kernel probe)
// bypass the graph because they record events between the launches.
template <class F>
int run_graphed(Engine* e, GraphCache& gc, F&& enqueue) {
  if (!e->graphs_on || e->stats_on || e->probe_on || gc.failed) return enqueue();
  if (!gc.exec) {
    if (cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      gc.failed = true;
      return enqueue();
    }
    const int64_t l0 = e->launches;
    const int rc = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->st, &graph);
    gc.launches = e->launches - l0;
    e->launches = l0;
    if (rc != 0 || ce != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      gc.failed = true;
      return rc != 0 ? rc : enqueue();
    }
    const cudaError_t ci = cudaGraphInstantiate(&gc.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess) {
      cudaGetLastError();
      gc.exec = nullptr;
      gc.failed = true;
      log_msg(2, "CUDA graph instantiation failed (%s): this launch sequence stays eager", cudaGetErrorString(ci));
      return enqueue();
    }
    log_msg(4, "captured a CUDA graph of %lld kernel launches", (long long)gc.launches);
  }
  e->launches += gc.launches;
  DV_CUDA_OK(cudaGraphLaunch(gc.exec, e->st));
  return DV_OK;
}

// stage timing scope (no-op unless stats are enabled)
struct ProbeScope {   // event pair around one kernel launch (dominant-kernel roofline probe)
  Engine* e; cudaEvent_t a = nullptr, b = nullptr;
  explicit ProbeScope(Engine* e_, int which = 0);
  ~ProbeScope();
};
struct StageScope {
  Engine* e; int stage; cudaEvent_t a = nullptr, b = nullptr;
  StageScope(Engine* e_, int s);
  ~StageScope();
};
#define DV_LAUNCHED(e, n) ((e)->launches += (n))
#define DV_TRY(expr) do { int _rc = (expr); if (_rc) return (dv_status)_rc; } while (0)

// helper kernels (util.cu)
void f32_to_f16(const float* src, __half* dst, int64_t n, cudaStream_t st);
void f16_to_f32(const __half* src, float* dst, int64_t n, cudaStream_t st);

// subsystem entry points --------------------------------------------------------------------------
int sp_init(Engine* e);                       // sp.cu
void sp_free(Engine* e);
int sp_run_encoder(Engine* e, int b);         // gray -> conv1a .. heads (logits + dense descriptor map)
int sp_dbg_refresh(Engine* e, const char* name);     // lazily produced debug tensors ("gray")
int sp_run_detect(Engine* e, int b);          // softmax/d2s, NMS, select, sample -> device results
int sp_run_describe(Engine* e, int b, const float* d_kpts, const int* d_n, int cap, float* d_desc);
void sp_device_results(Engine* e, int** kpts, float** kpts_f, float** scores, int** n, float** desc, float** re_kpts,
                       int** re_n, float** re_desc);
int sp_nms_select_dbg(Engine* e, const float* h_smap, int h8, int w8, float* h_nms, int32_t* kp, float* sc, int32_t* n);

int mix_init(Engine* e);                      // mix.cu
void mix_free(Engine* e);
int mix_run(Engine* e, int b);                // -> global descriptors [b,512] on device
float* mix_gdesc(Engine* e);

int lg_init(Engine* e);                       // lg.cu
void lg_free(Engine* e);

int bank_init(Engine* e);                     // knn.cu
void bank_free(Engine* e);

int store_init(Engine* e);                    // store.cu
void store_free(Engine* e);
void comm_free(Engine* e);                    // comm.cpp
int store_exchange_peers(Engine* e);          // store.cu (called by dv_comm_init)

}  // namespace dv
