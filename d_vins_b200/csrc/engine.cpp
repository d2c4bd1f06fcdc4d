// C-ABI implementation: engine lifetime, per-keyframe and batched entry points, measurement hooks.
// The heavy lifting lives in sp.cu / mix.cu / lg.cu / knn.cu / gemm_umma.cu.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "engine.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>


namespace dv {

static thread_local std::string g_err;
static std::atomic<int> g_log_level{[] { const char* e = getenv("DV_LOG"); return e ? atoi(e) : 2; }()};
static std::mutex g_log_mu;
static dv_log_sink g_log_sink = nullptr;
static void* g_log_user = nullptr;
bool log_enabled(int level) { return level <= g_log_level.load(std::memory_order_relaxed); }
void log_msg(int level, const char* fmt, ...) {
  if (!log_enabled(level)) return;
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_log_mu);
  if (g_log_sink) { g_log_sink(level, buf, g_log_user); return; }
  static const char* names[5] = {"", "error", "warn", "info", "debug"};
  fprintf(stderr, "[dvins %s] %s\n", names[level < 1 ? 1 : (level > 4 ? 4 : level)], buf);
}
void set_error(const std::string& msg) { g_err = msg; log_msg(1, "%s", msg.c_str()); }
const char* get_error() { return g_err.c_str(); }

// Device scratch of the dv_dbg_* entry points: released on every return path (DV_CUDA_OK returns early on failure).
struct Scratch {
  std::vector<void*> bufs;
  ~Scratch() { for (void* q : bufs) cudaFree(q); }
  template <class T> cudaError_t get(T** out, size_t bytes) {
    void* q = nullptr;
    const cudaError_t err = cudaMalloc(&q, bytes);
    if (err == cudaSuccess) bufs.push_back(q);
    *out = reinterpret_cast<T*>(q);
    return err;
  }
};

static int drain_stats(Engine* e) {
  for (auto& p : e->pending) {
    float ms = 0.f;
    cudaEventSynchronize(p.b);
    cudaEventElapsedTime(&ms, p.a, p.b);
    e->stage_ms[p.stage] += ms;
    e->ev_pool.push_back(p.a);
    e->ev_pool.push_back(p.b);
  }
  e->pending.clear();
  return DV_OK;
}

}  // namespace dv

using namespace dv;

#define DV_CHECK_ENGINE(e) do { if (!(e)) { dv::set_error("null engine"); return DV_ERR_INVALID; } } while (0)

extern "C" {

void dv_config_default(dv_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(dv_config);
  c->device = 0;
  c->height = 480; c->width = 752;
  c->max_batch = 1;
  c->max_kpts = 512; c->nms_radius = 4; c->det_thresh = 0.0005f; c->border = 4;
  c->max_vio = 256;
  c->knn_k = 3; c->exclude_recent = 50;
  c->lg_filter_thresh = 0.1f; c->lg_max_kpts = 1024;
  c->bank_capacity = 65536;
  c->store_capacity = 64;
  c->world_size = 1; c->rank = 0;
  c->weights_path = nullptr;
}

const char* dv_last_error(void) { return dv::get_error(); }
void dv_log_set_level(int32_t level) { dv::g_log_level.store(level < 0 ? 0 : (level > 4 ? 4 : level)); }
int32_t dv_log_get_level(void) { return dv::g_log_level.load(); }
void dv_log_set_sink(dv_log_sink sink, void* user) {
  std::lock_guard<std::mutex> lk(dv::g_log_mu);
  dv::g_log_sink = sink;
  dv::g_log_user = user;
}
const char* dv_version(void) { return "d_vins_b200 0.1 (sm_100a)"; }

dv_status dv_create(const dv_config* cfg, dv_engine** out) {
  if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(dv_config)) {
    set_error("dv_create: null argument or dv_config size mismatch");
    return DV_ERR_INVALID;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("dv_create: no CUDA device visible - this engine has no CPU fallback");
    return DV_ERR_NOGPU;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { set_error("dv_create: bad device ordinal"); return DV_ERR_INVALID; }
  cudaDeviceProp prop;
  DV_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    set_error(std::string("dv_create: device is not Blackwell sm_100 (") + prop.name + "); kernels are sm_100a-only");
    return DV_ERR_NOGPU;
  }
  if (cfg->height < 64 || cfg->width < 64 || cfg->max_batch < 1 || cfg->max_kpts < 1 || cfg->max_kpts > 1024 ||
      cfg->max_vio < 1 || cfg->max_vio > 512 || cfg->lg_max_kpts < 16 || cfg->lg_max_kpts > 2048 ||
      cfg->knn_k < 1 || cfg->knn_k > 8 || cfg->nms_radius != 4 || cfg->world_size < 1 || cfg->rank < 0 ||
      cfg->rank >= cfg->world_size || cfg->bank_capacity < 1 || cfg->store_capacity < 1 || cfg->exclude_recent < 0) {
    set_error("dv_create: configuration out of the supported range");
    return DV_ERR_UNSUPPORTED;
  }
  // LightGlue sees max_vio window points on the query side and max_kpts + max_vio points on the old keyframe's side
  if (cfg->max_vio > cfg->lg_max_kpts || cfg->max_kpts + cfg->max_vio > cfg->lg_max_kpts) {
    set_error("dv_create: need max_kpts + max_vio <= lg_max_kpts (LightGlue token capacity per image)");
    return DV_ERR_UNSUPPORTED;
  }
  if (cfg->store_capacity < cfg->max_batch) {
    set_error("dv_create: store_capacity must hold at least one batch (max_batch keyframes)");
    return DV_ERR_UNSUPPORTED;
  }
  DV_CUDA_OK(cudaSetDevice(cfg->device));
  Engine* e = new Engine();
  e->cfg = *cfg;
  if (cfg->weights_path) { e->weights_path = cfg->weights_path; e->cfg.weights_path = e->weights_path.c_str(); }
  e->B = cfg->max_batch; e->H = cfg->height; e->W = cfg->width;
  { const char* env = getenv("DV_GRAPHS"); e->graphs_on = !(env && env[0] == '0'); }
  e->h8 = e->H / 8; e->w8 = e->W / 8;
  auto fail = [&](int rc) { dv_destroy(reinterpret_cast<dv_engine*>(e)); return (dv_status)rc; };
  if (cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); return fail(DV_ERR_CUDA); }
  cudaEventCreate(&e->ev_t0); cudaEventCreate(&e->ev_t1);
  int rc = gemm_init();
  if (rc) return fail(rc);
  size_t img_bytes = (size_t)e->B * e->H * e->W * 3;
  if ((rc = e->alloc(&e->d_img_buf[0], img_bytes))) return fail(rc);
  if ((rc = e->alloc(&e->d_img_buf[1], img_bytes))) return fail(rc);
  e->d_img = e->d_img_buf[0];
  if (cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking) != cudaSuccess) { set_error("copy stream create failed"); return fail(DV_ERR_CUDA); }
  cudaEventCreateWithFlags(&e->ev_img_ready, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_img_free[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_img_free[1], cudaEventDisableTiming);
  if ((rc = e->alloc_pinned(&e->h_img, img_bytes))) return fail(rc);
  if (!e->weights_path.empty()) {
    if ((rc = load_weight_file(e->weights_path.c_str(), &e->weights))) return fail(rc);
    if ((rc = sp_init(e))) return fail(rc);
    if ((rc = mix_init(e))) return fail(rc);
    if ((rc = lg_init(e))) return fail(rc);
  }
  if ((rc = bank_init(e))) return fail(rc);
  if ((rc = store_init(e))) return fail(rc);
  if (cudaStreamSynchronize(e->st) != cudaSuccess) { set_error("engine init: device error"); return fail(DV_ERR_CUDA); }
  e->weights.clear();   // host copies no longer needed
  {
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    log_msg(3, "engine created: %dx%d, max_batch %d, max_kpts %d, max_vio %d, bank %lld rows, store %d keyframes, rank %d/%d, "
            "weights %s, device memory in use %.1f GB of %.1f GB", e->H, e->W, e->B, e->cfg.max_kpts, e->cfg.max_vio,
         (long long)e->cfg.bank_capacity, e->cfg.store_capacity, e->cfg.rank, e->cfg.world_size,
         e->cfg.weights_path ? e->cfg.weights_path : "(none: stage-level entry points only)", (tot - fr) / 1e9, tot / 1e9);
  }
  *out = reinterpret_cast<dv_engine*>(e);
  return DV_OK;
}

void dv_destroy(dv_engine* h) {
  if (!h) return;
  Engine* e = reinterpret_cast<Engine*>(h);
  cudaSetDevice(e->cfg.device);
  if (e->st) cudaStreamSynchronize(e->st);
  comm_free(e);
  sp_free(e); mix_free(e); lg_free(e); bank_free(e); store_free(e);
  for (void* p : e->allocs) cudaFree(p);
  for (void* p : e->pinned) cudaFreeHost(p);
  for (auto& p : e->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto& p : e->probe_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  if (e->ev_t0) cudaEventDestroy(e->ev_t0);
  if (e->ev_t1) cudaEventDestroy(e->ev_t1);
  if (e->st_copy) { cudaStreamSynchronize(e->st_copy); cudaStreamDestroy(e->st_copy); }
  if (e->ev_img_ready) cudaEventDestroy(e->ev_img_ready);
  for (int i = 0; i < 2; ++i) if (e->ev_img_free[i]) cudaEventDestroy(e->ev_img_free[i]);
  if (e->st) cudaStreamDestroy(e->st);
  cudaGetLastError();
  delete e;
}

// ------------------------------------------------------------------------------------------ measurement
dv_status dv_timer_start(dv_engine* h) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  DV_CUDA_OK(cudaEventRecord(e->ev_t0, e->st));
  return DV_OK;
}
dv_status dv_timer_stop(dv_engine* h, float* ms) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  // uploads queued inside the timed region count even if nothing has consumed them yet
  if (e->img_pending) DV_CUDA_OK(cudaStreamWaitEvent(e->st, e->ev_img_ready, 0));
  DV_CUDA_OK(cudaEventRecord(e->ev_t1, e->st));
  DV_CUDA_OK(cudaEventSynchronize(e->ev_t1));
  float t = 0.f;
  DV_CUDA_OK(cudaEventElapsedTime(&t, e->ev_t0, e->ev_t1));
  if (ms) *ms = t;
  return DV_OK;
}
dv_status dv_sync(dv_engine* h) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  DV_CUDA_OK(cudaStreamSynchronize(e->st_copy));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  return DV_OK;
}
dv_status dv_stats_enable(dv_engine* h, int32_t on) {
  DV_CHECK_ENGINE(h);
  reinterpret_cast<Engine*>(h)->stats_on = on != 0;
  return DV_OK;
}
dv_status dv_stats_reset(dv_engine* h) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  drain_stats(e);
  for (double& v : e->stage_ms) v = 0;
  e->launches = 0;
  return DV_OK;
}
dv_status dv_stats_read(dv_engine* h, double* stage_ms6, int64_t* launches) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  drain_stats(e);
  if (stage_ms6) for (int i = 0; i < ST_COUNT; ++i) stage_ms6[i] = e->stage_ms[i];
  if (launches) *launches = e->launches;
  return DV_OK;
}

dv_status dv_probe_enable(dv_engine* h, int32_t on) {
  DV_CHECK_ENGINE(h);
  reinterpret_cast<Engine*>(h)->probe_on = on != 0;
  return DV_OK;
}
dv_status dv_probe_select(dv_engine* h, int32_t which) {
  DV_CHECK_ENGINE(h);
  if (which < 0 || which > 1) { set_error("dv_probe_select: 0 = conv1a+conv1b kernel, 1 = kNN scan"); return DV_ERR_INVALID; }
  reinterpret_cast<Engine*>(h)->probe_sel = which;
  return DV_OK;
}
dv_status dv_probe_read(dv_engine* h, double* ms, int64_t* launches, int32_t reset) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  for (auto& p : e->probe_pending) {
    float t = 0.f;
    cudaEventElapsedTime(&t, p.a, p.b);
    e->probe_ms += t;
    e->probe_n++;
    e->ev_pool.push_back(p.a);
    e->ev_pool.push_back(p.b);
  }
  e->probe_pending.clear();
  if (ms) *ms = e->probe_ms;
  if (launches) *launches = e->probe_n;
  if (reset) { e->probe_ms = 0; e->probe_n = 0; }
  return DV_OK;
}

// ------------------------------------------------------------------------------------------ stage-level debug
dv_status dv_dbg_gemm(dv_engine* h, const float* A, const float* Bm, const float* bias, int32_t M, int32_t N, int32_t K,
                      int32_t relu, float* D) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Scratch sc;
  if (!A || !Bm || !D || M <= 0 || N <= 0 || K <= 0 || (K % 8) || (N % 4)) {
    set_error("dv_dbg_gemm: need K % 8 == 0, N % 4 == 0");
    return DV_ERR_INVALID;
  }
  float *dA32 = nullptr, *dB32 = nullptr, *dD = nullptr, *dbias = nullptr;
  __half *dA = nullptr, *dB = nullptr;
  DV_CUDA_OK(sc.get(&dA32, (size_t)M * K * 4));
  DV_CUDA_OK(sc.get(&dB32, (size_t)N * K * 4));
  DV_CUDA_OK(sc.get(&dA, (size_t)M * K * 2));
  DV_CUDA_OK(sc.get(&dB, (size_t)N * K * 2));
  DV_CUDA_OK(sc.get(&dD, (size_t)M * N * 4));
  DV_CUDA_OK(cudaMemcpyAsync(dA32, A, (size_t)M * K * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dB32, Bm, (size_t)N * K * 4, cudaMemcpyHostToDevice, e->st));
  if (bias) {
    DV_CUDA_OK(sc.get(&dbias, (size_t)N * 4));
    DV_CUDA_OK(cudaMemcpyAsync(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice, e->st));
  }
  f32_to_f16(dA32, dA, (int64_t)M * K, e->st);
  f32_to_f16(dB32, dB, (int64_t)N * K, e->st);
  EpiParams ep;
  ep.out32 = dD; ep.ld32 = N; ep.bias = dbias; ep.relu = relu;
  GemmPlan pl;
  int rc = plan_gemm(&pl, dA, K, M, dB, K, N, K, ep);
  if (!rc) rc = launch_gemm(pl, M, e->st);
  if (!rc) {
    cudaError_t ce = cudaMemcpyAsync(D, dD, (size_t)M * N * 4, cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("dv_dbg_gemm: ") + cudaGetErrorString(ce)); rc = DV_ERR_CUDA; }
  }
  return (dv_status)rc;
}

dv_status dv_dbg_gemm_ex(dv_engine* h, const float* A, const float* Bm, const float* bias, const float* res,
                         int32_t res_is_f16, int32_t M, int32_t N, int32_t K, int32_t relu, float* D32, float* D16) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Scratch sc;
  if (!A || !Bm || (!D32 && !D16) || M <= 0 || N <= 0 || K <= 0 || (K % 8) || (N % 8)) {
    set_error("dv_dbg_gemm_ex: need K % 8 == 0, N % 8 == 0 and at least one output");
    return DV_ERR_INVALID;
  }
  const size_t mk = (size_t)M * K, nk = (size_t)N * K, mn = (size_t)M * N;
  float *dA32 = nullptr, *dB32 = nullptr, *dD = nullptr, *dbias = nullptr, *dtmp = nullptr;
  __half *dA = nullptr, *dB = nullptr, *dD16 = nullptr, *dR16 = nullptr;
  DV_CUDA_OK(sc.get(&dA32, mk * 4));
  DV_CUDA_OK(sc.get(&dB32, nk * 4));
  DV_CUDA_OK(sc.get(&dA, mk * 2));
  DV_CUDA_OK(sc.get(&dB, nk * 2));
  DV_CUDA_OK(sc.get(&dD, mn * 4));
  DV_CUDA_OK(sc.get(&dtmp, mn * 4));
  DV_CUDA_OK(sc.get(&dD16, mn * 2));
  DV_CUDA_OK(sc.get(&dR16, mn * 2));
  DV_CUDA_OK(cudaMemcpyAsync(dA32, A, mk * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dB32, Bm, nk * 4, cudaMemcpyHostToDevice, e->st));
  if (bias) {
    DV_CUDA_OK(sc.get(&dbias, (size_t)N * 4));
    DV_CUDA_OK(cudaMemcpyAsync(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice, e->st));
  }
  f32_to_f16(dA32, dA, (int64_t)mk, e->st);
  f32_to_f16(dB32, dB, (int64_t)nk, e->st);
  EpiParams ep;
  ep.bias = dbias; ep.relu = relu;
  if (D32) { ep.out32 = dD; ep.ld32 = N; }
  if (D16) { ep.out16 = dD16; ep.ld16 = N; }
  if (res) {
    if (res_is_f16) {
      DV_CUDA_OK(cudaMemcpyAsync(dtmp, res, mn * 4, cudaMemcpyHostToDevice, e->st));
      f32_to_f16(dtmp, dR16, (int64_t)mn, e->st);
      ep.res16 = dR16; ep.ldr16 = N;
    } else {
      // fp32 residual: in place in the output buffer when there is an fp32 output (LightGlue's x += ffn(x))
      float* r = D32 ? dD : dtmp;
      DV_CUDA_OK(cudaMemcpyAsync(r, res, mn * 4, cudaMemcpyHostToDevice, e->st));
      ep.res32 = r; ep.ldr32 = N;
    }
  }
  GemmPlan pl;
  int rc = plan_gemm(&pl, dA, K, M, dB, K, N, K, ep);
  if (!rc) rc = launch_gemm(pl, M, e->st);
  if (!rc) {
    cudaError_t ce = cudaSuccess;
    if (D32) ce = cudaMemcpyAsync(D32, dD, mn * 4, cudaMemcpyDeviceToHost, e->st);
    if (D16 && ce == cudaSuccess) {
      f16_to_f32(dD16, dtmp, (int64_t)mn, e->st);
      ce = cudaMemcpyAsync(D16, dtmp, mn * 4, cudaMemcpyDeviceToHost, e->st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("dv_dbg_gemm_ex: ") + cudaGetErrorString(ce)); rc = DV_ERR_CUDA; }
  }
  return (dv_status)rc;
}

dv_status dv_dbg_conv3x3(dv_engine* h, const float* x, const float* wgt, const float* bias, int32_t n, int32_t hh,
                         int32_t ww, int32_t cin, int32_t cout, int32_t relu, int32_t pool, float* y) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Scratch sc;
  if (!x || !wgt || !y || n <= 0 || hh <= 0 || ww <= 0 || (cin % 64) || (cout % 8)) {
    set_error("dv_dbg_conv3x3: need cin % 64 == 0, cout % 8 == 0");
    return DV_ERR_INVALID;
  }
  const size_t nx = (size_t)n * hh * ww * cin;
  const int ho = pool ? hh / 2 : hh, wo = pool ? ww / 2 : ww;
  const size_t ny = (size_t)n * ho * wo * cout;
  // repack torch [cout,cin,3,3] -> [cout, (r*3+s)*cin + c]
  std::vector<float> wr((size_t)cout * 9 * cin);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * cin + c] = wgt[((size_t)o * cin + c) * 9 + t];
  float *dx32 = nullptr, *dw32 = nullptr, *dy32 = nullptr, *dbias = nullptr;
  __half *dx = nullptr, *dw = nullptr, *dy = nullptr;
  DV_CUDA_OK(sc.get(&dx32, nx * 4)); DV_CUDA_OK(sc.get(&dx, nx * 2));
  DV_CUDA_OK(sc.get(&dw32, wr.size() * 4)); DV_CUDA_OK(sc.get(&dw, wr.size() * 2));
  DV_CUDA_OK(sc.get(&dy, ny * 2)); DV_CUDA_OK(sc.get(&dy32, ny * 4));
  DV_CUDA_OK(cudaMemsetAsync(dy, 0, ny * 2, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dx32, x, nx * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dw32, wr.data(), wr.size() * 4, cudaMemcpyHostToDevice, e->st));
  if (bias) {
    DV_CUDA_OK(sc.get(&dbias, (size_t)cout * 4));
    DV_CUDA_OK(cudaMemcpyAsync(dbias, bias, (size_t)cout * 4, cudaMemcpyHostToDevice, e->st));
  }
  f32_to_f16(dx32, dx, (int64_t)nx, e->st);
  f32_to_f16(dw32, dw, (int64_t)wr.size(), e->st);
  EpiParams ep;
  ep.out16 = dy; ep.ld16 = cout; ep.bias = dbias; ep.relu = relu; ep.pool = pool;
  GemmPlan pl;
  int rc = plan_conv3x3(&pl, dx, n, hh, ww, cin, dw, cout, ep);
  if (!rc) rc = launch_gemm(pl, n, e->st);
  if (!rc) {
    f16_to_f32(dy, dy32, (int64_t)ny, e->st);
    cudaError_t ce = cudaMemcpyAsync(y, dy32, ny * 4, cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("dv_dbg_conv3x3: ") + cudaGetErrorString(ce)); rc = DV_ERR_CUDA; }
  }
  return (dv_status)rc;
}

dv_status dv_dbg_conv3x3_halo64(dv_engine* h, const float* x, const float* wgt, const float* bias, int32_t n, int32_t hh,
                                int32_t ww, int32_t relu, int32_t pool, int32_t out_blocked, float* y) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Scratch sc;
  if (!x || !wgt || !bias || !y || n <= 0 || hh <= 0 || ww <= 0) { set_error("dv_dbg_conv3x3_halo64: bad argument"); return DV_ERR_INVALID; }
  const int C = 64;
  const size_t nx = (size_t)n * hh * ww * C;
  const int ho = pool ? hh / 2 : hh, wo = pool ? ww / 2 : ww;
  const size_t ny = (size_t)n * ho * wo * C;
  std::vector<float> xb(nx), wr((size_t)C * 9 * C), yb(ny);
  for (int b = 0; b < n; ++b)
    for (int yy = 0; yy < hh; ++yy)
      for (int xx = 0; xx < ww; ++xx)
        for (int c = 0; c < C; ++c)
          xb[((((size_t)b * 8 + c / 8) * hh + yy) * ww + xx) * 8 + c % 8] = x[(((size_t)b * hh + yy) * ww + xx) * C + c];
  for (int o = 0; o < C; ++o)
    for (int c = 0; c < C; ++c)
      for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * C + c] = wgt[((size_t)o * C + c) * 9 + t];
  float *dx32 = nullptr, *dw32 = nullptr, *dy32 = nullptr, *dbias = nullptr;
  __half *dx = nullptr, *dw = nullptr, *dy = nullptr;
  DV_CUDA_OK(sc.get(&dx32, nx * 4)); DV_CUDA_OK(sc.get(&dx, nx * 2));
  DV_CUDA_OK(sc.get(&dw32, wr.size() * 4)); DV_CUDA_OK(sc.get(&dw, wr.size() * 2));
  DV_CUDA_OK(sc.get(&dy, ny * 2 + 64)); DV_CUDA_OK(sc.get(&dy32, ny * 4));
  DV_CUDA_OK(sc.get(&dbias, C * 4));
  DV_CUDA_OK(cudaMemsetAsync(dy, 0, ny * 2, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dx32, xb.data(), nx * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dw32, wr.data(), wr.size() * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dbias, bias, C * 4, cudaMemcpyHostToDevice, e->st));
  f32_to_f16(dx32, dx, (int64_t)nx, e->st);
  f32_to_f16(dw32, dw, (int64_t)wr.size(), e->st);
  HaloPlan pl;
  int rc = plan_conv3x3_halo64(&pl, dx, n, hh, ww, dw, dbias, dy, out_blocked, relu, pool);
  if (!rc) rc = launch_conv_halo64(pl, n, e->st);
  if (!rc) {
    f16_to_f32(dy, dy32, (int64_t)ny, e->st);
    cudaError_t ce = cudaMemcpyAsync(yb.data(), dy32, ny * 4, cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("dv_dbg_conv3x3_halo64: ") + cudaGetErrorString(ce)); rc = DV_ERR_CUDA; }
  }
  if (rc) return (dv_status)rc;
  if (out_blocked) {
    for (int b = 0; b < n; ++b)
      for (int yy = 0; yy < ho; ++yy)
        for (int xx = 0; xx < wo; ++xx)
          for (int c = 0; c < C; ++c)
            y[(((size_t)b * ho + yy) * wo + xx) * C + c] = yb[((((size_t)b * 8 + c / 8) * ho + yy) * wo + xx) * 8 + c % 8];
  } else {
    memcpy(y, yb.data(), ny * 4);
  }
  return DV_OK;
}

dv_status dv_dbg_conv3x3_halo128(dv_engine* h, const float* x, const float* wgt, const float* bias, int32_t n, int32_t hh,
                                 int32_t ww, int32_t cin, int32_t cout, int32_t relu, int32_t pool, int32_t out_blocked,
                                 float* y) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Scratch sc;
  if (!x || !wgt || !bias || !y || n <= 0 || hh <= 0 || ww <= 0) { set_error("dv_dbg_conv3x3_halo128: bad argument"); return DV_ERR_INVALID; }
  const size_t nx = (size_t)n * hh * ww * cin;
  const int ho = pool ? hh / 2 : hh, wo = pool ? ww / 2 : ww;
  const size_t ny = (size_t)n * ho * wo * cout;
  std::vector<float> xb(nx), wr((size_t)cout * 9 * cin), yb(ny);
  for (int b = 0; b < n; ++b)
    for (int yy = 0; yy < hh; ++yy)
      for (int xx = 0; xx < ww; ++xx)
        for (int c = 0; c < cin; ++c)
          xb[((((size_t)b * (cin / 8) + c / 8) * hh + yy) * ww + xx) * 8 + c % 8] = x[(((size_t)b * hh + yy) * ww + xx) * cin + c];
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 9; ++t) wr[((size_t)o * 9 + t) * cin + c] = wgt[((size_t)o * cin + c) * 9 + t];
  float *dx32 = nullptr, *dw32 = nullptr, *dy32 = nullptr, *dbias = nullptr;
  __half *dx = nullptr, *dw = nullptr, *dy = nullptr;
  DV_CUDA_OK(sc.get(&dx32, nx * 4)); DV_CUDA_OK(sc.get(&dx, nx * 2));
  DV_CUDA_OK(sc.get(&dw32, wr.size() * 4)); DV_CUDA_OK(sc.get(&dw, wr.size() * 2));
  DV_CUDA_OK(sc.get(&dy, ny * 2 + 64)); DV_CUDA_OK(sc.get(&dy32, ny * 4));
  DV_CUDA_OK(sc.get(&dbias, (size_t)cout * 4));
  DV_CUDA_OK(cudaMemsetAsync(dy, 0, ny * 2, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dx32, xb.data(), nx * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dw32, wr.data(), wr.size() * 4, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dbias, bias, (size_t)cout * 4, cudaMemcpyHostToDevice, e->st));
  f32_to_f16(dx32, dx, (int64_t)nx, e->st);
  f32_to_f16(dw32, dw, (int64_t)wr.size(), e->st);
  Halo128Plan pl;
  int rc = plan_conv3x3_halo128(&pl, dx, n, hh, ww, cin, dw, cout, dbias, dy, out_blocked, relu, pool);
  if (!rc) rc = launch_conv_halo128(pl, n, e->st);
  if (!rc) {
    f16_to_f32(dy, dy32, (int64_t)ny, e->st);
    cudaError_t ce = cudaMemcpyAsync(yb.data(), dy32, ny * 4, cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("dv_dbg_conv3x3_halo128: ") + cudaGetErrorString(ce)); rc = DV_ERR_CUDA; }
  }
  if (rc) return (dv_status)rc;
  if (out_blocked) {
    for (int b = 0; b < n; ++b)
      for (int yy = 0; yy < ho; ++yy)
        for (int xx = 0; xx < wo; ++xx)
          for (int c = 0; c < cout; ++c)
            y[(((size_t)b * ho + yy) * wo + xx) * cout + c] =
                yb[((((size_t)b * (cout / 8) + c / 8) * ho + yy) * wo + xx) * 8 + c % 8];
  } else {
    memcpy(y, yb.data(), ny * 4);
  }
  return DV_OK;
}

dv_status dv_dbg_read(dv_engine* h, const char* name, float* dst, int64_t capacity, int64_t* count) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Scratch sc;
  auto it = e->dbg.find(name ? name : "");
  if (it == e->dbg.end()) { set_error(std::string("dv_dbg_read: unknown tensor ") + (name ? name : "(null)")); return DV_ERR_INVALID; }
  const Engine::Dbg& d = it->second;
  if (count) *count = d.n;
  if (!dst) return DV_OK;
  DV_TRY(sp_dbg_refresh(e, name));
  if (capacity < d.n) { set_error("dv_dbg_read: capacity too small"); return DV_ERR_CAPACITY; }
  if (d.dtype == 0 || d.dtype == 2) {
    DV_CUDA_OK(cudaMemcpyAsync(dst, d.p, (size_t)d.n * 4, cudaMemcpyDeviceToHost, e->st));
  } else {
    float* tmp = nullptr;
    DV_CUDA_OK(sc.get(&tmp, (size_t)d.n * 4));
    f16_to_f32(reinterpret_cast<const __half*>(d.p), tmp, d.n, e->st);
    cudaError_t ce = cudaMemcpyAsync(dst, tmp, (size_t)d.n * 4, cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("dv_dbg_read: ") + cudaGetErrorString(ce)); return DV_ERR_CUDA; }
  }
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  return DV_OK;
}

}  // extern "C"
