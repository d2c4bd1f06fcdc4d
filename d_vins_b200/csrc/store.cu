// Device-resident keyframe feature store + the batched keyframe-round API (throughput path, SURVEY §8(e),(f)-1).
// A keyframe's local features never leave the GPU between extraction and matching: kpts = SP ++ VIO points,
// desc = SP descriptors ++ SP_RE descriptors - the concatenation contract of keyframe.cpp:401-432.
#include <string.h>

#include <algorithm>

#include "engine.h"
#include "lg.h"

namespace dv {

int comm_allgather(Engine* e, const float* send, float* recv, size_t count_per_rank);   // comm.cpp
int comm_allgather_bytes(Engine* e, const void* send, void* recv, size_t bytes_per_rank);
float* bank_rows(Engine* e);
float* bank_query_buf(Engine* e);
int64_t& bank_size_ref(Engine* e);
int bank_search_device(Engine* e, int nq, const int64_t* nb_limit, int k, float* D_host, int64_t* I_host);

struct Store {
  int slots = 0, cap = 0;          // cap = max_kpts + max_vio rows per keyframe
  float* kpts = nullptr;           // [slots, cap, 2]
  float* desc = nullptr;           // [slots, cap, 256]
  std::vector<int64_t> frame_id;   // host metadata
  std::vector<int> n_sp, n_vio;
  // batch staging
  float* h_vio = nullptr; int* h_nvio = nullptr; int* h_nsp = nullptr; int* d_slot = nullptr; int* h_slot = nullptr;
  std::vector<int64_t> cur_ids;
  int cur_b = 0;
  // ---- multi-GPU: every rank's store is mapped here through CUDA IPC, and each round's all-gather also carries the
  // (frame id, n_sp, n_vio) of the new keyframes, so LightGlue can PULL an old keyframe's features from its owner rank
  // with a one-sided peer copy over NVLink (SURVEY §8(e) "LightGlue operand locality", option (ii)).
  std::vector<float*> peer_kpts, peer_desc;                 // [world]; own rank = local pointers
  std::vector<std::vector<int64_t>> r_frame_id;             // [world][slots]
  std::vector<std::vector<int>> r_n_sp, r_n_vio;            // [world][slots]
  float *send516 = nullptr, *recv516 = nullptr;             // [b,516] / [world*b,516] round buffers
  int *d_meta = nullptr, *h_meta = nullptr;                 // this rank's [b,4] meta (device / pinned)
  int *d_meta_all = nullptr, *h_meta_all = nullptr;         // gathered [world*b,4]
  float *cache_kpts = nullptr, *cache_desc = nullptr;       // [B, cap, *] landing zone of pulled keyframes
  bool peers_open = false;
};

// send row i = [ gdesc_i (512 f32) | frame id lo, hi, n_sp, n_vio (bit patterns) ]
__global__ void k_round_pack(const float* __restrict__ gdesc, const int* __restrict__ meta, float* __restrict__ send) {
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < 516; c += blockDim.x)
    send[(int64_t)i * 516 + c] = c < 512 ? gdesc[(int64_t)i * 512 + c] : __int_as_float(meta[i * 4 + c - 512]);
}
__global__ void k_round_unpack(const float* __restrict__ recv, float* __restrict__ bank_dst, int* __restrict__ meta_all) {
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < 516; c += blockDim.x) {
    const float v = recv[(int64_t)i * 516 + c];
    if (c < 512) bank_dst[(int64_t)i * 512 + c] = v;
    else meta_all[i * 4 + c - 512] = __float_as_int(v);
  }
}

// one block per (frame, row chunk): rows [0,n_sp) from the SuperPoint outputs, rows [n_sp, n_sp+n_vio) from SP_RE
__global__ void k_store_write(const float* __restrict__ sp_kpts, const float* __restrict__ sp_desc,
                              const int* __restrict__ n_sp, int K, const float* __restrict__ re_kpts,
                              const float* __restrict__ re_desc, const int* __restrict__ n_vio, int V,
                              const int* __restrict__ slot, float* __restrict__ st_kpts, float* __restrict__ st_desc,
                              int cap) {
  const int f = blockIdx.y;
  const int ns = n_sp[f], nv = n_vio[f];
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  float* ok = st_kpts + (int64_t)slot[f] * cap * 2;
  float* od = st_desc + (int64_t)slot[f] * cap * 256;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < ns + nv; r += gridDim.x * warps) {
    const float* sk; const float* sd;
    if (r < ns) { sk = sp_kpts + ((int64_t)f * K + r) * 2; sd = sp_desc + ((int64_t)f * K + r) * 256; }
    else { sk = re_kpts + ((int64_t)f * V + (r - ns)) * 2; sd = re_desc + ((int64_t)f * V + (r - ns)) * 256; }
    if (lane < 2) ok[r * 2 + lane] = sk[lane];
    const float4* s4 = reinterpret_cast<const float4*>(sd);
    float4* d4 = reinterpret_cast<float4*>(od + (int64_t)r * 256);
    d4[lane] = s4[lane];
    d4[lane + 32] = s4[lane + 32];
  }
}

int store_init(Engine* e) {
  Store* s = new Store();
  e->store = s;
  s->slots = e->cfg.store_capacity;
  s->cap = e->cfg.max_kpts + e->cfg.max_vio;
  s->frame_id.assign(s->slots, -1);
  s->n_sp.assign(s->slots, 0);
  s->n_vio.assign(s->slots, 0);
  DV_TRY(e->alloc(&s->kpts, (size_t)s->slots * s->cap * 2));
  DV_TRY(e->alloc(&s->desc, (size_t)s->slots * s->cap * 256));
  DV_TRY(e->alloc_pinned(&s->h_vio, (size_t)e->B * e->cfg.max_vio * 2));
  DV_TRY(e->alloc_pinned(&s->h_nvio, (size_t)e->B));
  DV_TRY(e->alloc_pinned(&s->h_nsp, (size_t)e->B));
  DV_TRY(e->alloc_pinned(&s->h_slot, (size_t)e->B));
  DV_TRY(e->alloc(&s->d_slot, (size_t)e->B));
  const int ws = e->cfg.world_size;
  s->peer_kpts.assign(ws, nullptr); s->peer_desc.assign(ws, nullptr);
  s->peer_kpts[e->cfg.rank] = s->kpts; s->peer_desc[e->cfg.rank] = s->desc;
  s->r_frame_id.assign(ws, std::vector<int64_t>(s->slots, -1));
  s->r_n_sp.assign(ws, std::vector<int>(s->slots, 0));
  s->r_n_vio.assign(ws, std::vector<int>(s->slots, 0));
  if (ws > 1) {
    DV_TRY(e->alloc(&s->send516, (size_t)e->B * 516));
    DV_TRY(e->alloc(&s->recv516, (size_t)ws * e->B * 516));
    DV_TRY(e->alloc(&s->d_meta, (size_t)e->B * 4));
    DV_TRY(e->alloc_pinned(&s->h_meta, (size_t)e->B * 4));
    DV_TRY(e->alloc(&s->d_meta_all, (size_t)ws * e->B * 4));
    DV_TRY(e->alloc_pinned(&s->h_meta_all, (size_t)ws * e->B * 4));
    DV_TRY(e->alloc(&s->cache_kpts, (size_t)e->B * s->cap * 2));
    DV_TRY(e->alloc(&s->cache_desc, (size_t)e->B * s->cap * 256));
  }
  return DV_OK;
}
void store_free(Engine* e) {
  Store* s = e->store;
  if (s && s->peers_open)
    for (int r = 0; r < e->cfg.world_size; ++r)
      if (r != e->cfg.rank) {
        if (s->peer_kpts[r]) cudaIpcCloseMemHandle(s->peer_kpts[r]);
        if (s->peer_desc[r]) cudaIpcCloseMemHandle(s->peer_desc[r]);
      }
  delete e->store;
  e->store = nullptr;
}

// Exchange CUDA-IPC handles of the store buffers through the (already initialised) NCCL communicator and map every
// peer's store into this process.
int store_exchange_peers(Engine* e) {
  Store* s = e->store;
  const int ws = e->cfg.world_size, me = e->cfg.rank;
  if (ws <= 1 || s->peers_open) return DV_OK;
  struct Handles { cudaIpcMemHandle_t k, d; };
  Handles mine;
  DV_CUDA_OK(cudaIpcGetMemHandle(&mine.k, s->kpts));
  DV_CUDA_OK(cudaIpcGetMemHandle(&mine.d, s->desc));
  Handles *d_send = nullptr, *d_recv = nullptr;
  DV_CUDA_OK(cudaMalloc(&d_send, sizeof(Handles)));
  DV_CUDA_OK(cudaMalloc(&d_recv, sizeof(Handles) * ws));
  DV_CUDA_OK(cudaMemcpyAsync(d_send, &mine, sizeof(Handles), cudaMemcpyHostToDevice, e->st));
  int rc = comm_allgather_bytes(e, d_send, d_recv, sizeof(Handles));
  std::vector<Handles> all(ws);
  if (!rc) {
    cudaError_t ce = cudaMemcpyAsync(all.data(), d_recv, sizeof(Handles) * ws, cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    if (ce != cudaSuccess) { set_error(std::string("store_exchange_peers: ") + cudaGetErrorString(ce)); rc = DV_ERR_CUDA; }
  }
  cudaFree(d_send); cudaFree(d_recv);
  if (rc) return rc;
  for (int r = 0; r < ws; ++r) {
    if (r == me) continue;
    void *pk = nullptr, *pd = nullptr;
    DV_CUDA_OK(cudaIpcOpenMemHandle(&pk, all[r].k, cudaIpcMemLazyEnablePeerAccess));
    DV_CUDA_OK(cudaIpcOpenMemHandle(&pd, all[r].d, cudaIpcMemLazyEnablePeerAccess));
    s->peer_kpts[r] = reinterpret_cast<float*>(pk);
    s->peer_desc[r] = reinterpret_cast<float*>(pd);
  }
  s->peers_open = true;
  return DV_OK;
}

}  // namespace dv

using namespace dv;
#define DV_CHECK_ENGINE(e) do { if (!(e)) { dv::set_error("null engine"); return DV_ERR_INVALID; } } while (0)

extern "C" {

dv_status dv_batch_upload(dv_engine* h, int32_t b, const uint8_t* imgs, int64_t frame_stride, int32_t stride) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!imgs || b < 1 || b > e->B || stride < e->W || frame_stride < (int64_t)stride * e->H) { set_error("dv_batch_upload: bad arguments"); return DV_ERR_INVALID; }
  const size_t fb = (size_t)e->H * e->W;
  // asynchronous on the copy stream into the buffer the compute stream is not using: the caller may queue round R+1's
  // frames while round R is still matching (the source must stay valid until the next dv_batch_extract / dv_sync)
  uint8_t* dst = e->image_begin_upload();
  if (stride == e->W && frame_stride == (int64_t)fb) {
    // contiguous (typically already pinned by the caller): one async copy straight from the caller's buffer
    DV_CUDA_OK(cudaMemcpyAsync(dst, imgs, fb * b, cudaMemcpyHostToDevice, e->st_copy));
  } else {
    DV_CUDA_OK(cudaMemcpy2DAsync(dst, e->W, imgs, stride, e->W, (size_t)e->H, cudaMemcpyHostToDevice, e->st_copy));
    for (int i = 1; i < b; ++i)
      DV_CUDA_OK(cudaMemcpy2DAsync(dst + fb * i, e->W, imgs + frame_stride * i, stride, e->W, (size_t)e->H, cudaMemcpyHostToDevice, e->st_copy));
  }
  e->image_end_upload();
  e->img_ch = 1;
  e->next_b = b;
  e->next_pending = true;
  return DV_OK;
}

dv_status dv_batch_extract(dv_engine* h, int32_t b, const float* vio_xy, const int32_t* n_vio, const int64_t* frame_ids) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (!e->sp || !e->mix) { set_error("dv_batch_extract: engine created without weights"); return DV_ERR_INVALID; }
  e->adopt_upload();
  if (b < 1 || b != e->cur_b || !vio_xy || !n_vio || !frame_ids) { set_error("dv_batch_extract: b must equal the uploaded batch"); return DV_ERR_INVALID; }
  const int V = e->cfg.max_vio, K = e->cfg.max_kpts;
  for (int i = 0; i < b; ++i) {
    if (n_vio[i] < 0 || n_vio[i] > V || frame_ids[i] < 0) { set_error("dv_batch_extract: n_vio / frame id out of range"); return DV_ERR_INVALID; }
    s->h_nvio[i] = n_vio[i];
    s->h_slot[i] = (int)(frame_ids[i] % s->slots);
  }
  for (int i = 0; i < b; ++i)
    for (int j = i + 1; j < b; ++j)
      if (s->h_slot[i] == s->h_slot[j]) { set_error("dv_batch_extract: store_capacity too small for this batch"); return DV_ERR_CAPACITY; }
  memcpy(s->h_vio, vio_xy, sizeof(float) * 2 * V * b);
  float *d_rk, *d_rd, *d_kf, *d_de; int *d_rn, *d_n;
  sp_device_results(e, nullptr, &d_kf, nullptr, &d_n, &d_de, &d_rk, &d_rn, &d_rd);
  {
    StageScope sc(e, ST_COPY);
    DV_CUDA_OK(cudaMemcpyAsync(d_rk, s->h_vio, sizeof(float) * 2 * V * b, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_rn, s->h_nvio, sizeof(int) * b, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(s->d_slot, s->h_slot, sizeof(int) * b, cudaMemcpyHostToDevice, e->st));
  }
  DV_TRY(sp_run_encoder(e, b));
  DV_TRY(sp_run_detect(e, b));
  DV_TRY(sp_run_describe(e, b, d_rk, d_rn, V, d_rd));     // same encoder pass (the reference runs it twice)
  DV_TRY(mix_run(e, b));
  e->enc_done = e->det_done = e->mix_done = true;
  {
    StageScope sc(e, ST_SP_POST);
    k_store_write<<<dim3(8, b), 256, 0, e->st>>>(d_kf, d_de, d_n, K, d_rk, d_rd, d_rn, V, s->d_slot, s->kpts, s->desc, s->cap);
    DV_CUDA_OK(cudaGetLastError());
    DV_LAUNCHED(e, 1);
  }
  {
    StageScope sc(e, ST_COPY);
    DV_CUDA_OK(cudaMemcpyAsync(s->h_nsp, d_n, sizeof(int) * b, cudaMemcpyDeviceToHost, e->st));
  }
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  s->cur_ids.assign(frame_ids, frame_ids + b);
  s->cur_b = b;
  for (int i = 0; i < b; ++i) {
    const int sl = s->h_slot[i];
    s->frame_id[sl] = frame_ids[i];
    s->n_sp[sl] = s->h_nsp[i];
    s->n_vio[sl] = n_vio[i];
  }
  return DV_OK;
}

dv_status dv_batch_commit(dv_engine* h, int32_t b, int64_t* first_row) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!e->mix || b < 1 || b != e->cur_b || !e->mix_done) { set_error("dv_batch_commit: no extracted batch of this size"); return DV_ERR_INVALID; }
  const int ws = e->cfg.world_size;
  int64_t& size = bank_size_ref(e);
  if (size + (int64_t)ws * b > e->cfg.bank_capacity) { set_error("dv_batch_commit: bank full"); return DV_ERR_CAPACITY; }
  float* g = mix_gdesc(e);
  StageScope sc(e, ST_KNN);
  // this round's queries
  DV_CUDA_OK(cudaMemcpyAsync(bank_query_buf(e), g, sizeof(float) * 512 * b, cudaMemcpyDeviceToDevice, e->st));
  float* dst = bank_rows(e) + size * 512;
  if (ws == 1) {
    DV_CUDA_OK(cudaMemcpyAsync(dst, g, sizeof(float) * 512 * b, cudaMemcpyDeviceToDevice, e->st));
  } else {
    // the single collective of the path: [b, 512 + 4] rows per rank (global descriptor + keyframe id / point counts),
    // gathered rank-major == global frame order; unpacked straight into the bank tail.
    Store* s = e->store;
    for (int i = 0; i < b; ++i) {
      const int64_t fid = s->cur_ids[i];
      const int sl = (int)(fid % s->slots);
      s->h_meta[i * 4 + 0] = (int)(fid & 0xffffffffll); s->h_meta[i * 4 + 1] = (int)(fid >> 32);
      s->h_meta[i * 4 + 2] = s->n_sp[sl]; s->h_meta[i * 4 + 3] = s->n_vio[sl];
    }
    DV_CUDA_OK(cudaMemcpyAsync(s->d_meta, s->h_meta, sizeof(int) * 4 * b, cudaMemcpyHostToDevice, e->st));
    k_round_pack<<<b, 128, 0, e->st>>>(g, s->d_meta, s->send516);
    DV_TRY(comm_allgather(e, s->send516, s->recv516, (size_t)516 * b));
    k_round_unpack<<<ws * b, 128, 0, e->st>>>(s->recv516, dst, s->d_meta_all);
    DV_CUDA_OK(cudaGetLastError());
    DV_LAUNCHED(e, 3);
    DV_CUDA_OK(cudaMemcpyAsync(s->h_meta_all, s->d_meta_all, sizeof(int) * 4 * ws * b, cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaStreamSynchronize(e->st));
    for (int r = 0; r < ws; ++r)
      for (int i = 0; i < b; ++i) {
        const int* m = s->h_meta_all + ((size_t)r * b + i) * 4;
        const int64_t fid = ((int64_t)m[1] << 32) | (uint32_t)m[0];
        const int sl = (int)(fid % s->slots);
        s->r_frame_id[r][sl] = fid; s->r_n_sp[r][sl] = m[2]; s->r_n_vio[r][sl] = m[3];
      }
  }
  if (first_row) *first_row = size + (int64_t)e->cfg.rank * b;
  size += (int64_t)ws * b;
  return DV_OK;
}

dv_status dv_batch_search(dv_engine* h, int32_t b, const int64_t* nb_limit, float* D, int64_t* I) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (b < 1 || b > e->B || !nb_limit || !D || !I) { set_error("dv_batch_search: bad arguments"); return DV_ERR_INVALID; }
  return (dv_status)bank_search_device(e, b, nb_limit, e->cfg.knn_k, D, I);
}

dv_status dv_batch_read_global(dv_engine* h, int32_t i, float* des512) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!e->mix || !e->mix_done || i < 0 || i >= e->cur_b || !des512) { set_error("dv_batch_read_global: no such frame"); return DV_ERR_INVALID; }
  DV_CUDA_OK(cudaMemcpyAsync(des512, mix_gdesc(e) + (size_t)i * 512, sizeof(float) * 512, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  return DV_OK;
}

dv_status dv_store_read(dv_engine* h, int64_t frame_id, float* kpts_xy, float* desc, int32_t* n_total, int32_t* n_sp) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (frame_id < 0) { set_error("dv_store_read: bad frame id"); return DV_ERR_INVALID; }
  const int sl = (int)(frame_id % s->slots);
  if (s->frame_id[sl] != frame_id) { set_error("dv_store_read: keyframe not resident in this rank's store"); return DV_ERR_INVALID; }
  const int n = s->n_sp[sl] + s->n_vio[sl];
  if (kpts_xy) DV_CUDA_OK(cudaMemcpyAsync(kpts_xy, s->kpts + (size_t)sl * s->cap * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, e->st));
  if (desc) DV_CUDA_OK(cudaMemcpyAsync(desc, s->desc + (size_t)sl * s->cap * 256, sizeof(float) * 256 * n, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  if (n_total) *n_total = n;
  if (n_sp) *n_sp = s->n_sp[sl];
  return DV_OK;
}

dv_status dv_store_put(dv_engine* h, int64_t frame_id, const float* kpts_xy, const float* desc, int32_t n_total,
                       int32_t n_sp) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (frame_id < 0 || !kpts_xy || !desc || n_sp < 0 || n_total < n_sp || n_total > s->cap ||
      n_sp > e->cfg.max_kpts || n_total - n_sp > e->cfg.max_vio) {
    set_error("dv_store_put: bad arguments (n_sp <= max_kpts, n_total - n_sp <= max_vio)");
    return DV_ERR_INVALID;
  }
  const int sl = (int)(frame_id % s->slots);
  DV_CUDA_OK(cudaMemcpyAsync(s->kpts + (size_t)sl * s->cap * 2, kpts_xy, sizeof(float) * 2 * n_total, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(s->desc + (size_t)sl * s->cap * 256, desc, sizeof(float) * 256 * n_total, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  s->frame_id[sl] = frame_id;
  s->n_sp[sl] = n_sp;
  s->n_vio[sl] = n_total - n_sp;
  return DV_OK;
}

dv_status dv_batch_match(dv_engine* h, int32_t b, const int64_t* query_ids, const int64_t* old_ids, int32_t* matches,
                         float* mscores, int32_t* k_out) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (!e->lg) { set_error("dv_batch_match: engine created without weights"); return DV_ERR_INVALID; }
  if (b < 1 || b > e->B || !query_ids || !old_ids || !matches || !mscores || !k_out) { set_error("dv_batch_match: bad arguments"); return DV_ERR_INVALID; }
  const int V = e->cfg.max_vio;
  std::vector<LgSeg> segs;
  std::vector<int> which;
  for (int i = 0; i < b; ++i) {
    k_out[i] = 0;
    if (query_ids[i] < 0 || old_ids[i] < 0) { set_error("dv_batch_match: negative keyframe id"); return DV_ERR_INVALID; }
    const int qs = (int)(query_ids[i] % s->slots), os = (int)(old_ids[i] % s->slots);
    if (s->frame_id[qs] != query_ids[i]) { set_error("dv_batch_match: query keyframe not resident in this rank's store"); return DV_ERR_INVALID; }
    // the old keyframe: local store, else the owner rank's store (one-sided pull over NVLink)
    int owner = -1, n_sp_old = 0, n_vio_old = 0;
    if (s->frame_id[os] == old_ids[i]) { owner = e->cfg.rank; n_sp_old = s->n_sp[os]; n_vio_old = s->n_vio[os]; }
    else if (s->peers_open)
      for (int r = 0; r < e->cfg.world_size && owner < 0; ++r)
        if (r != e->cfg.rank && s->r_frame_id[r][os] == old_ids[i]) { owner = r; n_sp_old = s->r_n_sp[r][os]; n_vio_old = s->r_n_vio[r][os]; }
    if (owner < 0) { set_error("dv_batch_match: old keyframe not resident on any rank's store"); return DV_ERR_INVALID; }
    const int m = s->n_vio[qs], n = n_sp_old + n_vio_old;
    // keyframe.cpp:373,:935 - the reference skips SP_RE / LightGlue for <= 20 window points; engine floor is 10
    if (m < 10 || n < 10) continue;
    if (n > e->cfg.lg_max_kpts) { set_error("dv_batch_match: old keyframe exceeds lg_max_kpts"); return DV_ERR_CAPACITY; }
    const float* qk = s->kpts + ((size_t)qs * s->cap + s->n_sp[qs]) * 2;
    const float* qd = s->desc + ((size_t)qs * s->cap + s->n_sp[qs]) * 256;
    segs.push_back({qk, qd, m, e->W, e->H, 0});
    const float* ok = s->kpts + (size_t)os * s->cap * 2;
    const float* od = s->desc + (size_t)os * s->cap * 256;
    if (owner != e->cfg.rank) {
      StageScope sc(e, ST_COPY);
      float* ck = s->cache_kpts + (size_t)which.size() * s->cap * 2;
      float* cd = s->cache_desc + (size_t)which.size() * s->cap * 256;
      DV_CUDA_OK(cudaMemcpyAsync(ck, s->peer_kpts[owner] + (size_t)os * s->cap * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToDevice, e->st));
      DV_CUDA_OK(cudaMemcpyAsync(cd, s->peer_desc[owner] + (size_t)os * s->cap * 256, sizeof(float) * 256 * n, cudaMemcpyDeviceToDevice, e->st));
      ok = ck; od = cd;
    }
    segs.push_back({ok, od, n, e->W, e->H, 0});
    which.push_back(i);
  }
  if (which.empty()) return DV_OK;
  DV_TRY(lg_run(e, (int)which.size(), segs.data()));
  return (dv_status)lg_fetch_batch(e, (int)which.size(), V, which.data(), matches, mscores, k_out);
}

}  // extern "C"
