// Device-resident keyframe feature store + the batched keyframe-round API (throughput path, SURVEY §8(e),(f)-1).
// A keyframe's local features never leave the GPU between extraction and matching: kpts = SP ++ VIO points,
// desc = SP descriptors ++ SP_RE descriptors - the concatenation contract of keyframe.cpp:401-432.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <unordered_map>
#include <vector>

#include "engine.h"
#include "lg.h"

#define DV_LG_CHUNK_DEFAULT 0

namespace dv {

int comm_allgather(Engine* e, const float* send, float* recv, size_t count_per_rank);   // comm.cpp
int comm_allgather_bytes(Engine* e, const void* send, void* recv, size_t bytes_per_rank);
int comm_poll(Engine* e);                                                                // ncclCommGetAsyncError
float* bank_rows(Engine* e);
float* bank_query_buf(Engine* e);
int64_t& bank_size_ref(Engine* e);
int bank_search_device(Engine* e, int nq, const int64_t* nb_limit, int k, float* D_host, int64_t* I_host);

struct Loc { int rank, slot, n_sp, n_vio; };

struct Store {
  int slots = 0, cap = 0;          // cap = max_kpts + max_vio rows per keyframe
  float* kpts = nullptr;           // [slots, cap, 2]
  float* desc = nullptr;           // [slots, cap, 256]
  // Per-slot header on the device {frame id lo, hi, n_sp, n_vio}; id = -1 while the slot is being rewritten.  Peers
  // read it through the CUDA-IPC mapping AFTER they consumed the slot's features (seqlock: see k_pull_check).
  int* hdr = nullptr;              // [slots, 4]
  // host directory of every keyframe this rank knows about (its own and, via the round all-gather / dv_store_sync,
  // every peer's): frame id -> (rank, slot, point counts).  Slots are handed out round-robin (ring); the keyframe a
  // slot held before is forgotten when the slot is reused, locally and - through the metadata - on every peer.
  std::unordered_map<int64_t, Loc> where;
  std::vector<std::vector<int64_t>> slot_fid;               // [world][slots]: frame id held by (rank, slot), -1 = none
  int next_slot = 0;
  // batch staging
  float* h_vio = nullptr; int* h_nvio = nullptr; int* h_nsp = nullptr; int* d_slot = nullptr; int* h_slot = nullptr;
  int* d_fid = nullptr; int* h_fid = nullptr;               // [B,2] frame ids of the batch as (lo, hi)
  std::vector<int64_t> cur_ids;
  int cur_b = 0;
  // ---- multi-GPU: every rank's store is mapped here through CUDA IPC, and each round's all-gather also carries the
  // (frame id, slot, n_sp, n_vio) of the new keyframes, so LightGlue can read an old keyframe's features straight from
  // its owner rank's memory over NVLink (SURVEY §8(e) "LightGlue operand locality", option (ii)) - no staging copy.
  std::vector<float*> peer_kpts, peer_desc;                 // [world]; own rank = local pointers
  std::vector<int*> peer_hdr;
  float *send516 = nullptr, *recv516 = nullptr;             // [b,516] / [world*b,516] round buffers
  int *d_meta = nullptr, *h_meta = nullptr;                 // this rank's [b,4] meta (device / pinned)
  int *d_meta_all = nullptr, *h_meta_all = nullptr;         // gathered [world*b,4]
  int *d_hdr_all = nullptr, *h_hdr_all = nullptr;           // dv_store_sync: gathered headers [world, slots, 4]
  // seqlock verification of remote reads: per pair {peer header pointer, expected id} -> flag
  // dv_batch_match_begin .. dv_batch_match_end: the pairs in flight
  bool mp_pending = false;
  int mp_b = 0, mp_cap = 0, mp_npull = 0;
  std::vector<int> mp_kout, mp_which, mp_pull_pair;
  struct PullJob { const int* hdr; int lo, hi, pad; };
  PullJob *d_pull = nullptr, *h_pull = nullptr;             // [2B]
  int *d_pull_bad = nullptr, *h_pull_bad = nullptr;         // [2B]
  bool peers_open = false;

  void forget(int rank, int slot) {
    int64_t& f = slot_fid[rank][slot];
    if (f >= 0) {
      auto it = where.find(f);
      if (it != where.end() && it->second.rank == rank && it->second.slot == slot) where.erase(it);
    }
    f = -1;
  }
  void remember(int rank, int slot, int64_t fid, int n_sp, int n_vio) {
    if (slot_fid[rank][slot] != fid) forget(rank, slot);
    auto it = where.find(fid);
    if (it != where.end() && (it->second.rank != rank || it->second.slot != slot)) slot_fid[it->second.rank][it->second.slot] = -1;
    slot_fid[rank][slot] = fid;
    where[fid] = Loc{rank, slot, n_sp, n_vio};
  }
  // slot for a keyframe this rank is about to (re)write: its current slot if it is already resident here, else the
  // next slot of the ring
  int assign_slot(int rank, int64_t fid) {
    auto it = where.find(fid);
    if (it != where.end() && it->second.rank == rank) return it->second.slot;
    const int sl = next_slot;
    next_slot = (next_slot + 1) % slots;
    return sl;
  }
};

// send row i = [ gdesc_i (512 f32) | frame id lo, hi, slot, n_sp << 16 | n_vio (bit patterns) ]
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

// Seqlock around a slot rewrite.  k_store_begin runs BEFORE the encoder of the round (milliseconds before the first
// byte of the slot changes) and marks the slots invalid; k_store_commit runs after k_store_write and publishes the new
// keyframe.  A peer that read the slot and THEN still finds the expected id in the header (k_pull_check) therefore read
// it entirely before the rewrite began.
__global__ void k_store_begin(int* __restrict__ hdr, const int* __restrict__ slot, int b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  hdr[slot[i] * 4 + 0] = -1;
  hdr[slot[i] * 4 + 1] = -1;
  __threadfence_system();
}
__global__ void k_store_commit(int* __restrict__ hdr, const int* __restrict__ slot, const int* __restrict__ fid,
                               const int* __restrict__ n_sp, const int* __restrict__ n_vio, int* __restrict__ meta,
                               int b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  int* h = hdr + slot[i] * 4;
  h[2] = n_sp[i]; h[3] = n_vio[i];
  __threadfence_system();
  h[0] = fid[2 * i]; h[1] = fid[2 * i + 1];
  if (meta) { meta[i * 4 + 0] = fid[2 * i]; meta[i * 4 + 1] = fid[2 * i + 1]; meta[i * 4 + 2] = slot[i]; meta[i * 4 + 3] = (n_sp[i] << 16) | n_vio[i]; }
}
__global__ void k_pull_check(const Store::PullJob* __restrict__ jobs, int n, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const volatile int* h = jobs[i].hdr;
  bad[i] = (h[0] != jobs[i].lo || h[1] != jobs[i].hi) ? 1 : 0;
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
  const int ws = e->cfg.world_size;
  s->slot_fid.assign(ws, std::vector<int64_t>(s->slots, -1));
  DV_TRY(e->alloc(&s->kpts, (size_t)s->slots * s->cap * 2));
  DV_TRY(e->alloc(&s->desc, (size_t)s->slots * s->cap * 256));
  DV_TRY(e->alloc(&s->hdr, (size_t)s->slots * 4));
  DV_CUDA_OK(cudaMemsetAsync(s->hdr, 0xff, sizeof(int) * 4 * s->slots, e->st));
  DV_TRY(e->alloc_pinned(&s->h_vio, (size_t)e->B * e->cfg.max_vio * 2));
  DV_TRY(e->alloc_pinned(&s->h_nvio, (size_t)e->B));
  DV_TRY(e->alloc_pinned(&s->h_nsp, (size_t)e->B));
  DV_TRY(e->alloc_pinned(&s->h_slot, (size_t)e->B));
  DV_TRY(e->alloc_pinned(&s->h_fid, (size_t)e->B * 2));
  DV_TRY(e->alloc(&s->d_slot, (size_t)e->B));
  DV_TRY(e->alloc(&s->d_fid, (size_t)e->B * 2));
  DV_TRY(e->alloc(&s->d_meta, (size_t)e->B * 4));
  s->peer_kpts.assign(ws, nullptr); s->peer_desc.assign(ws, nullptr); s->peer_hdr.assign(ws, nullptr);
  s->peer_kpts[e->cfg.rank] = s->kpts; s->peer_desc[e->cfg.rank] = s->desc; s->peer_hdr[e->cfg.rank] = s->hdr;
  if (ws > 1) {
    DV_TRY(e->alloc(&s->send516, (size_t)e->B * 516));
    DV_TRY(e->alloc(&s->recv516, (size_t)ws * e->B * 516));
    DV_TRY(e->alloc(&s->d_meta_all, (size_t)ws * e->B * 4));
    DV_TRY(e->alloc_pinned(&s->h_meta_all, (size_t)ws * e->B * 4));
    DV_TRY(e->alloc(&s->d_hdr_all, (size_t)ws * s->slots * 4));
    DV_TRY(e->alloc_pinned(&s->h_hdr_all, (size_t)ws * s->slots * 4));
    DV_TRY(e->alloc(&s->d_pull, (size_t)e->B * 2));
    DV_TRY(e->alloc_pinned(&s->h_pull, (size_t)e->B * 2));
    DV_TRY(e->alloc(&s->d_pull_bad, (size_t)e->B * 2));
    DV_TRY(e->alloc_pinned(&s->h_pull_bad, (size_t)e->B * 2));
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
        if (s->peer_hdr[r]) cudaIpcCloseMemHandle(s->peer_hdr[r]);
      }
  delete e->store;
  e->store = nullptr;
}

namespace {
struct DevTmp {   // scope-owned cudaMalloc (early returns must not leak)
  void* p = nullptr;
  ~DevTmp() { if (p) cudaFree(p); }
};
}  // namespace

// Collective: every rank publishes its slot headers; each rank rebuilds its directory of REMOTE keyframes from them.
// Needed after dv_store_put / session reload (keyframes that never went through a round all-gather).
int store_sync_directory(Engine* e) {
  Store* s = e->store;
  const int ws = e->cfg.world_size, me = e->cfg.rank;
  if (ws <= 1) return DV_OK;
  DV_TRY(comm_allgather_bytes(e, s->hdr, s->d_hdr_all, sizeof(int) * 4 * s->slots));
  DV_CUDA_OK(cudaMemcpyAsync(s->h_hdr_all, s->d_hdr_all, sizeof(int) * 4 * s->slots * ws, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  DV_TRY(comm_poll(e));
  for (int r = 0; r < ws; ++r) {
    if (r == me) continue;
    for (int sl = 0; sl < s->slots; ++sl) {
      const int* h = s->h_hdr_all + ((size_t)r * s->slots + sl) * 4;
      const int64_t fid = ((int64_t)h[1] << 32) | (uint32_t)h[0];
      if (h[1] < 0) s->forget(r, sl);
      else s->remember(r, sl, fid, h[2], h[3]);
    }
  }
  return DV_OK;
}

// Exchange CUDA-IPC handles of the store buffers through the (already initialised) NCCL communicator and map every
// peer's store into this process.
int store_exchange_peers(Engine* e) {
  Store* s = e->store;
  const int ws = e->cfg.world_size, me = e->cfg.rank;
  if (ws <= 1 || s->peers_open) return DV_OK;
  struct Handles { cudaIpcMemHandle_t k, d, h; };
  Handles mine;
  DV_CUDA_OK(cudaIpcGetMemHandle(&mine.k, s->kpts));
  DV_CUDA_OK(cudaIpcGetMemHandle(&mine.d, s->desc));
  DV_CUDA_OK(cudaIpcGetMemHandle(&mine.h, s->hdr));
  DevTmp d_send, d_recv;
  DV_CUDA_OK(cudaMalloc(&d_send.p, sizeof(Handles)));
  DV_CUDA_OK(cudaMalloc(&d_recv.p, sizeof(Handles) * ws));
  DV_CUDA_OK(cudaMemcpyAsync(d_send.p, &mine, sizeof(Handles), cudaMemcpyHostToDevice, e->st));
  DV_TRY(comm_allgather_bytes(e, d_send.p, d_recv.p, sizeof(Handles)));
  std::vector<Handles> all(ws);
  DV_CUDA_OK(cudaMemcpyAsync(all.data(), d_recv.p, sizeof(Handles) * ws, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  for (int r = 0; r < ws; ++r) {
    if (r == me) continue;
    void *pk = nullptr, *pd = nullptr, *ph = nullptr;
    DV_CUDA_OK(cudaIpcOpenMemHandle(&pk, all[r].k, cudaIpcMemLazyEnablePeerAccess));
    s->peer_kpts[r] = reinterpret_cast<float*>(pk);
    DV_CUDA_OK(cudaIpcOpenMemHandle(&pd, all[r].d, cudaIpcMemLazyEnablePeerAccess));
    s->peer_desc[r] = reinterpret_cast<float*>(pd);
    DV_CUDA_OK(cudaIpcOpenMemHandle(&ph, all[r].h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_hdr[r] = reinterpret_cast<int*>(ph);
    s->peers_open = true;
  }
  s->peers_open = true;
  return store_sync_directory(e);
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
  if (b > s->slots) { set_error("dv_batch_extract: store_capacity smaller than the batch"); return DV_ERR_CAPACITY; }
  const int V = e->cfg.max_vio, K = e->cfg.max_kpts, me = e->cfg.rank;
  for (int i = 0; i < b; ++i) {
    if (n_vio[i] < 0 || n_vio[i] > V || frame_ids[i] < 0) { set_error("dv_batch_extract: n_vio / frame id out of range"); return DV_ERR_INVALID; }
    for (int j = 0; j < i; ++j)
      if (frame_ids[j] == frame_ids[i]) { set_error("dv_batch_extract: duplicate frame id in the batch"); return DV_ERR_INVALID; }
    s->h_nvio[i] = n_vio[i];
    s->h_fid[2 * i] = (int)(frame_ids[i] & 0xffffffffll); s->h_fid[2 * i + 1] = (int)(frame_ids[i] >> 32);
  }
  // slots: a keyframe already resident here keeps its slot; the others take the next slots of the ring, skipping
  // slots claimed by this very batch
  {
    std::vector<char> taken(s->slots, 0);
    for (int i = 0; i < b; ++i) {
      auto it = s->where.find(frame_ids[i]);
      s->h_slot[i] = (it != s->where.end() && it->second.rank == me) ? it->second.slot : -1;
      if (s->h_slot[i] >= 0) taken[s->h_slot[i]] = 1;
    }
    for (int i = 0; i < b; ++i) {
      if (s->h_slot[i] >= 0) continue;
      while (taken[s->next_slot]) s->next_slot = (s->next_slot + 1) % s->slots;
      s->h_slot[i] = s->next_slot;
      taken[s->next_slot] = 1;
      s->next_slot = (s->next_slot + 1) % s->slots;
    }
  }
  memcpy(s->h_vio, vio_xy, sizeof(float) * 2 * V * b);
  float *d_rk, *d_rd, *d_kf, *d_de; int *d_rn, *d_n;
  sp_device_results(e, nullptr, &d_kf, nullptr, &d_n, &d_de, &d_rk, &d_rn, &d_rd);
  {
    StageScope sc(e, ST_COPY);
    DV_CUDA_OK(cudaMemcpyAsync(d_rk, s->h_vio, sizeof(float) * 2 * V * b, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_rn, s->h_nvio, sizeof(int) * b, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(s->d_slot, s->h_slot, sizeof(int) * b, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(s->d_fid, s->h_fid, sizeof(int) * 2 * b, cudaMemcpyHostToDevice, e->st));
  }
  k_store_begin<<<cdiv(b, 128), 128, 0, e->st>>>(s->hdr, s->d_slot, b);
  DV_CUDA_OK(cudaGetLastError());
  DV_TRY(sp_run_encoder(e, b));
  DV_TRY(sp_run_detect(e, b));
  DV_TRY(sp_run_describe(e, b, d_rk, d_rn, V, d_rd));     // same encoder pass (the reference runs it twice)
  DV_TRY(mix_run(e, b));
  e->enc_done = e->det_done = e->mix_done = true;
  {
    StageScope sc(e, ST_SP_POST);
    k_store_write<<<dim3(8, b), 256, 0, e->st>>>(d_kf, d_de, d_n, K, d_rk, d_rd, d_rn, V, s->d_slot, s->kpts, s->desc, s->cap);
    k_store_commit<<<cdiv(b, 128), 128, 0, e->st>>>(s->hdr, s->d_slot, s->d_fid, d_n, d_rn, s->d_meta, b);
    DV_CUDA_OK(cudaGetLastError());
    DV_LAUNCHED(e, 3);
  }
  {
    StageScope sc(e, ST_COPY);
    DV_CUDA_OK(cudaMemcpyAsync(s->h_nsp, d_n, sizeof(int) * b, cudaMemcpyDeviceToHost, e->st));
  }
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  s->cur_ids.assign(frame_ids, frame_ids + b);
  s->cur_b = b;
  for (int i = 0; i < b; ++i) s->remember(me, s->h_slot[i], frame_ids[i], s->h_nsp[i], n_vio[i]);
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
    // the single collective of the path: [b, 512 + 4] rows per rank (global descriptor + keyframe id / slot / point
    // counts, written on the device by k_store_commit), gathered rank-major == global frame order; unpacked straight
    // into the bank tail.
    Store* s = e->store;
    if (s->cur_b != b) { set_error("dv_batch_commit: world_size > 1 needs a dv_batch_extract of the same batch first"); return DV_ERR_INVALID; }
    k_round_pack<<<b, 128, 0, e->st>>>(g, s->d_meta, s->send516);
    DV_TRY(comm_allgather(e, s->send516, s->recv516, (size_t)516 * b));
    k_round_unpack<<<ws * b, 128, 0, e->st>>>(s->recv516, dst, s->d_meta_all);
    DV_CUDA_OK(cudaGetLastError());
    DV_LAUNCHED(e, 3);
    DV_CUDA_OK(cudaMemcpyAsync(s->h_meta_all, s->d_meta_all, sizeof(int) * 4 * ws * b, cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaStreamSynchronize(e->st));
    DV_TRY(comm_poll(e));
    for (int r = 0; r < ws; ++r) {
      if (r == e->cfg.rank) continue;
      for (int i = 0; i < b; ++i) {
        const int* m = s->h_meta_all + ((size_t)r * b + i) * 4;
        const int64_t fid = ((int64_t)m[1] << 32) | (uint32_t)m[0];
        if (m[1] < 0 || m[2] < 0 || m[2] >= s->slots) continue;
        s->remember(r, m[2], fid, m[3] >> 16, m[3] & 0xffff);
      }
    }
  }
  if (first_row) *first_row = size + (int64_t)e->cfg.rank * b;
  size += (int64_t)ws * b;
  return DV_OK;
}

dv_status dv_batch_search(dv_engine* h, int32_t b, const int64_t* nb_limit, float* D, int64_t* I) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (b < 1 || b > e->B || !D || !I) { set_error("dv_batch_search: bad arguments"); return DV_ERR_INVALID; }
  if (nb_limit) return (dv_status)bank_search_device(e, b, nb_limit, e->cfg.knn_k, D, I);
  // nb_limit == NULL: the reference's own window (keyframe.cpp:274-282) for the rows the last dv_batch_commit gave this
  // rank's b frames - bank row `index` searches rows [0, index - exclude_recent] (all rows incl. itself while
  // index < exclude_recent)
  const int64_t size = bank_size_ref(e);
  const int64_t first = size - (int64_t)e->cfg.world_size * b + (int64_t)e->cfg.rank * b;
  if (first < 0) { set_error("dv_batch_search: no committed batch of this size"); return DV_ERR_INVALID; }
  std::vector<int64_t> lim(b);
  const int64_t ex = e->cfg.exclude_recent;
  for (int i = 0; i < b; ++i) { const int64_t idx = first + i; lim[i] = idx >= ex ? idx - ex + 1 : idx + 1; }
  return (dv_status)bank_search_device(e, b, lim.data(), e->cfg.knn_k, D, I);
}

dv_status dv_batch_read_global(dv_engine* h, int32_t i, float* des512) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!e->mix || !e->mix_done || i < 0 || i >= e->cur_b || !des512) { set_error("dv_batch_read_global: no such frame"); return DV_ERR_INVALID; }
  DV_CUDA_OK(cudaMemcpyAsync(des512, mix_gdesc(e) + (size_t)i * 512, sizeof(float) * 512, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  return DV_OK;
}

dv_status dv_store_lookup(dv_engine* h, int64_t frame_id, int32_t* owner_rank, int32_t* n_total, int32_t* n_sp) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  auto it = s->where.find(frame_id);
  const bool ok = it != s->where.end();
  if (owner_rank) *owner_rank = ok ? it->second.rank : -1;
  if (n_total) *n_total = ok ? it->second.n_sp + it->second.n_vio : 0;
  if (n_sp) *n_sp = ok ? it->second.n_sp : 0;
  return DV_OK;
}

dv_status dv_store_lookup_many(dv_engine* h, int32_t n, const int64_t* frame_ids, int32_t* owner_rank) {
  DV_CHECK_ENGINE(h);
  Store* s = reinterpret_cast<Engine*>(h)->store;
  if (n < 0 || (n > 0 && (!frame_ids || !owner_rank))) { set_error("dv_store_lookup_many: bad arguments"); return DV_ERR_INVALID; }
  for (int i = 0; i < n; ++i) {
    auto it = s->where.find(frame_ids[i]);
    owner_rank[i] = it == s->where.end() ? -1 : it->second.rank;
  }
  return DV_OK;
}

dv_status dv_batch_describe_global(dv_engine* h, int32_t b) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (!e->mix) { set_error("dv_batch_describe_global: engine created without weights"); return DV_ERR_INVALID; }
  e->adopt_upload();
  if (b < 1 || b != e->cur_b) { set_error("dv_batch_describe_global: b must equal the uploaded batch"); return DV_ERR_INVALID; }
  DV_TRY(mix_run(e, b));
  e->mix_done = true;
  // no keyframe metadata travels with this round's all-gather rows
  DV_CUDA_OK(cudaMemsetAsync(s->d_meta, 0xff, sizeof(int) * 4 * b, e->st));
  s->cur_ids.assign(b, -1);
  s->cur_b = b;
  return DV_OK;
}

dv_status dv_store_sync(dv_engine* h) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (e->cfg.world_size > 1 && !e->store->peers_open) { set_error("dv_store_sync: dv_comm_init was not called"); return DV_ERR_COMM; }
  return (dv_status)store_sync_directory(e);
}

dv_status dv_store_read(dv_engine* h, int64_t frame_id, float* kpts_xy, float* desc, int32_t* n_total, int32_t* n_sp) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  auto it = s->where.find(frame_id);
  if (frame_id < 0 || it == s->where.end()) { set_error("dv_store_read: keyframe not resident in any rank's store"); return DV_ERR_INVALID; }
  const Loc L = it->second;
  if (L.rank != e->cfg.rank && !s->peers_open) { set_error("dv_store_read: keyframe lives on another rank and peers are not mapped"); return DV_ERR_INVALID; }
  const int n = L.n_sp + L.n_vio;
  // remote keyframes are read through the CUDA-IPC mapping of the owner's store
  if (kpts_xy) DV_CUDA_OK(cudaMemcpyAsync(kpts_xy, s->peer_kpts[L.rank] + (size_t)L.slot * s->cap * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, e->st));
  if (desc) DV_CUDA_OK(cudaMemcpyAsync(desc, s->peer_desc[L.rank] + (size_t)L.slot * s->cap * 256, sizeof(float) * 256 * n, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  if (n_total) *n_total = n;
  if (n_sp) *n_sp = L.n_sp;
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
  const int me = e->cfg.rank;
  const int sl = s->assign_slot(me, frame_id);
  int inval[4] = {-1, -1, 0, 0};
  int hd[4] = {(int)(frame_id & 0xffffffffll), (int)(frame_id >> 32), n_sp, n_total - n_sp};
  DV_CUDA_OK(cudaMemcpyAsync(s->hdr + sl * 4, inval, sizeof(inval), cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(s->kpts + (size_t)sl * s->cap * 2, kpts_xy, sizeof(float) * 2 * n_total, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(s->desc + (size_t)sl * s->cap * 256, desc, sizeof(float) * 256 * n_total, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(s->hdr + sl * 4, hd, sizeof(hd), cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  s->remember(me, sl, frame_id, n_sp, n_total - n_sp);
  return DV_OK;
}

dv_status dv_batch_match(dv_engine* h, int32_t b, const int64_t* query_ids, const int64_t* old_ids, int32_t* matches,
                         float* mscores, int32_t* k_out) {
  DV_CHECK_ENGINE(h);
  return dv_batch_match_ex(h, b, query_ids, old_ids, DV_PART_WINDOW, DV_PART_ALL, reinterpret_cast<Engine*>(h)->cfg.max_vio,
                           matches, mscores, k_out);
}

dv_status dv_batch_match_begin(dv_engine* h, int32_t b, const int64_t* query_ids, const int64_t* old_ids,
                               int32_t query_part, int32_t old_part, int32_t out_cap) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (!e->lg) { set_error("dv_batch_match: engine created without weights"); return DV_ERR_INVALID; }
  if (b < 1 || b > e->B || !query_ids || !old_ids || query_part < 0 || query_part > 2 ||
      old_part < 0 || old_part > 2 || out_cap < 1) { set_error("dv_batch_match: bad arguments"); return DV_ERR_INVALID; }
  if (s->mp_pending) { set_error("dv_batch_match_begin: the previous match has not been collected (dv_batch_match_end)"); return DV_ERR_INVALID; }
  std::vector<int>& k_out = s->mp_kout;          // per-pair status until dv_batch_match_end
  k_out.assign((size_t)b, 0);
  std::vector<int>& which = s->mp_which;
  std::vector<int>& pull_pair = s->mp_pull_pair;
  which.clear(); pull_pair.clear();
  const int V = out_cap, me = e->cfg.rank;
  auto part = [](const Loc& L, int which, int* first, int* count) {
    *first = which == DV_PART_WINDOW ? L.n_sp : 0;
    *count = which == DV_PART_WINDOW ? L.n_vio : which == DV_PART_SP ? L.n_sp : L.n_sp + L.n_vio;
  };
  std::vector<LgSeg> segs;
  int n_pull = 0;                        // pull_pair: index into `which` of every pair whose old keyframe is remote
  for (int i = 0; i < b; ++i) {
    k_out[i] = 0;
    // per-pair status: -1 = one of the two keyframes is not (or no longer) resident anywhere; the other pairs still run
    auto qi = s->where.find(query_ids[i]);
    auto oi = s->where.find(old_ids[i]);
    if (query_ids[i] < 0 || old_ids[i] < 0 || qi == s->where.end() || oi == s->where.end() ||
        (qi->second.rank != me && !s->peers_open) || (oi->second.rank != me && !s->peers_open)) { k_out[i] = -1; continue; }
    const Loc Q = qi->second, O = oi->second;
    int qf, m, of, n;
    part(Q, query_part, &qf, &m);
    part(O, old_part, &of, &n);
    // keyframe.cpp:373,:935 - the reference skips SP_RE / LightGlue for <= 20 window points; engine floor is 10
    if (m < 10 || n < 10) continue;
    if (n > e->cfg.lg_max_kpts || m > e->cfg.lg_max_kpts) { set_error("dv_batch_match: keyframe exceeds lg_max_kpts"); return DV_ERR_CAPACITY; }
    if (m > out_cap) { set_error("dv_batch_match: out_cap smaller than the query side"); return DV_ERR_CAPACITY; }
    // remote keyframes are read in place from the owner's store (CUDA-IPC mapping, NVLink loads inside k_lg_load)
    const float* qk = s->peer_kpts[Q.rank] + ((size_t)Q.slot * s->cap + qf) * 2;
    const float* qd = s->peer_desc[Q.rank] + ((size_t)Q.slot * s->cap + qf) * 256;
    segs.push_back({qk, qd, m, e->W, e->H, 0});
    const float* ok = s->peer_kpts[O.rank] + ((size_t)O.slot * s->cap + of) * 2;
    const float* od = s->peer_desc[O.rank] + ((size_t)O.slot * s->cap + of) * 256;
    segs.push_back({ok, od, n, e->W, e->H, 0});
    for (int side = 0; side < 2; ++side) {
      const Loc& L = side ? O : Q;
      const int64_t fid = side ? old_ids[i] : query_ids[i];
      if (L.rank == me) continue;
      if (n_pull >= 2 * e->B) break;
      s->h_pull[n_pull] = {s->peer_hdr[L.rank] + L.slot * 4, (int)(fid & 0xffffffffll), (int)(fid >> 32), 0};
      pull_pair.push_back((int)which.size());
      ++n_pull;
    }
    which.push_back(i);
  }
  s->mp_b = b; s->mp_cap = V; s->mp_npull = 0;
  if (which.empty()) { s->mp_pending = true; e->match_pending = true; return DV_OK; }
  std::function<int()> after_load = [&]() -> int {
    if (!n_pull) return DV_OK;
    DV_CUDA_OK(cudaMemcpyAsync(s->d_pull, s->h_pull, sizeof(Store::PullJob) * n_pull, cudaMemcpyHostToDevice, e->st));
    k_pull_check<<<cdiv(n_pull, 128), 128, 0, e->st>>>(s->d_pull, n_pull, s->d_pull_bad);
    DV_CUDA_OK(cudaGetLastError());
    DV_CUDA_OK(cudaMemcpyAsync(s->h_pull_bad, s->d_pull_bad, sizeof(int) * n_pull, cudaMemcpyDeviceToHost, e->st));
    DV_LAUNCHED(e, 1);
    return DV_OK;
  };
  // DV_LG_CHUNK pairs per LightGlue pass (0 = all): smaller chunks keep the token buffers of a pass inside the L2
  static const int lg_chunk = [] { const char* env = getenv("DV_LG_CHUNK"); return env ? atoi(env) : DV_LG_CHUNK_DEFAULT; }();
  const int Pn = (int)which.size();
  const int step = (lg_chunk > 0 && lg_chunk < Pn) ? lg_chunk : Pn;
  for (int p0 = 0; p0 < Pn; p0 += step)
    DV_TRY(lg_run(e, std::min(step, Pn - p0), segs.data() + 2 * p0, (n_pull && p0 + step >= Pn) ? &after_load : nullptr, p0));   // seqlock check after the LAST load
  s->mp_npull = n_pull;
  DV_TRY(lg_fetch_batch_begin(e, (int)which.size(), V));               // result copies queued behind the kernels
  s->mp_pending = true;
  e->match_pending = true;
  return DV_OK;
}

dv_status dv_batch_match_end(dv_engine* h, int32_t* matches, float* mscores, int32_t* k_out) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Store* s = e->store;
  if (!s->mp_pending) { set_error("dv_batch_match_end: no match in flight"); return DV_ERR_INVALID; }
  if (!matches || !mscores || !k_out) { set_error("dv_batch_match_end: null output"); return DV_ERR_INVALID; }
  s->mp_pending = false;
  e->match_pending = false;
  for (int i = 0; i < s->mp_b; ++i) k_out[i] = s->mp_kout[i];
  if (s->mp_which.empty()) return DV_OK;
  DV_TRY(lg_fetch_batch_end(e, (int)s->mp_which.size(), s->mp_cap, s->mp_which.data(), matches, mscores, k_out));
  // a remote keyframe whose slot was being rewritten while it was read is reported like a non-resident one
  for (int j = 0; j < s->mp_npull; ++j)
    if (s->h_pull_bad[j]) k_out[s->mp_which[s->mp_pull_pair[j]]] = -1;
  return DV_OK;
}

dv_status dv_batch_match_ex(dv_engine* h, int32_t b, const int64_t* query_ids, const int64_t* old_ids, int32_t query_part,
                            int32_t old_part, int32_t out_cap, int32_t* matches, float* mscores, int32_t* k_out) {
  if (!matches || !mscores || !k_out) { set_error("dv_batch_match: bad arguments"); return DV_ERR_INVALID; }
  const dv_status rc = dv_batch_match_begin(h, b, query_ids, old_ids, query_part, old_part, out_cap);
  if (rc != DV_OK) return rc;
  return dv_batch_match_end(h, matches, mscores, k_out);
}

}  // extern "C"
