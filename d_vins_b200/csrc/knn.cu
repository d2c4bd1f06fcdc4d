// Exact cosine / inner-product kNN over the keyframe global-descriptor bank (replaces faiss IndexFlatIP,
// reference keyframe.cpp:262-346).  HBM-bound: the bank [rows,512] f32 is streamed once per query batch with
// 128-bit coalesced loads; one warp per bank row, the queries live in registers, warp-shuffle dot products,
// per-block partial top-k in shared memory, then a single-block merge.  Tie rule: lowest index first.
#include <float.h>

#include <algorithm>

#include "engine.h"

namespace dv {

#define KNN_MAXK 8
#define KNN_QB 4             // queries processed per pass over the bank
#define KNN_WARPS 8
#define KNN_ROWS_PER_BLOCK 64

struct Bank {
  float* rows = nullptr;       // [capacity, 512]
  int64_t capacity = 0;
  int64_t size = 0;
  float* q = nullptr;          // [B, 512] query staging
  float* partD = nullptr;      // [B, nblocks_cap, K]
  long long* partI = nullptr;
  float* outD = nullptr;       // [B, K]
  long long* outI = nullptr;
  long long* d_nb = nullptr;   // [B]
  int nblocks_cap = 0;
  float* h_D = nullptr; long long* h_I = nullptr; long long* h_nb = nullptr; float* h_q = nullptr;   // pinned
};

// candidate ordering: larger D first, then smaller index
__device__ __forceinline__ bool better(float d1, long long i1, float d2, long long i2) {
  return d1 > d2 || (d1 == d2 && i1 < i2);
}

__device__ __forceinline__ void topk_insert(float* D, long long* I, int k, float d, long long i) {
  if (!better(d, i, D[k - 1], I[k - 1])) return;
  int p = k - 1;
  while (p > 0 && better(d, i, D[p - 1], I[p - 1])) { D[p] = D[p - 1]; I[p] = I[p - 1]; --p; }
  D[p] = d; I[p] = i;
}

// grid: (nblocks, ceil(nq / KNN_QB)).  Each block scans KNN_ROWS_PER_BLOCK rows for up to KNN_QB queries.
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_scan(const float* __restrict__ bank,
                                                            const float* __restrict__ q,
                                                            const long long* __restrict__ nb_limit, int nq, int k,
                                                            float* __restrict__ partD, long long* __restrict__ partI,
                                                            int nblocks) {
  __shared__ float sD[KNN_QB][KNN_WARPS][KNN_MAXK];
  __shared__ long long sI[KNN_QB][KNN_WARPS][KNN_MAXK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * KNN_QB;
  const int nqb = min(KNN_QB, nq - q0);
  // each lane owns 16 of the 512 dims: 4 float4 at lane*4 + {0,128,256,384}
  float4 qr[KNN_QB][4];
  long long lim[KNN_QB];
  long long maxlim = 0;
#pragma unroll
  for (int j = 0; j < KNN_QB; ++j) {
    lim[j] = (j < nqb) ? nb_limit[q0 + j] : 0;
    maxlim = max(maxlim, lim[j]);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      qr[j][c] = (j < nqb) ? __ldg(reinterpret_cast<const float4*>(q + (int64_t)(q0 + j) * 512) + c * 32 + lane)
                           : make_float4(0, 0, 0, 0);
  }
  float D[KNN_QB][KNN_MAXK];
  long long I[KNN_QB][KNN_MAXK];
#pragma unroll
  for (int j = 0; j < KNN_QB; ++j)
#pragma unroll
    for (int t = 0; t < KNN_MAXK; ++t) { D[j][t] = -INFINITY; I[j][t] = -1; }
  const long long r0 = (long long)blockIdx.x * KNN_ROWS_PER_BLOCK;
  const long long r1 = min(r0 + KNN_ROWS_PER_BLOCK, maxlim);
  for (long long r = r0 + warp; r < r1; r += KNN_WARPS) {
    const float4* row = reinterpret_cast<const float4*>(bank + r * 512);
    float4 v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = __ldg(row + c * 32 + lane);
#pragma unroll
    for (int j = 0; j < KNN_QB; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        s += v[c].x * qr[j][c].x + v[c].y * qr[j][c].y + v[c].z * qr[j][c].z + v[c].w * qr[j][c].w;
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (j < nqb && r < lim[j]) topk_insert(D[j], I[j], k, s, r);   // every lane keeps the same (uniform) list
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < KNN_QB; ++j)
      for (int t = 0; t < k; ++t) { sD[j][warp][t] = D[j][t]; sI[j][warp][t] = I[j][t]; }
  }
  __syncthreads();
  if (threadIdx.x < nqb) {
    const int j = threadIdx.x;
    float bD[KNN_MAXK]; long long bI[KNN_MAXK];
    for (int t = 0; t < k; ++t) { bD[t] = -INFINITY; bI[t] = -1; }
    for (int w = 0; w < KNN_WARPS; ++w)
      for (int t = 0; t < k; ++t)
        if (sI[j][w][t] >= 0) topk_insert(bD, bI, k, sD[j][w][t], sI[j][w][t]);
    for (int t = 0; t < k; ++t) {
      partD[((int64_t)(q0 + j) * nblocks + blockIdx.x) * k + t] = bD[t];
      partI[((int64_t)(q0 + j) * nblocks + blockIdx.x) * k + t] = bI[t];
    }
  }
}

// one block per query: merge nblocks partial lists
__global__ void __launch_bounds__(256) k_knn_merge(const float* __restrict__ partD, const long long* __restrict__ partI,
                                                   int nblocks, int k, float* __restrict__ outD,
                                                   long long* __restrict__ outI) {
  __shared__ float sD[256][KNN_MAXK];
  __shared__ long long sI[256][KNN_MAXK];
  const int qi = blockIdx.x, tid = threadIdx.x;
  float D[KNN_MAXK]; long long I[KNN_MAXK];
  for (int t = 0; t < k; ++t) { D[t] = -INFINITY; I[t] = -1; }
  for (int b = tid; b < nblocks; b += 256)
    for (int t = 0; t < k; ++t) {
      const long long i = partI[((int64_t)qi * nblocks + b) * k + t];
      if (i >= 0) topk_insert(D, I, k, partD[((int64_t)qi * nblocks + b) * k + t], i);
    }
  for (int t = 0; t < k; ++t) { sD[tid][t] = D[t]; sI[tid][t] = I[t]; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 256; ++w)
      for (int t = 0; t < k; ++t)
        if (sI[w][t] >= 0) topk_insert(D, I, k, sD[w][t], sI[w][t]);
    for (int t = 0; t < k; ++t) { outD[(int64_t)qi * k + t] = D[t]; outI[(int64_t)qi * k + t] = I[t]; }
  }
}

// ------------------------------------------------------------------------------------------------ host
int bank_init(Engine* e) {
  Bank* b = new Bank();
  e->bank = b;
  b->capacity = e->cfg.bank_capacity;
  const int B = std::max(e->B * e->cfg.world_size, 1), K = e->cfg.knn_k;
  DV_TRY(e->alloc(&b->rows, (size_t)b->capacity * 512));
  b->nblocks_cap = (int)cdiv64(b->capacity, KNN_ROWS_PER_BLOCK);
  DV_TRY(e->alloc(&b->q, (size_t)B * 512));
  DV_TRY(e->alloc(&b->partD, (size_t)B * b->nblocks_cap * K));
  DV_TRY(e->alloc(&b->partI, (size_t)B * b->nblocks_cap * K));
  DV_TRY(e->alloc(&b->outD, (size_t)B * K));
  DV_TRY(e->alloc(&b->outI, (size_t)B * K));
  DV_TRY(e->alloc(&b->d_nb, (size_t)B));
  DV_TRY(e->alloc_pinned(&b->h_D, (size_t)B * K));
  DV_TRY(e->alloc_pinned(&b->h_I, (size_t)B * K));
  DV_TRY(e->alloc_pinned(&b->h_nb, (size_t)B));
  DV_TRY(e->alloc_pinned(&b->h_q, (size_t)B * 512));
  return DV_OK;
}
void bank_free(Engine* e) { delete e->bank; e->bank = nullptr; }

float* bank_rows(Engine* e) { return e->bank->rows; }
float* bank_query_buf(Engine* e) { return e->bank->q; }
int64_t& bank_size_ref(Engine* e) { return e->bank->size; }

// queries already in bank->q (device); nb_limit host array [nq]
int bank_search_device(Engine* e, int nq, const int64_t* nb_limit, int k, float* D_host, int64_t* I_host) {
  Bank* b = e->bank;
  if (nq <= 0) return DV_OK;
  if (k < 1 || k > KNN_MAXK) { set_error("knn: k out of range"); return DV_ERR_INVALID; }
  long long maxlim = 0;
  for (int i = 0; i < nq; ++i) {
    long long l = std::max<long long>(0, std::min<long long>(nb_limit[i], b->size));
    b->h_nb[i] = l;
    maxlim = std::max(maxlim, l);
  }
  StageScope sc(e, ST_KNN);
  if (maxlim == 0) {
    for (int i = 0; i < nq * k; ++i) { D_host[i] = -INFINITY; I_host[i] = -1; }
    return DV_OK;
  }
  DV_CUDA_OK(cudaMemcpyAsync(b->d_nb, b->h_nb, sizeof(long long) * nq, cudaMemcpyHostToDevice, e->st));
  const int nblocks = (int)cdiv64(maxlim, KNN_ROWS_PER_BLOCK);
  k_knn_scan<<<dim3(nblocks, cdiv(nq, KNN_QB)), KNN_WARPS * 32, 0, e->st>>>(b->rows, b->q, b->d_nb, nq, k, b->partD,
                                                                          b->partI, nblocks);
  k_knn_merge<<<nq, 256, 0, e->st>>>(b->partD, b->partI, nblocks, k, b->outD, b->outI);
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, 2);
  DV_CUDA_OK(cudaMemcpyAsync(b->h_D, b->outD, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(b->h_I, b->outI, sizeof(long long) * nq * k, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  for (int i = 0; i < nq * k; ++i) { D_host[i] = b->h_D[i]; I_host[i] = (int64_t)b->h_I[i]; }
  return DV_OK;
}

}  // namespace dv

using namespace dv;
#define DV_CHECK_ENGINE(e) do { if (!(e)) { dv::set_error("null engine"); return DV_ERR_INVALID; } } while (0)

extern "C" {

dv_status dv_bank_append(dv_engine* h, const float* des512, int64_t* row) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Bank* b = e->bank;
  if (!des512) { set_error("dv_bank_append: null descriptor"); return DV_ERR_INVALID; }
  if (b->size >= b->capacity) { set_error("dv_bank_append: bank full"); return DV_ERR_CAPACITY; }
  DV_CUDA_OK(cudaMemcpyAsync(b->rows + b->size * 512, des512, 512 * sizeof(float), cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  if (row) *row = b->size;
  b->size++;
  return DV_OK;
}

dv_status dv_bank_size(dv_engine* h, int64_t* rows) {
  DV_CHECK_ENGINE(h);
  if (rows) *rows = reinterpret_cast<Engine*>(h)->bank->size;
  return DV_OK;
}

dv_status dv_bank_search(dv_engine* h, const float* q512, int64_t nb_limit, int32_t k, float* D, int64_t* I) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!q512 || !D || !I || k < 1 || k > KNN_MAXK) { set_error("dv_bank_search: bad argument"); return DV_ERR_INVALID; }
  DV_CUDA_OK(cudaMemcpyAsync(e->bank->q, q512, 512 * sizeof(float), cudaMemcpyHostToDevice, e->st));
  return (dv_status)bank_search_device(e, 1, &nb_limit, k, D, I);
}

dv_status dv_bank_export(dv_engine* h, float* dst, int64_t max_rows, int64_t* rows) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Bank* b = e->bank;
  if (!dst || max_rows < b->size) { set_error("dv_bank_export: buffer too small"); return DV_ERR_CAPACITY; }
  DV_CUDA_OK(cudaMemcpyAsync(dst, b->rows, (size_t)b->size * 512 * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  if (rows) *rows = b->size;
  return DV_OK;
}

dv_status dv_bank_import(dv_engine* h, const float* src, int64_t rows) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  Bank* b = e->bank;
  if (rows < 0 || (rows > 0 && !src)) { set_error("dv_bank_import: bad argument"); return DV_ERR_INVALID; }
  if (rows > b->capacity) { set_error("dv_bank_import: exceeds bank capacity"); return DV_ERR_CAPACITY; }
  DV_CUDA_OK(cudaMemcpyAsync(b->rows, src, (size_t)rows * 512 * sizeof(float), cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  b->size = rows;
  return DV_OK;
}

}  // extern "C"
