// Exact cosine / inner-product kNN over the keyframe global-descriptor bank (replaces faiss IndexFlatIP,
// reference keyframe.cpp:262-346).  HBM-bound: the bank [rows,512] f32 is streamed once per query batch with
// 128-bit coalesced loads; one warp per bank row, the queries live in registers, warp-shuffle dot products,
// per-block partial top-k in shared memory, then a single-block merge.  Tie rule: lowest index first.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "engine.h"

namespace dv {

#define KNN_MAXK 8
#define KNN_QB 8             // queries processed per pass over the bank (query vectors live in shared memory)
#define KNN_WARPS 8
#define KNN_RPW 4            // bank rows in flight per warp (16 x 128-bit loads per lane)
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
  unsigned int* d_ticket = nullptr;   // [ceil(B / KNN_QB)] "blocks finished" counters of the fused merge (self-resetting)
  int nblocks_cap = 0;
  bool fused = true;           // DV_KNN_FUSED=0: two launches + limit upload + result copies (A/B)
  float* h_D = nullptr; long long* h_I = nullptr; long long* h_nb = nullptr; float* h_q = nullptr;   // pinned
};

// candidate ordering: larger D first, then smaller index
__device__ __forceinline__ bool better(float d1, long long i1, float d2, long long i2) {
  return d1 > d2 || (d1 == d2 && i1 < i2);
}

template <int K>
__device__ __forceinline__ void topk_insert(float (&D)[K], long long (&I)[K], float d, long long i) {
  if (!better(d, i, D[K - 1], I[K - 1])) return;
  D[K - 1] = d; I[K - 1] = i;
#pragma unroll
  for (int p = K - 1; p > 0; --p) {
    if (better(D[p], I[p], D[p - 1], I[p - 1])) {
      const float td = D[p]; D[p] = D[p - 1]; D[p - 1] = td;
      const long long ti = I[p]; I[p] = I[p - 1]; I[p - 1] = ti;
    }
  }
}

// Streaming scan: grid (nblocks, ceil(nq / KNN_QB)).  Each warp keeps KNN_RPW bank rows (4 x 2 KB) in flight in
// registers; the 8 query vectors sit in shared memory.  The 32 partial dot products (4 rows x 8 queries) of a lane
// are reduced across the warp with a 31-shuffle butterfly that leaves lane L holding the total of pair L, so lane L
// owns the candidate list of query (L & 7) for row slot (L >> 3).
// Search windows by value (kernel arguments) for up to 64 queries: no host->device copy before the launch.
struct KnnLimits { long long v[64]; };

// FUSED == true: the block that finishes last for its query group (ticket counter) merges the per-block candidate
// lists and writes the final top-k - to device memory or straight into mapped pinned host memory - so a search is ONE
// launch with no limit upload and no result copy (r01: 41 us per 10k-row search, 37 us of it fixed cost).
template <int K, bool FUSED>
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_scan(const float* __restrict__ bank,
                                                            const float* __restrict__ q,
                                                            const long long* __restrict__ nb_limit_dev,
                                                            const KnnLimits lim_args, int nq,
                                                            float* __restrict__ partD, long long* __restrict__ partI,
                                                            int nblocks, unsigned int* __restrict__ ticket,
                                                            float* __restrict__ outD, long long* __restrict__ outI) {
  const long long* __restrict__ nb_limit = nb_limit_dev ? nb_limit_dev : lim_args.v;
  __shared__ __align__(16) float sq[KNN_QB][512];
  __shared__ float sD[KNN_QB][KNN_WARPS * 4][K];
  __shared__ long long sI[KNN_QB][KNN_WARPS * 4][K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * KNN_QB;
  const int nqb = min(KNN_QB, nq - q0);
  for (int i = threadIdx.x; i < KNN_QB * 128; i += blockDim.x) {
    const int j = i >> 7, c = i & 127;
    reinterpret_cast<float4*>(&sq[j][0])[c] =
        (j < nqb) ? __ldg(reinterpret_cast<const float4*>(q + (int64_t)(q0 + j) * 512) + c) : make_float4(0, 0, 0, 0);
  }
  long long maxlim = 0;
  for (int j = 0; j < nqb; ++j) maxlim = max(maxlim, nb_limit[q0 + j]);
  const int myq = lane & 7, myslot = lane >> 3;
  const long long mylim = (myq < nqb) ? nb_limit[q0 + myq] : 0;
  float D[K]; long long I[K];
#pragma unroll
  for (int t = 0; t < K; ++t) { D[t] = -INFINITY; I[t] = -1; }
  __syncthreads();
  // grid-stride over 4-row groups: warp w of block b takes groups (b * 8 + w) + i * (gridDim.x * 8), so neighbouring warps
  // stream neighbouring 8 KB and the per-block prologue (query staging) / epilogue (candidate merge) is paid once per
  // block, not once per 64 rows (ncu: 40 % of the samples sat in the merge tail, the scan ran at 2.5 TB/s)
  const long long r1 = maxlim;
  const long long rstep = (long long)gridDim.x * KNN_WARPS * KNN_RPW;
  for (long long r = ((long long)blockIdx.x * KNN_WARPS + warp) * KNN_RPW; r < r1; r += rstep) {
    float4 v[KNN_RPW][4];
#pragma unroll
    for (int rr = 0; rr < KNN_RPW; ++rr) {
      const float4* row = reinterpret_cast<const float4*>(bank + (r + rr) * 512);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[rr][c] = (r + rr < r1) ? __ldg(row + c * 32 + lane) : make_float4(0, 0, 0, 0);
    }
    float s[32];   // index = rr * 8 + j
#pragma unroll
    for (int j = 0; j < KNN_QB; ++j) {
      if (j >= nqb) {                                 // block-uniform: a single-query search skips 7/8 of the FMAs
#pragma unroll
        for (int rr = 0; rr < KNN_RPW; ++rr) s[rr * 8 + j] = 0.f;
        continue;
      }
      float4 qv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) qv[c] = reinterpret_cast<const float4*>(&sq[j][0])[c * 32 + lane];
#pragma unroll
      for (int rr = 0; rr < KNN_RPW; ++rr) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          a += v[rr][c].x * qv[c].x + v[rr][c].y * qv[c].y + v[rr][c].z * qv[c].z + v[rr][c].w * qv[c].w;
        s[rr * 8 + j] = a;
      }
    }
    // butterfly transpose-reduce: after the 5 steps lane L holds sum over lanes of s[L]
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
      const bool up = (lane & h) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const float keep = up ? s[i + h] : s[i];
        const float send = up ? s[i] : s[i + h];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
      }
    }
    const long long row = r + myslot;
    if (row < mylim) topk_insert<K>(D, I, s[0], row);
  }
#pragma unroll
  for (int t = 0; t < K; ++t) { sD[myq][warp * 4 + myslot][t] = D[t]; sI[myq][warp * 4 + myslot][t] = I[t]; }
  __syncthreads();
  // per-block merge: warp j merges the 32 candidate lists (8 warps x 4 row slots) of query j - one list per lane, then a
  // shuffle tree (the serial 8-thread version cost more than the scan itself on 10k-row banks)
  if (warp < nqb) {
    const int j = warp;
    float bD[K]; long long bI[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { bD[t] = sD[j][lane][t]; bI[t] = sI[j][lane][t]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float oD[K]; long long oI[K];
#pragma unroll
      for (int t = 0; t < K; ++t) { oD[t] = __shfl_xor_sync(0xffffffffu, bD[t], o); oI[t] = __shfl_xor_sync(0xffffffffu, bI[t], o); }
#pragma unroll
      for (int t = 0; t < K; ++t)
        if (oI[t] >= 0) topk_insert<K>(bD, bI, oD[t], oI[t]);
    }
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < K; ++t) {
        partD[((int64_t)(q0 + j) * nblocks + blockIdx.x) * K + t] = bD[t];
        partI[((int64_t)(q0 + j) * nblocks + blockIdx.x) * K + t] = bI[t];
      }
    }
  }
  if (!FUSED) return;
  // ---- fused merge: last block of this query group
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ticket[blockIdx.y], 1u) == (unsigned)gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (warp < nqb) {                                  // one warp per query of the group
    const int j = warp;
    float mD[K]; long long mI[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { mD[t] = -INFINITY; mI[t] = -1; }
    for (int b = lane; b < nblocks; b += 32)
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const long long i = __ldcg(&partI[((int64_t)(q0 + j) * nblocks + b) * K + t]);
        if (i >= 0) topk_insert<K>(mD, mI, __ldcg(&partD[((int64_t)(q0 + j) * nblocks + b) * K + t]), i);
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float oD[K]; long long oI[K];
#pragma unroll
      for (int t = 0; t < K; ++t) { oD[t] = __shfl_xor_sync(0xffffffffu, mD[t], o); oI[t] = __shfl_xor_sync(0xffffffffu, mI[t], o); }
#pragma unroll
      for (int t = 0; t < K; ++t)
        if (oI[t] >= 0) topk_insert<K>(mD, mI, oD[t], oI[t]);
    }
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < K; ++t) { outD[(int64_t)(q0 + j) * K + t] = mD[t]; outI[(int64_t)(q0 + j) * K + t] = mI[t]; }
      __threadfence_system();                        // the outputs may live in mapped host memory
    }
  }
  if (threadIdx.x == 0) ticket[blockIdx.y] = 0u;     // ready for the next launch
}

// one block per query: merge nblocks partial lists (strided gather, then a shared-memory tree)
template <int K>
__global__ void __launch_bounds__(256) k_knn_merge(const float* __restrict__ partD, const long long* __restrict__ partI,
                                                   int nblocks, float* __restrict__ outD,
                                                   long long* __restrict__ outI) {
  __shared__ float sD[256][K];
  __shared__ long long sI[256][K];
  const int qi = blockIdx.x, tid = threadIdx.x;
  float D[K]; long long I[K];
#pragma unroll
  for (int t = 0; t < K; ++t) { D[t] = -INFINITY; I[t] = -1; }
  for (int b = tid; b < nblocks; b += 256)
#pragma unroll
    for (int t = 0; t < K; ++t) {
      const long long i = partI[((int64_t)qi * nblocks + b) * K + t];
      if (i >= 0) topk_insert<K>(D, I, partD[((int64_t)qi * nblocks + b) * K + t], i);
    }
  for (int stride = 128; stride >= 1; stride >>= 1) {
#pragma unroll
    for (int t = 0; t < K; ++t) { sD[tid][t] = D[t]; sI[tid][t] = I[t]; }
    __syncthreads();
    if (tid < stride) {
#pragma unroll
      for (int t = 0; t < K; ++t)
        if (sI[tid + stride][t] >= 0) topk_insert<K>(D, I, sD[tid + stride][t], sI[tid + stride][t]);
    }
    __syncthreads();
  }
  if (tid == 0)
#pragma unroll
    for (int t = 0; t < K; ++t) { outD[(int64_t)qi * K + t] = D[t]; outI[(int64_t)qi * K + t] = I[t]; }
}

template <int K>
static void knn_launch(Engine* e, const float* rows, const float* q, const long long* d_nb, const KnnLimits& lim, int nq,
                       int nblocks, float* partD, long long* partI, unsigned int* ticket, float* outD, long long* outI,
                       bool fused, cudaStream_t st) {
  ProbeScope pr(e, 1);   // roofline probe of the HBM-bound scan (bench.py --config mix_knn_10k)
  if (fused) {
    k_knn_scan<K, true><<<dim3(nblocks, cdiv(nq, KNN_QB)), KNN_WARPS * 32, 0, st>>>(rows, q, d_nb, lim, nq, partD, partI,
                                                                                 nblocks, ticket, outD, outI);
  } else {
    k_knn_scan<K, false><<<dim3(nblocks, cdiv(nq, KNN_QB)), KNN_WARPS * 32, 0, st>>>(rows, q, d_nb, lim, nq, partD, partI,
                                                                                  nblocks, ticket, outD, outI);
    k_knn_merge<K><<<nq, 256, 0, st>>>(partD, partI, nblocks, outD, outI);
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
  DV_TRY(e->alloc(&b->d_ticket, (size_t)cdiv(B, KNN_QB) + 1));
  { const char* env = getenv("DV_KNN_FUSED"); b->fused = !(env && env[0] == '0'); }
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

// queries at `d_q` (device memory, or mapped pinned host memory); nb_limit host array [nq]
static int bank_search_impl(Engine* e, const float* d_q, int nq, const int64_t* nb_limit, int k, float* D_host, int64_t* I_host) {
  Bank* b = e->bank;
  if (nq <= 0) return DV_OK;
  if (k < 1 || k > KNN_MAXK) { set_error("knn: k out of range"); return DV_ERR_INVALID; }
  KnnLimits lim;
  long long maxlim = 0;
  for (int i = 0; i < nq; ++i) {
    long long l = std::max<long long>(0, std::min<long long>(nb_limit[i], b->size));
    b->h_nb[i] = l;
    if (i < 64) lim.v[i] = l;
    maxlim = std::max(maxlim, l);
  }
  StageScope sc(e, ST_KNN);
  if (maxlim == 0) {
    for (int i = 0; i < nq * k; ++i) { D_host[i] = -INFINITY; I_host[i] = -1; }
    return DV_OK;
  }
  // fused path: limits by value, merge in the scan's last block, results written straight into pinned host memory
  const bool fused = b->fused && nq <= 64;
  const long long* d_nb = nullptr;
  if (!fused) {
    DV_CUDA_OK(cudaMemcpyAsync(b->d_nb, b->h_nb, sizeof(long long) * nq, cudaMemcpyHostToDevice, e->st));
    d_nb = b->d_nb;
  }
  float* oD = fused ? b->h_D : b->outD;
  long long* oI = fused ? b->h_I : b->outI;
  static int dev_sms = 0;
  if (!dev_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev); }
  // 128 registers x 256 threads: two resident blocks per SM.  ONE wave over all query groups: a batch of 64 queries is
  // 8 query groups, each scanning the bank with (2 * SMs) / 8 fat blocks instead of 8 x (rows / 64) thin ones whose
  // prologue (query staging) and merge dominated on 10k-row banks (r02: 100 us per 64-query search at 12k rows).
  const int groups = cdiv(nq, KNN_QB);
  const int wave = std::max(1, (dev_sms * 2) / groups);
  const int nblocks = (int)std::min<int64_t>(cdiv64(maxlim, KNN_ROWS_PER_BLOCK), (int64_t)wave);
  switch (k) {
#define DV_KNN_CASE(KK) case KK: knn_launch<KK>(e, b->rows, d_q, d_nb, lim, nq, nblocks, b->partD, b->partI, b->d_ticket, oD, oI, fused, e->st); break;
    DV_KNN_CASE(1) DV_KNN_CASE(2) DV_KNN_CASE(3) DV_KNN_CASE(4) DV_KNN_CASE(5) DV_KNN_CASE(6) DV_KNN_CASE(7)
    default: knn_launch<8>(e, b->rows, d_q, d_nb, lim, nq, nblocks, b->partD, b->partI, b->d_ticket, oD, oI, fused, e->st); break;
#undef DV_KNN_CASE
  }
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, fused ? 1 : 2);
  if (!fused) {
    DV_CUDA_OK(cudaMemcpyAsync(b->h_D, b->outD, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(b->h_I, b->outI, sizeof(long long) * nq * k, cudaMemcpyDeviceToHost, e->st));
  }
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  for (int i = 0; i < nq * k; ++i) { D_host[i] = b->h_D[i]; I_host[i] = (int64_t)b->h_I[i]; }
  return DV_OK;
}

int bank_search_device(Engine* e, int nq, const int64_t* nb_limit, int k, float* D_host, int64_t* I_host) {
  return bank_search_impl(e, e->bank->q, nq, nb_limit, k, D_host, I_host);
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
  if (e->bank->fused) {
    // the single query is read by the kernel straight from mapped pinned host memory (2 KB over PCIe): no H2D copy
    memcpy(e->bank->h_q, q512, 512 * sizeof(float));
    return (dv_status)bank_search_impl(e, e->bank->h_q, 1, &nb_limit, k, D, I);
  }
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
