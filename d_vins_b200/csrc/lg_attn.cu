// LightGlue attention on tcgen05 (self- and cross-attention share it): one CTA per (job, head, 128-query tile).
//
//   S[128 x 128] = Q K^T        Q, K tiles: 64-wide (head_dim) K-major SWIZZLE_128B rows brought by TMA straight from the
//                               packed qkv buffer [T,768]; accumulator in TMEM columns [0,128)
//   softmax (online)            4 warps, thread == query row: two passes of tcgen05.ld over the S row (max, then exp2 /
//                               sum); P is written as fp16 into shared memory in the K-major SWIZZLE_128B layout of an
//                               A operand (two 128x64 tiles)
//   O_chunk[128 x 64] = P V     V tile rows are keys (the MMA's K) with head_dim contiguous: an MN-major B operand, i.e.
//                               the tile exactly as TMA delivers it - no transpose; accumulator in TMEM columns [128,192)
//   O = O * corr + O_chunk      in registers of the row-owning thread (no TMEM read-modify-write)
//
// Replaces the mma.sync flash kernel of lg.cu (kept behind DV_LG_ATTN=mma): r01 launch lists had it at 77 us per launch
// (~130 TFLOP/s, legacy HMMA pipe) = 34 % of the LightGlue stage.  Here the tensor work is 12 tcgen05.mma per 128-key
// chunk and the bound becomes the 128x128 exp2 per chunk on the MUFU pipe.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "gemm.h"
#include "lg.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

namespace {
constexpr int TILE_BYTES = 128 * 128;          // 128 rows x 64 halves
constexpr int OFF_Q = 0;
constexpr int OFF_K = TILE_BYTES;              // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;  // 2 stages
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;  // two 128x64 tiles
constexpr int OFF_BAR = OFF_P + 2 * TILE_BYTES;
// 115 328 B: TWO CTAs per SM (2 x (115 328 + 1 KB reserved) <= 228 KB) - each CTA is one serial QK -> softmax -> PV chain,
// the second CTA's MMAs and TMEM loads fill the first one's softmax latency.  The dynamic-smem base sits right after the
// 1 KB reserved region, i.e. 1024-aligned; 512 B of slack cover a smaller alignment, anything worse traps.
constexpr int ALIGN_SLACK = 512;
constexpr int SMEM_BYTES = OFF_BAR + 128 + ALIGN_SLACK;
}  // namespace

__device__ __forceinline__ float ex2_approx(float x) {     // one MUFU op (exp2f adds range fix-ups around it)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// LAZY = true: O accumulates in TMEM across all key chunks (the PV MMAs keep their accumulate flag) and is only rescaled -
// read, multiplied, written back with tcgen05.st - when some row's running maximum grows by more than 2^8 over the value
// the probabilities are scaled with (P <= 256 stays exact enough in fp16).  The softmax threads then never wait for a PV
// product nor read O per chunk: the wait for PV(j-1) moves behind pass A of chunk j.
template <bool LAZY>
__global__ void __launch_bounds__(192, 2) lg_attn_umma_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                              const AttnJobU* __restrict__ jobs,
                                                              __half* __restrict__ ctx, int ldo, float sl2, long long* dbg) {
  const AttnJobU jb = jobs[blockIdx.z];
  const int q0 = blockIdx.x * 128;
  if (q0 >= jb.nq) return;                     // uniform per CTA
  const int head = blockIdx.y;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if ((int)(smem - smem_raw) > ALIGN_SLACK) __trap();
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = q_full + 1;              // [2]
  uint64_t* kv_empty = kv_full + 2;            // [2]
  uint64_t* s_full = kv_empty + 2;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = (jb.nk + 127) >> 7;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQKV, q_full, jb.q_col + head * 64, jb.q_row + q0);
      for (int j = 0; j < n_chunks; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(smem + OFF_K + s * TILE_BYTES, &tmQKV, &kv_full[s], jb.k_col + head * 64, jb.k_row + j * 128);
        tma_load_2d(smem + OFF_V + s * TILE_BYTES, &tmQKV, &kv_full[s], jb.v_col + head * 64, jb.k_row + j * 128);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc_f16_f32(128, 128);                 // A, B K-major
      constexpr uint32_t idesc_pv = make_idesc_f16_f32(128, 64) | (1u << 16);    // B (= V) MN-major
      const uint64_t dq = make_desc_sw128(smem_u32(smem + OFF_Q));
      const uint64_t dp = make_desc_sw128(smem_u32(smem + OFF_P));
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_chunks; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = make_desc_sw128(smem_u32(smem + OFF_K + s * TILE_BYTES));
        const uint64_t dv = make_desc_sw128(smem_u32(smem + OFF_V + s * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)                                             // head_dim 64 = 4 k-steps
          tc_mma_f16(tmem_base, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
        tc_commit(s_full);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const int pv_steps = min(8, (jb.nk - j * 128 + 15) >> 4);                // only the k-steps that hold valid keys
        for (int k = 0; k < pv_steps; ++k)                                      // up to 128 keys = 8 k-steps of 16
          // A: P tile (k >> 2), 32-byte step inside its 128-byte rows.  B: V rows [16k, 16k+16) = +2048 bytes.
          tc_mma_f16(tmem_base + 128, dp + (uint64_t)((k >> 2) * (TILE_BYTES >> 4) + (k & 3) * 2),
                     dv + (uint64_t)(k * 128), idesc_pv, (uint32_t)(LAZY ? (j | k) != 0 : k != 0));
        tc_commit(o_full);
        tc_commit(&kv_empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction: thread == query row
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = smem + OFF_P + row * 128;
    const bool warp_live = q0 + q * 32 < jb.nq;
    const bool dbgt = dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 1 && threadIdx.x == 64;
    for (int j = 0; j < n_chunks; ++j) {
      const long long t0 = dbgt ? clock64() : 0;
      mbar_wait(s_full, j & 1);
      const long long t1 = dbgt ? clock64() : 0;
      tc_fence_after();
      const int kbase = j * 128;
      const int ngrp = min(4, (jb.nk - kbase + 31) >> 5);    // 32-key groups holding valid keys (PV reads no further)
      if (!warp_live) {                                      // every row of this warp is >= nq: nothing to compute
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cnt(p_ready);
        if (!LAZY) mbar_wait(o_full, j & 1);
        continue;
      }
      // one sweep over this warp's valid 32-key groups of S: P = exp2(s * sl2 - mb) to shared memory (fp16, swizzled A
      // operand), row sum, and the row maximum of the raw scores on the way
      auto sweep = [&](float mb, float& sum, float& mx) {
#pragma unroll 1
        for (int c = 0; c < ngrp; ++c) {
          uint32_t r[32];
          tmem_ld32(tl + c * 32, r);
          tmem_ld_wait();
          __align__(16) __half2 hv[16];
          if (kbase + c * 32 + 32 <= jb.nk) {                 // full 32-key group (warp-uniform): no masking
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s0 = __uint_as_float(r[2 * i]), s1 = __uint_as_float(r[2 * i + 1]);
              mx = fmaxf(mx, fmaxf(s0, s1));
              const float p0 = ex2_approx(fmaf(s0, sl2, -mb));
              const float p1 = ex2_approx(fmaf(s1, sl2, -mb));
              hv[i] = __floats2half2_rn(p0, p1);
              sum += p0 + p1;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int key = kbase + c * 32 + 2 * i;
              const float s0 = __uint_as_float(r[2 * i]), s1 = __uint_as_float(r[2 * i + 1]);
              if (key < jb.nk) mx = fmaxf(mx, s0);
              if (key + 1 < jb.nk) mx = fmaxf(mx, s1);
              const float p0 = key < jb.nk ? ex2_approx(fmaf(s0, sl2, -mb)) : 0.f;
              const float p1 = key + 1 < jb.nk ? ex2_approx(fmaf(s1, sl2, -mb)) : 0.f;
              hv[i] = __floats2half2_rn(p0, p1);
              sum += p0 + p1;
            }
          }
          uint8_t* tile = prow + (c >> 1) * TILE_BYTES;
          const int ch0 = (c & 1) * 4;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(tile + (((ch0 + g) ^ (row & 7)) << 4)) = reinterpret_cast<const uint4*>(hv)[g];
        }
      };
      float corr = 1.f, sum = 0.f;
      long long t2 = t1;
      if (LAZY && j > 0) {
        // PV(j-1) has retired (it precedes S(j) on the tensor pipe): the P tiles may be overwritten, O is stable.
        // SINGLE sweep with the current scaling reference m_run; only if some row's maximum outgrew it by 2^8 (rare
        // after the first chunk) is O rescaled in TMEM and the sweep repeated with the new reference.
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        float mx = -INFINITY;
        sweep(m_run * sl2, sum, mx);
        t2 = dbgt ? clock64() : 0;
        const bool grow = (mx - m_run) * sl2 > 8.f;
        if (__any_sync(0xffffffffu, grow)) {
          if (grow) { corr = ex2_approx((m_run - mx) * sl2); m_run = mx; l_run *= corr; }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tl + 128 + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st32(tl + 128 + c * 32, r);
          }
          tmem_st_wait();
          sum = 0.f;
          float unused = -INFINITY;
          sweep(m_run * sl2, sum, unused);
        }
      } else {
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < ngrp; ++c) {
          uint32_t r[32];
          tmem_ld32(tl + c * 32, r);
          tmem_ld_wait();
          if (kbase + c * 32 + 32 <= jb.nk) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kbase + c * 32 + i < jb.nk) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
        }
        t2 = dbgt ? clock64() : 0;
        if (LAZY) {
          m_run = mx;                                         // first chunk: the reference is its own maximum
        } else {
          const float m_new = fmaxf(m_run, mx);               // finite: every chunk holds >= 1 valid key
          corr = ex2_approx((m_run - m_new) * sl2);
          m_run = m_new;
          l_run *= corr;
        }
        float unused = -INFINITY;
        sweep(m_run * sl2, sum, unused);
      }
      l_run += sum;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(p_ready);
      const long long t3 = dbgt ? clock64() : 0;
      long long t4 = t3;
      if (!LAZY) {
        if (__any_sync(0xffffffffu, corr != 1.f)) {          // the running maxima settle after the first chunks
#pragma unroll
          for (int i = 0; i < 64; ++i) o[i] *= corr;
        }
        mbar_wait(o_full, j & 1);
        t4 = dbgt ? clock64() : 0;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tl + 128 + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(r[i]);
        }
        tc_fence_before();     // order these TMEM reads before the next chunk's MMAs (released through p_ready / s_full)
      }
      if (dbgt) {
        const long long t5 = clock64();
        dbg[0] += t1 - t0; dbg[1] += t2 - t1; dbg[2] += t3 - t2; dbg[3] += t4 - t3; dbg[4] += t5 - t4; dbg[5] += 1;
      }
    }
    if (LAZY && warp_live) {                                 // the whole sum sits in TMEM: one read at the end
      mbar_wait(o_full, (n_chunks - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tl + 128 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(r[i]);
      }
      tc_fence_before();
    }
    if (q0 + row < jb.nq) {
      const float inv = 1.f / l_run;
      __half* op = ctx + (int64_t)(jb.q_row + q0 + row) * ldo + head * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        __align__(16) __half2 hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hv[i] = __floats2half2_rn(o[g * 8 + 2 * i] * inv, o[g * 8 + 2 * i + 1] * inv);
        reinterpret_cast<uint4*>(op)[g] = *reinterpret_cast<const uint4*>(hv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (default, r02).  ncu of the kernel above at batch 32 / 64 (profiles/README.md): MUFU pipe 30 % busy,
// tensor pipe 19 %; a third of the 3072 CTAs exit at once, and the live ones spend their time in per-CTA set-up (barrier
// init, TMEM allocation, first Q/K/V round trip), the O read-out tail, the wave quantisation of CTAs whose work differs
// 4 x (150- vs 662-token images) and - in steady state - in the serial S -> softmax -> PV chain of a single row-owning
// thread (stall samples: 25 % waiting for S, the exp2 sweep itself latency- not MUFU-bound at two softmax warps per SMSP).
//  * TWO resident CTAs per SM walk a host-built list of (job, head, query tile) items, most expensive first, in snake
//    order over the CTAs (round r: CTA b takes item r*G + b, odd rounds G-1-b): barriers / TMEM are set up once, the TMA
//    warp prefetches the next item's Q/K/V under the current item's tail.  Barriers run on global counters (items n,
//    chunks g); q_empty releases the Q tile.  Item descriptors are complete (no dependent job-table load) and fetched
//    one item ahead.
//  * EIGHT softmax warps: the two warps of a TMEM lane quarter share its 32 query rows, warp `half` owning the 32-key
//    groups {2 half, 2 half + 1} of every 128-key chunk (its own P tile) and O columns [32 half, +32).  A row's running
//    maximum / sum are combined through two spare TMEM columns (tcgen05.st -> named barrier -> tcgen05.ld): the chain
//    per chunk halves and four softmax warps per SMSP keep the MUFU pipe fed.
struct AttnItem { int q_row, rows, k_row, nk, q_col, k_col, v_col, o_col; };   // 32 bytes, see lg_attn_items()

__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}
// value of the partner warp (same TMEM lanes, other column half) for this thread's row
__device__ __forceinline__ float row_exchange(uint32_t col_own, uint32_t col_other, float v, int bar_id) {
  tmem_st1(col_own, v);
  tmem_st_wait();
  tc_fence_before();
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  tc_fence_after();
  const float o = tmem_ld1(col_other);
  tmem_ld_wait();
  return o;
}

__global__ void __launch_bounds__(320, 2) lg_attn_persist_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                 const AttnItem* __restrict__ items, int n_items,
                                                                 __half* __restrict__ ctx, int ldo, float sl2,
                                                                 int flags, float grow_thr, long long* dbg) {
  pdl_trigger();
  // DV_ATTN_DBG timeline capture (dbg[8] != 0): one thread per CTA stamps its phases into dbg[16 + 8 bid ..]
  long long* dbgc = nullptr;
  if (dbg && threadIdx.x == 64 && dbg[8] != 0) {
    dbgc = dbg + 16 + (long long)blockIdx.x * 8;
    unsigned long long gt; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    dbgc[0] = clock64(); dbgc[6] = (long long)gt; dbgc[7] = (long long)sm;
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if ((int)(smem - smem_raw) > ALIGN_SLACK) __trap();
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = q_full + 1;              // [2]
  uint64_t* kv_empty = kv_full + 2;            // [2]
  uint64_t* s_full = kv_empty + 2;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint64_t* q_empty = o_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x, bid = (int)blockIdx.x;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 8); mbar_init(o_full, 1); mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (dbgc) dbgc[1] = clock64();
  // index of this CTA's item in round r (-1: none; then no later round has one either); descriptors are constants
  // written by a host copy that precedes the whole LightGlue pass, so they may be read before the dependency wait
  auto item_idx = [&](int r) -> int {
    const int idx = r * G + ((r & 1) ? G - 1 - bid : bid);
    return idx < n_items ? idx : -1;
  };
  auto load_item = [&](int idx, AttnItem& it) {
    if (idx < 0) return;
    const int4* p4 = reinterpret_cast<const int4*>(items + idx);
    const int4 a = __ldg(p4), b = __ldg(p4 + 1);
    it.q_row = a.x; it.rows = a.y; it.k_row = a.z; it.nk = a.w;
    it.q_col = b.x; it.k_col = b.y; it.v_col = b.z; it.o_col = b.w;
  };
  int idx = item_idx(0);
  AttnItem cur = {0, 0, 0, 0, 0, 0, 0, 0}, nxt = cur;
  load_item(idx, cur);
  pdl_wait();
  if (dbgc) dbgc[2] = clock64();

  if (warp == 0) {
    if (elect_one_sync()) {
      int g = 0;
      for (int n = 0; idx >= 0; ++n) {
        const int nidx = item_idx(n + 1);
        load_item(nidx, nxt);
        const int n_chunks = (cur.nk + 127) >> 7;
        if (n > 0) mbar_wait(q_empty, (n - 1) & 1);          // every S product of the previous item has read Q
        mbar_arrive_expect_tx(q_full, TILE_BYTES);
        tma_load_2d(smem + OFF_Q, &tmQKV, q_full, cur.q_col, cur.q_row);
        for (int j = 0; j < n_chunks; ++j, ++g) {
          const int s = g & 1;
          mbar_wait(&kv_empty[s], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
          tma_load_2d(smem + OFF_K + s * TILE_BYTES, &tmQKV, &kv_full[s], cur.k_col, cur.k_row + j * 128);
          tma_load_2d(smem + OFF_V + s * TILE_BYTES, &tmQKV, &kv_full[s], cur.v_col, cur.k_row + j * 128);
        }
        // A/B switch (DV_ATTN_PREFETCH=1, off by default): pull the next item's Q and first two K/V chunks into L2 once
        // every K/V load of this item is queued.  Item boundaries wait ~1.1 k clk for the next first S although its
        // operands were requested a sweep earlier (DV_ATTN_DBG), but the prefetch changes nothing: the 4-6 k clk a TMA
        // load takes under this kernel's load is L2 -> SM delivery (15 B/clk per SM with <= 128 KB in flight), not DRAM.
        if ((flags & 1) && nidx >= 0) {
          tma_prefetch_2d(&tmQKV, nxt.q_col, nxt.q_row);
          tma_prefetch_2d(&tmQKV, nxt.k_col, nxt.k_row);
          tma_prefetch_2d(&tmQKV, nxt.v_col, nxt.k_row);
          if (nxt.nk > 128) {
            tma_prefetch_2d(&tmQKV, nxt.k_col, nxt.k_row + 128);
            tma_prefetch_2d(&tmQKV, nxt.v_col, nxt.k_row + 128);
          }
        }
        cur = nxt; idx = nidx;
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc_f16_f32(128, 128);                 // A, B K-major
      constexpr uint32_t idesc_pv = make_idesc_f16_f32(128, 64) | (1u << 16);    // B (= V) MN-major
      const uint64_t dq = make_desc_sw128(smem_u32(smem + OFF_Q));
      const uint64_t dp = make_desc_sw128(smem_u32(smem + OFF_P));
      int g = 0;
      for (int n = 0; idx >= 0; ++n) {
        const int nidx = item_idx(n + 1);
        load_item(nidx, nxt);
        const int n_chunks = (cur.nk + 127) >> 7;
        mbar_wait(q_full, n & 1);
        for (int j = 0; j < n_chunks; ++j, ++g) {
          const int s = g & 1;
          mbar_wait(&kv_full[s], (g >> 1) & 1);
          tc_fence_after();
          const uint64_t dk = make_desc_sw128(smem_u32(smem + OFF_K + s * TILE_BYTES));
          const uint64_t dv = make_desc_sw128(smem_u32(smem + OFF_V + s * TILE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)                                             // head_dim 64 = 4 k-steps
            tc_mma_f16(tmem_base, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
          tc_commit(s_full);
          if (j == n_chunks - 1) tc_commit(q_empty);
          mbar_wait(p_ready, g & 1);
          tc_fence_after();
          const int pv_steps = min(8, (cur.nk - j * 128 + 15) >> 4);               // only the k-steps that hold valid keys
          for (int k = 0; k < pv_steps; ++k)
            // A: P tile (k >> 2), 32-byte step inside its 128-byte rows.  B: V rows [16k, 16k+16) = +2048 bytes.
            tc_mma_f16(tmem_base + 128, dp + (uint64_t)((k >> 2) * (TILE_BYTES >> 4) + (k & 3) * 2),
                       dv + (uint64_t)(k * 128), idesc_pv, (uint32_t)((j | k) != 0));
          tc_commit(o_full);
          tc_commit(&kv_empty[s]);
        }
        cur = nxt; idx = nidx;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: two threads (warps) per query row
    const int q = warp & 3;                     // TMEM lane quarter
    const int half = (warp - 2) >> 2;           // key groups {2 half, 2 half + 1} of a chunk; O columns [32 half, +32)
    const int row = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t x_own = tl + 192 + half, x_oth = tl + 192 + (half ^ 1);     // exchange columns: row maximum
    const uint32_t y_own = tl + 194 + half, y_oth = tl + 194 + (half ^ 1);     // ... row sum
    const int bar_id = 1 + q;
    uint8_t* ptile = smem + OFF_P + half * TILE_BYTES + row * 128;             // this warp's P tile row
    const bool dbgt = dbg && bid == 3 && warp == 4 && lane == 0;               // DV_ATTN_DBG: one thread of lane quarter 0 (always live)
    long long d_s = 0, d_o = 0, d_sw = 0, d_x = 0, d_item = 0, d_ch = 0, d_end = 0;
    int g = 0;
    for (int n = 0; idx >= 0; ++n) {
      const int nidx = item_idx(n + 1);
      load_item(nidx, nxt);
      const int n_chunks = (cur.nk + 127) >> 7;
      const bool warp_live = q * 32 < cur.rows;
      float m_run = -INFINITY, l_run = 0.f;
      if (dbgt) d_item += 1;
      for (int j = 0; j < n_chunks; ++j, ++g) {
        const long long t0 = dbgt ? clock64() : 0;
        if (dbgt && j == 0 && n > 0) dbg[14] += t0 - dbg[13];            // previous item's end -> this item's first wait
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        if (dbgt) { d_s += clock64() - t0; d_ch += 1; if (j == 0) dbg[15] += clock64() - t0; }
        if (dbgc && g == 0) dbgc[3] = clock64();
        const int kbase = j * 128;
        const int ngrp = min(4, (cur.nk - kbase + 31) >> 5);   // 32-key groups holding valid keys (PV reads no further)
        if (!warp_live) {                                      // every row of this lane quarter is padding
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cnt(p_ready);
          continue;
        }
        const int c0 = 2 * half, c1 = min(c0 + 2, ngrp);       // this warp's groups (none: chunk tail of <= 64 keys)
        // P = exp2(s * sl2 - mb) of the own groups to shared memory (fp16, swizzled A operand), row sum, raw row maximum
        auto sweep = [&](float mb, float& sum, float& mx) {
#pragma unroll 1
          for (int c = c0; c < c1; ++c) {
            uint32_t r[32];
            tmem_ld32(tl + c * 32, r);
            tmem_ld_wait();
            __align__(16) __half2 hv[16];
            if (kbase + c * 32 + 32 <= cur.nk) {                // full 32-key group (warp-uniform): no masking
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float s0 = __uint_as_float(r[2 * i]), s1 = __uint_as_float(r[2 * i + 1]);
                mx = fmaxf(mx, fmaxf(s0, s1));
                const float p0 = ex2_approx(fmaf(s0, sl2, -mb));
                const float p1 = ex2_approx(fmaf(s1, sl2, -mb));
                hv[i] = __floats2half2_rn(p0, p1);
                sum += p0 + p1;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int key = kbase + c * 32 + 2 * i;
                const float s0 = __uint_as_float(r[2 * i]), s1 = __uint_as_float(r[2 * i + 1]);
                if (key < cur.nk) mx = fmaxf(mx, s0);
                if (key + 1 < cur.nk) mx = fmaxf(mx, s1);
                const float p0 = key < cur.nk ? ex2_approx(fmaf(s0, sl2, -mb)) : 0.f;
                const float p1 = key + 1 < cur.nk ? ex2_approx(fmaf(s1, sl2, -mb)) : 0.f;
                hv[i] = __floats2half2_rn(p0, p1);
                sum += p0 + p1;
              }
            }
            const int ch0 = (c & 1) * 4;
#pragma unroll
            for (int gg = 0; gg < 4; ++gg)
              *reinterpret_cast<uint4*>(ptile + (((ch0 + gg) ^ (row & 7)) << 4)) = reinterpret_cast<const uint4*>(hv)[gg];
          }
        };
        float sum = 0.f;
        if (j > 0) {
          // PV(j-1) has retired (it precedes S(j) on the tensor pipe): the P tiles may be overwritten, O is stable.
          // SINGLE sweep with the current scaling reference m_run; only if some row's maximum outgrew it by 2^grow_thr
          // (default 2^12: P stays far inside fp16's 2^16; 1-4 % of the chunks, 1.9 k clk each) is O rescaled in TMEM
          // and the sweep repeated with the new reference.
          const long long t1 = dbgt ? clock64() : 0;
          mbar_wait(o_full, (g - 1) & 1);
          tc_fence_after();
          const long long t2 = dbgt ? clock64() : 0;
          float mx = -INFINITY;
          sweep(m_run * sl2, sum, mx);
          const long long t3 = dbgt ? clock64() : 0;
          mx = fmaxf(mx, row_exchange(x_own, x_oth, mx, bar_id));
          if (dbgt) { const long long t4 = clock64(); d_o += t2 - t1; d_sw += t3 - t2; d_x += t4 - t3; }
          const bool grow = (mx - m_run) * sl2 > grow_thr;     // identical in both warps of the row
          if (__any_sync(0xffffffffu, grow)) {
            const long long tr = dbgt ? clock64() : 0;
            float corr = 1.f;
            if (grow) { corr = ex2_approx((m_run - mx) * sl2); m_run = mx; l_run *= corr; }
            uint32_t r[32];
            tmem_ld32(tl + 128 + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st32(tl + 128 + half * 32, r);
            tmem_st_wait();
            sum = 0.f;
            float unused = -INFINITY;
            sweep(m_run * sl2, sum, unused);
            if (dbgt) { dbg[11] += 1; dbg[12] += clock64() - tr; }
          }
        } else {
          const long long tf = dbgt ? clock64() : 0;
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = c0; c < c1; ++c) {
            uint32_t r[32];
            tmem_ld32(tl + c * 32, r);
            tmem_ld_wait();
            if (kbase + c * 32 + 32 <= cur.nk) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (kbase + c * 32 + i < cur.nk) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
          }
          m_run = fmaxf(mx, row_exchange(x_own, x_oth, mx, bar_id));   // first chunk: the reference is its own maximum
          float unused = -INFINITY;
          sweep(m_run * sl2, sum, unused);
          if (dbgt) dbg[9] += clock64() - tf;
        }
        l_run += sum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cnt(p_ready);
      }
      // PV of the item's last chunk has retired: O is complete, and the P tiles / the O columns may be rewritten by the
      // next item.  EVERY warp waits (a warp that idled through this item may be live in the next one).
      const long long te = dbgt ? clock64() : 0;
      mbar_wait(o_full, (g - 1) & 1);
      tc_fence_after();
      if (dbgt) d_end += clock64() - te;
      const long long tp = dbgt ? clock64() : 0;
      if (warp_live) {
        const float l_tot = l_run + row_exchange(y_own, y_oth, l_run, bar_id);
        uint32_t r[32];
        tmem_ld32(tl + 128 + half * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        if (row < cur.rows) {
          const float inv = 1.f / l_tot;
          __half* op = ctx + (int64_t)(cur.q_row + row) * ldo + cur.o_col + half * 32;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            __align__(16) __half2 hv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              hv[i] = __floats2half2_rn(__uint_as_float(r[gg * 8 + 2 * i]) * inv, __uint_as_float(r[gg * 8 + 2 * i + 1]) * inv);
            reinterpret_cast<uint4*>(op)[gg] = *reinterpret_cast<const uint4*>(hv);
          }
        }
      } else {
        tc_fence_before();
      }
      cur = nxt; idx = nidx;
      if (dbgt) { dbg[10] += clock64() - tp; dbg[13] = clock64(); }
    }
    if (dbgt) {
      dbg[0] += d_s; dbg[1] += d_o; dbg[2] += d_sw; dbg[3] += d_x; dbg[4] += d_ch; dbg[5] += d_item; dbg[6] += d_end;
    }
    if (dbgc) dbgc[4] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (dbgc) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    dbgc[5] = clock64();
    dbgc[6] = (long long)gt - dbgc[6];          // CTA lifetime in ns
    dbg[16 + 8 * 2 * 148 * 4 + blockIdx.x] = (long long)gt;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int g_attn_sms = 148;

int lg_attn_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(lg_attn_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(lg_attn_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(lg_attn_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_attn_sms, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

int plan_lg_attn(CUtensorMap* tm, const __half* qkv, int T_cap) {
  const uint64_t dims[2] = {768, (uint64_t)T_cap}, strides[1] = {1536};
  const uint32_t box[2] = {64, 128};
  return tmap_encode_f16(tm, qkv, 2, dims, strides, box, true);
}

int launch_lg_attn(const CUtensorMap& tm, const AttnJobU* jobs, int n_jobs, int max_nq, __half* ctx, int ldo, float scale,
                   cudaStream_t st) {
  if (n_jobs <= 0) return DV_OK;
  const dim3 grid(cdiv(max_nq, 128), 4, n_jobs);
  static long long* d_dbg = nullptr;
  static const bool want = getenv("DV_ATTN_DBG") != nullptr;       // diagnostics: softmax-thread cycle counters
  static int calls = 0;
  if (want && !d_dbg) { cudaMalloc(&d_dbg, 64); cudaMemset(d_dbg, 0, 64); }
  static const bool lazy = [] { const char* e = getenv("DV_ATTN_LAZY"); return !(e && e[0] == '0'); }();   // A/B switch
  if (lazy)
    lg_attn_umma_kernel<true><<<grid, 192, SMEM_BYTES, st>>>(tm, jobs, ctx, ldo, scale * 1.4426950408889634f, want ? d_dbg : nullptr);
  else
    lg_attn_umma_kernel<false><<<grid, 192, SMEM_BYTES, st>>>(tm, jobs, ctx, ldo, scale * 1.4426950408889634f, want ? d_dbg : nullptr);
  DV_CUDA_OK(cudaGetLastError());
  if (want && ++calls == 18) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[attn dbg] chunks %lld | cycles/chunk: wait S %lld, pass A (max) %lld, pass B (exp+P) %lld, wait O %lld, O acc %lld\n",
            h[5], h[0] / h[5], h[1] / h[5], h[2] / h[5], h[3] / h[5], h[4] / h[5]);
  }
  return DV_OK;
}

// Persistent launch: `items` = device list built by lg_attn_items(), most expensive first.
int launch_lg_attn_persist(const CUtensorMap& tm, const void* items, int n_items, __half* ctx, int ldo, float scale,
                           cudaStream_t st) {
  if (n_items <= 0) return DV_OK;
  const int grid = n_items < 2 * g_attn_sms ? n_items : 2 * g_attn_sms;
  static long long* d_dbg = nullptr;
  static const bool want = getenv("DV_ATTN_DBG") != nullptr;       // diagnostics: cycle counters of one softmax thread
  static int calls = 0;
  constexpr int DBG_WORDS = 16 + 8 * 2 * 148 * 4 + 2 * 148 * 4;    // counters, per-CTA timeline [grid][8], end timestamps
  if (want && !d_dbg) { cudaMalloc(&d_dbg, DBG_WORDS * 8); cudaMemset(d_dbg, 0, DBG_WORDS * 8); }
  // timeline capture of two launches (a self- and a cross-attention of the third LightGlue pass)
  const bool cap = want && (calls == 38 || calls == 39) && grid <= 2 * 148 * 4;
  if (cap) { const long long one = 1; cudaMemcpyAsync(d_dbg + 8, &one, 8, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); }
  // DV_ATTN_PREFETCH=1: L2 prefetch of the next item's first tiles (measured: no gain - 3.83 vs 3.80 ms per 64 pairs - the
  // loaded TMA latency is L2 -> SM delivery, not DRAM; off by default).  DV_ATTN_GROW: log2 of the growth of a row maximum
  // over the scaling reference that triggers a rescale of O in TMEM (default 12: P <= 2^12 in fp16, whose range ends at
  // 2^16; r02 A/B: 8 -> 12 = -0.06 ms per 64 pairs, every parity test unchanged)
  static const int flags = [] { const char* e = getenv("DV_ATTN_PREFETCH"); return (e && e[0] == '1') ? 1 : 0; }();
  static const float grow_thr = [] { const char* e = getenv("DV_ATTN_GROW"); const float v = e ? (float)atof(e) : 12.f;
                                     return v < 1.f ? 1.f : (v > 14.f ? 14.f : v); }();
  DV_CUDA_OK(launch_pdl(lg_attn_persist_kernel, dim3(grid), dim3(320), (size_t)SMEM_BYTES, st, tm,
                        reinterpret_cast<const AttnItem*>(items), n_items, ctx, ldo, scale * 1.4426950408889634f,
                        flags, grow_thr, want ? d_dbg : (long long*)nullptr));
  DV_CUDA_OK(cudaGetLastError());
  if (cap) {
    cudaStreamSynchronize(st);
    std::vector<long long> h(DBG_WORDS);
    cudaMemcpy(h.data(), d_dbg, DBG_WORDS * 8, cudaMemcpyDeviceToHost);
    const long long zero = 0;
    cudaMemcpy(d_dbg + 8, &zero, 8, cudaMemcpyHostToDevice);
    const long long* ends = h.data() + 16 + 8 * 2 * 148 * 4;
    long long end_min = ends[0], end_max = ends[0];
    std::vector<long long> ph[6];
    std::vector<int> per_sm(1024, 0);
    for (int b = 0; b < grid; ++b) {
      const long long* c = h.data() + 16 + 8 * b;
      ph[0].push_back(c[1] - c[0]); ph[1].push_back(c[2] - c[1]); ph[2].push_back(c[3] - c[2]);
      ph[3].push_back(c[4] - c[3]); ph[4].push_back(c[5] - c[4]); ph[5].push_back(c[6]);
      end_min = std::min(end_min, ends[b]); end_max = std::max(end_max, ends[b]);
      if (c[7] >= 0 && c[7] < 1024) per_sm[(int)c[7]]++;
    }
    long long start_min = ends[0] - (h.data() + 16)[6], start_max = start_min;
    for (int b = 0; b < grid; ++b) {
      const long long s0 = ends[b] - (h.data() + 16 + 8 * b)[6];
      start_min = std::min(start_min, s0); start_max = std::max(start_max, s0);
    }
    int sm_hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int v : per_sm) sm_hist[v < 7 ? v : 7]++;
    auto q = [](std::vector<long long>& v, double f) { std::sort(v.begin(), v.end()); return v[(size_t)(f * (v.size() - 1))]; };
    fprintf(stderr, "[attn timeline] launch %d: grid %d items %d | kernel span %lld ns (first CTA start -> last CTA end), starts spread %lld ns, "
            "ends spread %lld ns | SMs holding 0/1/2/3+ CTAs: %d/%d/%d/%d\n", calls, grid, n_items, end_max - start_min, start_max - start_min,
            end_max - end_min, sm_hist[0], sm_hist[1], sm_hist[2], sm_hist[3] + sm_hist[4] + sm_hist[5] + sm_hist[6] + sm_hist[7]);
    const char* nm[6] = {"setup (clk)", "pdl wait (clk)", "first S (clk)", "item loop (clk)", "exit barrier (clk)", "CTA lifetime (ns)"};
    for (int k2 = 0; k2 < 6; ++k2)
      fprintf(stderr, "[attn timeline]   %-18s min %lld  p10 %lld  median %lld  p90 %lld  max %lld\n", nm[k2], q(ph[k2], 0.0), q(ph[k2], 0.1),
              q(ph[k2], 0.5), q(ph[k2], 0.9), q(ph[k2], 1.0));
  }
  ++calls;
  if (want && calls % 18 == 0) {
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, 128, cudaMemcpyDeviceToHost);
    cudaMemset(d_dbg, 0, 128);
    const long long it = h[5] ? h[5] : 1;
    fprintf(stderr, "[attn dbg] CTA 3 warp 4 over 18 launches: chunks %lld items %lld | cycles/chunk: wait S %lld, wait O %lld, sweep %lld, "
            "exchange %lld | per item: first-chunk wait S %lld, first-chunk body %lld, item-end O wait %lld, epilogue %lld, item turn-over %lld | "
            "rescales %lld (%lld clk each)\n", h[4], h[5], h[0] / (h[4] ? h[4] : 1), h[1] / (h[4] ? h[4] : 1),
            h[2] / (h[4] ? h[4] : 1), h[3] / (h[4] ? h[4] : 1), h[15] / it, h[9] / it, h[6] / it, h[10] / it, h[14] / it,
            h[11], h[12] / (h[11] ? h[11] : 1));
  }
  return DV_OK;
}

// Host side of the item list: every (job, head, 128-query tile) of `jobs` as a complete 32-byte descriptor, sorted by
// decreasing cost.  A tile occupies its CTA for (32-key groups of the key side) sweeps whatever its number of live rows,
// so the key count ranks first and the live 32-row quarters (MUFU work) break ties.  Returns the number of items.
int lg_attn_item_bytes() { return (int)sizeof(AttnItem); }
int lg_attn_items(const AttnJobU* jobs, int n_jobs, void* items_out) {
  struct It { int cost; AttnItem it; };
  static thread_local std::vector<It> tmp;
  tmp.clear();
  for (int jn = 0; jn < n_jobs; ++jn) {
    const AttnJobU& jb = jobs[jn];
    const int groups = (jb.nk + 31) >> 5;
    for (int qt = 0; qt * 128 < jb.nq; ++qt) {
      const int rows = jb.nq - qt * 128 < 128 ? jb.nq - qt * 128 : 128;
      const int cost = groups * 8 + ((rows + 31) >> 5);
      for (int h = 0; h < 4; ++h)
        tmp.push_back({cost, {jb.q_row + qt * 128, rows, jb.k_row, jb.nk, jb.q_col + h * 64, jb.k_col + h * 64,
                              jb.v_col + h * 64, h * 64}});
    }
  }
  std::stable_sort(tmp.begin(), tmp.end(), [](const It& x, const It& y) { return x.cost > y.cost; });
  AttnItem* out = reinterpret_cast<AttnItem*>(items_out);
  for (size_t i = 0; i < tmp.size(); ++i) out[i] = tmp[i].it;
  return (int)tmp.size();
}

}  // namespace dv
