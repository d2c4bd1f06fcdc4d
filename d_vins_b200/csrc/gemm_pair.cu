// Weights-resident tcgen05 GEMM on CTA PAIRS (cta_group::2):  C[M,N] = A[M,K] * W[N,K]^T (+bias, +residual, ReLU,
// rotary) for the K <= 512 linear layers of LightGlue / MixVPR whose N is a multiple of 256.
//
// Why (r02 measurements, profiles/README.md): with one CTA per SM the K = 512 layers can only keep a 128-column slab
// of W resident (128 KB), so every 16 KB A stage feeds just 256 cycles of N = 128 MMAs and the MMA warp starves on A
// (accumulator wait 46 k of 72 k cycles per CTA for the 53248 x 512 x 512 FFN GEMM: ~21 B/clk of A per SM is all the
// 4-stage ring pulls through L2/HBM latency).  A CTA pair shares the B operand: each CTA keeps HALF of a 256-column
// slab (128 columns x K, <= 128 KB) and the pair issues M = 256, N = 256 MMAs - every A stage now feeds 512 cycles of
// full-rate (N = 256) tensor work, the slab count and with it the L2 -> SM traffic of A halve, and the B-operand
// shared-memory reads per SM halve as well.
//
// Pair p (cluster of two CTAs on one TPC) owns slab (p % n_slabs) and walks the 256-row super-tiles (p / n_slabs) +
// i * group; CTA r of the pair loads / drains rows [256 t + 128 r, + 128).  Both CTAs run a TMA producer (warp 0;
// completion bytes of both land on the LEADER's "full" barrier) and the staged epilogue of gemm_wres.cu (warps 2-9,
// own TMEM lanes, TMA stores); only the leader (cluster rank 0) issues MMAs, and its tcgen05.commit multicasts the
// "stage free" / "accumulator full" arrivals to both CTAs.  Accumulators: two buffers of 256 TMEM columns.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

#define DV_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

struct PairCfg {
  int s_tiles;       // 256-row super-tiles
  int n_slabs, group, stages;
  int epi_warps;     // 8: warp (q, h) drains column half h of every 128-column sub-tile; 4: warp q drains both halves
  int uw;            // columns per epilogue unit: 64, or 32 (fp32 + fp16 outputs: eight warps fit beside a 128 KB half-slab)
  uint32_t a_off, sb32_off, sb16_off, bar_off, bias_off, smem_bytes;
};

// UW: columns per epilogue unit (one staging box per warp).  64: the staged epilogue of gemm_wres.cu.  32 (fp32 + fp16
// outputs with an fp32 residual, K = 512): 6 KB of staging per warp instead of 12, so EIGHT warps drain the tile and
// twice as many residual boxes are in flight (the 4-warp version spent its time waiting for them one at a time).
template <bool F32, bool F16, int UW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
    umma_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                          const __grid_constant__ CUtensorMap tmR32, const __grid_constant__ CUtensorMap tmR16,
                          const GemmParams p, const PairCfg c) {
  pdl_trigger();
  const long long t_entry = p.dbg ? clock64() : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + c.bar_off);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* rbar = acc_empty + 2;                  // one per epilogue warp: residual boxes landed
  uint64_t* w_bar = rbar + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* sbias = reinterpret_cast<float*>(smem + c.bias_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const EpiParams& ep = p.epi;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1;
  const int slab_idx = pair % c.n_slabs;
  const int j0 = pair / c.n_slabs;
  const int col_base = slab_idx * 256;             // first output column of this pair

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (ep.out32) prefetch_tmap(&tmO32);
    if (ep.out16) prefetch_tmap(&tmO16);
    if (ep.res32) prefetch_tmap(&tmR32);
    if (ep.res16) prefetch_tmap(&tmR16);
    for (int s = 0; s < c.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], (uint32_t)(2 * c.epi_warps)); }
    for (int e = 0; e < 8; ++e) mbar_init(&rbar[e], 1);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sbias[i] = ep.bias ? ep.bias[col_base + i] : 0.f;
  if (warp == 1) tmem_alloc_pair(tmem_ptr, 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      // the half-slab of W is a constant: its load overlaps the previous kernel's tail (PDL).  Both CTAs' bytes are
      // credited to the leader's w_bar, which the MMA thread waits on.
      const uint32_t wbar_leader = mapa_u32(smem_u32(w_bar), 0);
      if (rank == 0) mbar_arrive_expect_tx(w_bar, (uint32_t)p.num_kb * 32768u);
      for (int kb = 0; kb < p.num_kb; ++kb)
        tma_load_2d_pair(smem + (uint32_t)kb * 16384u, &tmB, wbar_leader, kb * 64, col_base + (int)rank * 128);
      pdl_wait();
      int kc = 0;
      for (int st = j0; st < c.s_tiles; st += c.group) {
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % c.stages;
          mbar_wait(&empty[s], ((kc / c.stages) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 32768u);
          tma_load_2d_pair(smem + c.a_off + (uint32_t)s * 16384u, &tmA, mapa_u32(smem_u32(&full[s]), 0), kb * 64,
                           st * 256 + (int)rank * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(256, 256);
      mbar_wait(w_bar, 0);
      int kc = 0, it = 0;
      for (int st = j0; st < c.s_tiles; st += c.group, ++it) {
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * 256);
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % c.stages;
          mbar_wait(&full[s], (kc / c.stages) & 1);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(smem_u32(smem + c.a_off + (uint32_t)s * 16384u));
          const uint64_t db = make_desc_sw128(smem_u32(smem + (uint32_t)kb * 16384u));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          tc_commit_pair(&empty[s], 3);
        }
        tc_commit_pair(&acc_full[a], 3);
      }
    }
  } else if (UW == 16) {
    if constexpr (UW == 16 && F32 && F16) {
      // -------------------------------------------------------------- epilogue with residual PREFETCH (ffn.3 signature:
      // fp32 residual in, fp32 + fp16 out, possibly in place).  16-column units: per warp a residual box R and separate
      // output boxes S (fp32) / Hb (fp16) - 5 KB - so the residual of unit u + 1 (or of the next tile's unit 0) is
      // requested as soon as unit u's has been read into registers and travels under unit u's TMEM read, arithmetic and
      // stores.  With the in-place boxes of the wider units the load could only be issued after the previous store had
      // finished reading the box, and its full HBM latency sat on every unit's critical path.
      const int e = warp - 2;
      const int q = warp & 3;                    // TMEM lane quarter
      const int h0 = e >> 2;                     // column half [128 h0, +128) of the slab: eight units of 16 columns
      uint8_t* R = smem + c.sb32_off + e * 4096;
      uint8_t* S = R + 2048;
      uint8_t* Hb = smem + c.sb16_off + e * 1024;
      uint64_t* rb = &rbar[e];
      const uint32_t sw64 = (((uint32_t)lane >> 1) & 3u);        // SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3
      uint32_t rphase = 0;
      auto issue_res = [&](int st, int u) {
        const int r0 = st * 256 + (int)rank * 128 + q * 32;
        if (r0 < p.M && lane == 0) {
          mbar_arrive_expect_tx(rb, 2048u);
          tma_load_2d(R, &tmR32, rb, col_base + h0 * 128 + u * 16, r0);
        }
      };
      int it = 0;
      if (j0 < c.s_tiles) issue_res(j0, 0);
      for (int st = j0; st < c.s_tiles; st += c.group, ++it) {
        const int a = it & 1;
        const int row0 = st * 256 + (int)rank * 128 + q * 32;
        const bool active = row0 < p.M;                  // warp-uniform
        mbar_wait(&acc_full[a], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + h0 * 128);
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
          const int lcol = h0 * 128 + u * 16;
          float v[16];
          if (active) {
            mbar_wait(rb, rphase);
            rphase ^= 1u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 t = *reinterpret_cast<const float4*>(R + (uint32_t)lane * 64u + ((((uint32_t)g) ^ sw64) << 4));
              v[g * 4 + 0] = t.x; v[g * 4 + 1] = t.y; v[g * 4 + 2] = t.z; v[g * 4 + 3] = t.w;
            }
          }
          __syncwarp();                                  // every lane has read R: the next box may land in it
          if (u < 7) issue_res(st, u + 1);
          else if (st + c.group < c.s_tiles) issue_res(st + c.group, 0);
          if (active) {
            uint32_t r[16];
            tmem_ld16(taddr + u * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b4 = *reinterpret_cast<const float4*>(&sbias[lcol + g * 4]);   // broadcast
              v[g * 4 + 0] += __uint_as_float(r[g * 4 + 0]) + b4.x;
              v[g * 4 + 1] += __uint_as_float(r[g * 4 + 1]) + b4.y;
              v[g * 4 + 2] += __uint_as_float(r[g * 4 + 2]) + b4.z;
              v[g * 4 + 3] += __uint_as_float(r[g * 4 + 3]) + b4.w;
            }
            if (ep.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
          }
          if (u == 7) tc_fence_before();                 // last TMEM read of this accumulator buffer
          if (lane == 0) bulk_wait_read0();              // the previous stores finished READING S / Hb
          __syncwarp();
          if (active) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<float4*>(S + (uint32_t)lane * 64u + ((((uint32_t)g) ^ sw64) << 4)) =
                  make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              __align__(16) __half2 hv[4];
#pragma unroll
              for (int x = 0; x < 4; ++x) hv[x] = __floats2half2_rn(v[g * 8 + 2 * x], v[g * 8 + 2 * x + 1]);
              *reinterpret_cast<uint4*>(Hb + (uint32_t)lane * 32u + (uint32_t)g * 16u) = *reinterpret_cast<const uint4*>(hv);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (u == 7) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[a]), 0));
            if (active) {
              tma_store_2d(&tmO32, S, col_base + lcol, row0);
              tma_store_2d(&tmO16, Hb, col_base + lcol, row0);
              bulk_commit();
            }
          }
        }
      }
      if (lane == 0) bulk_wait0();
    }
  } else if (warp - 2 < c.epi_warps) {
    // ------------------------------------------------------------------ epilogue: 8 (or 4) warps per CTA
    const int e = warp - 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    constexpr int UWG = UW < 32 ? 32 : UW;     // (the 16-column prefetch epilogue above never reaches this branch)
    constexpr int NCI = UWG / 32;              // 32-column TMEM loads per unit
    constexpr uint32_t B32_BYTES = UWG * 32 * 4, B16_BYTES = UWG * 32 * 2;
    const int nh = c.epi_warps == 8 ? 1 : 2;   // UW = 64: column halves this warp walks per 128-column sub-tile
    const int h0 = c.epi_warps == 8 ? (e >> 2) : 0;
    uint8_t* b32 = smem + c.sb32_off + e * B32_BYTES;
    uint8_t* b16 = smem + c.sb16_off + e * B16_BYTES;
    uint64_t* rb = &rbar[e];
    const uint32_t rowoff = (uint32_t)lane * 128u;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    // fp16 staging offset of 16-byte chunk `ch` of this lane's row: 128-byte rows / SWIZZLE_128B (UW = 64) or 64-byte
    // rows / SWIZZLE_64B (UW = 32: chunk ^= address bits [7:8] = (row >> 1) & 3)
    auto off16 = [&](int ch) -> uint32_t {
      if (UW == 64) return rowoff + (((uint32_t)ch << 4) ^ swz);
      return (uint32_t)lane * 64u + ((((uint32_t)ch) ^ (((uint32_t)lane >> 1) & 3u)) << 4);
    };
    const bool has_res = (F32 && ep.res32) || (F16 && ep.res16);
    const int n_sub = UW == 64 ? 2 * nh : 4;   // drain units per tile of this warp
    uint32_t rphase = 0;
    int it = 0;
    long long t_rd = 0, t_acc = 0, t_work = 0, t_all0 = p.dbg ? clock64() : 0;     // DV_GEMM_DBG cycle counters
    for (int st = j0; st < c.s_tiles; st += c.group, ++it) {
      const int a = it & 1;
      const int row0 = st * 256 + (int)rank * 128 + q * 32;   // first row of this warp
      const bool active = row0 < p.M;                    // warp-uniform
      const bool writer = row0 + lane < p.M;
      const bool rope_tile = ep.rope16 && col_base < ep.rope_cols && active;
      // (cos_j, sin_j) of this row's 32 rotary angles as 64 fp16 values: 8 x 16-byte loads of one 128-byte line
      uint4 rt[8];
      if (rope_tile && writer) {
        const uint4* rp = reinterpret_cast<const uint4*>(ep.rope16 + (long)(row0 + lane) * 64);
#pragma unroll
        for (int g = 0; g < 8; ++g) rt[g] = __ldg(rp + g);
      } else {
        const __half2 id = __floats2half2_rn(1.f, 0.f);
        const uint32_t idu = *reinterpret_cast<const uint32_t*>(&id);
#pragma unroll
        for (int g = 0; g < 8; ++g) rt[g] = make_uint4(idu, idu, idu, idu);
      }
      for (int sub = 0; sub < n_sub; ++sub) {
        // first column of this unit inside the slab
        const int lcol = UW == 64 ? (sub / nh) * 128 + (h0 + sub % nh) * 64 : h0 * 128 + sub * 32;
        const int colw = col_base + lcol;                // ... and in the output matrix
        long long t0 = p.dbg ? clock64() : 0;
        if (lane == 0) bulk_wait_read0();                // previous TMA stores finished READING the staging boxes
        __syncwarp();
        if (p.dbg) { const long long t1 = clock64(); t_rd += t1 - t0; t0 = t1; }
        if (active && has_res && lane == 0) {
          uint32_t bytes = 0;
          if (F32 && ep.res32) bytes += B32_BYTES;
          if (F16 && ep.res16) bytes += B16_BYTES;
          mbar_arrive_expect_tx(rb, bytes);
          if (F32 && ep.res32) {
#pragma unroll
            for (int ci = 0; ci < NCI; ++ci) tma_load_2d(b32 + ci * 4096, &tmR32, rb, colw + ci * 32, row0);
          }
          if (F16 && ep.res16) tma_load_2d(b16, &tmR16, rb, colw, row0);
        }
        if (sub == 0) {
          mbar_wait(&acc_full[a], (it >> 1) & 1);
          tc_fence_after();
        }
        if (p.dbg) { const long long t1 = clock64(); t_acc += t1 - t0; t0 = t1; }
        const bool rope = rope_tile && colw < ep.rope_cols;
        if (active) {
          if (has_res) { mbar_wait(rb, rphase); rphase ^= 1u; }
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + lcol);
#pragma unroll
          for (int ci = 0; ci < NCI; ++ci) {
            uint32_t r[32];
            tmem_ld32(taddr + ci * 32, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b4 = *reinterpret_cast<const float4*>(&sbias[lcol + ci * 32 + g * 4]);   // broadcast
              v[g * 4 + 0] = __uint_as_float(r[g * 4 + 0]) + b4.x;
              v[g * 4 + 1] = __uint_as_float(r[g * 4 + 1]) + b4.y;
              v[g * 4 + 2] = __uint_as_float(r[g * 4 + 2]) + b4.z;
              v[g * 4 + 3] = __uint_as_float(r[g * 4 + 3]) + b4.w;
            }
            if (F32 && ep.res32) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 t = *reinterpret_cast<const float4*>(b32 + ci * 4096 + rowoff + (((uint32_t)g << 4) ^ swz));
                v[g * 4 + 0] += t.x; v[g * 4 + 1] += t.y; v[g * 4 + 2] += t.z; v[g * 4 + 3] += t.w;
              }
            }
            if (F16 && ep.res16) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 t = *reinterpret_cast<const uint4*>(b16 + off16(ci * 4 + g));
                const __half2* h2 = reinterpret_cast<const __half2*>(&t);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const float2 f = __half22float2(h2[x]);
                  v[g * 8 + 2 * x] += f.x; v[g * 8 + 2 * x + 1] += f.y;
                }
              }
            }
            if (ep.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (rope) {
              const int cr = UW == 64 ? ci : (lcol >> 5) & 1;      // which half of the head's 64 columns (16 angles each)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 t4 = cr ? rt[4 + g] : rt[g];            // angles cr * 16 + g * 4 + {0..3}
                const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const float2 csn = __half22float2(*reinterpret_cast<const __half2*>(&tw[x]));   // (cos, sin)
                  const int jj = g * 4 + x;
                  const float x0 = v[2 * jj], x1 = v[2 * jj + 1];
                  v[2 * jj] = x0 * csn.x - x1 * csn.y;
                  v[2 * jj + 1] = x1 * csn.x + x0 * csn.y;
                }
              }
            }
            if (F32 && ep.out32) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(b32 + ci * 4096 + rowoff + (((uint32_t)g << 4) ^ swz)) =
                    make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            }
            if (F16 && ep.out16) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                __align__(16) __half2 hv[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) hv[x] = __floats2half2_rn(v[g * 8 + 2 * x], v[g * 8 + 2 * x + 1]);
                *reinterpret_cast<uint4*>(b16 + off16(ci * 4 + g)) = *reinterpret_cast<const uint4*>(hv);
              }
            }
          }
        }
        // staging writes become visible to the async proxy; after the last unit the accumulator buffer is free:
        // every epilogue warp of BOTH CTAs arrives on the leader's barrier (the MMA thread lives there)
        if (sub == n_sub - 1) tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (sub == n_sub - 1) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[a]), 0));
          if (active) {
            if (F32 && ep.out32) {
#pragma unroll
              for (int ci = 0; ci < NCI; ++ci) tma_store_2d(&tmO32, b32 + ci * 4096, colw + ci * 32, row0);
            }
            if (F16 && ep.out16) tma_store_2d(&tmO16, b16, colw, row0);
            bulk_commit();
          }
        }
        if (p.dbg) t_work += clock64() - t0;
      }
    }
    if (p.dbg && blockIdx.x == 0 && e == 0 && lane == 0) {
      p.dbg[0] = t_rd; p.dbg[1] = t_acc; p.dbg[2] = t_work; p.dbg[3] = clock64() - t_all0; p.dbg[4] = it;
      p.dbg[5] = t_all0 - t_entry;
    }
    if (lane == 0) bulk_wait0();
    if (p.dbg && blockIdx.x == 0 && e == 0 && lane == 0) p.dbg[6] = clock64() - t_entry;
  }
  // neither CTA may exit (or free TMEM) while the other can still signal its barriers / multicast into it
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512u);
  }
}

static int g_sms_pair = 148;
static int g_use_pair = 2;                             // DV_GEMM_PAIR: 0 off, 1 all eligible shapes, 2 only K > 256
static int g_pair_narrow = 1;                          // DV_GEMM_PAIR_NARROW: 0 never, 1 fp32 + fp16 outputs, 2 fp16-only too
static int g_pair_prefetch = 1;                        // DV_GEMM_PAIR_PREFETCH=0: no 16-column residual-prefetch epilogue (A/B)
static constexpr uint32_t PAIR_SMEM_MAX = 232448;      // 227 KB opt-in limit per CTA

int gemm_pair_init() {
  { const char* e = getenv("DV_GEMM_PAIR"); if (e) g_use_pair = atoi(e); }
  { const char* e = getenv("DV_GEMM_PAIR_NARROW"); if (e) g_pair_narrow = atoi(e); }
  { const char* e = getenv("DV_GEMM_PAIR_PREFETCH"); if (e) g_pair_prefetch = atoi(e); }
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_pair_kernel<true, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_pair_kernel<true, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_pair_kernel<true, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_pair_kernel<true, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_pair_kernel<false, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_pair_kernel<false, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_MAX));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_sms_pair, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

static bool pair_config(const GemmPlan& pl, long m_tiles, PairCfg* c) {
  const GemmParams& p = pl.p;
  const EpiParams& ep = p.epi;
  if (!g_use_pair || !pl.staged) return false;              // same operand / alignment rules as the staged kernel
  if ((p.K & 63) || p.K > 512 || (p.N & 255)) return false;
  if (g_use_pair == 2 && p.K <= 256) return false;
  if (ep.rope_cs && (!ep.rope16 || (ep.rope_cols & 127))) return false;
  if (m_tiles < 16) return false;
  const bool f32 = ep.out32 || ep.res32, f16 = ep.out16 || ep.res16;
  const uint32_t fixed = 1024u /*alignment slack*/ + 512u /*barriers*/ + 1024u /*bias*/;
  const uint32_t wbytes = 128u * (uint32_t)p.K * 2u;        // this CTA's half of the 256-column slab
  const int pairs = g_sms_pair / 2;
  // 32-column units need 64-byte fp16 boxes: no fp16 residual (never used with fp32 + fp16 outputs) and no rotary
  // with fp16-only outputs kept simple; rope_cols handled for both widths
  const bool narrow_ok = !ep.res16 && ((f32 && f16 && g_pair_narrow >= 1) || (!f32 && f16 && g_pair_narrow >= 2));
  // ffn.3 signature (fp32 residual in, fp32 + fp16 out, no fp16 residual / rotary): 16-column units with residual prefetch
  if (g_pair_prefetch && f32 && f16 && ep.res32 && ep.out32 && ep.out16 && !ep.res16 && !ep.rope16 && !ep.rope_cs) {
    const uint32_t staging = 8u * 4096u + 8u * 1024u;
    if (wbytes + staging + fixed + 2u * 16384u <= PAIR_SMEM_MAX) {
      int stages = (int)((PAIR_SMEM_MAX - wbytes - staging - fixed) / 16384u);
      if (stages > 8) stages = 8;
      c->epi_warps = 8; c->uw = 16;
      c->n_slabs = p.N / 256;
      if (c->n_slabs > pairs) return false;
      c->s_tiles = (int)((m_tiles + 1) / 2);
      c->group = pairs / c->n_slabs;
      if (c->group > c->s_tiles) c->group = c->s_tiles;
      c->stages = stages;
      c->a_off = wbytes;
      c->sb32_off = c->a_off + (uint32_t)stages * 16384u;
      c->sb16_off = c->sb32_off + 8u * 4096u;
      c->bar_off = c->sb16_off + 8u * 1024u;
      c->bias_off = c->bar_off + 512u;
      c->smem_bytes = c->bias_off + 1024u + 1024u;
      return true;
    }
  }
  for (int cand = 0; cand < 3; ++cand) {                    // (8 warps, 64), (8 warps, 32), (4 warps, 64)
    const int epi = cand == 2 ? 4 : 8, uw = cand == 1 ? 32 : 64;
    if (uw == 32 && !narrow_ok) continue;
    if (uw == 64 && cand == 0 && narrow_ok && f32 && f16) continue;   // prefer narrow units for fp32 + fp16
    if (uw == 64 && cand == 0 && narrow_ok && g_pair_narrow >= 2 && p.K > 256) continue;
    const uint32_t st32 = f32 ? (uint32_t)(uw * 128) * (uint32_t)epi : 0u, st16 = f16 ? (uint32_t)(uw * 64) * (uint32_t)epi : 0u;
    const uint32_t staging = st32 + st16;
    if (wbytes + staging + fixed + 3u * 16384u > PAIR_SMEM_MAX) continue;
    int stages = (int)((PAIR_SMEM_MAX - wbytes - staging - fixed) / 16384u);
    if (stages > 8) stages = 8;
    c->epi_warps = epi;
    c->uw = uw;
    c->n_slabs = p.N / 256;
    if (c->n_slabs > pairs) return false;
    c->s_tiles = (int)((m_tiles + 1) / 2);
    c->group = pairs / c->n_slabs;
    if (c->group > c->s_tiles) c->group = c->s_tiles;
    c->stages = stages;
    c->a_off = wbytes;
    c->sb32_off = c->a_off + (uint32_t)stages * 16384u;
    c->sb16_off = c->sb32_off + st32;
    c->bar_off = c->sb16_off + st16;
    c->bias_off = c->bar_off + 512u;
    c->smem_bytes = c->bias_off + 1024u + 1024u;
    return true;
  }
  return false;
}

bool gemm_pair_eligible(const GemmPlan& pl, long m_tiles) {
  PairCfg c;
  return pair_config(pl, m_tiles, &c);
}

int launch_gemm_pair(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st) {
  PairCfg c;
  if (!pair_config(pl, m_tiles, &c)) { set_error("launch_gemm_pair: shape not eligible"); return DV_ERR_INVALID; }
  const EpiParams& ep = p.epi;
  if (pl.staged_rows != p.M) {
    // exact row count: the TMA engine clips the last row tile, so rows >= M are neither read nor written
    const int w32 = c.uw == 16 ? 16 : 32;
    if (ep.out32) DV_RC(tmap_encode_rows(&pl.tmO32, ep.out32, 4, p.N, p.M, (long)ep.ld32 * 4, w32, 32));
    if (ep.out16) DV_RC(tmap_encode_rows(&pl.tmO16, ep.out16, 2, p.N, p.M, (long)ep.ld16 * 2, c.uw, 32));
    if (ep.res32) DV_RC(tmap_encode_rows(&pl.tmR32, ep.res32, 4, p.N, p.M, (long)ep.ldr32 * 4, w32, 32));
    if (ep.res16) DV_RC(tmap_encode_rows(&pl.tmR16, ep.res16, 2, p.N, p.M, (long)ep.ldr16 * 2, c.uw, 32));
    pl.staged_rows = p.M;
  }
  const int grid = 2 * c.group * c.n_slabs;
  const bool f32 = ep.out32 || ep.res32, f16 = ep.out16 || ep.res16;
  static long long* d_dbg = nullptr;
  static const bool want_dbg = getenv("DV_GEMM_DBG") != nullptr;     // diagnostics: epilogue cycle counters of CTA 0
  GemmParams pd = p;
  if (want_dbg) {
    if (!d_dbg) cudaMalloc(&d_dbg, 64);
    cudaMemsetAsync(d_dbg, 0, 64, st);
    pd.dbg = d_dbg;
  }
#define DV_PAIR_LAUNCH(F32_, F16_, UW_)                                                                              \
  DV_CUDA_OK(launch_pdl(umma_gemm_pair_kernel<F32_, F16_, UW_>, dim3(grid), dim3(320), c.smem_bytes, st, pl.tmA, pl.tmB, \
                        pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, pd, c))
  if (f32 && f16) {
    if (c.uw == 16) DV_PAIR_LAUNCH(true, true, 16);
    else if (c.uw == 32) DV_PAIR_LAUNCH(true, true, 32);
    else DV_PAIR_LAUNCH(true, true, 64);
  }
  else if (f32) DV_PAIR_LAUNCH(true, false, 64);
  else { if (c.uw == 32) DV_PAIR_LAUNCH(false, true, 32); else DV_PAIR_LAUNCH(false, true, 64); }
#undef DV_PAIR_LAUNCH
  DV_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[pair dbg] M %d N %d K %d stages %d group %d epi %d uw %d | CTA0 epi warp: tiles %lld total %lld cyc: "
            "store-read wait %lld, acc wait %lld, work %lld | prologue %lld, entry->stores drained %lld\n", p.M, p.N, p.K,
            c.stages, c.group, c.epi_warps, c.uw, h[4], h[3], h[0], h[1], h[2], h[5], h[6]);
  }
  return DV_OK;
}

}  // namespace dv
