// Persistent variant of the tcgen05 GEMM / tap-per-TMA implicit-GEMM conv (same operands, same fused epilogue as
// gemm_umma.cu): one CTA per SM loops over output tiles, so barrier init / TMEM allocation / tensor-map fetch are
// paid once per SM instead of once per 128 x BN tile, and the accumulator is double-buffered in TMEM so the epilogue
// of tile i (8 warps) overlaps the TMA + MMA main loop of tile i+1.  Round-1 launch lists showed the one-tile-per-CTA
// kernel spending 8-17 us on GEMMs whose MMA time is 1-2 us (LightGlue / MixVPR layers): fixed per-CTA cost.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

template <int BN>
struct PCfg {
  static constexpr int STAGES = (BN == 64) ? 8 : 6;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 512;                // whole bias vector (N <= 2048 fp32), loaded once per CTA
  static constexpr int SMEM_BYTES = BIAS_OFF + 8192 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;          // two accumulators
};

struct TileXY {
  int m_tile, n_tile;     // tile indices inside the (batch) problem
  int a_row, b_row;       // first A / B row of the problem (batched mode), else 0
  int m_lim, n_lim;       // rows / columns of the problem
  long out_off;           // element offset of the problem's output
  bool valid;
};
template <int BN>
__device__ __forceinline__ TileXY decode_tile(const GemmParams& p, int tile) {
  TileXY t;
  if (p.batch) {
    const int per = p.batch_m_tiles * p.n_tiles;
    const int b = tile / per;
    const int rem = tile - b * per;
    t.m_tile = rem / p.n_tiles;
    t.n_tile = rem - t.m_tile * p.n_tiles;
    const int4 d = __ldg(p.batch + b);
    t.a_row = d.x; t.b_row = d.y; t.m_lim = d.z; t.n_lim = (d.w + 3) & ~3;
    t.out_off = (long)b * p.out_bstride;
    t.valid = t.m_tile * 128 < d.z && t.n_tile * BN < d.w;
  } else {
    t.n_tile = tile % p.n_tiles;
    t.m_tile = tile / p.n_tiles;
    t.a_row = 0; t.b_row = 0; t.m_lim = p.M; t.n_lim = p.N; t.out_off = 0; t.valid = true;
  }
  return t;
}

template <int BN>
__global__ void __launch_bounds__(320, 1) umma_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const GemmParams p, const int m_tiles) {
  using C = PCfg<BN>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + C::STAGES;
  uint64_t* acc_full = empty + C::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
    fence_barrier_init();
  }
  // dbg counters (r01): broadcasting the bias with 32 shuffles per 32-column chunk cost ~1500 cycles per 128x128 tile;
  // the vector is staged in shared memory once and read back as 128-bit broadcasts.
  float* sbias = reinterpret_cast<float*>(smem + C::BIAS_OFF);
  const bool bias_smem = p.epi.bias != nullptr && p.N <= 2048 && !p.batch;
  if (bias_smem)
    for (int i = threadIdx.x; i < ((p.N + 127) & ~127); i += blockDim.x) sbias[i] = i < p.N ? p.epi.bias[i] : 0.f;
  if (warp == 1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();      // prologue above reads only weights (bias); activations / batch descriptors from here on

  if (warp == 0) {
    if (elect_one_sync()) {
      int kc = 0;                                     // running k-block counter across tiles (ring position)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileXY tx = decode_tile<BN>(p, tile);
        if (!tx.valid) continue;
        const int n_tile = tx.n_tile, m_tile = tx.m_tile;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int tw_i = m_tile % p.tiles_w;
          const int t2 = m_tile / p.tiles_w;
          const int th_i = t2 % p.tiles_h;
          img = t2 / p.tiles_h;
          w0 = tw_i << p.tw_log2;
          h0 = th_i * (128 >> p.tw_log2);
        }
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % C::STAGES;
          mbar_wait(&empty[s], ((kc / C::STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
          uint8_t* sA = smem + s * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          int kB;
          if (p.conv) {
            const int tap = kb / p.cin_blocks;
            const int cb = kb - tap * p.cin_blocks;
            const int r3 = tap / 3, s3 = tap - r3 * 3;
            tma_load_4d(sA, &tmA, &full[s], cb * 64, w0 * p.cstride + s3 - 1, h0 * p.cstride + r3 - 1, img);
            kB = tap * p.cin + cb * 64;
          } else {
            tma_load_2d(sA, &tmA, &full[s], kb * 64, tx.a_row + m_tile * 128);
            kB = kb * 64;
          }
          tma_load_2d(sB, &tmB, &full[s], kB, tx.b_row + n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, BN);
      int kc = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (!decode_tile<BN>(p, tile).valid) continue;
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % C::STAGES;
          mbar_wait(&full[s], (kc / C::STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * C::STAGE_BYTES);
          const uint64_t da = make_desc_sw128(a_addr);
          const uint64_t db = make_desc_sw128(a_addr + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[a]);
        ++it;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps
    const int q = warp & 3;                    // TMEM lane quarter
    const int half = (warp - 2) >> 2;          // this warp takes chunks half, half+2, ...
    const int row = q * 32 + lane;
    const EpiParams& ep = p.epi;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileXY tx = decode_tile<BN>(p, tile);
      if (!tx.valid) continue;
      const int a = it & 1;
      const int n_tile = tx.n_tile, m_tile = tx.m_tile;
      const int n0 = n_tile * BN;
      const int Nlim = tx.n_lim;
      long out_row = -1;
      bool writer = false;
      if (p.conv) {
        const int tw_i = m_tile % p.tiles_w;
        const int t2 = m_tile / p.tiles_w;
        const int th_i = t2 % p.tiles_h;
        const int img = t2 / p.tiles_h;
        const int TW = 1 << p.tw_log2;
        const int hl = row >> p.tw_log2, wl = row & (TW - 1);
        const int h = th_i * (128 >> p.tw_log2) + hl, w = (tw_i << p.tw_log2) + wl;
        if (ep.pool) {
          const int Ho = p.H >> 1, Wo = p.W >> 1;
          writer = !(hl & 1) && !(wl & 1) && (h >> 1) < Ho && (w >> 1) < Wo;
          out_row = ((long)img * Ho + (h >> 1)) * Wo + (w >> 1);
        } else {
          writer = h < p.H && w < p.W;
          out_row = ((long)img * p.H + h) * p.W + w;
        }
      } else {
        const long g = (long)m_tile * 128 + row;
        writer = g < tx.m_lim;
        out_row = g;
      }
      // bias of this warp's chunks is fetched BEFORE waiting for the accumulator: its (loaded-system) latency would
      // otherwise be exposed once per chunk on the epilogue's critical path
      float blv[BN / 64];
#pragma unroll
      for (int ci = 0; ci < BN / 64; ++ci) {
        const int cc = n0 + (half + 2 * ci) * 32 + lane;
        blv[ci] = (ep.bias && !bias_smem && cc < Nlim) ? __ldg(ep.bias + cc) : 0.f;
      }
      const long long te0 = p.dbg ? clock64() : 0;
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      const long long te1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int col0 = n0 + c * 32;
        if (col0 >= Nlim) break;         // warp-uniform
        uint32_t r[32];
        if (p.dbg_mode & 4) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0;
        } else {
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
        }
        const int ncols = min(32, Nlim - col0);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (bias_smem) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(&sbias[col0 + g * 4]);   // same address in all lanes
            v[g * 4] += b4.x; v[g * 4 + 1] += b4.y; v[g * 4 + 2] += b4.z; v[g * 4 + 3] += b4.w;
          }
        } else if (ep.bias && !(p.dbg_mode & 2)) {
          // lane j holds bias[col0 + j] (prefetched above); broadcast to every row-owning lane
          const float bl = blv[(c - half) >> 1];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
        }
        if (ep.res32 && writer && !ep.pool) {
          const float4* rp = reinterpret_cast<const float4*>(ep.res32 + out_row * ep.ldr32 + col0);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g * 4 + 4 <= ncols) {
              const float4 t = rp[g];   // plain load: may alias out32 (in-place residual)
              v[g * 4 + 0] += t.x; v[g * 4 + 1] += t.y; v[g * 4 + 2] += t.z; v[g * 4 + 3] += t.w;
            }
        }
        if (ep.res16 && writer && !ep.pool) {
          const uint4* rp = reinterpret_cast<const uint4*>(ep.res16 + out_row * ep.ldr16 + col0);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (g * 8 + 8 <= ncols) {
              const uint4 t = __ldg(rp + g);
              const __half2* h2 = reinterpret_cast<const __half2*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                v[g * 8 + 2 * e] += f.x; v[g * 8 + 2 * e + 1] += f.y;
              }
            }
        }
        if (ep.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ep.rope_cs && col0 < ep.rope_cols && writer) {      // warp-uniform except `writer`
          const int j0 = (col0 & 63) >> 1;
          const float4* cp = reinterpret_cast<const float4*>(ep.rope_cs + out_row * 32 + j0);
          const float4* sp = reinterpret_cast<const float4*>(ep.rope_sn + out_row * 32 + j0);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 c4 = __ldg(cp + g), s4 = __ldg(sp + g);
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = g * 4 + e;
              const float x0 = v[2 * jj], x1 = v[2 * jj + 1];
              v[2 * jj] = x0 * cc[e] - x1 * ss[e];
              v[2 * jj + 1] = x1 * cc[e] + x0 * ss[e];
            }
          }
        }
        if (ep.out32 && writer && !ep.pool && !(p.dbg_mode & 1)) {
          float4* op = reinterpret_cast<float4*>(ep.out32 + tx.out_off + out_row * ep.ld32 + col0);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g * 4 + 4 <= ncols) op[g] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
        if (ep.out16) {
          __align__(16) __half2 hv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) hv[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          if (ep.pool) {
            const int TW = 1 << p.tw_log2;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint32_t u = *reinterpret_cast<uint32_t*>(&hv[j]);
              uint32_t o = __shfl_xor_sync(0xffffffffu, u, 1);
              __half2 m = __hmax2(*reinterpret_cast<__half2*>(&u), *reinterpret_cast<__half2*>(&o));
              u = *reinterpret_cast<uint32_t*>(&m);
              o = __shfl_xor_sync(0xffffffffu, u, TW);
              hv[j] = __hmax2(m, *reinterpret_cast<__half2*>(&o));
            }
          }
          if (writer) {
            const uint4* src = reinterpret_cast<const uint4*>(hv);
            if (ep.blocked_hw) {
              const long img = out_row / ep.blocked_hw, pix = out_row - img * ep.blocked_hw;
              const int ngrp = p.N >> 3;
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (g * 8 + 8 <= ncols)
                  *reinterpret_cast<uint4*>(ep.out16 + ((img * ngrp + (col0 >> 3) + g) * ep.blocked_hw + pix) * 8) = src[g];
            } else {
              uint4* op = reinterpret_cast<uint4*>(ep.out16 + out_row * ep.ld16 + col0);
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (g * 8 + 8 <= ncols) op[g] = src[g];
            }
          }
        }
      }
      // all of this warp's TMEM reads of buffer `a` are complete (tmem_ld_wait above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(&acc_empty[a]);
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[3] += te1 - te0; p.dbg[4] += clock64() - te1; p.dbg[5] += 1; }
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

static int g_sms = 148;

int gemm_persistent_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_persist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PCfg<64>::SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_persist_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PCfg<128>::SMEM_BYTES));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

int launch_gemm_persistent(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st) {
  const long total = m_tiles * p.n_tiles;
  const int grid = (int)(total < g_sms ? total : g_sms);
  static long long* d_dbg = nullptr;
  static const bool want_dbg = getenv("DV_GEMM_DBG") != nullptr;     // diagnostics: epilogue cycle counters of CTA 0
  GemmParams pp = p;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (want_dbg) {
    if (!d_dbg) cudaMalloc(&d_dbg, 64);
    cudaMemsetAsync(d_dbg, 0, 64, st);
    pp.dbg = d_dbg;
    pp.dbg_mode = atoi(getenv("DV_GEMM_DBG"));
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st);
  }
  if (pl.bn == 64)
    DV_CUDA_OK(launch_pdl(umma_gemm_persist_kernel<64>, dim3(grid), dim3(320), PCfg<64>::SMEM_BYTES, st, pl.tmA, pl.tmB, pp,
                          (int)m_tiles));
  else
    DV_CUDA_OK(launch_pdl(umma_gemm_persist_kernel<128>, dim3(grid), dim3(320), PCfg<128>::SMEM_BYTES, st, pl.tmA, pl.tmB, pp,
                          (int)m_tiles));
  DV_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    long long h[8];
    float ms = 0.f;
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gemm dbg] m_tiles %ld n_tiles %d kb %d bn %d grid %d: %.1f us | CTA0 epilogue: wait %lld work %lld tiles %lld\n",
            m_tiles, p.n_tiles, p.num_kb, pl.bn, grid, ms * 1e3, h[3], h[4], h[5]);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return DV_OK;
}

int launch_gemm_batched(const GemmPlan& pl, const int4* desc, int count, int max_m, int max_n, long out_bstride,
                        cudaStream_t st) {
  if (count <= 0) return DV_OK;
  GemmParams p = pl.p;
  if (p.conv || !p.epi.out32 || p.epi.out16 || p.epi.res32 || p.epi.res16) {
    set_error("launch_gemm_batched: plain fp32-output plans only");
    return DV_ERR_INVALID;
  }
  p.batch = desc; p.batch_count = count; p.out_bstride = out_bstride;
  p.batch_m_tiles = cdiv(max_m, 128);
  p.n_tiles = cdiv(max_n, pl.bn);
  return launch_gemm_persistent(pl, p, (long)count * p.batch_m_tiles, st);
}

}  // namespace dv
