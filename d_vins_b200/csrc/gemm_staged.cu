// Persistent tcgen05 GEMM with a shared-memory staged epilogue: C[M,N] = A[M,K] * W[N,K]^T (+bias, +residual, ReLU,
// rotary), plain row-major operands, 128 x 128 tiles.
//
// Why a second epilogue: the register epilogue of gemm_persistent.cu stores (and reads the residual) with one thread
// per output ROW, i.e. every warp-wide access touches 32 different 128-byte lines.  Its cycle counters
// (profiles/README.md, DV_GEMM_DBG) put 2/3 of the tile time into those stores; the LightGlue / MixVPR GEMMs with
// K <= 512 ran at 170-490 TFLOP/s because of it.  Here the epilogue warps only move TMEM -> registers -> swizzled
// shared memory; the global side is done by the TMA engine in both directions:
//   * residual tiles (fp32 and/or fp16) are fetched by TMA into the warp's staging boxes BEFORE the warp waits for
//     the accumulator, so their latency hides behind the tile's MMAs;
//   * results are written back to the same boxes and leave with cp.async.bulk.tensor stores (full-line writes, clipped
//     at the matrix edge by the tensor map, asynchronous to the warp).
// Staging boxes are [32 rows x 128 bytes] with the 128-byte swizzle, so the row-per-lane shared-memory accesses are
// conflict-free (8 consecutive rows cover all 32 banks).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue; warp (q = warp % 4,
// h = (warp - 2) / 4) owns TMEM lanes [32q, 32q + 32) x accumulator columns [64h, 64h + 64).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

#define DV_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

template <bool F32, bool F16>
struct SCfg {
  static constexpr int STAGES = (F32 && F16) ? 3 : (F32 ? 4 : 5);
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = 128 * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SB32_OFF = STAGES * STAGE_BYTES;              // 8 warps x 2 boxes x 4 KB
  static constexpr int SB16_OFF = SB32_OFF + (F32 ? 65536 : 0);      // 8 warps x 1 box x 4 KB
  static constexpr int BAR_OFF = SB16_OFF + (F16 ? 32768 : 0);
  static constexpr int BIAS_OFF = BAR_OFF + 512;                     // bias vector, N <= 1024
  static constexpr int SMEM_BYTES = BIAS_OFF + 4096 + 1024;
  static constexpr int TMEM_COLS = 256;                              // two 128-column accumulators
};

template <bool F32, bool F16>
__global__ void __launch_bounds__(320, 1)
    umma_gemm_staged_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                            const __grid_constant__ CUtensorMap tmR32, const __grid_constant__ CUtensorMap tmR16,
                            const GemmParams p, const int m_tiles) {
  using C = SCfg<F32, F16>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + C::STAGES;
  uint64_t* acc_full = empty + C::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* rbar = acc_empty + 2;                  // one per epilogue warp: residual boxes landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rbar + 8);
  float* sbias = reinterpret_cast<float*>(smem + C::BIAS_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = m_tiles * p.n_tiles;
  const EpiParams& ep = p.epi;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (ep.out32) prefetch_tmap(&tmO32);
    if (ep.out16) prefetch_tmap(&tmO16);
    if (ep.res32) prefetch_tmap(&tmR32);
    if (ep.res16) prefetch_tmap(&tmR16);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
    for (int e = 0; e < 8; ++e) mbar_init(&rbar[e], 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < ((p.N + 127) & ~127); i += blockDim.x) sbias[i] = (ep.bias && i < p.N) ? ep.bias[i] : 0.f;
  if (warp == 1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();      // everything above touched only weights / on-chip state; operands of earlier kernels from here on

  if (warp == 0) {
    if (elect_one_sync()) {
      int kc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % C::STAGES;
          mbar_wait(&empty[s], ((kc / C::STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
          uint8_t* sA = smem + s * C::STAGE_BYTES;
          tma_load_2d(sA, &tmA, &full[s], kb * 64, m_tile * 128);
          tma_load_2d(sA + C::A_BYTES, &tmB, &full[s], kb * 64, n_tile * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, 128);
      int kc = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * 128);
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
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps
    const int e = warp - 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int h = e >> 2;                      // column half
    uint8_t* b32 = smem + C::SB32_OFF + e * 8192;
    uint8_t* b16 = smem + C::SB16_OFF + e * 4096;
    uint64_t* rb = &rbar[e];
    const uint32_t rowoff = (uint32_t)lane * 128u;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const bool has_res = (F32 && ep.res32) || (F16 && ep.res16);
    uint32_t rphase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
      const int colw = n_tile * 128 + h * 64;            // first column of this warp
      const int row0 = m_tile * 128 + q * 32;            // first row of this warp
      const bool active = colw < p.N && row0 < p.M;      // warp-uniform
      const bool writer = row0 + lane < p.M;
      // (1) the previous tile's TMA stores must have finished READING the staging boxes
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      // (2) residual boxes: in flight while the tile's MMAs run
      if (active && has_res && lane == 0) {
        uint32_t bytes = 0;
        if (F32 && ep.res32) bytes += 4096u + (colw + 32 < p.N ? 4096u : 0u);
        if (F16 && ep.res16) bytes += 4096u;
        mbar_arrive_expect_tx(rb, bytes);
        if (F32 && ep.res32) {
          tma_load_2d(b32, &tmR32, rb, colw, row0);
          if (colw + 32 < p.N) tma_load_2d(b32 + 4096, &tmR32, rb, colw + 32, row0);
        }
        if (F16 && ep.res16) tma_load_2d(b16, &tmR16, rb, colw, row0);
      }
      // rotary tables of this row: 32 cos + 32 sin, one head (64 columns) per warp; fetched before the wait
      const bool rope = ep.rope_cs && colw < ep.rope_cols && active;
      float4 cs4[8], sn4[8];
      if (rope && writer) {
        const float4* cp = reinterpret_cast<const float4*>(ep.rope_cs + (long)(row0 + lane) * 32);
        const float4* sp = reinterpret_cast<const float4*>(ep.rope_sn + (long)(row0 + lane) * 32);
#pragma unroll
        for (int g = 0; g < 8; ++g) { cs4[g] = __ldg(cp + g); sn4[g] = __ldg(sp + g); }
      } else {
#pragma unroll
        for (int g = 0; g < 8; ++g) { cs4[g] = make_float4(1.f, 1.f, 1.f, 1.f); sn4[g] = make_float4(0.f, 0.f, 0.f, 0.f); }
      }
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      if (active) {
        if (has_res) { mbar_wait(rb, rphase); rphase ^= 1u; }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 128 + h * 64);
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int col0 = colw + ci * 32;
          if (col0 < p.N) {                              // warp-uniform
            uint32_t r[32];
            tmem_ld32(taddr + ci * 32, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b4 = *reinterpret_cast<const float4*>(&sbias[col0 + g * 4]);   // broadcast
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
                const uint4 t = *reinterpret_cast<const uint4*>(b16 + rowoff + (((uint32_t)(ci * 4 + g) << 4) ^ swz));
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
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 c4 = cs4[ci * 4 + g], s4 = sn4[ci * 4 + g];
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int jj = g * 4 + x;
                  const float x0 = v[2 * jj], x1 = v[2 * jj + 1];
                  v[2 * jj] = x0 * cc[x] - x1 * ss[x];
                  v[2 * jj + 1] = x1 * cc[x] + x0 * ss[x];
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
                *reinterpret_cast<uint4*>(b16 + rowoff + (((uint32_t)(ci * 4 + g) << 4) ^ swz)) =
                    *reinterpret_cast<const uint4*>(hv);
              }
            }
          }
        }
      }
      // TMEM reads of buffer `a` are complete (tmem_ld_wait); staging writes become visible to the async proxy
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cnt(&acc_empty[a]);
        if (active) {
          if (F32 && ep.out32) {
            tma_store_2d(&tmO32, b32, colw, row0);
            if (colw + 32 < p.N) tma_store_2d(&tmO32, b32 + 4096, colw + 32, row0);
          }
          if (F16 && ep.out16) tma_store_2d(&tmO16, b16, colw, row0);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

static int g_sms_staged = 148;

int gemm_staged_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_staged_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SCfg<true, true>::SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_staged_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SCfg<true, false>::SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_staged_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SCfg<false, true>::SMEM_BYTES));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_sms_staged, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

// plain row-major GEMM, 128-column tiles, every output / residual 16-byte aligned with a 16-byte-multiple pitch
bool gemm_staged_eligible(const GemmPlan& pl) {
  const GemmParams& p = pl.p;
  const EpiParams& ep = p.epi;
  if (p.conv || pl.bn != 128 || ep.pool || ep.blocked_hw || p.batch) return false;
  if ((p.N & 7) || p.N > 1024) return false;
  if (!ep.out16 && !ep.out32) return false;
  auto ok = [](const void* ptr, long ld, int es) {
    return !ptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((ld * es) & 15) == 0);
  };
  if (!ok(ep.out16, ep.ld16, 2) || !ok(ep.out32, ep.ld32, 4) || !ok(ep.res16, ep.ldr16, 2) ||
      !ok(ep.res32, ep.ldr32, 4))
    return false;
  if (ep.rope_cs && (ep.rope_cols & 63)) return false;
  return true;
}

int launch_gemm_staged(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st) {
  const EpiParams& ep = p.epi;
  if (pl.staged_rows != p.M) {
    // exact row count: the TMA engine clips the last row tile, so rows >= M are neither read nor written
    if (ep.out32) DV_RC(tmap_encode_rows(&pl.tmO32, ep.out32, 4, p.N, p.M, (long)ep.ld32 * 4, 32, 32));
    if (ep.out16) DV_RC(tmap_encode_rows(&pl.tmO16, ep.out16, 2, p.N, p.M, (long)ep.ld16 * 2, 64, 32));
    if (ep.res32) DV_RC(tmap_encode_rows(&pl.tmR32, ep.res32, 4, p.N, p.M, (long)ep.ldr32 * 4, 32, 32));
    if (ep.res16) DV_RC(tmap_encode_rows(&pl.tmR16, ep.res16, 2, p.N, p.M, (long)ep.ldr16 * 2, 64, 32));
    pl.staged_rows = p.M;
  }
  const long total = m_tiles * p.n_tiles;
  const int grid = (int)(total < g_sms_staged ? total : g_sms_staged);
  const bool f32 = ep.out32 || ep.res32, f16 = ep.out16 || ep.res16;
  if (f32 && f16)
    DV_CUDA_OK(launch_pdl(umma_gemm_staged_kernel<true, true>, dim3(grid), dim3(320), SCfg<true, true>::SMEM_BYTES, st,
                          pl.tmA, pl.tmB, pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, p, (int)m_tiles));
  else if (f32)
    DV_CUDA_OK(launch_pdl(umma_gemm_staged_kernel<true, false>, dim3(grid), dim3(320), SCfg<true, false>::SMEM_BYTES, st,
                          pl.tmA, pl.tmB, pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, p, (int)m_tiles));
  else
    DV_CUDA_OK(launch_pdl(umma_gemm_staged_kernel<false, true>, dim3(grid), dim3(320), SCfg<false, true>::SMEM_BYTES, st,
                          pl.tmA, pl.tmB, pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, p, (int)m_tiles));
  DV_CUDA_OK(cudaGetLastError());
  return DV_OK;
}

}  // namespace dv
