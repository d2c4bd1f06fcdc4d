// Fused LightGlue FFN block on tcgen05:   x <- x + W3 . GELU(LayerNorm512(W0 . [x | msg] + b0)) + b3
// (cvg/LightGlue TransformerLayer ffn = Linear(512,512) -> LayerNorm(512) -> GELU -> Linear(512,256), residual).
//
// Round-1 profile: as three kernels (GEMM, LayerNorm+GELU, GEMM) the block moves the [T,512] intermediate through HBM
// twice and is bandwidth/latency bound (~87 us per block at T = 32k tokens).  Here one persistent CTA per SM owns a
// 128-token row block end to end:
//   GEMM1  acc[128 x 512] fp32 = all 512 TMEM columns; per 64-wide k-block TMA brings the A tile (16 KB) and the W0
//          tile as two 256-row boxes (64 KB); two N=256 MMAs per k-step.
//   E1     each epilogue thread owns one token row: LayerNorm statistics over the 512 TMEM columns (two warps share a
//          row quarter and exchange partial sums through smem), then GELU and an fp16 write of the row into shared
//          memory in the SWIZZLE_128B K-major layout the second GEMM's A descriptor expects (overlaying the drained
//          GEMM1 ring).
//   GEMM2  acc[128 x 256] (TMEM columns 0..255 reused) from the smem-resident H and W3 tiles streamed by TMA
//          (prefetched during GEMM1 / E1).
//   E2     + b3 + residual (fp32 master, in place) and the fp16 copy that feeds the next layer's GEMMs.
#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

namespace {
constexpr int P1_A = 16384, P1_B = 65536, P1_STAGE = P1_A + P1_B;   // 80 KB
constexpr int OFF_P1 = 0;                                           // 2 stages: [0, 160K)
constexpr int OFF_H = 0;                                            // H tiles 8 x 16 KB overlay [0, 128K)
constexpr int OFF_PAR = 131072;                                     // parameter vectors live in [128K, 136K) during E1..E2
constexpr int OFF_W3 = 2 * P1_STAGE;                                // W3 ring: 2 x 32 KB at [160K, 224K)
constexpr int W3_STAGE = 32768;
constexpr int OFF_BAR = OFF_W3 + 2 * W3_STAGE;                      // 229376
constexpr int OFF_B3 = OFF_BAR + 128;                               // b3[256] fp32, loaded once (never overlaid)
constexpr int OFF_STAT = OFF_PAR + 6144;                            // [128 rows][2 halves] float2, E1 only (overlay)
constexpr int SMEM_BYTES = OFF_B3 + 1024 + 1024;                    // + alignment slack = 231552 <= 227 KB
}  // namespace

struct FfnParams {
  int T;                       // token rows
  const float *b0, *ln_g, *ln_b, *b3;
  float* x32;                  // [T,256] residual master (in/out)
  __half* x16; int ld16;       // fp16 copy of x (X2[:, 0:256], ld 512)
};

// exact (erf) GELU via Abramowitz-Stegun 7.1.26 on bare MUFU rcp / ex2 (|error| 3e-7; see lg.cu): ~17 instructions
// instead of the ~50 of libm's erff - the E1 pass of this kernel is issue-bound on exactly this.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_z = poly * t * e;
  const float half_x = 0.5f * x;
  return x >= 0.f ? fmaf(-half_x, erfc_z, x) : half_x * erfc_z;
}

__global__ void __launch_bounds__(320, 1) lg_ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX,
                                                              const __grid_constant__ CUtensorMap tmW0,
                                                              const __grid_constant__ CUtensorMap tmW3,
                                                              const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full1 = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [2]
  uint64_t* empty1 = full1 + 2;                                    // [2]
  uint64_t* full3 = empty1 + 2;                                    // [2]
  uint64_t* empty3 = full3 + 2;                                    // [2]
  uint64_t* acc1_full = empty3 + 2;
  uint64_t* h_ready = acc1_full + 1;
  uint64_t* acc2_full = h_ready + 1;
  uint64_t* h_free = acc2_full + 1;
  uint64_t* tmem_free = h_free + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_free + 1);
  float2* stat = reinterpret_cast<float2*>(smem + OFF_STAT);
  float* par = reinterpret_cast<float*>(smem + OFF_PAR);           // b0[512] | g[512] | b[512]  (E1 only)
  float* sb3 = reinterpret_cast<float*>(smem + OFF_B3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rb = (p.T + 127) >> 7;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmW0); prefetch_tmap(&tmW3);
    for (int s = 0; s < 2; ++s) { mbar_init(&full1[s], 1); mbar_init(&empty1[s], 1); mbar_init(&full3[s], 1); mbar_init(&empty3[s], 1); }
    mbar_init(acc1_full, 1); mbar_init(h_ready, 8); mbar_init(acc2_full, 1); mbar_init(h_free, 1); mbar_init(tmem_free, 8);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64) sb3[threadIdx.x - 64] = p.b3[threadIdx.x - 64];
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      int it = 0, k1 = 0, k3 = 0;
      for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x, ++it) {
        if (it > 0) mbar_wait(h_free, (it - 1) & 1);            // GEMM2 of the previous row block has consumed H
        for (int kb = 0; kb < 8; ++kb, ++k1) {
          const int s = k1 & 1;
          mbar_wait(&empty1[s], ((k1 >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&full1[s], P1_STAGE);
          uint8_t* st = smem + OFF_P1 + s * P1_STAGE;
          tma_load_2d(st, &tmX, &full1[s], kb * 64, rb * 128);
          tma_load_2d(st + P1_A, &tmW0, &full1[s], kb * 64, 0);
          tma_load_2d(st + P1_A + 32768, &tmW0, &full1[s], kb * 64, 256);
        }
        for (int kb = 0; kb < 8; ++kb, ++k3) {
          const int s = k3 & 1;
          mbar_wait(&empty3[s], ((k3 >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&full3[s], W3_STAGE);
          tma_load_2d(smem + OFF_W3 + s * W3_STAGE, &tmW3, &full3[s], kb * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, 256);
      int it = 0, k1 = 0, k3 = 0;
      for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x, ++it) {
        if (it > 0) mbar_wait(tmem_free, (it - 1) & 1);         // E2 of the previous row block has drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb, ++k1) {
          const int s = k1 & 1;
          mbar_wait(&full1[s], (k1 >> 1) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + OFF_P1 + s * P1_STAGE);
          const uint64_t da = make_desc_sw128(a_addr);
          const uint64_t db0 = make_desc_sw128(a_addr + P1_A);
          const uint64_t db1 = make_desc_sw128(a_addr + P1_A + 32768);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
            tc_mma_f16(tmem_base + 256, da + (uint64_t)(k * 2), db1 + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          }
          tc_commit(&empty1[s]);
        }
        tc_commit(acc1_full);
        mbar_wait(h_ready, it & 1);
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb, ++k3) {
          const int s = k3 & 1;
          mbar_wait(&full3[s], (k3 >> 1) & 1);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(smem_u32(smem + OFF_H + kb * 16384));
          const uint64_t db = make_desc_sw128(smem_u32(smem + OFF_W3 + s * W3_STAGE));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          tc_commit(&empty3[s]);
        }
        tc_commit(acc2_full);
        tc_commit(h_free);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;                                   // 0..255 among the epilogue threads
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x, ++it) {
      const long grow = (long)rb * 128 + row;
      const bool live = grow < p.T;
      mbar_wait(acc1_full, it & 1);
      tc_fence_after();
      // parameter vectors -> smem (the region is free: GEMM1 has drained its ring)
      for (int i = et; i < 1536; i += 256) {
        float v;
        if (i < 512) v = __ldg(p.b0 + i);
        else if (i < 1024) v = __ldg(p.ln_g + i - 512);
        else v = __ldg(p.ln_b + i - 1024);
        par[i] = v;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- E1 pass A: row statistics over this warp's 256 columns
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int col0 = half * 256 + c * 32;
        uint32_t r[32];
        tmem_ld32(tlane + col0, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = *reinterpret_cast<const float4*>(&par[col0 + g * 4]);      // broadcast
          const float v0 = __uint_as_float(r[g * 4]) + b4.x, v1 = __uint_as_float(r[g * 4 + 1]) + b4.y;
          const float v2 = __uint_as_float(r[g * 4 + 2]) + b4.z, v3 = __uint_as_float(r[g * 4 + 3]) + b4.w;
          s1 += (v0 + v1) + (v2 + v3);
          s2 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, s2))));
        }
      }
      stat[row * 2 + half] = make_float2(s1, s2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 o = stat[row * 2 + (half ^ 1)];
      const float mean = (s1 + o.x) * (1.f / 512.f);
      const float var = fmaxf((s2 + o.y) * (1.f / 512.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float nmr = -mean * rstd;
      // ---- E1 pass B: normalise, GELU, fp16 -> H (SWIZZLE_128B K-major tiles of 64 columns)
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int col0 = half * 256 + c * 32;
        uint32_t r[32];
        tmem_ld32(tlane + col0, r);
        tmem_ld_wait();
        __align__(16) __half2 hv[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int k0 = col0 + g * 4;
          const float4 b4 = *reinterpret_cast<const float4*>(&par[k0]);                 // b0, gamma, beta: broadcast
          const float4 g4 = *reinterpret_cast<const float4*>(&par[512 + k0]);
          const float4 e4 = *reinterpret_cast<const float4*>(&par[1024 + k0]);
          // ((v - mean) * rstd) * gamma + beta as two FMAs
          const float y0 = fmaf(fmaf(__uint_as_float(r[g * 4]) + b4.x, rstd, nmr), g4.x, e4.x);
          const float y1 = fmaf(fmaf(__uint_as_float(r[g * 4 + 1]) + b4.y, rstd, nmr), g4.y, e4.y);
          const float y2 = fmaf(fmaf(__uint_as_float(r[g * 4 + 2]) + b4.z, rstd, nmr), g4.z, e4.z);
          const float y3 = fmaf(fmaf(__uint_as_float(r[g * 4 + 3]) + b4.w, rstd, nmr), g4.w, e4.w);
          hv[g * 2] = __floats2half2_rn(gelu_erf(y0), gelu_erf(y1));
          hv[g * 2 + 1] = __floats2half2_rn(gelu_erf(y2), gelu_erf(y3));
        }
        uint8_t* tile = smem + OFF_H + (col0 >> 6) * 16384 + row * 128;
        const int ch0 = (col0 & 63) >> 3;                              // first 16-byte chunk of this 32-column group
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(tile + (((ch0 + g) ^ (row & 7)) << 4)) = reinterpret_cast<const uint4*>(hv)[g];
      }
      fence_proxy_async_smem();           // generic-proxy writes of H -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(h_ready);
      // ---- E2: x += acc2 + b3 ; fp32 master + fp16 copy
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < 8; c += 2) {
        const int col0 = c * 32;
        uint32_t r[32];
        tmem_ld32(tlane + col0, r);
        tmem_ld_wait();
        if (live) {
          float4* xp = reinterpret_cast<float4*>(p.x32 + grow * 256 + col0);
          __align__(16) __half2 hv[16];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 t = xp[g];
            t.x += __uint_as_float(r[g * 4 + 0]) + sb3[col0 + g * 4 + 0];
            t.y += __uint_as_float(r[g * 4 + 1]) + sb3[col0 + g * 4 + 1];
            t.z += __uint_as_float(r[g * 4 + 2]) + sb3[col0 + g * 4 + 2];
            t.w += __uint_as_float(r[g * 4 + 3]) + sb3[col0 + g * 4 + 3];
            xp[g] = t;
            hv[g * 2] = __floats2half2_rn(t.x, t.y);
            hv[g * 2 + 1] = __floats2half2_rn(t.z, t.w);
          }
          uint4* op = reinterpret_cast<uint4*>(p.x16 + grow * p.ld16 + col0);
#pragma unroll
          for (int g = 0; g < 4; ++g) op[g] = reinterpret_cast<const uint4*>(hv)[g];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(tmem_free);
      // par / stat overlay the GEMM1 ring: they are only touched between acc1_full and h_ready of the same row block,
      // while the producer cannot refill that ring before h_free (GEMM2 finished), which follows h_ready.
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int g_ffn_sms = 148;

int lg_ffn_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(lg_ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_ffn_sms, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

int plan_lg_ffn(FfnPlan* pl, const __half* X2, int T_cap, const __half* W0, const __half* W3, const float* b0,
                const float* ln_g, const float* ln_b, const float* b3, float* x32, __half* x16, int ld16) {
  pl->b0 = b0; pl->ln_g = ln_g; pl->ln_b = ln_b; pl->b3 = b3; pl->x32 = x32; pl->x16 = x16; pl->ld16 = ld16;
  pl->T_cap = T_cap;
  const uint64_t xd[2] = {512, (uint64_t)T_cap}, xs[1] = {1024};
  const uint32_t xb[2] = {64, 128};
  int rc = tmap_encode_f16(&pl->tmX, X2, 2, xd, xs, xb, true);
  if (rc) return rc;
  const uint64_t w0d[2] = {512, 512}, w0s[1] = {1024};
  const uint32_t wb[2] = {64, 256};
  rc = tmap_encode_f16(&pl->tmW0, W0, 2, w0d, w0s, wb, true);
  if (rc) return rc;
  const uint64_t w3d[2] = {512, 256};
  return tmap_encode_f16(&pl->tmW3, W3, 2, w3d, w0s, wb, true);
}

int launch_lg_ffn(const FfnPlan& pl, int T, cudaStream_t st) {
  if (T <= 0) return DV_OK;
  if (T > pl.T_cap) { set_error("launch_lg_ffn: rows exceed plan capacity"); return DV_ERR_CAPACITY; }
  FfnParams p;
  p.T = T; p.b0 = pl.b0; p.ln_g = pl.ln_g; p.ln_b = pl.ln_b; p.b3 = pl.b3; p.x32 = pl.x32; p.x16 = pl.x16; p.ld16 = pl.ld16;
  const int n_rb = (T + 127) / 128;
  const int grid = n_rb < g_ffn_sms ? n_rb : g_ffn_sms;
  lg_ffn_fused_kernel<<<grid, 320, SMEM_BYTES, st>>>(pl.tmX, pl.tmW0, pl.tmW3, p);
  DV_CUDA_OK(cudaGetLastError());
  return DV_OK;
}

}  // namespace dv
