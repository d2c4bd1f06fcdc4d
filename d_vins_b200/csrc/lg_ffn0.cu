// LightGlue FFN, first half, in ONE kernel:   g = GELU(LayerNorm512(W0' . [x | ctx] + b0'))      (fp16 in, fp16 out)
// (cvg/LightGlue TransformerLayer.ffn[0..2]; W0' carries the folded out-projection, lg.cu).  Replaces the K = 512 GEMM
// (umma_gemm_pair_kernel, 34 us per 53 k tokens) + k_lg_ln_gelu (32 us): the [T,512] fp16 pre-LayerNorm activation is
// no longer written to and re-read from HBM, and one launch per transformer block disappears.
//
// A cluster of FOUR CTAs (two cta_group::2 pairs on two TPCs) owns 256 token rows x all 512 hidden columns:
//   pair s = rank >> 1 keeps the 256-column slab s of W0' resident (each CTA half of it: 128 columns x 512 = 128 KB) and
//   issues M = 256 / N = 256 MMAs exactly as gemm_pair.cu does; CTA rank & 1 of either pair holds rows [128 (rank & 1), +128).
// LayerNorm needs the statistics of a whole 512-column row, of which a CTA has 256 columns in TMEM: every epilogue warp
// sweeps its part of the accumulator once for (sum, sum of squares) - of the values ROUNDED TO FP16, i.e. exactly the
// numbers the unfused path stored and normalised - the two warps of a row combine through shared memory, and the CTA
// sends its 128 row partials to the CTA with the same rows in the other pair (rank ^ 2) through distributed shared
// memory (st.async: remote store that completes on the receiver's mbarrier; double-buffered by tile parity).  A second sweep normalises,
// applies the exact (erf) GELU and stages fp16 for the TMA store.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

namespace {
constexpr int F0_K = 512, F0_N = 512, F0_KB = F0_K / 64;
constexpr int F0_STAGES = 3;
constexpr uint32_t F0_W = 0;                                   // 8 k-blocks x [128 x 64] fp16 = 128 KB
constexpr uint32_t F0_A = 131072;                              // 3 x 16 KB
constexpr uint32_t F0_ST = F0_A + F0_STAGES * 16384;           // staging: 16 warps x 2 KB (32-column fp16 boxes)
constexpr uint32_t F0_BAR = F0_ST + 16 * 2048;                  // barriers (256 B)
constexpr uint32_t F0_PAR = F0_BAR + 256;                      // bias | gamma | beta of this CTA's 256 columns (3 KB)
constexpr uint32_t F0_PART = F0_PAR + 3072;                    // [2 parities][4 quarters][128 rows] float2: intra-CTA row partials
constexpr uint32_t F0_XBUF = F0_PART + 8192;                   // [2 parities][128 rows] float2: partner CTA's partials
constexpr uint32_t F0_SMEM = F0_XBUF + 2048 + 1024;            // + alignment slack = 227 584
}  // namespace

struct Ffn0Params {
  int M;                          // token rows
  const float *bias, *gamma, *beta;
  long long* dbg;                 // DV_FFN0_DBG: cycle counters of one epilogue thread (diagnostics only)
};

__device__ __forceinline__ float ffn0_gelu_erf(float x) {      // the same Abramowitz-Stegun erf as k_lg_ln_gelu (lg.cu)
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
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(576, 1)
    lg_ffn0_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO16, const Ffn0Params p, int s_tiles, int n_clusters) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + F0_BAR);       // [3]
  uint64_t* empty = full + F0_STAGES;                                // [3]
  uint64_t* acc_full = empty + F0_STAGES;                            // [2]
  uint64_t* acc_empty = acc_full + 2;                                // [2]
  uint64_t* w_bar = acc_empty + 2;
  uint64_t* xbar = w_bar + 1;                                        // [2 parities][4 quarters]: partner partials landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xbar + 8);
  float* spar = reinterpret_cast<float*>(smem + F0_PAR);             // [0,256) bias, [256,512) gamma, [512,768) beta
  float2* part = reinterpret_cast<float2*>(smem + F0_PART);
  float2* xbuf = reinterpret_cast<float2*>(smem + F0_XBUF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();         // 0..3
  const uint32_t leader = rank & ~1u;              // MMA-issuing CTA of this pair
  const uint32_t rhalf = rank & 1u;                // row half of the 256-row super-tile
  const int slab = (int)(rank >> 1);               // hidden columns [256 slab, +256)
  const uint16_t pair_mask = (uint16_t)(3u << leader);
  const int cluster = (int)blockIdx.x >> 2;
  const int col_base = slab * 256;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmO16);
    for (int s = 0; s < F0_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 32); }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&xbar[i], 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    spar[i] = p.bias[col_base + i];
    spar[256 + i] = p.gamma[col_base + i];
    spar[512 + i] = p.beta[col_base + i];
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // every CTA's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t wbar_leader = mapa_u32(smem_u32(w_bar), leader);
      if (rhalf == 0) mbar_arrive_expect_tx(w_bar, (uint32_t)F0_KB * 32768u);
      for (int kb = 0; kb < F0_KB; ++kb)
        tma_load_2d_pair(smem + F0_W + (uint32_t)kb * 16384u, &tmB, wbar_leader, kb * 64, col_base + (int)rhalf * 128);
      pdl_wait();
      int kc = 0;
      for (int st = cluster; st < s_tiles; st += n_clusters) {
        for (int kb = 0; kb < F0_KB; ++kb, ++kc) {
          const int s = kc % F0_STAGES;
          mbar_wait(&empty[s], ((kc / F0_STAGES) & 1) ^ 1);
          if (rhalf == 0) mbar_arrive_expect_tx(&full[s], 32768u);
          tma_load_2d_pair(smem + F0_A + (uint32_t)s * 16384u, &tmA, mapa_u32(smem_u32(&full[s]), leader), kb * 64,
                           st * 256 + (int)rhalf * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (rhalf == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(256, 256);
      mbar_wait(w_bar, 0);
      int kc = 0, it = 0;
      for (int st = cluster; st < s_tiles; st += n_clusters, ++it) {
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * 256);
        for (int kb = 0; kb < F0_KB; ++kb, ++kc) {
          const int s = kc % F0_STAGES;
          mbar_wait(&full[s], (kc / F0_STAGES) & 1);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(smem_u32(smem + F0_A + (uint32_t)s * 16384u));
          const uint64_t db = make_desc_sw128(smem_u32(smem + F0_W + (uint32_t)kb * 16384u));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          tc_commit_pair(&empty[s], pair_mask);
        }
        tc_commit_pair(&acc_full[a], pair_mask);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 16 warps, two sweeps over the accumulator
    // (the sweeps are instruction-issue bound - ~33 instructions per element with the exact GELU - so four warps per
    // scheduler instead of two; warp (q = warp % 4, hq = (warp - 2) / 4) owns TMEM lanes [32q, +32) x columns [64 hq, +64))
    const int e = warp - 2;
    const int q = warp & 3;
    const int hq = e >> 2;
    const int row = q * 32 + lane;
    uint8_t* b16 = smem + F0_ST + e * 2048;        // one 32-column fp16 box (64-byte rows, SWIZZLE_64B)
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), leader);
    const uint32_t partner = rank ^ 2u;            // the CTA with the same rows in the other pair
    const uint32_t xbuf_partner = mapa_u32(smem_u32(xbuf), partner);
    const uint32_t xbar_partner = mapa_u32(smem_u32(xbar), partner);
    const int bar_id = 1 + q;
    const bool dbgt = p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0;
    long long d_acc = 0, d_s1 = 0, d_x = 0, d_s2 = 0;
    int it = 0;
    for (int st = cluster; st < s_tiles; st += n_clusters, ++it) {
      const int a = it & 1, par = it & 1;
      const long long t0 = dbgt ? clock64() : 0;
      const int row0 = st * 256 + (int)rhalf * 128 + q * 32;     // first row of this warp
      const bool active = row0 < p.M;                            // warp-uniform (partials are exchanged regardless)
      // this quarter's 32 partner partials (8 bytes each) arrive as asynchronous remote stores that complete on xbar
      if (hq == 0 && lane == 0) mbar_arrive_expect_tx(&xbar[par * 4 + q], 256u);
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      const long long t1 = dbgt ? clock64() : 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + hq * 64);
      // ---- sweep 1: row sum / sum of squares of (acc + bias) rounded to fp16
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        uint32_t r[32];
        tmem_ld32(taddr + ci * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = *reinterpret_cast<const float4*>(&spar[hq * 64 + ci * 32 + g * 4]);   // broadcast
          const float2 f0 = __half22float2(__floats2half2_rn(__uint_as_float(r[g * 4 + 0]) + b4.x, __uint_as_float(r[g * 4 + 1]) + b4.y));
          const float2 f1 = __half22float2(__floats2half2_rn(__uint_as_float(r[g * 4 + 2]) + b4.z, __uint_as_float(r[g * 4 + 3]) + b4.w));
          s1 += (f0.x + f0.y) + (f1.x + f1.y);
          s2 = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, s2))));
        }
      }
      const long long t2 = dbgt ? clock64() : 0;
      float2* pt = part + par * 512;                               // [4 column quarters][128 rows], double-buffered by tile
      pt[hq * 128 + row] = make_float2(s1, s2);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // the four warps of this lane quarter
      const float2 p0 = pt[row], p1 = pt[128 + row], p2 = pt[256 + row], p3 = pt[384 + row];
      const float c1 = (p0.x + p1.x) + (p2.x + p3.x), c2 = (p0.y + p1.y) + (p2.y + p3.y);   // this CTA's 256 columns
      if (hq == 0) {
        const uint32_t dst = xbuf_partner + (uint32_t)((par * 128 + row) * 8);
        const uint32_t bar = xbar_partner + (uint32_t)((par * 4 + q) * 8);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(dst),
                     "f"(c1), "f"(c2), "r"(bar)
                     : "memory");
      }
      mbar_wait(&xbar[par * 4 + q], (it >> 1) & 1);
      const float2 px = xbuf[par * 128 + row];
      // the partner adds in the other order; fp32 addition is commutative, so both CTAs normalise with identical statistics
      const float mean = (c1 + px.x) * (1.f / 512.f);
      const float var = fmaxf((c2 + px.y) * (1.f / 512.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float nmr = -mean * rstd;
      const long long t3 = dbgt ? clock64() : 0;
      // ---- sweep 2: normalise, GELU, fp16 -> staging -> TMA store, 32 columns at a time
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const int lcol = hq * 64 + u * 32;                       // first column of this unit inside the CTA's 256
        if (lane == 0) bulk_wait_read0();                        // the previous TMA store finished READING the staging box
        __syncwarp();
        uint32_t r[32];
        tmem_ld32(taddr + u * 32, r);
        tmem_ld_wait();
        if (u == 1) {                                            // last TMEM read of this accumulator buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)a * 8u);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __align__(16) __half2 hv[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int cc = lcol + g * 8 + 2 * x;
            const float2 bb = *reinterpret_cast<const float2*>(&spar[cc]);
            const float2 gg = *reinterpret_cast<const float2*>(&spar[256 + cc]);
            const float2 be = *reinterpret_cast<const float2*>(&spar[512 + cc]);
            const float2 f = __half22float2(__floats2half2_rn(__uint_as_float(r[g * 8 + 2 * x]) + bb.x,
                                                              __uint_as_float(r[g * 8 + 2 * x + 1]) + bb.y));
            const float y0 = ffn0_gelu_erf(fmaf(fmaf(f.x, rstd, nmr), gg.x, be.x));
            const float y1 = ffn0_gelu_erf(fmaf(fmaf(f.y, rstd, nmr), gg.y, be.y));
            hv[x] = __floats2half2_rn(y0, y1);
          }
          // 64-byte rows, SWIZZLE_64B: 16-byte chunk ^= address bits [7:8] = (row >> 1) & 3
          *reinterpret_cast<uint4*>(b16 + (uint32_t)lane * 64u + ((((uint32_t)g) ^ (((uint32_t)lane >> 1) & 3u)) << 4)) =
              *reinterpret_cast<const uint4*>(hv);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && active) {
          tma_store_2d(&tmO16, b16, col_base + lcol, row0);
          bulk_commit();
        }
      }
      if (dbgt) { const long long t4 = clock64(); d_acc += t1 - t0; d_s1 += t2 - t1; d_x += t3 - t2; d_s2 += t4 - t3; }
    }
    if (dbgt) { p.dbg[0] = d_acc; p.dbg[1] = d_s1; p.dbg[2] = d_x; p.dbg[3] = d_s2; p.dbg[4] = it; }
    if (lane == 0) bulk_wait0();
  }
  // no CTA may exit (or free TMEM) while another can still signal its barriers / write its shared memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512u);
  }
}

static int g_ffn0_clusters = 32;     // co-resident clusters of four (GPCs whose SM count is not a multiple of 4 strand SMs)

int lg_ffn0_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(lg_ffn0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F0_SMEM));
  int dev = 0, sms = 148;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * (sms / 4)); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = F0_SMEM;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lg_ffn0_kernel, &cfg) == cudaSuccess && n > 0) g_ffn0_clusters = n < sms / 4 ? n : sms / 4;
  else { cudaGetLastError(); g_ffn0_clusters = sms / 4 > 4 ? sms / 4 - 4 : 1; }
  return DV_OK;
}
int lg_ffn0_clusters() { return g_ffn0_clusters; }

// A = X2 [T_cap, 512] fp16 (pitch lda), W [512, 512] fp16 row-major (K-major), out [T_cap, 512] fp16 (pitch ldo)
int plan_lg_ffn0(Ffn0Plan* pl, const __half* A, int lda, int T_cap, const __half* W, const float* bias, const float* gamma,
                 const float* beta, __half* out, int ldo) {
  pl->bias = bias; pl->gamma = gamma; pl->beta = beta; pl->out = out; pl->ldo = ldo; pl->rows_cap = T_cap;
  pl->out_rows = -1;
  const uint64_t ad[2] = {(uint64_t)F0_K, (uint64_t)T_cap}, as[1] = {(uint64_t)lda * 2};
  const uint32_t box[2] = {64, 128};
  int rc = tmap_encode_f16(&pl->tmA, A, 2, ad, as, box, true);
  if (rc) return rc;
  const uint64_t wd[2] = {(uint64_t)F0_K, (uint64_t)F0_N}, ws[1] = {(uint64_t)F0_K * 2};
  return tmap_encode_f16(&pl->tmB, W, 2, wd, ws, box, true);
}

int launch_lg_ffn0(const Ffn0Plan& pl, int rows, cudaStream_t st) {
  if (rows <= 0) return DV_OK;
  if (rows > pl.rows_cap) { set_error("launch_lg_ffn0: rows exceed plan capacity"); return DV_ERR_CAPACITY; }
  if (pl.out_rows != rows) {     // exact row count: the TMA engine clips the last row tile
    int rc = tmap_encode_rows(&pl.tmO16, pl.out, 2, F0_N, rows, (long)pl.ldo * 2, 32, 32);
    if (rc) return rc;
    pl.out_rows = rows;
  }
  const int s_tiles = (rows + 255) / 256;
  int n_clusters = g_ffn0_clusters;
  if (n_clusters > s_tiles) n_clusters = s_tiles;
  Ffn0Params p;
  p.M = rows; p.bias = pl.bias; p.gamma = pl.gamma; p.beta = pl.beta;
  static long long* d_dbg = nullptr;
  static const bool want_dbg = getenv("DV_FFN0_DBG") != nullptr;
  if (want_dbg && !d_dbg) { cudaMalloc(&d_dbg, 64); cudaMemset(d_dbg, 0, 64); }
  p.dbg = want_dbg ? d_dbg : nullptr;
  DV_CUDA_OK(launch_pdl(lg_ffn0_kernel, dim3(4 * n_clusters), dim3(576), (size_t)F0_SMEM, st, pl.tmA, pl.tmB, pl.tmO16, p,
                        s_tiles, n_clusters));
  DV_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[ffn0 dbg] rows %d clusters %d | CTA 0 warp 2: tiles %lld, cycles per tile: accumulator wait %lld, sweep 1 %lld, "
            "combine + DSMEM exchange %lld, sweep 2 %lld\n", rows, n_clusters, h[4], h[0] / (h[4] ? h[4] : 1), h[1] / (h[4] ? h[4] : 1),
            h[2] / (h[4] ? h[4] : 1), h[3] / (h[4] ? h[4] : 1));
  }
  return DV_OK;
}

}  // namespace dv
