// Weights-resident persistent tcgen05 GEMM: C[M,N] = A[M,K] * W[N,K]^T (+bias, +residual, ReLU, rotary) for the
// small-K linear layers of LightGlue / MixVPR (K <= 512).
//
// Why: ncu on the streaming kernel (gemm_staged.cu, 128 x 128 tiles) shows l1tex__m_xbar2l1tex_read_bytes = 218 MB for
// the 26112 x 512 x 512 FFN GEMM - every tile re-fetches 128 KB of A and 128 KB of W through the L2 -> SM fabric for
// 2048 cycles of MMA (128 B/clk/SM against a measured fabric ceiling of ~40 B/clk/SM), so the tensor pipe sits at
// 20-40 %.  Here a CTA keeps a SLAB of the weight matrix (SLAB = 256 or 128 output columns x all of K, <= 128 KB)
// resident in shared memory for its whole lifetime and streams only A tiles: 128 x K x 2 bytes per 128 x SLAB x K MACs
// = 32 B/clk/SM at SLAB = 256 (16 KB per 512 MMA cycles), 64 B/clk/SM at SLAB = 128.  The same idea as the
// weights-stationary halo convolution (conv_halo.cu), for plain row-major operands.
//
// CTA c owns slab (c % n_slabs) and walks the row tiles (c / n_slabs) + i * group.  Accumulators: two buffers of SLAB
// TMEM columns (512 columns at SLAB = 256, the whole TMEM - one CTA per SM).  Epilogue: the staged one of
// gemm_staged.cu (TMEM -> registers -> swizzled smem boxes -> TMA store; residuals prefetched by TMA), run once per
// 128-column sub-tile of the slab.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue; warp (q = warp % 4,
// h = (warp - 2) / 4) owns TMEM lanes [32q, 32q + 32) x columns [64h, 64h + 64) of every 128-column sub-tile.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

#define DV_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

struct WresCfg {
  int m_tiles, slab, n_slabs, group, stages;
  int epi_warps;     // 8: warp (q, h) drains column half h of every 128-column sub-tile; 4: warp q drains both halves in
                     // turn through one set of staging boxes (fp32 + fp16 outputs with a 128 KB slab: 48 KB of staging)
  uint32_t a_off, sb32_off, sb16_off, bar_off, bias_off, smem_bytes;
};

template <bool F32, bool F16>
__global__ void __launch_bounds__(320, 1)
    umma_gemm_wres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                          const __grid_constant__ CUtensorMap tmR32, const __grid_constant__ CUtensorMap tmR16,
                          const GemmParams p, const WresCfg c) {
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
  const bool live = (int)blockIdx.x < c.group * c.n_slabs;
  const int slab_idx = (int)blockIdx.x % c.n_slabs;
  const int j0 = live ? (int)blockIdx.x / c.n_slabs : c.m_tiles;
  const int col_base = slab_idx * c.slab;          // first output column of this CTA
  const uint32_t w_kb_bytes = (uint32_t)c.slab * 128u;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (ep.out32) prefetch_tmap(&tmO32);
    if (ep.out16) prefetch_tmap(&tmO16);
    if (ep.res32) prefetch_tmap(&tmR32);
    if (ep.res16) prefetch_tmap(&tmR16);
    for (int s = 0; s < c.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], (uint32_t)c.epi_warps); }
    for (int e = 0; e < 8; ++e) mbar_init(&rbar[e], 1);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < c.slab; i += blockDim.x) sbias[i] = ep.bias ? ep.bias[col_base + i] : 0.f;
  if (warp == 1) tmem_alloc(tmem_ptr, (uint32_t)(2 * c.slab));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      // the weight slab is a constant: its load overlaps the previous kernel's tail (PDL)
      if (live) {
        mbar_arrive_expect_tx(w_bar, (uint32_t)p.num_kb * w_kb_bytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          for (int hb = 0; hb < c.slab; hb += 128)
            tma_load_2d(smem + (uint32_t)kb * w_kb_bytes + (uint32_t)hb * 128u, &tmB, w_bar, kb * 64, col_base + hb);
      }
      pdl_wait();
      int kc = 0;
      for (int m_tile = j0; m_tile < c.m_tiles; m_tile += c.group) {
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % c.stages;
          mbar_wait(&empty[s], ((kc / c.stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], 16384u);
          tma_load_2d(smem + c.a_off + (uint32_t)s * 16384u, &tmA, &full[s], kb * 64, m_tile * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync() && live) {
      const uint32_t idesc = make_idesc_f16_f32(128, c.slab);
      mbar_wait(w_bar, 0);
      int kc = 0, it = 0;
      for (int m_tile = j0; m_tile < c.m_tiles; m_tile += c.group, ++it) {
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * c.slab);
        for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
          const int s = kc % c.stages;
          mbar_wait(&full[s], (kc / c.stages) & 1);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(smem_u32(smem + c.a_off + (uint32_t)s * 16384u));
          const uint64_t db = make_desc_sw128(smem_u32(smem + (uint32_t)kb * w_kb_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[a]);
      }
    }
  } else if (warp - 2 < c.epi_warps) {
    // ------------------------------------------------------------------ epilogue: 8 (or 4) warps
    const int e = warp - 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int nh = c.epi_warps == 8 ? 1 : 2;   // column halves this warp walks per 128-column sub-tile
    const int h0 = c.epi_warps == 8 ? (e >> 2) : 0;
    uint8_t* b32 = smem + c.sb32_off + e * 8192;
    uint8_t* b16 = smem + c.sb16_off + e * 4096;
    uint64_t* rb = &rbar[e];
    const uint32_t rowoff = (uint32_t)lane * 128u;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const bool has_res = (F32 && ep.res32) || (F16 && ep.res16);
    const int n_sub = (c.slab >> 7) * nh;      // drain units per tile: (sub-tile, half) pairs of this warp
    uint32_t rphase = 0;
    int it = 0;
    long long t_rd = 0, t_acc = 0, t_work = 0, t_all0 = p.dbg ? clock64() : 0;     // DV_GEMM_DBG cycle counters
    for (int m_tile = j0; m_tile < c.m_tiles; m_tile += c.group, ++it) {
      const int a = it & 1;
      const int row0 = m_tile * 128 + q * 32;            // first row of this warp
      const bool active = row0 < p.M;                    // warp-uniform
      const bool writer = row0 + lane < p.M;
      // rotary tables of this row (32 cos + 32 sin), shared by every head of the slab; fetched before the wait
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
        const int lcol = (sub / nh) * 128 + (h0 + sub % nh) * 64;   // first column of this unit inside the slab
        const int colw = col_base + lcol;                // ... and in the output matrix
        // (1) the previous TMA stores must have finished READING the staging boxes
        long long t0 = p.dbg ? clock64() : 0;
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        if (p.dbg) { const long long t1 = clock64(); t_rd += t1 - t0; t0 = t1; }
        // (2) residual boxes: in flight while the tile's MMAs run (sub 0) / while the previous sub-tile drains
        if (active && has_res && lane == 0) {
          uint32_t bytes = 0;
          if (F32 && ep.res32) bytes += 8192u;
          if (F16 && ep.res16) bytes += 4096u;
          mbar_arrive_expect_tx(rb, bytes);
          if (F32 && ep.res32) {
            tma_load_2d(b32, &tmR32, rb, colw, row0);
            tma_load_2d(b32 + 4096, &tmR32, rb, colw + 32, row0);
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
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * c.slab + lcol);
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
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
                const uint4 t4 = rt[ci * 4 + g];                    // angles ci * 16 + g * 4 + {0..3}
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
                *reinterpret_cast<uint4*>(b16 + rowoff + (((uint32_t)(ci * 4 + g) << 4) ^ swz)) =
                    *reinterpret_cast<const uint4*>(hv);
              }
            }
          }
        }
        // staging writes become visible to the async proxy; after the last sub-tile the accumulator buffer is free
        if (sub == n_sub - 1) tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (sub == n_sub - 1) mbar_arrive_cnt(&acc_empty[a]);
          if (active) {
            if (F32 && ep.out32) {
              tma_store_2d(&tmO32, b32, colw, row0);
              tma_store_2d(&tmO32, b32 + 4096, colw + 32, row0);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(2 * c.slab));
  }
}

static int g_sms_wres = 148;
static bool g_use_wres = true;
static constexpr uint32_t WRES_SMEM_MAX = 232448;      // 227 KB opt-in limit per CTA

int gemm_wres_init() {
  { const char* e = getenv("DV_GEMM_WRES"); g_use_wres = !(e && e[0] == '0'); }
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_wres_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WRES_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_wres_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WRES_SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_wres_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WRES_SMEM_MAX));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_sms_wres, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

// Picks the slab width / stage count for a plan; false when the shape does not fit (then the streaming kernel runs).
static bool wres_config(const GemmPlan& pl, long m_tiles, WresCfg* c) {
  const GemmParams& p = pl.p;
  const EpiParams& ep = p.epi;
  if (!g_use_wres || !pl.staged) return false;              // same operand / alignment rules as the staged kernel
  if ((p.K & 63) || p.K > 512 || (p.N & 127)) return false;
  if (ep.rope_cs && (!ep.rope16 || (ep.rope_cols & 127))) return false;
  if (m_tiles < 8) return false;
  const bool f32 = ep.out32 || ep.res32, f16 = ep.out16 || ep.res16;
  const uint32_t fixed = 1024u /*alignment slack*/ + 512u /*barriers*/ + 1024u /*bias*/;
  for (int cand = 0; cand < 4; ++cand) {                    // (slab 256, 8 warps), (256, 4), (128, 8), (128, 4)
    const int slab = cand < 2 ? 256 : 128, epi = (cand & 1) ? 4 : 8;
    const uint32_t st32 = f32 ? 8192u * (uint32_t)epi : 0u, st16 = f16 ? 4096u * (uint32_t)epi : 0u;
    const uint32_t staging = st32 + st16;
    if (p.N % slab) continue;
    const uint32_t wbytes = (uint32_t)slab * (uint32_t)p.K * 2u;
    if (wbytes + staging + fixed + 3u * 16384u > WRES_SMEM_MAX) continue;
    c->epi_warps = epi;
    int stages = (int)((WRES_SMEM_MAX - wbytes - staging - fixed) / 16384u);
    if (stages > 8) stages = 8;
    c->slab = slab;
    c->n_slabs = p.N / slab;
    if (c->n_slabs > g_sms_wres) return false;
    c->group = g_sms_wres / c->n_slabs;
    if (c->group > m_tiles) c->group = (int)m_tiles;
    c->stages = stages;
    c->m_tiles = (int)m_tiles;
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

bool gemm_wres_eligible(const GemmPlan& pl, long m_tiles) {
  WresCfg c;
  return wres_config(pl, m_tiles, &c);
}

int launch_gemm_wres(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st) {
  WresCfg c;
  if (!wres_config(pl, m_tiles, &c)) { set_error("launch_gemm_wres: shape not eligible"); return DV_ERR_INVALID; }
  const EpiParams& ep = p.epi;
  if (pl.staged_rows != p.M) {
    // exact row count: the TMA engine clips the last row tile, so rows >= M are neither read nor written
    if (ep.out32) DV_RC(tmap_encode_rows(&pl.tmO32, ep.out32, 4, p.N, p.M, (long)ep.ld32 * 4, 32, 32));
    if (ep.out16) DV_RC(tmap_encode_rows(&pl.tmO16, ep.out16, 2, p.N, p.M, (long)ep.ld16 * 2, 64, 32));
    if (ep.res32) DV_RC(tmap_encode_rows(&pl.tmR32, ep.res32, 4, p.N, p.M, (long)ep.ldr32 * 4, 32, 32));
    if (ep.res16) DV_RC(tmap_encode_rows(&pl.tmR16, ep.res16, 2, p.N, p.M, (long)ep.ldr16 * 2, 64, 32));
    pl.staged_rows = p.M;
  }
  const int grid = c.group * c.n_slabs;
  const bool f32 = ep.out32 || ep.res32, f16 = ep.out16 || ep.res16;
  static long long* d_dbg = nullptr;
  static const bool want_dbg = getenv("DV_GEMM_DBG") != nullptr;     // diagnostics: epilogue cycle counters of CTA 0
  GemmParams pd = p;
  if (want_dbg) {
    if (!d_dbg) cudaMalloc(&d_dbg, 64);
    cudaMemsetAsync(d_dbg, 0, 64, st);
    pd.dbg = d_dbg;
  }
  if (f32 && f16)
    DV_CUDA_OK(launch_pdl(umma_gemm_wres_kernel<true, true>, dim3(grid), dim3(320), c.smem_bytes, st, pl.tmA, pl.tmB,
                          pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, pd, c));
  else if (f32)
    DV_CUDA_OK(launch_pdl(umma_gemm_wres_kernel<true, false>, dim3(grid), dim3(320), c.smem_bytes, st, pl.tmA, pl.tmB,
                          pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, pd, c));
  else
    DV_CUDA_OK(launch_pdl(umma_gemm_wres_kernel<false, true>, dim3(grid), dim3(320), c.smem_bytes, st, pl.tmA, pl.tmB,
                          pl.tmO32, pl.tmO16, pl.tmR32, pl.tmR16, pd, c));
  DV_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[wres dbg] M %d N %d K %d slab %d stages %d group %d | CTA0 epi warp: tiles %lld total %lld cyc: "
            "store-read wait %lld, acc wait %lld, work %lld | prologue %lld, entry->stores drained %lld\n", p.M, p.N, p.K,
            c.slab, c.stages, c.group, h[4], h[3], h[0], h[1], h[2], h[5], h[6]);
  }
  return DV_OK;
}

}  // namespace dv
