// MixVPR / ResNet-50 stem: 7x7 stride-2 pad-3 convolution 3 -> 64 (+ folded BatchNorm bias, ReLU) on 320 x 320 frames as
// ONE implicit-GEMM kernel: the im2col operand is built in shared memory by producer warps, never in HBM.
//
// Why: the two-kernel version wrote a [B*25600, 192] fp16 im2col matrix (9.8 MB per frame, 314 MB per 32 frames) and read
// it back in a GEMM: 161 us + 67 us per 32 frames for 0.24 GMAC/frame, both HBM-bound.  Here a persistent CTA per SM
//   * keeps the [64 x 192] filter matrix (k = (r*7+s)*3 + c, zero-padded 147 -> 192) resident (24 KB, SWIZZLE_128B);
//   * per 8 x 16-pixel output tile (M = 128): eight producer warps stage the 21 x 37 x 3 input patch (4.6 KB, coalesced
//     row reads of the fp16 NHWC frame) and scatter it into the K-major SWIZZLE_128B A operand [128 x 192] - a filter
//     row's 21 values (7 pixels x 3 channels) are contiguous in the patch, so k -> patch[(2*oy + k/21) * 111 + 6*ox + k%21];
//   * 12 MMAs (M=128, N=64, K=16) per tile into one of two TMEM accumulators; four epilogue warps add the bias, ReLU and
//     store NHWC fp16 (128 contiguous bytes per pixel).
// The same in-kernel im2col idea as SuperPoint's conv1a (conv_halo.cu, FUSE == 2).
#include <type_traits>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

namespace {
constexpr int IMG = 320, OUT = 160;                 // input / output side
constexpr int TOH = 8, TOW = 16;                    // output tile -> M = 128
constexpr int PH = TOH * 2 + 5, PW = TOW * 2 + 5;   // input patch 21 x 37 pixels
constexpr int PROW = PW * 3;                        // 111 halfs per patch row
constexpr int PATCH = PH * PROW;                    // 2331 halfs
constexpr int KB = 3;                               // 64-wide K blocks (192)
constexpr int A_STAGE = KB * 128 * 128;             // 49152 B
constexpr int W_BYTES = KB * 64 * 128;              // 24576 B
constexpr int OFF_W = 0;
constexpr int OFF_A = W_BYTES;                      // 24576 (1024-aligned)
constexpr int OFF_PATCH = OFF_A + 2 * A_STAGE;      // 122880
constexpr int PATCH_STRIDE = 4736;                  // >= 2331 * 2, 64-byte multiple
constexpr int OFF_BAR = OFF_PATCH + 2 * PATCH_STRIDE;
constexpr int OFF_BIAS = OFF_BAR + 128;
constexpr int SMEM_BYTES = OFF_BIAS + 256 + 1024;
constexpr int NPRODW = 8;                           // producer warps
constexpr int THREADS = 32 * (2 + 4 + NPRODW);      // 448
constexpr int TILES_X = OUT / TOW, TILES_Y = OUT / TOH, TILES_IMG = TILES_X * TILES_Y;   // 10 x 20
}  // namespace

struct StemParams {
  const __half* img;    // [N, 320, 320, 3]
  const float* bias;    // [64]
  __half* out;          // [N, 160, 160, 64]
  int total_tiles;
};

__global__ void __launch_bounds__(THREADS, 1) stem_conv_kernel(const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [2]
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_bar = acc_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmW);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], NPRODW); mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) sbias[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  {  // K padding (k >= 152: chunks 19..23 of the third K block) of both A stages: zero once, never written again
    uint4* az = reinterpret_cast<uint4*>(smem + OFF_A);
    for (int i = threadIdx.x; i < 2 * A_STAGE / 16; i += blockDim.x) az[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one_sync()) {                          // the filter matrix is a constant: loaded before the PDL wait
      mbar_arrive_expect_tx(w_bar, W_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem + OFF_W + kb * 8192, &tmW, w_bar, kb * 64, 0);
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, 64);
      mbar_wait(w_bar, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&acc_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_wait(&a_full[s], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(s * 64);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = make_desc_sw128(smem_u32(smem + OFF_A + s * A_STAGE + kb * 16384));
          const uint64_t db = make_desc_sw128(smem_u32(smem + OFF_W + kb * 8192));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
        }
        tc_commit(&a_empty[s]);
        tc_commit(&acc_full[s]);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue: 4 warps, one TMEM lane quarter each
    pdl_wait();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int oyl = row >> 4, oxl = row & 15;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const int img = tile / TILES_IMG;
      const int rem = tile - img * TILES_IMG;
      const int ty = rem / TILES_X, tx = rem - ty * TILES_X;
      mbar_wait(&acc_full[s], (it >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 64), r0);
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 64 + 32), r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(&acc_empty[s]);
      __half* dst = p.out + ((((int64_t)img * OUT + ty * TOH + oyl) * OUT) + tx * TOW + oxl) * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        __align__(16) __half2 hv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g * 8 + 2 * j;
          const uint32_t u0 = c < 32 ? r0[c] : r1[c - 32], u1 = c < 32 ? r0[c + 1] : r1[c - 31];
          hv[j] = __floats2half2_rn(fmaxf(__uint_as_float(u0) + sbias[c], 0.f), fmaxf(__uint_as_float(u1) + sbias[c + 1], 0.f));
        }
        *reinterpret_cast<uint4*>(dst + g * 8) = *reinterpret_cast<const uint4*>(hv);
      }
    }
  } else {
    // ------------------------------------------------------------------ producers: frame patch -> im2col A operand
    pdl_wait();
    const int pt = threadIdx.x - 32 * 6;             // 0..255
    constexpr int PPT = (PATCH + 32 * NPRODW - 1) / (32 * NPRODW);   // patch halfs per thread (10)
    // the patch of tile it+1 is fetched into registers while the A operand of tile it is built: the L2 latency of the
    // 2-byte gathers stays off the per-tile chain
    auto fetch = [&](int tile, __half (&v)[PPT]) {
      const int img = tile / TILES_IMG;
      const int rem = tile - img * TILES_IMG;
      const int ty = rem / TILES_X, tx = rem - ty * TILES_X;
      const int iy0 = ty * TOH * 2 - 3, ix0 = tx * TOW * 2 - 3;
      const __half* src = p.img + (int64_t)img * IMG * IMG * 3;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int i = pt + j * 32 * NPRODW;
        const int r = i / PROW, q = i - r * PROW;
        const int iy = iy0 + r, ix = ix0 + q / 3;
        v[j] = __float2half_rn(0.f);
        if (i < PATCH && iy >= 0 && iy < IMG && ix >= 0 && ix < IMG) v[j] = src[((int64_t)iy * IMG + ix0) * 3 + q];
      }
    };
    __half cur[PPT];
    if ((int)blockIdx.x < p.total_tiles) fetch(blockIdx.x, cur);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      __half* patch = reinterpret_cast<__half*>(smem + OFF_PATCH + s * PATCH_STRIDE);
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int i = pt + j * 32 * NPRODW;
        if (i < PATCH) patch[i] = cur[j];
      }
      if (tile + (int)gridDim.x < p.total_tiles) fetch(tile + gridDim.x, cur);
      asm volatile("bar.sync 2, 256;" ::: "memory");   // patch visible to all producer warps
      mbar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);     // the MMAs that read this A stage have retired
      uint8_t* a = smem + OFF_A + s * A_STAGE;
      // 128 rows x 19 real chunks of 8 halfs (k < 152; chunks 19..23 are the K padding and stay zero from the start).
      // A thread owns one row and a compile-time set of chunks (warps 0-3: chunks 0..9, warps 4-7: chunks 10..18), so
      // every k -> (filter row, offset) split is a constant and a chunk is 8 LDS.U16 at fixed offsets + one STS.128.
      {
        const int row = pt & 127;
        const int oyl = row >> 4, oxl = row & 15;
        const __half* pr = patch + (oyl * 2) * PROW + oxl * 6;
        uint8_t* arow = a + row * 128;
        const int sw = row & 7;
        auto build = [&](auto first, auto count) {
          constexpr int C0 = decltype(first)::value, CN = decltype(count)::value;
#pragma unroll
          for (int n = 0; n < CN; ++n) {
            const int c = C0 + n;
            __align__(16) __half v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = c * 8 + j;
              v[j] = k < 147 ? pr[(k / 21) * PROW + (k % 21)] : __float2half_rn(0.f);
            }
            *reinterpret_cast<uint4*>(arow + (c >> 3) * 16384 + (((c & 7) ^ sw) << 4)) = *reinterpret_cast<const uint4*>(v);
          }
        };
        if (pt < 128) build(std::integral_constant<int, 0>{}, std::integral_constant<int, 10>{});
        else build(std::integral_constant<int, 10>{}, std::integral_constant<int, 9>{});
      }
      fence_proxy_async_smem();                        // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(&a_full[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int g_sms_stem = 148;

int stem_conv_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_sms_stem, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

int plan_stem_conv(StemPlan* pl, const __half* img16, int n_cap, const __half* w, const float* bias, __half* out) {
  pl->img = img16; pl->n_cap = n_cap; pl->bias = bias; pl->out = out;
  const uint64_t wd[2] = {192, 64};
  const uint64_t ws[1] = {192 * 2};
  const uint32_t wb[2] = {64, 64};
  return tmap_encode_f16(&pl->tmW, w, 2, wd, ws, wb, /*swizzle128=*/true);
}

int launch_stem_conv(const StemPlan& pl, int n_img, cudaStream_t st) {
  if (n_img <= 0) return DV_OK;
  if (n_img > pl.n_cap) { set_error("launch_stem_conv: batch exceeds plan capacity"); return DV_ERR_CAPACITY; }
  StemParams p;
  p.img = pl.img; p.bias = pl.bias; p.out = pl.out;
  p.total_tiles = n_img * TILES_IMG;
  const int grid = p.total_tiles < g_sms_stem ? p.total_tiles : g_sms_stem;
  DV_CUDA_OK(launch_pdl(stem_conv_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, st, pl.tmW, p));
  DV_CUDA_OK(cudaGetLastError());
  return DV_OK;
}

}  // namespace dv
