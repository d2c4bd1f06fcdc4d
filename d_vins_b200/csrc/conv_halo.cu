// Weights-stationary, halo-tile 3x3 convolution (64 -> 64 channels) on tcgen05 - the SuperPoint conv1b / conv2a /
// conv2b layers (65 % of SuperPoint's MACs).
//
// Why: the tap-per-TMA implicit GEMM (gemm_umma.cu) re-reads every input pixel nine times and the filter bank once
// per tile; with 128x64 tiles that is 24 KB of L2->SM traffic per 128x64x64 MACs and the SM's L2 port, not the tensor
// pipe, is the limit (measured 26 % of the bf16 peak, ncu r01).  Here a persistent CTA per SM keeps all nine 64x64
// filter taps in shared memory (72 KB) and brings each 16x8-pixel output tile's 18x10 halo ONCE (23 KB) with a single
// 5-D TMA box over the channel-blocked activation layout [N][C/8][H][W][8].  In shared memory the box is
// [c/8][18][10][8ch]: every run of 8 pixels along w is a contiguous 128-byte UMMA core matrix, so the A operand of tap
// (r,s) is just the same buffer viewed through a no-swizzle descriptor whose start address is shifted by
// (r*10 + s) * 16 bytes (SBO = one halo row = 160 B, LBO = one channel group = 2880 B).  36 MMAs (9 taps x 4 k-steps,
// M=128, N=64, K=16) per tile accumulate into one of two TMEM buffers while the epilogue warps drain the other.
#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

namespace {
constexpr int TH = 16, TW = 8;                      // output tile (pixels) -> M = 128
constexpr int HH = TH + 2, HW = TW + 2;             // halo tile
constexpr int CG = 8;                               // channel groups of 8 (C_in = 64)
constexpr int HALO_BYTES = CG * HH * HW * 16;       // 23040
constexpr int HALO_STRIDE = 23552;                  // padded to a multiple of 512 B
constexpr int STAGES = 3;
constexpr int W_BYTES = 9 * 64 * 128;               // nine [64 x 64] fp16 taps, SWIZZLE_128B rows
constexpr int OFF_W = 0;
constexpr int OFF_HALO = W_BYTES;                   // 73728 (1024-aligned)
constexpr int OFF_BAR = OFF_HALO + STAGES * HALO_STRIDE;
constexpr int OFF_BIAS = OFF_BAR + 256;
constexpr int OFF_GRAY = OFF_BIAS + 512;            // FUSE1A: two (TH+4) x (TW+4) fp32 windows of the gray frame
constexpr int GRAY_H = TH + 4, GRAY_W = TW + 4;
constexpr int NPROD = 8;                            // FUSE == 1 producer warps (one per output channel group of conv1a)
// FUSE == 2 (conv1a on the tensor cores): im2col operand A1 [256 halo-pixel rows x 16 taps] fp16, no-swizzle K-major
// ([k/8][row][8]: SBO 128 B, LBO 4096 B), double-buffered; W1a' [64 x 16] in the same canonical layout (LBO 1024 B)
constexpr int OFF_A1 = (OFF_GRAY + 2 * GRAY_H * GRAY_W * 4 + 127) & ~127;
constexpr int A1_BYTES = 2 * 256 * 16;              // 8192
constexpr int OFF_W1A = OFF_A1 + 2 * A1_BYTES;
constexpr int SMEM_BYTES = OFF_W1A + 2048 + 1024;
constexpr int NMID = 8;                             // FUSE == 2: warps 10-17 move conv1a from TMEM into the halo buffer
constexpr int NIM2COL = 2;                          // FUSE == 2: warps 18-19 build A1 from the u8 frame
constexpr int NPIX = HH * HW;                       // 180 halo pixels
}  // namespace

struct HaloParams {
  int H, W, tiles_w, tiles_h, total_tiles;
  const float* bias;
  __half* out;
  int out_blocked, relu, pool;
  // FUSE1A: conv1a (1 -> 64, 3x3, ReLU) is evaluated by the producer warps straight into the halo buffer
  const float* gray;   // [N, H, W] fp32
  const float* w1a;    // [64, 9]
  const float* b1a;    // [64]
  const uint8_t* img8; // FUSE == 2: the 1-channel u8 frame itself [N, H, W]
};

// FUSE1A = true: SuperPoint conv1a + conv1b in one kernel.  The 64-channel full-resolution conv1a activation (46 MB per
// 480x752 frame, written and re-read through HBM by the two-kernel version) never exists in memory: eight producer
// warps (one per group of 8 output channels, lanes = halo pixels) compute the 18x10x64 halo of conv1a on CUDA cores
// from a 20x12 window of the gray frame and store it as fp16 in exactly the [c/8][18][10][8] layout the TMA box
// would have produced; pixels outside the image are zeros (conv1b's zero padding), not conv1a evaluated out of range.
//
// FUSE == 2: conv1a itself runs on the tensor cores.  Two warps gather each halo pixel's 3x3 neighbourhood from the u8
// frame into an im2col operand A1 [180 (-> 256) x 16] (u8 / 256 is exact in fp16; the 256/255 lands in the weights),
// the MMA thread issues D1[256 x 64] = A1 x W1a'^T (two M=128, N=64, K=16 instructions, 64 tensor cycles) one tile
// ahead of conv1b, four "mid" warps read D1 from TMEM, add the bias, ReLU, convert and store it as the
// [c/8][18][10][8] fp16 halo of conv1b (zeros outside the image).  ~100 instructions per thread per tile instead of the
// ~700 of the CUDA-core producer, so the conv1b MMAs stay the critical path.
template <int FUSE>
__global__ void __launch_bounds__(FUSE == 1 ? 320 + 32 * NPROD : (FUSE == 2 ? 320 + 32 * (NMID + NIM2COL) : 320), 1)
conv3x3_halo64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const HaloParams p) {
  constexpr bool FUSE1A = FUSE == 1;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_bar = acc_empty + 2;
  uint64_t* a1_full = w_bar + 1;                   // FUSE == 2: [2] im2col operand built / consumed, D1 ready / drained
  uint64_t* a1_empty = a1_full + 2;
  uint64_t* d1_full = a1_empty + 2;
  uint64_t* d1_empty = d1_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d1_empty + 2);
  constexpr uint32_t TMEM_COLS = FUSE == 2 ? 512 : 128;   // D2: 2 x 64 columns; FUSE == 2 adds D1: 2 x (2 x 64) at column 128
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], FUSE1A ? NPROD : (FUSE == 2 ? NMID : 1)); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a1_full[b], NIM2COL); mbar_init(&a1_empty[b], 1); mbar_init(&d1_full[b], 1); mbar_init(&d1_empty[b], NMID);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  if (threadIdx.x >= 64 && threadIdx.x < 128) sbias[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  if (FUSE == 2 && threadIdx.x >= 128 && threadIdx.x < 192) sbias[threadIdx.x - 64] = p.b1a[threadIdx.x - 128];   // [64..127]
  if (FUSE == 2) {
    // W1a' = w1a * 256/255 as fp16 in the canonical no-swizzle K-major layout; taps 9..15 and A1's spare rows are zero
    __half* w1 = reinterpret_cast<__half*>(smem + OFF_W1A);
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
      const int n = i >> 4, k = i & 15;
      const float v = k < 9 ? p.w1a[n * 9 + k] * (256.0f / 255.0f) : 0.f;
      w1[(k >> 3) * 512 + n * 8 + (k & 7)] = __float2half_rn(v);
    }
    uint4* a1z = reinterpret_cast<uint4*>(smem + OFF_A1);
    for (int i = threadIdx.x; i < 2 * A1_BYTES / 16; i += blockDim.x) a1z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      // the 72 KB of filter taps are constants: their load overlaps the previous kernel's tail (PDL), then wait
      mbar_arrive_expect_tx(w_bar, W_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem + OFF_W + t * 8192, &tmW, w_bar, t * 64, 0);
      pdl_wait();
      int it = 0;
      if (FUSE == 0)
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        const int img = tile / tiles_per_img;
        const int rem = tile - img * tiles_per_img;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        mbar_arrive_expect_tx(&full[s], HALO_BYTES);
        tma_load_5d(smem + OFF_HALO + s * HALO_STRIDE, &tmX, &full[s], 0, tw_i * TW - 1, th_i * TH - 1, 0, img);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, 64);
      mbar_wait(w_bar, 0);
      const uint64_t db_base = make_desc_sw128(smem_u32(smem + OFF_W));
      // FUSE == 2: conv1a of tile j, D1[b] = A1[b] x W1a'^T, issued one tile ahead of the conv1b MMAs that consume it
      const uint64_t dw1 = make_desc_noswz(smem_u32(smem + OFF_W1A), 1024, 128);
      auto issue_conv1a = [&](int j) {
        const int b = j & 1;
        mbar_wait(&a1_full[b], (j >> 1) & 1);
        mbar_wait(&d1_empty[b], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a1 = smem_u32(smem + OFF_A1 + b * A1_BYTES);
        const uint32_t d1 = tmem_base + 128u + (uint32_t)(b * 128);
        tc_mma_f16(d1, make_desc_noswz(a1, 4096, 128), dw1, idesc, 0u);
        tc_mma_f16(d1 + 64u, make_desc_noswz(a1 + 128 * 16, 4096, 128), dw1, idesc, 0u);
        tc_commit(&a1_empty[b]);
        tc_commit(&d1_full[b]);
      };
      const int my_tiles = ((int)blockIdx.x < p.total_tiles) ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      if (FUSE == 2 && my_tiles > 0) issue_conv1a(0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES, a = it & 1;
        if (FUSE == 2 && it + 1 < my_tiles) issue_conv1a(it + 1);
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        // descriptors differ from the per-stage / per-tap bases only in the 14-bit start-address field (addr >> 4)
        const uint64_t da_base = make_desc_noswz(smem_u32(smem + OFF_HALO + s * HALO_STRIDE), HH * HW * 16, HW * 16);
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * 64);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int r = t / 3, sx = t - r * 3;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma_f16(d_tmem, da_base + (uint64_t)(2 * kk * HH * HW + r * HW + sx), db_base + (uint64_t)(t * 512 + kk * 2),
                       idesc, (uint32_t)((t | kk) != 0));
        }
        tc_commit(&empty[s]);
        tc_commit(&acc_full[a]);
      }
    }
  } else if (FUSE == 2 && warp >= 10 && warp < 10 + NMID) {
    // ------------------------------------------------------------------ mid epilogue: conv1a TMEM -> conv1b halo
    const int q = warp & 3;                         // TMEM lane quarter
    const int half = (warp - 10) >> 2;              // 32-channel half of conv1a's 64 outputs
    float bv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bv[j] = sbias[64 + half * 32 + j];
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it % STAGES, b = it & 1;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
      mbar_wait(&d1_full[b], (it >> 1) & 1);
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);                // the conv1b MMAs that read this halo slot have retired
      tc_fence_after();
      uint8_t* halo = smem + OFF_HALO + s * HALO_STRIDE;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        if (blk * 128 + q * 32 >= NPIX) continue;                   // warp-uniform
        const int px = blk * 128 + q * 32 + lane;
        const int hh = px / HW, ww = px - hh * HW;
        const int y = th_i * TH - 1 + hh, x = tw_i * TW - 1 + ww;   // image position of this halo pixel
        const bool inside = px < NPIX && y >= 0 && y < p.H && x >= 0 && x < p.W;
        {
          uint32_t rr[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)(b * 128 + blk * 64 + half * 32), rr);
          tmem_ld_wait();
          if (px < NPIX) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              __align__(16) __half2 hv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float v0 = fmaxf(__uint_as_float(rr[g * 8 + 2 * j]) + bv[g * 8 + 2 * j], 0.f);
                const float v1 = fmaxf(__uint_as_float(rr[g * 8 + 2 * j + 1]) + bv[g * 8 + 2 * j + 1], 0.f);
                hv[j] = inside ? __floats2half2_rn(v0, v1) : __floats2half2_rn(0.f, 0.f);
              }
              *reinterpret_cast<uint4*>(halo + (half * 4 + g) * (NPIX * 16) + px * 16) = *reinterpret_cast<const uint4*>(hv);
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();             // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) { mbar_arrive_cnt(&full[s]); mbar_arrive_cnt(&d1_empty[b]); }
    }
  } else if (FUSE == 2 && warp >= 10 + NMID) {
    // ------------------------------------------------------------------ im2col producers: u8 frame -> A1
    const int pt = threadIdx.x - 32 * (10 + NMID);  // 0..63
    __half* gbuf = reinterpret_cast<__half*>(smem + OFF_GRAY);
    constexpr int WPT = (GRAY_H * GRAY_W + 32 * NIM2COL - 1) / (32 * NIM2COL);   // window pixels per thread (4)
    // the (TH+4) x (TW+4) u8 window of a tile; pixels outside the image are conv1a's own zero padding.  The window of
    // tile it+1 is fetched into registers while A1 of tile it is built, so the L2 latency stays off the per-tile chain.
    auto fetch = [&](int tile, float (&v)[WPT]) {
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
      const int h0 = th_i * TH - 2, w0 = tw_i * TW - 2;
#pragma unroll
      for (int j = 0; j < WPT; ++j) {
        const int i = pt + j * 32 * NIM2COL;
        const int gy = i / GRAY_W, gx = i - gy * GRAY_W;
        const int y = h0 + gy, x = w0 + gx;
        const bool ok = i < GRAY_H * GRAY_W && y >= 0 && y < p.H && x >= 0 && x < p.W;
        v[j] = ok ? (float)__ldg(p.img8 + ((int64_t)img * p.H + y) * p.W + x) : 0.f;
      }
    };
    float cur[WPT];
    if ((int)blockIdx.x < p.total_tiles) fetch(blockIdx.x, cur);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      __half* g = gbuf + b * (GRAY_H * GRAY_W);
#pragma unroll
      for (int j = 0; j < WPT; ++j) {
        const int i = pt + j * 32 * NIM2COL;
        if (i < GRAY_H * GRAY_W) g[i] = __float2half_rn(cur[j] * (1.0f / 256.0f));      // u8 / 256: exact in fp16
      }
      if (tile + (int)gridDim.x < p.total_tiles) fetch(tile + gridDim.x, cur);
      asm volatile("bar.sync 2, 64;" ::: "memory");                 // window visible to both producer warps
      mbar_wait(&a1_empty[b], ((it >> 1) & 1) ^ 1);                 // conv1a MMAs of tile it-2 have retired
      uint8_t* a1 = smem + OFF_A1 + b * A1_BYTES;
      for (int px = pt; px < NPIX; px += 32 * NIM2COL) {
        const int hh = px / HW, ww = px - hh * HW;
        __align__(16) __half v[16];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q3 = 0; q3 < 3; ++q3) v[r * 3 + q3] = g[(hh + r) * GRAY_W + ww + q3];
#pragma unroll
        for (int k = 9; k < 16; ++k) v[k] = __float2half_rn(0.f);
        *reinterpret_cast<uint4*>(a1 + px * 16) = *reinterpret_cast<const uint4*>(&v[0]);
        *reinterpret_cast<uint4*>(a1 + 4096 + px * 16) = *reinterpret_cast<const uint4*>(&v[8]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(&a1_full[b]);
    }
  } else if (FUSE1A && warp >= 10) {
    // ------------------------------------------------------------------ conv1a producers (CUDA cores)
    const int cg = warp - 10;                       // output channel group of conv1a owned by this warp
    const int pt = threadIdx.x - 320;               // 0..255 among the producer threads
    float wr[8][9], br[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      br[c] = p.b1a[cg * 8 + c];
#pragma unroll
      for (int t = 0; t < 9; ++t) wr[c][t] = p.w1a[(cg * 8 + c) * 9 + t];
    }
    float* gbuf = reinterpret_cast<float*>(smem + OFF_GRAY);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it % STAGES;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
      const int h0 = th_i * TH - 2, w0 = tw_i * TW - 2;            // top-left of the gray window
      float* g = gbuf + (it & 1) * (GRAY_H * GRAY_W);
      if (pt < GRAY_H * GRAY_W) {
        const int gy = pt / GRAY_W, gx = pt - gy * GRAY_W;
        const int y = h0 + gy, x = w0 + gx;
        g[pt] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? __ldg(p.gray + ((int64_t)img * p.H + y) * p.W + x) : 0.f;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");                // window visible to all producer warps
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);                // the MMAs that read this halo slot have retired
      uint8_t* halo = smem + OFF_HALO + s * HALO_STRIDE + cg * (HH * HW * 16);
      for (int px = lane; px < HH * HW; px += 32) {
        const int hh = px / HW, ww = px - hh * HW;
        const int y = th_i * TH - 1 + hh, x = tw_i * TW - 1 + ww;  // image position of this halo pixel
        __align__(16) __half2 hv[4];
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
          float in[9];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3) in[r * 3 + q3] = g[(hh + r) * GRAY_W + ww + q3];
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            float a0 = br[c], a1 = br[c + 1];
#pragma unroll
            for (int t = 0; t < 9; ++t) { a0 = fmaf(wr[c][t], in[t], a0); a1 = fmaf(wr[c + 1][t], in[t], a1); }
            hv[c >> 1] = __floats2half2_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) hv[c] = __floats2half2_rn(0.f, 0.f);
        }
        *reinterpret_cast<uint4*>(halo + px * 16) = *reinterpret_cast<const uint4*>(hv);
      }
      fence_proxy_async_smem();             // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(&full[s]);
    }
  } else {
    // 8 epilogue warps: warp w owns TMEM lane quarter (w & 3) and the 32-column half ((w - 2) >> 2)
    const int q = warp & 3;
    const int c = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    const int Ho = p.pool ? (p.H >> 1) : p.H, Wo = p.pool ? (p.W >> 1) : p.W;
    float bv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bv[j] = sbias[c * 32 + j];
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
      const int h = th_i * TH + hl, w = tw_i * TW + wl;
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      uint32_t rr[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c * 32), rr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(&acc_empty[a]);      // TMEM buffer free: the next tile's MMAs may start
      bool writer; int ho, wo;
      if (p.pool) {
        ho = h >> 1; wo = w >> 1;
        writer = !(hl & 1) && !(wl & 1) && ho < Ho && wo < Wo;
      } else {
        ho = h; wo = w;
        writer = h < p.H && w < p.W;
      }
      __align__(16) __half2 hv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v0 = __uint_as_float(rr[2 * j]) + bv[2 * j];
        float v1 = __uint_as_float(rr[2 * j + 1]) + bv[2 * j + 1];
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        hv[j] = __floats2half2_rn(v0, v1);
      }
      if (p.pool) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t u = *reinterpret_cast<uint32_t*>(&hv[j]);
          uint32_t o = __shfl_xor_sync(0xffffffffu, u, 1);
          __half2 m = __hmax2(*reinterpret_cast<__half2*>(&u), *reinterpret_cast<__half2*>(&o));
          u = *reinterpret_cast<uint32_t*>(&m);
          o = __shfl_xor_sync(0xffffffffu, u, 8);
          hv[j] = __hmax2(m, *reinterpret_cast<__half2*>(&o));
        }
      }
      if (writer) {
        const uint4* src = reinterpret_cast<const uint4*>(hv);
        if (p.out_blocked) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int64_t off = ((((int64_t)img * 8 + c * 4 + g) * Ho + ho) * Wo + wo) * 8;
            *reinterpret_cast<uint4*>(p.out + off) = src[g];
          }
        } else {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (((int64_t)img * Ho + ho) * Wo + wo) * 64 + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) dst[g] = src[g];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int g_num_sms = 148;

int conv_halo_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(conv3x3_halo64_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(conv3x3_halo64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(conv3x3_halo64_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

int plan_conv3x3_halo64(HaloPlan* pl, const __half* x, int n_cap, int H, int W, const __half* w, const float* bias,
                        __half* out, int out_blocked, int relu, int pool) {
  pl->H = H; pl->W = W; pl->n_cap = n_cap;
  pl->tiles_w = cdiv(W, TW); pl->tiles_h = cdiv(H, TH);
  pl->bias = bias; pl->out = out; pl->out_blocked = out_blocked; pl->relu = relu; pl->pool = pool;
  // activations [N][8][H][W][8]
  const uint64_t dims[5] = {8, (uint64_t)W, (uint64_t)H, 8, (uint64_t)n_cap};
  const uint64_t strides[4] = {16, (uint64_t)W * 16, (uint64_t)H * W * 16, (uint64_t)8 * H * W * 16};
  const uint32_t box[5] = {8, HW, HH, CG, 1};
  int rc = tmap_encode_f16(&pl->tmX, x, 5, dims, strides, box, /*swizzle128=*/false);
  if (rc) return rc;
  const uint64_t wd[2] = {576, 64};
  const uint64_t ws[1] = {576 * 2};
  const uint32_t wb[2] = {64, 64};
  return tmap_encode_f16(&pl->tmW, w, 2, wd, ws, wb, /*swizzle128=*/true);
}

int launch_conv_halo64(const HaloPlan& pl, int n_img, cudaStream_t st) {
  if (n_img <= 0) return DV_OK;
  if (n_img > pl.n_cap) { set_error("launch_conv_halo64: batch exceeds plan capacity"); return DV_ERR_CAPACITY; }
  HaloParams p;
  p.H = pl.H; p.W = pl.W; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.total_tiles = n_img * pl.tiles_w * pl.tiles_h;
  p.bias = pl.bias; p.out = pl.out; p.out_blocked = pl.out_blocked; p.relu = pl.relu; p.pool = pl.pool;
  p.gray = pl.gray; p.w1a = pl.w1a; p.b1a = pl.b1a; p.img8 = pl.img8;
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  if (pl.img8)
    DV_CUDA_OK(launch_pdl(conv3x3_halo64_kernel<2>, dim3(grid), dim3(320 + 32 * (NMID + NIM2COL)), SMEM_BYTES, st, pl.tmX,
                          pl.tmW, p));
  else if (pl.gray)
    DV_CUDA_OK(launch_pdl(conv3x3_halo64_kernel<1>, dim3(grid), dim3(320 + 32 * NPROD), SMEM_BYTES, st, pl.tmX, pl.tmW, p));
  else
    DV_CUDA_OK(launch_pdl(conv3x3_halo64_kernel<0>, dim3(grid), dim3(320), SMEM_BYTES, st, pl.tmX, pl.tmW, p));
  DV_CUDA_OK(cudaGetLastError());
  return DV_OK;
}

}  // namespace dv
