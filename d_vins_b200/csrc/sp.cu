// SuperPoint on B200: VGG encoder + detector/descriptor heads on tcgen05 implicit-GEMM tiles (gemm_umma.cu),
// and the HBM-bound post-net as coalesced kernels: softmax-65 + depth-to-space, fused 3-round 9x9 NMS tile kernel,
// border/threshold/candidate emission, radix-select top-k + bitonic sort, bilinear descriptor sampling + L2.
// Semantics follow the reference's export/superpoint.py:52-224 and export/ultrapoint.py:101-127 (see oracle/).
#include <stdlib.h>

#include <algorithm>

#include "engine.h"

namespace dv {

#define DV_NMS_DEFAULT_LARGE false

struct SpNet {
  // weights
  float *w1a = nullptr, *b1a = nullptr;                 // conv1a fp32 [64,9],[64]
  __half* w[9] = {};                                    // 1b,2a,2b,3a,3b,4a,4b,PD(512),(unused)
  float* bias[9] = {};
  __half *wPb = nullptr, *wDb = nullptr;
  float *bPb = nullptr, *bDb = nullptr;
  // activations (NHWC fp16)
  float* gray = nullptr;                                // [B,H,W] f32
  __half *a1a = nullptr, *a1b = nullptr, *a2a = nullptr, *a2b = nullptr, *a3a = nullptr, *a3b = nullptr,
         *a4a = nullptr, *a4b = nullptr, *aPD = nullptr;
  float *logits = nullptr, *dmap = nullptr;             // [B*h8*w8, 80], [B*h8*w8, 256]
  GemmPlan p1b, p2a, p2b, p3a, p3b, p4a, p4b, pPD, pPb, pDb;
  HaloPlan h1b, h2a, h2b;                               // weights-stationary halo kernels for the 64->64 layers
  Halo128Plan h3a, h3b, h4a, h4b, hPD;                  // 256-pixel halo tiles, streamed weights (conv_halo128.cu)
  bool use_halo128 = false;
  bool gray_valid = false;                              // the fp32 gray frame of the current batch exists
  bool nms_large = false;                               // DV_NMS_TILE=L / S: 128x64 or 64x32 NMS tiles (A/B)
  bool nms_full = false;                                // DV_NMS_TILE=F: 64x32 tiles, every pool over the whole region (r01)
  bool use_halo = true;                                 // DV_SP_HALO=0: generic tap-per-TMA kernel (debug toggle)
  int fuse1a_tc = 0;                                    // DV_SP_FUSE1A=2: conv1a on the tensor cores inside conv1b (conv_halo.cu FUSE == 2)
  bool fuse1a = false;                                  // DV_SP_FUSE1A=1: conv1a inside conv1b's producer (correct, but the
                                                        // CUDA-core producer is 3x slower than the tensor main loop - r01)
  // post
  float *smap = nullptr, *nms = nullptr;                // [B,H8,W8]
  unsigned long long* cand = nullptr;                   // [B, H8*W8]
  int* cand_cnt = nullptr;                              // [B]
  int* kpts = nullptr;                                  // [B,K,2] int32 (x,y)
  float* kpts_f = nullptr;                              // [B,K,2] float (x,y)
  float* scores = nullptr;                              // [B,K]
  int* n_kpts = nullptr;                                // [B]
  float* desc = nullptr;                                // [B,K,256]
  // SP_RE
  float* re_kpts = nullptr;                             // [B,max_vio,2]
  int* re_n = nullptr;                                  // [B]
  float* re_desc = nullptr;                             // [B,max_vio,256]
  int H8 = 0, W8 = 0;
  GraphCache g_enc[2][2], g_det;                        // B = 1 launch sequences: [1- / 3-channel][frame buffer]; detection post-net
};

// ------------------------------------------------------------------------------------------------ kernels
// preprocess_kernel.cu:193-346 under the identity-affine precondition: u8 * (1/255.f); 3-ch: BGR->RGB, gray mix.
__global__ void k_gray(const uint8_t* __restrict__ img, float* __restrict__ out, int64_t npix, int ch) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float a = 1.0f / 255.0f;
  if (ch == 1) {
    out[i] = __fmul_rn((float)img[i], a);
  } else {
    const float b = (float)img[i * 3 + 0], g = (float)img[i * 3 + 1], r = (float)img[i * 3 + 2];
    const float c0 = __fmul_rn(r, a), c1 = __fmul_rn(g, a), c2 = __fmul_rn(b, a);
    // preprocess_kernel.cu:280 writes `0.299 * c0 + 0.587 * c1 + 0.114 * c2` with DOUBLE literals: the mix is evaluated
    // in fp64 and rounded to fp32 once (pinned against the compiled reference kernel, tests/test_refpre_gpu.py).
    out[i] = (float)__dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)c0), __dmul_rn(0.587, (double)c1)),
                              __dmul_rn(0.114, (double)c2));
  }
}

// conv1a: C_in = 1 is not an MMA shape -> CUDA cores.  8 threads per pixel, 8 output channels each: a warp writes
// 4 pixels x 128 B = 512 contiguous bytes of the NHWC fp16 output.  A block covers 128 x CONV1A_ROWS pixels so the
// 72 filter taps each thread needs are loaded once per block.
#define CONV1A_ROWS 8
// Channel-blocked variant for the halo-tile conv1b (conv_halo.cu): output [B][8][H][W][8]; warp = channel group,
// lane = pixel, so a warp writes 32 consecutive pixels x 16 B = 512 contiguous bytes.
__global__ void __launch_bounds__(256) k_conv1a_blocked(const float* __restrict__ gray, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __half* __restrict__ out,
                                                        int H, int W) {
  __shared__ float tile[CONV1A_ROWS + 2][130];
  const int b = blockIdx.z, y0 = blockIdx.y * CONV1A_ROWS, x0 = blockIdx.x * 128;
  const float* g = gray + (int64_t)b * H * W;
  for (int i = threadIdx.x; i < (CONV1A_ROWS + 2) * 130; i += 256) {
    const int r = i / 130, c = i - r * 130;
    const int yy = y0 + r - 1, xx = x0 + c - 1;
    tile[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? g[(int64_t)yy * W + xx] : 0.f;
  }
  const int cg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = w[(cg * 8 + c) * 9 + t];
  }
  __syncthreads();
  for (int ry = 0; ry < CONV1A_ROWS; ++ry) {
    const int y = y0 + ry;
    if (y >= H) break;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int xl = lane + it * 32;
      const int x = x0 + xl;
      if (x >= W) continue;
      float in[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) in[r * 3 + q] = tile[ry + r][xl + q];
      __align__(16) __half2 hv[4];
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        float a0 = br[c], a1 = br[c + 1];
#pragma unroll
        for (int t = 0; t < 9; ++t) { a0 = fmaf(wr[c][t], in[t], a0); a1 = fmaf(wr[c + 1][t], in[t], a1); }
        hv[c >> 1] = __floats2half2_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
      }
      *reinterpret_cast<uint4*>(out + ((((int64_t)b * 8 + cg) * H + y) * W + x) * 8) = *reinterpret_cast<uint4*>(hv);
    }
  }
}

__global__ void __launch_bounds__(256) k_conv1a(const float* __restrict__ gray, const float* __restrict__ w,
                                                const float* __restrict__ bias, __half* __restrict__ out, int H,
                                                int W) {
  __shared__ float tile[CONV1A_ROWS + 2][130];
  const int b = blockIdx.z, y0 = blockIdx.y * CONV1A_ROWS, x0 = blockIdx.x * 128;
  const float* g = gray + (int64_t)b * H * W;
  for (int i = threadIdx.x; i < (CONV1A_ROWS + 2) * 130; i += 256) {
    const int r = i / 130, c = i - r * 130;
    const int yy = y0 + r - 1, xx = x0 + c - 1;
    tile[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? g[(int64_t)yy * W + xx] : 0.f;
  }
  const int cg = threadIdx.x & 7, px = threadIdx.x >> 3;
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = w[(cg * 8 + c) * 9 + t];
  }
  __syncthreads();
  for (int ry = 0; ry < CONV1A_ROWS; ++ry) {
    const int y = y0 + ry;
    if (y >= H) break;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int xl = px + it * 32;
      const int x = x0 + xl;
      if (x >= W) continue;
      float in[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) in[r * 3 + q] = tile[ry + r][xl + q];
      __align__(16) __half2 hv[4];
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        float a0 = br[c], a1 = br[c + 1];
#pragma unroll
        for (int t = 0; t < 9; ++t) { a0 = fmaf(wr[c][t], in[t], a0); a1 = fmaf(wr[c + 1][t], in[t], a1); }
        hv[c >> 1] = __floats2half2_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
      }
      *reinterpret_cast<uint4*>(out + (((int64_t)b * H + y) * W + x) * 64 + cg * 8) = *reinterpret_cast<uint4*>(hv);
    }
  }
}

// softmax over 65 logits, drop the dustbin, scatter the 64 cell scores to the 8x8 pixel block
// (export/superpoint.py:174-177).  One warp per cell.
__global__ void k_softmax_d2s(const float* __restrict__ logits, int ld, float* __restrict__ smap, int ncells, int h8,
                              int w8) {
  const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncells) return;
  const float* l = logits + (int64_t)cell * ld;
  const float v0 = l[lane], v1 = l[lane + 32], v2 = l[64];
  float m = fmaxf(fmaxf(v0, v1), v2);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e0 = expf(v0 - m), e1 = expf(v1 - m), e2 = expf(v2 - m);
  float s = e0 + e1;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  s += e2;
  const int b = cell / (h8 * w8);
  const int rem = cell - b * h8 * w8;
  const int cy = rem / w8, cx = rem - cy * w8;
  const int W8 = w8 * 8;
  float* o = smap + (int64_t)b * h8 * 8 * W8 + (int64_t)cy * 8 * W8 + cx * 8;
  o[(lane >> 3) * W8 + (lane & 7)] = e0 / s;
  o[((lane + 32) >> 3) * W8 + (lane & 7)] = e1 / s;
}

// simple_nms (export/superpoint.py:52-69), radius 4, two refinement rounds, fused into one tile kernel.
// Dependency radius = 4 + 8 + 8 = 20, so each 64x32 output tile loads a 104x72 region.  Separable 9-tap max in
// shared memory; pixels outside the image are -inf and never "maxima" (== torch's implicit -inf padding).
#define NMS_HALO 20
// Tile configuration: TW x TH output pixels per CTA, (TW + 40) x (TH + 40) region in shared memory (10 bytes / pixel),
// ROW_STRIP / COL_STRIP outputs per thread in the row / column passes (must divide the region width / height).
//   NmsCfgS: 64 x 32 tiles, 104 x 72 region (3.65x the tile), 75 KB, 320 threads, three CTAs per SM          (r01)
//   NmsCfgL: 128 x 64 tiles, 168 x 104 region (2.13x the tile), 171 KB, 768 threads, one CTA per SM: 1.7x fewer
//            region pixels per frame through the ten separable passes
struct NmsCfgS { static constexpr int TW = 64, TH = 32, THREADS = 320, ROW_STRIP = 26, COL_STRIP = 24; static constexpr bool APRON = true; };
struct NmsCfgF { static constexpr int TW = 64, TH = 32, THREADS = 320, ROW_STRIP = 26, COL_STRIP = 24; static constexpr bool APRON = false; };   // r01: full region in every pass
struct NmsCfgL { static constexpr int TW = 128, TH = 64, THREADS = 768, ROW_STRIP = 24, COL_STRIP = 26; static constexpr bool APRON = false; };
template <class C> struct NmsDims {
  static constexpr int RW = C::TW + 2 * NMS_HALO, RH = C::TH + 2 * NMS_HALO, RN = RW * RH;
  static constexpr int SMEM = RN * (2 * 4 + 2);
  static_assert(RW % C::ROW_STRIP == 0 && RH % C::COL_STRIP == 0, "strips must tile the region");
  static_assert(RH * (RW / C::ROW_STRIP) <= C::THREADS && RW * (RH / C::COL_STRIP) <= C::THREADS, "one strip per thread");
};

// 9-tap running max over a strip of L outputs in registers: 4 max ops per output (pairwise doubling) and ONE shared
// memory load per input instead of 9.  Taps outside [0, n) are -inf (region edge == torch's implicit -inf padding).
// `ld(p)` produces input p, `st(p, m)` consumes the 9-tap max at p: the elementwise steps between NMS's max-pools
// (mask -> 0/1, suppressed scores, equality tests) ride inside the passes instead of being separate sweeps.
template <int L, class LD, class ST>
__device__ __forceinline__ void strip_max9(LD ld, ST st, int p0, int n) {
  float v[L + 8];
#pragma unroll
  for (int i = 0; i < L + 8; ++i) {
    const int p = p0 + i - 4;
    v[i] = (p >= 0 && p < n) ? ld(p) : -INFINITY;
  }
  float m2[L + 7];
#pragma unroll
  for (int i = 0; i < L + 7; ++i) m2[i] = fmaxf(v[i], v[i + 1]);
  float m4[L + 5];
#pragma unroll
  for (int i = 0; i < L + 5; ++i) m4[i] = fmaxf(m2[i], m2[i + 2]);
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const float m8 = fmaxf(m4[i], m4[i + 4]);
    if (p0 + i < n) st(p0 + i, fmaxf(m8, v[i + 8]));
  }
}


// 9x9 max-pool of the region: T1 = rowmax(in(i)); out(i, colmax(T1)); both functors take the linear region index.
template <class C, class IN, class OUT>
__device__ __forceinline__ void maxpool9(IN in, OUT out, float* T1) {
  constexpr int NMS_RW = NmsDims<C>::RW, NMS_RH = NmsDims<C>::RH, NMS_ROW_STRIP = C::ROW_STRIP, NMS_COL_STRIP = C::COL_STRIP;
  {
    const int t = threadIdx.x;
    if (t < NMS_RH * (NMS_RW / NMS_ROW_STRIP)) {
      const int row = t / (NMS_RW / NMS_ROW_STRIP), sp = t - row * (NMS_RW / NMS_ROW_STRIP);
      const int base = row * NMS_RW;
      strip_max9<NMS_ROW_STRIP>([&](int p) { return in(base + p); }, [&](int p, float m) { T1[base + p] = m; },
                                sp * NMS_ROW_STRIP, NMS_RW);
    }
  }
  __syncthreads();
  {
    const int t = threadIdx.x;
    if (t < NMS_RW * (NMS_RH / NMS_COL_STRIP)) {
      const int sp = t / NMS_RW, col = t - sp * NMS_RW;
      strip_max9<NMS_COL_STRIP>([&](int p) { return T1[p * NMS_RW + col]; },
                                [&](int p, float m) { out(p * NMS_RW + col, m); }, sp * NMS_COL_STRIP, NMS_RH);
    }
  }
  __syncthreads();
}

// The same 9x9 max-pool restricted to the part of the region that can still influence the tile (NmsCfgS only, r02).
// The five pools of simple_nms consume 4 pixels of apron each: pool k only has to be right on the region shrunk by
// a = 4k pixels per side ([a, RH - a) x [a, RW - a)), and its row pass only on the rows the column pass reads
// ([a - 4, RH - a + 4)).  Summed over the ten passes that is 42 k region pixels per tile instead of 74 k.  LR / LC are the
// strip lengths of the row / column pass (chosen per pool so that the 320 threads stay busy; odd row-strip lengths keep
// the strips of neighbouring lanes off the same banks).  Values outside the shrinking rectangle are stale but never read
// by anything that reaches the tile.
template <int L, class LD, class ST>
__device__ __forceinline__ void strip_max9_b(LD ld, ST st, int p0, int n, int hi) {
  float v[L + 8];
#pragma unroll
  for (int i = 0; i < L + 8; ++i) {
    const int p = p0 + i - 4;
    v[i] = (p >= 0 && p < n) ? ld(p) : -INFINITY;
  }
  float m2[L + 7];
#pragma unroll
  for (int i = 0; i < L + 7; ++i) m2[i] = fmaxf(v[i], v[i + 1]);
  float m4[L + 5];
#pragma unroll
  for (int i = 0; i < L + 5; ++i) m4[i] = fmaxf(m2[i], m2[i + 2]);
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const float m8 = fmaxf(m4[i], m4[i + 4]);
    if (p0 + i < hi) st(p0 + i, fmaxf(m8, v[i + 8]));
  }
}
template <class C, int A, int LR, int LC, class IN, class OUT>
__device__ __forceinline__ void maxpool9_apron(IN in, OUT out, float* T1) {
  constexpr int RW = NmsDims<C>::RW, RH = NmsDims<C>::RH;
  constexpr int W = RW - 2 * A, H = RH - 2 * A;              // output rectangle
  constexpr int NS = (W + LR - 1) / LR, ROWS = H + 8;        // row pass: strips per row, rows [A - 4, RH - A + 4)
  constexpr int NC = (H + LC - 1) / LC;                      // column pass: strips per column
  static_assert(A >= 4 && ROWS * NS <= C::THREADS && W * NC <= C::THREADS, "one strip per thread");
  {
    const int t = threadIdx.x;
    if (t < ROWS * NS) {
      const int row = A - 4 + t / NS, sp = t % NS;
      const int base = row * RW;
      strip_max9_b<LR>([&](int p) { return in(base + p); }, [&](int p, float m) { T1[base + p] = m; }, A + sp * LR, RW,
                       RW - A);
    }
  }
  __syncthreads();
  {
    const int t = threadIdx.x;
    if (t < W * NC) {
      const int sp = t / W, col = A + t % W;
      strip_max9_b<LC>([&](int p) { return T1[p * RW + col]; }, [&](int p, float m) { out(p * RW + col, m); },
                       A + sp * LC, RH, RH - A);
    }
  }
  __syncthreads();
}

template <class C>
__global__ void __launch_bounds__(C::THREADS) k_nms_select(const float* __restrict__ smap, float* __restrict__ nms_out,
                                                    unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt,
                                                    int H8, int W8, int border, float thresh) {
  constexpr int NMS_TW = C::TW, NMS_TH = C::TH, NMS_RW = NmsDims<C>::RW, NMS_RN = NmsDims<C>::RN;
  extern __shared__ float sm[];
  float* S = sm;
  float* T1 = S + NMS_RN;
  uint8_t* MM = reinterpret_cast<uint8_t*>(T1 + NMS_RN);
  uint8_t* SUPP = MM + NMS_RN;
  const int b = blockIdx.z;
  const int gx0 = blockIdx.x * NMS_TW - NMS_HALO, gy0 = blockIdx.y * NMS_TH - NMS_HALO;
  const float* src = smap + (int64_t)b * H8 * W8;
  const float NEG = -INFINITY;
  for (int i = threadIdx.x; i < NMS_RN; i += blockDim.x) {
    const int y = i / NMS_RW, x = i - y * NMS_RW;
    const int gy = gy0 + y, gx = gx0 + x;
    const bool in = gy >= 0 && gy < H8 && gx >= 0 && gx < W8;
    S[i] = in ? src[(int64_t)gy * W8 + gx] : NEG;
  }
  __syncthreads();
  // export/superpoint.py:52-66 simple_nms: max_mask = scores == max_pool(scores); two rounds of
  // supp = max_pool(mask) > 0; supp_scores = where(supp, 0, scores); mask |= (supp_scores == max_pool(supp_scores)) & ~supp
  auto in_s = [&](int i) { return S[i]; };
  auto out_m0 = [&](int i, float m) {
    const float v = S[i];
    MM[i] = (v != NEG) && (v == m);
  };
  auto in_m = [&](int i) { return MM[i] ? 1.f : 0.f; };
  auto out_supp = [&](int i, float m) { SUPP[i] = m > 0.f; };
  auto in_ss = [&](int i) {
    const float v = S[i];
    return (v == NEG) ? NEG : (SUPP[i] ? 0.f : v);
  };
  auto out_m = [&](int i, float m) {
    const float v = S[i];
    if (v != NEG && !SUPP[i] && v == m) MM[i] = 1;
  };
  if constexpr (C::APRON) {
    // shrinking rectangles: each pool is evaluated only where it can still reach the tile (see maxpool9_apron)
    maxpool9_apron<C, 4, 25, 22>(in_s, out_m0, T1);
    maxpool9_apron<C, 8, 18, 19>(in_m, out_supp, T1);
    maxpool9_apron<C, 12, 17, 12>(in_ss, out_m, T1);
    maxpool9_apron<C, 16, 13, 10>(in_m, out_supp, T1);
    maxpool9_apron<C, 20, 9, 8>(in_ss, out_m, T1);
  } else {
    maxpool9<C>(in_s, out_m0, T1);
    for (int round = 0; round < 2; ++round) {
      maxpool9<C>(in_m, out_supp, T1);
      maxpool9<C>(in_ss, out_m, T1);
    }
  }
  for (int i = threadIdx.x; i < NMS_TW * NMS_TH; i += blockDim.x) {
    const int ty = i / NMS_TW, tx = i - ty * NMS_TW;
    const int gy = blockIdx.y * NMS_TH + ty, gx = blockIdx.x * NMS_TW + tx;
    if (gy >= H8 || gx >= W8) continue;
    const int ri = (ty + NMS_HALO) * NMS_RW + tx + NMS_HALO;
    const float v = MM[ri] ? S[ri] : 0.f;
    const int64_t lin = (int64_t)gy * W8 + gx;
    if (nms_out) nms_out[(int64_t)b * H8 * W8 + lin] = v;
    // export/superpoint.py:183-195: border -> -1, strict threshold
    const bool inb = gy >= border && gy < H8 - border && gx >= border && gx < W8 - border;
    if (inb && v > thresh) {
      const int slot = atomicAdd(&cand_cnt[b], 1);
      // key: score bits (positive floats order as unsigned) then ~index => descending key = score desc, index asc
      cand[(int64_t)b * H8 * W8 + slot] =
          ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)lin);
    }
  }
}

// top-k selection per frame: exact k-th key by 8-bit radix select, then bitonic sort of the <= 1024 survivors.
// <= k candidates: all kept, in row-major order (export/superpoint.py:76-77 returns them unsorted).
__global__ void __launch_bounds__(1024) k_topk(const unsigned long long* __restrict__ cand,
                                               const int* __restrict__ cand_cnt, int cap, int K, int W8,
                                               int* __restrict__ kpts, float* __restrict__ kpts_f,
                                               float* __restrict__ scores, int* __restrict__ n_out) {
  __shared__ unsigned long long sel[1024];
  __shared__ int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_k, s_nsel;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(cand_cnt[b], cap);
  const unsigned long long* c = cand + (int64_t)b * cap;
  const bool all = n <= K;
  int nsel;
  if (all) {
    nsel = n;
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long k = c[i];
      sel[i] = (k << 32) | (k >> 32);   // sort by ~index descending == index ascending
    }
    __syncthreads();
  } else {
    if (tid == 0) { s_prefix = 0ull; s_k = K; s_nsel = 0; }
    __syncthreads();
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long k = c[i];
        if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(int)((k >> shift) & 255ull)], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, kk = s_k;
        for (int d = 255; d >= 0; --d) {
          if (cum + hist[d] >= kk) { s_k = kk - cum; s_prefix = (prefix << 8) | (unsigned long long)d; break; }
          cum += hist[d];
        }
      }
      __syncthreads();
    }
    const unsigned long long T = s_prefix;
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long k = c[i];
      if (k >= T) { const int p = atomicAdd(&s_nsel, 1); if (p < 1024) sel[p] = k; }
    }
    __syncthreads();
    nsel = min(s_nsel, K);
  }
  int P = 1;
  while (P < nsel) P <<= 1;
  for (int i = nsel + tid; i < P; i += blockDim.x) sel[i] = 0ull;
  __syncthreads();
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sel[i], d = sel[ixj];
          const bool desc = (i & k2) == 0;
          if (desc ? (a < d) : (a > d)) { sel[i] = d; sel[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < nsel; i += blockDim.x) {
    unsigned long long k = sel[i];
    if (all) k = (k << 32) | (k >> 32);
    const unsigned lin = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
    const float sc = __uint_as_float((unsigned)(k >> 32));
    const int y = (int)(lin / (unsigned)W8), x = (int)(lin - (unsigned)y * (unsigned)W8);
    kpts[((int64_t)b * K + i) * 2 + 0] = x;
    kpts[((int64_t)b * K + i) * 2 + 1] = y;
    kpts_f[((int64_t)b * K + i) * 2 + 0] = (float)x;
    kpts_f[((int64_t)b * K + i) * 2 + 1] = (float)y;
    scores[(int64_t)b * K + i] = sc;
  }
  if (tid == 0) n_out[b] = nsel;
}

// sample_descriptors (export/superpoint.py:83-98): per-pixel L2-normalised dense map, bilinear grid_sample with
// align_corners=True and zero padding, L2-normalise again.  One warp per keypoint, 8 channels per lane.
__global__ void k_sample_desc(const float* __restrict__ dmap, int h8, int w8, const float* __restrict__ kpts,
                              const int* __restrict__ n_kpts, int cap, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int kp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (kp >= n_kpts[b]) return;
  const float kx = kpts[((int64_t)b * cap + kp) * 2 + 0], ky = kpts[((int64_t)b * cap + kp) * 2 + 1];
  // keypoints - s/2 + 0.5; / (w*s - s/2 - 0.5); *2 - 1; grid_sample unnormalise ((g+1)/2)*(size-1)
  const float ux = __fdiv_rn(kx - 3.5f, (float)(w8 * 8) - 4.5f), uy = __fdiv_rn(ky - 3.5f, (float)(h8 * 8) - 4.5f);
  const float gx = __fsub_rn(__fmul_rn(ux, 2.f), 1.f), gy = __fsub_rn(__fmul_rn(uy, 2.f), 1.f);
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w8 - 1));
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h8 - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
  const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};   // nw, ne, sw, se
  const int cx[4] = {x0, x1, x0, x1}, cy[4] = {y0, y0, y1, y1};
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (cx[c] < 0 || cx[c] >= w8 || cy[c] < 0 || cy[c] >= h8) continue;   // warp-uniform
    const float4* p =
        reinterpret_cast<const float4*>(dmap + (((int64_t)b * h8 + cy[c]) * w8 + cx[c]) * 256 + lane * 8);
    const float4 a = __ldg(p), d = __ldg(p + 1);
    float v[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = wgt[c] / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], inv, acc[j]);
  }
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss += acc[j] * acc[j];
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  float4* o = reinterpret_cast<float4*>(out + ((int64_t)b * cap + kp) * 256 + lane * 8);
  o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}

// ------------------------------------------------------------------------------------------------ host
static int get_conv(Engine* e, const std::string& name, int cout, int cin, std::vector<float>* w_packed,
                    std::vector<float>* b) {
  const HostTensor* w = e->weight("sp." + name + ".weight");
  const HostTensor* bb = e->weight("sp." + name + ".bias");
  if (!w || !bb || w->dims.size() != 4 || w->dims[0] != cout || w->dims[1] != cin || w->dims[2] != 3 ||
      w->dims[3] != 3 || bb->numel() != cout) {
    set_error("SuperPoint weights: missing or mis-shaped tensor sp." + name);
    return DV_ERR_WEIGHTS;
  }
  // torch [cout,cin,3,3] -> [cout, (r*3+s)*cin + c]  (K-major rows for the implicit GEMM)
  w_packed->resize((size_t)cout * 9 * cin);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 9; ++t) (*w_packed)[((size_t)o * 9 + t) * cin + c] = w->data[((size_t)o * cin + c) * 9 + t];
  *b = bb->data;
  return DV_OK;
}

int sp_init(Engine* e) {
  SpNet* s = new SpNet();
  e->sp = s;
  const int B = e->B, H = e->H, W = e->W, h8 = e->h8, w8 = e->w8;
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  s->H8 = h8 * 8; s->W8 = w8 * 8;
  // ---- weights
  {
    const HostTensor* w = e->weight("sp.conv1a.weight");
    const HostTensor* b = e->weight("sp.conv1a.bias");
    if (!w || !b || w->numel() != 64 * 9 || b->numel() != 64) { set_error("SuperPoint weights: sp.conv1a"); return DV_ERR_WEIGHTS; }
    DV_TRY(e->upload_f32(w->data, &s->w1a));
    DV_TRY(e->upload_f32(b->data, &s->b1a));
  }
  struct L { const char* name; int cin, cout; };
  const L layers[7] = {{"conv1b", 64, 64}, {"conv2a", 64, 64}, {"conv2b", 64, 64}, {"conv3a", 64, 128},
                       {"conv3b", 128, 128}, {"conv4a", 128, 128}, {"conv4b", 128, 128}};
  for (int i = 0; i < 7; ++i) {
    std::vector<float> wp, b;
    DV_TRY(get_conv(e, layers[i].name, layers[i].cout, layers[i].cin, &wp, &b));
    DV_TRY(e->upload_f16(wp, &s->w[i]));
    DV_TRY(e->upload_f32(b, &s->bias[i]));
  }
  {  // convPa ++ convDa share their input: one N=512 implicit GEMM
    std::vector<float> wp, bp, wd, bd;
    DV_TRY(get_conv(e, "convPa", 256, 128, &wp, &bp));
    DV_TRY(get_conv(e, "convDa", 256, 128, &wd, &bd));
    wp.insert(wp.end(), wd.begin(), wd.end());
    bp.insert(bp.end(), bd.begin(), bd.end());
    DV_TRY(e->upload_f16(wp, &s->w[7]));
    DV_TRY(e->upload_f32(bp, &s->bias[7]));
  }
  {
    const HostTensor* w = e->weight("sp.convPb.weight");
    const HostTensor* b = e->weight("sp.convPb.bias");
    const HostTensor* wd = e->weight("sp.convDb.weight");
    const HostTensor* bd = e->weight("sp.convDb.bias");
    if (!w || !b || !wd || !bd || w->numel() != 65 * 256 || b->numel() != 65 || wd->numel() != 256 * 256 ||
        bd->numel() != 256) { set_error("SuperPoint weights: sp.convPb / sp.convDb"); return DV_ERR_WEIGHTS; }
    std::vector<float> wp(80 * 256, 0.f), bp(80, 0.f);       // N padded 65 -> 80 (zero rows)
    std::copy(w->data.begin(), w->data.end(), wp.begin());
    std::copy(b->data.begin(), b->data.end(), bp.begin());
    DV_TRY(e->upload_f16(wp, &s->wPb));
    DV_TRY(e->upload_f32(bp, &s->bPb));
    DV_TRY(e->upload_f16(wd->data, &s->wDb));
    DV_TRY(e->upload_f32(bd->data, &s->bDb));
  }
  // ---- activations
  const size_t P1 = (size_t)B * H * W, P2 = (size_t)B * H2 * W2, P4 = (size_t)B * H4 * W4, P8 = (size_t)B * h8 * w8;
  DV_TRY(e->alloc(&s->gray, P1));
  DV_TRY(e->alloc(&s->a1a, P1 * 64));
  DV_TRY(e->alloc(&s->a1b, P2 * 64));
  DV_TRY(e->alloc(&s->a2a, P2 * 64));
  DV_TRY(e->alloc(&s->a2b, P4 * 64));
  DV_TRY(e->alloc(&s->a3a, P4 * 128));
  DV_TRY(e->alloc(&s->a3b, P8 * 128));
  DV_TRY(e->alloc(&s->a4a, P8 * 128));
  DV_TRY(e->alloc(&s->a4b, P8 * 128));
  DV_TRY(e->alloc(&s->aPD, P8 * 512));
  DV_TRY(e->alloc(&s->logits, P8 * 80));
  DV_TRY(e->alloc(&s->dmap, P8 * 256));
  const size_t PS = (size_t)B * s->H8 * s->W8;
  const int K = e->cfg.max_kpts, V = e->cfg.max_vio;
  DV_TRY(e->alloc(&s->smap, PS));
  DV_TRY(e->alloc(&s->nms, PS));
  DV_TRY(e->alloc(&s->cand, PS));
  DV_TRY(e->alloc(&s->cand_cnt, (size_t)B));
  DV_TRY(e->alloc(&s->kpts, (size_t)B * K * 2));
  DV_TRY(e->alloc(&s->kpts_f, (size_t)B * K * 2));
  DV_TRY(e->alloc(&s->scores, (size_t)B * K));
  DV_TRY(e->alloc(&s->n_kpts, (size_t)B));
  DV_TRY(e->alloc(&s->desc, (size_t)B * K * 256));
  DV_TRY(e->alloc(&s->re_kpts, (size_t)B * V * 2));
  DV_TRY(e->alloc(&s->re_n, (size_t)B));
  DV_TRY(e->alloc(&s->re_desc, (size_t)B * V * 256));
  // ---- plans
  auto ep16 = [](__half* out, int ld, const float* bias, int relu, int pool) {
    EpiParams ep; ep.out16 = out; ep.ld16 = ld; ep.bias = bias; ep.relu = relu; ep.pool = pool; return ep;
  };
  {
    const char* env = getenv("DV_SP_HALO");
    s->use_halo = !(env && env[0] == '0');
  }
  // halo path: conv1a -> (blocked) -> conv1b+pool -> (blocked) -> conv2a -> (blocked) -> conv2b+pool -> NHWC
  DV_TRY(plan_conv3x3_halo64(&s->h1b, s->a1a, B, H, W, s->w[0], s->bias[0], s->a1b, 1, 1, 1));
  {
    const char* env = getenv("DV_SP_FUSE1A");
    s->fuse1a = (env && env[0] == '1');            // 1: CUDA-core conv1a in conv1b's producer (slower; kept for A/B)
    s->fuse1a_tc = (!env || env[0] == '2') ? 1 : 0;   // default: conv1a on the tensor cores inside conv1b; 0: separate kernel
    if (s->fuse1a) { s->h1b.gray = s->gray; s->h1b.w1a = s->w1a; s->h1b.b1a = s->b1a; }   // conv1a inside conv1b's producer
  }
  DV_TRY(plan_conv3x3_halo64(&s->h2a, s->a1b, B, H2, W2, s->w[1], s->bias[1], s->a2a, 1, 1, 0));
  {
    const char* env = getenv("DV_SP_HALO128");           // 0: tap-per-TMA implicit GEMM for the 128-channel layers (A/B)
    s->use_halo128 = s->use_halo && !(env && env[0] == '0');
  }
  // conv2b+pool -> (blocked when the 128-channel layers run on the halo kernel, else NHWC)
  DV_TRY(plan_conv3x3_halo64(&s->h2b, s->a2a, B, H2, W2, s->w[2], s->bias[2], s->a2b, s->use_halo128 ? 1 : 0, 1, 1));
  // conv3a -> conv3b+pool -> conv4a -> conv4b stay channel-blocked; convPa ++ convDa writes NHWC for the 1x1 heads
  DV_TRY(plan_conv3x3_halo128(&s->h3a, s->a2b, B, H4, W4, 64, s->w[3], 128, s->bias[3], s->a3a, 1, 1, 0));
  DV_TRY(plan_conv3x3_halo128(&s->h3b, s->a3a, B, H4, W4, 128, s->w[4], 128, s->bias[4], s->a3b, 1, 1, 1));
  DV_TRY(plan_conv3x3_halo128(&s->h4a, s->a3b, B, h8, w8, 128, s->w[5], 128, s->bias[5], s->a4a, 1, 1, 0));
  DV_TRY(plan_conv3x3_halo128(&s->h4b, s->a4a, B, h8, w8, 128, s->w[6], 128, s->bias[6], s->a4b, 1, 1, 0));
  DV_TRY(plan_conv3x3_halo128(&s->hPD, s->a4b, B, h8, w8, 128, s->w[7], 512, s->bias[7], s->aPD, 0, 1, 0));
  DV_TRY(plan_conv3x3(&s->p1b, s->a1a, B, H, W, 64, s->w[0], 64, ep16(s->a1b, 64, s->bias[0], 1, 1)));
  DV_TRY(plan_conv3x3(&s->p2a, s->a1b, B, H2, W2, 64, s->w[1], 64, ep16(s->a2a, 64, s->bias[1], 1, 0)));
  DV_TRY(plan_conv3x3(&s->p2b, s->a2a, B, H2, W2, 64, s->w[2], 64, ep16(s->a2b, 64, s->bias[2], 1, 1)));
  DV_TRY(plan_conv3x3(&s->p3a, s->a2b, B, H4, W4, 64, s->w[3], 128, ep16(s->a3a, 128, s->bias[3], 1, 0)));
  DV_TRY(plan_conv3x3(&s->p3b, s->a3a, B, H4, W4, 128, s->w[4], 128, ep16(s->a3b, 128, s->bias[4], 1, 1)));
  DV_TRY(plan_conv3x3(&s->p4a, s->a3b, B, h8, w8, 128, s->w[5], 128, ep16(s->a4a, 128, s->bias[5], 1, 0)));
  DV_TRY(plan_conv3x3(&s->p4b, s->a4a, B, h8, w8, 128, s->w[6], 128, ep16(s->a4b, 128, s->bias[6], 1, 0)));
  DV_TRY(plan_conv3x3(&s->pPD, s->a4b, B, h8, w8, 128, s->w[7], 512, ep16(s->aPD, 512, s->bias[7], 1, 0)));
  {
    EpiParams ep; ep.out32 = s->logits; ep.ld32 = 80; ep.bias = s->bPb;
    DV_TRY(plan_gemm(&s->pPb, s->aPD, 512, (int)P8, s->wPb, 256, 80, 256, ep));
    EpiParams ed; ed.out32 = s->dmap; ed.ld32 = 256; ed.bias = s->bDb;
    DV_TRY(plan_gemm(&s->pDb, s->aPD + 256, 512, (int)P8, s->wDb, 256, 256, 256, ed));
  }
  DV_CUDA_OK(cudaFuncSetAttribute(k_nms_select<NmsCfgS>, cudaFuncAttributeMaxDynamicSharedMemorySize, NmsDims<NmsCfgS>::SMEM));
  DV_CUDA_OK(cudaFuncSetAttribute(k_nms_select<NmsCfgF>, cudaFuncAttributeMaxDynamicSharedMemorySize, NmsDims<NmsCfgF>::SMEM));
  DV_CUDA_OK(cudaFuncSetAttribute(k_nms_select<NmsCfgL>, cudaFuncAttributeMaxDynamicSharedMemorySize, NmsDims<NmsCfgL>::SMEM));
  { const char* env = getenv("DV_NMS_TILE"); s->nms_large = env ? env[0] == 'L' : DV_NMS_DEFAULT_LARGE; s->nms_full = env && env[0] == 'F'; }
  // debug views
  e->dbg["gray"] = {s->gray, (int64_t)H * W, 0};
  e->dbg[s->use_halo ? "conv1a_blocked" : "conv1a"] = {s->a1a, (int64_t)H * W * 64, 1};
  e->dbg[s->use_halo ? "conv1b_pool_blocked" : "conv1b_pool"] = {s->a1b, (int64_t)H2 * W2 * 64, 1};
  e->dbg[s->use_halo ? "conv2a_blocked" : "conv2a"] = {s->a2a, (int64_t)H2 * W2 * 64, 1};
  {
    const char* sfx = s->use_halo128 ? "_blocked" : "";   // [C/8][H][W][8] on the halo path, NHWC otherwise
    e->dbg[std::string("conv2b_pool") + sfx] = {s->a2b, (int64_t)H4 * W4 * 64, 1};
    e->dbg[std::string("conv3a") + sfx] = {s->a3a, (int64_t)H4 * W4 * 128, 1};
    e->dbg[std::string("conv3b_pool") + sfx] = {s->a3b, (int64_t)h8 * w8 * 128, 1};
    e->dbg[std::string("conv4a") + sfx] = {s->a4a, (int64_t)h8 * w8 * 128, 1};
    e->dbg[std::string("conv4b") + sfx] = {s->a4b, (int64_t)h8 * w8 * 128, 1};
  }
  e->dbg["convPD"] = {s->aPD, (int64_t)h8 * w8 * 512, 1};
  e->dbg["logits"] = {s->logits, (int64_t)h8 * w8 * 80, 0};
  e->dbg["dmap"] = {s->dmap, (int64_t)h8 * w8 * 256, 0};
  e->dbg["score_map"] = {s->smap, (int64_t)s->H8 * s->W8, 0};
  e->dbg["nms"] = {s->nms, (int64_t)s->H8 * s->W8, 0};
  return DV_OK;
}

void sp_free(Engine* e) {
  delete e->sp;
  e->sp = nullptr;
}

static int sp_enqueue_encoder(Engine* e, int b, bool release_inline);

int sp_run_encoder(Engine* e, int b) {
  SpNet* s = e->sp;
  if (!s) { set_error("SuperPoint not initialised (engine created without weights)"); return DV_ERR_INVALID; }
  StageScope sc(e, ST_SP_CONV);
  e->image_acquire();
  // host-side state (graph replays do not run the enqueue code): is the fp32 gray frame produced on this path?
  s->gray_valid = !(s->use_halo && s->fuse1a_tc && e->img_ch == 1);
  if (b == 1) {
    // per-keyframe latency path: 13 launches replayed as one CUDA graph; the frame-buffer events stay outside of it.
    // The frame pointer is a kernel argument baked into the graph and frames are double-buffered: one graph per buffer.
    const int rc = run_graphed(e, s->g_enc[e->img_ch == 3 ? 1 : 0][e->img_idx], [&]() { return sp_enqueue_encoder(e, 1, false); });
    e->image_release();
    return rc;
  }
  return sp_enqueue_encoder(e, b, true);
}

static int sp_enqueue_encoder(Engine* e, int b, bool release_inline) {
  SpNet* s = e->sp;
  const int H = e->H, W = e->W;
  const int64_t npix = (int64_t)b * H * W;
  // 1-channel frames: conv1a runs on the tensor cores inside conv1b, straight from the u8 frame (3-channel frames
  // go through the fp64 gray mix of k_gray and the separate conv1a kernel); the fp32 gray frame is then not needed at
  // all (92 MB per 64 frames) and is only produced on demand for dv_dbg_read("gray") (sp_dbg_refresh)
  const bool tc1a_path = s->use_halo && s->fuse1a_tc && e->img_ch == 1;
  if (!tc1a_path) k_gray<<<(unsigned)cdiv64(npix, 256), 256, 0, e->st>>>(e->d_img, s->gray, npix, e->img_ch);
  if (s->use_halo) {
    const bool tc1a = tc1a_path;
    if (!s->fuse1a && !tc1a) {
      k_conv1a_blocked<<<dim3(cdiv(W, 128), cdiv(H, CONV1A_ROWS), b), 256, 0, e->st>>>(s->gray, s->w1a, s->b1a, s->a1a, H, W);
      DV_CUDA_OK(cudaGetLastError());
    }
    {
      ProbeScope pr(e);
      if (tc1a) {
        HaloPlan hp = s->h1b;
        hp.img8 = e->d_img; hp.w1a = s->w1a; hp.b1a = s->b1a;
        DV_TRY(launch_conv_halo64(hp, b, e->st));
      } else {
        DV_TRY(launch_conv_halo64(s->h1b, b, e->st));
      }
    }
    DV_TRY(launch_conv_halo64(s->h2a, b, e->st));
    DV_TRY(launch_conv_halo64(s->h2b, b, e->st));
  } else {
    k_conv1a<<<dim3(cdiv(W, 128), cdiv(H, CONV1A_ROWS), b), 256, 0, e->st>>>(s->gray, s->w1a, s->b1a, s->a1a, H, W);
    DV_CUDA_OK(cudaGetLastError());
    {
      ProbeScope pr(e);
      DV_TRY(launch_gemm(s->p1b, b, e->st));
    }
    DV_TRY(launch_gemm(s->p2a, b, e->st));
    DV_TRY(launch_gemm(s->p2b, b, e->st));
  }
  if (release_inline) e->image_release();     // k_gray / the fused conv1a+conv1b kernel were the encoder's only readers of the u8 frames
  if (s->use_halo128) {
    DV_TRY(launch_conv_halo128(s->h3a, b, e->st));
    DV_TRY(launch_conv_halo128(s->h3b, b, e->st));
    DV_TRY(launch_conv_halo128(s->h4a, b, e->st));
    DV_TRY(launch_conv_halo128(s->h4b, b, e->st));
    DV_TRY(launch_conv_halo128(s->hPD, b, e->st));
  } else {
    DV_TRY(launch_gemm(s->p3a, b, e->st));
    DV_TRY(launch_gemm(s->p3b, b, e->st));
    DV_TRY(launch_gemm(s->p4a, b, e->st));
    DV_TRY(launch_gemm(s->p4b, b, e->st));
    DV_TRY(launch_gemm(s->pPD, b, e->st));
  }
  const int rows = b * e->h8 * e->w8;
  DV_TRY(launch_gemm(s->pPb, rows, e->st));
  DV_TRY(launch_gemm(s->pDb, rows, e->st));
  DV_LAUNCHED(e, 12);
  return DV_OK;
}

static int run_post(Engine* e, int b, const float* smap, float* nms_out) {
  SpNet* s = e->sp;
  const int H8 = s->H8, W8 = s->W8, K = e->cfg.max_kpts;
  DV_CUDA_OK(cudaMemsetAsync(s->cand_cnt, 0, sizeof(int) * b, e->st));
  if (s->nms_large)
    k_nms_select<NmsCfgL><<<dim3(cdiv(W8, NmsCfgL::TW), cdiv(H8, NmsCfgL::TH), b), NmsCfgL::THREADS, NmsDims<NmsCfgL>::SMEM, e->st>>>(
        smap, nms_out, s->cand, s->cand_cnt, H8, W8, e->cfg.border, e->cfg.det_thresh);
  else if (s->nms_full)
    k_nms_select<NmsCfgF><<<dim3(cdiv(W8, NmsCfgF::TW), cdiv(H8, NmsCfgF::TH), b), NmsCfgF::THREADS, NmsDims<NmsCfgF>::SMEM, e->st>>>(
        smap, nms_out, s->cand, s->cand_cnt, H8, W8, e->cfg.border, e->cfg.det_thresh);
  else
    k_nms_select<NmsCfgS><<<dim3(cdiv(W8, NmsCfgS::TW), cdiv(H8, NmsCfgS::TH), b), NmsCfgS::THREADS, NmsDims<NmsCfgS>::SMEM, e->st>>>(
        smap, nms_out, s->cand, s->cand_cnt, H8, W8, e->cfg.border, e->cfg.det_thresh);
  k_topk<<<b, 1024, 0, e->st>>>(s->cand, s->cand_cnt, H8 * W8, K, W8, s->kpts, s->kpts_f, s->scores, s->n_kpts);
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, 2);
  return DV_OK;
}

// dv_dbg_read("gray"): the pre-processed frame is skipped on the tensor-core conv1a path; produce it on demand
int sp_dbg_refresh(Engine* e, const char* name) {
  SpNet* s = e->sp;
  if (!s || strcmp(name, "gray") != 0 || s->gray_valid || e->cur_b <= 0) return DV_OK;
  const int64_t npix = (int64_t)e->cur_b * e->H * e->W;
  k_gray<<<(unsigned)cdiv64(npix, 256), 256, 0, e->st>>>(e->d_img, s->gray, npix, e->img_ch);
  DV_CUDA_OK(cudaGetLastError());
  s->gray_valid = true;
  return DV_OK;
}

static int sp_enqueue_detect(Engine* e, int b);

int sp_run_detect(Engine* e, int b) {
  SpNet* s = e->sp;
  StageScope sc(e, ST_SP_POST);
  if (b == 1) return run_graphed(e, s->g_det, [&]() { return sp_enqueue_detect(e, 1); });
  return sp_enqueue_detect(e, b);
}

static int sp_enqueue_detect(Engine* e, int b) {
  SpNet* s = e->sp;
  const int ncells = b * e->h8 * e->w8;
  k_softmax_d2s<<<cdiv(ncells, 8), 256, 0, e->st>>>(s->logits, 80, s->smap, ncells, e->h8, e->w8);
  DV_TRY(run_post(e, b, s->smap, s->nms));
  const int K = e->cfg.max_kpts;
  k_sample_desc<<<dim3(cdiv(K, 8), b), 256, 0, e->st>>>(s->dmap, e->h8, e->w8, s->kpts_f, s->n_kpts, K, s->desc);
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, 2);
  return DV_OK;
}

int sp_run_describe(Engine* e, int b, const float* d_kpts, const int* d_n, int cap, float* d_desc) {
  SpNet* s = e->sp;
  StageScope sc(e, ST_SP_POST);
  k_sample_desc<<<dim3(cdiv(cap, 8), b), 256, 0, e->st>>>(s->dmap, e->h8, e->w8, d_kpts, d_n, cap, d_desc);
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, 1);
  return DV_OK;
}

// accessors used by engine.cpp / store.cu
void sp_device_results(Engine* e, int** kpts, float** kpts_f, float** scores, int** n, float** desc, float** re_kpts,
                       int** re_n, float** re_desc) {
  SpNet* s = e->sp;
  if (kpts) *kpts = s->kpts;
  if (kpts_f) *kpts_f = s->kpts_f;
  if (scores) *scores = s->scores;
  if (n) *n = s->n_kpts;
  if (desc) *desc = s->desc;
  if (re_kpts) *re_kpts = s->re_kpts;
  if (re_n) *re_n = s->re_n;
  if (re_desc) *re_desc = s->re_desc;
}

int sp_nms_select_dbg(Engine* e, const float* h_smap, int h8x8, int w8x8, float* h_nms, int32_t* kp, float* sc,
                      int32_t* n) {
  SpNet* s = e->sp;
  if (!s) { set_error("dv_dbg_nms_select needs an engine created with weights"); return DV_ERR_INVALID; }
  if (h8x8 != s->H8 || w8x8 != s->W8) { set_error("dv_dbg_nms_select: score map must be [8*(H/8), 8*(W/8)]"); return DV_ERR_INVALID; }
  const size_t np = (size_t)h8x8 * w8x8;
  DV_CUDA_OK(cudaMemcpyAsync(s->smap, h_smap, np * 4, cudaMemcpyHostToDevice, e->st));
  DV_TRY(run_post(e, 1, s->smap, s->nms));
  int cnt = 0;
  DV_CUDA_OK(cudaMemcpyAsync(&cnt, s->n_kpts, 4, cudaMemcpyDeviceToHost, e->st));
  if (h_nms) DV_CUDA_OK(cudaMemcpyAsync(h_nms, s->nms, np * 4, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  DV_CUDA_OK(cudaMemcpyAsync(kp, s->kpts, (size_t)cnt * 8, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(sc, s->scores, (size_t)cnt * 4, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  *n = cnt;
  e->enc_done = e->det_done = false;
  return DV_OK;
}

}  // namespace dv
