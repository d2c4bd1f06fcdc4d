// MixVPR on B200: reference pre-processing (preprocess_kernel.cu:348-435 semantics), torchvision ResNet-50 v1.5
// conv1..layer3 with folded BatchNorm on the tcgen05 GEMM / implicit-GEMM kernels (1x1 = plain GEMM over NHWC,
// 3x3 stride-1 = TMA tap-shifted implicit GEMM, the three stride-2 convs + the 7x7 stem = im2col + GEMM), and the
// feature-mixer aggregator (LayerNorm -> GEMM+ReLU -> GEMM+residual, channel_proj GEMM, row_proj + L2 kernel).
// Architecture restated from amaralibey/MixVPR + torchvision (un-vendored; see oracle/mixvpr.py).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>

#include "engine.h"

namespace dv {

#define DV_MIX_CHUNK_DEFAULT 0

struct MixNet {
  // parameters (device)
  std::vector<__half*> w16;
  std::vector<float*> b32;
  float d2i[6] = {0, 0, 0, 0, 0, 0};
  float *ln_g[4] = {}, *ln_b[4] = {};
  float *row_w = nullptr, *row_b = nullptr;
  float *chanT = nullptr, *chan_b = nullptr;   // reordered tail: channel_proj weight transposed [1024][256] fp32, bias [256]
  float* u32 = nullptr;                        // [B*1024, 2] row_proj applied to the mixer state
  float row_wsum[2] = {0.f, 0.f};              // sum_p row_proj.weight[r, p]
  bool reorder_tail = true;                    // DV_MIX_TAIL=0: transpose + channel_proj GEMM + row_proj kernel (A/B)
  // buffers
  __half* img16 = nullptr;      // [B,320,320,3]
  __half* col = nullptr;        // im2col scratch
  __half *xa = nullptr, *xb = nullptr, *t1 = nullptr, *t2 = nullptr, *ds = nullptr, *sub = nullptr;
  float* x32 = nullptr;         // [B*1024, 400] mixer state
  __half *ln16 = nullptr, *h16 = nullptr, *xT16 = nullptr;
  float* y32 = nullptr;         // [B*400, 256]
  float* gdesc = nullptr;       // [B, 512]
  std::vector<GemmPlan> plans;
  StemPlan stem;                       // 7x7/2 stem as one implicit GEMM with in-kernel im2col (stem_conv.cu)
  bool fused_stem = true;              // DV_MIX_STEM=0: im2col kernel + GEMM (A/B)
  bool strided_conv = true;            // DV_MIX_S2CONV=0: im2col kernel + GEMM for the two 3x3/2 convolutions (A/B)
  std::vector<HaloPlan> hplans;        // layer1's 64->64 3x3 convs run on the weights-stationary halo kernel
  std::vector<std::function<int(Engine*, int)>> ops;
  int n_launch = 0;
  int frame_off = 0;                   // first frame of the chunk being processed (mix_run chunks the batch)
  int chunk = 0;                       // DV_MIX_CHUNK: frames per pass over the op list (0 = the whole batch)
  GraphCache g_b1[2][2];               // the whole op list for one frame as a CUDA graph (per-keyframe latency path),
                                       // one per [frame buffer][1- / 3-channel]: the frame pointer is baked into the graph
};

// ------------------------------------------------------------------------------------------------ kernels
// u8 HxW(xch) -> fp16 NHWC [B,320,320,3]: non-centred inverse affine, bilinear with const 114 outside,
// floorf(v + .5f), BGR->RGB swap, (x/255 - mean)/std with the reference's BGR-ordered constants on RGB planes.
// Explicit _rn intrinsics keep the evaluation order of the restated oracle (no FMA contraction before floorf).
__global__ void k_mix_pre(const uint8_t* __restrict__ img, int H, int W, int ch, float m0, float m1, float m2,
                          float m3, float m4, float m5, __half* __restrict__ out, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / (320 * 320);
  const int p = i - b * 320 * 320;
  const int dy = p / 320, dx = p - dy * 320;
  const uint8_t* src = img + (int64_t)b * H * W * ch;
  const float sx = __fadd_rn(__fadd_rn(__fmul_rn(m0, (float)dx), __fmul_rn(m1, (float)dy)), m2);
  const float sy = __fadd_rn(__fadd_rn(__fmul_rn(m3, (float)dx), __fmul_rn(m4, (float)dy)), m5);
  float c[3];
  if (sx <= -1.f || sx >= (float)W || sy <= -1.f || sy >= (float)H) {
    c[0] = c[1] = c[2] = 114.f;
  } else {
    const int yl = (int)floorf(sy), xl = (int)floorf(sx);
    const int yh = yl + 1, xh = xl + 1;
    const float ly = __fsub_rn(sy, (float)yl), lx = __fsub_rn(sx, (float)xl);
    const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
    const bool ok1 = yl >= 0 && xl >= 0, ok2 = yl >= 0 && xh < W, ok3 = yh < H && xl >= 0, ok4 = yh < H && xh < W;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int kk = (ch == 3) ? k : 0;     // gray frames are replicated to BGR first (deep_net.cpp:1259-1262)
      const float v1 = ok1 ? (float)src[((int64_t)yl * W + xl) * ch + kk] : 114.f;
      const float v2 = ok2 ? (float)src[((int64_t)yl * W + xh) * ch + kk] : 114.f;
      const float v3 = ok3 ? (float)src[((int64_t)yh * W + xl) * ch + kk] : 114.f;
      const float v4 = ok4 ? (float)src[((int64_t)yh * W + xh) * ch + kk] : 114.f;
      const float s = __fadd_rn(
          __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)), __fmul_rn(w4, v4)),
          0.5f);
      c[k] = floorf(s);
    }
  }
  { const float t = c[2]; c[2] = c[0]; c[0] = t; }   // "Invert"
  const float a = 1.0f / 255.0f;
  const float mean[3] = {0.406f, 0.456f, 0.485f}, sd[3] = {0.225f, 0.224f, 0.229f};   // deep_net.cpp:1298-1300
#pragma unroll
  for (int k = 0; k < 3; ++k)
    out[(int64_t)i * 3 + k] = __float2half_rn(__fdiv_rn(__fsub_rn(__fmul_rn(c[k], a), mean[k]), sd[k]));
}

// 7x7 stride-2 pad-3 im2col of the 3-channel image: out [B*160*160, 192], k = (r*7+s)*3 + c, zero-padded to 192.
// One block per (frame, output row, 32 output columns): the 7 x 69 x 3 input window is staged once in shared memory
// (each input row is one contiguous 414-byte read); a pixel's 21 values of filter row r are contiguous there, so
// k -> smem[r*207 + ox*6 + k%21].  Stores are 16-byte, 384 contiguous bytes per output row.
#define STEM_OXB 32
#define STEM_ROWLEN ((2 * STEM_OXB + 5) * 3)   // 207
__global__ void __launch_bounds__(256) k_im2col_stem(const __half* __restrict__ img, __half* __restrict__ out) {
  __shared__ __half tile[7 * STEM_ROWLEN];
  const int b = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * STEM_OXB;
  const int ix0 = ox0 * 2 - 3;
  for (int i = threadIdx.x; i < 7 * STEM_ROWLEN; i += blockDim.x) {
    const int r = i / STEM_ROWLEN, q = i - r * STEM_ROWLEN;
    const int iy = oy * 2 + r - 3, ix = ix0 + q / 3;
    __half v = __float2half(0.f);
    if (iy >= 0 && iy < 320 && ix >= 0 && ix < 320) v = img[(((int64_t)b * 320 + iy) * 320 + ix0) * 3 + q];
    tile[i] = v;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < STEM_OXB * 24; q += blockDim.x) {
    const int px = q / 24, k0 = (q - px * 24) * 8;
    __align__(16) __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      const int r = k / 21;
      v[j] = (k < 147) ? tile[r * STEM_ROWLEN + px * 6 + (k - r * 21)] : __float2half(0.f);
    }
    const int64_t m = ((int64_t)b * 160 + oy) * 160 + ox0 + px;
    *reinterpret_cast<uint4*>(out + m * 192 + k0) = *reinterpret_cast<uint4*>(v);
  }
}

// 3x3 stride-2 pad-1 im2col over NHWC (C % 8 == 0): out [B*Ho*Wo, 9*C], k = (r*3+s)*C + c.
__global__ void k_im2col_s2(const __half* __restrict__ x, __half* __restrict__ out, int H, int W, int C, int Ho,
                            int Wo, int64_t total8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int cg = (int)(i % c8);
  int64_t t = i / c8;
  const int tap = (int)(t % 9);
  t /= 9;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int b = (int)(t / Ho);
  const int r = tap / 3, s = tap - r * 3;
  const int iy = oy * 2 + r - 1, ix = ox * 2 + s - 1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W)
    v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + iy) * W + ix) * C) + cg);
  reinterpret_cast<uint4*>(out)[i] = v;
}

// stride-2 spatial subsample (input of the 1x1/2 downsample convs)
__global__ void k_subsample2(const __half* __restrict__ x, __half* __restrict__ out, int H, int W, int C, int Ho,
                             int Wo, int64_t total8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int cg = (int)(i % c8);
  int64_t t = i / c8;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int b = (int)(t / Ho);
  reinterpret_cast<uint4*>(out)[i] =
      __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + oy * 2) * W + ox * 2) * C) + cg);
}

// 3x3 stride-2 pad-1 max-pool over NHWC fp16 (implicit -inf padding)
__global__ void k_maxpool3s2(const __half* __restrict__ x, __half* __restrict__ out, int H, int W, int C, int Ho,
                             int Wo, int64_t total8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int cg = (int)(i % c8);
  int64_t t = i / c8;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int b = (int)(t / Ho);
  __half2 m[4];
  const __half2 ninf = __float2half2_rn(-INFINITY);
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = ninf;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int iy = oy * 2 + r - 1, ix = ox * 2 + s - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + iy) * W + ix) * C) + cg);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], h[j]);
    }
  reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<uint4*>(m);
}

// [B, P, C] fp16 (NHWC feature map) -> [B, C, P] fp32 (mixer state), tiled through shared memory.
// Tile = 64 channels x 80 positions (P = 400 = 5 x 80): 16-byte loads along the channels (128 contiguous bytes per
// position), 16-byte stores along the positions (320 contiguous bytes per channel).  The 32 x 32 scalar version moved
// 157 MB per 64 frames at 2.5 TB/s.
#define TR_C 64
#define TR_P 80
__global__ void __launch_bounds__(256) k_transpose_h2f(const __half* __restrict__ in, float* __restrict__ out, int P, int C) {
  __shared__ float tile[TR_C][TR_P + 1];
  const int b = blockIdx.z, p0 = blockIdx.x * TR_P, c0 = blockIdx.y * TR_C;
  for (int i = threadIdx.x; i < TR_P * (TR_C / 8); i += 256) {
    const int pl = i >> 3, c8 = i & 7;
    const int p = p0 + pl;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (p < P) v = __ldg(reinterpret_cast<const uint4*>(in + ((int64_t)b * P + p) * C + c0 + c8 * 8));
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      tile[c8 * 8 + 2 * j][pl] = f.x;
      tile[c8 * 8 + 2 * j + 1][pl] = f.y;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TR_C * (TR_P / 4); i += 256) {
    const int cl = i / (TR_P / 4), p4 = i - cl * (TR_P / 4);
    const int p = p0 + p4 * 4;
    if (p + 3 < P) {
      *reinterpret_cast<float4*>(out + ((int64_t)b * C + c0 + cl) * P + p) =
          make_float4(tile[cl][p4 * 4], tile[cl][p4 * 4 + 1], tile[cl][p4 * 4 + 2], tile[cl][p4 * 4 + 3]);
    } else {
      for (int j = 0; j < 4 && p + j < P; ++j) out[((int64_t)b * C + c0 + cl) * P + p + j] = tile[cl][p4 * 4 + j];
    }
  }
}
// [B, C, P] fp32 -> [B, P, C] fp16: the mirror image (16-byte loads along the positions, 16-byte stores along the channels)
__global__ void __launch_bounds__(256) k_transpose_f2h(const float* __restrict__ in, __half* __restrict__ out, int C, int P) {
  __shared__ float tile[TR_C][TR_P + 1];
  const int b = blockIdx.z, c0 = blockIdx.x * TR_C, p0 = blockIdx.y * TR_P;
  for (int i = threadIdx.x; i < TR_C * (TR_P / 4); i += 256) {
    const int cl = i / (TR_P / 4), p4 = i - cl * (TR_P / 4);
    const int p = p0 + p4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p + 3 < P) {
      v = *reinterpret_cast<const float4*>(in + ((int64_t)b * C + c0 + cl) * P + p);
    } else {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < 4 && p + j < P; ++j) t[j] = in[((int64_t)b * C + c0 + cl) * P + p + j];
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
    tile[cl][p4 * 4] = v.x; tile[cl][p4 * 4 + 1] = v.y; tile[cl][p4 * 4 + 2] = v.z; tile[cl][p4 * 4 + 3] = v.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TR_P * (TR_C / 8); i += 256) {
    const int pl = i >> 3, c8 = i & 7;
    const int p = p0 + pl;
    if (p >= P) continue;
    __align__(16) __half2 hv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hv[j] = __floats2half2_rn(tile[c8 * 8 + 2 * j][pl], tile[c8 * 8 + 2 * j + 1][pl]);
    *reinterpret_cast<uint4*>(out + ((int64_t)b * P + p) * C + c0 + c8 * 8) = *reinterpret_cast<const uint4*>(hv);
  }
}

// LayerNorm over rows of length D <= 512, D % 4 == 0 (eps 1e-5, affine), fp32 in -> fp16 out.  One warp per row, the row
// held in registers (one pass over memory, 128-bit loads; the three-pass scalar version ran at 2.4 TB/s of L2 traffic).
__global__ void __launch_bounds__(256) k_layernorm_f2h(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ bta, __half* __restrict__ out,
                                                       int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const int nv = D >> 2;                                  // float4 per row
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (lane + 32 * k < nv) ? xr[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < nv) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q / (float)D + 1e-5f);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < nv) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * k);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bta) + lane + 32 * k);
        __align__(8) __half2 h[2] = {
            __floats2half2_rn((v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y),
            __floats2half2_rn((v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w)};
        *reinterpret_cast<uint2*>(out + row * D + (lane + 32 * k) * 4) = *reinterpret_cast<const uint2*>(h);
      }
  }
}

// row_proj (400 -> 2) on y [B,400,256], flatten index c*2 + r, L2-normalise -> [B,512].  One block per frame:
// 4 thread groups split the 400 positions (coalesced 1 KB rows), shared-memory combine, block-wide norm.
__global__ void __launch_bounds__(1024) k_rowproj_norm(const float* __restrict__ y, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out, int P) {
  __shared__ float part[4][256][2];
  __shared__ float red[8];
  const int b = blockIdx.x, o = threadIdx.x & 255, grp = threadIdx.x >> 8;
  float a0 = 0.f, a1 = 0.f;
  const int per = (P + 3) / 4;
  const int p0 = grp * per, p1 = min(P, p0 + per);
#pragma unroll 4
  for (int p = p0; p < p1; ++p) {
    const float v = y[((int64_t)b * P + p) * 256 + o];
    a0 = fmaf(v, __ldg(w + p), a0);
    a1 = fmaf(v, __ldg(w + P + p), a1);
  }
  part[grp][o][0] = a0;
  part[grp][o][1] = a1;
  __syncthreads();
  if (grp == 0) {
    a0 = ((part[0][o][0] + part[1][o][0]) + (part[2][o][0] + part[3][o][0])) + bias[0];
    a1 = ((part[0][o][1] + part[1][o][1]) + (part[2][o][1] + part[3][o][1])) + bias[1];
    float ss = a0 * a0 + a1 * a1;
#pragma unroll
    for (int off = 16; off; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((o & 31) == 0) red[o >> 5] = ss;
  }
  __syncthreads();
  if (grp == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
    out[(int64_t)b * 512 + o * 2 + 0] = a0 * inv;
    out[(int64_t)b * 512 + o * 2 + 1] = a1 * inv;
  }
}

// Reordered aggregator tail (r02).  channel_proj (1024 -> 256 over the channels) and row_proj (400 -> 2 over the
// positions) are both linear and act on different axes, so
//   z[c', r] = sum_c Wc[c', c] * (sum_p Wr[r, p] * x[c, p])  +  bc[c'] * sum_p Wr[r, p]  +  br[r]
// i.e. row_proj FIRST: one streaming pass over the fp32 mixer state (105 MB per 64 frames) leaves [1024, 2] per frame, and
// channel_proj shrinks to a 256 x 1024 x 2 product per frame.  The [C,P] fp32 -> [P,C] fp16 transpose (157 MB), the
// 400 x 256 x 1024 GEMM per frame and its fp32 output round trip disappear; the state is never rounded to fp16.
// k_rowproj_x: one warp per (frame, channel) row of 400 floats, the row in registers (128-bit loads).
__global__ void __launch_bounds__(256) k_rowproj_x(const float* __restrict__ x, const float* __restrict__ w,
                                                   float* __restrict__ u, int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const int nv = D >> 2;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4 w0[4], w1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool in = lane + 32 * k < nv;
    w0[k] = in ? __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    w1[k] = in ? __ldg(reinterpret_cast<const float4*>(w + D) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < nv) {
        const float4 v = xr[lane + 32 * k];
        a0 = fmaf(v.x, w0[k].x, fmaf(v.y, w0[k].y, fmaf(v.z, w0[k].z, fmaf(v.w, w0[k].w, a0))));
        a1 = fmaf(v.x, w1[k].x, fmaf(v.y, w1[k].y, fmaf(v.z, w1[k].z, fmaf(v.w, w1[k].w, a1))));
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == 0) *reinterpret_cast<float2*>(u + row * 2) = make_float2(a0, a1);
  }
}
// k_chanproj_norm: one block per frame; thread (o, grp) sums channels [256 grp, +256) of output o for both rows
// (coalesced rows of the transposed weight), shared-memory combine, bias terms, block-wide L2 norm, index o*2 + r.
__global__ void __launch_bounds__(1024) k_chanproj_norm(const float* __restrict__ u, const float* __restrict__ wT,
                                                        const float* __restrict__ bc, const float* __restrict__ br,
                                                        float ws0, float ws1, float* __restrict__ out) {
  __shared__ float2 us[1024];
  __shared__ float part[4][256][2];
  __shared__ float red[8];
  const int b = blockIdx.x, o = threadIdx.x & 255, grp = threadIdx.x >> 8;
  us[threadIdx.x] = *reinterpret_cast<const float2*>(u + ((int64_t)b * 1024 + threadIdx.x) * 2);
  __syncthreads();
  float a0 = 0.f, a1 = 0.f;
  const float* wp = wT + (int64_t)grp * 256 * 256 + o;
#pragma unroll 8
  for (int c = 0; c < 256; ++c) {
    const float wv = __ldg(wp + (int64_t)c * 256);
    const float2 uv = us[grp * 256 + c];
    a0 = fmaf(wv, uv.x, a0);
    a1 = fmaf(wv, uv.y, a1);
  }
  part[grp][o][0] = a0;
  part[grp][o][1] = a1;
  __syncthreads();
  if (grp == 0) {
    const float bo = bc[o];
    a0 = ((part[0][o][0] + part[1][o][0]) + (part[2][o][0] + part[3][o][0])) + fmaf(bo, ws0, br[0]);
    a1 = ((part[0][o][1] + part[1][o][1]) + (part[2][o][1] + part[3][o][1])) + fmaf(bo, ws1, br[1]);
    float ss = a0 * a0 + a1 * a1;
#pragma unroll
    for (int off = 16; off; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((o & 31) == 0) red[o >> 5] = ss;
  }
  __syncthreads();
  if (grp == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
    out[(int64_t)b * 512 + o * 2 + 0] = a0 * inv;
    out[(int64_t)b * 512 + o * 2 + 1] = a1 * inv;
  }
}

// ------------------------------------------------------------------------------------------------ host
namespace {

struct Folded { std::vector<float> w, b; };

// fold eval-mode BatchNorm (eps 1e-5) into the preceding bias-free conv
int fold_bn(Engine* e, const std::string& conv, const std::string& bn, int cout, int per_out, Folded* f) {
  const HostTensor* w = e->weight(conv + ".weight");
  const HostTensor* g = e->weight(bn + ".weight");
  const HostTensor* bt = e->weight(bn + ".bias");
  const HostTensor* mu = e->weight(bn + ".running_mean");
  const HostTensor* var = e->weight(bn + ".running_var");
  if (!w || !g || !bt || !mu || !var || w->numel() != (int64_t)cout * per_out || g->numel() != cout) {
    set_error("MixVPR weights: missing or mis-shaped " + conv + " / " + bn);
    return DV_ERR_WEIGHTS;
  }
  f->w.resize(w->data.size());
  f->b.resize(cout);
  for (int o = 0; o < cout; ++o) {
    const float sc = g->data[o] / sqrtf(var->data[o] + 1e-5f);
    for (int k = 0; k < per_out; ++k) f->w[(size_t)o * per_out + k] = w->data[(size_t)o * per_out + k] * sc;
    f->b[o] = bt->data[o] - mu->data[o] * sc;
  }
  return DV_OK;
}

// torch [cout,cin,kh,kw] -> [cout, (r*kw+s)*cin + c], optionally K-padded
std::vector<float> repack_khwc(const std::vector<float>& w, int cout, int cin, int kh, int kw, int kpad) {
  const int K = kh * kw * cin;
  std::vector<float> o((size_t)cout * kpad, 0.f);
  for (int oc = 0; oc < cout; ++oc)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < kh * kw; ++t) o[(size_t)oc * kpad + t * cin + c] = w[((size_t)oc * cin + c) * kh * kw + t];
  (void)K;
  return o;
}

}  // namespace

int mix_init(Engine* e) {
  MixNet* m = new MixNet();
  e->mix = m;
  const int B = e->B;
  // d2i exactly as cv::invertAffineTransform's CV_32F branch (double math, rounded to float)
  {
    const float sx = 320.f / (float)e->W, sy = 320.f / (float)e->H;
    double D = (double)sx * (double)sy;
    D = D != 0 ? 1.0 / D : 0.0;
    m->d2i[0] = (float)((double)sy * D); m->d2i[1] = -0.f * 0.f; m->d2i[2] = 0.f;
    m->d2i[3] = 0.f; m->d2i[4] = (float)((double)sx * D); m->d2i[5] = 0.f;
    m->d2i[1] = 0.f;
  }
  // ---- buffers
  const size_t P160 = (size_t)B * 160 * 160, P80 = (size_t)B * 80 * 80, P40 = (size_t)B * 40 * 40, P20 = (size_t)B * 400;
  DV_TRY(e->alloc(&m->img16, (size_t)B * 320 * 320 * 3 + 8));
  DV_TRY(e->alloc(&m->col, std::max({P160 * 192, P40 * 1152, P20 * 2304})));
  const size_t xmax = std::max({P160 * 64, P80 * 256, P40 * 512, P20 * 1024});
  DV_TRY(e->alloc(&m->xa, xmax));
  DV_TRY(e->alloc(&m->xb, xmax));
  DV_TRY(e->alloc(&m->t1, std::max({P80 * 128, P40 * 256, P80 * 64})));
  DV_TRY(e->alloc(&m->t2, std::max({P80 * 64, P40 * 128, P20 * 256})));
  DV_TRY(e->alloc(&m->ds, std::max({P80 * 256, P40 * 512, P20 * 1024})));
  DV_TRY(e->alloc(&m->sub, std::max({P40 * 256, P20 * 512})));
  DV_TRY(e->alloc(&m->x32, (size_t)B * 1024 * 400));
  DV_TRY(e->alloc(&m->ln16, (size_t)B * 1024 * 400));
  DV_TRY(e->alloc(&m->h16, (size_t)B * 1024 * 400));
  DV_TRY(e->alloc(&m->xT16, (size_t)B * 400 * 1024));
  DV_TRY(e->alloc(&m->y32, (size_t)B * 400 * 256));
  DV_TRY(e->alloc(&m->gdesc, (size_t)B * 512));
  DV_TRY(e->alloc(&m->u32, (size_t)B * 1024 * 2));
  m->plans.reserve(128);
  m->hplans.reserve(8);

  auto add_w = [&](const std::vector<float>& w, const std::vector<float>& b, __half** dw, float** db) -> int {
    DV_TRY(e->upload_f16(w, dw));
    DV_TRY(e->upload_f32(b, db));
    m->w16.push_back(*dw);
    m->b32.push_back(*db);
    return DV_OK;
  };
  auto add_gemm = [&](const __half* A, int lda, int Mcap, const __half* Wt, int ldb, int N, int K,
                      const EpiParams& ep, int rows_per_frame) -> int {
    m->plans.emplace_back();
    DV_TRY(plan_gemm(&m->plans.back(), A, lda, Mcap, Wt, ldb, N, K, ep));
    const int idx = (int)m->plans.size() - 1;
    m->ops.push_back([idx, rows_per_frame](Engine* en, int b) { return launch_gemm(en->mix->plans[idx], b * rows_per_frame, en->st); });
    m->n_launch++;
    return DV_OK;
  };
  auto add_conv = [&](const __half* x, int Hh, int Ww, int cin, const __half* Wt, int cout, const EpiParams& ep,
                      int stride = 1) -> int {
    m->plans.emplace_back();
    DV_TRY(plan_conv3x3(&m->plans.back(), x, B, Hh, Ww, cin, Wt, cout, ep, stride));
    const int idx = (int)m->plans.size() - 1;
    m->ops.push_back([idx](Engine* en, int b) { return launch_gemm(en->mix->plans[idx], b, en->st); });
    m->n_launch++;
    return DV_OK;
  };
  auto epi16 = [](__half* out, int ld, const float* bias, int relu, const __half* res = nullptr, int ldr = 0) {
    EpiParams ep; ep.out16 = out; ep.ld16 = ld; ep.bias = bias; ep.relu = relu; ep.res16 = res; ep.ldr16 = ldr;
    return ep;
  };

  { const char* env = getenv("DV_MIX_CHUNK"); m->chunk = env ? atoi(env) : DV_MIX_CHUNK_DEFAULT; }
  { const char* env = getenv("DV_MIX_S2CONV"); m->strided_conv = !(env && env[0] == '0'); }
  { const char* env = getenv("DV_MIX_TAIL"); m->reorder_tail = !(env && env[0] == '0'); }
  const std::string pre = "mix.backbone.model.";
  // ---- pre-processing + stem
  m->ops.push_back([](Engine* en, int b) {
    MixNet* mm = en->mix;
    const int total = b * 320 * 320;
    k_mix_pre<<<cdiv(total, 256), 256, 0, en->st>>>(en->d_img + (size_t)mm->frame_off * en->H * en->W * en->img_ch, en->H, en->W, en->img_ch, mm->d2i[0], mm->d2i[1],
                                                   mm->d2i[2], mm->d2i[3], mm->d2i[4], mm->d2i[5], mm->img16, total);
    if (mm->fused_stem) return launch_stem_conv(mm->stem, b, en->st);
    k_im2col_stem<<<dim3(160 / STEM_OXB, 160, b), 256, 0, en->st>>>(mm->img16, mm->col);
    return (int)DV_OK;
  });
  m->n_launch += 2;
  {
    const char* env = getenv("DV_MIX_STEM");
    m->fused_stem = !(env && env[0] == '0');
    Folded f;
    DV_TRY(fold_bn(e, pre + "conv1", pre + "bn1", 64, 147, &f));
    std::vector<float> wp = repack_khwc(f.w, 64, 3, 7, 7, 192);
    __half* dw; float* db;
    DV_TRY(add_w(wp, f.b, &dw, &db));
    if (m->fused_stem) {
      DV_TRY(plan_stem_conv(&m->stem, m->img16, B, dw, db, m->xa));
    } else {
      DV_TRY(add_gemm(m->col, 192, (int)P160, dw, 192, 64, 192, epi16(m->xa, 64, db, 1), 25600));
    }
  }
  m->ops.push_back([](Engine* en, int b) {
    MixNet* mm = en->mix;
    const int64_t t8 = (int64_t)b * 80 * 80 * 8;
    k_maxpool3s2<<<(unsigned)cdiv64(t8, 256), 256, 0, en->st>>>(mm->xa, mm->xb, 160, 160, 64, 80, 80, t8);
    return (int)DV_OK;
  });
  m->n_launch++;
  // ---- layer1..3
  __half* x = m->xb;       // current block input
  __half* xo = m->xa;      // current block output
  int inpl = 64, Hc = 80;
  const int layer_planes[3] = {64, 128, 256}, layer_blocks[3] = {3, 4, 6}, layer_stride[3] = {1, 2, 2};
  for (int li = 0; li < 3; ++li) {
    const int planes = layer_planes[li], outc = planes * 4;
    for (int bi = 0; bi < layer_blocks[li]; ++bi) {
      const std::string q = pre + "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
      const int stride = (bi == 0) ? layer_stride[li] : 1;
      const int Hin = Hc, Hout = Hc / stride;
      const int rows_in = Hin * Hin, rows_out = Hout * Hout;
      Folded f1, f2, f3;
      DV_TRY(fold_bn(e, q + "conv1", q + "bn1", planes, inpl, &f1));
      DV_TRY(fold_bn(e, q + "conv2", q + "bn2", planes, planes * 9, &f2));
      DV_TRY(fold_bn(e, q + "conv3", q + "bn3", outc, planes, &f3));
      __half *w1, *w2, *w3; float *b1, *b2, *b3;
      DV_TRY(add_w(f1.w, f1.b, &w1, &b1));
      DV_TRY(add_w(repack_khwc(f2.w, planes, planes, 3, 3, 9 * planes), f2.b, &w2, &b2));
      DV_TRY(add_w(f3.w, f3.b, &w3, &b3));
      // conv1 1x1 + ReLU
      const bool halo = gemm_is_persistent() && planes == 64 && stride == 1;
      {
        EpiParams e1 = epi16(m->t1, planes, b1, 1);
        if (halo) e1.blocked_hw = rows_in;            // channel-blocked output feeds the halo-tile 3x3
        DV_TRY(add_gemm(x, inpl, B * rows_in, w1, inpl, planes, inpl, e1, rows_in));
      }
      // conv2 3x3 (+stride) + ReLU
      if (halo) {
        m->hplans.emplace_back();
        DV_TRY(plan_conv3x3_halo64(&m->hplans.back(), m->t1, B, Hin, Hin, w2, b2, m->t2, /*out_blocked=*/0, 1, 0));
        const int hidx = (int)m->hplans.size() - 1;
        m->ops.push_back([hidx](Engine* en, int b) { return launch_conv_halo64(en->mix->hplans[hidx], b, en->st); });
        m->n_launch++;
      } else if (stride == 1) {
        DV_TRY(add_conv(m->t1, Hin, Hin, planes, w2, planes, epi16(m->t2, planes, b2, 1)));
      } else if (m->strided_conv) {
        // 3x3 / 2: implicit GEMM whose tap boxes sample every other input pixel (TMA traversal stride 2) - no im2col matrix
        DV_TRY(add_conv(m->t1, Hin, Hin, planes, w2, planes, epi16(m->t2, planes, b2, 1), 2));
      } else {
        const int C = planes, Hi = Hin, Ho = Hout;
        m->ops.push_back([C, Hi, Ho](Engine* en, int b) {
          MixNet* mm = en->mix;
          const int64_t t8 = (int64_t)b * Ho * Ho * 9 * (C / 8);
          k_im2col_s2<<<(unsigned)cdiv64(t8, 256), 256, 0, en->st>>>(mm->t1, mm->col, Hi, Hi, C, Ho, Ho, t8);
          return (int)DV_OK;
        });
        m->n_launch++;
        DV_TRY(add_gemm(m->col, 9 * planes, B * rows_out, w2, 9 * planes, planes, 9 * planes,
                        epi16(m->t2, planes, b2, 1), rows_out));
      }
      // identity / downsample
      const __half* res = x;
      if (bi == 0) {
        Folded fd;
        DV_TRY(fold_bn(e, q + "downsample.0", q + "downsample.1", outc, inpl, &fd));
        __half* wd; float* bd;
        DV_TRY(add_w(fd.w, fd.b, &wd, &bd));
        const __half* dsin = x;
        if (stride == 2) {
          const int C = inpl, Hi = Hin, Ho = Hout;
          const __half* xin = x;
          m->ops.push_back([C, Hi, Ho, xin](Engine* en, int b) {
            MixNet* mm = en->mix;
            const int64_t t8 = (int64_t)b * Ho * Ho * (C / 8);
            k_subsample2<<<(unsigned)cdiv64(t8, 256), 256, 0, en->st>>>(xin, mm->sub, Hi, Hi, C, Ho, Ho, t8);
            return (int)DV_OK;
          });
          m->n_launch++;
          dsin = m->sub;
        }
        DV_TRY(add_gemm(dsin, inpl, B * rows_out, wd, inpl, outc, inpl, epi16(m->ds, outc, bd, 0), rows_out));
        res = m->ds;
      }
      // conv3 1x1 + residual + ReLU
      DV_TRY(add_gemm(m->t2, planes, B * rows_out, w3, planes, outc, planes, epi16(xo, outc, b3, 1, res, outc), rows_out));
      std::swap(x, xo);
      inpl = outc;
      Hc = Hout;
    }
  }
  const __half* feat = x;   // [B,400,1024] NHWC
  // ---- aggregator
  m->ops.push_back([feat](Engine* en, int b) {
    k_transpose_h2f<<<dim3(cdiv(400, TR_P), 1024 / TR_C, b), 256, 0, en->st>>>(feat, en->mix->x32, 400, 1024);
    return (int)DV_OK;
  });
  m->n_launch++;
  for (int i = 0; i < 4; ++i) {
    const std::string p = "mix.aggregator.mix." + std::to_string(i) + ".mix.";
    const HostTensor *g = e->weight(p + "0.weight"), *bt = e->weight(p + "0.bias");
    const HostTensor *w1 = e->weight(p + "1.weight"), *b1 = e->weight(p + "1.bias");
    const HostTensor *w2 = e->weight(p + "3.weight"), *b2 = e->weight(p + "3.bias");
    if (!g || !bt || !w1 || !b1 || !w2 || !b2 || w1->numel() != 160000 || w2->numel() != 160000 || g->numel() != 400) {
      set_error("MixVPR weights: aggregator.mix." + std::to_string(i));
      return DV_ERR_WEIGHTS;
    }
    DV_TRY(e->upload_f32(g->data, &m->ln_g[i]));
    DV_TRY(e->upload_f32(bt->data, &m->ln_b[i]));
    __half *dw1, *dw2; float *db1, *db2;
    DV_TRY(add_w(w1->data, b1->data, &dw1, &db1));
    DV_TRY(add_w(w2->data, b2->data, &dw2, &db2));
    m->ops.push_back([i](Engine* en, int b) {
      MixNet* mm = en->mix;
      const int64_t rows = (int64_t)b * 1024;
      k_layernorm_f2h<<<(unsigned)std::min<int64_t>(cdiv64(rows, 8), 148 * 8), 256, 0, en->st>>>(mm->x32, mm->ln_g[i], mm->ln_b[i],
                                                                                            mm->ln16, rows, 400);
      return (int)DV_OK;
    });
    m->n_launch++;
    DV_TRY(add_gemm(m->ln16, 400, B * 1024, dw1, 400, 400, 400, epi16(m->h16, 400, db1, 1), 1024));
    EpiParams ep; ep.out32 = m->x32; ep.ld32 = 400; ep.res32 = m->x32; ep.ldr32 = 400; ep.bias = db2;
    DV_TRY(add_gemm(m->h16, 400, B * 1024, dw2, 400, 400, 400, ep, 1024));
  }
  {
    const HostTensor *wc = e->weight("mix.aggregator.channel_proj.weight"), *bc = e->weight("mix.aggregator.channel_proj.bias");
    const HostTensor *wr = e->weight("mix.aggregator.row_proj.weight"), *br = e->weight("mix.aggregator.row_proj.bias");
    if (!wc || !bc || !wr || !br || wc->numel() != 256 * 1024 || wr->numel() != 800 || bc->numel() != 256 || br->numel() != 2) {
      set_error("MixVPR weights: aggregator.channel_proj / row_proj");
      return DV_ERR_WEIGHTS;
    }
    DV_TRY(e->upload_f32(wr->data, &m->row_w));
    DV_TRY(e->upload_f32(br->data, &m->row_b));
    if (m->reorder_tail) {
      // row_proj first, channel_proj on [1024, 2] (see k_rowproj_x)
      std::vector<float> wT((size_t)1024 * 256);
      for (int o = 0; o < 256; ++o)
        for (int c = 0; c < 1024; ++c) wT[(size_t)c * 256 + o] = wc->data[(size_t)o * 1024 + c];
      DV_TRY(e->upload_f32(wT, &m->chanT));
      DV_TRY(e->upload_f32(bc->data, &m->chan_b));
      for (int r = 0; r < 2; ++r) {
        double acc = 0.0;
        for (int p = 0; p < 400; ++p) acc += (double)wr->data[(size_t)r * 400 + p];
        m->row_wsum[r] = (float)acc;
      }
      m->ops.push_back([](Engine* en, int b) {
        MixNet* mm = en->mix;
        const int64_t rows = (int64_t)b * 1024;
        k_rowproj_x<<<(unsigned)std::min<int64_t>(cdiv64(rows, 8), 148 * 8), 256, 0, en->st>>>(mm->x32, mm->row_w, mm->u32, rows, 400);
        k_chanproj_norm<<<b, 1024, 0, en->st>>>(mm->u32, mm->chanT, mm->chan_b, mm->row_b, mm->row_wsum[0], mm->row_wsum[1],
                                                mm->gdesc + (size_t)mm->frame_off * 512);
        return (int)DV_OK;
      });
      m->n_launch += 2;
    } else {
      m->ops.push_back([](Engine* en, int b) {
        k_transpose_f2h<<<dim3(1024 / TR_C, cdiv(400, TR_P), b), 256, 0, en->st>>>(en->mix->x32, en->mix->xT16, 1024, 400);
        return (int)DV_OK;
      });
      m->n_launch++;
      __half* dwc; float* dbc;
      DV_TRY(add_w(wc->data, bc->data, &dwc, &dbc));
      EpiParams ep; ep.out32 = m->y32; ep.ld32 = 256; ep.bias = dbc;
      DV_TRY(add_gemm(m->xT16, 1024, B * 400, dwc, 1024, 256, 1024, ep, 400));
      m->ops.push_back([](Engine* en, int b) {
        MixNet* mm = en->mix;
        k_rowproj_norm<<<b, 1024, 0, en->st>>>(mm->y32, mm->row_w, mm->row_b, mm->gdesc + (size_t)mm->frame_off * 512, 400);
        return (int)DV_OK;
      });
      m->n_launch++;
    }
  }
  e->dbg["mix_img"] = {m->img16, (int64_t)320 * 320 * 3, 1};
  e->dbg["mix_feat"] = {feat, (int64_t)400 * 1024, 1};
  e->dbg["mix_x32"] = {m->x32, (int64_t)1024 * 400, 0};
  e->dbg["mix_gdesc"] = {m->gdesc, 512, 0};
  return DV_OK;
}

void mix_free(Engine* e) {
  delete e->mix;
  e->mix = nullptr;
}

int mix_run(Engine* e, int b) {
  MixNet* m = e->mix;
  if (!m) { set_error("MixVPR not initialised (engine created without weights)"); return DV_ERR_INVALID; }
  StageScope sc(e, ST_MIX);
  e->image_acquire();
  // The op list runs over the batch in chunks of `chunk` frames: every ResNet activation of a chunk then fits the
  // 126 MB L2 between producer and consumer kernels (at 64 frames a layer1 tensor alone is 210 MB and every 1x1 / 3x3
  // kernel streams its operands from HBM).  Only the first op (frame pointer) and the last (descriptor row) see the offset.
  auto enqueue = [&]() -> int {
    const int step = (m->chunk > 0 && m->chunk < b) ? m->chunk : b;
    for (int f0 = 0; f0 < b; f0 += step) {
      m->frame_off = f0;
      const int nb = std::min(step, b - f0);
      for (auto& op : m->ops) DV_TRY(op(e, nb));
      DV_LAUNCHED(e, m->n_launch);
    }
    m->frame_off = 0;
    DV_CUDA_OK(cudaGetLastError());
    return DV_OK;
  };
  const int rc = b == 1 ? run_graphed(e, m->g_b1[e->img_idx][e->img_ch == 3 ? 1 : 0], enqueue) : enqueue();
  e->image_release();       // k_mix_pre (first op) was the last reader queued for this buffer
  return rc;
}

float* mix_gdesc(Engine* e) { return e->mix->gdesc; }

}  // namespace dv

using namespace dv;

extern "C" dv_status dv_mix_describe(dv_engine* h, float* des512) {
  if (!h) { dv::set_error("null engine"); return DV_ERR_INVALID; }
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!des512) { set_error("dv_mix_describe: null output"); return DV_ERR_INVALID; }
  e->adopt_upload();
  if (e->cur_b <= 0) { set_error("no frame uploaded"); return DV_ERR_INVALID; }
  if (!e->mix_done) { DV_TRY(mix_run(e, e->cur_b)); e->mix_done = true; }
  DV_CUDA_OK(cudaMemcpyAsync(des512, e->mix->gdesc, 512 * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  return DV_OK;
}
