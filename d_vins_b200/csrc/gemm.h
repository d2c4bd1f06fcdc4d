// tcgen05 GEMM / implicit-GEMM 3x3 convolution: host-side plan + launch API.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dv {

// Fused epilogue, applied per accumulator row (= output pixel / token) in this order:
//   v = acc + bias[col]; v += res16 / res32 (optional); relu (optional); 2x2 max-pool (conv only, optional);
//   store fp16 (out16) and/or fp32 (out32).
struct EpiParams {
  __half* out16 = nullptr; int ld16 = 0;
  float* out32 = nullptr; int ld32 = 0;
  const float* bias = nullptr;
  const __half* res16 = nullptr; int ldr16 = 0;
  const float* res32 = nullptr; int ldr32 = 0;
  int relu = 0;
  int pool = 0;
  // rotary embedding fused into the store (LightGlue Wqkv; persistent kernel only): columns [0, rope_cols) are
  // (2j, 2j+1) pairs rotated by angle j = (col % 64) / 2 of the row: tables [rows, 32] fp32.
  const float* rope_cs = nullptr;
  const float* rope_sn = nullptr;
  int rope_cols = 0;
  // the same tables as one fp16 array [rows, 64] = (cos_j, sin_j) interleaved: 128 contiguous bytes per row instead of
  // two 128-byte fp32 rows (weights-resident kernel: the per-thread row fetch was its L1 bottleneck, profiles/README.md)
  const __half* rope16 = nullptr;
  // out16 in the channel-blocked layout [rows / blocked_hw][N/8][blocked_hw][8] (input format of conv_halo.cu);
  // persistent kernel only, plain (non-conv) mode.  0 = row-major.
  int blocked_hw = 0;
};
bool gemm_is_persistent();

struct GemmParams {
  int M, N, K;          // M rows actually computed (set at launch), N output columns, K reduction length
  int num_kb;           // K blocks of 64
  int n_tiles;
  int conv;             // 0: plain row-major A; 1: 3x3 pad-1 conv over NHWC A (H, W = OUTPUT size)
  int cstride = 1;      // conv: spatial stride 1 or 2 (stride 2: the TMA box samples every other input pixel)
  int H, W, cin, cin_blocks, tw_log2, tiles_w, tiles_h;
  EpiParams epi;
  // batched mode (persistent kernel only): tile -> (batch b, m_tile, n_tile); operands are row windows of the same two
  // tensors: A rows [batch[b].x, +batch[b].z), B rows [batch[b].y, +batch[b].w); output b at out32 + b * out_bstride.
  const int4* batch = nullptr;
  int batch_count = 0, batch_m_tiles = 0;
  long out_bstride = 0;
  long long* dbg = nullptr;   // DV_GEMM_DBG=1: epilogue cycle counters of CTA 0 (diagnostics only)
  int dbg_mode = 0;           // diagnostics: 1 skip global stores, 2 skip bias broadcast, 4 skip TMEM loads
};

struct GemmPlan {
  CUtensorMap tmA, tmB;
  GemmParams p;
  int bn = 0;           // 64 or 128
  int rows_cap = 0;     // plain: max M; conv: max images
  // staged epilogue (gemm_staged.cu): outputs / residuals move through shared memory with TMA.  The output maps are
  // encoded for the exact row count of a launch (TMA clips the last tile) and cached until it changes.
  int staged = 0;
  mutable CUtensorMap tmO16, tmO32, tmR16, tmR32;
  mutable int staged_rows = -1;
};

// A [M_cap, K] fp16 row-major (row pitch lda elements), B = weights [N, K] fp16 row-major (pitch ldb).
int plan_gemm(GemmPlan* pl, const __half* A, int lda, int M_cap, const __half* B, int ldb, int N, int K,
              const EpiParams& epi, int bn = 0);
// x NHWC [n_cap, H, W, cin] fp16 (cin % 64 == 0), w [cout, 9*cin] fp16 with k = (r*3+s)*cin + c.
// stride 2: H, W are the INPUT size; the output is (H/2) x (W/2) (pad 1, torchvision ResNet downsampling blocks)
int plan_conv3x3(GemmPlan* pl, const __half* x, int n_cap, int H, int W, int cin, const __half* w, int cout,
                 const EpiParams& epi, int stride = 1);
// rows = M (plain) or number of images (conv).
int launch_gemm(const GemmPlan& pl, int rows, cudaStream_t st);
// batched: `desc` device array of {a_row_off, b_row_off, m, n}; max_m / max_n bound the tile grid (fp32 output only)
int launch_gemm_batched(const GemmPlan& pl, const int4* desc, int count, int max_m, int max_n, long out_bstride,
                        cudaStream_t st);
int gemm_init();      // resolves cuTensorMapEncodeTiled, sets kernel attributes
int tmap_encode_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                    const uint32_t* box, bool swizzle128);
// 2-D row-major [rows, cols] map of 2- or 4-byte elements, SWIZZLE_128B boxes of box_cols x box_rows
int tmap_encode_rows(CUtensorMap* tm, const void* base, int elem_bytes, long cols, long rows, long pitch_bytes,
                     int box_cols, int box_rows);
bool gemm_staged_eligible(const GemmPlan& pl);
int gemm_staged_init();
int launch_gemm_staged(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st);
// weights-resident variant for K <= 512 (gemm_wres.cu): a CTA keeps a 128/256-column slab of W in shared memory
bool gemm_wres_eligible(const GemmPlan& pl, long m_tiles);
int gemm_wres_init();
int launch_gemm_wres(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st);
// the same on CTA pairs (gemm_pair.cu, cta_group::2): a pair shares a 256-column slab, M = 256 / N = 256 MMAs
bool gemm_pair_eligible(const GemmPlan& pl, long m_tiles);
int gemm_pair_init();
int launch_gemm_pair(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st);

// ---- LightGlue FFN first half in one kernel (lg_ffn0.cu): GELU(LayerNorm512(W0 . [x | ctx] + b0)), clusters of four CTAs
struct Ffn0Plan {
  CUtensorMap tmA, tmB;
  mutable CUtensorMap tmO16;
  mutable int out_rows = -1;
  const float *bias = nullptr, *gamma = nullptr, *beta = nullptr;
  __half* out = nullptr;
  int ldo = 0, rows_cap = 0;
};
int lg_ffn0_init();
int plan_lg_ffn0(Ffn0Plan* pl, const __half* A, int lda, int T_cap, const __half* W, const float* bias, const float* gamma,
                 const float* beta, __half* out, int ldo);
int launch_lg_ffn0(const Ffn0Plan& pl, int rows, cudaStream_t st);

// ---- weights-stationary halo-tile 3x3 convolution, 64 -> 64 channels (conv_halo.cu) -------------------------------
// Activations in the channel-blocked layout [N][C/8][H][W][8] ("NC/8HWC8"): one TMA box per 16x8-pixel output tile
// brings the 18x10 halo ONCE; the nine tap operands are views of it (no-swizzle UMMA descriptors with shifted start
// addresses).  All nine 64x64 filter taps stay resident in shared memory for the CTA's lifetime (persistent kernel).
struct HaloPlan {
  CUtensorMap tmX, tmW;
  int H = 0, W = 0, n_cap = 0, tiles_w = 0, tiles_h = 0;
  const float* bias = nullptr;
  __half* out = nullptr;
  int out_blocked = 1;     // 1: [N][8][Ho][Wo][8]; 0: NHWC [N][Ho][Wo][64]
  int relu = 1, pool = 0;
  // optional fusion of SuperPoint's conv1a (1->64) into the producer: the input is then the gray frame itself
  const float* gray = nullptr;   // [N,H,W] fp32
  const float* w1a = nullptr;    // [64,9]
  const float* b1a = nullptr;    // [64]
  // ... or, on the tensor cores, straight from the 1-channel u8 frame (takes precedence over `gray`)
  const uint8_t* img8 = nullptr; // [N,H,W] u8
};
int plan_conv3x3_halo64(HaloPlan* pl, const __half* x_blocked, int n_cap, int H, int W, const __half* w /*[64,576]*/,
                        const float* bias, __half* out, int out_blocked, int relu, int pool);
int launch_conv_halo64(const HaloPlan& pl, int n_img, cudaStream_t st);
int conv_halo_init();

// ---- halo-tile 3x3 convolution for 64/128 -> 128k channels (conv_halo128.cu): 32x8-pixel tiles (two M=128 row blocks),
// activations fetched once per tile, weights streamed through a ring of [128 x 64] blocks.  Input channel-blocked.
struct Halo128Plan {
  CUtensorMap tmX, tmW, tmWh;     // tmWh: [64 x 64] half-blocks for the CTA-pair kernel
  int H = 0, W = 0, n_cap = 0, tiles_w = 0, tiles_h = 0, cin = 0, cout = 0;
  const float* bias = nullptr;
  __half* out = nullptr;
  int out_blocked = 1;     // 1: [N][cout/8][Ho][Wo][8]; 0: NHWC [N][Ho][Wo][cout]
  int relu = 1, pool = 0;
};
int plan_conv3x3_halo128(Halo128Plan* pl, const __half* x_blocked, int n_cap, int H, int W, int cin,
                         const __half* w /*[cout, 9*cin]*/, int cout, const float* bias, __half* out, int out_blocked,
                         int relu, int pool);
int launch_conv_halo128(const Halo128Plan& pl, int n_img, cudaStream_t st);
int conv_halo128_init();

// ---- ResNet stem 7x7/2 3->64 as one implicit GEMM with in-kernel im2col (stem_conv.cu) ---------------------------------
struct StemPlan {
  CUtensorMap tmW;
  const __half* img = nullptr;   // [N,320,320,3] fp16
  const float* bias = nullptr;
  __half* out = nullptr;         // [N,160,160,64] fp16
  int n_cap = 0;
};
int plan_stem_conv(StemPlan* pl, const __half* img16, int n_cap, const __half* w /*[64,192], k=(r*7+s)*3+c*/, const float* bias,
                   __half* out);
int launch_stem_conv(const StemPlan& pl, int n_img, cudaStream_t st);
int stem_conv_init();

}  // namespace dv
