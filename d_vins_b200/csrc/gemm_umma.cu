// Warp-specialised tcgen05 GEMM for sm_100a.
//   D[M,N] = A[M,K] * B[N,K]^T  (fp16 operands, fp32 accumulation in TMEM)
// One 128 x BN output tile per CTA; K streamed in 64-element (128-byte, SWIZZLE_128B) blocks through a
// TMA -> mbarrier -> tcgen05.mma ring.  Roles: warp 0 = TMA producer (1 lane), warp 1 = TMEM allocator + MMA
// issuer (1 lane), warps 2..5 = epilogue (tcgen05.ld -> registers -> fused bias/residual/ReLU/2x2-pool -> global).
// Two or more CTAs are resident per SM (<= 97 KB smem, <= 128 TMEM columns each), so one CTA's epilogue overlaps
// another's main loop.  The 3x3 convolution variant is an implicit GEMM: the A tile of K-block (tap, channel
// block) is a 4-D TMA box over the NHWC activation tensor shifted by the tap offset; out-of-bounds (including
// negative) coordinates are zero-filled by the TMA unit, which IS the conv zero padding.
#include <stdlib.h>

#include "gemm.h"
#include "umma.cuh"
#include "common.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 64) ? 4 : 3;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;   // + barriers + 1024-B alignment slack
};

template <int BN>
__global__ void __launch_bounds__(192) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int n0 = n_tile * BN;

  // conv tile coordinates
  int img = 0, h0 = 0, w0 = 0;
  if (p.conv) {
    const int tw_i = m_tile % p.tiles_w;
    const int t2 = m_tile / p.tiles_w;
    const int th_i = t2 % p.tiles_h;
    img = t2 / p.tiles_h;
    w0 = tw_i << p.tw_log2;
    h0 = th_i * (128 >> p.tw_log2);
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const int r = kb / C::STAGES;
        mbar_wait(&empty[s], (r & 1) ^ 1);
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
          tma_load_2d(sA, &tmA, &full[s], kb * 64, m_tile * 128);
          kB = kb * 64;
        }
        tma_load_2d(sB, &tmB, &full[s], kB, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const int r = kb / C::STAGES;
        mbar_wait(&full[s], r & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * C::STAGE_BYTES);
        const uint64_t da = make_desc_sw128(a_addr);
        const uint64_t db = make_desc_sw128(a_addr + C::A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x K=16 per 64-wide block: +32 B along K inside the swizzle atom
          tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
        tc_commit(&empty[s]);          // smem slot reusable once these MMAs retire
      }
      tc_commit(tmem_full);            // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // accumulator row == tile-local output row
    const EpiParams& ep = p.epi;
    long out_row = -1;                 // index of the output row (pixel/token); -1 = nothing to store
    bool writer = false;
    if (p.conv) {
      const int TW = 1 << p.tw_log2;
      const int hl = row >> p.tw_log2, wl = row & (TW - 1);
      const int h = h0 + hl, w = w0 + wl;
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
      writer = g < p.M;
      out_row = g;
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;          // warp-uniform
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      const int ncols = min(32, p.N - col0);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += (j < ncols) ? __ldg(ep.bias + col0 + j) : 0.f;
      }
      // residuals are only read for rows that will be stored (conv+pool never carries a residual)
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
      if (ep.out32 && writer && !ep.pool) {
        float4* op = reinterpret_cast<float4*>(ep.out32 + out_row * ep.ld32 + col0);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (g * 4 + 4 <= ncols) op[g] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      }
      if (ep.out16) {
        __align__(16) __half2 hv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hv[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        if (ep.pool) {
          // 2x2 max over accumulator rows {row, row^1, row^TW, row^TW^1}: all inside this warp (TW <= 16).
          // fp16 rounding is monotonic, so max-after-round == round-after-max.
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
          uint4* op = reinterpret_cast<uint4*>(ep.out16 + out_row * ep.ld16 + col0);
          const uint4* src = reinterpret_cast<const uint4*>(hv);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (g * 8 + 8 <= ncols) op[g] = src[g];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled g_encode = nullptr;
static bool g_use_persistent = true;
static bool g_use_staged = true;
bool gemm_is_persistent() { return g_use_persistent; }
int gemm_persistent_init();                                                                    // gemm_persistent.cu
int launch_gemm_persistent(const GemmPlan& pl, const GemmParams& p, long m_tiles, cudaStream_t st);

int gemm_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DV_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return DV_ERR_CUDA;
    }
    g_encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
  }
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<64>::SMEM_BYTES));
  DV_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<128>::SMEM_BYTES));
  {
    const char* env = getenv("DV_GEMM_PERSISTENT");      // debug toggle: 0 = one tile per CTA (original kernel)
    g_use_persistent = !(env && env[0] == '0');
    env = getenv("DV_GEMM_STAGED");                       // debug toggle: 0 = register epilogue for every GEMM
    g_use_staged = !(env && env[0] == '0');
  }
  int rc = gemm_persistent_init();
  if (rc) return rc;
  rc = gemm_staged_init();
  if (rc) return rc;
  rc = gemm_wres_init();
  if (rc) return rc;
  rc = gemm_pair_init();
  if (rc) return rc;
  rc = conv_halo128_init();
  if (rc) return rc;
  rc = stem_conv_init();
  if (rc) return rc;
  return conv_halo_init();
}

int conv_halo_init();   // conv_halo.cu

// fp16 tiled tensor map with zero OOB fill; swizzle128 selects SWIZZLE_128B (box inner = 128 B) or none.
int tmap_encode_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                    const uint32_t* box, bool swizzle128) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i];
  if (!g_encode || (reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tmap_encode_f16: driver entry point missing or base not 16-byte aligned");
    return DV_ERR_INVALID;
  }
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d) rank %d", (int)r, rank);
    set_error(b);
    return DV_ERR_CUDA;
  }
  return DV_OK;
}

int tmap_encode_rows(CUtensorMap* tm, const void* base, int elem_bytes, long cols, long rows, long pitch_bytes,
                     int box_cols, int box_rows) {
  if (!g_encode || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0 ||
      (box_cols * elem_bytes != 128 && box_cols * elem_bytes != 64 && box_cols * elem_bytes != 32)) {
    set_error("tmap_encode_rows: base / pitch must be 16-byte aligned and the box 128, 64 or 32 bytes wide");
    return DV_ERR_INVALID;
  }
  cuuint64_t d[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t st[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t bx[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_cols * elem_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : box_cols * elem_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled (rows map) failed (%d): cols %ld rows %ld pitch %ld", (int)r, cols,
             rows, pitch_bytes);
    set_error(b);
    return DV_ERR_CUDA;
  }
  return DV_OK;
}

static int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, const cuuint32_t* elem_strides = nullptr) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) es[i] = elem_strides[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tensor map base not 16-byte aligned");
    return DV_ERR_INVALID;
  }
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu box %u %u", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    set_error(b);
    return DV_ERR_CUDA;
  }
  return DV_OK;
}

static int pick_bn(int N, int bn) {
  if (bn == 64 || bn == 128) return bn;
  return (N <= 64) ? 64 : 128;
}

static int encode_weights(GemmPlan* pl, const __half* B, int ldb, int N, int K) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)pl->bn};
  return encode(&pl->tmB, B, 2, dims, strides, box);
}

int plan_gemm(GemmPlan* pl, const __half* A, int lda, int M_cap, const __half* B, int ldb, int N, int K,
              const EpiParams& epi, int bn) {
  if ((lda % 8) || (ldb % 8) || N <= 0 || K <= 0 || M_cap <= 0) {
    set_error("plan_gemm: lda/ldb must be multiples of 8 and extents positive");
    return DV_ERR_INVALID;
  }
  pl->bn = pick_bn(N, bn);
  pl->rows_cap = M_cap;
  GemmParams& p = pl->p;
  p = GemmParams{};
  p.M = M_cap; p.N = N; p.K = K;
  p.num_kb = cdiv(K, 64);
  p.n_tiles = cdiv(N, pl->bn);
  p.conv = 0;
  p.epi = epi;
  p.epi.pool = 0;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M_cap};
  cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
  cuuint32_t box[2] = {64, 128};
  int rc = encode(&pl->tmA, A, 2, dims, strides, box);
  if (rc) return rc;
  pl->staged = g_use_persistent && g_use_staged && gemm_staged_eligible(*pl);
  pl->staged_rows = -1;
  return encode_weights(pl, B, ldb, N, K);
}

int plan_conv3x3(GemmPlan* pl, const __half* x, int n_cap, int Hin, int Win, int cin, const __half* w, int cout,
                 const EpiParams& epi, int stride) {
  if (cin % 64) {
    set_error("plan_conv3x3: cin must be a multiple of 64");
    return DV_ERR_INVALID;
  }
  if (stride != 1 && (stride != 2 || (Hin & 1) || (Win & 1) || epi.pool)) {
    set_error("plan_conv3x3: stride must be 1, or 2 with even input size and no pooling");
    return DV_ERR_INVALID;
  }
  const int H = Hin / stride, W = Win / stride;      // output size: the tiles and the epilogue walk output pixels
  pl->bn = pick_bn(cout, 0);
  pl->rows_cap = n_cap;
  GemmParams& p = pl->p;
  p = GemmParams{};
  p.N = cout; p.K = 9 * cin;
  p.cin = cin; p.cin_blocks = cin / 64;
  p.num_kb = 9 * p.cin_blocks;
  p.n_tiles = cdiv(cout, pl->bn);
  p.conv = 1; p.H = H; p.W = W; p.cstride = stride;
  // tile = TH x TW pixels, TH*TW = 128.  Pooling needs TW <= 16 (2x2 partners inside one warp) and even TH, TW.
  int best = -1; long best_cost = 0;
  for (int l = (epi.pool ? 1 : 0); l <= (epi.pool ? 4 : 7); ++l) {
    const int TW = 1 << l, TH = 128 >> l;
    const long cost = (long)cdiv(W, TW) * cdiv(H, TH);
    if (best < 0 || cost < best_cost || (cost == best_cost && TW == 16)) { best = l; best_cost = cost; }
  }
  p.tw_log2 = best;
  p.tiles_w = cdiv(W, 1 << best);
  p.tiles_h = cdiv(H, 128 >> best);
  p.epi = epi;
  p.M = 0;
  // the tap operand of a TH x TW output tile is a box over the INPUT whose traversal stride is the convolution's stride:
  // boxDim counts un-strided positions (stride * T), the box delivers ceil(boxDim / stride) = T elements per dimension;
  // coordinates outside the image are zero-filled (= the convolution's padding)
  cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)n_cap};
  cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)Win * cin * 2, (cuuint64_t)Hin * Win * cin * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(stride << best), (cuuint32_t)(stride * (128 >> best)), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  int rc = encode(&pl->tmA, x, 4, dims, strides, box, es);
  if (rc) return rc;
  return encode_weights(pl, w, 9 * cin, cout, 9 * cin);
}

int launch_gemm(const GemmPlan& pl, int rows, cudaStream_t st) {
  if (rows <= 0) return DV_OK;
  if (rows > pl.rows_cap) {
    set_error("launch_gemm: rows exceed plan capacity");
    return DV_ERR_CAPACITY;
  }
  GemmParams p = pl.p;
  long m_tiles;
  if (p.conv) {
    m_tiles = (long)rows * p.tiles_w * p.tiles_h;
  } else {
    p.M = rows;
    m_tiles = cdiv(rows, 128);
  }
  static const bool want_time = getenv("DV_GEMM_TIME") != nullptr;   // diagnostics: per-launch event timing to stderr
  if (want_time && !p.conv) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const bool pair = pl.staged && gemm_pair_eligible(pl, m_tiles);
    const bool wres = pl.staged && gemm_wres_eligible(pl, m_tiles);
    cudaEventRecord(e0, st);
    int rc = pair ? launch_gemm_pair(pl, p, m_tiles, st)
             : wres ? launch_gemm_wres(pl, p, m_tiles, st)
                    : (pl.staged ? launch_gemm_staged(pl, p, m_tiles, st) : launch_gemm_persistent(pl, p, m_tiles, st));
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[gemm time] M %d N %d K %d %s: %.1f us = %.0f TFLOP/s\n", p.M, p.N, p.K,
            pair ? "pair" : wres ? "wres" : (pl.staged ? "staged" : "persist"), ms * 1e3, 2.0 * p.M * p.N * p.K / (ms * 1e-3) / 1e12);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
  }
  if (pl.staged) {
    if (gemm_pair_eligible(pl, m_tiles)) return launch_gemm_pair(pl, p, m_tiles, st);
    if (gemm_wres_eligible(pl, m_tiles)) return launch_gemm_wres(pl, p, m_tiles, st);
    return launch_gemm_staged(pl, p, m_tiles, st);
  }
  if (g_use_persistent) return launch_gemm_persistent(pl, p, m_tiles, st);
  const long grid = m_tiles * p.n_tiles;
  if (pl.bn == 64)
    umma_gemm_kernel<64><<<(unsigned)grid, 192, Cfg<64>::SMEM_BYTES, st>>>(pl.tmA, pl.tmB, p);
  else
    umma_gemm_kernel<128><<<(unsigned)grid, 192, Cfg<128>::SMEM_BYTES, st>>>(pl.tmA, pl.tmB, p);
  DV_CUDA_OK(cudaGetLastError());
  return DV_OK;
}

}  // namespace dv
