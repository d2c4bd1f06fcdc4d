// LightGlue (SuperPoint variant: 9 layers, d=256, 4 heads) on B200.
// Tokens of all images of all pairs are packed into one [T,*] buffer (each image segment padded to a multiple of 128
// rows), so every linear layer is ONE tcgen05 GEMM over all tokens (gemm_umma.cu) with fused bias / residual
// epilogues; attention is a flash-style kernel per (segment, head, 64-query tile) with online softmax;
// the assignment head (double log-softmax + matchability, mutual arg-max, ordered compaction) is a set of
// coalesced warp-shuffle kernels.  Architecture restated from cvg/LightGlue (un-vendored; see oracle/lightglue.py).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <utility>

#include "engine.h"
#include "lg.h"

namespace dv {

static constexpr int LG_D = 256, LG_HEADS = 4, LG_LAYERS = 9;
// Token segments (one per image) are packed back to back, each padded to LG_SEGPAD rows only: the row-wise kernels
// (GEMMs, LayerNorm) do not care about segment boundaries, the attention / assignment kernels bound every access by
// the segment's own (off, n).  (Padding to the 128-row GEMM tile, as before, cost 26 % extra rows at 150 + 662 tokens.)
static constexpr int LG_SEGPAD = 16;
__host__ __device__ static inline int lg_pad(int n) { return (n + LG_SEGPAD - 1) & ~(LG_SEGPAD - 1); }

struct AttnJob {
  const __half* q; const __half* k; const __half* v; __half* o;
  int nq, nk, ldq, ldk, ldv, ldo;
};

struct LgLayer {
  __half *wqkv, *wout, *wf0, *wf3, *cqkv, *cout, *cf0, *cf3;
  float *bqkv, *bout, *bf0, *bf3, *bcqkv, *bcout, *bcf0, *bcf3;
  float *ln_g, *ln_b, *cln_g, *cln_b;
  GemmPlan p_qkv, p_out, p_f0, p_f3, pc_qkv, pc_out, pc_f0, pc_f3;
  Ffn0Plan pf_f0, pfc_f0;              // fused ffn.0 + LayerNorm + GELU (lg_ffn0.cu)
};

struct LgNet {
  int P = 1, segcap = 1024, Tcap = 0;
  int ln_grid = 148 * 8;               // grid-stride LayerNorm+GELU kernel: 8 CTAs of 8 warps per SM
  // out_proj folded into the FFN's first linear (DV_LG_FOLD_OUT=0 keeps the separate GEMM, A/B):
  //   ffn.0([x | out_proj(ctx)]) = W0a x + (W0b Wout) ctx + (b0 + W0b bout)
  // two linear maps with nothing in between compose offline (like folding BatchNorm into a convolution), so the
  // attention writes its context straight into the second half of X2 and one GEMM launch per block disappears.
  bool fold_out = true;
  bool fuse_ffn0 = true;               // DV_LG_FUSE_FFN0=0: GEMM + k_lg_ln_gelu as two kernels (A/B); =2: fused at any size
  bool fuse_ffn0_always = false;
  float* Wr = nullptr;                 // [32,2]
  LgLayer L[LG_LAYERS];
  __half* wfinal = nullptr; float* bfinal = nullptr;   // pre-scaled by 256^-1/4
  float *wmatch = nullptr, *bmatch = nullptr;
  GemmPlan p_final, p_sim;
  // activations
  __half* X2 = nullptr;                // [T,512]: x (fp16 copy) | msg
  float* x32 = nullptr;                // [T,256] residual stream (fp32 master)
  __half* qkv = nullptr;               // [T,768]
  __half* ctx = nullptr;               // [T,256]
  __half* ffh = nullptr;               // [T,512] FFN pre-LayerNorm activations (fp16)
  __half* ffg = nullptr;               // [T,512]
  float *cs = nullptr, *sn = nullptr;  // [T,32] rotary cos / sin
  __half* rope16 = nullptr;            // [T,64] the same as fp16 (cos_j, sin_j) pairs (weights-resident GEMM epilogue)
  __half* md = nullptr;                // [T,256]
  float* z = nullptr;                  // [T] matchability logits
  float* sim = nullptr;                // [P, segcap, segcap]
  float* Lm = nullptr;                 // [P, segcap, segcap] log assignment
  float *rlse = nullptr, *clse = nullptr, *max0 = nullptr, *max1 = nullptr;   // [P, segcap]
  int *m0 = nullptr, *m1 = nullptr;    // [P, segcap]
  int* matches = nullptr; float* mscores = nullptr; float* mk0 = nullptr; float* mk1 = nullptr; int* kcount = nullptr;
  float* kpts = nullptr;               // [T,2] pixel keypoints (for mkpts)
  // per-call tables
  AttnJob *jobs_self = nullptr, *jobs_cross = nullptr, *h_jobs = nullptr;   // device x2, pinned host [4P]
  AttnJobU *ju_self = nullptr, *ju_cross = nullptr, *h_ju = nullptr;        // tcgen05 attention job tables
  // persistent attention: cost-sorted (job, head, query tile) lists, self then cross, 64 entries per pair slot each
  uint8_t *it_self = nullptr, *it_cross = nullptr, *h_items = nullptr;   // 32-byte descriptors
  bool attn_persist = true;            // DV_ATTN_PERSIST=0: one CTA per tile (A/B)
  int ipp = 64;                        // item-list entries per pair slot
  CUtensorMap tm_qkv;
  bool attn_umma = true;               // tcgen05 attention (lg_attn.cu: MN-major V operand, O accumulated in TMEM with lazy
                                       // rescale, two CTAs per SM); DV_LG_ATTN=mma: the mma.sync flash kernel of this file
  LgSeg *d_segs = nullptr, *h_segs = nullptr;                               // [2P]
  // staging for the host-vector API
  float *st_k = nullptr, *st_d = nullptr, *h_st_k = nullptr, *h_st_d = nullptr;   // [2*segcap,2], [2*segcap,256]
  int* h_out_i = nullptr; float* h_out_f = nullptr;
  cudaEvent_t ev_fetch = nullptr;      // results of the batched match in flight have landed in h_out_*
  std::map<std::pair<int, int>, GraphCache> graphs;     // P == 1 launch sequences by (m, n)
};

// ------------------------------------------------------------------------------------------------ kernels
// descriptors -> residual stream (fp32 master + fp16 copy in X2[:,0:256]); keypoints -> normalised -> rotary table.
// normalisation: deep_net.cpp:839-841,:874-880 ((kp - [W/2,H/2]) / max(W/2,H/2), integer halves).
__global__ void k_lg_load(const LgSeg* __restrict__ segs, const float* __restrict__ Wr, float* __restrict__ x32,
                          __half* __restrict__ X2, float* __restrict__ cs, float* __restrict__ sn,
                          __half* __restrict__ rope16, float* __restrict__ kpts_out) {
  const LgSeg sg = segs[blockIdx.y];
  const int rows = lg_pad(sg.n);
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
    const int64_t t = sg.off + r;
    float v[8];
    if (r < sg.n) {
      const float4* s = reinterpret_cast<const float4*>(sg.desc + (int64_t)r * 256 + lane * 8);
      const float4 a = __ldg(s), b = __ldg(s + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    float4* xo = reinterpret_cast<float4*>(x32 + t * 256 + lane * 8);
    xo[0] = make_float4(v[0], v[1], v[2], v[3]);
    xo[1] = make_float4(v[4], v[5], v[6], v[7]);
    __align__(16) __half2 hv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hv[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(X2 + t * 512 + lane * 8) = *reinterpret_cast<uint4*>(hv);
    float kx = 0.f, ky = 0.f;
    if (r < sg.n) { kx = sg.kpts[r * 2]; ky = sg.kpts[r * 2 + 1]; }
    if (lane == 0) { kpts_out[t * 2] = kx; kpts_out[t * 2 + 1] = ky; }
    const float sw = (float)(sg.w / 2), sh = (float)(sg.h / 2);
    const float sc = fmaxf(sw, sh);
    const float nx = __fdiv_rn(kx - sw, sc), ny = __fdiv_rn(ky - sh, sc);
    const float pr = nx * Wr[lane * 2] + ny * Wr[lane * 2 + 1];
    const float cv = cosf(pr), sv = sinf(pr);
    cs[t * 32 + lane] = cv;
    sn[t * 32 + lane] = sv;
    *reinterpret_cast<__half2*>(rope16 + t * 64 + lane * 2) = __floats2half2_rn(cv, sv);
  }
}

// rotary on q and k (columns [0,512) of qkv, 4 heads x 64 each): pairs (2j, 2j+1) rotated by angle j.
__global__ void k_lg_rope(__half* __restrict__ qkv, const float* __restrict__ cs, const float* __restrict__ sn,
                          int64_t T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (token, pair) over 256 pairs
  if (i >= T * 256) return;
  const int64_t t = i >> 8;
  const int pidx = (int)(i & 255);          // pair index within [q|k]: column = 2*pidx
  const int j = pidx & 31;                  // angle index within the head
  __half2* p = reinterpret_cast<__half2*>(qkv + t * 768) + pidx;
  const float2 x = __half22float2(*p);
  const float c = cs[t * 32 + j], s = sn[t * 32 + j];
  *p = __floats2half2_rn(x.x * c - x.y * s, x.y * c + x.x * s);
}

// ---- flash-style attention on mma.sync m16n8k16 (fp16 in, fp32 accumulate), head_dim 64, 64 queries per CTA.
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

#define ATT_LD 72   // padded smem row (halves): 144 B stride -> conflict-free ldmatrix
__device__ __forceinline__ float ex2_ftz(float x) {      // bare MUFU.EX2 (exp2f adds denormal range handling around it)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#define ATT_WARPS 4                 // warps per CTA: 16 query rows each (8 = 128-row tiles measured slower: coarser waves)
#define ATT_QT (ATT_WARPS * 16)
#define ATT_STAGES 3
#define ATT_SMEM (2 * ATT_STAGES * 64 * ATT_LD * 2)   // bytes
__global__ void __launch_bounds__(ATT_WARPS * 32) k_lg_attention(const AttnJob* __restrict__ jobs, float scale) {
  // three-stage K / V ring (cp.async, prefetch distance 2): ONE block barrier per 64-key chunk.  55 KB dynamic smem,
  // four CTAs per SM.  The Q tile is staged through K stage 2 (first written after the first barrier) and then lives
  // in registers, so its load overlaps the first two K / V prefetches.
  extern __shared__ __align__(16) __half att_smem[];
  __half (*sKb)[64 * ATT_LD] = reinterpret_cast<__half (*)[64 * ATT_LD]>(att_smem);
  __half (*sVb)[64 * ATT_LD] = reinterpret_cast<__half (*)[64 * ATT_LD]>(att_smem + ATT_STAGES * 64 * ATT_LD);
  __half* sQ = &sKb[ATT_STAGES - 1][0];
  pdl_trigger();
  pdl_wait();
  const AttnJob jb = jobs[blockIdx.z];
  const int q0 = blockIdx.x * ATT_QT;
  if (q0 >= jb.nq) return;
  const int head = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* Q = jb.q + head * 64;
  const __half* K = jb.k + head * 64;
  const __half* V = jb.v + head * 64;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float sl2 = scale * 1.4426950408889634f;   // softmax in base 2
  // stage loader: 16-byte cp.async per (row, 8-half column group); rows >= nk are zero-filled (src-size 0)
  auto load_kv = [&](int stage, int k0) {
    for (int i = tid; i < 64 * 8; i += ATT_WARPS * 32) {
      const int r = i >> 3, c = (i & 7) * 8;
      const bool ok = k0 + r < jb.nk;
      const int rr = ok ? k0 + r : 0;
      const uint32_t dk = (uint32_t)__cvta_generic_to_shared(&sKb[stage][r * ATT_LD + c]);
      const uint32_t dv = (uint32_t)__cvta_generic_to_shared(&sVb[stage][r * ATT_LD + c]);
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(K + (int64_t)rr * jb.ldk + c), "r"(sz) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dv), "l"(V + (int64_t)rr * jb.ldv + c), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int n_it = (jb.nk + 63) >> 6;
  load_kv(0, 0);
  if (n_it > 1) load_kv(1, 64);
  // Q tile -> smem (rows beyond nq read as zero)
  for (int i = tid; i < ATT_QT * 8; i += ATT_WARPS * 32) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < jb.nq) v = __ldg(reinterpret_cast<const uint4*>(Q + (int64_t)(q0 + r) * jb.ldq + c));
    *reinterpret_cast<uint4*>(&sQ[r * ATT_LD + c]) = v;
  }
  __syncthreads();
  uint32_t qa[4][4];   // A fragments of this warp's 16 query rows, 4 k-steps over d
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3],
            &sQ[(warp * 16 + (lane & 15)) * ATT_LD + ks * 16 + (lane >> 4) * 8]);
  for (int it = 0; it < n_it; ++it) {
    const int k0 = it * 64;
    if (it + 1 < n_it) asm volatile("cp.async.wait_group 1;" ::: "memory");   // chunk `it` landed (it+1 may be in flight)
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // chunk `it` visible to all warps; every warp is past chunk it-1 (and past its Q ldmatrix)
    if (it + 2 < n_it) load_kv((it + 2) % ATT_STAGES, k0 + 128);               // overwrites the stage of chunk it-1
    const __half* sK = sKb[it % ATT_STAGES];
    const __half* sV = sVb[it % ATT_STAGES];
    if (q0 + warp * 16 >= jb.nq) continue;                        // warp-uniform: this warp's 16 rows are all padding
    // S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {   // two 8-key tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, &sK[(np * 16 + (lane & 7) + (lane >> 4) * 8) * ATT_LD + ks * 16 + ((lane >> 3) & 1) * 8]);
        mma16816(s[np * 2], qa[ks], b0, b1);
        mma16816(s[np * 2 + 1], qa[ks], b2, b3);
      }
    }
    // mask keys >= nk (only the last chunk can be partial: block-uniform branch), online softmax (rows r0 = lane/4, r0+8)
    float mx[2] = {-INFINITY, -INFINITY};
    if (k0 + 64 > jb.nk) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = k0 + nt * 8 + (lane & 3) * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (key + (j & 1) >= jb.nk) s[nt][j] = -INFINITY;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) mx[j >> 1] = fmaxf(mx[j >> 1], s[nt][j]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], mnew[2], nms[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mnew[h] = fmaxf(mrow[h], mx[h]);            // finite: every chunk holds >= 1 valid key
      corr[h] = ex2_ftz((mrow[h] - mnew[h]) * sl2);
      nms[h] = -mnew[h] * sl2;                    // exp2(s * sl2 - m * sl2): one FFMA + MUFU per score
      mrow[h] = mnew[h];
      lrow[h] *= corr[h];
    }
    float ps[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = ex2_ftz(fmaf(s[nt][j], sl2, nms[j >> 1]));
        s[nt][j] = p;
        ps[j >> 1] += p;
      }
    lrow[0] += ps[0];
    lrow[1] += ps[1];
    // the running maxima settle after the first chunks: skip the 32 rescale multiplies when no row of the warp moved
    if (__any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f)) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        o[dt][0] *= corr[0]; o[dt][1] *= corr[0]; o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
      }
    }
    // O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {   // 16 keys per step
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_h2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_h2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_h2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {   // two 8-wide d tiles per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, &sV[(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LD + dp * 16 + (lane >> 4) * 8]);
        mma16816(o[dp * 2], pa, b0, b1);
        mma16816(o[dp * 2 + 1], pa, b2, b3);
      }
    }
  }
  // finalise: row sums across the quad, normalise, store fp16
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 1);
    lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 2);
  }
  const int r0 = q0 + warp * 16 + (lane >> 2);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + h * 8;
    if (r >= jb.nq) continue;
    const float inv = 1.f / lrow[h];
    __half* op = jb.o + (int64_t)r * jb.ldo + head * 64 + (lane & 3) * 2;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
      *reinterpret_cast<__half2*>(op + dt * 8) = __floats2half2_rn(o[dt][h * 2] * inv, o[dt][h * 2 + 1] * inv);
  }
}

// Exact (erf) GELU, 0.5 x (1 + erf(x / sqrt 2)), without libm's branchy erff: Abramowitz-Stegun 7.1.26,
// erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), |error| <= 1.5e-7 - three orders below the fp16
// rounding of the result.  ~16 instructions (2 MUFU) instead of ~30; the kernel is issue-bound (ncu: sm 62 %).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));        // bare MUFU (1 ulp): __frcp_rn /
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));     // exp2f expand to ~10 more each
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_z = poly * t * e;                                                   // 1 - erf(z), z >= 0
  const float half_x = 0.5f * x;
  // x >= 0: 0.5 x (2 - erfc);  x < 0: 0.5 x erfc
  return x >= 0.f ? fmaf(-half_x, erfc_z, x) : half_x * erfc_z;
}

// LayerNorm(512, eps 1e-5, affine) + exact (erf) GELU, fp32 in -> fp16 out.  One warp per token.
// Grid-stride over tokens, TWO tokens per warp per iteration with 128-bit loads (4 x 16 B in flight per lane: the
// one-token / 8-byte version sat at 2.5 TB/s of L2-resident traffic, latency-bound); gamma / beta of the lane's 16
// columns stay in registers.  Lane l owns columns [8l, 8l+8) and [256 + 8l, 256 + 8l + 8).
__global__ void __launch_bounds__(256) k_lg_ln_gelu(const __half* __restrict__ x, const float* __restrict__ g,
                                                    const float* __restrict__ b, __half* __restrict__ out, int64_t T) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  float gg[16], bb[16];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + h * 256 + lane * 8 + q * 4));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + h * 256 + lane * 8 + q * 4));
      gg[h * 8 + q * 4] = g4.x; gg[h * 8 + q * 4 + 1] = g4.y; gg[h * 8 + q * 4 + 2] = g4.z; gg[h * 8 + q * 4 + 3] = g4.w;
      bb[h * 8 + q * 4] = b4.x; bb[h * 8 + q * 4 + 1] = b4.y; bb[h * 8 + q * 4 + 2] = b4.z; bb[h * 8 + q * 4 + 3] = b4.w;
    }
  pdl_wait();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2; t0 < T; t0 += warps * 2) {
    uint4 raw[2][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const bool ok = t0 + k < T;
      const uint4* xr = reinterpret_cast<const uint4*>(x + (t0 + (ok ? k : 0)) * 512);
      raw[k][0] = xr[lane];
      raw[k][1] = xr[32 + lane];
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (t0 + k >= T) break;                          // warp-uniform
      float v[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[k][h]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __half22float2(h2[q]);
          v[h * 8 + 2 * q] = f.x; v[h * 8 + 2 * q + 1] = f.y;
        }
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.f / 512.f);
      float q2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q2 += d * d; }
#pragma unroll
      for (int o = 16; o; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
      const float rstd = rsqrtf(q2 * (1.f / 512.f) + 1e-5f);
      const float nmr = -mean * rstd;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        __align__(16) __half2 hv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = h * 8 + 2 * q;
          // ((v - mean) * rstd) * g + b as two FMAs: n = v * rstd - mean * rstd
          const float y0 = gelu_erf(fmaf(fmaf(v[i], rstd, nmr), gg[i], bb[i]));
          const float y1 = gelu_erf(fmaf(fmaf(v[i + 1], rstd, nmr), gg[i + 1], bb[i + 1]));
          hv[q] = __floats2half2_rn(y0, y1);
        }
        *reinterpret_cast<uint4*>(out + (t0 + k) * 512 + h * 256 + lane * 8) = *reinterpret_cast<const uint4*>(hv);
      }
    }
  }
}

// matchability logit z = w . x + b (fp32 residual stream).  One warp per token.
__global__ void k_lg_matchability(const float* __restrict__ x32, const float* __restrict__ w, const float* __restrict__ b,
                                  float* __restrict__ z, int64_t T) {
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += x32[t * 256 + c * 32 + lane] * __ldg(w + c * 32 + lane);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) z[t] = s + b[0];
}

__device__ __forceinline__ float logsigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

struct PairDesc { int off0, off1, m, n; };

// row log-sum-exp of sim [m,n] (ld = segcap).  One warp per row.
__global__ void k_lg_row_lse(const float* __restrict__ sim, int ld, int64_t pstride, const PairDesc* __restrict__ pd,
                             float* __restrict__ rlse, int segcap) {
  const PairDesc p = pd[blockIdx.y];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= p.m) return;
  const float* r = sim + blockIdx.y * pstride + (int64_t)i * ld;
  float mx = -INFINITY;
  for (int j = lane; j < p.n; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int j = lane; j < p.n; j += 32) s += expf(r[j] - mx);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) rlse[blockIdx.y * segcap + i] = mx + logf(s);
}

// column log-sum-exp: block = 32 columns x 8 row-lanes, coalesced row reads.
__global__ void __launch_bounds__(256) k_lg_col_lse(const float* __restrict__ sim, int ld, int64_t pstride,
                                                    const PairDesc* __restrict__ pd, float* __restrict__ clse,
                                                    int segcap) {
  __shared__ float sm[8][32], ss[8][32];
  const PairDesc p = pd[blockIdx.y];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  if (blockIdx.x * 32 >= p.n) return;
  const float* base = sim + blockIdx.y * pstride;
  float mx = -INFINITY, s = 0.f;
  if (j < p.n)
    for (int i = ty; i < p.m; i += 8) {
      const float v = base[(int64_t)i * ld + j];
      if (v > mx) { s = s * expf(mx - v) + 1.f; mx = v; } else { s += expf(v - mx); }
    }
  sm[ty][tx] = mx; ss[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < p.n) {
    float M = sm[0][tx];
    for (int k = 1; k < 8; ++k) M = fmaxf(M, sm[k][tx]);
    float S = 0.f;
    for (int k = 0; k < 8; ++k) if (ss[k][tx] > 0.f) S += ss[k][tx] * expf(sm[k][tx] - M);
    clse[blockIdx.y * segcap + j] = M + logf(S);
  }
}

// L = log_softmax_row + log_softmax_col + logsigmoid(z0_i) + logsigmoid(z1_j), and the row arg-max (lowest index on
// ties).  compute == 0: L is given (stage-isolated test) and only the arg-max runs.  One warp per row.
__global__ void k_lg_L_rowmax(const float* __restrict__ sim, float* __restrict__ Lm, int ld, int64_t pstride,
                              const PairDesc* __restrict__ pd, const float* __restrict__ rlse,
                              const float* __restrict__ clse, const float* __restrict__ z, int segcap, int compute,
                              int* __restrict__ m0, float* __restrict__ max0) {
  const PairDesc p = pd[blockIdx.y];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= p.m) return;
  const int64_t ro = blockIdx.y * pstride + (int64_t)i * ld;
  float best = -INFINITY; int bj = 0x7fffffff;
  float rl = 0.f, c0 = 0.f;
  if (compute) { rl = rlse[blockIdx.y * segcap + i]; c0 = logsigmoid(z[p.off0 + i]); }
  for (int j = lane; j < p.n; j += 32) {
    float v;
    if (compute) {
      const float s = sim[ro + j];
      v = ((s - rl) + (s - clse[blockIdx.y * segcap + j])) + (c0 + logsigmoid(z[p.off1 + j]));
      Lm[ro + j] = v;
    } else {
      v = Lm[ro + j];
    }
    if (v > best) { best = v; bj = j; }      // j ascending per lane: first max kept
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
  }
  if (lane == 0) { m0[blockIdx.y * segcap + i] = bj; max0[blockIdx.y * segcap + i] = best; }
}

__global__ void __launch_bounds__(256) k_lg_colmax(const float* __restrict__ Lm, int ld, int64_t pstride,
                                                   const PairDesc* __restrict__ pd, int segcap, int* __restrict__ m1) {
  __shared__ float sv[8][32];
  __shared__ int si[8][32];
  const PairDesc p = pd[blockIdx.y];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  if (blockIdx.x * 32 >= p.n) return;
  const float* base = Lm + blockIdx.y * pstride;
  float best = -INFINITY; int bi = 0x7fffffff;
  if (j < p.n)
    for (int i = ty; i < p.m; i += 8) {
      const float v = base[(int64_t)i * ld + j];
      if (v > best) { best = v; bi = i; }
    }
  sv[ty][tx] = best; si[ty][tx] = bi;
  __syncthreads();
  if (ty == 0 && j < p.n) {
    for (int k = 1; k < 8; ++k)
      if (sv[k][tx] > best || (sv[k][tx] == best && si[k][tx] < bi)) { best = sv[k][tx]; bi = si[k][tx]; }
    m1[blockIdx.y * segcap + j] = bi;
  }
}

// mutual check + threshold + ordered compaction (pairs ascending in i0: keyframe.cpp:623-654 relies on it) and the
// matched-keypoint gather (preprocess_kernel.cu:104-137: de-normalise (kn*scale + shift)).  One block per pair.
__global__ void __launch_bounds__(1024) k_lg_extract(const PairDesc* __restrict__ pd, const LgSeg* __restrict__ segs,
                                                     const int* __restrict__ m0, const int* __restrict__ m1,
                                                     const float* __restrict__ max0, int segcap, float thresh,
                                                     const float* __restrict__ kpts, int* __restrict__ matches,
                                                     float* __restrict__ mscores, float* __restrict__ mk0,
                                                     float* __restrict__ mk1, int* __restrict__ kcount, int out_base) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int pi = blockIdx.x;
  const PairDesc p = pd[pi];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < p.m; i0 += 1024) {
    const int i = i0 + tid;
    bool valid = false; int j = -1; float ms = 0.f;
    if (i < p.m) {
      j = m0[pi * segcap + i];
      ms = expf(max0[pi * segcap + i]);
      valid = (j >= 0 && j < p.n) && (m1[pi * segcap + j] == i) && (ms > thresh);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    const int wpre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const int base = base_s;
    if (valid) {
      const int k = base + woff + wpre;
      const int64_t o = (int64_t)(pi + out_base) * segcap + k;
      matches[o * 2] = i; matches[o * 2 + 1] = j;
      mscores[o] = ms;
      const LgSeg s0 = segs[pi * 2], s1 = segs[pi * 2 + 1];
      // normalise then de-normalise, as the reference's kernels do
      const float sw0 = (float)(s0.w / 2), sh0 = (float)(s0.h / 2), sc0 = fmaxf(sw0, sh0);
      const float sw1 = (float)(s1.w / 2), sh1 = (float)(s1.h / 2), sc1 = fmaxf(sw1, sh1);
      const float x0 = kpts[(int64_t)(p.off0 + i) * 2], y0 = kpts[(int64_t)(p.off0 + i) * 2 + 1];
      const float x1 = kpts[(int64_t)(p.off1 + j) * 2], y1 = kpts[(int64_t)(p.off1 + j) * 2 + 1];
      // the reference's three kernels in sequence, IEEE op by op (no FMA contraction): normalize_kpts (:52-65),
      // recover_normkpts (:104-137: kn * scale + shift), kpts_post_process (:68-101: (v + .5f) / s - .5f with
      // s = width_adj / width = 1 under the equal-size precondition).  Pinned by tests/test_refpre_gpu.py.
      auto rec = [](float v, float sh, float sc) {
        const float kn = __fdiv_rn(__fsub_rn(v, sh), sc);
        const float px = __fadd_rn(__fmul_rn(kn, sc), sh);
        return __fsub_rn(__fdiv_rn(__fadd_rn(px, 0.5f), 1.0f), 0.5f);
      };
      mk0[o * 2] = rec(x0, sw0, sc0); mk0[o * 2 + 1] = rec(y0, sh0, sc0);
      mk1[o * 2] = rec(x1, sw1, sc1); mk1[o * 2 + 1] = rec(y1, sh1, sc1);
    }
    __syncthreads();
    if (tid == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += wsum[w]; base_s = base + tot; }
    __syncthreads();
  }
  if (tid == 0) kcount[pi + out_base] = base_s;
}

// ------------------------------------------------------------------------------------------------ host
namespace {

// [W0a | W0b] (512 x 512), Wout (256 x 256)  ->  [W0a | W0b Wout], b0 + W0b bout.  fp64 accumulation so the fp16
// rounding of the folded weights is reproducible by the oracle (oracle/quant.py does the same product in float64).
void fold_out_proj(const std::vector<float>& w0, const std::vector<float>& b0, const std::vector<float>& wout,
                   const std::vector<float>& bout, std::vector<float>* wf, std::vector<float>* bf) {
  wf->assign(w0.begin(), w0.end());
  bf->assign(b0.begin(), b0.end());
  std::vector<double> acc(256);
  for (int o = 0; o < 512; ++o) {
    std::fill(acc.begin(), acc.end(), 0.0);
    double bacc = (double)b0[o];
    for (int m = 0; m < 256; ++m) {
      const double wv = (double)w0[(size_t)o * 512 + 256 + m];
      const float* wr = &wout[(size_t)m * 256];
      for (int c = 0; c < 256; ++c) acc[c] += wv * (double)wr[c];
      bacc += wv * (double)bout[m];
    }
    for (int c = 0; c < 256; ++c) (*wf)[(size_t)o * 512 + 256 + c] = (float)acc[c];
    (*bf)[o] = (float)bacc;
  }
}

int get_lin(Engine* e, const std::string& name, int cout, int cin, const HostTensor** w, const HostTensor** b) {
  *w = e->weight("lg." + name + ".weight");
  *b = e->weight("lg." + name + ".bias");
  if (!*w || !*b || (*w)->numel() != (int64_t)cout * cin || (*b)->numel() != cout) {
    set_error("LightGlue weights: missing or mis-shaped lg." + name);
    return DV_ERR_WEIGHTS;
  }
  return DV_OK;
}

}  // namespace

int lg_init(Engine* e) {
  LgNet* g = new LgNet();
  e->lg = g;
  g->P = e->B;
  g->segcap = (e->cfg.lg_max_kpts + 127) & ~127;
  g->Tcap = g->P * 2 * g->segcap;
  { const char* env = getenv("DV_LG_FOLD_OUT"); g->fold_out = !(env && env[0] == '0'); }
  { const char* env = getenv("DV_LG_FUSE_FFN0"); g->fuse_ffn0 = !(env && env[0] == '0'); g->fuse_ffn0_always = env && env[0] == '2'; }
  DV_TRY(lg_ffn0_init());
  DV_CUDA_OK(cudaFuncSetAttribute(k_lg_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  const int T = g->Tcap, P = g->P, SC = g->segcap;
  {
    const HostTensor* wr = e->weight("lg.posenc.Wr.weight");
    if (!wr || wr->numel() != 64) { set_error("LightGlue weights: lg.posenc.Wr.weight"); return DV_ERR_WEIGHTS; }
    DV_TRY(e->upload_f32(wr->data, &g->Wr));
  }
  DV_TRY(e->alloc(&g->X2, (size_t)T * 512));
  DV_TRY(e->alloc(&g->x32, (size_t)T * 256));
  DV_TRY(e->alloc(&g->qkv, (size_t)T * 768));
  DV_TRY(e->alloc(&g->ctx, (size_t)T * 256));
  DV_TRY(e->alloc(&g->ffh, (size_t)T * 512));
  DV_TRY(e->alloc(&g->ffg, (size_t)T * 512));
  DV_TRY(e->alloc(&g->cs, (size_t)T * 32));
  DV_TRY(e->alloc(&g->sn, (size_t)T * 32));
  DV_TRY(e->alloc(&g->rope16, (size_t)T * 64));
  DV_TRY(e->alloc(&g->md, (size_t)T * 256));
  DV_TRY(e->alloc(&g->z, (size_t)T));
  DV_TRY(e->alloc(&g->kpts, (size_t)T * 2));
  DV_TRY(e->alloc(&g->sim, (size_t)P * SC * SC));
  DV_TRY(e->alloc(&g->Lm, (size_t)P * SC * SC));
  DV_TRY(e->alloc(&g->rlse, (size_t)P * SC));
  DV_TRY(e->alloc(&g->clse, (size_t)P * SC));
  DV_TRY(e->alloc(&g->max0, (size_t)P * SC));
  DV_TRY(e->alloc(&g->max1, (size_t)P * SC));
  DV_TRY(e->alloc(&g->m0, (size_t)P * SC));
  DV_TRY(e->alloc(&g->m1, (size_t)P * SC));
  DV_TRY(e->alloc(&g->matches, (size_t)P * SC * 2));
  DV_TRY(e->alloc(&g->mscores, (size_t)P * SC));
  DV_TRY(e->alloc(&g->mk0, (size_t)P * SC * 2));
  DV_TRY(e->alloc(&g->mk1, (size_t)P * SC * 2));
  DV_TRY(e->alloc(&g->kcount, (size_t)P));
  DV_TRY(e->alloc(&g->jobs_self, (size_t)2 * P));
  DV_TRY(e->alloc(&g->jobs_cross, (size_t)2 * P));
  DV_TRY(e->alloc_pinned(&g->h_jobs, (size_t)4 * P));
  DV_TRY(e->alloc(&g->ju_self, (size_t)2 * P));
  DV_TRY(e->alloc(&g->ju_cross, (size_t)2 * P));
  DV_TRY(e->alloc_pinned(&g->h_ju, (size_t)4 * P));
  { const char* env = getenv("DV_ATTN_PERSIST"); g->attn_persist = !(env && env[0] == '0'); }
  g->ipp = 8 * (g->segcap / 128);      // 2 images x 4 heads x query tiles
  const size_t ib = (size_t)lg_attn_item_bytes();
  DV_TRY(e->alloc(&g->it_self, ib * g->ipp * P));
  DV_TRY(e->alloc(&g->it_cross, ib * g->ipp * P));
  DV_TRY(e->alloc_pinned(&g->h_items, ib * 2 * g->ipp * P));
  // default: the tcgen05 attention (lg_attn.cu, lazy rescale, two CTAs per SM); DV_LG_ATTN=mma selects the mma.sync kernel
  { const char* env = getenv("DV_LG_ATTN"); g->attn_umma = !(env && env[0] == 'm'); }
  DV_TRY(lg_attn_init());
  DV_TRY(plan_lg_attn(&g->tm_qkv, g->qkv, T));
  DV_TRY(e->alloc(&g->d_segs, (size_t)2 * P + (size_t)P));     // LgSeg[2P] followed by PairDesc[P] (same size class)
  DV_TRY(e->alloc_pinned(&g->h_segs, (size_t)2 * P + (size_t)P));
  DV_TRY(e->alloc(&g->st_k, (size_t)2 * SC * 2));
  DV_TRY(e->alloc(&g->st_d, (size_t)2 * SC * 256));
  DV_TRY(e->alloc_pinned(&g->h_st_k, (size_t)2 * SC * 2));
  DV_TRY(e->alloc_pinned(&g->h_st_d, (size_t)2 * SC * 256));
  DV_TRY(e->alloc_pinned(&g->h_out_i, (size_t)P * SC * 2 + P));
  DV_TRY(e->alloc_pinned(&g->h_out_f, (size_t)P * SC * 5));

  auto ep16 = [](__half* out, int ld, const float* bias) { EpiParams ep; ep.out16 = out; ep.ld16 = ld; ep.bias = bias; return ep; };
  for (int i = 0; i < LG_LAYERS; ++i) {
    LgLayer& L = g->L[i];
    const std::string ps = "transformers." + std::to_string(i) + ".self_attn.";
    const std::string pc = "transformers." + std::to_string(i) + ".cross_attn.";
    const HostTensor *w, *b;
    // Wqkv: de-interleave rows f = h*192 + d*3 + {q,k,v}  ->  f' = which*256 + h*64 + d
    DV_TRY(get_lin(e, ps + "Wqkv", 768, 256, &w, &b));
    {
      std::vector<float> wp(768 * 256), bp(768);
      for (int h = 0; h < 4; ++h)
        for (int d = 0; d < 64; ++d)
          for (int which = 0; which < 3; ++which) {
            const int src = h * 192 + d * 3 + which, dst = which * 256 + h * 64 + d;
            std::copy(w->data.begin() + (size_t)src * 256, w->data.begin() + (size_t)(src + 1) * 256, wp.begin() + (size_t)dst * 256);
            bp[dst] = b->data[src];
          }
      DV_TRY(e->upload_f16(wp, &L.wqkv)); DV_TRY(e->upload_f32(bp, &L.bqkv));
    }
    const HostTensor *wo, *bo;
    DV_TRY(get_lin(e, ps + "out_proj", 256, 256, &wo, &bo));
    DV_TRY(e->upload_f16(wo->data, &L.wout)); DV_TRY(e->upload_f32(bo->data, &L.bout));
    DV_TRY(get_lin(e, ps + "ffn.0", 512, 512, &w, &b));
    if (g->fold_out) {
      std::vector<float> wf, bf;
      fold_out_proj(w->data, b->data, wo->data, bo->data, &wf, &bf);
      DV_TRY(e->upload_f16(wf, &L.wf0)); DV_TRY(e->upload_f32(bf, &L.bf0));
    } else {
      DV_TRY(e->upload_f16(w->data, &L.wf0)); DV_TRY(e->upload_f32(b->data, &L.bf0));
    }
    DV_TRY(get_lin(e, ps + "ffn.3", 256, 512, &w, &b));
    DV_TRY(e->upload_f16(w->data, &L.wf3)); DV_TRY(e->upload_f32(b->data, &L.bf3));
    const HostTensor *lg_ = e->weight("lg." + ps + "ffn.1.weight"), *lb_ = e->weight("lg." + ps + "ffn.1.bias");
    const HostTensor *cg_ = e->weight("lg." + pc + "ffn.1.weight"), *cb_ = e->weight("lg." + pc + "ffn.1.bias");
    if (!lg_ || !lb_ || !cg_ || !cb_ || lg_->numel() != 512 || cg_->numel() != 512) { set_error("LightGlue weights: ffn.1 (LayerNorm)"); return DV_ERR_WEIGHTS; }
    DV_TRY(e->upload_f32(lg_->data, &L.ln_g)); DV_TRY(e->upload_f32(lb_->data, &L.ln_b));
    DV_TRY(e->upload_f32(cg_->data, &L.cln_g)); DV_TRY(e->upload_f32(cb_->data, &L.cln_b));
    {  // cross: [to_qk ; to_v] as one N=512 GEMM
      const HostTensor *wq, *bq, *wv, *bv;
      DV_TRY(get_lin(e, pc + "to_qk", 256, 256, &wq, &bq));
      DV_TRY(get_lin(e, pc + "to_v", 256, 256, &wv, &bv));
      std::vector<float> wp(wq->data), bp(bq->data);
      wp.insert(wp.end(), wv->data.begin(), wv->data.end());
      bp.insert(bp.end(), bv->data.begin(), bv->data.end());
      DV_TRY(e->upload_f16(wp, &L.cqkv)); DV_TRY(e->upload_f32(bp, &L.bcqkv));
    }
    DV_TRY(get_lin(e, pc + "to_out", 256, 256, &wo, &bo));
    DV_TRY(e->upload_f16(wo->data, &L.cout)); DV_TRY(e->upload_f32(bo->data, &L.bcout));
    DV_TRY(get_lin(e, pc + "ffn.0", 512, 512, &w, &b));
    if (g->fold_out) {
      std::vector<float> wf, bf;
      fold_out_proj(w->data, b->data, wo->data, bo->data, &wf, &bf);
      DV_TRY(e->upload_f16(wf, &L.cf0)); DV_TRY(e->upload_f32(bf, &L.bcf0));
    } else {
      DV_TRY(e->upload_f16(w->data, &L.cf0)); DV_TRY(e->upload_f32(b->data, &L.bcf0));
    }
    DV_TRY(get_lin(e, pc + "ffn.3", 256, 512, &w, &b));
    DV_TRY(e->upload_f16(w->data, &L.cf3)); DV_TRY(e->upload_f32(b->data, &L.bcf3));
    // plans (A operands are fixed buffers; rows are set at launch)
    { EpiParams ep = ep16(g->qkv, 768, L.bqkv);
      if (gemm_is_persistent()) { ep.rope_cs = g->cs; ep.rope_sn = g->sn; ep.rope16 = g->rope16; ep.rope_cols = 512; }   // rotary fused into the store
      DV_TRY(plan_gemm(&L.p_qkv, g->X2, 512, T, L.wqkv, 256, 768, 256, ep)); }
    DV_TRY(plan_gemm(&L.p_out, g->ctx, 256, T, L.wout, 256, 256, 256, ep16(g->X2 + 256, 512, L.bout)));
    DV_TRY(plan_gemm(&L.p_f0, g->X2, 512, T, L.wf0, 512, 512, 512, ep16(g->ffh, 512, L.bf0)));
    DV_TRY(plan_lg_ffn0(&L.pf_f0, g->X2, 512, T, L.wf0, L.bf0, L.ln_g, L.ln_b, g->ffg, 512));
    { EpiParams ep; ep.out32 = g->x32; ep.ld32 = 256; ep.res32 = g->x32; ep.ldr32 = 256; ep.out16 = g->X2; ep.ld16 = 512; ep.bias = L.bf3;
      DV_TRY(plan_gemm(&L.p_f3, g->ffg, 512, T, L.wf3, 512, 256, 512, ep)); }
    DV_TRY(plan_gemm(&L.pc_qkv, g->X2, 512, T, L.cqkv, 256, 512, 256, ep16(g->qkv, 768, L.bcqkv)));
    DV_TRY(plan_gemm(&L.pc_out, g->ctx, 256, T, L.cout, 256, 256, 256, ep16(g->X2 + 256, 512, L.bcout)));
    DV_TRY(plan_gemm(&L.pc_f0, g->X2, 512, T, L.cf0, 512, 512, 512, ep16(g->ffh, 512, L.bcf0)));
    DV_TRY(plan_lg_ffn0(&L.pfc_f0, g->X2, 512, T, L.cf0, L.bcf0, L.cln_g, L.cln_b, g->ffg, 512));
    { EpiParams ep; ep.out32 = g->x32; ep.ld32 = 256; ep.res32 = g->x32; ep.ldr32 = 256; ep.out16 = g->X2; ep.ld16 = 512; ep.bias = L.bcf3;
      DV_TRY(plan_gemm(&L.pc_f3, g->ffg, 512, T, L.cf3, 512, 256, 512, ep)); }
  }
  {
    const std::string pa = "log_assignment." + std::to_string(LG_LAYERS - 1) + ".";
    const HostTensor *w, *b;
    DV_TRY(get_lin(e, pa + "final_proj", 256, 256, &w, &b));
    std::vector<float> wp(w->data), bp(b->data);
    for (auto& v : wp) v *= 0.25f;     // / 256^(1/4): exact power-of-two scaling folded into the weights
    for (auto& v : bp) v *= 0.25f;
    DV_TRY(e->upload_f16(wp, &g->wfinal)); DV_TRY(e->upload_f32(bp, &g->bfinal));
    DV_TRY(plan_gemm(&g->p_final, g->X2, 512, T, g->wfinal, 256, 256, 256, ep16(g->md, 256, g->bfinal)));
    { EpiParams es; es.out32 = g->sim; es.ld32 = SC;
      // B = md too: its tensor map must span all T packed rows (the per-pair window is selected by row offsets)
      DV_TRY(plan_gemm(&g->p_sim, g->md, 256, T, g->md, 256, T, 256, es)); }
    const HostTensor *wm, *bm;
    DV_TRY(get_lin(e, pa + "matchability", 1, 256, &wm, &bm));
    DV_TRY(e->upload_f32(wm->data, &g->wmatch)); DV_TRY(e->upload_f32(bm->data, &g->bmatch));
  }
  e->dbg["lg_L"] = {g->Lm, (int64_t)SC * SC, 0};
  e->dbg["lg_sim"] = {g->sim, (int64_t)SC * SC, 0};
  e->dbg["lg_x32"] = {g->x32, (int64_t)2 * SC * 256, 0};
  return DV_OK;
}

void lg_free(Engine* e) {
  if (e->lg && e->lg->ev_fetch) cudaEventDestroy(e->lg->ev_fetch);
  delete e->lg;
  e->lg = nullptr;
}

static int lg_tail(Engine* e, int P, const PairDesc* d_pd, const LgSeg* d_segs, int max_m, int max_n, int compute, int out_base = 0) {
  LgNet* g = e->lg;
  const int SC = g->segcap;
  const int64_t ps = (int64_t)SC * SC;
  if (compute) {
    k_lg_row_lse<<<dim3(cdiv(max_m, 8), P), 256, 0, e->st>>>(g->sim, SC, ps, d_pd, g->rlse, SC);
    k_lg_col_lse<<<dim3(cdiv(max_n, 32), P), 256, 0, e->st>>>(g->sim, SC, ps, d_pd, g->clse, SC);
  }
  k_lg_L_rowmax<<<dim3(cdiv(max_m, 8), P), 256, 0, e->st>>>(g->sim, g->Lm, SC, ps, d_pd, g->rlse, g->clse, g->z, SC,
                                                          compute, g->m0, g->max0);
  k_lg_colmax<<<dim3(cdiv(max_n, 32), P), 256, 0, e->st>>>(g->Lm, SC, ps, d_pd, SC, g->m1);
  k_lg_extract<<<P, 1024, 0, e->st>>>(d_pd, d_segs, g->m0, g->m1, g->max0, SC, e->cfg.lg_filter_thresh, g->kpts,
                                      g->matches, g->mscores, g->mk0, g->mk1, g->kcount, out_base);
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, compute ? 5 : 3);
  return DV_OK;
}

// segs: host array [2P] with device pointers to keypoints / descriptors; results stay on the device.
int lg_run(Engine* e, int P, const LgSeg* segs_in, const std::function<int()>* after_load, int out_base) {
  LgNet* g = e->lg;
  if (!g) { set_error("LightGlue not initialised (engine created without weights)"); return DV_ERR_INVALID; }
  if (P < 1 || out_base < 0 || out_base + P > g->P) { set_error("lg_run: pair count exceeds max_batch"); return DV_ERR_CAPACITY; }
  StageScope sc(e, ST_LG);
  const int SC = g->segcap;
  // Per-call tables (pinned host + device): pair p of this call uses the entries of global pair slot out_base + p, so
  // the chunks of one batch never overwrite each other's tables while earlier chunks' copies are still queued.
  LgSeg* hs = g->h_segs + 2 * out_base;
  PairDesc* hp = reinterpret_cast<PairDesc*>(g->h_segs + 2 * g->P) + out_base;
  AttnJob* hj = g->h_jobs + 2 * out_base;                  // self jobs; cross jobs at + 2 * g->P
  AttnJobU* hu = g->h_ju + 2 * out_base;
  LgSeg* d_segs = g->d_segs + 2 * out_base;
  AttnJob *d_js = g->jobs_self + 2 * out_base, *d_jc = g->jobs_cross + 2 * out_base;
  AttnJobU *d_us = g->ju_self + 2 * out_base, *d_uc = g->ju_cross + 2 * out_base;
  const size_t ib = (size_t)lg_attn_item_bytes();
  uint8_t *hi_s = g->h_items + ib * g->ipp * out_base, *hi_c = g->h_items + ib * g->ipp * (g->P + out_base);
  uint8_t *d_is = g->it_self + ib * g->ipp * out_base, *d_ic = g->it_cross + ib * g->ipp * out_base;
  // attention output: the ctx buffer (separate out_proj GEMM), or - out_proj folded - the msg half of X2
  __half* const octx = g->fold_out ? g->X2 + 256 : g->ctx;
  const int oldo = g->fold_out ? 512 : 256;
  int off = 0, max_n_any = 0, max_m = 0, max_n = 0;
  for (int i = 0; i < 2 * P; ++i) {
    hs[i] = segs_in[i];
    if (hs[i].n < 1 || hs[i].n > e->cfg.lg_max_kpts) { set_error("lg_run: keypoint count out of range"); return DV_ERR_INVALID; }
    hs[i].off = off;
    off += lg_pad(hs[i].n);
    max_n_any = std::max(max_n_any, hs[i].n);
  }
  const int T = off;
  for (int p = 0; p < P; ++p) {
    const LgSeg &s0 = hs[2 * p], &s1 = hs[2 * p + 1];
    hp[p] = {s0.off, s1.off, s0.n, s1.n};
    max_m = std::max(max_m, s0.n);
    max_n = std::max(max_n, s1.n);
    for (int k = 0; k < 2; ++k) {
      const LgSeg& s = hs[2 * p + k];
      __half* base = g->qkv + (int64_t)s.off * 768;
      hj[2 * p + k] = {base, base + 256, base + 512, octx + (int64_t)s.off * oldo, s.n, s.n, 768, 768, 768, oldo};
      hu[2 * p + k] = {s.off, s.n, s.off, s.n, 0, 256, 512, 0};
    }
    __half* b0 = g->qkv + (int64_t)s0.off * 768;
    __half* b1 = g->qkv + (int64_t)s1.off * 768;
    hj[2 * g->P + 2 * p] = {b0, b1, b1 + 256, octx + (int64_t)s0.off * oldo, s0.n, s1.n, 768, 768, 768, oldo};
    hj[2 * g->P + 2 * p + 1] = {b1, b0, b0 + 256, octx + (int64_t)s1.off * oldo, s1.n, s0.n, 768, 768, 768, oldo};
    hu[2 * g->P + 2 * p] = {s0.off, s0.n, s1.off, s1.n, 0, 0, 256, 0};          // cross: qk of the other image, its v
    hu[2 * g->P + 2 * p + 1] = {s1.off, s1.n, s0.off, s0.n, 0, 0, 256, 0};
  }
  int n_is = 0, n_ic = 0;
  static const bool attn_lazy = [] { const char* en = getenv("DV_ATTN_LAZY"); return !(en && en[0] == '0'); }();
  if (g->attn_umma && g->attn_persist && attn_lazy) {
    n_is = lg_attn_items(hu, 2 * P, hi_s);
    n_ic = lg_attn_items(hu + 2 * g->P, 2 * P, hi_c);
  }
  // Everything below only queues work on e->st; the per-call tables above sit in pinned memory at fixed addresses.  A
  // single pair (the per-keyframe latency path: ~130 launches of a few microseconds each) is captured once per (m, n)
  // into a CUDA graph and replayed; batched calls amortise their launches over the batch and stay eager.
  auto enqueue = [&]() -> int {
    PairDesc* d_pd = reinterpret_cast<PairDesc*>(g->d_segs + 2 * g->P) + out_base;
    DV_CUDA_OK(cudaMemcpyAsync(d_segs, hs, sizeof(LgSeg) * 2 * P, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_pd, hp, sizeof(PairDesc) * P, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_js, hj, sizeof(AttnJob) * 2 * P, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_jc, hj + 2 * g->P, sizeof(AttnJob) * 2 * P, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_us, hu, sizeof(AttnJobU) * 2 * P, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(d_uc, hu + 2 * g->P, sizeof(AttnJobU) * 2 * P, cudaMemcpyHostToDevice, e->st));
    if (n_is) {
      DV_CUDA_OK(cudaMemcpyAsync(d_is, hi_s, ib * n_is, cudaMemcpyHostToDevice, e->st));
      DV_CUDA_OK(cudaMemcpyAsync(d_ic, hi_c, ib * n_ic, cudaMemcpyHostToDevice, e->st));
    }
    k_lg_load<<<dim3(8, 2 * P), 256, 0, e->st>>>(d_segs, g->Wr, g->x32, g->X2, g->cs, g->sn, g->rope16, g->kpts);
    DV_LAUNCHED(e, 1);
    if (after_load && *after_load) DV_TRY((*after_load)());
    const dim3 agrid(cdiv(max_n_any, ATT_QT), LG_HEADS, 2 * P);
    // the four-CTA-cluster kernel pays off once every cluster has a few row tiles; a single pair (the B = 1 latency path,
    // ~830 tokens) is 3 % faster with the two small kernels
    const bool fused0 = g->fuse_ffn0 && (T >= 2048 || g->fuse_ffn0_always);
    for (int i = 0; i < LG_LAYERS; ++i) {
      LgLayer& L = g->L[i];
      // self block
      DV_TRY(launch_gemm(L.p_qkv, T, e->st));
      if (!gemm_is_persistent())
        k_lg_rope<<<(unsigned)cdiv64((int64_t)T * 256, 256), 256, 0, e->st>>>(g->qkv, g->cs, g->sn, T);
      if (n_is) DV_TRY(launch_lg_attn_persist(g->tm_qkv, d_is, n_is, octx, oldo, 0.125f, e->st));
      else if (g->attn_umma) DV_TRY(launch_lg_attn(g->tm_qkv, d_us, 2 * P, max_n_any, octx, oldo, 0.125f, e->st));
      else DV_CUDA_OK(launch_pdl(k_lg_attention, agrid, dim3(ATT_WARPS * 32), ATT_SMEM, e->st, (const AttnJob*)d_js, 0.125f));
      if (!g->fold_out) DV_TRY(launch_gemm(L.p_out, T, e->st));
      if (fused0) {
        DV_TRY(launch_lg_ffn0(L.pf_f0, T, e->st));
      } else {
        DV_TRY(launch_gemm(L.p_f0, T, e->st));
        DV_CUDA_OK(launch_pdl(k_lg_ln_gelu, dim3(g->ln_grid), dim3(256), 0, e->st, (const __half*)g->ffh, (const float*)L.ln_g,
                              (const float*)L.ln_b, g->ffg, (int64_t)T));
      }
      DV_TRY(launch_gemm(L.p_f3, T, e->st));
      // cross block
      DV_TRY(launch_gemm(L.pc_qkv, T, e->st));
      if (n_ic) DV_TRY(launch_lg_attn_persist(g->tm_qkv, d_ic, n_ic, octx, oldo, 0.125f, e->st));
      else if (g->attn_umma) DV_TRY(launch_lg_attn(g->tm_qkv, d_uc, 2 * P, max_n_any, octx, oldo, 0.125f, e->st));
      else DV_CUDA_OK(launch_pdl(k_lg_attention, agrid, dim3(ATT_WARPS * 32), ATT_SMEM, e->st, (const AttnJob*)d_jc, 0.125f));
      if (!g->fold_out) DV_TRY(launch_gemm(L.pc_out, T, e->st));
      if (fused0) {
        DV_TRY(launch_lg_ffn0(L.pfc_f0, T, e->st));
      } else {
        DV_TRY(launch_gemm(L.pc_f0, T, e->st));
        DV_CUDA_OK(launch_pdl(k_lg_ln_gelu, dim3(g->ln_grid), dim3(256), 0, e->st, (const __half*)g->ffh, (const float*)L.cln_g,
                              (const float*)L.cln_b, g->ffg, (int64_t)T));
      }
      DV_TRY(launch_gemm(L.pc_f3, T, e->st));
      DV_LAUNCHED(e, (g->fold_out ? 10 : 12) - (fused0 ? 2 : 0));
    }
    DV_CUDA_OK(cudaGetLastError());
    DV_TRY(launch_gemm(g->p_final, T, e->st));
    k_lg_matchability<<<cdiv(T, 8), 256, 0, e->st>>>(g->x32, g->wmatch, g->bmatch, g->z, T);
    DV_LAUNCHED(e, 2);
    // sim_p = md0_p md1_p^T for all pairs in ONE batched launch: operands are row windows of the packed md buffer
    DV_TRY(launch_gemm_batched(g->p_sim, reinterpret_cast<const int4*>(d_pd), P, max_m, max_n, (long)SC * SC, e->st));
    DV_LAUNCHED(e, 1);
    return lg_tail(e, P, d_pd, d_segs, max_m, max_n, 1, out_base);
  };
  const bool has_hook = after_load && *after_load;
  if (P == 1 && !has_hook && out_base == 0) {
    if (g->graphs.size() > 32) g->graphs.clear();
    return run_graphed(e, g->graphs[std::make_pair(hs[0].n, hs[1].n)], enqueue);
  }
  return enqueue();
}

// copies match results of pair p to host buffers (synchronises)
int lg_fetch(Engine* e, int p, int cap, int32_t* matches, float* mscores, float* mk0, float* mk1, int32_t* k_out) {
  LgNet* g = e->lg;
  const int SC = g->segcap;
  int k = 0;
  DV_CUDA_OK(cudaMemcpyAsync(&k, g->kcount + p, sizeof(int), cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  if (k > cap) { set_error("lg_fetch: output capacity too small"); return DV_ERR_CAPACITY; }
  const int64_t o = (int64_t)p * SC;
  if (k > 0) {
    DV_CUDA_OK(cudaMemcpyAsync(matches, g->matches + o * 2, sizeof(int) * 2 * k, cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(mscores, g->mscores + o, sizeof(float) * k, cudaMemcpyDeviceToHost, e->st));
    if (mk0) DV_CUDA_OK(cudaMemcpyAsync(mk0, g->mk0 + o * 2, sizeof(float) * 2 * k, cudaMemcpyDeviceToHost, e->st));
    if (mk1) DV_CUDA_OK(cudaMemcpyAsync(mk1, g->mk1 + o * 2, sizeof(float) * 2 * k, cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaStreamSynchronize(e->st));
  }
  *k_out = k;
  return DV_OK;
}

// All pairs of a batched match in three strided D2H copies and ONE synchronisation (the per-pair version cost two
// stream synchronisations per pair: ~1 ms of idle GPU per 32-pair round).  slot[p] = caller-side index of pair p.
// Split in two so that the copies can be queued right behind the kernels (lg_fetch_batch_begin) and collected later
// (lg_fetch_batch_end: waits for the event, unpacks the pinned staging buffers) - dv_batch_match_begin / _end.
int lg_fetch_batch_begin(Engine* e, int P, int cap) {
  LgNet* g = e->lg;
  const int SC = g->segcap;
  if (cap > SC) cap = SC;
  int* h_m = g->h_out_i;
  int* h_k = g->h_out_i + (size_t)P * cap * 2;
  float* h_s = g->h_out_f;
  DV_CUDA_OK(cudaMemcpyAsync(h_k, g->kcount, sizeof(int) * P, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaMemcpy2DAsync(h_m, sizeof(int) * 2 * cap, g->matches, sizeof(int) * 2 * SC, sizeof(int) * 2 * cap, P,
                               cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaMemcpy2DAsync(h_s, sizeof(float) * cap, g->mscores, sizeof(float) * SC, sizeof(float) * cap, P,
                               cudaMemcpyDeviceToHost, e->st));
  if (!g->ev_fetch) DV_CUDA_OK(cudaEventCreateWithFlags(&g->ev_fetch, cudaEventDisableTiming));
  DV_CUDA_OK(cudaEventRecord(g->ev_fetch, e->st));
  return DV_OK;
}

int lg_fetch_batch_end(Engine* e, int P, int cap, const int* slot, int32_t* matches, float* mscores, int32_t* k_out) {
  LgNet* g = e->lg;
  const int SC = g->segcap;
  if (cap > SC) cap = SC;
  const int* h_m = g->h_out_i;
  const int* h_k = g->h_out_i + (size_t)P * cap * 2;
  const float* h_s = g->h_out_f;
  DV_CUDA_OK(cudaEventSynchronize(g->ev_fetch));
  for (int p = 0; p < P; ++p) {
    const int k = h_k[p], i = slot[p];
    if (k > cap) { set_error("lg_fetch_batch: output capacity too small"); return DV_ERR_CAPACITY; }
    memcpy(matches + (size_t)i * cap * 2, h_m + (size_t)p * cap * 2, sizeof(int) * 2 * k);
    memcpy(mscores + (size_t)i * cap, h_s + (size_t)p * cap, sizeof(float) * k);
    k_out[i] = k;
  }
  return DV_OK;
}

int lg_fetch_batch(Engine* e, int P, int cap, const int* slot, int32_t* matches, float* mscores, int32_t* k_out) {
  DV_TRY(lg_fetch_batch_begin(e, P, cap));
  return lg_fetch_batch_end(e, P, cap, slot, matches, mscores, k_out);
}

}  // namespace dv

using namespace dv;
#define DV_CHECK_ENGINE(e) do { if (!(e)) { dv::set_error("null engine"); return DV_ERR_INVALID; } } while (0)

extern "C" {

dv_status dv_lg_match(dv_engine* h, const float* kpts0, int32_t m, const float* kpts1, int32_t n, const float* desc0,
                      const float* desc1, int32_t h0, int32_t w0, int32_t h1, int32_t w1, int32_t* matches,
                      float* mscores, float* mkpts0, float* mkpts1, int32_t* k_out) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  LgNet* g = e->lg;
  if (!g) { set_error("dv_lg_match: engine created without weights"); return DV_ERR_INVALID; }
  if (e->match_pending) { set_error("dv_lg_match: a batched match is in flight (collect it with dv_batch_match_end first)"); return DV_ERR_INVALID; }
  if (!kpts0 || !kpts1 || !desc0 || !desc1 || !matches || !mscores || !k_out) { set_error("dv_lg_match: null argument"); return DV_ERR_INVALID; }
  // engine profile of the reference: 10 <= kpts <= 1024 (README.md:180)
  if (m < 10 || n < 10 || m > e->cfg.lg_max_kpts || n > e->cfg.lg_max_kpts) { set_error("dv_lg_match: need 10 <= m,n <= lg_max_kpts"); return DV_ERR_INVALID; }
  const int SC = g->segcap;
  {
    StageScope sc(e, ST_COPY);
    memcpy(g->h_st_k, kpts0, sizeof(float) * 2 * m);
    memcpy(g->h_st_k + 2 * SC, kpts1, sizeof(float) * 2 * n);
    memcpy(g->h_st_d, desc0, sizeof(float) * 256 * m);
    memcpy(g->h_st_d + (size_t)256 * SC, desc1, sizeof(float) * 256 * n);
    DV_CUDA_OK(cudaMemcpyAsync(g->st_k, g->h_st_k, sizeof(float) * 2 * m, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(g->st_k + 2 * SC, g->h_st_k + 2 * SC, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(g->st_d, g->h_st_d, sizeof(float) * 256 * m, cudaMemcpyHostToDevice, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(g->st_d + (size_t)256 * SC, g->h_st_d + (size_t)256 * SC, sizeof(float) * 256 * n, cudaMemcpyHostToDevice, e->st));
  }
  LgSeg segs[2] = {{g->st_k, g->st_d, m, w0, h0, 0}, {g->st_k + 2 * SC, g->st_d + (size_t)256 * SC, n, w1, h1, 0}};
  DV_TRY(lg_run(e, 1, segs));
  return (dv_status)lg_fetch(e, 0, std::min(m, n), matches, mscores, mkpts0, mkpts1, k_out);
}

dv_status dv_dbg_match_extract(dv_engine* h, const float* L, int32_t m, int32_t n, int32_t* matches, float* mscores,
                               int32_t* k_out) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  LgNet* g = e->lg;
  if (!g) { set_error("dv_dbg_match_extract: engine created without weights"); return DV_ERR_INVALID; }
  if (e->match_pending) { set_error("dv_dbg_match_extract: a batched match is in flight (dv_batch_match_end first)"); return DV_ERR_INVALID; }
  const int SC = g->segcap;
  if (!L || m < 1 || n < 1 || m > SC || n > SC) { set_error("dv_dbg_match_extract: bad shape"); return DV_ERR_INVALID; }
  DV_CUDA_OK(cudaMemcpy2DAsync(g->Lm, sizeof(float) * SC, L, sizeof(float) * n, sizeof(float) * n, m, cudaMemcpyHostToDevice, e->st));
  LgSeg* hs = g->h_segs;
  hs[0] = {g->st_k, g->st_d, m, 2, 2, 0};
  hs[1] = {g->st_k, g->st_d, n, 2, 2, SC};
  PairDesc* hp = reinterpret_cast<PairDesc*>(hs + 2 * g->P);
  hp[0] = {0, SC, m, n};
  DV_CUDA_OK(cudaMemcpyAsync(g->d_segs, hs, sizeof(LgSeg) * 3 * g->P, cudaMemcpyHostToDevice, e->st));
  DV_TRY(lg_tail(e, 1, reinterpret_cast<PairDesc*>(g->d_segs + 2 * g->P), g->d_segs, m, n, 0));
  return (dv_status)lg_fetch(e, 0, std::min(m, n), matches, mscores, nullptr, nullptr, k_out);
}

}  // extern "C"
