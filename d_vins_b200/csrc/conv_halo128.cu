// Halo-tile 3x3 convolution for the 128-channel SuperPoint layers (conv3a 64->128, conv3b / conv4a / conv4b 128->128,
// convPa ++ convDa 128->512) on tcgen05: 256-pixel output tiles, activations fetched once, weights streamed.
//
// Why: ncu / event timing of the tap-per-TMA implicit GEMM (gemm_persistent.cu, 128 x 128 tiles) puts these layers at
// 53-62 % of the measured bf16 peak: every tile pulls 9 x (16 KB of A + 16 KB of W) per 64 input channels through the
// SM's L2 port - 125 B per tensor-pipe clock against a measured ceiling of ~40 B/clk/SM (profiles/README.md).  The
// 64-channel layers solved that by keeping all nine filter taps resident (conv_halo.cu); 128 x 128 x 9 taps = 288 KB do
// not fit.  Here the CTA instead
//   * brings the (32+2) x (8+2) halo of a 32 x 8-pixel output tile ONCE with one 5-D TMA box over the channel-blocked
//     activation layout [N][C/8][H][W][8] (87 KB for 128 channels, two stages) and reuses it for all nine taps and for
//     every 128-column chunk of the output channels (4 chunks for the 512-channel head layer);
//   * streams the filter bank through a ring of [128 cout x 64 cin] fp16 blocks (16 KB, SWIZZLE_128B), each consumed by
//     EIGHT MMAs: 4 k-steps x 2 row blocks (the tile is two M = 128 accumulators, 2 x 128 TMEM columns per buffer,
//     two buffers = all 512 columns);
// = (87 + 288) KB per 2 x 72 MMAs (9216 tensor clocks) = 41 B/clk.  The A operand of tap (r, s) is the halo buffer seen
// through a no-swizzle K-major descriptor shifted by (r * 10 + s) * 16 bytes, exactly as in conv_halo.cu (SBO = one halo
// row = 160 B, LBO = one channel group = 34 * 10 * 16 B).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (bias, ReLU, optional 2x2
// max-pool, fp16 store in the channel-blocked or the NHWC layout); warp (q = warp % 4, h = (warp - 2) / 4) owns TMEM
// lanes [32q, 32q + 32) x columns [64h, 64h + 64) of both row blocks.
#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"
#include "../../include/dvins_perception.h"

namespace dv {

namespace {
constexpr int TH2 = 32, TW2 = 8;                     // output tile (pixels): two M = 128 row blocks of 16 x 8
constexpr int HH2 = TH2 + 2, HW2 = TW2 + 2;          // halo tile 34 x 10
constexpr int CGS = HH2 * HW2 * 16;                  // bytes of one 8-channel group of the halo: 5440
constexpr int WBLK = 128 * 128;                      // one weight block: 128 cout rows x 64 cin (128 B), SWIZZLE_128B
constexpr uint32_t SMEM_MAX = 232448;

struct H128Cfg {
  int halo_bytes, halo_stride, w_stages;
  uint32_t off_w, off_bar, smem_bytes;
};
__host__ __device__ inline H128Cfg h128_cfg(int cin) {
  H128Cfg c;
  c.halo_bytes = (cin / 8) * CGS;                                  // 87040 (cin 128) / 43520 (cin 64)
  c.halo_stride = (c.halo_bytes + 1023) & ~1023;
  c.off_w = 2u * (uint32_t)c.halo_stride;                          // two halo stages, then the weight ring
  const uint32_t fixed = 1024u /*alignment slack*/ + 256u /*barriers*/;
  int ws = (int)((SMEM_MAX - c.off_w - fixed) / (uint32_t)WBLK);
  c.w_stages = ws > 8 ? 8 : ws;                                    // 3 (cin 128) / 8 (cin 64)
  c.off_bar = c.off_w + (uint32_t)c.w_stages * (uint32_t)WBLK;
  c.smem_bytes = c.off_bar + 256u + 1024u;
  return c;
}
struct H128PairCfg {
  int halo_bytes, halo_stride, w_stages;
  uint32_t off_w, off_bar, smem_bytes;
};
__host__ __device__ inline H128PairCfg h128_pair_cfg(int cin) {
  H128PairCfg c;
  c.halo_bytes = (cin / 8) * CGS;
  c.halo_stride = (c.halo_bytes + 1023) & ~1023;
  c.off_w = 2u * (uint32_t)c.halo_stride;                          // two halo stages, then the ring of 8 KB half-blocks
  const uint32_t fixed = 1024u /*alignment slack*/ + 512u /*barriers*/;
  int ws = (int)((SMEM_MAX - c.off_w - fixed) / (uint32_t)(WBLK / 2));
  c.w_stages = ws > 16 ? 16 : ws;                                  // 6 (cin 128) / 16 (cin 64)
  c.off_bar = c.off_w + (uint32_t)c.w_stages * (uint32_t)(WBLK / 2);
  c.smem_bytes = c.off_bar + 512u + 1024u;
  return c;
}
}  // namespace

struct H128Params {
  int H, W, tiles_w, tiles_h, total_tiles, cin, cout, n_chunks;
  const float* bias;
  __half* out;
  int out_blocked, relu, pool;
};

// Epilogue shared by the single-CTA and the CTA-pair kernel: bias, ReLU, optional 2x2 max-pool, fp16 store (channel-
// blocked or NHWC).  Warp (q = warp % 4, h = (warp - 2) / 4) owns TMEM lanes [32q, 32q + 32) x columns [64h, 64h + 64) of
// both row blocks.  A CTA walks tiles first, first + stride, ...; in pair mode (`pair_rank_stride` = 2) CTA r of the
// pair owns tile 2 * i + r of pair-iteration i and releases the accumulator on the LEADER's barrier
// (`acc_empty_cluster` = its shared::cluster address, else 0).
__device__ __forceinline__ void h128_epilogue(const H128Params& p, uint32_t tmem_base, uint64_t* acc_full,
                                              uint64_t* acc_empty, uint32_t acc_empty_cluster, int first, int stride,
                                              int tile_mul, int tile_add, int warp, int lane) {
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int q = warp & 3;
  const int h = (warp - 2) >> 2;
  const int row = q * 32 + lane;                 // row inside a 128-row block: 16 x 8 pixels
  const int hl = row >> 3, wl = row & 7;
  const int Ho = p.pool ? (p.H >> 1) : p.H, Wo = p.pool ? (p.W >> 1) : p.W;
  int u = 0;
  for (int it_tile = first; it_tile * tile_mul < p.total_tiles; it_tile += stride) {
    const int tile = it_tile * tile_mul + tile_add;
    const bool tile_ok = tile < p.total_tiles;   // pair mode: the odd CTA of the last pair may have no tile
    const int img = tile / tiles_per_img;
    const int rem = tile - img * tiles_per_img;
    const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
    for (int ch = 0; ch < p.n_chunks; ++ch, ++u) {
      const int a = u & 1;
      mbar_wait(&acc_full[a], (u >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const int y = th_i * TH2 + mb * 16 + hl, x = tw_i * TW2 + wl;
        bool writer; int ho, wo;
        if (p.pool) {
          ho = y >> 1; wo = x >> 1;
          writer = !(hl & 1) && !(wl & 1) && ho < Ho && wo < Wo;
        } else {
          ho = y; wo = x;
          writer = y < p.H && x < p.W;
        }
        writer = writer && tile_ok;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c0 = ch * 128 + h * 64 + ci * 32;            // first output channel of these 32 columns
          uint32_t rr[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + mb * 128 + h * 64 + ci * 32), rr);
          tmem_ld_wait();
          if (mb == 1 && ci == 1) {                              // last TMEM read of this accumulator buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (acc_empty_cluster) mbar_arrive_cluster(acc_empty_cluster + (uint32_t)a * 8u);
              else mbar_arrive_cnt(&acc_empty[a]);
            }
          }
          __align__(16) __half2 hv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias + c0 + 2 * j));
            float v0 = __uint_as_float(rr[2 * j]) + b2.x;
            float v1 = __uint_as_float(rr[2 * j + 1]) + b2.y;
            if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            hv[j] = __floats2half2_rn(v0, v1);
          }
          if (p.pool) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint32_t uu = *reinterpret_cast<uint32_t*>(&hv[j]);
              uint32_t oo = __shfl_xor_sync(0xffffffffu, uu, 1);
              __half2 m = __hmax2(*reinterpret_cast<__half2*>(&uu), *reinterpret_cast<__half2*>(&oo));
              uu = *reinterpret_cast<uint32_t*>(&m);
              oo = __shfl_xor_sync(0xffffffffu, uu, 8);
              hv[j] = __hmax2(m, *reinterpret_cast<__half2*>(&oo));
            }
          }
          if (writer) {
            const uint4* src = reinterpret_cast<const uint4*>(hv);
            if (p.out_blocked) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int64_t off = ((((int64_t)img * (p.cout >> 3) + (c0 >> 3) + g) * Ho + ho) * Wo + wo) * 8;
                *reinterpret_cast<uint4*>(p.out + off) = src[g];
              }
            } else {
              uint4* dst = reinterpret_cast<uint4*>(p.out + (((int64_t)img * Ho + ho) * Wo + wo) * p.cout + c0);
#pragma unroll
              for (int g = 0; g < 4; ++g) dst[g] = src[g];
            }
          }
        }
      }
    }
  }
}


__global__ void __launch_bounds__(320, 1)
conv3x3_halo128_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                       const H128Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const H128Cfg c = h128_cfg(p.cin);
  uint64_t* hfull = reinterpret_cast<uint64_t*>(smem + c.off_bar);   // [2] halo landed / released
  uint64_t* hempty = hfull + 2;
  uint64_t* wfull = hempty + 2;                                      // [8] weight block landed / consumed
  uint64_t* wempty = wfull + 8;
  uint64_t* acc_full = wempty + 8;                                   // [2]
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int kh_n = p.cin >> 6;                                       // 64-channel halves per tap

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < 2; ++s) { mbar_init(&hfull[s], 1); mbar_init(&hempty[s], 1); }
    for (int s = 0; s < c.w_stages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      auto load_halo = [&](int tile, int it) {
        const int hs = it & 1;
        mbar_wait(&hempty[hs], ((it >> 1) & 1) ^ 1);
        const int img = tile / tiles_per_img;
        const int rem = tile - img * tiles_per_img;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        mbar_arrive_expect_tx(&hfull[hs], (uint32_t)c.halo_bytes);
        tma_load_5d(smem + hs * c.halo_stride, &tmX, &hfull[hs], 0, tw_i * TW2 - 1, th_i * TH2 - 1, 0, img);
      };
      int it = 0, wc = 0;
      if ((int)blockIdx.x < p.total_tiles) load_halo(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        // The next tile's halo is requested once this tile's MMAs have started (its first weight block was consumed, so
        // the previous tile - whose halo slot is being recycled - has retired): 87 KB land behind ~9000 MMA clocks
        // without ever stalling the weight stream.
        int j = 0;
        for (int ch = 0; ch < p.n_chunks; ++ch)
          for (int t = 0; t < 9; ++t)
            for (int kh = 0; kh < kh_n; ++kh, ++wc, ++j) {
              const int ws = wc % c.w_stages;
              mbar_wait(&wempty[ws], ((wc / c.w_stages) & 1) ^ 1);
              if (j == c.w_stages && tile + (int)gridDim.x < p.total_tiles) load_halo(tile + gridDim.x, it + 1);
              mbar_arrive_expect_tx(&wfull[ws], (uint32_t)WBLK);
              tma_load_2d(smem + c.off_w + (uint32_t)ws * WBLK, &tmW, &wfull[ws], t * p.cin + kh * 64, ch * 128);
            }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(128, 128);
      int it = 0, wc = 0, u = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int hs = it & 1;
        mbar_wait(&hfull[hs], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t da_base = make_desc_noswz(smem_u32(smem + hs * c.halo_stride), CGS, HW2 * 16);
        for (int ch = 0; ch < p.n_chunks; ++ch, ++u) {
          const int a = u & 1;
          mbar_wait(&acc_empty[a], ((u >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * 256);
          for (int t = 0; t < 9; ++t) {
            const int r = t / 3, sx = t - r * 3;
            for (int kh = 0; kh < kh_n; ++kh, ++wc) {
              const int ws = wc % c.w_stages;
              mbar_wait(&wfull[ws], (wc / c.w_stages) & 1);
              tc_fence_after();
              const uint64_t db = make_desc_sw128(smem_u32(smem + c.off_w + (uint32_t)ws * WBLK));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                // 16-byte units: channel group (kh * 8 + kk * 2) * 340, halo row (mb * 16 + r) * 10, column sx
                const uint64_t a_off = (uint64_t)((kh * 8 + kk * 2) * (HH2 * HW2) + r * HW2 + sx);
                const uint32_t acc = (uint32_t)((t | kh | kk) != 0);
                tc_mma_f16(d_tmem, da_base + a_off, db + (uint64_t)(kk * 2), idesc, acc);
                tc_mma_f16(d_tmem + 128u, da_base + a_off + (uint64_t)(16 * HW2), db + (uint64_t)(kk * 2), idesc, acc);
              }
              tc_commit(&wempty[ws]);
            }
          }
          tc_commit(&acc_full[a]);
        }
        tc_commit(&hempty[hs]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps
    h128_epilogue(p, tmem_base, acc_full, acc_empty, 0u, blockIdx.x, gridDim.x, 1, 0, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2, default): the two CTAs of a cluster work on two neighbouring 256-pixel tiles in
// lockstep and SHARE the weight stream.  Every MMA is M = 256 (128 pixels of each CTA's halo) x N = 128, the B operand
// split over the pair - each CTA loads and keeps only HALF of every weight block (64 cout rows x 64 cin, 8 KB):
//   * L2 -> SM weight traffic per SM halves (32 -> 16 B per tensor clock) and the ring holds twice as many blocks, i.e.
//     twice the prefetch distance in time (ncu r01/r02: tensor pipe 35-49 % with the 3-stage 16 KB ring at cin = 128);
//   * B-operand shared-memory reads per MMA halve (fact 2 of DESIGN.md: N <= 128 single-CTA MMAs are smem-read bound).
// Only the leader (cluster rank 0) issues MMAs; its commits are multicast to both CTAs' barriers; both CTAs' TMA bytes
// are credited to the leader's "full" barriers; every epilogue warp of both CTAs releases the accumulator on the leader.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv3x3_halo128_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWh,
                            const H128Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const H128PairCfg c = h128_pair_cfg(p.cin);
  uint64_t* hfull = reinterpret_cast<uint64_t*>(smem + c.off_bar);   // [2] halo landed / released
  uint64_t* hempty = hfull + 2;
  uint64_t* wfull = hempty + 2;                                      // [16] weight half-block landed / consumed
  uint64_t* wempty = wfull + 16;
  uint64_t* acc_full = wempty + 16;                                  // [2]
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int kh_n = p.cin >> 6;                                       // 64-channel halves per tap
  const int pair_tiles = (p.total_tiles + 1) >> 1;                   // pair-iteration i covers tiles 2i, 2i + 1

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmWh);
    for (int s = 0; s < 2; ++s) { mbar_init(&hfull[s], 1); mbar_init(&hempty[s], 1); }
    for (int s = 0; s < c.w_stages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 16); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      auto load_halo = [&](int pt, int it) {
        const int hs = it & 1;
        mbar_wait(&hempty[hs], ((it >> 1) & 1) ^ 1);
        // the odd CTA of the last pair may point one tile past the end: image index n_img -> the box lies outside the
        // tensor (n_cap may equal n_img) and is zero-filled, or reads a stale frame of the buffer; nothing is written
        const int tile = 2 * pt + (int)rank;
        const int img = tile / tiles_per_img;
        const int rem = tile - img * tiles_per_img;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        if (rank == 0) mbar_arrive_expect_tx(&hfull[hs], 2u * (uint32_t)c.halo_bytes);
        tma_load_5d_pair(smem + hs * c.halo_stride, &tmX, mapa_u32(smem_u32(&hfull[hs]), 0), 0, tw_i * TW2 - 1,
                         th_i * TH2 - 1, 0, img);
      };
      // The next tile's halo is requested once this tile's MMAs have started (a ring's worth of weight blocks was
      // consumed, so the previous tile - whose halo slot is recycled - has retired); tiles with fewer blocks than ring
      // stages request it at their last block (the producer may then wait for the previous tile, never deadlocks)
      const int blocks_per_tile = p.n_chunks * 9 * kh_n;
      const int halo_trigger = c.w_stages < blocks_per_tile - 1 ? c.w_stages : blocks_per_tile - 1;
      int it = 0, wc = 0;
      if (pair < pair_tiles) load_halo(pair, 0);
      for (int pt = pair; pt < pair_tiles; pt += n_pairs, ++it) {
        int j = 0;
        for (int ch = 0; ch < p.n_chunks; ++ch)
          for (int t = 0; t < 9; ++t)
            for (int kh = 0; kh < kh_n; ++kh, ++wc, ++j) {
              const int ws = wc % c.w_stages;
              mbar_wait(&wempty[ws], ((wc / c.w_stages) & 1) ^ 1);
              if (j == halo_trigger && pt + n_pairs < pair_tiles) load_halo(pt + n_pairs, it + 1);
              if (rank == 0) mbar_arrive_expect_tx(&wfull[ws], (uint32_t)WBLK);     // 2 x 8 KB
              tma_load_2d_pair(smem + c.off_w + (uint32_t)ws * (WBLK / 2), &tmWh, mapa_u32(smem_u32(&wfull[ws]), 0),
                               t * p.cin + kh * 64, ch * 128 + (int)rank * 64);
            }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16_f32(256, 128);
      int it = 0, wc = 0, u = 0;
      for (int pt = pair; pt < pair_tiles; pt += n_pairs, ++it) {
        const int hs = it & 1;
        mbar_wait(&hfull[hs], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t da_base = make_desc_noswz(smem_u32(smem + hs * c.halo_stride), CGS, HW2 * 16);
        for (int ch = 0; ch < p.n_chunks; ++ch, ++u) {
          const int a = u & 1;
          mbar_wait(&acc_empty[a], ((u >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * 256);
          for (int t = 0; t < 9; ++t) {
            const int r = t / 3, sx = t - r * 3;
            for (int kh = 0; kh < kh_n; ++kh, ++wc) {
              const int ws = wc % c.w_stages;
              mbar_wait(&wfull[ws], (wc / c.w_stages) & 1);
              tc_fence_after();
              const uint64_t db = make_desc_sw128(smem_u32(smem + c.off_w + (uint32_t)ws * (WBLK / 2)));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t a_off = (uint64_t)((kh * 8 + kk * 2) * (HH2 * HW2) + r * HW2 + sx);
                const uint32_t acc = (uint32_t)((t | kh | kk) != 0);
                tc_mma_f16_pair(d_tmem, da_base + a_off, db + (uint64_t)(kk * 2), idesc, acc);
                tc_mma_f16_pair(d_tmem + 128u, da_base + a_off + (uint64_t)(16 * HW2), db + (uint64_t)(kk * 2), idesc, acc);
              }
              tc_commit_pair(&wempty[ws], 3);
            }
          }
          tc_commit_pair(&acc_full[a], 3);
        }
        tc_commit_pair(&hempty[hs], 3);
      }
    }
  } else {
    h128_epilogue(p, tmem_base, acc_full, acc_empty, mapa_u32(smem_u32(&acc_empty[0]), 0), pair, n_pairs, 2, (int)rank,
                  warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int g_sms_h128 = 148;
static bool g_h128_pair = true;      // DV_SP_HALO128_PAIR=0: one CTA per tile stream (A/B)

int conv_halo128_init() {
  DV_CUDA_OK(cudaFuncSetAttribute(conv3x3_halo128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
  DV_CUDA_OK(cudaFuncSetAttribute(conv3x3_halo128_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
  { const char* e = getenv("DV_SP_HALO128_PAIR"); g_h128_pair = !(e && e[0] == '0'); }
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  DV_CUDA_OK(cudaDeviceGetAttribute(&g_sms_h128, cudaDevAttrMultiProcessorCount, dev));
  return DV_OK;
}

int plan_conv3x3_halo128(Halo128Plan* pl, const __half* x_blocked, int n_cap, int H, int W, int cin, const __half* w,
                         int cout, const float* bias, __half* out, int out_blocked, int relu, int pool) {
  if ((cin != 64 && cin != 128) || (cout & 127)) {
    set_error("plan_conv3x3_halo128: cin must be 64 or 128 and cout a multiple of 128");
    return DV_ERR_INVALID;
  }
  pl->H = H; pl->W = W; pl->n_cap = n_cap; pl->cin = cin; pl->cout = cout;
  pl->tiles_w = cdiv(W, TW2); pl->tiles_h = cdiv(H, TH2);
  pl->bias = bias; pl->out = out; pl->out_blocked = out_blocked; pl->relu = relu; pl->pool = pool;
  // activations [N][cin/8][H][W][8]
  const uint64_t dims[5] = {8, (uint64_t)W, (uint64_t)H, (uint64_t)(cin / 8), (uint64_t)n_cap};
  const uint64_t strides[4] = {16, (uint64_t)W * 16, (uint64_t)H * W * 16, (uint64_t)(cin / 8) * H * W * 16};
  const uint32_t box[5] = {8, HW2, HH2, (uint32_t)(cin / 8), 1};
  int rc = tmap_encode_f16(&pl->tmX, x_blocked, 5, dims, strides, box, /*swizzle128=*/false);
  if (rc) return rc;
  const uint64_t wd[2] = {(uint64_t)9 * cin, (uint64_t)cout};
  const uint64_t ws[1] = {(uint64_t)9 * cin * 2};
  const uint32_t wb[2] = {64, 128};
  rc = tmap_encode_f16(&pl->tmW, w, 2, wd, ws, wb, /*swizzle128=*/true);
  if (rc) return rc;
  const uint32_t wbh[2] = {64, 64};             // CTA-pair kernel: each CTA loads half of a weight block
  return tmap_encode_f16(&pl->tmWh, w, 2, wd, ws, wbh, /*swizzle128=*/true);
}

int launch_conv_halo128(const Halo128Plan& pl, int n_img, cudaStream_t st) {
  if (n_img <= 0) return DV_OK;
  if (n_img > pl.n_cap) { set_error("launch_conv_halo128: batch exceeds plan capacity"); return DV_ERR_CAPACITY; }
  H128Params p;
  p.H = pl.H; p.W = pl.W; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.total_tiles = n_img * pl.tiles_w * pl.tiles_h;
  p.cin = pl.cin; p.cout = pl.cout; p.n_chunks = pl.cout / 128;
  p.bias = pl.bias; p.out = pl.out; p.out_blocked = pl.out_blocked; p.relu = pl.relu; p.pool = pl.pool;
  if (g_h128_pair && p.total_tiles >= 4) {
    const int pair_tiles = (p.total_tiles + 1) / 2;
    const int pairs = pair_tiles < g_sms_h128 / 2 ? pair_tiles : g_sms_h128 / 2;
    const H128PairCfg cp = h128_pair_cfg(pl.cin);
    DV_CUDA_OK(launch_pdl(conv3x3_halo128_pair_kernel, dim3(2 * pairs), dim3(320), cp.smem_bytes, st, pl.tmX, pl.tmWh, p));
    DV_CUDA_OK(cudaGetLastError());
    return DV_OK;
  }
  const int grid = p.total_tiles < g_sms_h128 ? p.total_tiles : g_sms_h128;
  const H128Cfg c = h128_cfg(pl.cin);
  DV_CUDA_OK(launch_pdl(conv3x3_halo128_kernel, dim3(grid), dim3(320), c.smem_bytes, st, pl.tmX, pl.tmW, p));
  DV_CUDA_OK(cudaGetLastError());
  return DV_OK;
}

}  // namespace dv
