/*
 * dvins_perception.h - C ABI of the B200-native loop-closure perception engine.
 *
 * Drop-in boundary for D_VINS' loop_fusion deep-perception facade (upstream paths relative to
 * kajo-kurisu/D_VINS):
 *   Estimator_net::Estimator  loop_fusion/src/deep_net/deep_net.h:141-171   (sp_extractor x2, lg_matcher)
 *   MixVPR_net::MixVPR        loop_fusion/src/deep_net/deep_net.h:117-135   (mix_extractor)
 *   KeyFrame::sort_vec_faiss  loop_fusion/src/keyframe.cpp:262-346          (faiss IndexFlatIP kNN)
 * Plain C: opaque handle, POD arguments, caller-allocated outputs, no torch / OpenCV / STL types.
 * Every entry point returns dv_status (0 = OK); the failing thread's message is in dv_last_error().
 * Calls on ONE engine are serialised by contract (the reference calls everything from its single
 * `process` thread, pose_graph_node.cpp:264-397); independent engines are fully concurrent.
 * There is NO CPU fallback: dv_create fails with DV_ERR_NOGPU when no sm_100 device is visible.
 *
 * Precondition carried over from every shipped D_VINS config: network input size == image size
 * (width_adj == image_width, height_adj == image_height); other sizes -> DV_ERR_UNSUPPORTED.
 */
#ifndef DVINS_PERCEPTION_H_
#define DVINS_PERCEPTION_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DV_DESC_DIM 256        /* SuperPoint local descriptor width   (export/superpoint.py:108) */
#define DV_GLOBAL_DIM 512      /* MixVPR global descriptor width      (keyframe.cpp:264)          */
#define DV_MIX_HW 320          /* MixVPR network input side           (deep_net.cpp:1273)         */

typedef struct dv_engine dv_engine;

typedef enum dv_status {
  DV_OK = 0,
  DV_ERR_INVALID = 1,      /* bad argument / shape                                           */
  DV_ERR_CUDA = 2,         /* CUDA runtime / driver failure (sticky)                         */
  DV_ERR_UNSUPPORTED = 3,  /* configuration outside the precondition                         */
  DV_ERR_WEIGHTS = 4,      /* weight file missing / malformed / tensor missing               */
  DV_ERR_CAPACITY = 5,     /* bank / feature store / batch capacity exceeded                 */
  DV_ERR_COMM = 6,         /* NCCL failure                                                   */
  DV_ERR_NOGPU = 7         /* no Blackwell (sm_100) device: the engine refuses to run on CPU */
} dv_status;

/* Replaces the reference's compile-time #defines (deep_net.h:33-51), YAML keys (pose_graph_node.cpp:445-500)
 * and export-time constants (export/superpoint.py:110-116; keyframe.cpp:264,274,294). */
typedef struct dv_config {
  int32_t struct_size;       /* = sizeof(dv_config), ABI guard                                   */
  int32_t device;            /* CUDA device ordinal                                              */
  int32_t height, width;     /* frame size == network size, e.g. 480x752 (euroc yaml:13-14)       */
  int32_t max_batch;         /* frames (and LightGlue pairs) per batched call, >= 1              */
  int32_t max_kpts;          /* SuperPoint top-k, 512                                            */
  int32_t nms_radius;        /* 4                                                                */
  float det_thresh;          /* 0.0005                                                           */
  int32_t border;            /* 4                                                                */
  int32_t max_vio;           /* capacity for caller-supplied (VIO / window) points per frame, <=512 */
  int32_t knn_k;             /* 3  (keyframe.cpp:294)                                            */
  int32_t exclude_recent;    /* 50 (keyframe.cpp:274)                                            */
  float lg_filter_thresh;    /* 0.1                                                              */
  int32_t lg_max_kpts;       /* per-image LightGlue capacity, 1024 (README.md:180)               */
  int64_t bank_capacity;     /* rows of the global-descriptor bank                               */
  int32_t store_capacity;    /* keyframes whose local features stay device-resident (ring)       */
  int32_t world_size, rank;  /* frame-sharded ranks (1,0 = single GPU)                           */
  const char* weights_path;  /* DVWGT001 file: sp.*, lg.*, mix.* tensors (see oracle/weights.py) */
} dv_config;

void dv_config_default(dv_config* cfg);
const char* dv_last_error(void);
const char* dv_version(void);

/* Levelled logger (replaces iLogger, tensorrt_tools/ilogger.hpp:24-29 / ilogger.cpp:385-440: printf-style INFO* macros with
 * an optional file sink).  Levels: 0 off, 1 error, 2 warning, 3 info, 4 debug; default 2, or the DV_LOG environment
 * variable at load time.  Every failure reported through dv_last_error() is also logged at level 1.  Messages go to
 * stderr unless a sink is installed; the sink is called on the thread that logs. */
typedef void (*dv_log_sink)(int32_t level, const char* message, void* user);
void dv_log_set_level(int32_t level);
int32_t dv_log_get_level(void);
void dv_log_set_sink(dv_log_sink sink, void* user);

/* Factories single_init / creat_mix / creat_estimator (deep_net.cpp:1184-1203, :1445) load four TensorRT engines
 * at static-init time and return unchecked null pointers on failure; here: one explicit, checked call. */
dv_status dv_create(const dv_config* cfg, dv_engine** out);
void dv_destroy(dv_engine* e);

/* ------------------------------------------------------------------------------------------------
 * Per-keyframe API (latency path; mirrors the reference call sites 1:1).
 * ---------------------------------------------------------------------------------------------- */

/* One H2D copy of the frame, shared by SP, SP_RE and MixVPR (the reference uploads it three times:
 * deep_net.cpp:575, :748, :1294).  img: u8, `channels` 1 (gray) or 3 (BGR), row pitch `stride` bytes.  The frame is
 * staged through an internal pinned buffer and copied asynchronously: `img` may be reused as soon as the call returns. */
dv_status dv_frame_upload(dv_engine* e, const uint8_t* img, int32_t height, int32_t width, int32_t stride,
                          int32_t channels);

/* Estimator::sp_extractor(img)  deep_net.cpp:527-688.  Outputs (host, caller-allocated, capacity max_kpts):
 * kpts_xy [n,2] int32 pixel (x,y); scores [n]; desc [n,256] unit rows; kpts_norm [n,2] (may be NULL). */
dv_status dv_sp_detect(dv_engine* e, int32_t* kpts_xy, float* scores, float* desc, float* kpts_norm, int32_t* n);

/* Estimator::sp_extractor(img, kpts)  deep_net.cpp:690-812 (SP "recover"): describe at the caller's float pixel
 * keypoints.  Shares the encoder pass of the uploaded frame (the reference runs the encoder a 2nd time). */
dv_status dv_sp_describe(dv_engine* e, const float* kpts_xy, int32_t n, float* desc);

/* MixVPR::mix_extractor(img)  deep_net.cpp:1254-1323 -> 512-d unit vector. */
dv_status dv_mix_describe(dv_engine* e, float* des512);

/* KeyFrame::compute_mix_des_test's append (keyframe.cpp:353): bank row id returned in *row. */
dv_status dv_bank_append(dv_engine* e, const float* des512, int64_t* row);
dv_status dv_bank_size(dv_engine* e, int64_t* rows);
/* KeyFrame::sort_vec_faiss  keyframe.cpp:262-346: exact inner-product top-k over bank rows [0, nb_limit).
 * D [k] descending, I [k]; padded with (-inf, -1) when nb_limit < k (faiss convention). */
dv_status dv_bank_search(dv_engine* e, const float* q512, int64_t nb_limit, int32_t k, float* D, int64_t* I);
/* "next" row SURVEY §8(f)-3: flat f32 [rows,512] persistence replacing pose_graph.cpp:1042-1069 / :1156-1194. */
dv_status dv_bank_export(dv_engine* e, float* dst, int64_t max_rows, int64_t* rows);
dv_status dv_bank_import(dv_engine* e, const float* src, int64_t rows);

/* Estimator::lg_matcher(kpts0,kpts1,desc0,desc1,h0,w0,h1,w1)  deep_net.cpp:814-1000.
 * kpts in pixels; 10 <= m,n <= lg_max_kpts.  matches [K,2] int32 ascending in i0 (keyframe.cpp:623-654 relies on
 * it), mscores [K]; mkpts0/mkpts1 [K,2] de-normalised matched keypoints (may be NULL; the caller only uses their
 * count, keyframe.cpp:621-632).  Output capacity: min(m,n). */
dv_status dv_lg_match(dv_engine* e, const float* kpts0, int32_t m, const float* kpts1, int32_t n, const float* desc0,
                      const float* desc1, int32_t h0, int32_t w0, int32_t h1, int32_t w1, int32_t* matches,
                      float* mscores, float* mkpts0, float* mkpts1, int32_t* k_out);

/* ------------------------------------------------------------------------------------------------
 * Batched keyframe-round API (throughput path): b frames per rank per round, features stay
 * device-resident in a per-rank feature store keyed by global keyframe index.
 * ---------------------------------------------------------------------------------------------- */

/* imgs: b gray frames, frame i at imgs + i*frame_stride, row pitch `stride`.  Asynchronous and double-buffered: the copy is
 * queued on the engine's copy stream and the call returns at once; `imgs` (ideally pinned) must stay valid until the next
 * dv_batch_extract or dv_sync.  May be called one round ahead - after dv_batch_extract of round R, before its
 * dv_batch_commit / dv_batch_search / dv_batch_match - so the transfer of round R+1 overlaps the matching of round R. */
dv_status dv_batch_upload(dv_engine* e, int32_t b, const uint8_t* imgs, int64_t frame_stride, int32_t stride);
/* SP + SP_RE + MixVPR for the b uploaded frames.  vio_xy [b, max_vio, 2] pixel coords, n_vio [b];
 * frame_ids [b] global keyframe indices.  Local features (kpts = SP ++ VIO, desc = SP ++ SP_RE, the concatenation
 * contract of keyframe.cpp:401-432) are written into the feature store slot frame_id % store_capacity. */
dv_status dv_batch_extract(dv_engine* e, int32_t b, const float* vio_xy, const int32_t* n_vio,
                           const int64_t* frame_ids);
/* Append the b new global descriptors to the bank.  world_size > 1: ONE ncclAllGather of [b,512] per rank, landing
 * rank-major in every rank's bank (row = round_base + rank*b + i).  first_row = row of this rank's frame 0. */
dv_status dv_batch_commit(dv_engine* e, int32_t b, int64_t* first_row);
/* kNN for the b query descriptors of this round; nb_limit [b] (keyframe.cpp:274-282: index>=50 ? index-49 : index+1).
 * nb_limit == NULL: the engine applies that rule itself with cfg.exclude_recent to the rows dv_batch_commit assigned. */
dv_status dv_batch_search(dv_engine* e, int32_t b, const int64_t* nb_limit, float* D, int64_t* I);
/* LightGlue for b pairs: query = frame query_ids[i]'s window points + SP_RE descriptors (kpts0/desc0 of
 * keyframe.cpp:605-618), old = frame old_ids[i]'s full keypoint set.  A keyframe that lives in ANOTHER rank's store is
 * read in place over NVLink (CUDA-IPC mapping set up by dv_comm_init; seqlock-verified against concurrent slot reuse).
 * matches [b, max_vio, 2], mscores [b, max_vio], k_out [b]; k_out[i] = -1 marks a pair whose query or old keyframe is
 * not (or no longer) resident on any rank - the other pairs of the batch are still matched. */
dv_status dv_batch_match(dv_engine* e, int32_t b, const int64_t* query_ids, const int64_t* old_ids, int32_t* matches,
                         float* mscores, int32_t* k_out);
/* Which of a stored keyframe's points form the QUERY side of a match (the old side is always its full point set). */
#define DV_PART_WINDOW 0   /* the caller-supplied window / VIO points + SP_RE descriptors (keyframe.cpp:605-618) */
#define DV_PART_SP 1       /* its SuperPoint keypoints + descriptors (BASELINE config "SP+LG 512-kpt pair match") */
#define DV_PART_ALL 2      /* SP ++ window points */
/* dv_batch_match with a selectable query part and old part; out_cap = row capacity per pair of matches / mscores
 * (>= the largest query-side point count).  dv_batch_match == (DV_PART_WINDOW, DV_PART_ALL, max_vio). */
dv_status dv_batch_match_ex(dv_engine* e, int32_t b, const int64_t* query_ids, const int64_t* old_ids,
                            int32_t query_part, int32_t old_part, int32_t out_cap, int32_t* matches, float* mscores,
                            int32_t* k_out);
/* The same in two halves: _begin validates, queues the LightGlue pass and the result copies on the engine's stream and
 * returns; _end waits for them and fills matches [b, out_cap, 2] / mscores [b, out_cap] / k_out [b] of that call.  In
 * between the caller may upload and extract the NEXT round (dv_batch_upload / dv_batch_extract): its kernels queue right
 * behind the match, so the GPU never idles while the host collects results and prepares the next round (the reference
 * runs LightGlue synchronously inside findConnection, keyframe.cpp:583-632).  One match may be in flight; until it is
 * collected a second dv_batch_match* / dv_lg_match is refused (DV_ERR_INVALID) - they share the result staging buffers. */
dv_status dv_batch_match_begin(dv_engine* e, int32_t b, const int64_t* query_ids, const int64_t* old_ids,
                               int32_t query_part, int32_t old_part, int32_t out_cap);
dv_status dv_batch_match_end(dv_engine* e, int32_t* matches, float* mscores, int32_t* k_out);
/* MixVPR only for the b uploaded frames (BASELINE config "MixVPR + kNN"): MixVPR::mix_extractor, deep_net.cpp:1254-1323,
 * batched; follow with dv_batch_commit / dv_batch_search.  No local features are extracted or stored. */
dv_status dv_batch_describe_global(dv_engine* e, int32_t b);
/* Read one stored keyframe back (tests / persistence): kpts [n,2] f32, desc [n,256], n_sp = SuperPoint share. */
dv_status dv_store_read(dv_engine* e, int64_t frame_id, float* kpts_xy, float* desc, int32_t* n_total, int32_t* n_sp);
/* Restores a keyframe's local features into the store (the inverse of dv_store_read): rows [0, n_sp) are its SuperPoint
 * keypoints / descriptors, rows [n_sp, n_total) its window points (keyframe.cpp:401-432 concatenation order).  Together
 * with dv_bank_import this reloads a saved session - the reference's own load path (pose_graph.cpp:1156-1194) never
 * restores the deep features it saved (:1042-1069), so a reloaded map cannot close loops there. */
dv_status dv_store_put(dv_engine* e, int64_t frame_id, const float* kpts_xy, const float* desc, int32_t n_total,
                       int32_t n_sp);
/* Where does keyframe frame_id live?  owner_rank = -1 if no rank holds it (never stored, or its ring slot was reused).
 * The store is a ring of store_capacity keyframes per rank (about (max_kpts + max_vio) * 1032 bytes each): size it to the
 * number of keyframes that may still be loop candidates - the whole session at 0.7 MB per EuRoC keyframe fits HBM. */
dv_status dv_store_lookup(dv_engine* e, int64_t frame_id, int32_t* owner_rank, int32_t* n_total, int32_t* n_sp);
dv_status dv_store_lookup_many(dv_engine* e, int32_t n, const int64_t* frame_ids, int32_t* owner_rank);
/* Collective (all ranks): re-publishes every rank's slot directory, e.g. after dv_store_put / a session reload, so
 * peers can match against keyframes that never went through a round all-gather.  No-op for world_size == 1. */
dv_status dv_store_sync(dv_engine* e);
/* Results of the last dv_batch_extract for frame slot i (host copies). */
dv_status dv_batch_read_global(dv_engine* e, int32_t i, float* des512);

/* ------------------------------------------------------------------------------------------------
 * Loop decision + geometric verification (the step right after LightGlue).
 * ---------------------------------------------------------------------------------------------- */
/* Replaces the YAML keys read in pose_graph_node.cpp:480-486 and the constants of keyframe.cpp:835 / pose_graph.cpp:470. */
typedef struct dv_loop_params {
  int32_t struct_size;         /* = sizeof(dv_loop_params)                                        */
  int32_t min_loop_num;        /* MIN_LOOP_NUM 18                                                 */
  int32_t ransac_iters;        /* 200 hypotheses (keyframe.cpp:835); all are evaluated, in parallel */
  int32_t min_frame_index;     /* 50: detectLoop only fires for frame_index > 50 (pose_graph.cpp:470) */
  double pnp_inflation;        /* PNP_INFLATION 3.5: reprojection threshold = pnp_inflation / 460  */
  double max_theta_diff;       /* MAX_THETA_DIFF 40 (degrees)                                     */
  double max_pose_diff;        /* MAX_POSE_DIFF 25 (metres)                                       */
  double loop_top_thres;       /* 0.45                                                            */
  double loop_back_thres;      /* 0.40                                                            */
  double qic[9], tic[3];       /* camera -> body extrinsics (row-major rotation), keyframe.cpp:817 */
  uint64_t seed;               /* hypothesis sampling stream (counter-based, reproducible)        */
} dv_loop_params;
typedef struct dv_loop_result {
  int32_t has_loop;            /* KeyFrame::findConnection's return value                          */
  int32_t n_inliers;           /* matches surviving PnP-RANSAC                                    */
  double pnp_t_old[3], pnp_r_old[9];   /* PnP_T_old / PnP_R_old (body pose of the old keyframe, row-major) */
  double relative_t[3], relative_q[4], relative_yaw;   /* loop_info: t, q (w,x,y,z), yaw in degrees        */
} dv_loop_result;
void dv_loop_params_default(dv_loop_params* p);
/* PoseGraph::detectLoop  pose_graph.cpp:451-509: top_sim / top_sim_index = the keyframe's kNN result (dv_bank_search);
 * returns the loop candidate's keyframe index (the SMALLEST qualifying one) or -1.  Pure host arithmetic. */
int64_t dv_detect_loop(const dv_loop_params* p, const float* top_sim, const int64_t* top_sim_index, int32_t k,
                       int64_t frame_index);
/* KeyFrame::PnPRANSAC  keyframe.cpp:805-868 + the acceptance test of findConnection :1094-1183, for b pairs at once.
 * Per pair i: n_pts[i] correspondences - pts3d [b,cap,3] the CURRENT keyframe's matched 3-D points (world frame),
 * pts2d_norm [b,cap,2] the matched OLD keypoints in normalised image coordinates - and the current keyframe's VIO pose
 * vio_R [b,9] (row-major), vio_T [b,3] (the extrinsic guess).  status [b,cap] receives the inlier mask (the reference's
 * `status` vector), out [b] the poses and the loop decision.  Pairs with n_pts <= min_loop_num are skipped like the
 * reference does (has_loop = 0). */
dv_status dv_verify_loop(dv_engine* e, int32_t b, const int32_t* n_pts, int32_t cap, const double* pts3d,
                         const double* pts2d_norm, const double* vio_R, const double* vio_T, const dv_loop_params* p,
                         uint8_t* status, dv_loop_result* out);

/* Multi-GPU plumbing: the host passes the 128-byte ncclUniqueId it broadcast over its own channel
 * (torch.distributed in bench.py).  rank/world_size come from dv_config. */
dv_status dv_comm_unique_id(void* id128);
dv_status dv_comm_init(dv_engine* e, const void* id128);

/* ------------------------------------------------------------------------------------------------
 * Measurement hooks: CUDA events on the engine's own stream (torch.cuda.Event cannot see it).
 * ---------------------------------------------------------------------------------------------- */
dv_status dv_timer_start(dv_engine* e);
dv_status dv_timer_stop(dv_engine* e, float* ms);      /* records, synchronises, returns elapsed */
dv_status dv_sync(dv_engine* e);
/* Per-stage accumulated device time (ms) and kernel-launch count since the last reset.
 * stage: 0 sp_convs, 1 sp_post, 2 mixvpr, 3 knn, 4 lightglue, 5 copies. */
dv_status dv_stats_reset(dv_engine* e);
dv_status dv_stats_read(dv_engine* e, double* stage_ms6, int64_t* launches);
dv_status dv_stats_enable(dv_engine* e, int32_t on);   /* stage timing costs event records; off by default */
/* Event pair around every launch of the dominant kernel (conv1b implicit GEMM, 43 % of SuperPoint's MACs): its
 * accumulated device time and launch count inside the caller's timed region -> roofline.achieved in bench.py. */
dv_status dv_probe_enable(dv_engine* e, int32_t on);
/* which kernel the probe brackets: 0 = conv1a+conv1b implicit GEMM (default), 1 = kNN bank scan (HBM-bound). */
dv_status dv_probe_select(dv_engine* e, int32_t which);
dv_status dv_probe_read(dv_engine* e, double* ms, int64_t* launches, int32_t reset);

/* ------------------------------------------------------------------------------------------------
 * Stage-level entry points (parity tests address every kernel family in isolation through these).
 * ---------------------------------------------------------------------------------------------- */
/* D[M,N] = A[M,K] * B[N,K]^T (+bias[N]) (relu?) : fp32 host in/out, fp16 operands, fp32 accumulate on tcgen05. */
dv_status dv_dbg_gemm(dv_engine* e, const float* A, const float* B, const float* bias, int32_t M, int32_t N, int32_t K,
                      int32_t relu, float* D);
/* Same GEMM with the other fused-epilogue operands: an optional residual [M,N] (fp32 - added in place in the fp32
 * output buffer, as LightGlue's x += ffn(x) does - or rounded to fp16 when res_is_f16), and fp32 and / or fp16 outputs
 * (D16 is returned widened to fp32).  N % 8 == 0.  Either output may be NULL, not both. */
dv_status dv_dbg_gemm_ex(dv_engine* e, const float* A, const float* B, const float* bias, const float* res,
                         int32_t res_is_f16, int32_t M, int32_t N, int32_t K, int32_t relu, float* D32, float* D16);
/* 3x3 pad-1 conv on NHWC fp16 via the implicit-GEMM tcgen05 kernel: x [n,h,w,cin], wgt [cout,cin,3,3] (torch
 * layout), y [n,h',w',cout] with h' = pool ? h/2 : h. */
dv_status dv_dbg_conv3x3(dv_engine* e, const float* x, const float* wgt, const float* bias, int32_t n, int32_t h,
                         int32_t w, int32_t cin, int32_t cout, int32_t relu, int32_t pool, float* y);
/* Same op through the weights-stationary halo-tile kernel (cin = cout = 64; conv_halo.cu).  x / y are NHWC on the host;
 * the entry point converts to / from the channel-blocked device layout.  out_blocked selects the device output layout. */
dv_status dv_dbg_conv3x3_halo64(dv_engine* e, const float* x, const float* wgt, const float* bias, int32_t n, int32_t h,
                                int32_t w, int32_t relu, int32_t pool, int32_t out_blocked, float* y);
/* Same op through the 256-pixel halo-tile kernel for cin in {64, 128}, cout a multiple of 128 (conv_halo128.cu: the
 * SuperPoint conv3a..convPa/Da layers, deep_net.cpp:527-688 / export/superpoint.py:159-173). */
dv_status dv_dbg_conv3x3_halo128(dv_engine* e, const float* x, const float* wgt, const float* bias, int32_t n, int32_t h,
                                 int32_t w, int32_t cin, int32_t cout, int32_t relu, int32_t pool, int32_t out_blocked,
                                 float* y);
/* NMS + border + threshold + top-k on a caller-supplied f32 score map [h8,w8] (integer stage in isolation). */
dv_status dv_dbg_nms_select(dv_engine* e, const float* score_map, int32_t h8, int32_t w8, float* nms_out,
                            int32_t* kpts_xy, float* scores, int32_t* n);
/* Match extraction on a caller-supplied log-assignment matrix L [m,n]. */
dv_status dv_dbg_match_extract(dv_engine* e, const float* L, int32_t m, int32_t n, int32_t* matches, float* mscores,
                               int32_t* k_out);
/* Intermediate tensors of the last per-frame run, by name ("score_map", "logits", "conv1b", ...), as fp32. */
dv_status dv_dbg_read(dv_engine* e, const char* name, float* dst, int64_t capacity, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* DVINS_PERCEPTION_H_ */
