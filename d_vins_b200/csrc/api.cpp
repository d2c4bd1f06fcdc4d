// Per-keyframe C-ABI entry points (latency path): one upload, lazily shared encoder pass, D2H of the results.
#include <string.h>

#include "engine.h"

using namespace dv;

#define DV_CHECK_ENGINE(e) do { if (!(e)) { dv::set_error("null engine"); return DV_ERR_INVALID; } } while (0)

namespace dv {
int ensure_encoder(Engine* e) {
  e->adopt_upload();
  if (e->cur_b <= 0) { set_error("no frame uploaded"); return DV_ERR_INVALID; }
  if (!e->enc_done) { DV_TRY(sp_run_encoder(e, e->cur_b)); e->enc_done = true; }
  return DV_OK;
}
int ensure_detect(Engine* e) {
  DV_TRY(ensure_encoder(e));
  if (!e->det_done) { DV_TRY(sp_run_detect(e, e->cur_b)); e->det_done = true; }
  return DV_OK;
}
}  // namespace dv

extern "C" {

dv_status dv_frame_upload(dv_engine* h, const uint8_t* img, int32_t height, int32_t width, int32_t stride,
                          int32_t channels) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!img || (channels != 1 && channels != 3) || stride < width * channels) { set_error("dv_frame_upload: bad image"); return DV_ERR_INVALID; }
  if (height != e->H || width != e->W) {
    // the reference resizes the image but keeps feeding original-resolution keypoints (latent bug): reject
    set_error("dv_frame_upload: frame size differs from the configured network size");
    return DV_ERR_UNSUPPORTED;
  }
  const size_t row = (size_t)width * channels;
  DV_CUDA_OK(cudaEventSynchronize(e->ev_img_ready));      // the previous upload has left the pinned staging buffer
  for (int y = 0; y < height; ++y) memcpy(e->h_img + (size_t)y * row, img + (size_t)y * stride, row);
  uint8_t* dst = e->image_begin_upload();
  DV_CUDA_OK(cudaMemcpyAsync(dst, e->h_img, row * height, cudaMemcpyHostToDevice, e->st_copy));
  e->image_end_upload();
  e->img_ch = channels;
  e->cur_b = 1;
  e->next_pending = false;
  e->enc_done = e->det_done = e->mix_done = false;
  return DV_OK;
}

dv_status dv_sp_detect(dv_engine* h, int32_t* kpts_xy, float* scores, float* desc, float* kpts_norm, int32_t* n) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!kpts_xy || !scores || !desc || !n) { set_error("dv_sp_detect: null output"); return DV_ERR_INVALID; }
  DV_TRY(ensure_detect(e));
  int *d_kp, *d_n; float *d_sc, *d_de;
  sp_device_results(e, &d_kp, nullptr, &d_sc, &d_n, &d_de, nullptr, nullptr, nullptr);
  int cnt = 0;
  {
    StageScope sc(e, ST_COPY);
    DV_CUDA_OK(cudaMemcpyAsync(&cnt, d_n, sizeof(int), cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaStreamSynchronize(e->st));
    // only the n live rows travel (the reference always copies the pre-sized 512x256 tensor, trt_tensor.cpp:400-430)
    DV_CUDA_OK(cudaMemcpyAsync(kpts_xy, d_kp, (size_t)cnt * 2 * sizeof(int), cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(scores, d_sc, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, e->st));
    DV_CUDA_OK(cudaMemcpyAsync(desc, d_de, (size_t)cnt * 256 * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  }
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  *n = cnt;
  if (kpts_norm) {
    // deep_net.cpp:633-659: integer halves, scale = max of the halves
    const float sw = (float)(e->W / 2), sh = (float)(e->H / 2);
    const float scl = sw > sh ? sw : sh;
    for (int i = 0; i < cnt; ++i) {
      kpts_norm[2 * i] = ((float)kpts_xy[2 * i] - sw) / scl;
      kpts_norm[2 * i + 1] = ((float)kpts_xy[2 * i + 1] - sh) / scl;
    }
  }
  return DV_OK;
}

dv_status dv_sp_describe(dv_engine* h, const float* kpts_xy, int32_t n, float* desc) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!kpts_xy || !desc || n < 0 || n > e->cfg.max_vio) { set_error("dv_sp_describe: n must be in [0, max_vio]"); return DV_ERR_INVALID; }
  if (n == 0) return DV_OK;
  DV_TRY(ensure_encoder(e));
  float *d_rk, *d_rd; int* d_rn;
  sp_device_results(e, nullptr, nullptr, nullptr, nullptr, nullptr, &d_rk, &d_rn, &d_rd);
  DV_CUDA_OK(cudaMemcpyAsync(d_rk, kpts_xy, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(d_rn, &n, sizeof(int), cudaMemcpyHostToDevice, e->st));
  DV_TRY(sp_run_describe(e, 1, d_rk, d_rn, e->cfg.max_vio, d_rd));
  DV_CUDA_OK(cudaMemcpyAsync(desc, d_rd, (size_t)n * 256 * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  return DV_OK;
}

dv_status dv_dbg_nms_select(dv_engine* h, const float* score_map, int32_t h8, int32_t w8, float* nms_out,
                            int32_t* kpts_xy, float* scores, int32_t* n) {
  DV_CHECK_ENGINE(h);
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!score_map || !kpts_xy || !scores || !n) { set_error("dv_dbg_nms_select: null argument"); return DV_ERR_INVALID; }
  return (dv_status)sp_nms_select_dbg(e, score_map, h8, w8, nms_out, kpts_xy, scores, n);
}

}  // extern "C"
