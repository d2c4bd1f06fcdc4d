// ROS-free keyframe stream driver over the C ABI - see loop_closure.h for the reference call sites it mirrors.
#include "loop_closure.h"

#include <math.h>
#include <string.h>

namespace dv {

// PinholeCamera::liftProjective (camera_models/src/camera_models/PinholeCamera.cc:450-505): lift to the normalised
// plane, then the recursive radtan undistortion (8 fixed-point iterations of :646-662).
void PinholeCamera::liftProjective(double u, double v, double* xn, double* yn) const {
  const double mx_d = (1.0 / fx) * u + (-cx / fx);
  const double my_d = (1.0 / fy) * v + (-cy / fy);
  if (k1 == 0.0 && k2 == 0.0 && p1 == 0.0 && p2 == 0.0) { *xn = mx_d; *yn = my_d; return; }   // m_noDistortion
  auto distortion = [&](double x, double y, double* dx, double* dy) {
    const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
    const double rad = k1 * rho2 + k2 * rho2 * rho2;
    *dx = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2);
    *dy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2);
  };
  double dx, dy;
  distortion(mx_d, my_d, &dx, &dy);
  double mx_u = mx_d - dx, my_u = my_d - dy;
  for (int i = 1; i < 8; ++i) {
    distortion(mx_u, my_u, &dx, &dy);
    mx_u = mx_d - dx;
    my_u = my_d - dy;
  }
  *xn = mx_u; *yn = my_u;
}

LoopClosure::LoopClosure(dv_engine* engine, const PinholeCamera& cam, const dv_loop_params& params, int max_vio,
                         int max_kpts, bool use_sp, double skip_dis)
    : e_(engine), cam_(cam), prm_(params), max_vio_(max_vio), max_kpts_(max_kpts), use_sp_(use_sp), skip_dis_(skip_dis) {
  last_t_[0] = last_t_[1] = last_t_[2] = -100.0;      // pose_graph_node.cpp:56
  vio_buf_.assign((size_t)max_vio_ * 2, 0.f);
}

// Eigen::Quaterniond(w, x, y, z).toRotationMatrix(), row-major
static void quat_to_rot(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
               tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

dv_status LoopClosure::process(const ImageMsg& img, const PoseMsg& pose, const PointCloudMsg& pts, LoopResult* out,
                               bool* is_keyframe) {
  if (!out || !is_keyframe || !img.data) return DV_ERR_INVALID;
  *out = LoopResult();
  *is_keyframe = false;
  // ---- pose_graph_node.cpp:347-355: pose, SKIP_DIS gate
  const double* T = pose.position;
  double R[9];
  quat_to_rot(pose.orientation, R);
  const double d0 = T[0] - last_t_[0], d1 = T[1] - last_t_[1], d2 = T[2] - last_t_[2];
  if (!(sqrt(d0 * d0 + d1 * d1 + d2 * d2) > skip_dis_)) return DV_OK;
  // ---- :357-382: unpack the point cloud (xyz + channels [norm_x, norm_y, u, v, id])
  const int n = (int)pts.points.size();
  if ((int)pts.channels.size() != n) return DV_ERR_INVALID;
  if (n > max_vio_) return DV_ERR_CAPACITY;
  std::vector<double> p3((size_t)n * 3);
  for (int i = 0; i < n; ++i) {
    if (pts.channels[i].values.size() < 5) return DV_ERR_INVALID;
    p3[3 * i] = pts.points[i].x; p3[3 * i + 1] = pts.points[i].y; p3[3 * i + 2] = pts.points[i].z;
    vio_buf_[2 * i] = pts.channels[i].values[2];
    vio_buf_[2 * i + 1] = pts.channels[i].values[3];
  }
  const int64_t index = frame_index_;
  out->index = (int)index;
  // ---- KeyFrame ctor (keyframe.cpp:74-81): SP_RE (only for > 20 window points, :373) -> SP -> MixVPR -> kNN.  One
  // upload, one encoder pass; features go straight into the device-resident store.
  const int32_t n_vio = n > 20 ? n : 0;
  const uint8_t* src = img.data;
  if (img.step != img.width) {                               // cv_bridge hands out a tightly packed MONO8 clone
    img_buf_.resize((size_t)img.height * img.width);
    for (int y = 0; y < img.height; ++y) memcpy(&img_buf_[(size_t)y * img.width], img.data + (size_t)y * img.step, (size_t)img.width);
    src = img_buf_.data();
  }
  dv_status rc;
  if ((rc = dv_batch_upload(e_, 1, src, (int64_t)img.height * img.width, img.width)) != DV_OK) return rc;
  if ((rc = dv_batch_extract(e_, 1, vio_buf_.data(), &n_vio, &index)) != DV_OK) return rc;
  int64_t row = -1;
  if ((rc = dv_batch_commit(e_, 1, &row)) != DV_OK) return rc;
  if (row != index) return DV_ERR_INVALID;                   // bank row == keyframe index (keyframe.cpp:353)
  float D[8]; int64_t I[8];
  if ((rc = dv_batch_search(e_, 1, nullptr, D, I)) != DV_OK) return rc;      // sort_vec_faiss, keyframe.cpp:262-346
  for (int j = 0; j < 3; ++j) { out->top_sim[j] = D[j]; out->top_sim_index[j] = I[j]; }
  Kf kf;
  kf.kpts.assign((size_t)(max_kpts_ + max_vio_) * 2, 0.f);
  int32_t n_total = 0, n_sp = 0;
  if ((rc = dv_store_read(e_, index, kf.kpts.data(), nullptr, &n_total, &n_sp)) != DV_OK) return rc;
  kf.kpts.resize((size_t)n_total * 2);
  out->n_sp = n_sp; out->n_window = n_vio;
  // ---- PoseGraph::addKeyFrame (pose_graph.cpp:95-100): detectLoop
  const int64_t cand = dv_detect_loop(&prm_, D, I, 3, index);
  out->loop_candidate = cand;
  auto old_it = cand >= 0 ? kfs_.find(cand) : kfs_.end();
  // ---- KeyFrame::findConnection (keyframe.cpp:871-1186)
  if (old_it != kfs_.end()) {
    const Kf& old = old_it->second;
    const int old_first = use_sp_ ? 0 : -1;                  // USE_SP = 0: the old keyframe offers its window points only
    const int n_cur_kpts = use_sp_ ? n_total : n_vio;
    int32_t n_old_total = 0, n_old_sp = 0, owner = -1;
    dv_store_lookup(e_, cand, &owner, &n_old_total, &n_old_sp);
    const int n_old = use_sp_ ? n_old_total : n_old_total - n_old_sp;
    if (n > 20 && n_cur_kpts > 20 && n_old > 20 && owner >= 0) {          // keyframe.cpp:935
      std::vector<int32_t> m((size_t)max_vio_ * 2);
      std::vector<float> ms((size_t)max_vio_);
      int32_t k = 0;
      rc = dv_batch_match_ex(e_, 1, &index, &cand, DV_PART_WINDOW, use_sp_ ? DV_PART_ALL : DV_PART_WINDOW, max_vio_,
                             m.data(), ms.data(), &k);
      if (rc != DV_OK) return rc;
      if (k > 0) {
        out->n_matches = k;
        // light_glue_matcher (:623-654): status[i0] = 1, matched_2d_old(_norm) pushed in MATCH order; reduceVector then
        // filters matched_3d in INDEX order - the two line up because the pairs arrive ascending in i0
        std::vector<double> X((size_t)k * 3), U((size_t)k * 2);
        const int off = (old_first < 0) ? n_old_sp : 0;
        for (int q = 0; q < k; ++q) {
          const int i0 = m[2 * q], j = m[2 * q + 1] + off;
          X[3 * q] = p3[3 * i0]; X[3 * q + 1] = p3[3 * i0 + 1]; X[3 * q + 2] = p3[3 * i0 + 2];
          double xn, yn;
          cam_.liftProjective(old.kpts[2 * j], old.kpts[2 * j + 1], &xn, &yn);     // keyframe.cpp:887-894
          U[2 * q] = (double)(float)xn; U[2 * q + 1] = (double)(float)yn;           // cv::Point2f
        }
        if (k > prm_.min_loop_num) {                                                // keyframe.cpp:1094
          std::vector<uint8_t> status((size_t)k);
          dv_loop_result lr;
          rc = dv_verify_loop(e_, 1, &k, k, X.data(), U.data(), R, T, &prm_, status.data(), &lr);
          if (rc != DV_OK) return rc;
          out->n_inliers = lr.n_inliers;
          out->has_loop = lr.has_loop != 0;
          if (out->has_loop) {                                                      // keyframe.cpp:1176-1180: loop_info
            for (int q = 0; q < 3; ++q) out->loop_info[q] = lr.relative_t[q];
            for (int q = 0; q < 4; ++q) out->loop_info[3 + q] = lr.relative_q[q];
            out->loop_info[7] = lr.relative_yaw;
          }
        }
      }
    }
  }
  kfs_[index] = std::move(kf);
  last_t_[0] = T[0]; last_t_[1] = T[1]; last_t_[2] = T[2];
  ++frame_index_;
  *is_keyframe = true;
  return DV_OK;
}

}  // namespace dv
