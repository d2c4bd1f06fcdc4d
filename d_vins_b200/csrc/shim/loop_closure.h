// ROS-free keyframe stream driver over the C ABI: the deep-perception + loop-decision half of
// loop_fusion's `process()` thread with the reference's message shapes on the input side.
//
//   pose_graph_node.cpp:330-388   process(): image / pose / point-cloud messages -> KeyFrame(...)
//   vins_estimator/src/utility/visualization.cpp:399-429   pubKeyframe(): PointCloud layout - points[i] = world xyz,
//                                 channels[i].values = [norm_x, norm_y, u, v, id]
//   keyframe.cpp:49-87            KeyFrame ctor: SP_RE -> SP -> MixVPR -> kNN
//   pose_graph.cpp:71-170         addKeyFrame(): detectLoop -> findConnection -> loop_info
//   keyframe.cpp:871-1186         findConnection(): LightGlue, reduceVector, PnPRANSAC, acceptance gates
//   camera_models/.../PinholeCamera.cc:450-505   liftProjective (recursive radtan undistortion, 8 iterations)
//
// What stays outside (out of scope): the pose-graph optimisation thread, drift correction, ROS publishers.
// The struct types below carry exactly the fields `process()` reads from sensor_msgs / nav_msgs, so a loop_fusion build
// forwards its messages field by field.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../../include/dvins_perception.h"

namespace dv {

struct ImageMsg {            // sensor_msgs::Image (MONO8 / 8UC1)
  double stamp = 0.0;
  int height = 0, width = 0, step = 0;
  const uint8_t* data = nullptr;
};
struct PoseMsg {             // nav_msgs::Odometry: pose.pose.position / orientation (w, x, y, z)
  double stamp = 0.0;
  double position[3] = {0, 0, 0};
  double orientation[4] = {1, 0, 0, 0};
};
struct Point32 { float x, y, z; };
struct ChannelFloat32 { std::vector<float> values; };
struct PointCloudMsg {       // sensor_msgs::PointCloud as published by pubKeyframe
  double stamp = 0.0;
  std::vector<Point32> points;               // world-frame 3-D points of the tracked features
  std::vector<ChannelFloat32> channels;      // per point: [norm_x, norm_y, u, v, id]
};

// camodocal PinholeCamera (the model every D_VINS EuRoC / KITTI config uses)
struct PinholeCamera {
  double fx = 1, fy = 1, cx = 0, cy = 0, k1 = 0, k2 = 0, p1 = 0, p2 = 0;
  void liftProjective(double u, double v, double* xn, double* yn) const;    // PinholeCamera.cc:450-505
};

struct LoopResult {
  int index = -1;                 // this keyframe's index
  int n_sp = 0, n_window = 0;     // SuperPoint keypoints, window (VIO) points
  float top_sim[3] = {0, 0, 0};   // KeyFrame::top_sim / top_sim_index (sort_vec_faiss)
  int64_t top_sim_index[3] = {-1, -1, -1};
  int64_t loop_candidate = -1;    // PoseGraph::detectLoop
  int n_matches = 0;              // LightGlue matches against the candidate
  int n_inliers = 0;              // after PnP-RANSAC
  bool has_loop = false;          // KeyFrame::findConnection
  double loop_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // relative_t (3), relative_q w,x,y,z (4), relative_yaw
};

class LoopClosure {
 public:
  // `engine` must have been created for the stream's image size with max_vio >= the largest window-point count and a
  // store_capacity covering every keyframe that may still become a loop candidate.  use_sp = the USE_SP YAML switch.
  LoopClosure(dv_engine* engine, const PinholeCamera& cam, const dv_loop_params& params, int max_vio, int max_kpts,
              bool use_sp = true, double skip_dis = 0.0);
  // One iteration of process() (pose_graph_node.cpp:264-397) for a time-aligned (image, pose, points) triple.
  // Returns DV_OK and fills `out`; *is_keyframe = false when the SKIP_DIS test dropped the frame.
  dv_status process(const ImageMsg& img, const PoseMsg& pose, const PointCloudMsg& pts, LoopResult* out, bool* is_keyframe);
  int keyframes() const { return frame_index_; }

 private:
  struct Kf {                       // what findConnection needs from an OLD keyframe on the host
    std::vector<float> kpts;        // [n,2] pixel keypoints: SuperPoint ++ window points (keyframe.cpp:401-432)
  };
  dv_engine* e_;
  PinholeCamera cam_;
  dv_loop_params prm_;
  int max_vio_, max_kpts_;
  bool use_sp_;
  double skip_dis_;
  int frame_index_ = 0;
  bool have_last_ = false;
  double last_t_[3] = {0, 0, 0};
  std::map<int64_t, Kf> kfs_;
  std::vector<uint8_t> img_buf_;
  std::vector<float> vio_buf_;
};

}  // namespace dv
