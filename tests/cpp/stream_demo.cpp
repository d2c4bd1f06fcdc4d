// Drives the ROS-free stream driver (csrc/shim/loop_closure.h) with a recorded keyframe stream, in the message shapes
// loop_fusion's process() consumes (pose_graph_node.cpp:330-388), and dumps one line per keyframe so the pytest driver
// can compare it with a step-by-step replication through the C ABI + the oracle.
//
// stream file (little endian): i32 H, W, T | per frame: f64 stamp, f64 pos[3], f64 quat_wxyz[4], i32 n,
//   n x { f32 xyz[3], f32 ch[5] = norm_x, norm_y, u, v, id }, H*W u8 pixels
// params file: f64 fx, fy, cx, cy, k1, k2, p1, p2, qic[9], tic[3], top_thres, back_thres, pnp_inflation, i32 min_loop_num
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../d_vins_b200/csrc/shim/loop_closure.h"

template <class T> static bool rd(FILE* f, T* v, size_t n = 1) { return fread(v, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: stream_demo weights stream.bin params.bin out.txt\n"); return 2; }
  FILE* fs = fopen(argv[2], "rb");
  FILE* fp = fopen(argv[3], "rb");
  if (!fs || !fp) { fprintf(stderr, "cannot open inputs\n"); return 2; }
  int32_t H, W, T;
  if (!rd(fs, &H) || !rd(fs, &W) || !rd(fs, &T)) return 2;
  dv::PinholeCamera cam;
  dv_loop_params prm;
  dv_loop_params_default(&prm);
  double c8[8], thr[3]; int32_t mln;
  if (!rd(fp, c8, 8) || !rd(fp, prm.qic, 9) || !rd(fp, prm.tic, 3) || !rd(fp, thr, 3) || !rd(fp, &mln)) return 2;
  cam.fx = c8[0]; cam.fy = c8[1]; cam.cx = c8[2]; cam.cy = c8[3]; cam.k1 = c8[4]; cam.k2 = c8[5]; cam.p1 = c8[6]; cam.p2 = c8[7];
  prm.loop_top_thres = thr[0]; prm.loop_back_thres = thr[1]; prm.pnp_inflation = thr[2]; prm.min_loop_num = mln;
  fclose(fp);
  dv_config cfg;
  dv_config_default(&cfg);
  cfg.height = H; cfg.width = W; cfg.max_batch = 1; cfg.max_vio = 64; cfg.store_capacity = T + 4; cfg.bank_capacity = T + 4;
  cfg.weights_path = argv[1];
  dv_engine* e = nullptr;
  if (dv_create(&cfg, &e) != DV_OK) { fprintf(stderr, "dv_create: %s\n", dv_last_error()); return 1; }
  dv::LoopClosure lc(e, cam, prm, cfg.max_vio, cfg.max_kpts);
  FILE* out = fopen(argv[4], "w");
  std::vector<uint8_t> px((size_t)H * W);
  for (int t = 0; t < T; ++t) {
    dv::ImageMsg im; dv::PoseMsg po; dv::PointCloudMsg pc;
    int32_t n;
    if (!rd(fs, &po.stamp) || !rd(fs, po.position, 3) || !rd(fs, po.orientation, 4) || !rd(fs, &n)) return 2;
    pc.points.resize(n); pc.channels.resize(n);
    for (int i = 0; i < n; ++i) {
      float v[8];
      if (!rd(fs, v, 8)) return 2;
      pc.points[i] = {v[0], v[1], v[2]};
      pc.channels[i].values.assign(v + 3, v + 8);
    }
    if (!rd(fs, px.data(), px.size())) return 2;
    im.stamp = po.stamp; im.height = H; im.width = W; im.step = W; im.data = px.data();
    dv::LoopResult r; bool kf = false;
    const dv_status rc = lc.process(im, po, pc, &r, &kf);
    if (rc != DV_OK) { fprintf(stderr, "process(%d): status %d: %s\n", t, (int)rc, dv_last_error()); return 1; }
    if (!kf) { fprintf(out, "skip %d\n", t); continue; }
    fprintf(out, "kf %d n_sp %d n_win %d top %lld %lld %lld sim %.9g %.9g %.9g cand %lld matches %d inliers %d loop %d info", r.index,
            r.n_sp, r.n_window, (long long)r.top_sim_index[0], (long long)r.top_sim_index[1], (long long)r.top_sim_index[2],
            r.top_sim[0], r.top_sim[1], r.top_sim[2], (long long)r.loop_candidate, r.n_matches, r.n_inliers, r.has_loop ? 1 : 0);
    for (int q = 0; q < 8; ++q) fprintf(out, " %.17g", r.loop_info[q]);
    fprintf(out, "\n");
  }
  fclose(out);
  fclose(fs);
  dv_destroy(e);
  return 0;
}
