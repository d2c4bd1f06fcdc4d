// Exercises the reference-facing C++ facade in the order loop_fusion/src/keyframe.cpp uses it
// (ctor :74-81 -> computeWindowSuperpoint, computeSuperpoint, compute_mix_des_test, sort_vec_faiss; then
// light_glue_matcher :583-632) and dumps the results so the pytest driver can compare them with the C-ABI path.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../d_vins_b200/csrc/shim/deep_net_shim.h"

static std::vector<uint8_t> read_file(const char* p) {
  FILE* f = fopen(p, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", p); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> b(n);
  if (fread(b.data(), 1, n, f) != (size_t)n) exit(2);
  fclose(f);
  return b;
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: shim_demo weights frameA.raw frameB.raw vio.f32 out.txt\n"); return 2; }
  const int H = 480, W = 752;
  auto est = Estimator_net::single_init(argv[1], 0);
  auto est_w = Estimator_net::single_init(argv[1], 1);
  auto est_lg = Estimator_net::single_init(argv[1], 2);
  auto mix = MixVPR_net::creat_mix(argv[1], 0);
  if (!est || !est_w || !est_lg || !mix) { fprintf(stderr, "factory failed: %s\n", dv_last_error()); return 1; }
  std::vector<uint8_t> fa = read_file(argv[2]), fb = read_file(argv[3]), vraw = read_file(argv[4]);
  std::vector<dv::Pt> vio;
  const float* vf = reinterpret_cast<const float*>(vraw.data());
  for (size_t i = 0; i + 1 < vraw.size() / 4; i += 2) vio.emplace_back(vf[i], vf[i + 1]);
  FILE* out = fopen(argv[5], "w");
  std::vector<dv::Pt> kp_prev; std::vector<float> de_prev;
  for (int t = 0; t < 2; ++t) {
    dv::Image img{t == 0 ? fa.data() : fb.data(), H, W, 1, W};
    est_w->sp_re_desc.clear();
    if (vio.size() > 20) est_w->sp_extractor(img, vio);                 // computeWindowSuperpoint
    est->sp_kpts.clear(); est->sp_kpts_norm.clear(); est->sp_scores.clear(); est->sp_desc.clear();
    est->sp_extractor(img);                                              // computeSuperpoint
    std::vector<dv::Pt> kp = est->sp_kpts;
    kp.insert(kp.end(), vio.begin(), vio.end());
    std::vector<float> de = est->sp_desc;
    de.insert(de.end(), est_w->sp_re_desc.begin(), est_w->sp_re_desc.end());
    mix->mix_des.clear();
    mix->mix_extractor(img);                                             // compute_mix_des_test
    long row = mix->append_to_bank();
    mix->top_sim_index.clear(); mix->top_sim.clear();
    mix->sort_in_bank((int)row);                                         // sort_vec_faiss
    fprintf(out, "frame %d n_sp %zu n_total %zu mix0 %.6f row %ld top %d %.6f\n", t, est->sp_kpts.size(), kp.size(),
            mix->mix_des.empty() ? 0.f : mix->mix_des[0], row, mix->top_sim_index.empty() ? -9 : mix->top_sim_index[0],
            mix->top_sim.empty() ? 0.f : mix->top_sim[0]);
    for (size_t i = 0; i < est->sp_kpts.size(); ++i) fprintf(out, "kp %d %d\n", (int)est->sp_kpts[i].x, (int)est->sp_kpts[i].y);
    if (t == 1) {                                                        // light_glue_matcher against frame 0
      est_lg->lg_matches.clear(); est_lg->lg_scores.clear(); est_lg->lg_mkpts0.clear(); est_lg->lg_mkpts1.clear();
      est_lg->lg_matcher(vio, kp_prev, est_w->sp_re_desc, de_prev, H, W, H, W);
      fprintf(out, "matches %zu\n", est_lg->lg_mkpts0.size());
      for (size_t i = 0; i < est_lg->lg_mkpts0.size(); ++i)
        fprintf(out, "m %d %d %.6f\n", est_lg->lg_matches[2 * i], est_lg->lg_matches[2 * i + 1], est_lg->lg_scores[i]);
    }
    kp_prev = kp; de_prev = de;
  }
  fclose(out);
  return 0;
}
