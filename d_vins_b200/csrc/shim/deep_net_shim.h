// C++ host mirror of the reference's deep-perception facade, implemented over the C ABI (dvins_perception.h).
//
// Re-declares Estimator_net::Estimator / MixVPR_net::MixVPR with the SAME public result members and method names as
// loop_fusion/src/deep_net/deep_net.h:115-177 so the call sites in loop_fusion/src/keyframe.cpp (:348-354, :356-378,
// :380-399, :583-632) port 1:1.  Differences, all deliberate:
//   * images are passed as dv::Image (pointer + rows/cols/step/channels); define DV_SHIM_WITH_OPENCV before including
//     this header to get cv::Mat / cv::Point2f overloads identical to the reference signatures;
//   * factories return nullptr *and* leave a message in dv_last_error() (the reference never checks its factories);
//   * all four reference singletons (SP, SP_RE, LG, MixVPR; keyframe.cpp:25-28) share ONE dv_engine, so the frame is
//     uploaded once and the SuperPoint encoder runs once per keyframe;
//   * the kNN lives behind MixVPR::sort_in_bank(index) instead of a per-query faiss rebuild (keyframe.cpp:262-346).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/dvins_perception.h"

#ifdef DV_SHIM_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace dv {

struct Point2f {
  float x = 0.f, y = 0.f;
  Point2f() = default;
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct Image {
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0, channels = 1, step = 0;   // step = row pitch in bytes
};

#ifdef DV_SHIM_WITH_OPENCV
using Pt = cv::Point2f;
inline Image as_image(const cv::Mat& m) { return Image{m.data, m.rows, m.cols, m.channels(), (int)m.step}; }
#else
using Pt = Point2f;
#endif

// One engine per (process, GPU); shared by the estimator / mix facades below.
std::shared_ptr<dv_engine> shared_engine(const std::string& weights_path, int height, int width, int gpuid = 0);

}  // namespace dv

namespace MixVPR_net {
using namespace std;
class MixVPR {
 public:
  MixVPR() = default;
  virtual ~MixVPR() = default;
  virtual void mix_extractor(const dv::Image& img) = 0;                 // deep_net.h:121
#ifdef DV_SHIM_WITH_OPENCV
  void mix_extractor(const cv::Mat& img) { mix_extractor(dv::as_image(img)); }
#endif
  // keyframe.cpp:262-346 (sort_vec_faiss): appends nothing; searches bank rows [0, index-50] (or [0,index] if
  // index < 50) for the current mix_des and fills top_sim_index / top_sim.
  virtual void sort_in_bank(int index) = 0;
  // deep_net.h:123 declares (db, xq, n) but the implementation (deep_net.cpp:1325) and its only caller (:1399) pass
  // the QUERY first: sort_in_faiss(xq, db, nb).  Behaviour follows the implementation: import db [nb,512], search xq,
  // append the top-3 to top_sim_index / top_sim and to sim_map[call counter].
  virtual void sort_in_faiss(float* xq, float* db, int nb) = 0;
  // deep_net.h:122 / deep_net.cpp:1386-1428: run every image of a folder through mix_extractor + sort_in_faiss (the
  // newest 15 excluded) filling descriptors_database / sim_map / top_sim.  The reference then cv::imshow()s the top-3;
  // here the results stay in the members.  Without OpenCV the folder is scanned for binary PGM (P5) files.
  virtual void test_in_dataset(const std::string filepath) = 0;
  virtual long append_to_bank() = 0;                                    // keyframe.cpp:353

  std::vector<float> descriptors_database;   // kept for source compatibility; the live bank is device-resident
  std::vector<float> mix_des;                // 512
  std::vector<int> top_sim_index;
  std::vector<float> top_sim;
  std::map<int, vector<int>> sim_map;        // deep_net.h:134
};
shared_ptr<MixVPR> creat_mix(const std::string& weights_path, const int& engine_type, int gpuid = 0, int height = 480,
                             int width = 752);
}  // namespace MixVPR_net

namespace Estimator_net {
using namespace std;
class Estimator {
 public:
  Estimator() = default;
  virtual ~Estimator() = default;
  virtual void sp_extractor(const dv::Image& img) = 0;                                   // deep_net.h:144
  virtual void sp_extractor(const dv::Image& img, vector<dv::Pt>& sp_kpts) = 0;          // deep_net.h:145
  virtual void lg_matcher(std::vector<dv::Pt>& lg_in_kpts0, std::vector<dv::Pt>& lg_in_kpts1,
                          std::vector<float>& lg_in_desc0, std::vector<float>& lg_in_desc1, const int& height_0,
                          const int& width_0, const int& height_1, const int& width_1) = 0;   // deep_net.h:146-151
  // deep_net.h:153 / deep_net.cpp:1003-1133 (legacy, no caller in the reference): LightGlue on the keypoints the last
  // sp_extractor(img) produced, used for BOTH sides (the SuperPoint output tensors are bound as kpts0 and kpts1).
  virtual void lg_matcher() = 0;
#ifdef DV_SHIM_WITH_OPENCV
  void sp_extractor(const cv::Mat& img) { sp_extractor(dv::as_image(img)); }
  void sp_extractor(const cv::Mat& img, vector<cv::Point2f>& k) { sp_extractor(dv::as_image(img), k); }
#endif

  int width = 0;
  int height = 0;
  vector<float> sp_desc;
  vector<float> sp_scores;
  vector<dv::Pt> sp_kpts_norm;
  vector<dv::Pt> sp_kpts;

  vector<float> sp_re_desc;
  vector<float> sp_re_scores;

  vector<int> lg_matches;
  vector<float> lg_scores;
  vector<dv::Pt> lg_mkpts0;
  vector<dv::Pt> lg_mkpts1;
#ifdef DV_SHIM_WITH_OPENCV
  cv::Mat image;                             // deep_net.h:169
#else
  dv::Image image;                           // deep_net.h:169 (non-owning view of the last frame)
#endif
};

// engine_type 0: SP, 1: SP_RE, 2: LG (deep_net.cpp:1198-1203) - all three views share one dv_engine.
shared_ptr<Estimator> single_init(const std::string& weights_path, const int& engine_type, int gpuid = 0,
                                  int height = 480, int width = 752);
shared_ptr<Estimator> creat_estimator(const std::string& weights_path, const std::string& unused_lg_path, int gpuid = 0);
shared_ptr<Estimator> recover_estimator(const std::string& weights_path, int gpuid = 0);
}  // namespace Estimator_net
