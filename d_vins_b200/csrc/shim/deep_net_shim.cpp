// Implementation of the reference-facing C++ facade over the C ABI.  Mirrors the caller contract of
// loop_fusion/src/keyframe.cpp: the caller clear()s the result vectors, calls, then copies them out.
#include "deep_net_shim.h"

#include <map>
#include <mutex>
#include <tuple>

namespace dv {

std::shared_ptr<dv_engine> shared_engine(const std::string& weights_path, int height, int width, int gpuid) {
  static std::mutex mu;
  static std::map<std::tuple<std::string, int, int, int>, std::weak_ptr<dv_engine>> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple(weights_path, height, width, gpuid);
  if (auto sp = cache[key].lock()) return sp;
  dv_config cfg;
  dv_config_default(&cfg);
  cfg.device = gpuid;
  cfg.height = height;
  cfg.width = width;
  cfg.weights_path = weights_path.c_str();
  dv_engine* e = nullptr;
  if (dv_create(&cfg, &e) != DV_OK) return nullptr;   // message in dv_last_error()
  std::shared_ptr<dv_engine> sp(e, [](dv_engine* p) { dv_destroy(p); });
  cache[key] = sp;
  return sp;
}

}  // namespace dv

namespace {

// The frame most recently uploaded to an engine: SP, SP_RE and MixVPR are called back to back on the same image
// (keyframe.cpp:74-81), so the upload happens once.
struct UploadCache {
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0;
};
std::map<dv_engine*, UploadCache>& upload_cache() {
  static std::map<dv_engine*, UploadCache> m;
  return m;
}
bool ensure_uploaded(dv_engine* e, const dv::Image& img, bool force) {
  UploadCache& c = upload_cache()[e];
  if (!force && c.data == img.data && c.rows == img.rows && c.cols == img.cols) return true;
  if (dv_frame_upload(e, img.data, img.rows, img.cols, img.step ? img.step : img.cols * img.channels, img.channels) != DV_OK)
    return false;
  c = UploadCache{img.data, img.rows, img.cols};
  return true;
}

class EstimatorImpl : public Estimator_net::Estimator {
 public:
  explicit EstimatorImpl(std::shared_ptr<dv_engine> e) : e_(std::move(e)) {}
  void sp_extractor(const dv::Image& img) override {        // deep_net.cpp:527-688
    width = img.cols; height = img.rows;
    if (!ensure_uploaded(e_.get(), img, /*force=*/true)) return;
    std::vector<int32_t> kp(512 * 2);
    std::vector<float> sc(512), de(512 * 256), kn(512 * 2);
    int32_t n = 0;
    if (dv_sp_detect(e_.get(), kp.data(), sc.data(), de.data(), kn.data(), &n) != DV_OK) return;
    for (int i = 0; i < n; ++i) {                             // the reference's push_back loop (deep_net.cpp:650-662)
      sp_kpts.emplace_back((float)kp[2 * i], (float)kp[2 * i + 1]);
      sp_kpts_norm.emplace_back(kn[2 * i], kn[2 * i + 1]);
      sp_scores.push_back(sc[i]);
    }
    sp_desc.insert(sp_desc.end(), de.begin(), de.begin() + (size_t)n * 256);
  }
  void sp_extractor(const dv::Image& img, std::vector<dv::Pt>& pts) override {   // deep_net.cpp:690-812
    width = img.cols; height = img.rows;
    if (!ensure_uploaded(e_.get(), img, /*force=*/false)) return;
    const int n = (int)pts.size();
    std::vector<float> k((size_t)n * 2), de((size_t)n * 256);
    for (int i = 0; i < n; ++i) { k[2 * i] = pts[i].x; k[2 * i + 1] = pts[i].y; }
    if (dv_sp_describe(e_.get(), k.data(), n, de.data()) != DV_OK) return;
    sp_re_desc.insert(sp_re_desc.end(), de.begin(), de.end());
    sp_re_scores.assign((size_t)n, 0.f);                      // export/ultrapoint.py:117: scores are zeros
  }
  void lg_matcher(std::vector<dv::Pt>& k0, std::vector<dv::Pt>& k1, std::vector<float>& d0, std::vector<float>& d1,
                  const int& h0, const int& w0, const int& h1, const int& w1) override {   // deep_net.cpp:814-1000
    const int m = (int)k0.size(), n = (int)k1.size();
    std::vector<float> a((size_t)m * 2), b((size_t)n * 2);
    for (int i = 0; i < m; ++i) { a[2 * i] = k0[i].x; a[2 * i + 1] = k0[i].y; }
    for (int i = 0; i < n; ++i) { b[2 * i] = k1[i].x; b[2 * i + 1] = k1[i].y; }
    const int cap = m < n ? m : n;
    std::vector<int32_t> ma((size_t)cap * 2);
    std::vector<float> ms(cap), mk0((size_t)cap * 2), mk1((size_t)cap * 2);
    int32_t k = 0;
    if (dv_lg_match(e_.get(), a.data(), m, b.data(), n, d0.data(), d1.data(), h0, w0, h1, w1, ma.data(), ms.data(),
                    mk0.data(), mk1.data(), &k) != DV_OK)
      return;
    lg_matches.insert(lg_matches.end(), ma.begin(), ma.begin() + (size_t)k * 2);   // [i0, i1] pairs, i0 ascending
    lg_scores.insert(lg_scores.end(), ms.begin(), ms.begin() + k);
    for (int i = 0; i < k; ++i) {
      lg_mkpts0.emplace_back(mk0[2 * i], mk0[2 * i + 1]);
      lg_mkpts1.emplace_back(mk1[2 * i], mk1[2 * i + 1]);
    }
  }

 private:
  std::shared_ptr<dv_engine> e_;
};

class MixVPRImpl : public MixVPR_net::MixVPR {
 public:
  explicit MixVPRImpl(std::shared_ptr<dv_engine> e) : e_(std::move(e)) {}
  void mix_extractor(const dv::Image& img) override {        // deep_net.cpp:1254-1323
    if (!ensure_uploaded(e_.get(), img, /*force=*/false)) return;
    float d[512];
    if (dv_mix_describe(e_.get(), d) != DV_OK) return;
    mix_des.assign(d, d + 512);
  }
  long append_to_bank() override {                            // keyframe.cpp:353
    int64_t row = -1;
    if (mix_des.size() != 512 || dv_bank_append(e_.get(), mix_des.data(), &row) != DV_OK) return -1;
    return (long)row;
  }
  void sort_in_bank(int index) override {                     // keyframe.cpp:262-346
    const int64_t nb = index >= 50 ? index - 49 : index + 1;
    float D[3]; int64_t I[3];
    if (mix_des.size() != 512 || dv_bank_search(e_.get(), mix_des.data(), nb, 3, D, I) != DV_OK) return;
    for (int j = 0; j < 3; ++j) { top_sim_index.push_back((int)I[j]); top_sim.push_back(D[j]); }
  }
  void sort_in_faiss(float* db, float* xq, int n) override {  // deep_net.cpp:1325-1384
    float D[3]; int64_t I[3];
    if (dv_bank_import(e_.get(), db, n) != DV_OK || dv_bank_search(e_.get(), xq, n, 3, D, I) != DV_OK) return;
    for (int j = 0; j < 3; ++j) { top_sim_index.push_back((int)I[j]); top_sim.push_back(D[j]); }
  }

 private:
  std::shared_ptr<dv_engine> e_;
};

}  // namespace

namespace MixVPR_net {
shared_ptr<MixVPR> creat_mix(const std::string& weights_path, const int&, int gpuid, int height, int width) {
  auto e = dv::shared_engine(weights_path, height, width, gpuid);
  if (!e) return nullptr;
  return std::make_shared<MixVPRImpl>(e);
}
}  // namespace MixVPR_net

namespace Estimator_net {
shared_ptr<Estimator> single_init(const std::string& weights_path, const int& engine_type, int gpuid, int height, int width) {
  if (engine_type < 0 || engine_type > 2) return nullptr;    // the reference falls off the end here (UB, deep_net.cpp:522-525)
  auto e = dv::shared_engine(weights_path, height, width, gpuid);
  if (!e) return nullptr;
  return std::make_shared<EstimatorImpl>(e);
}
shared_ptr<Estimator> creat_estimator(const std::string& weights_path, const std::string&, int gpuid) {
  return single_init(weights_path, 0, gpuid);
}
shared_ptr<Estimator> recover_estimator(const std::string& weights_path, int gpuid) {
  return single_init(weights_path, 1, gpuid);
}
}  // namespace Estimator_net
