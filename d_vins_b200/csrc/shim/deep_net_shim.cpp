// Implementation of the reference-facing C++ facade over the C ABI.  Mirrors the caller contract of
// loop_fusion/src/keyframe.cpp: the caller clear()s the result vectors, calls, then copies them out.
#include "deep_net_shim.h"

#include <dirent.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

namespace dv {

void shim_forget_engine(dv_engine* e);     // drops the upload-cache entry of a dying engine (defined below)

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
  std::shared_ptr<dv_engine> sp(e, [](dv_engine* p) { dv::shim_forget_engine(p); dv_destroy(p); });
  cache[key] = sp;
  return sp;
}

}  // namespace dv

namespace {

// The frame most recently uploaded to an engine: SP_RE, SP and MixVPR are called back to back on the same image
// (keyframe.cpp:74-81), so the upload happens once per keyframe.  "Same image" is decided by CONTENT (a 64-bit hash of
// every pixel + the geometry), never by the buffer address: a new keyframe's cv::Mat may well land at the address of
// the previous one (allocator reuse), and the first call of a keyframe is the SP_RE overload.  ~30 us per 480x752 frame.
struct UploadCache {
  uint64_t hash = 0;
  int rows = 0, cols = 0, channels = 0;
  bool valid = false;
};
std::mutex& upload_mu() { static std::mutex m; return m; }
std::map<dv_engine*, UploadCache>& upload_cache() {
  static std::map<dv_engine*, UploadCache> m;
  return m;
}
uint64_t frame_hash(const dv::Image& img) {
  uint64_t h = 0xcbf29ce484222325ull;
  const size_t row = (size_t)img.cols * img.channels;
  const int step = img.step ? img.step : (int)row;
  for (int y = 0; y < img.rows; ++y) {
    const uint8_t* p = img.data + (size_t)y * step;
    size_t i = 0;
    for (; i + 8 <= row; i += 8) {
      uint64_t w;
      memcpy(&w, p + i, 8);
      h = (h ^ w) * 0x100000001b3ull;
      h ^= h >> 29;
    }
    for (; i < row; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  }
  return h;
}
bool ensure_uploaded(dv_engine* e, const dv::Image& img, bool force) {
  const uint64_t hsh = frame_hash(img);
  std::lock_guard<std::mutex> lk(upload_mu());
  UploadCache& c = upload_cache()[e];
  if (!force && c.valid && c.hash == hsh && c.rows == img.rows && c.cols == img.cols && c.channels == img.channels) return true;
  c.valid = false;
  if (dv_frame_upload(e, img.data, img.rows, img.cols, img.step ? img.step : img.cols * img.channels, img.channels) != DV_OK)
    return false;
  c = UploadCache{hsh, img.rows, img.cols, img.channels, true};
  return true;
}
void forget_engine(dv_engine* e) {
  std::lock_guard<std::mutex> lk(upload_mu());
  upload_cache().erase(e);
}

class EstimatorImpl : public Estimator_net::Estimator {
 public:
  explicit EstimatorImpl(std::shared_ptr<dv_engine> e) : e_(std::move(e)) {}
  void sp_extractor(const dv::Image& img) override {        // deep_net.cpp:527-688
    width = img.cols; height = img.rows;
    set_image(img);
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
    set_image(img);
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
  void lg_matcher() override {                               // deep_net.cpp:1003-1133 (legacy self-match)
    std::vector<dv::Pt> k = sp_kpts;
    std::vector<float> d = sp_desc;
    if (k.size() < 10) return;
    lg_matcher(k, k, d, d, height, width, height, width);
  }

 private:
  void set_image(const dv::Image& img) {
#ifdef DV_SHIM_WITH_OPENCV
    image = cv::Mat(img.rows, img.cols, img.channels == 3 ? CV_8UC3 : CV_8UC1, const_cast<uint8_t*>(img.data),
                    (size_t)(img.step ? img.step : img.cols * img.channels));
#else
    image = img;
#endif
  }
  std::shared_ptr<dv_engine> e_;
};

// binary PGM (P5, maxval 255) reader for test_in_dataset without OpenCV
bool read_pgm(const std::string& path, std::vector<uint8_t>* px, int* rows, int* cols) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[3] = {0, 0, 0};
  int w = 0, h = 0, mx = 0;
  bool ok = fscanf(f, "%2s", magic) == 1 && strcmp(magic, "P5") == 0;
  auto next_int = [&](int* v) {
    int c = fgetc(f);
    while (c == '#' || c == ' ' || c == '\n' || c == '\r' || c == '\t') {
      if (c == '#') while (c != '\n' && c != EOF) c = fgetc(f);
      c = fgetc(f);
    }
    if (c == EOF) return false;
    ungetc(c, f);
    return fscanf(f, "%d", v) == 1;
  };
  ok = ok && next_int(&w) && next_int(&h) && next_int(&mx) && mx == 255 && w > 0 && h > 0;
  if (ok) {
    fgetc(f);                                     // the single whitespace after maxval
    px->resize((size_t)w * h);
    ok = fread(px->data(), 1, px->size(), f) == px->size();
  }
  fclose(f);
  *rows = h; *cols = w;
  return ok;
}

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
  void sort_in_faiss(float* xq, float* db, int nb) override {  // deep_net.cpp:1325-1384 (query first, see header)
    float D[3]; int64_t I[3];
    if (dv_bank_import(e_.get(), db, nb) != DV_OK || dv_bank_search(e_.get(), xq, nb, 3, D, I) != DV_OK) return;
    for (int j = 0; j < 3; ++j) {
      top_sim_index.push_back((int)I[j]);
      top_sim.push_back(D[j]);
      sim_map[flag_].push_back((int)I[j]);
    }
    ++flag_;
  }
  void test_in_dataset(const std::string filepath) override {  // deep_net.cpp:1386-1428 (without the imshow loop)
    std::vector<std::string> names;
    if (DIR* d = opendir(filepath.c_str())) {
      while (dirent* ent = readdir(d)) {
        const std::string n = ent->d_name;
        if (n.size() > 4 && n.substr(n.size() - 4) == ".pgm") names.push_back(filepath + "/" + n);
      }
      closedir(d);
    }
    std::sort(names.begin(), names.end());
    for (const std::string& path : names) {
      std::vector<uint8_t> px;
      int rows = 0, cols = 0;
      if (!read_pgm(path, &px, &rows, &cols)) continue;
      mix_des.clear();
      mix_extractor(dv::Image{px.data(), rows, cols, 1, cols});
      if (mix_des.size() != 512) continue;
      const size_t have = descriptors_database.size() / 512;
      if (have > 15) {
        std::vector<float> prefix(descriptors_database.begin(), descriptors_database.begin() + (have - 15) * 512);
        sort_in_faiss(mix_des.data(), prefix.data(), (int)(prefix.size() / 512));
      }
      descriptors_database.insert(descriptors_database.end(), mix_des.begin(), mix_des.end());
    }
  }

 private:
  std::shared_ptr<dv_engine> e_;
  int flag_ = 0;                                              // the reference's function-static call counter
};

}  // namespace

namespace dv {
void shim_forget_engine(dv_engine* e) { forget_engine(e); }
}  // namespace dv

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
