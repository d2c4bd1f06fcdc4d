// TEST INFRASTRUCTURE (oracle side) - thin C wrapper around the REFERENCE's own CUDA pre/post-processing kernels.
//
// The reference sources are compiled where they lie (never copied):
//   /root/reference/loop_fusion/src/deep_net/tensorrt_tools/preprocess_kernel.cu  (+ cuda_tools.cpp, ilogger.cpp)
// by oracle/ref_pre/build_ref.py into oracle/_ref/libdvins_refpre.so.  Every entry point below calls the reference's
// public CUDAKernel:: function exactly as its call site in loop_fusion/src/deep_net/deep_net.cpp does; this file only
// moves host buffers to the device and back.  Used by tests/test_refpre_gpu.py to pin (a) the oracle's restatement of
// these kernels and (b) the engine's own pre-processing kernels, bit for bit.  Never linked into the product.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "preprocess_kernel.cuh"

#define RP_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace {
template <class T> struct DevBuf {
  T* p = nullptr;
  explicit DevBuf(size_t n) { if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) p = nullptr; }
  ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

extern "C" {

// deep_net.cpp:527-585 (sp_extractor): Norm::alpha_beta(1/255.f, 0, Invert), const_value 114, dst [1,h_adj,w_adj].
int refpre_sp(const uint8_t* img, int channels, int rows, int cols, int h_adj, int w_adj, const float* d2i, float* out) {
  const size_t n_img = (size_t)rows * cols * channels, n_out = (size_t)h_adj * w_adj;
  DevBuf<uint8_t> d_img(n_img); DevBuf<float> d_m(8), d_out(n_out);
  if (!d_img.p || !d_m.p || !d_out.p) return -1;
  RP_CHECK(cudaMemcpy(d_img.p, img, n_img, cudaMemcpyHostToDevice));
  RP_CHECK(cudaMemcpy(d_m.p, d2i, 6 * sizeof(float), cudaMemcpyHostToDevice));
  auto norm = CUDAKernel::Norm::alpha_beta(1 / 255.f, 0.f, CUDAKernel::ChannelType::Invert);
  CUDAKernel::warp_affine_bilinear_and_normalize_plane(channels, d_img.p, cols * channels, cols, rows, d_out.p, w_adj, h_adj,
                                                       d_m.p, 114, norm, 0);
  RP_CHECK(cudaDeviceSynchronize());
  RP_CHECK(cudaMemcpy(out, d_out.p, n_out * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// deep_net.cpp:1254-1307 (mix_extractor): 3-channel image (gray frames are cvtColor'ed to BGR by the caller), mean/std
// literals in the reference's order, Norm::mean_std(..., 1/255.f, Invert), dst [3,320,320].
int refpre_mix(const uint8_t* img, int channels, int rows, int cols, const float* d2i, float* out) {
  const size_t n_img = (size_t)rows * cols * channels, n_out = (size_t)3 * 320 * 320;
  DevBuf<uint8_t> d_img(n_img); DevBuf<float> d_m(8), d_out(n_out);
  if (!d_img.p || !d_m.p || !d_out.p) return -1;
  RP_CHECK(cudaMemcpy(d_img.p, img, n_img, cudaMemcpyHostToDevice));
  RP_CHECK(cudaMemcpy(d_m.p, d2i, 6 * sizeof(float), cudaMemcpyHostToDevice));
  float mean[] = {0.406, 0.456, 0.485};
  float std[] = {0.225, 0.224, 0.229};
  auto norm = CUDAKernel::Norm::mean_std(mean, std, 1 / 255.f, CUDAKernel::ChannelType::Invert);
  CUDAKernel::warp_affine_bilinear_and_normalize_plane_mix(channels, d_img.p, cols * channels, cols, rows, d_out.p, 320, 320,
                                                           d_m.p, 114, norm, 0);
  RP_CHECK(cudaDeviceSynchronize());
  RP_CHECK(cudaMemcpy(out, d_out.p, n_out * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// deep_net.cpp:633-659 (int keypoints out of SuperPoint) and :874-880 (float keypoints into LightGlue); n = numel.
int refpre_normalize_kpts_i32(const int* src, int n, float shift_w, float shift_h, float scale, float* dst) {
  DevBuf<int> d_s(n); DevBuf<float> d_d(n);
  if (!d_s.p || !d_d.p) return -1;
  RP_CHECK(cudaMemcpy(d_s.p, src, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDAKernel::normalize_kpts(d_s.p, d_d.p, n, shift_w, shift_h, scale, 0);
  RP_CHECK(cudaDeviceSynchronize());
  RP_CHECK(cudaMemcpy(dst, d_d.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}
int refpre_normalize_kpts_f32(const float* src, int n, float shift_w, float shift_h, float scale, float* dst) {
  DevBuf<float> d_s(n), d_d(n);
  if (!d_s.p || !d_d.p) return -1;
  RP_CHECK(cudaMemcpy(d_s.p, src, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDAKernel::normalize_kpts(d_s.p, d_d.p, n, shift_w, shift_h, scale, 0);
  RP_CHECK(cudaDeviceSynchronize());
  RP_CHECK(cudaMemcpy(dst, d_d.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// deep_net.cpp:930-960 (lg_matcher post): matches [K,2] -> matched keypoints in pixels.  kn0 / kn1 are the NORMALISED
// keypoints the matcher consumed; kpts_num = 2 * M as at the call site.  The reference launches its second kernel over
// kpts_num threads although mkpts hold 2 * K floats (SURVEY §2.2 defect), so the device buffers here are padded to
// max(2K, 2M) floats; only the first 2K are returned.
int refpre_matches_post(const float* kn0, int m, const float* kn1, int n, const int* matches, int k, float shift_w,
                        float shift_h, float sw0, float sh0, float sw1, float sh1, float* mk0, float* mk1) {
  const size_t cap = (size_t)2 * std::max(std::max(m, n), k) + 2;
  DevBuf<float> d_k0((size_t)2 * m), d_k1((size_t)2 * n), d_m0(cap), d_m1(cap); DevBuf<int> d_ma((size_t)2 * k);
  if (!d_k0.p || !d_k1.p || !d_m0.p || !d_m1.p || !d_ma.p) return -1;
  RP_CHECK(cudaMemcpy(d_k0.p, kn0, (size_t)2 * m * 4, cudaMemcpyHostToDevice));
  RP_CHECK(cudaMemcpy(d_k1.p, kn1, (size_t)2 * n * 4, cudaMemcpyHostToDevice));
  RP_CHECK(cudaMemcpy(d_ma.p, matches, (size_t)2 * k * 4, cudaMemcpyHostToDevice));
  RP_CHECK(cudaMemset(d_m0.p, 0, cap * 4));
  RP_CHECK(cudaMemset(d_m1.p, 0, cap * 4));
  CUDAKernel::matches_post_process(d_k0.p, d_k1.p, d_ma.p, d_m0.p, d_m1.p, shift_w, shift_h, sw0, sh0, sw1, sh1, 2 * m, 2 * k, 0);
  RP_CHECK(cudaDeviceSynchronize());
  RP_CHECK(cudaMemcpy(mk0, d_m0.p, (size_t)2 * k * 4, cudaMemcpyDeviceToHost));
  RP_CHECK(cudaMemcpy(mk1, d_m1.p, (size_t)2 * k * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int refpre_version() { return 1; }

}  // extern "C"
