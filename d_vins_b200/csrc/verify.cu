// Loop decision + geometric verification on the GPU: the step right after LightGlue.
//   PoseGraph::detectLoop     loop_fusion/src/pose_graph.cpp:451-509   (threshold rule on the top-k similarities)
//   KeyFrame::PnPRANSAC       loop_fusion/src/keyframe.cpp:805-868     (cv::solvePnPRansac, K = I, extrinsic guess)
//   KeyFrame::findConnection  loop_fusion/src/keyframe.cpp:1094-1183   (inlier count / yaw / translation gates)
// OpenCV's RANSAC is a serial hypothesis stream driven by cv::RNG; here ALL hypotheses of a pair are generated, solved
// and scored in parallel - one CTA per (current, old) keyframe pair, one thread per hypothesis, fp64 throughout.  The
// algorithm (counter-based sampling, 5-point Gauss-Newton from the extrinsic guess, inlier count, lowest-index
// tie-break, refinement on the winner's inliers) is defined by oracle/pnp.py, which is pinned against
// cv2.solvePnPRansac (tests/golden/pnp_cv2.npz).  ~200 x (5 x 4 GN steps + n reprojections) fp64 operations per pair:
// latency-bound, not a roofline kernel - it exists so the keyframe pipe can close on the device without a CPU stage.
#include <math.h>
#include <string.h>

#include <vector>

#include "engine.h"

namespace dv {

namespace {

constexpr int VF_THREADS = 256;
constexpr int VF_MODEL = 5;
constexpr int VF_HYP_ITERS = 4;
constexpr int VF_REFINE_ITERS = 10;

struct VfPair {                // per pair, device
  double R0[9], t0[3];         // world -> camera extrinsic guess
  int n, off;                  // correspondences [off, off + n) of the packed arrays
};
struct VfOut {
  double R[9], t[3];
  int ok, n_inliers;
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ inline void exp_so3(const double* w, double* E) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;
  if (th < 1e-12) { a = 1.0; b = 0.0; }
  else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  // K = [w]x ; K^2 = w w^T - th2 I
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double k2 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
      E[i * 3 + j] = (i == j ? 1.0 : 0.0) + a * K[i * 3 + j] + b * k2;
    }
}

// accumulate one correspondence into the upper triangle of J^T J (21 values) and J^T r (6 values)
__device__ inline bool accum_point(const double* R, const double* t, const double* X, const double* u, double* A,
                                   double* g) {
  const double P0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const double P1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  const double P2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  if (!(P2 > 1e-9)) return false;
  const double iz = 1.0 / P2;
  const double r0 = P0 * iz - u[0], r1 = P1 * iz - u[1];
  const double a = iz, c0 = -P0 * iz * iz, c1 = -P1 * iz * iz;
  // Jp = [[a,0,c0],[0,a,c1]] ; J = [Jp * (-[P]x) | Jp]
  // -[P]x = [[0,P2,-P1],[-P2,0,P0],[P1,-P0,0]]
  const double J0[6] = {c0 * P1, a * P2 - c0 * P0, -a * P1, a, 0.0, c0};
  const double J1[6] = {-a * P2 + c1 * P1, -c1 * P0, a * P0, 0.0, a, c1};
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) A[k++] += J0[i] * J0[j] + J1[i] * J1[j];
    g[i] += J0[i] * r0 + J1[i] * r1;
  }
  return true;
}

// solve (A + 1e-12 I) d = -g by Cholesky; A upper triangle packed row-major (21).  false if not positive definite.
__device__ inline bool solve6(const double* Au, const double* g, double* d) {
  double M[6][6];
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) { M[i][j] = Au[k]; M[j][i] = Au[k]; ++k; }
#pragma unroll
  for (int i = 0; i < 6; ++i) M[i][i] += 1e-12;
  double L[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) L[i][j] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = M[i][j];
#pragma unroll
      for (int q = 0; q < j; ++q) s -= L[i][q] * L[j][q];
      if (i == j) {
        if (!(s > 1e-300)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = -g[i];
#pragma unroll
    for (int q = 0; q < i; ++q) s -= L[i][q] * y[q];
    y[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int q = i + 1; q < 6; ++q) s -= L[q][i] * d[q];
    d[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (!isfinite(d[i])) return false;
  return true;
}

__device__ inline void apply_update(double* R, double* t, const double* d) {
  double E[9];
  exp_so3(d, E);
  double Rn[9], tn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = E[i * 3] * R[j] + E[i * 3 + 1] * R[3 + j] + E[i * 3 + 2] * R[6 + j];
    tn[i] = E[i * 3] * t[0] + E[i * 3 + 1] * t[1] + E[i * 3 + 2] * t[2] + d[3 + i];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = tn[i];
}

__device__ inline bool is_inlier(const double* R, const double* t, const double* X, const double* u, double thr2) {
  const double P0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const double P1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  const double P2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  if (!(P2 > 1e-9)) return false;
  const double e0 = P0 / P2 - u[0], e1 = P1 / P2 - u[1];
  return e0 * e0 + e1 * e1 <= thr2;
}

// One CTA per pair.  Dynamic shared memory: X [n,3], u [n,2] (fp64).
__global__ void __launch_bounds__(VF_THREADS) k_verify_loop(const VfPair* __restrict__ pairs,
                                                            const double* __restrict__ X_all,
                                                            const double* __restrict__ u_all, int iters, double thr2,
                                                            unsigned long long seed, uint8_t* __restrict__ status,
                                                            VfOut* __restrict__ out) {
  extern __shared__ double vsm[];
  __shared__ int s_cnt[VF_THREADS];
  __shared__ double s_pose[12], s_pose0[12];
  __shared__ int s_best, s_ok;
  __shared__ double s_red[8][27];
  const VfPair pr = pairs[blockIdx.x];
  const int n = pr.n, tid = threadIdx.x;
  double* sX = vsm;
  double* sU = vsm + (size_t)n * 3;
  for (int i = tid; i < n * 3; i += VF_THREADS) sX[i] = X_all[(size_t)pr.off * 3 + i];
  for (int i = tid; i < n * 2; i += VF_THREADS) sU[i] = u_all[(size_t)pr.off * 2 + i];
  __syncthreads();
  // ---- hypotheses: thread h (+ k * 256) samples 5 correspondences, runs Gauss-Newton from the guess, counts inliers
  int my_cnt = -1, my_h = 0x7fffffff;
  double bR[9], bt[3];
  for (int h = tid; h < iters && n >= VF_MODEL; h += VF_THREADS) {
    int idx[VF_MODEL];
#pragma unroll
    for (int j = 0; j < VF_MODEL; ++j) {
      int id = 0;
      for (int tr = 0; tr < 16; ++tr) {
        id = (int)(splitmix64(seed + ((unsigned long long)h << 24) + ((unsigned long long)j << 16) + (unsigned long long)tr) %
                   (unsigned long long)n);
        bool dup = false;
        for (int q = 0; q < j; ++q) dup |= idx[q] == id;
        if (!dup) break;
      }
      idx[j] = id;
    }
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = pr.R0[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = pr.t0[i];
    bool ok = true;
    for (int it = 0; it < VF_HYP_ITERS && ok; ++it) {
      double A[21], g[6], d[6];
#pragma unroll
      for (int i = 0; i < 21; ++i) A[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) g[i] = 0.0;
#pragma unroll
      for (int j = 0; j < VF_MODEL; ++j) ok = ok && accum_point(R, t, sX + idx[j] * 3, sU + idx[j] * 2, A, g);
      ok = ok && solve6(A, g, d);
      if (ok) apply_update(R, t, d);
    }
    if (!ok) continue;
    int c = 0;
    for (int i = 0; i < n; ++i) c += is_inlier(R, t, sX + i * 3, sU + i * 2, thr2) ? 1 : 0;
    if (c > my_cnt) {          // h ascending per thread: the first maximum is kept
      my_cnt = c; my_h = h;
#pragma unroll
      for (int i = 0; i < 9; ++i) bR[i] = R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) bt[i] = t[i];
    }
  }
  // ---- best hypothesis: most inliers, ties -> lowest hypothesis index
  s_cnt[tid] = my_cnt;
  __syncthreads();
  if (tid == 0) { s_best = -1; s_ok = 0; }
  __syncthreads();
  {
    // pack (count, ~h) so that a max picks the highest count and, among equals, the lowest h
    long long key = my_cnt < 0 ? -1ll : (((long long)my_cnt << 32) | (long long)(0x7fffffff - my_h));
    __shared__ long long s_key[VF_THREADS];
    s_key[tid] = key;
    __syncthreads();
    for (int s = VF_THREADS / 2; s > 0; s >>= 1) {
      if (tid < s && s_key[tid + s] > s_key[tid]) s_key[tid] = s_key[tid + s];
      __syncthreads();
    }
    const long long best = s_key[0];
    if (best >= 0 && key == best && (best >> 32) >= VF_MODEL) {
#pragma unroll
      for (int i = 0; i < 9; ++i) s_pose[i] = bR[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) s_pose[9 + i] = bt[i];
#pragma unroll
      for (int i = 0; i < 12; ++i) s_pose0[i] = s_pose[i];
      s_best = (int)(best >> 32);
      s_ok = 1;
    }
    __syncthreads();
  }
  uint8_t* st = status + pr.off;
  if (!s_ok) {                 // OpenCV: returns false, pose stays at the guess, inlier list empty
    for (int i = tid; i < n; i += VF_THREADS) st[i] = 0;
    if (tid == 0) {
      VfOut o;
      for (int i = 0; i < 9; ++i) o.R[i] = pr.R0[i];
      for (int i = 0; i < 3; ++i) o.t[i] = pr.t0[i];
      o.ok = 0; o.n_inliers = 0;
      out[blockIdx.x] = o;
    }
    return;
  }
  // ---- the winner's inlier mask (what the reference copies into `status`)
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = s_pose[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = s_pose[9 + i];
  for (int i = tid; i < n; i += VF_THREADS) st[i] = is_inlier(R, t, sX + i * 3, sU + i * 2, thr2) ? 1 : 0;
  __syncthreads();
  // ---- refinement on the inliers: every thread accumulates its points, block tree-reduction of the 27 sums (fixed
  // order -> deterministic), thread 0 solves and broadcasts the update
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int it = 0; it < VF_REFINE_ITERS; ++it) {
    double A[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) A[i] = 0.0;
    bool ok = true;
    for (int i = tid; i < n; i += VF_THREADS)
      if (st[i]) ok = accum_point(R, t, sX + i * 3, sU + i * 2, A, A + 21) && ok;
    if (!ok) s_fail = 1;
    // warp reduce, then across the 8 warps
#pragma unroll
    for (int i = 0; i < 27; ++i) {
      double v = A[i];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      A[i] = v;
    }
    if ((tid & 31) == 0)
      for (int i = 0; i < 27; ++i) s_red[tid >> 5][i] = A[i];
    __syncthreads();
    if (tid == 0) {
      double S[27], d[6];
      for (int i = 0; i < 27; ++i) { double v = 0.0; for (int w = 0; w < 8; ++w) v += s_red[w][i]; S[i] = v; }
      double Rn[9], tn[3];
      for (int i = 0; i < 9; ++i) Rn[i] = s_pose[i];
      for (int i = 0; i < 3; ++i) tn[i] = s_pose[9 + i];
      if (!s_fail && solve6(S, S + 21, d)) {
        apply_update(Rn, tn, d);
        for (int i = 0; i < 9; ++i) s_pose[i] = Rn[i];
        for (int i = 0; i < 3; ++i) s_pose[9 + i] = tn[i];
      } else {
        s_fail = 1;
      }
    }
    __syncthreads();
    if (s_fail) break;
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = s_pose[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = s_pose[9 + i];
    __syncthreads();
  }
  if (tid == 0) {
    VfOut o;
    // a failed refinement keeps the winning hypothesis' own pose (oracle/pnp.py: `if ok: R, t = Rr, tr`)
    const double* src = s_fail ? s_pose0 : s_pose;
    for (int i = 0; i < 9; ++i) o.R[i] = src[i];
    for (int i = 0; i < 3; ++i) o.t[i] = src[9 + i];
    o.ok = 1; o.n_inliers = s_best;
    out[blockIdx.x] = o;
  }
}

// utility.h:75-91 (degrees) / :140-148
inline double r2yaw_deg(const double* R) { return atan2(R[3], R[0]) / M_PI * 180.0; }
inline double normalize_angle(double a) {
  if (a > 0) return a - 360.0 * floor((a + 180.0) / 360.0);
  return a + 360.0 * floor((-a + 180.0) / 360.0);
}
inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
inline void mat3_t(const double* A, double* T) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}
inline void mat3_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
// Eigen::Quaterniond(R) -> (w, x, y, z)
inline void rot_to_quat(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[7] - R[5]) * s; q[2] = (R[2] - R[6]) * s; q[3] = (R[3] - R[1]) * s;
    return;
  }
  int i = 0;
  if (R[4] > R[0]) i = 1;
  if (R[8] > R[i * 3 + i]) i = 2;
  const int j = (i + 1) % 3, k = (i + 2) % 3;
  double s = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
  q[1 + i] = 0.5 * s;
  s = 0.5 / s;
  q[0] = (R[k * 3 + j] - R[j * 3 + k]) * s;
  q[1 + j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
  q[1 + k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
}

struct DevTmp {
  void* p = nullptr;
  ~DevTmp() { if (p) cudaFree(p); }
};

}  // namespace

}  // namespace dv

using namespace dv;

extern "C" {

void dv_loop_params_default(dv_loop_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->struct_size = (int32_t)sizeof(dv_loop_params);
  p->min_loop_num = 18;            // euroc_stereo_imu_config.yaml:55-61
  p->ransac_iters = 200;           // keyframe.cpp:835
  p->pnp_inflation = 3.5;
  p->max_theta_diff = 40.0;
  p->max_pose_diff = 25.0;
  p->loop_top_thres = 0.45;
  p->loop_back_thres = 0.40;
  p->min_frame_index = 50;         // pose_graph.cpp:470
  p->qic[0] = p->qic[4] = p->qic[8] = 1.0;
  p->seed = 0;
}

int64_t dv_detect_loop(const dv_loop_params* p, const float* top_sim, const int64_t* top_sim_index, int32_t k,
                       int64_t frame_index) {
  if (!p || !top_sim || !top_sim_index || k < 1) return -1;
  // pose_graph.cpp:451-509.  faiss pads missing results with index -1 / similarity -inf: those never qualify.
  bool find_loop = false;
  if (top_sim_index[0] >= 0 && (double)top_sim[0] > p->loop_top_thres)
    for (int i = 1; i < k; ++i)
      if (top_sim_index[i] >= 0 && (double)top_sim[i] > p->loop_back_thres) find_loop = true;
  if (!find_loop || frame_index <= p->min_frame_index) return -1;
  int64_t min_index = -1;
  for (int i = 0; i < k; ++i) {
    if (top_sim_index[i] < 0) continue;
    if (min_index == -1 || (top_sim_index[i] < min_index && (double)top_sim[i] > p->loop_back_thres)) min_index = top_sim_index[i];
  }
  return min_index;
}

dv_status dv_verify_loop(dv_engine* h, int32_t b, const int32_t* n_pts, int32_t cap, const double* pts3d,
                         const double* pts2d_norm, const double* vio_R, const double* vio_T,
                         const dv_loop_params* p, uint8_t* status, dv_loop_result* out) {
  if (!h) { set_error("null engine"); return DV_ERR_INVALID; }
  Engine* e = reinterpret_cast<Engine*>(h);
  if (b < 1 || !n_pts || cap < 1 || !pts3d || !pts2d_norm || !vio_R || !vio_T || !p || !status || !out ||
      p->struct_size != (int32_t)sizeof(dv_loop_params) || p->ransac_iters < 1 || p->ransac_iters > 65536) {
    set_error("dv_verify_loop: bad arguments");
    return DV_ERR_INVALID;
  }
  std::vector<VfPair> pairs;
  std::vector<int> which;
  std::vector<double> X, U;
  int max_n = 0;
  for (int i = 0; i < b; ++i) {
    dv_loop_result& o = out[i];
    memset(&o, 0, sizeof(o));
    o.pnp_r_old[0] = o.pnp_r_old[4] = o.pnp_r_old[8] = 1.0;
    o.relative_q[0] = 1.0;
    memset(status + (size_t)i * cap, 0, (size_t)cap);
    const int n = n_pts[i];
    if (n < 0 || n > cap) { set_error("dv_verify_loop: n_pts out of range"); return DV_ERR_INVALID; }
    if (n <= p->min_loop_num) continue;                        // keyframe.cpp:1094: not enough matches to try
    VfPair pr;
    // R_w_c = vio_R * qic ; T_w_c = vio_T + vio_R * tic ; guess = inverse (keyframe.cpp:817-822)
    double Rwc[9], Twc[3], tmp[3];
    mat3_mul(vio_R + (size_t)i * 9, p->qic, Rwc);
    mat3_vec(vio_R + (size_t)i * 9, p->tic, tmp);
    for (int k = 0; k < 3; ++k) Twc[k] = vio_T[(size_t)i * 3 + k] + tmp[k];
    mat3_t(Rwc, pr.R0);
    mat3_vec(pr.R0, Twc, tmp);
    for (int k = 0; k < 3; ++k) pr.t0[k] = -tmp[k];
    pr.n = n; pr.off = (int)(X.size() / 3);
    X.insert(X.end(), pts3d + (size_t)i * cap * 3, pts3d + (size_t)i * cap * 3 + (size_t)n * 3);
    U.insert(U.end(), pts2d_norm + (size_t)i * cap * 2, pts2d_norm + (size_t)i * cap * 2 + (size_t)n * 2);
    pairs.push_back(pr);
    which.push_back(i);
    max_n = std::max(max_n, n);
  }
  if (pairs.empty()) return DV_OK;
  const size_t smem = (size_t)max_n * 5 * sizeof(double);
  if (smem > 200 * 1024) { set_error("dv_verify_loop: too many correspondences per pair"); return DV_ERR_CAPACITY; }
  const int P = (int)pairs.size();
  const size_t npt = X.size() / 3;
  DevTmp dP, dX, dU, dS, dO;
  DV_CUDA_OK(cudaMalloc(&dP.p, sizeof(VfPair) * P));
  DV_CUDA_OK(cudaMalloc(&dX.p, sizeof(double) * 3 * npt));
  DV_CUDA_OK(cudaMalloc(&dU.p, sizeof(double) * 2 * npt));
  DV_CUDA_OK(cudaMalloc(&dS.p, npt));
  DV_CUDA_OK(cudaMalloc(&dO.p, sizeof(VfOut) * P));
  DV_CUDA_OK(cudaMemcpyAsync(dP.p, pairs.data(), sizeof(VfPair) * P, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dX.p, X.data(), sizeof(double) * 3 * npt, cudaMemcpyHostToDevice, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(dU.p, U.data(), sizeof(double) * 2 * npt, cudaMemcpyHostToDevice, e->st));
  if (smem > 48 * 1024)
    DV_CUDA_OK(cudaFuncSetAttribute(k_verify_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double thr = p->pnp_inflation / 460.0;                 // keyframe.cpp:835
  k_verify_loop<<<P, VF_THREADS, smem, e->st>>>(reinterpret_cast<const VfPair*>(dP.p), reinterpret_cast<const double*>(dX.p),
                                                reinterpret_cast<const double*>(dU.p), p->ransac_iters, thr * thr,
                                                (unsigned long long)p->seed, reinterpret_cast<uint8_t*>(dS.p),
                                                reinterpret_cast<VfOut*>(dO.p));
  DV_CUDA_OK(cudaGetLastError());
  DV_LAUNCHED(e, 1);
  std::vector<VfOut> ho(P);
  std::vector<uint8_t> hs(npt);
  DV_CUDA_OK(cudaMemcpyAsync(ho.data(), dO.p, sizeof(VfOut) * P, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaMemcpyAsync(hs.data(), dS.p, npt, cudaMemcpyDeviceToHost, e->st));
  DV_CUDA_OK(cudaStreamSynchronize(e->st));
  for (int j = 0; j < P; ++j) {
    const int i = which[j];
    const VfPair& pr = pairs[j];
    dv_loop_result& o = out[i];
    memcpy(status + (size_t)i * cap, hs.data() + pr.off, (size_t)pr.n);
    o.n_inliers = ho[j].n_inliers;
    // keyframe.cpp:858-866: camera pose of the old keyframe -> body pose with the extrinsics
    double Rwco[9], Twco[3], nt[3], qicT[9], tmp[3];
    mat3_t(ho[j].R, Rwco);
    for (int k = 0; k < 3; ++k) nt[k] = -ho[j].t[k];
    mat3_vec(Rwco, nt, Twco);
    mat3_t(p->qic, qicT);
    mat3_mul(Rwco, qicT, o.pnp_r_old);
    mat3_vec(o.pnp_r_old, p->tic, tmp);
    for (int k = 0; k < 3; ++k) o.pnp_t_old[k] = Twco[k] - tmp[k];
    if (o.n_inliers > p->min_loop_num) {                       // keyframe.cpp:1163-1183
      double PRt[9], dT[3], rq[9];
      mat3_t(o.pnp_r_old, PRt);
      for (int k = 0; k < 3; ++k) dT[k] = vio_T[(size_t)i * 3 + k] - o.pnp_t_old[k];
      mat3_vec(PRt, dT, o.relative_t);
      mat3_mul(PRt, vio_R + (size_t)i * 9, rq);
      rot_to_quat(rq, o.relative_q);
      o.relative_yaw = normalize_angle(r2yaw_deg(vio_R + (size_t)i * 9) - r2yaw_deg(o.pnp_r_old));
      const double tn = sqrt(o.relative_t[0] * o.relative_t[0] + o.relative_t[1] * o.relative_t[1] + o.relative_t[2] * o.relative_t[2]);
      o.has_loop = (fabs(o.relative_yaw) < p->max_theta_diff && tn < p->max_pose_diff) ? 1 : 0;
    }
  }
  return DV_OK;
}

}  // extern "C"
