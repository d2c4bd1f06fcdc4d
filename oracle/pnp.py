"""CPU restatement (numpy float64) of the loop decision + geometric verification step that follows LightGlue
(TEST INFRASTRUCTURE).

Reference code this follows:
  * loop_fusion/src/pose_graph.cpp:451-509   PoseGraph::detectLoop - thresholds on the top-3 cosine similarities,
                                              returns the SMALLEST qualifying keyframe index
  * loop_fusion/src/keyframe.cpp:805-868     KeyFrame::PnPRANSAC - cv::solvePnPRansac(points3d, points2d_norm, K = I,
                                              no distortion, useExtrinsicGuess = true (current VIO pose), 200 iterations,
                                              reprojection threshold PNP_INFLATION / 460, confidence 0.99) and the
                                              camera -> body conversion with the extrinsics (qic, tic)
  * loop_fusion/src/keyframe.cpp:1094-1183   findConnection tail - inlier count > MIN_LOOP_NUM, relative pose, yaw /
                                              translation gates (MAX_THETA_DIFF, MAX_POSE_DIFF)
  * loop_fusion/src/utility/utility.h:75-91, :140-148  R2ypr (degrees), normalizeAngle

cv::solvePnPRansac itself lives in OpenCV (3.4.10, README.md:22 - un-vendored): a 5-point minimal EPnP model inside a
sequential RANSAC driven by cv::RNG, adaptive stopping, then an iterative refinement on the inliers.  Its hypothesis
stream is inherently serial, so the B200 engine evaluates ALL `ransac_iters` hypotheses in parallel and this oracle
defines the algorithm both sides implement:

  hypothesis h:  5 distinct correspondences chosen by a counter-based generator (splitmix64 of (seed, h, j, try)),
                 4 Gauss-Newton iterations on the 6-DoF pose starting from the extrinsic guess (the reference passes
                 useExtrinsicGuess = true, so the guess is the natural linearisation point), left-multiplicative update
  score:         inliers = points in front of the camera with squared reprojection error <= thresh^2
  selection:     most inliers, ties -> lowest hypothesis index; fewer than 5 inliers -> failure (OpenCV returns false and
                 leaves rvec / tvec at the guess with an empty inlier list)
  refinement:    10 Gauss-Newton iterations on the winner's inliers (OpenCV: solvePnP(ITERATIVE) on the inliers); the
                 reported inlier mask is the winning hypothesis' mask, like OpenCV's

Pinned against OpenCV: tests/golden/make_golden_pnp.py runs cv2.solvePnPRansac with the reference's arguments on seeded
synthetic scenes (true inliers + gross outliers) and stores its inlier masks and poses; tests/test_oracle_cpu.py requires
this oracle to find the same inlier sets and the same pose within 1e-6.
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1
MODEL_POINTS = 5          # OpenCV solvePnPRansac: model_points = 5 (EPnP kernel) unless P3P is requested
HYP_GN_ITERS = 4
REFINE_GN_ITERS = 10


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def sample_indices(seed: int, h: int, n: int):
    """5 distinct indices in [0, n) for hypothesis h: j-th index = first non-duplicate of
    splitmix64(seed + (h << 24) + (j << 16) + try) % n, try = 0..15 (then the duplicate is kept - only reachable for
    tiny n, and such a hypothesis simply scores badly)."""
    out = []
    for j in range(MODEL_POINTS):
        idx = 0
        for t in range(16):
            idx = splitmix64((seed + (h << 24) + (j << 16) + t) & MASK64) % n
            if idx not in out:
                break
        out.append(int(idx))
    return out


def _exp_so3(w):
    th = np.sqrt(w @ w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + (np.sin(th) / th) * K + ((1 - np.cos(th)) / (th * th)) * (K @ K)


def _cholesky_solve6(A, b):
    """In-place Cholesky (lower) of the 6x6 normal matrix; None if not positive definite."""
    L = np.zeros((6, 6))
    for i in range(6):
        for j in range(i + 1):
            s = A[i, j] - L[i, :j] @ L[j, :j]
            if i == j:
                if s <= 1e-300:
                    return None
                L[i, i] = np.sqrt(s)
            else:
                L[i, j] = s / L[j, j]
    y = np.zeros(6)
    for i in range(6):
        y[i] = (b[i] - L[i, :i] @ y[:i]) / L[i, i]
    x = np.zeros(6)
    for i in range(5, -1, -1):
        x[i] = (y[i] - L[i + 1:, i] @ x[i + 1:]) / L[i, i]
    return x


def gauss_newton(R, t, X, u, iters):
    """Minimise sum |proj(R X + t) - u|^2 with left-multiplicative updates.  Returns (R, t, ok)."""
    R = R.copy(); t = t.copy()
    for _ in range(iters):
        A = np.zeros((6, 6)); g = np.zeros(6)
        for Xi, ui in zip(X, u):
            P = R @ Xi + t
            if P[2] <= 1e-9:
                return R, t, False
            iz = 1.0 / P[2]
            r = np.array([P[0] * iz - ui[0], P[1] * iz - ui[1]])
            Jp = np.array([[iz, 0.0, -P[0] * iz * iz], [0.0, iz, -P[1] * iz * iz]])
            # d(P)/d(omega) = -[P]x ; d(P)/d(upsilon) = I
            Px = np.array([[0, -P[2], P[1]], [P[2], 0, -P[0]], [-P[1], P[0], 0]])
            J = np.concatenate([Jp @ (-Px), Jp], axis=1)        # 2 x 6
            A += J.T @ J
            g += J.T @ r
        A += 1e-12 * np.eye(6)
        d = _cholesky_solve6(A, -g)
        if d is None or not np.all(np.isfinite(d)):
            return R, t, False
        E = _exp_so3(d[:3])
        R = E @ R
        t = E @ t + d[3:]
    return R, t, True


def inlier_mask(R, t, X, u, thresh):
    P = X @ R.T + t
    z = P[:, 2]
    ok = z > 1e-9
    zs = np.where(ok, z, 1.0)
    e = (P[:, 0] / zs - u[:, 0]) ** 2 + (P[:, 1] / zs - u[:, 1]) ** 2
    return ok & (e <= thresh * thresh)


def pnp_ransac(X, u, R0, t0, thresh, iters=200, seed=0):
    """-> (ok, R, t, mask).  X [n,3] world points, u [n,2] normalised image points of the OLD camera, (R0, t0) the
    world->camera extrinsic guess."""
    X = np.asarray(X, np.float64); u = np.asarray(u, np.float64)
    n = X.shape[0]
    best_cnt, best = -1, None
    if n >= MODEL_POINTS:
        for h in range(iters):
            idx = sample_indices(seed, h, n)
            R, t, ok = gauss_newton(R0, t0, X[idx], u[idx], HYP_GN_ITERS)
            if not ok:
                continue
            m = inlier_mask(R, t, X, u, thresh)
            c = int(m.sum())
            if c > best_cnt:
                best_cnt, best = c, (R, t, m)
    if best is None or best_cnt < MODEL_POINTS:
        return False, R0.copy(), t0.copy(), np.zeros(n, bool)
    R, t, m = best
    Rr, tr, ok = gauss_newton(R, t, X[m], u[m], REFINE_GN_ITERS)
    if ok:
        R, t = Rr, tr
    return True, R, t, m


def r2ypr(R):
    """utility.h:75-91, degrees."""
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = np.arctan2(n[1], n[0])
    p = np.arctan2(-n[2], n[0] * np.cos(y) + n[1] * np.sin(y))
    r = np.arctan2(a[0] * np.sin(y) - a[1] * np.cos(y), -o[0] * np.sin(y) + o[1] * np.cos(y))
    return np.array([y, p, r]) / np.pi * 180.0


def normalize_angle(a):
    """utility.h:140-148"""
    if a > 0:
        return a - 360.0 * np.floor((a + 180.0) / 360.0)
    return a + 360.0 * np.floor((-a + 180.0) / 360.0)


def rot_to_quat(R):
    """Eigen::Quaterniond(R) (w, x, y, z), Shepperd's branches as Eigen implements them."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([w, (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s])
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[1 + i] = 0.5 * s
    s = 0.5 / s
    q[0] = (R[k, j] - R[j, k]) * s
    q[1 + j] = (R[j, i] + R[i, j]) * s
    q[1 + k] = (R[k, i] + R[i, k]) * s
    return q


def verify_loop(X, u, vio_R, vio_T, qic, tic, *, pnp_inflation=3.5, min_loop_num=18, max_theta_diff=40.0,
                max_pose_diff=25.0, iters=200, seed=0):
    """keyframe.cpp:1094-1183 + :805-868 for one (current, old) keyframe pair.
    X [n,3]: the current keyframe's matched 3-D points (world), u [n,2]: the matched OLD keypoints, normalised.
    Returns dict(has_loop, n_inliers, status [n] u8, pnp_T_old, pnp_R_old, relative_t, relative_q (w,x,y,z),
    relative_yaw)."""
    X = np.asarray(X, np.float64).reshape(-1, 3); u = np.asarray(u, np.float64).reshape(-1, 2)
    n = X.shape[0]
    vio_R = np.asarray(vio_R, np.float64).reshape(3, 3); vio_T = np.asarray(vio_T, np.float64).reshape(3)
    qic = np.asarray(qic, np.float64).reshape(3, 3); tic = np.asarray(tic, np.float64).reshape(3)
    out = dict(has_loop=False, n_inliers=0, status=np.zeros(n, np.uint8), pnp_T_old=np.zeros(3), pnp_R_old=np.eye(3),
               relative_t=np.zeros(3), relative_q=np.array([1.0, 0, 0, 0]), relative_yaw=0.0)
    if n <= min_loop_num:                                   # keyframe.cpp:1094
        return out
    R_w_c = vio_R @ qic
    T_w_c = vio_T + vio_R @ tic
    R0 = R_w_c.T                                            # a rotation: inverse == transpose (keyframe.cpp:820)
    t0 = -(R0 @ T_w_c)
    ok, R, t, mask = pnp_ransac(X, u, R0, t0, pnp_inflation / 460.0, iters, seed)
    out["status"] = mask.astype(np.uint8)
    out["n_inliers"] = int(mask.sum())
    R_w_c_old = R.T
    T_w_c_old = R_w_c_old @ (-t)
    PR = R_w_c_old @ qic.T
    PT = T_w_c_old - PR @ tic
    out["pnp_R_old"], out["pnp_T_old"] = PR, PT
    if out["n_inliers"] > min_loop_num:                     # keyframe.cpp:1163
        rt = PR.T @ (vio_T - PT)
        rq = PR.T @ vio_R
        yaw = normalize_angle(r2ypr(vio_R)[0] - r2ypr(PR)[0])
        out["relative_t"], out["relative_q"], out["relative_yaw"] = rt, rot_to_quat(rq), float(yaw)
        out["has_loop"] = bool(abs(yaw) < max_theta_diff and np.linalg.norm(rt) < max_pose_diff)
    return out


def detect_loop(top_sim, top_sim_index, frame_index, loop_top_thres=0.45, loop_back_thres=0.40, min_frame_index=50):
    """pose_graph.cpp:451-509.  top_sim / top_sim_index: the k kNN results of the keyframe (descending similarity).
    Returns the loop candidate's keyframe index or -1."""
    find_loop = False
    if len(top_sim_index) and top_sim[0] > loop_top_thres:
        for i in range(1, len(top_sim_index)):
            if top_sim[i] > loop_back_thres:
                find_loop = True
    if find_loop and frame_index > min_frame_index:
        min_index = -1
        for i in range(len(top_sim_index)):
            if min_index == -1 or (top_sim_index[i] < min_index and top_sim[i] > loop_back_thres):
                min_index = int(top_sim_index[i])
        return min_index
    return -1


# ------------------------------------------------------------------------------------------------ synthetic scenes
def synth_scene(n, n_out, seed, noise=0.0008, drift=0.15):
    """A loop-closure geometry: n world points seen from an OLD camera; the CURRENT keyframe's VIO pose is the old pose
    perturbed by `drift` (the extrinsic guess).  n_out correspondences are gross outliers.
    Returns dict(X, u, vio_R, vio_T, qic, tic, R_true, t_true, inliers)."""
    rng = np.random.default_rng(seed)

    def rot(v):
        return _exp_so3(np.asarray(v, np.float64))
    qic = rot([0.01, -0.02, 0.015]) @ np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]])     # body x-forward -> camera z-forward
    tic = np.array([0.05, -0.02, 0.01])
    R_wb_old = rot(rng.normal(0, 0.3, 3)); T_wb_old = rng.normal(0, 2.0, 3)
    R_wc = R_wb_old @ qic; T_wc = T_wb_old + R_wb_old @ tic
    # points in front of the old camera
    Pc = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2.0, 12.0, n)], 1)
    X = Pc @ R_wc.T + T_wc
    u = Pc[:, :2] / Pc[:, 2:3] + rng.normal(0, noise, (n, 2))
    out_idx = rng.choice(n, n_out, replace=False)
    u[out_idx] = rng.uniform(-0.8, 0.8, (n_out, 2))
    inl = np.ones(n, bool); inl[out_idx] = False
    # current keyframe's VIO pose: the old pose + drift
    vio_R = rot(rng.normal(0, drift * 0.3, 3)) @ R_wb_old
    vio_T = T_wb_old + rng.normal(0, drift, 3)
    R_true = R_wc.T; t_true = -(R_true @ T_wc)
    return dict(X=X, u=u, vio_R=vio_R, vio_T=vio_T, qic=qic, tic=tic, R_true=R_true, t_true=t_true, inliers=inl)
