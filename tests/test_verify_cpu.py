"""Geometric verification / loop decision: the oracle (oracle/pnp.py) against OpenCV's own solvePnPRansac outputs
(tests/golden/pnp_cv2.npz, made by tests/golden/make_golden_pnp.py with the reference's call arguments,
keyframe.cpp:835) and the detectLoop rule against hand-computed cases (pose_graph.cpp:451-509)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pnp_cv2.npz")


def _cases():
    g = np.load(GOLD)
    for c in range(len(g["cases"])):
        yield c, g


def test_oracle_matches_opencv_inliers_and_pose():
    from oracle import pnp
    for c, g in _cases():
        r = pnp.verify_loop(g["X_%d" % c], g["u_%d" % c], g["vio_R_%d" % c], g["vio_T_%d" % c], g["qic_%d" % c],
                            g["tic_%d" % c], seed=0)
        assert np.array_equal(r["status"], g["cv_mask_%d" % c]), c          # the same inlier set as OpenCV
        assert np.array_equal(r["status"], g["true_mask_%d" % c]), c        # == the planted inliers
        # pose: OpenCV's refined camera pose -> body pose (keyframe.cpp:858-866)
        Rc = g["cv_R_%d" % c]; tc = g["cv_t_%d" % c]; qic = g["qic_%d" % c]; tic = g["tic_%d" % c]
        PR = Rc.T @ qic.T
        PT = Rc.T @ (-tc) - PR @ tic
        assert np.abs(r["pnp_R_old"] - PR).max() < 1e-6, (c, np.abs(r["pnp_R_old"] - PR).max())
        assert np.abs(r["pnp_T_old"] - PT).max() < 1e-5, (c, np.abs(r["pnp_T_old"] - PT).max())
        assert r["has_loop"] and r["n_inliers"] == int(g["cv_mask_%d" % c].sum())
        q = r["relative_q"]
        assert abs(np.linalg.norm(q) - 1) < 1e-12


def test_oracle_gates():
    from oracle import pnp
    s = pnp.synth_scene(60, 10, seed=3)
    base = dict(X=s["X"], u=s["u"], vio_R=s["vio_R"], vio_T=s["vio_T"], qic=s["qic"], tic=s["tic"])
    assert pnp.verify_loop(**base)["has_loop"]
    assert not pnp.verify_loop(**base, max_pose_diff=1e-3)["has_loop"]               # translation gate
    assert not pnp.verify_loop(**base, max_theta_diff=1e-6)["has_loop"]              # yaw gate
    assert not pnp.verify_loop(**base, min_loop_num=60)["has_loop"]                  # n <= MIN_LOOP_NUM: PnP skipped
    r = pnp.verify_loop(**base, min_loop_num=60)
    assert r["n_inliers"] == 0 and not r["status"].any()
    # far too few true inliers for a 5-point model: OpenCV-style failure = empty inlier list, pose = the guess
    s2 = pnp.synth_scene(40, 37, seed=4)
    r2 = pnp.verify_loop(s2["X"], s2["u"], s2["vio_R"], s2["vio_T"], s2["qic"], s2["tic"])
    assert not r2["has_loop"] and r2["n_inliers"] < 19
    # sampling stream is reproducible and depends on the seed
    assert pnp.sample_indices(0, 7, 150) == pnp.sample_indices(0, 7, 150)
    assert pnp.sample_indices(0, 7, 150) != pnp.sample_indices(1, 7, 150)
    assert len(set(pnp.sample_indices(5, 3, 6))) == 5


def test_detect_loop_rule():
    """pose_graph.cpp:451-509: top_sim[0] > top_thres AND some top_sim[1..] > back_thres AND frame_index > 50; the
    candidate is the SMALLEST index among the first result and those above back_thres."""
    from oracle import pnp
    assert pnp.detect_loop([0.9, 0.8, 0.7], [400, 120, 300], 500) == 120
    assert pnp.detect_loop([0.9, 0.8, 0.3], [400, 420, 100], 500) == 400          # 100 is below back_thres: ignored
    assert pnp.detect_loop([0.9, 0.3, 0.3], [400, 120, 300], 500) == -1           # nothing backs the top hit
    assert pnp.detect_loop([0.44, 0.43, 0.42], [400, 120, 300], 500) == -1        # top below top_thres
    assert pnp.detect_loop([0.9, 0.8, 0.7], [4, 2, 3], 50) == -1                  # frame_index must exceed 50
    assert pnp.detect_loop([0.9, 0.8, 0.7], [4, 2, 3], 51) == 2
    assert pnp.detect_loop([], [], 500) == -1


def test_c_abi_detect_loop_matches_oracle():
    """dv_detect_loop is pure host arithmetic: callable without a GPU."""
    from d_vins_b200 import capi
    from oracle import pnp
    p = capi.loop_params()
    assert (p.min_loop_num, p.ransac_iters, p.min_frame_index) == (18, 200, 50)
    rng = np.random.default_rng(0)
    for _ in range(300):
        k = 3
        sim = np.sort(rng.uniform(0.2, 1.0, k))[::-1].astype(np.float32)
        idx = rng.integers(0, 1000, k).astype(np.int64)
        fi = int(rng.integers(0, 120))
        if rng.random() < 0.2:                       # faiss padding when nb < k
            sim[-1] = -np.inf; idx[-1] = -1
        want = pnp.detect_loop(list(sim[idx >= 0]), list(idx[idx >= 0]), fi)
        assert capi.detect_loop(p, sim, idx, fi) == want
