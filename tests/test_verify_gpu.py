"""dv_verify_loop (hand-written fp64 RANSAC kernel, one CTA per keyframe pair) vs the oracle (oracle/pnp.py, itself
pinned to cv2.solvePnPRansac): identical inlier masks, poses to 1e-9, same loop decisions; batched and ragged."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pnp_cv2.npz")


@pytest.fixture(scope="module")
def eng():
    from d_vins_b200 import capi
    e = capi.Engine(height=160, width=224)          # no weights needed for this stage
    yield e
    e.close()


def _check(r, o, tol=1e-9):
    assert np.array_equal(r["status"], o["status"])
    assert r["n_inliers"] == o["n_inliers"] and r["has_loop"] == o["has_loop"]
    assert np.abs(r["pnp_R_old"] - o["pnp_R_old"]).max() < tol
    assert np.abs(r["pnp_T_old"] - o["pnp_T_old"]).max() < tol * 10
    assert np.abs(r["relative_t"] - o["relative_t"]).max() < tol * 10
    assert np.abs(r["relative_q"] - o["relative_q"]).max() < tol
    assert abs(r["relative_yaw"] - o["relative_yaw"]) < 1e-7


def test_golden_scenes_batched(eng):
    """All OpenCV golden scenes in ONE call (ragged point counts; same extrinsics): engine == oracle == OpenCV masks."""
    from d_vins_b200 import capi
    from oracle import pnp
    g = np.load(GOLD)
    nc = len(g["cases"])
    X = [g["X_%d" % c] for c in range(nc)]; U = [g["u_%d" % c] for c in range(nc)]
    R = np.stack([g["vio_R_%d" % c] for c in range(nc)]); T = np.stack([g["vio_T_%d" % c] for c in range(nc)])
    p = capi.loop_params(qic=g["qic_0"], tic=g["tic_0"], seed=0)
    res = eng.verify_loop(X, U, R, T, p)
    for c in range(nc):
        o = pnp.verify_loop(X[c], U[c], R[c], T[c], g["qic_0"], g["tic_0"], seed=0)
        _check(res[c], o)
        assert np.array_equal(res[c]["status"], g["cv_mask_%d" % c])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_scenes_and_gates(eng, seed):
    from d_vins_b200 import capi
    from oracle import pnp
    rng = np.random.default_rng(seed)
    scenes = [pnp.synth_scene(int(rng.integers(19, 200)), 0, seed * 50 + i) for i in range(12)]
    for s in scenes:                                   # re-plant outliers with a common extrinsic pair
        pass
    q, t = scenes[0]["qic"], scenes[0]["tic"]
    X, U, R, T = [], [], [], []
    for i, s in enumerate(scenes):
        n = len(s["X"]); n_out = int(rng.integers(0, max(1, n // 2)))
        s2 = pnp.synth_scene(n, n_out, seed * 50 + i)
        X.append(s2["X"]); U.append(s2["u"]); R.append(s2["vio_R"]); T.append(s2["vio_T"])
    X.append(scenes[0]["X"][:10]); U.append(scenes[0]["u"][:10]); R.append(scenes[0]["vio_R"]); T.append(scenes[0]["vio_T"])   # n <= MIN_LOOP_NUM
    for kw in (dict(), dict(max_pose_diff=0.2), dict(max_theta_diff=0.5), dict(ransac_iters=37, seed=99), dict(pnp_inflation=1.0)):
        p = capi.loop_params(qic=q, tic=t, **kw)
        res = eng.verify_loop(X, U, np.stack(R), np.stack(T), p)
        okw = dict(pnp_inflation=p.pnp_inflation, max_theta_diff=p.max_theta_diff, max_pose_diff=p.max_pose_diff,
                   iters=p.ransac_iters, seed=p.seed)
        for i in range(len(X)):
            _check(res[i], pnp.verify_loop(X[i], U[i], R[i], T[i], q, t, **okw))
        assert res[-1]["n_inliers"] == 0 and not res[-1]["has_loop"]
