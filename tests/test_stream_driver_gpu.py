"""the ROS-free keyframe stream driver (csrc/shim/loop_closure.{h,cpp}, plain g++ over the C ABI)
fed with sensor_msgs-shaped records - image + odometry pose + PointCloud channels [norm_x, norm_y, u, v, id]
(pose_graph_node.cpp:330-388, visualization.cpp:399-429) - must produce, keyframe by keyframe, exactly what a
step-by-step replication of KeyFrame ctor -> detectLoop -> findConnection through the C ABI produces, and its
geometric verification must agree with the oracle (oracle/pnp.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H, W = 160, 224
PERIOD, T = 60, 72
N_PTS = 40
F, CX, CY, Z0 = 200.0, 112.0, 80.0, 5.0
DIST = (-0.05, 0.01, 0.0005, -0.0003)            # k1 k2 p1 p2 (radtan)
TOP_THRES, BACK_THRES, PNP_INFL, MIN_LOOP = 0.5, 0.3, 3.5, 4


def lift(u, v):
    """PinholeCamera::liftProjective (PinholeCamera.cc:450-505), float64"""
    k1, k2, p1, p2 = DIST
    mx_d = (1.0 / F) * u + (-CX / F); my_d = (1.0 / F) * v + (-CY / F)

    def dist(x, y):
        mx2, my2, mxy = x * x, y * y, x * y
        rho2 = mx2 + my2
        rad = k1 * rho2 + k2 * rho2 * rho2
        return x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2), y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2)
    dx, dy = dist(mx_d, my_d)
    mx_u, my_u = mx_d - dx, my_d - dy
    for _ in range(7):
        dx, dy = dist(mx_u, my_u)
        mx_u, my_u = mx_d - dx, my_d - dy
    return mx_u, my_u


def test_stream_driver_matches_c_abi_replication(weights_file, tmp_path):
    from d_vins_b200 import build, capi
    from oracle import pnp, synth
    exe = build.build_stream_demo()
    st = synth.Stream(H, W, period=PERIOD, margin=48)
    uv = synth.vio_points(N_PTS, H, W, 77, min_dist=12)
    rng = np.random.default_rng(5)
    qic = pnp._exp_so3(np.array([0.02, -0.01, 0.03])); tic = np.array([0.03, -0.01, 0.02])
    recs = []
    with open(tmp_path / "stream.bin", "wb") as f:
        f.write(struct.pack("<iii", H, W, T))
        for t in range(T):
            y0, x0 = st.offset(t)
            img = st.frame(t)
            # planar world at depth Z0, camera axes == world axes, camera centre follows the crop offset
            C = np.array([(x0 + CX) / F * Z0, (y0 + CY) / F * Z0, 0.0])
            R_wb = qic.T                                        # R_wc = I = R_wb qic
            T_wb = C - R_wb @ tic
            # window points: undistorted pixel -> world; the published uv are the (distorted) pixels themselves
            p3 = np.zeros((N_PTS, 3), np.float32); ch = np.zeros((N_PTS, 5), np.float32)
            for i, (u, v) in enumerate(uv):
                xn, yn = lift(float(u), float(v))
                p3[i] = C + np.array([xn * Z0, yn * Z0, Z0])
                ch[i] = (xn, yn, u, v, 1000 + i)
            drift = pnp._exp_so3(rng.normal(0, 0.004, 3))
            vio_R = drift @ R_wb; vio_T = T_wb + rng.normal(0, 0.01, 3)
            q = pnp.rot_to_quat(vio_R)
            f.write(struct.pack("<d3d4di", 0.1 * t, *vio_T, *q, N_PTS))
            f.write(np.concatenate([p3, ch], 1).astype("<f4").tobytes())
            f.write(img.tobytes())
            recs.append((img, p3, ch, vio_T, q))
    with open(tmp_path / "params.bin", "wb") as f:
        f.write(struct.pack("<8d", F, F, CX, CY, *DIST))
        f.write(qic.astype("<f8").tobytes()); f.write(tic.astype("<f8").tobytes())
        f.write(struct.pack("<3di", TOP_THRES, BACK_THRES, PNP_INFL, MIN_LOOP))
    out = tmp_path / "out.txt"
    r = subprocess.run([exe, weights_file, str(tmp_path / "stream.bin"), str(tmp_path / "params.bin"), str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = out.read_text().splitlines()
    assert len(lines) == T and all(l.startswith("kf ") for l in lines)

    # ---- replication through the C ABI (Python host) + oracle
    prm = capi.loop_params(qic=qic, tic=tic, loop_top_thres=TOP_THRES, loop_back_thres=BACK_THRES, pnp_inflation=PNP_INFL,
                           min_loop_num=MIN_LOOP)
    eng = capi.Engine(height=H, width=W, weights_path=weights_file, max_batch=1, max_vio=64, store_capacity=T + 4,
                      bank_capacity=T + 4)
    n_cand = n_verified = n_loops = 0
    kpts = {}
    try:
        for t, (img, p3, ch, vio_T, q) in enumerate(recs):
            tok = lines[t].split()
            vio = np.zeros((1, 64, 2), np.float32); vio[0, :N_PTS] = ch[:, 2:4]
            eng.batch_upload(img[None]); eng.batch_extract(vio, np.array([N_PTS], np.int32), np.array([t], np.int64))
            assert eng.batch_commit(1) == t
            D, I = eng.batch_search(None, 1)
            kp, _, nsp = eng.store_read(t)
            kpts[t] = kp
            assert int(tok[1]) == t and int(tok[3]) == nsp and int(tok[5]) == N_PTS
            assert [int(x) for x in tok[7:10]] == list(I[0])
            assert np.array_equal(np.array([float(x) for x in tok[11:14]], np.float32), D[0])
            cand = capi.detect_loop(prm, D[0], I[0], t)
            valid = I[0] >= 0
            assert cand == pnp.detect_loop(list(D[0][valid]), list(I[0][valid]), t, TOP_THRES, BACK_THRES)
            assert int(tok[15]) == cand
            n_m = n_in = 0; has = False; info = np.zeros(8)
            if cand >= 0:
                n_cand += 1
                m, _ = eng.batch_match_ex(np.array([t], np.int64), np.array([cand], np.int64), 0, 2, cap=64)[0]
                n_m = len(m)
                if n_m > MIN_LOOP:
                    X = p3[m[:, 0]].astype(np.float64)
                    U = np.array([lift(float(kpts[cand][j, 0]), float(kpts[cand][j, 1])) for j in m[:, 1]], np.float32).astype(np.float64)
                    # Eigen::Quaterniond(w,x,y,z).toRotationMatrix()
                    w_, x_, y_, z_ = q
                    Rq = np.array([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_)],
                                   [2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_)],
                                   [2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)]])
                    res = eng.verify_loop([X], [U], Rq[None], np.asarray(vio_T)[None], prm)[0]
                    o = pnp.verify_loop(X, U, Rq, vio_T, qic, tic, pnp_inflation=PNP_INFL, min_loop_num=MIN_LOOP)
                    assert np.array_equal(res["status"], o["status"]) and res["has_loop"] == o["has_loop"]
                    n_verified += 1
                    n_in, has = res["n_inliers"], res["has_loop"]
                    if has:
                        n_loops += 1
                        info = np.concatenate([res["relative_t"], res["relative_q"], [res["relative_yaw"]]])
                        assert np.abs(info - np.concatenate([o["relative_t"], o["relative_q"], [o["relative_yaw"]]])).max() < 1e-7
            assert int(tok[17]) == n_m and int(tok[19]) == n_in and int(tok[21]) == int(has), (t, tok[14:22], n_m, n_in, has)
            assert np.abs(np.array([float(x) for x in tok[23:31]]) - info).max() < 1e-12
    finally:
        eng.close()
    print("stream driver: %d keyframes, %d loop candidates, %d verified, %d loops" % (T, n_cand, n_verified, n_loops))
    assert n_cand >= 1, "detectLoop never fired: thresholds too strict for the synthetic stream"
