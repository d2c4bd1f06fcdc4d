"""Golden vectors for the geometric-verification step from OpenCV itself: cv2.solvePnPRansac called exactly as
KeyFrame::PnPRANSAC does (loop_fusion/src/keyframe.cpp:835: K = I, no distortion, useExtrinsicGuess = true with the
current VIO pose, 200 iterations, PNP_INFLATION / 460, confidence 0.99) on seeded synthetic loop geometries
(oracle.pnp.synth_scene).  Writes tests/golden/pnp_cv2.npz.  Run in the build container (cv2 4.13 python)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pnp      # noqa: E402

CASES = [(150, 50), (60, 25), (30, 8), (25, 3), (100, 10), (40, 15), (150, 0), (80, 30)]


def main():
    out = {"cases": np.array(CASES, np.int32), "cv2_version": np.array(cv2.__version__)}
    for c, (n, n_out) in enumerate(CASES):
        s = pnp.synth_scene(n, n_out, seed=100 + c)
        R_w_c = s["vio_R"] @ s["qic"]; T_w_c = s["vio_T"] + s["vio_R"] @ s["tic"]
        R0 = R_w_c.T; t0 = -(R0 @ T_w_c)
        rvec, _ = cv2.Rodrigues(R0)
        ok, rv, tv, inl = cv2.solvePnPRansac(s["X"].astype(np.float32), s["u"].astype(np.float32), np.eye(3), None,
                                             rvec.copy(), t0.reshape(3, 1).copy(), True, 200, 3.5 / 460.0, 0.99)
        assert ok
        m = np.zeros(n, np.uint8); m[inl.ravel()] = 1
        R, _ = cv2.Rodrigues(rv)
        for k in ("X", "u", "vio_R", "vio_T", "qic", "tic"):
            out["%s_%d" % (k, c)] = s[k]
        out["cv_mask_%d" % c] = m
        out["cv_R_%d" % c] = R
        out["cv_t_%d" % c] = tv.ravel()
        out["true_mask_%d" % c] = s["inliers"].astype(np.uint8)
        print(c, n, n_out, int(m.sum()), bool(np.array_equal(m.astype(bool), s["inliers"])))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pnp_cv2.npz"), **out)


if __name__ == "__main__":
    main()
