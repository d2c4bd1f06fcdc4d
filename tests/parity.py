"""Shared parity criteria.

Discrete outputs (keypoint indices, NMS survivors, match pairs, kNN ids) are *bit-exact* whenever the discrete
stage is fed identical floats (stage-isolated tests).  End to end, the floats feeding those stages come from fp16
tensor-core convolutions vs the fp32 CPU oracle, so decisions whose margin is below the float tolerance are
ambiguous by construction; the end-to-end criteria below require exact agreement on every decision whose oracle
margin exceeds the tolerance, and bound how many ambiguous ones may exist.
"""
import numpy as np

# stated fp16 tolerances (fp16 operands, fp32 accumulation; vs fp32 CPU oracle)
SCORE_ATOL = 6e-3          # detector score map, absolute (scores are softmax outputs in [0,1])
DESC_COS_MIN = 0.999       # per-descriptor cosine
DESC_ATOL = 6e-3           # per-element |diff| of unit-norm descriptors
GLOBAL_COS_MIN = 0.999     # MixVPR 512-d
LG_L_ATOL = 0.25           # log-assignment entries (log domain): |dL| <= LG_L_ATOL + LG_L_RTOL * |L|
LG_L_RTOL = 0.01           # (similarities reach several hundred with fp16 operands: error scales with |L|)


def nhwc(t):
    """torch NCHW tensor [1,C,H,W] -> numpy [H,W,C]."""
    return t[0].permute(1, 2, 0).contiguous().numpy()


def check_keypoints(oracle, got, tol=SCORE_ATOL, min_exact_frac=0.90):
    """oracle: dict from oracle.superpoint.superpoint (with 'score_map', 'nms'); got: dict kpts/scores.
    Returns (n_exact, n_robust_missing, n_unexplained, report)."""
    smap = oracle["score_map"]
    H, W = smap.shape
    o_set = {(int(x), int(y)) for x, y in oracle["kpts"]}
    g_set = {(int(x), int(y)) for x, y in got["kpts"]}
    assert len(g_set) == len(got["kpts"]), "duplicate keypoints"
    # oracle cut = lowest selected score (if the top-k cut is active)
    nms = oracle["nms"]
    cand = np.sort(nms[nms > 0.0005])[::-1]
    k = len(oracle["kpts"])
    cut = cand[k - 1] if len(cand) > k else 0.0005

    def local_margin(x, y):
        y0, y1, x0, x1 = max(y - 4, 0), min(y + 5, H), max(x - 4, 0), min(x + 5, W)
        win = smap[y0:y1, x0:x1].copy()
        win[y - y0, x - x0] = -np.inf
        return smap[y, x] - win.max()

    robust_missing = 0
    for (x, y) in o_set - g_set:
        if smap[y, x] > cut + 2 * tol and local_margin(x, y) > 2 * tol:
            robust_missing += 1
    unexplained = 0
    for (x, y) in g_set - o_set:
        near_max = local_margin(x, y) > -2 * tol
        near_cut = smap[y, x] > cut - 2 * tol
        inb = 4 <= x < W - 4 and 4 <= y < H - 4
        if not (near_max and near_cut and inb):
            unexplained += 1
    exact = len(o_set & g_set)
    rep = "exact %d/%d, robust-missing %d, unexplained %d" % (exact, len(o_set), robust_missing, unexplained)
    return exact, robust_missing, unexplained, rep
