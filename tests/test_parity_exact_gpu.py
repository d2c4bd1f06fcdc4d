"""End-to-end parity of the DISCRETE outputs against the quantisation-aware oracle (oracle/quant.py: the reference's
arithmetic with the engine's operand precisions, so only the summation order inside dot products differs).

What is asserted, on the three BASELINE frames (EuRoC 480x752, KITTI 376x1241, the config-2 pair) - the achieved numbers
of the last GPU run are committed in profiles/r02_parity_diag.json:
  * LightGlue match pairs: np.array_equal (pairs AND order) on every shape.
  * SuperPoint keypoints: the engine's keypoint SET equals the oracle's up to at most 1 of 512 (rate >= 0.998), nothing
    robustly missing / unexplained; keypoint ORDER is identical for every pair of keypoints whose oracle scores differ
    by more than the float tolerance (adjacent top-512 scores are as close as 3e-6 relative, below any achievable
    agreement of fp16 tensor-core convolutions - DESIGN.md section 2).
  * the float tensors behind them within the quantisation-aware tolerances.
The fp32-oracle tolerance tests stay in test_superpoint_gpu.py / test_lightglue_gpu.py."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

SCORE_ATOL_Q = 1.5e-3      # score map vs the quantisation-aware oracle (measured 0.8e-3 .. 1.0e-3)
SCORE_RTOL_Q = 3.0e-3      # selected keypoint scores, relative (measured 1.5e-3 .. 2.1e-3)
DESC_ATOL_Q = 1.0e-3       # sampled descriptors (measured 2.9e-4 .. 3.7e-4)
MIN_SET_RATE = 0.998       # >= 511 of 512 keypoints identical


def _frames():
    from oracle import synth
    a, b = synth.make_pair(shift=(8, 16))
    return {"euroc": synth.make_frame(480, 752, synth.BASE_SEED), "kitti": synth.make_frame(376, 1241, synth.BASE_SEED + 5),
            "pair_a": a, "pair_b": b}


@pytest.fixture(scope="module")
def engines(weights_file):
    from d_vins_b200 import capi
    es = {(480, 752): capi.Engine(height=480, width=752, weights_path=weights_file),
          (376, 1241): capi.Engine(height=376, width=1241, weights_path=weights_file)}
    yield es
    for e in es.values():
        e.close()


@pytest.mark.parametrize("name", ["euroc", "kitti", "pair_a", "pair_b"])
def test_keypoints_vs_quantisation_aware_oracle(engines, all_weights, name):
    from oracle import quant, weights
    img = _frames()[name]
    H, W = img.shape
    e = engines[(H, W)]
    oq = quant.superpoint_q(weights.sub(all_weights, "sp."), img)
    e.frame_upload(img)
    r = e.sp_detect()
    sm = e.dbg_read("score_map").reshape(oq["score_map"].shape)
    assert np.abs(sm - oq["score_map"]).max() < SCORE_ATOL_Q
    so = [tuple(k) for k in oq["kpts"]]; sr = [tuple(k) for k in r["kpts"]]
    common = set(so) & set(sr)
    rate = len(common) / len(so)
    exact, missing, unexplained, rep = parity.check_keypoints(oq, r, tol=SCORE_ATOL_Q)
    same_pos = sum(a == b for a, b in zip(so, sr))
    print("%s: keypoint set %d/%d (rate %.4f), same rank %d, %s" % (name, len(common), len(so), rate, same_pos, rep))
    assert len(sr) == len(so) == 512
    assert rate >= MIN_SET_RATE, rep
    assert missing == 0 and unexplained == 0, rep
    # order: every pair whose oracle scores are separated by more than the tolerance keeps its relative order
    pos_r = {k: i for i, k in enumerate(sr)}
    ks = [k for k in so if k in pos_r]
    sc = np.array([oq["scores"][so.index(k)] for k in ks]); pr = np.array([pos_r[k] for k in ks])
    sep = sc[:, None] > sc[None, :] * (1.0 + 2 * SCORE_RTOL_Q)          # i clearly above j in the oracle
    assert not np.any(sep & (pr[:, None] > pr[None, :])), "order differs beyond the float tolerance"
    # floats on the common keypoints
    io = np.array([so.index(k) for k in ks]); ig = pr
    assert (np.abs(oq["scores"][io] - r["scores"][ig]) / oq["scores"][io]).max() < SCORE_RTOL_Q
    assert np.abs(oq["desc"][io] - r["desc"][ig]).max() < DESC_ATOL_Q
    assert np.all(np.diff(r["scores"]) <= 0)


def _synthetic_pair(M, N, seed, noise=0.03):
    rng = np.random.default_rng(seed)
    d1 = rng.standard_normal((N, 256)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    perm = rng.permutation(N)[:M]
    d0 = d1[perm] + noise * rng.standard_normal((M, 256)).astype(np.float32)
    d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    k1 = np.stack([rng.uniform(8, 744, N), rng.uniform(8, 472, N)], 1).astype(np.float32)
    k0 = k1[perm] + rng.normal(0, 1, (M, 2)).astype(np.float32)
    return k0, k1, d0, d1


@pytest.mark.parametrize("M,N", [(300, 400), (150, 662), (512, 512), (37, 1000), (1024, 1024)])
def test_match_pairs_bit_exact_synthetic_descriptors(engines, all_weights, M, N):
    from oracle import quant, weights
    e = engines[(480, 752)]
    k0, k1, d0, d1 = _synthetic_pair(M, N, M + N)
    mq, sq = quant.lightglue_q(weights.sub(all_weights, "lg."), k0, k1, d0, d1, 480, 752, 480, 752)
    mg, sg = e.lg_match(k0, k1, d0, d1, 480, 752, 480, 752)
    print("LightGlue %dx%d: %d pairs (oracle %d)" % (M, N, len(mg), len(mq)))
    assert len(mq) >= 10
    # bit-exact pairs and order, except a pair whose match score sits on the acceptance threshold (0.1) within the float
    # tolerance of the score itself (1024 x 1024: one oracle pair at 0.1015): such a pair may fall on either side
    qs, gs = {tuple(m): s for m, s in zip(mq, sq)}, {tuple(m): s for m, s in zip(mg, sg)}
    for k in set(qs) ^ set(gs):
        sc = qs.get(k, gs.get(k))
        assert abs(np.log(sc) - np.log(0.1)) < 0.25, ("pair %s differs away from the threshold: score %.4f" % (k, sc))
    common = [tuple(m) for m in mq if tuple(m) in gs]
    assert common == [tuple(m) for m in mg if tuple(m) in qs], "order of the common pairs"
    sq = np.array([qs[k] for k in common]); sg = np.array([gs[k] for k in common])
    dlog = float(np.abs(np.log(sg) - np.log(sq)).max())
    print("max |dlog score| %.4f" % dlog)
    assert dlog < 0.2       # measured 0.05-0.11 (r02); the fp32-oracle bound on the log assignment is 0.25


def test_match_pairs_bit_exact_on_engine_features(engines, all_weights):
    """BASELINE config 2 (SP 512 x SP 512 of the translated pair) and the EuRoC shape (150 window points x 662): the
    engine's own features through the engine's LightGlue == the quantisation-aware oracle on the same features."""
    from oracle import quant, synth, weights
    e = engines[(480, 752)]
    wl = weights.sub(all_weights, "lg.")
    fr = _frames()
    vio = synth.vio_points(150, 480, 752, synth.BASE_SEED + 3)
    e.frame_upload(fr["pair_a"]); ra = e.sp_detect(); dre_a = e.sp_describe(vio)
    e.frame_upload(fr["pair_b"]); rb = e.sp_detect(); dre_b = e.sp_describe(vio)
    mg, sg = e.lg_match(ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752)
    mq, sq = quant.lightglue_q(wl, ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752)
    print("config-2 pair: %d pairs (oracle %d)" % (len(mg), len(mq)))
    assert len(mq) >= 20 and np.array_equal(mg, mq)
    kp_all = np.concatenate([rb["kpts"].astype(np.float32), vio]); de_all = np.concatenate([rb["desc"], dre_b])
    mg, sg = e.lg_match(vio, kp_all, dre_a, de_all, 480, 752, 480, 752)
    mq, sq = quant.lightglue_q(wl, vio, kp_all, dre_a, de_all, 480, 752, 480, 752)
    print("150 x 662: %d pairs (oracle %d)" % (len(mg), len(mq)))
    assert np.array_equal(mg, mq)
