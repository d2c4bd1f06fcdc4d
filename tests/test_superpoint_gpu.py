"""SuperPoint / SP_RE parity: CUDA path (through the C ABI) vs the CPU oracle on identical seeded frames."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(weights_file):
    from d_vins_b200 import capi
    e = capi.Engine(height=480, width=752, weights_path=weights_file)
    yield e
    e.close()


@pytest.fixture(scope="module")
def frame_and_oracle(all_weights):
    from oracle import superpoint as osp, synth, weights
    img = synth.make_frame(480, 752, synth.BASE_SEED)
    keep = {}
    o = osp.superpoint(weights.sub(all_weights, "sp."), img, keep=keep)
    return img, o, keep


def test_nms_topk_bit_exact_on_identical_scores(eng, frame_and_oracle):
    """Integer stage in isolation: same f32 score map in -> same NMS survivors, same keypoints, same order."""
    _, o, _ = frame_and_oracle
    nms, kp, sc = eng.dbg_nms_select(o["score_map"])
    assert np.array_equal(nms, o["nms"])
    assert np.array_equal(kp, o["kpts"])
    assert np.array_equal(sc, o["scores"])


def test_nms_topk_few_candidates_row_major(eng):
    """<= k candidates stay in row-major order, unsorted (export/superpoint.py:76-77); plateaus all survive."""
    from oracle import superpoint as osp
    import torch
    rng = np.random.default_rng(3)
    s = np.zeros((480, 752), np.float32)
    ys = rng.integers(10, 470, 200); xs = rng.integers(10, 740, 200)
    s[ys, xs] = rng.uniform(0.01, 0.9, 200).astype(np.float32)
    s[100:103, 200:203] = 0.5          # plateau: every pixel equals its window max
    s[0:6, 0:6] = 0.7                  # inside the border: removed
    nms_o = osp.simple_nms(torch.from_numpy(s))
    kp_o, sc_o, _ = osp.select_keypoints(nms_o, 512)
    nms, kp, sc = eng.dbg_nms_select(s)
    assert np.array_equal(nms, nms_o.numpy())
    assert np.array_equal(kp, kp_o.numpy().astype(np.int32))
    assert np.array_equal(sc, sc_o.numpy())


def test_encoder_activations(eng, frame_and_oracle):
    img, o, keep = frame_and_oracle
    eng.frame_upload(img)
    eng.sp_detect()
    import torch
    import torch.nn.functional as F
    # conv1a runs inside the conv1b kernel (tensor cores, conv_halo.cu FUSE == 2) and never reaches memory: the first
    # observable activation is conv1b after its fused 2x2 max-pool
    keep = dict(keep)
    keep["conv1b_pool"] = F.max_pool2d(torch.as_tensor(keep["conv1b"]), 2, 2)
    for name, key in (("conv1b_pool", "conv1b_pool"), ("conv2a", "conv2a"), ("conv3a", "conv3a"), ("conv4a", "conv4a"),
                      ("conv4b", "conv4b")):
        ref = parity.nhwc(keep[key])
        try:
            got = eng.dbg_read(name).reshape(ref.shape)
        except Exception:      # 64-channel layers live in the channel-blocked layout [C/8][H][W][8] (conv_halo.cu)
            hh, ww, cc = ref.shape
            got = eng.dbg_read(name + "_blocked").reshape(cc // 8, hh, ww, 8).transpose(1, 2, 0, 3).reshape(hh, ww, cc)
        err = np.abs(got - ref).max()
        assert err < 0.03 * max(1.0, np.abs(ref).max()), (name, err)
    ref = parity.nhwc(keep["logits"])
    got = eng.dbg_read("logits").reshape(60, 94, 80)[:, :, :65]
    assert np.abs(got - ref).max() < 0.08, np.abs(got - ref).max()
    sm = eng.dbg_read("score_map").reshape(480, 752)
    assert np.abs(sm - o["score_map"]).max() < parity.SCORE_ATOL


def test_conv1a_unfused_paths(weights_file, frame_and_oracle, monkeypatch):
    """DV_SP_FUSE1A=0: separate CUDA-core conv1a kernel (also the path 3-channel frames take); =1: CUDA-core conv1a in
    conv1b's producer.  Both must agree with the oracle like the default tensor-core fusion does."""
    from d_vins_b200 import capi
    img, o, keep = frame_and_oracle
    for mode in ("0", "1"):
        monkeypatch.setenv("DV_SP_FUSE1A", mode)
        e = capi.Engine(height=480, width=752, weights_path=weights_file)
        try:
            e.frame_upload(img)
            e.sp_detect()
            if mode == "0":
                ref = parity.nhwc(keep["conv1a"])
                hh, ww, cc = ref.shape
                got = e.dbg_read("conv1a_blocked").reshape(cc // 8, hh, ww, 8).transpose(1, 2, 0, 3).reshape(hh, ww, cc)
                assert np.abs(got - ref).max() < 0.03 * max(1.0, np.abs(ref).max())
            sm = e.dbg_read("score_map").reshape(480, 752)
            assert np.abs(sm - o["score_map"]).max() < parity.SCORE_ATOL
        finally:
            e.close()


def test_superpoint_end_to_end(eng, frame_and_oracle):
    img, o, _ = frame_and_oracle
    eng.frame_upload(img)
    r = eng.sp_detect()
    assert len(r["kpts"]) == len(o["kpts"]) == 512
    exact, missing, unexplained, rep = parity.check_keypoints(o, r)
    print(rep)
    assert missing == 0 and unexplained == 0, rep
    assert exact >= 0.9 * len(o["kpts"]), rep
    # floats on the common keypoints
    oi = {(int(x), int(y)): i for i, (x, y) in enumerate(o["kpts"])}
    pairs = [(oi[(int(x), int(y))], j) for j, (x, y) in enumerate(r["kpts"]) if (int(x), int(y)) in oi]
    io, ig = np.array(pairs).T
    assert np.abs(o["scores"][io] - r["scores"][ig]).max() < parity.SCORE_ATOL
    cos = (o["desc"][io] * r["desc"][ig]).sum(1)
    assert cos.min() > parity.DESC_COS_MIN, cos.min()
    assert np.abs(o["desc"][io] - r["desc"][ig]).max() < parity.DESC_ATOL
    assert np.allclose(np.linalg.norm(r["desc"], axis=1), 1.0, atol=1e-5)
    # normalised keypoints: integer halves (deep_net.cpp:633-659)
    from oracle import superpoint as osp
    assert np.array_equal(r["kpts_norm"], osp.normalize_kpts(r["kpts"], 752, 480))
    # scores sorted descending (top-k active)
    assert np.all(np.diff(r["scores"]) <= 0)


def test_sp_recover(eng, frame_and_oracle, all_weights):
    from oracle import superpoint as osp, synth, weights
    img, o, _ = frame_and_oracle
    vio = synth.vio_points(150, 480, 752, synth.BASE_SEED + 3)
    ref = osp.superpoint_recover(weights.sub(all_weights, "sp."), img, vio, feat=o["feat"])
    eng.frame_upload(img)
    got = eng.sp_describe(vio)
    cos = (ref * got).sum(1)
    assert cos.min() > parity.DESC_COS_MIN, cos.min()
    assert np.abs(ref - got).max() < parity.DESC_ATOL


def test_golden_fixture_matches_engine(eng):
    """Committed golden vectors produced by the reference's own export/superpoint.py (tests/golden/make_golden.py)."""
    import os
    from oracle import synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sp_euroc.npz"))
    img = synth.make_frame(480, 752, synth.BASE_SEED)
    eng.frame_upload(img)
    r = eng.sp_detect()
    gs = {(int(x), int(y)) for x, y in g["kpts"]}
    rs = {(int(x), int(y)) for x, y in r["kpts"]}
    assert len(gs & rs) >= 0.9 * len(gs)
    de = eng.sp_describe(g["vio"])
    cos = (de[:48] * g["desc_r_head"]).sum(1)
    assert cos.min() > parity.DESC_COS_MIN


def test_rejects_wrong_size(eng):
    from d_vins_b200 import capi
    with pytest.raises(capi.DvError):
        eng.frame_upload(np.zeros((100, 100), np.uint8))


def test_kitti_shape_376x1241(weights_file, all_weights):
    """BASELINE config 5 shape: W not divisible by 8 -> score map 376x1240, pooled widths 620/310/155."""
    import os
    from d_vins_b200 import capi
    from oracle import superpoint as osp, synth, weights
    e = capi.Engine(height=376, width=1241, weights_path=weights_file)
    try:
        img = synth.make_frame(376, 1241, synth.BASE_SEED + 5)
        o = osp.superpoint(weights.sub(all_weights, "sp."), img)
        e.frame_upload(img)
        r = e.sp_detect()
        sm = e.dbg_read("score_map").reshape(376, 1240)
        assert np.abs(sm - o["score_map"]).max() < parity.SCORE_ATOL
        exact, missing, unexplained, rep = parity.check_keypoints(o, r)
        assert missing == 0 and unexplained == 0 and exact >= 0.9 * len(o["kpts"]), rep
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sp_kitti.npz"))
        gs = {(int(x), int(y)) for x, y in g["kpts"]}
        assert len(gs & {(int(x), int(y)) for x, y in r["kpts"]}) >= 0.9 * len(gs)
        # integer halves: 1241 // 2 = 620
        assert np.array_equal(r["kpts_norm"], osp.normalize_kpts(r["kpts"], 1241, 376))
        nms, kp, sc = e.dbg_nms_select(o["score_map"])
        assert np.array_equal(kp, o["kpts"]) and np.array_equal(nms, o["nms"])
    finally:
        e.close()
