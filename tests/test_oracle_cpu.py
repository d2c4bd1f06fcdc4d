"""CPU suite: the oracle against the committed golden vectors (produced by the reference's own export modules),
oracle self-consistency, and host-side logic.  No GPU, no /root/reference."""
import os

import numpy as np
import pytest


def test_oracle_matches_reference_golden(all_weights):
    from oracle import superpoint as osp, synth, weights
    ws = weights.sub(all_weights, "sp.")
    for tag, (h, w), ds in (("euroc", (480, 752), 0), ("kitti", (376, 1241), 5)):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sp_%s.npz" % tag))
        assert int(g["ties"]) == 0
        img = synth.make_frame(h, w, synth.BASE_SEED + ds)
        r = osp.superpoint(ws, img)
        assert np.array_equal(r["kpts"], g["kpts"])            # integer outputs: bit exact
        assert np.array_equal(r["scores"], g["scores"])
        assert np.abs(r["desc"][:48] - g["desc_head"]).max() < 1e-6
        assert np.abs(r["desc"].sum(1) - g["desc_rowsum"]).max() < 1e-5
        dr = osp.superpoint_recover(ws, img, g["vio"], feat=r["feat"])
        assert np.abs(dr[:48] - g["desc_r_head"]).max() < 1e-6
        assert np.abs(dr.sum(1) - g["desc_r_rowsum"]).max() < 1e-5
        if tag == "kitti":
            assert r["score_map"].shape == (376, 1240)         # 8*floor(W/8)


def test_nms_properties():
    import torch
    from oracle import superpoint as osp
    rng = np.random.default_rng(0)
    s = torch.from_numpy(rng.random((64, 80)).astype(np.float32))
    n = osp.simple_nms(s)
    keep = n > 0
    ys, xs = np.nonzero(keep.numpy())
    for y, x in zip(ys, xs):       # no two survivors within the suppression radius
        d = np.maximum(np.abs(ys - y), np.abs(xs - x))
        assert np.sum(d <= 4) == 1
    assert np.array_equal(osp.simple_nms(n).numpy(), n.numpy())   # idempotent


def test_normalize_kpts_integer_halves():
    from oracle import superpoint as osp
    k = np.array([[620, 188], [0, 0]], np.float32)
    n = osp.normalize_kpts(k, 1241, 376)        # 1241 // 2 = 620, not 620.5 (deep_net.cpp:839-841)
    assert n[0, 0] == 0.0 and n[0, 1] == 0.0
    assert np.isclose(n[1, 0], -1.0) and np.isclose(n[1, 1], -188 / 620)


def test_knn_windows_and_padding():
    from oracle import knn
    assert knn.nb_limit(0) == 1 and knn.nb_limit(49) == 50 and knn.nb_limit(50) == 1 and knn.nb_limit(120) == 71
    bank = np.eye(4, 512, dtype=np.float32)
    D, I = knn.knn_ip(bank, bank[1], 2)
    assert list(I) == [1, 0, -1] and D[2] == -np.inf
    D2, I2 = knn.knn_reference_style(bank, bank[1], 4)
    assert I2[0] == 1


def test_match_filter_ordering_and_mutual():
    from oracle import lightglue as olg
    L = np.full((4, 5), -9.0, np.float32)
    L[0, 3] = -0.1; L[2, 1] = -0.2; L[3, 1] = -0.3      # row 3 also prefers col 1 but col 1 prefers row 2
    L[1, 4] = -5.0                                      # mutual but exp(-5) < 0.1
    m, s = olg.filter_matches(L)
    assert m.tolist() == [[0, 3], [2, 1]]
    assert np.allclose(s, np.exp([-0.1, -0.2]))


def test_mix_preprocess_quirks():
    from oracle import mixvpr as omix
    img = np.full((480, 752), 100, np.uint8)
    x = omix.preprocess_mix(img)
    a = np.float32(1) / np.float32(255)
    # plane 0 is normalised with the "B" statistics although it holds R (deep_net.cpp:1298-1300)
    assert np.isclose(x[0, 5, 5], (100 * a - 0.406) / 0.225)
    assert np.isclose(x[2, 5, 5], (100 * a - 0.485) / 0.229)
    assert omix.affine_d2i(752, 480)[0] == np.float32(2.35)


def test_weight_file_roundtrip(tmp_path):
    from oracle import weights
    w = weights.synth_superpoint(calibrate=False)
    p = str(tmp_path / "w.dvw")
    weights.save_weights(p, w)
    r = weights.load_weights(p)
    assert list(r.keys()) == list(w.keys())
    assert all(np.array_equal(r[k], w[k]) for k in w)


def test_synthetic_stream_revisits():
    from oracle import synth
    st = synth.Stream(120, 160, period=10, margin=40)
    assert st.offset(0) == st.offset(10)
    a, b = st.frame(0), st.frame(10)
    assert np.abs(a.astype(int) - b.astype(int)).mean() < 4      # same view, fresh noise
