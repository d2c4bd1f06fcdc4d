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


def test_lightglue_oracle_matches_hf_golden(all_weights):
    """tests/golden/lg_hf.npz was produced by HF transformers' LightGlue (an independent port of cvg/LightGlue) with
    the same synthetic weights (tests/golden/make_golden_lg_hf.py): pairs bit-exact, scores / descriptors to fp32
    round-off, for M == N, ragged M < N (HF padding mask) and the D_VINS window-vs-keyframe shape."""
    from oracle import lightglue as olg, weights
    wl = weights.sub(all_weights, "lg.")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lg_hf.npz"))
    h, w = int(g["h"]), int(g["w"])
    for tag in ("equal", "ragged", "window"):
        keep = {}
        m, s = olg.lightglue(wl, g[tag + "_k0"], g[tag + "_k1"], g[tag + "_d0"], g[tag + "_d1"], h, w, h, w, keep=keep)
        assert len(g[tag + "_matches"]) > 0
        assert np.array_equal(m, g[tag + "_matches"]), tag
        assert np.abs(s - g[tag + "_mscores"]).max() < 2e-4
        assert np.abs(keep["x0_8"] - g[tag + "_x0"]).max() < 5e-5
        assert np.abs(keep["x1_8"] - g[tag + "_x1"]).max() < 5e-5


def test_lightglue_oracle_matches_hf_live(all_weights):
    """Same comparison with the HF model executed in this process (skipped where transformers is not importable)."""
    pytest.importorskip("transformers")
    import importlib.util
    from oracle import lightglue as olg, weights
    spec = importlib.util.spec_from_file_location(
        "make_golden_lg_hf", os.path.join(os.path.dirname(__file__), "golden", "make_golden_lg_hf.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    wl = weights.sub(all_weights, "lg.")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lg_hf.npz"))
    h, w = int(g["h"]), int(g["w"])
    hf = mk.hf_model(wl)
    rng = np.random.default_rng(11)
    k0 = g["ragged_k0"] + rng.uniform(-2, 2, g["ragged_k0"].shape).astype(np.float32)   # a case NOT in the fixture
    k1, d0, d1 = g["ragged_k1"], g["ragged_d0"], g["ragged_d1"]
    m0, s0, _, _ = mk.run_hf(hf, k0, k1, d0, d1, h, w)
    valid = np.nonzero(m0 >= 0)[0]
    m, s = olg.lightglue(wl, k0, k1, d0, d1, h, w, h, w)
    assert np.array_equal(m, np.stack([valid, m0[valid]], 1))
    assert np.abs(s - s0[valid]).max() < 2e-4


def test_mixvpr_backbone_matches_torchvision(all_weights):
    """MixVPR's trunk is torchvision ResNet-50 cropped before layer4 (reference README.md:35-45): the oracle's
    restatement must equal torchvision's own module loaded with the same state_dict."""
    tv = pytest.importorskip("torchvision")
    import torch
    from oracle import mixvpr as omix, synth, weights
    wm = weights.sub(all_weights, "mix.")
    net = tv.models.resnet50(weights=None).eval()
    sd = {k[len("backbone.model."):]: torch.from_numpy(v) for k, v in wm.items() if k.startswith("backbone.model.")}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected
    assert all(k.startswith(("layer4.", "fc.")) or k.endswith("num_batches_tracked") for k in missing), missing
    x = torch.from_numpy(omix.preprocess_mix(synth.make_frame(480, 752, synth.BASE_SEED)))[None]
    with torch.no_grad():
        y = net.layer3(net.layer2(net.layer1(net.maxpool(net.relu(net.bn1(net.conv1(x)))))))
        o = omix.backbone(wm, x)
    assert o.shape == (1, 1024, 20, 20)
    assert torch.equal(o, y) or float((o - y).abs().max()) < 1e-5 * float(y.abs().max())


def test_knn_matches_torch_topk():
    """faiss IndexFlatIP (keyframe.cpp:302-314) = exact inner-product top-k; cross-check the oracle against torch."""
    import torch
    from oracle import knn, synth
    bank, q = synth.make_bank(3000, seed=5)
    D, I = knn.knn_ip(bank, q, 2500)
    Dt, It = torch.topk(torch.from_numpy(bank[:2500]) @ torch.from_numpy(q), 3)
    assert np.array_equal(I, It.numpy())
    assert np.abs(D - Dt.numpy()).max() < 1e-5


def test_mix_preprocess_matches_compiled_reference_golden():
    """tests/golden/refpre_mix.npz: outputs of the REFERENCE's own warp_affine_bilinear_and_normalize_plane_kernel_mix
    (preprocess_kernel.cu:348-435, compiled from /root/reference with IEEE flags and run on a B200 by
    tests/test_refpre_gpu.py).  The oracle's numpy restatement reproduces them bit for bit; the reference's shipped
    --use_fast_math build differs from both by one u8 level on the recorded (< 0.02 %) positions."""
    from oracle import mixvpr as omix
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "refpre_mix.npz"))
    for tag in ("euroc", "noise"):
        o = omix.preprocess_mix(g[tag + "_img"])
        assert np.array_equal(o, g[tag + "_ieee"])
        assert len(g[tag + "_fast_levels_diff"]) < 2e-3 * o.size


def test_mixvpr_aggregator_matches_published_module_form(all_weights):
    """amaralibey/MixVPR is not vendored in the reference (README.md:35-45 only names its config).  The oracle's functional
    aggregator is checked against the published module structure restated as nn.Modules - FeatureMixerLayer =
    x + Sequential(LayerNorm, Linear, ReLU, Linear)(x); MixVPR = flatten(2) -> mix -> permute -> channel_proj -> permute ->
    row_proj -> L2-normalise(flatten(1)) - loaded STRICTLY from the same state_dict keys the real checkpoint uses
    (`aggregator.mix.{i}.mix.{0,1,3}`, `aggregator.channel_proj`, `aggregator.row_proj`)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import mixvpr as omix, weights

    class FeatureMixerLayer(nn.Module):
        def __init__(self, in_dim, mlp_ratio=1):
            super().__init__()
            self.mix = nn.Sequential(nn.LayerNorm(in_dim), nn.Linear(in_dim, int(in_dim * mlp_ratio)), nn.ReLU(),
                                     nn.Linear(int(in_dim * mlp_ratio), in_dim))

        def forward(self, x):
            return x + self.mix(x)

    class MixVPR(nn.Module):
        def __init__(self, in_channels=1024, in_h=20, in_w=20, out_channels=256, mix_depth=4, mlp_ratio=1, out_rows=2):
            super().__init__()
            hw = in_h * in_w
            self.mix = nn.Sequential(*[FeatureMixerLayer(hw, mlp_ratio) for _ in range(mix_depth)])
            self.channel_proj = nn.Linear(in_channels, out_channels)
            self.row_proj = nn.Linear(hw, out_rows)

        def forward(self, x):
            x = self.mix(x.flatten(2))
            x = self.channel_proj(x.permute(0, 2, 1)).permute(0, 2, 1)
            x = self.row_proj(x)
            return F.normalize(x.flatten(1), p=2, dim=-1)

    wm = weights.sub(all_weights, "mix.")
    agg = MixVPR().eval()
    sd = {k[len("aggregator."):]: torch.from_numpy(v) for k, v in wm.items() if k.startswith("aggregator.")}
    agg.load_state_dict(sd, strict=True)                       # key names and shapes are exactly the module's
    feat = torch.from_numpy(np.random.default_rng(4).standard_normal((1, 1024, 20, 20)).astype(np.float32)).relu()
    with torch.no_grad():
        ref = agg(feat)[0].numpy()
        got = omix.aggregator(wm, feat).numpy()
    assert ref.shape == (512,)
    assert np.abs(ref - got).max() < 1e-6


def test_filter_matches_properties_randomised():
    """Size-independent properties of the match extraction (filter_matches + LightGlue-ONNX packing): every emitted pair
    is a mutual argmax above the threshold, i0 ascending, no keypoint of either image used twice, nothing valid left out."""
    from hypothesis import given, settings, strategies as st
    from oracle import lightglue as olg

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
    def check(m, n, seed):
        rng = np.random.default_rng(seed)
        L = np.log(rng.uniform(1e-4, 1.0, (m, n))).astype(np.float32)
        if seed % 3 == 0:                                   # plant exact ties: lowest index must win on both axes
            L[:, n // 2] = L[:, 0]
            L[m // 2, :] = L[0, :]
        pairs, ms = olg.filter_matches(L)
        assert np.all(np.diff(pairs[:, 0]) > 0)
        assert len(set(pairs[:, 1].tolist())) == len(pairs)
        m0, m1 = L.argmax(1), L.argmax(0)
        for (i, j), s in zip(pairs, ms):
            assert m0[i] == j and m1[j] == i and s > np.float32(0.1) and np.isclose(s, np.exp(L[i, j]), rtol=1e-6)
        valid = [(i, int(m0[i])) for i in range(m) if m1[m0[i]] == i and np.exp(L[i, m0[i]]) > np.float32(0.1)]
        assert [tuple(p) for p in pairs.tolist()] == valid
    check()


def test_knn_window_and_bruteforce_randomised():
    """kNN over the bank prefix the reference searches (keyframe.cpp:274-294): ids equal a brute-force sort with the
    lowest-index tie rule, padded with (-inf, -1) when fewer than k rows are visible."""
    from hypothesis import given, settings, strategies as st
    from oracle import knn

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 130), st.integers(0, 2 ** 31 - 1))
    def check(t, seed):
        rng = np.random.default_rng(seed)
        bank = rng.standard_normal((t + 1, 512)).astype(np.float32)
        bank /= np.linalg.norm(bank, axis=1, keepdims=True)
        if t > 3:
            bank[1] = bank[0]                                # duplicate rows: the lower index must rank first
        q = bank[t]
        nb = knn.nb_limit(t)
        assert nb == (t - 49 if t >= 50 else t + 1)
        D, I = knn.knn_ip(bank, q, nb)
        sims = bank[:nb] @ q
        order = sorted(range(nb), key=lambda r: (-sims[r], r))[:3]
        assert list(I[:len(order)]) == order
        assert all(i == -1 for i in I[len(order):]) and all(d == -np.inf for d in D[len(order):])
    check()


def test_mixvpr_tail_reorder_identity(all_weights):
    """The engine applies row_proj BEFORE channel_proj (mix.cu k_rowproj_x / k_chanproj_norm):
    z[c', r] = sum_c Wc[c', c] (sum_p Wr[r, p] x[c, p]) + bc[c'] sum_p Wr[r, p] + br[r].  Check that identity against the
    oracle's published order (channel_proj, then row_proj) on the oracle's own weights and a random mixer state."""
    import torch
    from oracle import weights
    wm = weights.sub(all_weights, "mix.")
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(1, 1024, 20, 20, generator=g, dtype=torch.float64)
    t = lambda k: torch.as_tensor(np.asarray(wm[k]), dtype=torch.float64)
    Wc, bc = t("aggregator.channel_proj.weight"), t("aggregator.channel_proj.bias")
    Wr, br = t("aggregator.row_proj.weight"), t("aggregator.row_proj.bias")
    x = feat.flatten(2)[0]                                     # [1024, 400]
    z_pub = torch.nn.functional.linear(torch.nn.functional.linear(x.t(), Wc, bc).t(), Wr, br)   # [256, 2]
    u = x @ Wr.t()                                             # [1024, 2]
    z_eng = Wc @ u + bc[:, None] * Wr.sum(1)[None, :] + br[None, :]
    assert z_pub.shape == z_eng.shape == (256, 2)
    assert float((z_pub - z_eng).abs().max()) < 1e-10 * max(1.0, float(z_pub.abs().max()))
