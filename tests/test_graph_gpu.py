"""The per-keyframe (B = 1) paths replay CUDA graphs captured on first use (SuperPoint encoder, detection post-net,
MixVPR, LightGlue per (m, n)).  DV_GRAPHS=0 runs the same launch sequences eagerly: every output must be bit-identical,
across repeated calls, changing shapes and changing inputs (the graphs' memcpy nodes re-read the pinned tables)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(weights_file):
    from d_vins_b200 import capi
    from oracle import synth
    H, W = 160, 224
    e = capi.Engine(height=H, width=W, weights_path=weights_file, max_vio=64, bank_capacity=64)
    out = []
    try:
        st = synth.Stream(H, W, period=12, margin=48)
        prev = None
        for t in range(5):
            img = st.frame(t)
            vio = synth.vio_points(30 + 5 * (t % 2), H, W, 40 + t, min_dist=10)
            e.frame_upload(img)
            dre = e.sp_describe(vio)
            r = e.sp_detect()
            g = e.mix_describe()
            kp = np.concatenate([r["kpts"].astype(np.float32), vio]); de = np.concatenate([r["desc"], dre])
            old = prev if prev is not None else (kp, de)
            m, s = e.lg_match(vio, old[0], dre, old[1], H, W, H, W)
            m2, s2 = e.lg_match(vio[:20], old[0][:300], dre[:20], old[1][:300], H, W, H, W)     # another (m, n) graph
            m3, s3 = e.lg_match(vio, old[0], dre, old[1], H, W, H, W)                             # replay of the first
            assert np.array_equal(m, m3) and np.array_equal(s, s3)
            out.append((r["kpts"], r["scores"], r["desc"], dre, g, m, s, m2, s2))
            prev = (kp, de)
        if True:   # 3-channel frame: separate encoder graph
            img3 = np.repeat(st.frame(1)[:, :, None], 3, 2)
            e.frame_upload(img3)
            r3 = e.sp_detect()
            out.append((r3["kpts"], r3["scores"], r3["desc"]))
    finally:
        e.close()
    return out


def test_graph_replay_equals_eager(weights_file, monkeypatch):
    monkeypatch.setenv("DV_GRAPHS", "1")
    a = _run(weights_file)
    monkeypatch.setenv("DV_GRAPHS", "0")
    b = _run(weights_file)
    assert len(a) == len(b)
    for ta, tb in zip(a, b):
        for xa, xb in zip(ta, tb):
            assert np.array_equal(xa, xb)
    assert sum(len(t[5]) for t in a[:5]) > 0
