"""BASELINE config 4 in miniature: a revisiting synthetic keyframe stream through the batched round API (extract ->
commit -> search -> match) against the oracle run SEQUENTIALLY, one keyframe at a time, in the reference's order
(keyframe.cpp:348-445 extract, :262-346 search over all but the newest 50, :583-673 match against the retrieved one).

Checks, per keyframe: global descriptor cosine, retrieved ids identical wherever the oracle's margin exceeds the float
tolerance, the planted revisit (t -> t - PERIOD) retrieved first on both sides, and LightGlue against the retrieved
keyframe agreeing with the oracle run on the same stored features."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

H, W = 160, 224
PERIOD, T, B = 60, 72, 4
N_VIO = 40


def test_revisiting_stream_matches_sequential_oracle(weights_file, all_weights):
    from d_vins_b200 import capi
    from oracle import knn, lightglue as olg, mixvpr as omix, synth, weights
    wm, wl = weights.sub(all_weights, "mix."), weights.sub(all_weights, "lg.")
    st = synth.Stream(H, W, period=PERIOD, margin=48)
    frames = [st.frame(t) for t in range(T)]
    assert np.array_equal(frames[0], frames[PERIOD]) or np.abs(frames[0].astype(int) - frames[PERIOD].astype(int)).mean() < 8
    vio_pts = synth.vio_points(N_VIO, H, W, 77, min_dist=12)

    eng = capi.Engine(height=H, width=W, weights_path=weights_file, max_batch=B, max_vio=64, store_capacity=T + B,
                      bank_capacity=256)
    try:
        eng.bank_import(np.zeros((0, 512), np.float32))
        vio = np.zeros((B, 64, 2), np.float32); vio[:, :N_VIO] = vio_pts
        nv = np.full((B,), N_VIO, np.int32)
        g_eng, I_eng, D_eng = [], [], []
        for r in range(T // B):
            ids = np.arange(B, dtype=np.int64) + r * B
            eng.batch_upload(np.stack([frames[t] for t in ids]))
            eng.batch_extract(vio, nv, ids)
            first = eng.batch_commit(B)
            assert first == r * B                                   # bank row == keyframe index
            g_eng += [eng.batch_read_global(i) for i in range(B)]
            D, I = eng.batch_search([knn.nb_limit(int(t)) for t in ids])
            D_eng += list(D); I_eng += list(I)

        # ---- sequential oracle: descriptor -> bank -> search, keyframe by keyframe
        bank_o = np.zeros((0, 512), np.float32)
        n_checked = n_revisit = 0
        for t in range(T):
            g = omix.mixvpr(wm, frames[t])
            assert float(g @ g_eng[t]) > parity.GLOBAL_COS_MIN, (t, float(g @ g_eng[t]))
            bank_o = np.concatenate([bank_o, g[None]])
            nb = knn.nb_limit(t)
            Do, Io = knn.knn_ip(bank_o, g, nb)
            assert np.array_equal(I_eng[t] >= 0, Io >= 0)            # same padding with (-inf, -1)
            k = int((Io >= 0).sum())
            assert np.abs(D_eng[t][:k] - Do[:k]).max(initial=0) < 2e-3
            # ids are exact wherever the oracle's ranking margin exceeds the descriptor tolerance
            for j in range(k):
                lo = Do[j] - Do[j + 1] if j + 1 < k else np.inf
                hi = Do[j - 1] - Do[j] if j > 0 else np.inf
                if min(lo, hi) > 4e-3:
                    assert I_eng[t][j] == Io[j], (t, j, I_eng[t], Io, Do)
                    n_checked += 1
            if t >= PERIOD:                                          # the planted loop: t revisits t - PERIOD
                assert Io[0] == t - PERIOD and I_eng[t][0] == t - PERIOD, (t, Io, I_eng[t])
                n_revisit += 1
        assert n_checked >= T and n_revisit == T - PERIOD

        # ---- LightGlue against the retrieved keyframe (the reference matches the window points of the new keyframe
        # against all points of the old one): engine vs the oracle on the engine's own stored features
        q_ids = np.arange(PERIOD, PERIOD + B, dtype=np.int64)
        old_ids = np.array([I_eng[t][0] for t in q_ids], np.int64)
        res = eng.batch_match(q_ids, old_ids)
        for (m, s), t, o in zip(res, q_ids, old_ids):
            kq, dq, nspq = eng.store_read(int(t))
            ko, do, _ = eng.store_read(int(o))
            mo, so = olg.lightglue(wl, kq[nspq:], ko, dq[nspq:], do, H, W, H, W)
            so_set = {tuple(p) for p in mo}; sg_set = {tuple(p) for p in m}
            assert len(so_set) >= 3
            assert len(so_set & sg_set) >= 0.8 * len(so_set), (t, len(so_set & sg_set), len(so_set))
            assert np.all(np.diff(m[:, 0]) > 0)                      # ascending in i0 (keyframe.cpp:623-654 relies on it)
    finally:
        eng.close()
