"""Batched keyframe-round API (throughput path) vs the per-keyframe API and the oracle's orchestration contract."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(weights_file):
    from d_vins_b200 import capi
    e = capi.Engine(height=480, width=752, weights_path=weights_file, max_batch=4, max_vio=160, store_capacity=16,
                    bank_capacity=4096)
    yield e
    e.close()


def test_batch_equals_per_frame(eng):
    from oracle import knn, synth
    st = synth.Stream(480, 752, period=12, margin=64)
    b = 4
    frames = np.stack([st.frame(t) for t in range(b)])
    vio = np.zeros((b, 160, 2), np.float32); nv = np.array([150, 120, 21, 0], np.int32)
    for i in range(b):
        vio[i, :max(nv[i], 1)] = synth.vio_points(max(nv[i], 1), 480, 752, 50 + i)[:max(nv[i], 1)]
    ids = np.arange(b, dtype=np.int64)
    eng.bank_import(np.zeros((0, 512), np.float32))
    eng.batch_upload(frames)
    eng.batch_extract(vio, nv, ids)
    first = eng.batch_commit(b)
    assert first == 0 and eng.bank_size() == b
    g_batch = [eng.batch_read_global(i) for i in range(b)]
    stored = [eng.store_read(i) for i in range(b)]
    D, I = eng.batch_search([knn.nb_limit(int(t)) for t in ids])
    bank = eng.bank_export()
    for i in range(b):
        eng.frame_upload(frames[i])
        r = eng.sp_detect()
        g = eng.mix_describe()
        kp, de, nsp = stored[i]
        assert nsp == len(r["kpts"])
        # same kernels, same inputs: batched and per-frame paths agree bit for bit
        assert np.array_equal(kp[:nsp], r["kpts"].astype(np.float32))
        assert np.array_equal(de[:nsp], r["desc"])
        assert np.array_equal(g, g_batch[i]) and np.array_equal(bank[i], g)
        assert len(kp) == nsp + nv[i]
        if nv[i]:
            re = eng.sp_describe(vio[i, :nv[i]])
            assert np.array_equal(de[nsp:], re)                      # keyframe.cpp:401-432: SP ++ SP_RE
            assert np.array_equal(kp[nsp:], vio[i, :nv[i]])
        Do, Io = knn.knn_ip(bank, g_batch[i], knn.nb_limit(i))
        assert np.array_equal(I[i], Io)


def test_batch_match_uses_store(eng):
    from oracle import synth
    st = synth.Stream(480, 752, period=12, margin=64)
    # frames 0 and 12 see the same scene (loop); reuse ids 0..3 from the previous test + add frame 12
    frames = np.stack([st.frame(12)])
    vio = np.zeros((1, 160, 2), np.float32); nv = np.array([150], np.int32)
    vio[0, :150] = synth.vio_points(150, 480, 752, 50)       # same window points as frame 0
    eng.batch_upload(frames)
    eng.batch_extract(vio, nv, np.array([12], np.int64))
    res = eng.batch_match(np.array([12], np.int64), np.array([0], np.int64))
    m, s = res[0]
    kq, dq, nspq = eng.store_read(12)
    ko, do, _ = eng.store_read(0)
    m2, s2 = eng.lg_match(kq[nspq:], ko, dq[nspq:], do, 480, 752, 480, 752)
    assert np.array_equal(m, m2) and np.array_equal(s, s2)
    assert len(m) >= 3
    # window point i of frame 12 should match the *same* VIO point of frame 0 (appended after its SP points);
    # the seeded random-weight matcher is weak, so only a majority is required
    _, _, nsp0 = eng.store_read(0)
    ok = np.mean(m[:, 1] == nsp0 + m[:, 0])
    assert ok >= 0.5, ok


def test_upload_runs_one_round_ahead(eng):
    """dv_batch_upload is asynchronous and double-buffered: round R+1's frames may be queued right after round R's
    extraction; commit / search / match of round R keep working on R, and R+1's extraction sees R+1's frames."""
    from oracle import knn, synth
    st = synth.Stream(480, 752, period=12, margin=64)
    b = 4
    fa = np.stack([st.frame(t) for t in range(20, 20 + b)])
    fb = np.stack([st.frame(t) for t in range(30, 30 + b)])
    vio = np.zeros((b, 160, 2), np.float32); nv = np.full((b,), 30, np.int32)
    for i in range(b):
        vio[i, :30] = synth.vio_points(30, 480, 752, 90 + i)
    ids_a, ids_b = np.arange(100, 100 + b, dtype=np.int64), np.arange(104, 104 + b, dtype=np.int64)
    eng.bank_import(np.zeros((0, 512), np.float32))
    # reference: strictly sequential
    eng.batch_upload(fa); eng.batch_extract(vio, nv, ids_a)
    ga = [eng.batch_read_global(i) for i in range(b)]
    eng.batch_upload(fb); eng.batch_extract(vio, nv, ids_b)
    gb = [eng.batch_read_global(i) for i in range(b)]
    assert not np.array_equal(ga[0], gb[0])
    # pipelined: B uploaded while A is still the current round
    eng.bank_import(np.zeros((0, 512), np.float32))
    eng.batch_upload(fa); eng.batch_extract(vio, nv, ids_a)
    eng.batch_upload(fb)                                            # prefetch
    assert all(np.array_equal(eng.batch_read_global(i), ga[i]) for i in range(b))
    assert eng.batch_commit(b) == 0
    D, I = eng.batch_search([knn.nb_limit(int(t)) for t in range(b)])
    assert np.array_equal(I[:, 0], np.arange(b))                    # every frame retrieves itself (nb = t + 1)
    res = eng.batch_match(ids_a, ids_a)
    assert len(res) == b
    eng.batch_extract(vio, nv, ids_b)                               # consumes the prefetched frames
    assert all(np.array_equal(eng.batch_read_global(i), gb[i]) for i in range(b))


def test_session_save_load_roundtrip(eng, weights_file, tmp_path):
    """the descriptor bank and the keyframes' local features survive a save / load into a NEW engine
    (dv_bank_export/import + dv_store_read/put): retrieval and a LightGlue match against a restored keyframe are
    bit-identical to the original session's."""
    from d_vins_b200 import capi
    from oracle import knn, synth
    st = synth.Stream(480, 752, period=12, margin=64)
    b = 4
    frames = np.stack([st.frame(t) for t in range(b)])
    vio = np.zeros((b, 160, 2), np.float32); nv = np.full((b,), 40, np.int32)
    for i in range(b):
        vio[i, :40] = synth.vio_points(40, 480, 752, 70 + i)
    ids = np.arange(b, dtype=np.int64)
    eng.bank_import(np.zeros((0, 512), np.float32))
    eng.batch_upload(frames); eng.batch_extract(vio, nv, ids)
    assert eng.batch_commit(b) == 0
    q = eng.batch_read_global(2)
    D0, I0 = eng.bank_search(q, b)
    m0 = eng.batch_match(np.array([3], np.int64), np.array([1], np.int64))[0]
    path = str(tmp_path / "session.dvs")
    eng.save_session(path, ids)
    e2 = capi.Engine(height=480, width=752, weights_path=weights_file, max_batch=4, max_vio=160, store_capacity=16,
                     bank_capacity=4096)
    try:
        assert list(e2.load_session(path)) == list(ids)
        assert e2.bank_size() == b
        D1, I1 = e2.bank_search(q, b)
        assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
        for t in ids:
            k0, d0, n0 = eng.store_read(int(t)); k1, d1, n1 = e2.store_read(int(t))
            assert n0 == n1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)
        m1 = e2.batch_match(np.array([3], np.int64), np.array([1], np.int64))[0]
        assert np.array_equal(m0[0], m1[0]) and np.array_equal(m0[1], m1[1])
    finally:
        e2.close()


def test_store_ring_eviction_is_per_pair_status(weights_file):
    """The feature store is a ring: a keyframe whose slot was reused is reported as k_out = -1 for ITS pair only, the
    other pairs of the batch are still matched (the whole batch used to fail with DV_ERR_INVALID)."""
    from d_vins_b200 import capi
    from oracle import synth
    H, W, b = 160, 224, 2
    e = capi.Engine(height=H, width=W, weights_path=weights_file, max_batch=b, max_vio=64, store_capacity=4,
                    bank_capacity=64)
    try:
        st = synth.Stream(H, W, period=12, margin=48)
        vio = np.zeros((b, 64, 2), np.float32); nv = np.full((b,), 40, np.int32)
        vio[:, :40] = synth.vio_points(40, H, W, 5, min_dist=12)
        for r in range(3):                                   # 6 keyframes through a 4-slot ring: 0 and 1 are evicted
            ids = np.arange(b, dtype=np.int64) + r * b
            e.batch_upload(np.stack([st.frame(int(t)) for t in ids]))
            e.batch_extract(vio, nv, ids)
            e.batch_commit(b)
        assert e.store_lookup(0)[0] == -1 and e.store_lookup(1)[0] == -1
        assert e.store_lookup(2)[0] == 0 and e.store_lookup(5) == (0, e.store_lookup(5)[1], e.store_lookup(5)[2])
        assert list(e.store_lookup_many([0, 3, 99, 5])) == [-1, 0, -1, 0]
        res = e.batch_match(np.array([4, 5], np.int64), np.array([0, 3], np.int64))
        assert list(e.last_match_status >= 0) == [False, True]
        assert len(res[0][0]) == 0 and e.last_match_status[0] == -1
        ref = e.batch_match(np.array([5], np.int64), np.array([3], np.int64))[0]
        assert np.array_equal(res[1][0], ref[0]) and np.array_equal(res[1][1], ref[1])
        with pytest.raises(capi.DvError):
            e.store_read(0)
    finally:
        e.close()


def test_engine_window_rule_and_global_only_round(eng):
    """dv_batch_search(nb_limit = NULL) applies keyframe.cpp:274-282 with cfg.exclude_recent itself;
    dv_batch_describe_global runs MixVPR alone and yields the same descriptors as the full extraction."""
    from oracle import knn, synth
    st = synth.Stream(480, 752, period=12, margin=64)
    b = 4
    frames = np.stack([st.frame(t) for t in range(40, 40 + b)])
    vio = np.zeros((b, 160, 2), np.float32); nv = np.zeros((b,), np.int32)
    bank, _ = synth.make_bank(300, seed=11)
    eng.bank_import(bank)
    eng.batch_upload(frames); eng.batch_extract(vio, nv, np.arange(200, 200 + b, dtype=np.int64))
    g_full = [eng.batch_read_global(i) for i in range(b)]
    first = eng.batch_commit(b)
    assert first == 300
    D0, I0 = eng.batch_search([knn.nb_limit(300 + i) for i in range(b)])
    D1, I1 = eng.batch_search(None, b)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
    eng.bank_import(bank)
    eng.batch_upload(frames); eng.batch_describe_global(b)
    assert all(np.array_equal(eng.batch_read_global(i), g_full[i]) for i in range(b))
    assert eng.batch_commit(b) == 300
    D2, I2 = eng.batch_search(None, b)
    assert np.array_equal(I0, I2) and np.array_equal(D0, D2)


def test_match_sp_part_equals_host_vector_match(eng):
    """BASELINE config 2 path: dv_batch_match_ex(DV_PART_SP, DV_PART_SP) on stored keyframes == dv_lg_match on the same
    SuperPoint keypoints / descriptors passed as host vectors."""
    from oracle import synth
    a, bimg = synth.make_pair(shift=(8, 16))
    vio = np.zeros((2, 160, 2), np.float32); nv = np.zeros((2,), np.int32)
    eng.batch_upload(np.stack([a, bimg])); eng.batch_extract(vio, nv, np.array([900, 901], np.int64))
    m, s = eng.batch_match_sp(np.array([900], np.int64), np.array([901], np.int64))[0]
    ka, da, na = eng.store_read(900); kb, db, nb = eng.store_read(901)
    assert na == len(ka) == 512 and nb == len(kb) == 512
    m2, s2 = eng.lg_match(ka, kb, da, db, 480, 752, 480, 752)
    assert len(m) > 20 and np.array_equal(m, m2) and np.array_equal(s, s2)


def test_config_validation(weights_file):
    from d_vins_b200 import capi
    for kw in (dict(max_vio=600), dict(max_kpts=1024, max_vio=64), dict(max_batch=8, store_capacity=4)):
        with pytest.raises(capi.DvError):
            capi.Engine(height=160, width=224, weights_path=weights_file, **kw)


def test_match_begin_end_equals_synchronous_match(eng):
    """dv_batch_match_begin / _end: the next round is uploaded and extracted between the two halves (its kernels queue
    behind the match); the collected results equal the synchronous dv_batch_match of the same pairs, and a second begin
    without an end is refused."""
    from d_vins_b200 import capi
    from oracle import synth
    st = synth.Stream(480, 752, period=12, margin=64)
    b = 4
    vio = np.zeros((b, 160, 2), np.float32); nv = np.full((b,), 150, np.int32)
    for i in range(b):
        vio[i, :150] = synth.vio_points(150, 480, 752, 70 + i)
    eng.bank_import(np.zeros((0, 512), np.float32))
    ids0, ids1, ids2 = (np.arange(b, dtype=np.int64) + k * b + 1000 for k in range(3))
    for ids, t0 in ((ids0, 0), (ids1, 4)):
        eng.batch_upload(np.stack([st.frame(t0 + t) for t in range(b)]))
        eng.batch_extract(vio, nv, ids)
        eng.batch_commit(b)
    ref = eng.batch_match(ids1, ids0)
    eng.batch_match_begin(ids1, ids0)
    with pytest.raises(capi.DvError):
        eng.batch_match_begin(ids1, ids0)
    eng.batch_upload(np.stack([st.frame(8 + t) for t in range(b)]))      # the next round, queued behind the match
    eng.batch_extract(vio, nv, ids2)
    got = eng.batch_match_end()
    assert len(got) == len(ref) and sum(len(m) for m, _ in ref) > 0
    for (m0, s0), (m1, s1) in zip(ref, got):
        assert np.array_equal(m0, m1) and np.array_equal(s0, s1)
