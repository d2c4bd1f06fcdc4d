"""Cosine kNN over the descriptor bank vs the numpy oracle (exact ids, D within fp32 summation-order tolerance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from d_vins_b200 import capi
    e = capi.Engine(height=64, width=64, bank_capacity=60000)
    yield e
    e.close()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 50, 1000, 10000, 50000])
def test_knn_matches_oracle(eng, n):
    from oracle import knn, synth
    bank, q = synth.make_bank(max(n, 1), seed=100 + n)
    bank = bank[:n]
    eng.bank_import(bank)
    assert eng.bank_size() == n
    for nb in sorted({n, max(n - 49, 0), n // 2}):
        D, I = eng.bank_search(q, nb)
        Do, Io = knn.knn_ip(bank, q, nb)
        assert np.array_equal(I, Io), (n, nb, I, Io)
        fin = np.isfinite(Do)
        assert np.array_equal(np.isfinite(D), fin)
        assert np.abs(D[fin] - Do[fin]).max(initial=0) < 1e-5


def test_knn_ties_lowest_index(eng):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(512).astype(np.float32); v /= np.linalg.norm(v)
    bank = rng.standard_normal((300, 512)).astype(np.float32)
    bank /= np.linalg.norm(bank, axis=1, keepdims=True)
    bank[[7, 130, 255]] = v            # three identical rows -> identical inner products
    eng.bank_import(bank)
    D, I = eng.bank_search(v, 300)
    assert list(I) == [7, 130, 255]


def test_append_and_window(eng):
    from oracle import knn
    rng = np.random.default_rng(2)
    eng.bank_import(np.zeros((0, 512), np.float32))
    rows = []
    for t in range(120):
        d = rng.standard_normal(512).astype(np.float32); d /= np.linalg.norm(d)
        assert eng.bank_append(d) == t
        rows.append(d)
        nb = knn.nb_limit(t)
        D, I = eng.bank_search(d, nb)
        Do, Io = knn.knn_ip(np.stack(rows), d, nb)
        assert np.array_equal(I, Io)
    assert np.allclose(eng.bank_export(), np.stack(rows))


def test_fused_single_launch_equals_two_launch_path(monkeypatch):
    """Default search = ONE launch (limits as kernel arguments, merge in the scan's last block, results written into
    mapped pinned host memory).  DV_KNN_FUSED=0 selects the r01 path (limit upload + scan + merge + two result copies);
    both must return bit-identical (D, I) - repeatedly (the ticket counters reset themselves)."""
    from d_vins_b200 import capi
    from oracle import knn, synth
    bank, q = synth.make_bank(20000, seed=7)
    rng = np.random.default_rng(3)
    qs = [q] + [bank[i] + 0.05 * rng.standard_normal(512).astype(np.float32) for i in (5, 777, 19999)]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DV_KNN_FUSED", mode)
        e = capi.Engine(height=64, width=64, bank_capacity=20000)
        try:
            e.bank_import(bank)
            out = []
            for rep in range(3):
                for x in qs:
                    for nb in (20000, 9951, 130, 3, 1):
                        out.append(e.bank_search(x, nb))
            res[mode] = out
        finally:
            e.close()
    for (D1, I1), (D0, I0) in zip(res["1"], res["0"]):
        assert np.array_equal(I1, I0) and np.array_equal(D1, D0)
    Do, Io = knn.knn_ip(bank, q, 20000)
    assert np.array_equal(res["1"][0][1], Io)
