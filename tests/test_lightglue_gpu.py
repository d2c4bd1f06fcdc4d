"""LightGlue parity: match extraction bit-exact on identical log-assignment matrices; end-to-end matches vs the oracle
with the margin-aware criterion; log-assignment matrix within a stated tolerance."""
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


def _synthetic_pair(M, N, seed, noise=0.03):
    rng = np.random.default_rng(seed)
    d1 = rng.standard_normal((N, 256)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    perm = rng.permutation(N)[:M]
    d0 = d1[perm] + noise * rng.standard_normal((M, 256)).astype(np.float32)
    d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    k1 = np.stack([rng.uniform(8, 744, N), rng.uniform(8, 472, N)], 1).astype(np.float32)
    k0 = k1[perm] + rng.normal(0, 1, (M, 2)).astype(np.float32)
    return k0, k1, d0, d1, perm


@pytest.mark.parametrize("m,n,seed", [(64, 64, 0), (300, 400, 1), (150, 662, 2), (1024, 1024, 3), (10, 10, 4), (777, 130, 5)])
def test_match_extraction_bit_exact(eng, m, n, seed):
    """Integer stage in isolation: identical L in -> identical (i0-ascending) pairs and scores' exp."""
    from oracle import lightglue as olg
    rng = np.random.default_rng(seed)
    L = (rng.standard_normal((m, n)) * 3 - 6).astype(np.float32)
    k = min(m, n)
    idx = rng.permutation(k)[: k // 2]
    L[idx, idx[::-1]] = rng.uniform(-2.0, -0.01, len(idx)).astype(np.float32)   # planted mutual maxima
    L[3 % m, :] = L[3 % m, 0]                                                  # a fully tied row -> lowest index
    mo, so = olg.filter_matches(L)
    mg, sg = eng.dbg_match_extract(L)
    assert np.array_equal(mo, mg)
    assert np.allclose(so, sg, rtol=1e-6, atol=0)
    assert np.all(np.diff(mg[:, 0]) > 0) if len(mg) > 1 else True


@pytest.mark.parametrize("M,N", [(300, 400), (150, 662), (512, 512), (37, 1000)])
def test_lightglue_synthetic_descriptors(eng, all_weights, M, N):
    from oracle import lightglue as olg, weights
    wl = weights.sub(all_weights, "lg.")
    k0, k1, d0, d1, perm = _synthetic_pair(M, N, M + N)
    keep = {}
    mo, so = olg.lightglue(wl, k0, k1, d0, d1, 480, 752, 480, 752, keep)
    mg, sg = eng.lg_match(k0, k1, d0, d1, 480, 752, 480, 752)
    Lg = eng.dbg_read("lg_L").reshape(1024, 1024)[:M, :N]
    Lo = keep["L"]
    # tolerance where it matters: around each row's maximum
    rows = np.arange(M)
    jo = Lo.argmax(1)
    dL = np.abs(Lg[rows, jo] - Lo[rows, jo]) - parity.LG_L_RTOL * np.abs(Lo[rows, jo])
    assert dL.max() < parity.LG_L_ATOL, dL.max()
    dall = np.abs(Lg - Lo) - parity.LG_L_RTOL * np.abs(Lo)
    assert dall.max() < 2 * parity.LG_L_ATOL, dall.max()
    # margin-aware pair agreement: every oracle pair whose row/column margins exceed the tolerance must be found
    so_set = {(int(i), int(j)) for i, j in mo}
    sg_set = {(int(i), int(j)) for i, j in mg}
    srt = np.sort(Lo, 1)
    row_gap = srt[:, -1] - srt[:, -2] if N > 1 else np.full(M, np.inf)
    tol = 2 * parity.LG_L_ATOL
    miss = [(i, j) for (i, j) in so_set - sg_set
            if row_gap[i] > tol and np.exp(Lo[i, j]) > 0.1 * np.exp(tol)]
    extra = [(i, j) for (i, j) in sg_set - so_set if np.exp(Lo[i, j]) < 0.1 * np.exp(-tol) or Lo[i].max() - Lo[i, j] > tol]
    assert not miss and not extra, (miss[:5], extra[:5])
    assert len(so_set & sg_set) >= 0.95 * len(so_set), (len(so_set & sg_set), len(so_set))
    assert np.all(np.diff(mg[:, 0]) > 0)
    common = {p: s for p, s in zip(map(tuple, mo), so)}
    for p, s in zip(map(tuple, mg), sg):
        if p in common:
            assert abs(np.log(s) - np.log(common[p])) < parity.LG_L_ATOL


def test_lightglue_on_superpoint_features(eng, all_weights):
    """Config 2 of BASELINE.json: SuperPoint + LightGlue, 512-kpt pair, 480x752."""
    from oracle import lightglue as olg, superpoint as osp, synth, weights
    ws, wl = weights.sub(all_weights, "sp."), weights.sub(all_weights, "lg.")
    a, b = synth.make_pair(shift=(8, 16))
    ra, rb = osp.superpoint(ws, a), osp.superpoint(ws, b)
    mo, so = olg.lightglue(wl, ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752)
    # LightGlue stage on identical (oracle) features
    mg, sg, mk0, mk1 = eng.lg_match(ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752, want_mkpts=True)
    so_set = {(int(i), int(j)) for i, j in mo}; sg_set = {(int(i), int(j)) for i, j in mg}
    assert len(so_set) > 20
    assert len(so_set & sg_set) >= 0.9 * len(so_set), (len(so_set & sg_set), len(so_set), len(sg_set))
    assert np.abs(mk0 - ra["kpts"][mg[:, 0]]).max() < 1e-3 and np.abs(mk1 - rb["kpts"][mg[:, 1]]).max() < 1e-3


def test_lightglue_rejects_small_inputs(eng):
    from d_vins_b200 import capi
    z = np.zeros((5, 2), np.float32); d = np.zeros((5, 256), np.float32)
    with pytest.raises(capi.DvError):
        eng.lg_match(z, z, d, d, 480, 752, 480, 752)


def test_fused_ffn0_matches_two_kernel_path():
    """lg_ffn0.cu (ffn.0 + LayerNorm + GELU in one four-CTA-cluster kernel, row statistics exchanged through
    distributed shared memory) against the GEMM + k_lg_ln_gelu pair on the same inputs (150 x 662 and 1024 x 1024):
    the same match pairs up to pairs whose score sits on the acceptance threshold, scores within the float tolerance
    (the fused kernel sums the LayerNorm statistics in a different order)."""
    import json
    import os
    import subprocess
    import sys
    code = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
from d_vins_b200 import capi
e = capi.Engine(height=480, width=752, weights_path=bench.make_weights())
out = []
for (M, N, seed) in [(150, 662, 5), (1024, 1024, 2048)]:
    rng = np.random.default_rng(seed)
    d1 = rng.standard_normal((N, 256)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    perm = rng.permutation(N)[:M]
    d0 = d1[perm] + 0.03 * rng.standard_normal((M, 256)).astype(np.float32); d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    k1 = np.stack([rng.uniform(8, 744, N), rng.uniform(8, 472, N)], 1).astype(np.float32)
    k0 = k1[perm] + rng.normal(0, 1, (M, 2)).astype(np.float32)
    m, s = e.lg_match(k0, k1, d0, d1, 480, 752, 480, 752)
    out.append({"m": m.tolist(), "s": s.tolist()})
print(json.dumps(out))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for mode in ("0", "2"):
        env = dict(os.environ, DV_LG_FUSE_FFN0=mode)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    for a, b in zip(*outs):
        sa = {tuple(m): s for m, s in zip(a["m"], a["s"])}
        sb = {tuple(m): s for m, s in zip(b["m"], b["s"])}
        assert len(sa) >= 10
        common = sorted(set(sa) & set(sb))
        dlog = max(abs(np.log(sa[k]) - np.log(sb[k])) for k in common)
        print("pairs %d / %d, common %d, max |dlog score| %.4f" % (len(sa), len(sb), len(common), dlog))
        # two engine variants differ like engine and oracle do: one fp16 rounding that falls the other way cascades
        # through nine layers (measured 0.02-0.11); the documented bound on the log assignment is 0.25
        assert dlog < 0.2
        for k in set(sa) ^ set(sb):
            sc = sa.get(k, sb.get(k))
            assert abs(np.log(sc) - np.log(0.1)) < 0.25, (k, sc)
