"""MixVPR global descriptor: CUDA path vs the CPU oracle (pre-processing exact, descriptor within fp16 tolerance)."""
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


def test_preprocess_matches_reference_kernel_semantics(eng):
    from oracle import mixvpr as omix, synth
    img = synth.make_frame(480, 752, synth.BASE_SEED)
    eng.frame_upload(img)
    eng.mix_describe()
    got = eng.dbg_read("mix_img").reshape(320, 320, 3)
    ref = omix.preprocess_mix(img).transpose(1, 2, 0)
    # the device stores the pre-processed image as fp16: identical after the same rounding
    assert np.array_equal(got, ref.astype(np.float16).astype(np.float32))


def test_mixvpr_descriptor(eng, all_weights):
    from oracle import mixvpr as omix, synth, weights
    wm = weights.sub(all_weights, "mix.")
    for seed in (synth.BASE_SEED, synth.BASE_SEED + 17):
        img = synth.make_frame(480, 752, seed)
        keep = {}
        ref = omix.mixvpr(wm, img, keep)
        eng.frame_upload(img)
        got = eng.mix_describe()
        feat = eng.dbg_read("mix_feat").reshape(400, 1024)
        rf = parity.nhwc(keep["layer3.5"]).reshape(400, 1024)
        assert np.abs(feat - rf).max() < 0.03 * np.abs(rf).max(), np.abs(feat - rf).max()
        assert abs(np.linalg.norm(got) - 1.0) < 1e-5
        cos = float(ref @ got)
        assert cos > parity.GLOBAL_COS_MIN, cos
        assert np.abs(ref - got).max() < 4e-3, np.abs(ref - got).max()


def test_mixvpr_bgr_input(eng, all_weights):
    """3-channel frames: the BGR->RGB swap + BGR-ordered mean/std quirk (deep_net.cpp:1298-1300) is replicated."""
    from oracle import mixvpr as omix, weights
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (480, 752, 3), dtype=np.uint8)
    eng.frame_upload(img)
    got = eng.mix_describe()
    pre = eng.dbg_read("mix_img").reshape(320, 320, 3)
    ref_pre = omix.preprocess_mix(img).transpose(1, 2, 0)
    assert np.array_equal(pre, ref_pre.astype(np.float16).astype(np.float32))
    ref = omix.mixvpr(weights.sub(all_weights, "mix."), img)
    assert float(ref @ got) > parity.GLOBAL_COS_MIN


def test_reordered_tail_matches_gemm_tail():
    """The aggregator tail with row_proj applied first (k_rowproj_x + k_chanproj_norm, fp32 throughout; the default)
    against the transpose + channel_proj GEMM + row_proj kernels (DV_MIX_TAIL=0) on the same frames: the two are the same
    linear map evaluated in a different order, so the unit descriptors agree to fp16-operand precision."""
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
from oracle import synth
e = capi.Engine(height=480, width=752, weights_path=bench.make_weights())
out = []
for seed in (synth.BASE_SEED, synth.BASE_SEED + 5):
    e.frame_upload(synth.make_frame(480, 752, seed))
    out.append(e.mix_describe().tolist())
print(json.dumps(out))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, DV_MIX_TAIL=mode),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.asarray(json.loads(r.stdout.strip().splitlines()[-1]), np.float32))
    a, b = outs
    assert a.shape == b.shape == (2, 512)
    for x, y in zip(a, b):
        assert float(x @ y) > 0.99999, float(x @ y)
        assert np.abs(x - y).max() < 1e-3, np.abs(x - y).max()
