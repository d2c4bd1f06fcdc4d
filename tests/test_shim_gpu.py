"""The C++ facade (Estimator_net / MixVPR_net mirror) driven in keyframe.cpp's call order produces exactly what the
C-ABI path produces (same library, same kernels)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_matches_c_abi(weights_file, tmp_path):
    from d_vins_b200 import build, capi
    from oracle import synth
    exe = build.build_shim_demo()
    a, b = synth.make_pair(shift=(8, 16))
    vio = synth.vio_points(150, 480, 752, synth.BASE_SEED + 3)
    pa, pb, pv, po = (str(tmp_path / n) for n in ("a.raw", "b.raw", "vio.f32", "out.txt"))
    a.tofile(pa); b.tofile(pb); vio.astype(np.float32).tofile(pv)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(capi.LIB_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    subprocess.run([exe, weights_file, pa, pb, pv, po], check=True, env=env, timeout=300)
    lines = open(po).read().splitlines()
    eng = capi.Engine(height=480, width=752, weights_path=weights_file)
    frames = [l.split() for l in lines if l.startswith("frame")]
    kps = [[], []]
    cur = -1
    for l in lines:
        if l.startswith("frame"):
            cur += 1
        elif l.startswith("kp "):
            kps[cur].append(tuple(map(int, l.split()[1:])))
    feats = []
    for t, img in enumerate((a, b)):
        eng.frame_upload(img)
        re = eng.sp_describe(vio)
        r = eng.sp_detect()
        g = eng.mix_describe()
        assert int(frames[t][3]) == len(r["kpts"]) and int(frames[t][5]) == len(r["kpts"]) + 150
        assert kps[t] == [tuple(k) for k in r["kpts"]]
        assert abs(float(frames[t][7]) - g[0]) < 1e-5
        feats.append((np.concatenate([r["kpts"].astype(np.float32), vio]), np.concatenate([r["desc"], re]), re))
    m, s = eng.lg_match(vio, feats[0][0], feats[1][2], feats[0][1], 480, 752, 480, 752)
    got = [tuple(map(int, l.split()[1:3])) for l in lines if l.startswith("m ")]
    assert got == [tuple(p) for p in m]
    eng.close()
