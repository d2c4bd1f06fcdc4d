"""Small ncu / compute-sanitizer target: a few keyframe rounds (extract -> commit -> search -> match) on synthetic
EuRoC-shaped frames.  Used as:  ncu --set full -k regex:... python tools/profile_target.py [rounds] [batch]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from d_vins_b200 import capi  # noqa: E402
from oracle import synth  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (480, 752)
eng = capi.Engine(height=H, width=W, max_batch=b, max_vio=160, weights_path=bench.make_weights(),
                  bank_capacity=20000, store_capacity=(rounds + 1) * b)
st = synth.Stream(H, W, period=600, margin=400)
frames = np.stack([st.frame(i) for i in range(b)])
vio = np.zeros((b, 160, 2), np.float32)
nv = np.full((b,), 150, np.int32)
for i in range(b):
    vio[i, :150] = synth.vio_points(150, H, W, synth.BASE_SEED + 3 + i, min_dist=30 if H >= 400 else 8)
bank, _ = synth.make_bank(10000, seed=synth.BASE_SEED + 9)
eng.bank_import(bank)
eng.batch_upload(frames)
for R in range(rounds):
    ids = np.arange(b, dtype=np.int64) + R * b
    eng.batch_extract(vio, nv, ids)
    eng.batch_commit(b)
    eng.batch_search(None, b)
    eng.batch_match(ids, ids - b if R > 0 else ids)
# the per-keyframe (B = 1) calls as well: graphs / fused kNN / host-vector LightGlue
eng.frame_upload(frames[0])
r = eng.sp_detect(); d = eng.sp_describe(vio[0, :150]); g = eng.mix_describe()
eng.bank_search(g, 10000)
eng.lg_match(vio[0, :150], r["kpts"], d, r["desc"], H, W, H, W)
eng.sync()
print("done", rounds, b)
