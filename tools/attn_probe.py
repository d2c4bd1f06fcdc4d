"""LightGlue single-pair probe: python tools/attn_probe.py [M N]  (prints lg_match time and the match count)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from d_vins_b200 import capi
e = capi.Engine(height=480, width=752, weights_path=bench.make_weights())
rng = np.random.default_rng(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d0 = rng.standard_normal((M, 256)).astype(np.float32); d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
d1 = rng.standard_normal((N, 256)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
k0 = rng.uniform(8, 470, (M, 2)).astype(np.float32); k1 = rng.uniform(8, 470, (N, 2)).astype(np.float32)
for i in range(3):
    e.timer_start(); m, s = e.lg_match(k0, k1, d0, d1, 480, 752, 480, 752); print("lg_match ms", e.timer_stop(), len(m))
