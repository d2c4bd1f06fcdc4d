import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d_vins_b200 import capi
e = capi.Engine(height=64, width=64)
rng = np.random.default_rng(0)
for (M, N, K) in [(32768, 512, 512), (32768, 256, 256)]:
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32) / 16
    bias = rng.standard_normal(N).astype(np.float32)
    for _ in range(2):
        e.dbg_gemm(A, B, bias)
