"""GEMM shape probe: DV_GEMM_TIME=1 [DV_GEMM_WRES=0|1] python tools/gemm_probe.py  (per-launch event timing on stderr)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d_vins_b200 import capi
e = capi.Engine(height=64, width=64)
rng = np.random.default_rng(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 26112
for (N, K) in [(768, 256), (512, 256), (256, 256), (512, 512), (256, 512), (1024, 256), (256, 1024)]:
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32) / 16
    bias = rng.standard_normal(N).astype(np.float32)
    for _ in range(3):
        e.dbg_gemm_ex(A, B, bias, want32=False, want16=True)
