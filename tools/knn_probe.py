"""kNN bank-scan probe (BASELINE config 3 / 5 shapes): dv_bank_search over 10k / 50k rows, one query per pass, timed with
the engine's per-stage CUDA events.  python tools/knn_probe.py -> us per search and GB/s of bank bytes (rows * 2048 B)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d_vins_b200 import capi
from oracle import synth

for rows in (10000, 50000, 200000):
    bank, q = synth.make_bank(rows, seed=synth.BASE_SEED + 9)
    e = capi.Engine(height=64, width=64, bank_capacity=rows + 64)
    e.bank_import(bank)
    for _ in range(3):
        e.bank_search(q, rows)
    e.stats_enable(True); e.stats_reset()
    reps = 50
    for i in range(reps):
        e.bank_search(bank[(7 * i) % rows], rows)
    st, _ = e.stats_read()
    us = st["knn"] * 1e3 / reps
    print("rows %6d: %8.1f us per search (scan + merge + result copy) -> %7.1f GB/s of bank bytes" % (
        rows, us, rows * 2048.0 / (us * 1e-6) / 1e9))
    e.close()
