#!/usr/bin/env python
"""bench.py - loop-closure keyframe throughput (SP + SP_RE + MixVPR + kNN + LightGlue) on 1..8 B200.

One "step" = `--rounds-per-step` keyframe rounds of the hot path; in every round each rank takes `--batch` synthetic
frames (default config: EuRoC-shaped 480x752 gray, 150 VIO points each), runs SuperPoint (512 kpts) + SP-recover
(shared encoder) + MixVPR, appends the global descriptors to the bank (world_size > 1: the path's single NCCL
all-gather), searches the bank (k=3, newest 50 excluded, 10 000 pre-filled rows + the stream so far) and runs one
LightGlue match per frame against the keyframe the kNN RETRIEVED (its features live in the device-resident store of
whichever rank extracted it - at N > 1 most of them are read from a peer GPU over NVLink).

  value : frames/s, whole job, frames resident in HBM when the timed region starts (CUDA events on the engine stream,
          max over ranks)
  e2e   : the same rounds through the C ABI with host buffers: pinned-host -> device frame upload and all result
          read-backs inside the timed region
  roofline : dominant kernel (event pair around every launch inside the timed region) + the dominant stage
  latency_b1 : ONE keyframe at a time through the reference-shaped per-keyframe calls (keyframe.cpp:74-81 order) with
          host inputs and host outputs - the number the reference's README publishes (12.5 ms on an RTX 2070S)
  cpu_baseline / --impl reference : the CPU oracle (PyTorch fp32 port of the reference's arithmetic) on the host cores
  parity_multi (N > 1): pre-flight on a small engine pair - banks identical on all ranks, remote match == local match

`--config` selects the BASELINE.json configuration: euroc_full (configs[3], default), kitti_50k (configs[4]),
sp_lg_512 (configs[1]), mix_knn_10k (configs[2]).

Contract: `python bench.py --gpus N --steps K --warmup W`; under torchrun one rank per GPU.  ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loop-closure frames/sec (SP+LG+MixVPR+kNN)"
UNIT = "frames/s"


def sp_gflop(h, w):
    """SuperPoint MACs from the layer shapes; 2 flop per MAC."""
    p1 = h * w; h2, w2 = h // 2, w // 2; h4, w4 = h2 // 2, w2 // 2; h8, w8 = h4 // 2, w4 // 2
    mac = p1 * 64 * 9 + p1 * 64 * 576 + 2 * (h2 * w2) * 64 * 576 + (h4 * w4) * 128 * 576 + (h4 * w4) * 128 * 1152 \
        + 2 * (h8 * w8) * 128 * 1152 + 2 * (h8 * w8) * 256 * 1152 + (h8 * w8) * 65 * 256 + (h8 * w8) * 256 * 256
    return 2.0 * mac / 1e9


def lg_gflop(m, n, d=256):
    """LightGlue(M,N) MACs: 9 layers of self(M) + self(N) + cross, + the assignment head."""
    def self_(t):
        return 10 * t * d * d + 2 * t * t * d
    cross = 9 * (m + n) * d * d + 3 * m * n * d
    return 2.0 * (9 * (self_(m) + self_(n) + cross) + (m + n) * d * d + m * n * d) / 1e9


MIX_GFLOP = 16.21   # ResNet50[:layer3] @320x320 + mixer, SURVEY §8(d)

CONFIGS = {
    # BASELINE.json configs[3]
    "euroc_full": dict(H=480, W=752, n_vio=150, max_vio=160, bank=10000, mode="full",
                       workload="full loop_fusion keyframe pipe, EuRoC-shaped synthetic stream: 480x752 gray, SuperPoint "
                                "512 kpts + SP_RE 150 pts (shared encoder) + MixVPR 320x320 + cosine kNN k=3 over 10k+ rows "
                                "+ LightGlue 150x662 every frame against the kNN-retrieved keyframe"),
    # BASELINE.json configs[4]
    "kitti_50k": dict(H=376, W=1241, n_vio=200, max_vio=208, bank=50000, mode="full",
                      workload="KITTI-odom-shaped stream: 376x1241 gray (score map 376x1240), SuperPoint 512 kpts + SP_RE "
                               "200 pts + MixVPR + cosine kNN k=3 over a 50k-keyframe bank + LightGlue 200x712 every frame"),
    # BASELINE.json configs[1]
    "sp_lg_512": dict(H=480, W=752, n_vio=0, max_vio=16, bank=0, mode="pair",
                      workload="SuperPoint + LightGlue pair match: two 480x752 frames (B = A translated), 512 keypoints "
                               "each, LightGlue 512x512; one unit = one PAIR (2 SuperPoint frames + 1 match)"),
    # BASELINE.json configs[2]
    "mix_knn_10k": dict(H=480, W=752, n_vio=0, max_vio=16, bank=10000, mode="global",
                        workload="MixVPR (ResNet50[:layer3] 320x320 from a 480x752 gray frame -> 512-d) + cosine kNN k=3 over "
                                 "a 10k-row bank"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("bf16_tflops"), d.get("hbm_gbs"),
                "measured (MEASURED_PEAKS.json; tensor = sustained cuBLAS bf16, kernel timed inside a long step)")
    return 1400.0, 1650.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_weights():
    from oracle import weights
    path = os.path.join(tempfile.gettempdir(), "dvins_synth_%d.dvw" % os.getuid())
    if not os.path.exists(path):
        tmp = path + ".%d" % os.getpid()
        weights.save_weights(tmp, weights.synth_all())
        os.replace(tmp, path)
    return path


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pipe(cfg, W_all, frames, vio, bank, steps, warmup):
    """The CPU oracle over `steps` units of the config's workload (one unit per step).  -> (units/s, cores, seconds)."""
    import torch
    from oracle import weights, superpoint as osp, mixvpr as omix, lightglue as olg, knn
    ws, wl, wm = weights.sub(W_all, "sp."), weights.sub(W_all, "lg."), weights.sub(W_all, "mix.")
    # PyTorch CPU convs at these sizes stop scaling past ~32 threads (measured on the 128-thread B200 host: 0.37 s/frame
    # SuperPoint at 16-32 threads, 0.61 s at 64, worse at 128), so the baseline uses the fastest setting it can.
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    H, W, mode = cfg["H"], cfg["W"], cfg["mode"]
    prev = None
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        img = frames[i % len(frames)]
        if mode == "global":
            g = omix.mixvpr(wm, img)
            knn.knn_reference_style(bank, g, len(bank))
            continue
        if mode == "pair":
            ra = osp.superpoint(ws, img)
            rb = osp.superpoint(ws, frames[(i + 1) % len(frames)])
            olg.lightglue(wl, ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], H, W, H, W)
            continue
        r = osp.superpoint(ws, img)
        dre = osp.superpoint_recover(ws, img, vio, feat=r["feat"])
        g = omix.mixvpr(wm, img)
        knn.knn_reference_style(bank, g, len(bank))
        kp_all = np.concatenate([r["kpts"].astype(np.float32), vio]); de_all = np.concatenate([r["desc"], dre])
        if prev is not None:
            olg.lightglue(wl, vio, prev[0], dre, prev[1], H, W, H, W)
        else:
            olg.lightglue(wl, vio, kp_all, dre, de_all, H, W, H, W)
        prev = (kp_all, de_all)
    dt = time.perf_counter() - t0
    return steps / dt, cores, dt


def cpu_inputs(cfg, nframes=4):
    from oracle import synth
    H, W = cfg["H"], cfg["W"]
    st = synth.Stream(H, W, period=40, margin=96)
    frames = [st.frame(t) for t in range(nframes)]
    vio = synth.vio_points(max(cfg["n_vio"], 1), H, W, synth.BASE_SEED + 3)[:cfg["n_vio"]] if cfg["n_vio"] else \
        np.zeros((0, 2), np.float32)
    bank, _ = synth.make_bank(max(cfg["bank"], 3), seed=synth.BASE_SEED + 9)
    return frames, vio, bank


def run_reference(args, cfg, rank):
    """--impl reference: the reference's own CPU arithmetic for this path.  The TensorRT/ROS C++ path cannot be built
    here; the arm therefore times the oracle port (kind "port"), one unit of the workload per step.
    Under torchrun rank 0 alone runs; the other ranks exit without work."""
    if rank != 0:
        return
    from oracle import weights
    W_all = weights.load_weights(make_weights())
    frames, vio, bank = cpu_inputs(cfg)
    steps = max(1, args.steps)
    warm = min(max(args.warmup, 1), 2)
    fps, cores, dt = cpu_pipe(cfg, W_all, frames, vio, bank, steps, warm)
    unit = "pairs/s" if cfg["mode"] == "pair" else UNIT
    sample = "%d units (1 per step) of the same workload, PyTorch fp32 CPU oracle, %d threads" % (steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": unit, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "units_per_step": 1},
            "cpu_baseline": {"value": fps, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on stdout.  Everything
    else is sent to stderr: fd 1 is pointed at fd 2 for the run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------ multi-GPU pre-flight
def multi_gpu_preflight(dist, rank, world, local, wpath):
    """Multi-GPU data-path parity on a small engine (2 frames per rank per round; every rank takes part):
      banks_equal     - after two rounds every rank's bank holds all ranks' descriptors, bit-identical, in global order
      remote_eq_local - a LightGlue match whose OLD keyframe lives in a peer's feature store (read over NVLink through
                        the CUDA-IPC mapping) equals the match against a locally extracted copy of the same keyframe
    Returns a dict (same on every rank)."""
    from d_vins_b200 import capi, sharding
    from oracle import synth
    H, W, b = 480, 752, 2
    eng = capi.Engine(device=local, height=H, width=W, max_batch=b, max_vio=160, weights_path=wpath, bank_capacity=256,
                      store_capacity=64, world_size=world, rank=rank)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0])
    st = synth.Stream(H, W, period=12, margin=64)

    def extract(ids, frame_of):
        frames = np.stack([st.frame(frame_of(int(t))) for t in ids])
        vio = np.zeros((b, 160, 2), np.float32)
        for i, t in enumerate(ids):
            vio[i, :150] = synth.vio_points(150, H, W, 500 + frame_of(int(t)))
        eng.batch_upload(frames)
        eng.batch_extract(vio, np.full((b,), 150, np.int32), ids)
        return eng.batch_commit(b)

    ok_rows = True
    globals_seen = []
    for R in range(2):
        ids = sharding.round_frame_ids(R, rank, world, b)
        first = extract(ids, lambda t: t)
        ok_rows &= first == int(ids[0])
        globals_seen.extend(eng.batch_read_global(i) for i in range(b))
    bank = eng.bank_export()
    ok_rows &= bank.shape[0] == 2 * world * b
    mine = [int(t) for R in range(2) for t in sharding.round_frame_ids(R, rank, world, b)]
    for g, t in zip(globals_seen, mine):
        ok_rows &= bool(np.array_equal(bank[t], g))
    gathered = [None] * world
    dist.all_gather_object(gathered, bank.tobytes())
    banks_equal = bool(ok_rows) and all(g == gathered[0] for g in gathered)
    # remote old keyframe: this rank's first frame of round 1 against the NEXT rank's first frame of round 0
    peer = (rank + 1) % world
    q = int(sharding.round_frame_ids(1, rank, world, b)[0])
    old_remote = int(sharding.round_frame_ids(0, peer, world, b)[0])
    owner, n_tot, _ = eng.store_lookup(old_remote)
    m_remote, s_remote = eng.batch_match(np.array([q], np.int64), np.array([old_remote], np.int64))[0]
    st_remote = int(eng.last_match_status[0])
    # local copy of the same keyframe under fresh ids (every rank takes part in the round's all-gather)
    copy_ids = np.array([1000 + rank * b, 1001 + rank * b], np.int64)
    extract(copy_ids, lambda t: old_remote if t == copy_ids[0] else old_remote + 1)
    m_local, s_local = eng.batch_match(np.array([q], np.int64), copy_ids[:1])[0]
    same = (owner == peer and st_remote >= 0 and len(m_remote) > 0 and np.array_equal(m_remote, m_local)
            and np.array_equal(s_remote, s_local))
    # a keyframe nobody holds is reported per pair, not as a failed batch
    res = eng.batch_match(np.array([q, q], np.int64), np.array([987654, copy_ids[0]], np.int64))
    per_pair = int(eng.last_match_status[0]) == -1 and np.array_equal(res[1][0], m_local)
    eng.close()
    flags = [None] * world
    dist.all_gather_object(flags, (bool(same), bool(per_pair), int(len(m_remote))))
    return {"remote_eq_local": all(f[0] for f in flags), "banks_equal": banks_equal,
            "missing_keyframe_is_per_pair": all(f[1] for f in flags), "pairs_matched_remote": [f[2] for f in flags],
            "frames_per_rank_per_round": b, "rounds": 3}


# ------------------------------------------------------------------------------------------------ B=1 latency
def latency_b1(cfg, wpath, local, frames, vio_pts, bank, n=200, warm=20):
    """ONE keyframe at a time through the reference-shaped per-keyframe entry points in keyframe.cpp:74-81 order
    (computeWindowSuperpoint -> computeSuperpoint -> compute_mix_des_test -> sort_vec_faiss) followed by the
    light_glue_matcher call of findConnection (:583-632), HOST buffers in and out on every call (what the reference's
    facade returns: 512x256 f32 descriptors etc.).  Wall clock per keyframe, p50 over n keyframes."""
    from d_vins_b200 import capi
    H, W = cfg["H"], cfg["W"]
    eng = capi.Engine(device=local, height=H, width=W, max_batch=1, max_vio=cfg["max_vio"], weights_path=wpath,
                      bank_capacity=bank.shape[0] + n + warm + 8, store_capacity=4)
    eng.bank_import(bank)
    names = ["upload", "sp_re", "sp", "mixvpr", "knn", "lightglue", "total"]
    T = {k: [] for k in names}
    prev = None
    for i in range(warm + n):
        img = frames[i % len(frames)]
        t = [time.perf_counter()]
        eng.frame_upload(img); t.append(time.perf_counter())
        dre = eng.sp_describe(vio_pts); t.append(time.perf_counter())
        r = eng.sp_detect(); t.append(time.perf_counter())
        g = eng.mix_describe(); t.append(time.perf_counter())
        row = eng.bank_append(g)
        eng.bank_search(g, row - 49 if row >= 50 else row + 1); t.append(time.perf_counter())
        kp_all = np.concatenate([r["kpts"].astype(np.float32), vio_pts]); de_all = np.concatenate([r["desc"], dre])
        old = prev if prev is not None else (kp_all, de_all)
        eng.lg_match(vio_pts, old[0], dre, old[1], H, W, H, W); t.append(time.perf_counter())
        prev = (kp_all, de_all)
        if i >= warm:
            for k, a, b_ in zip(names[:-1], t[:-1], t[1:]):
                T[k].append((b_ - a) * 1e3)
            T["total"].append((t[-1] - t[0]) * 1e3)
    eng.close()
    out = {"p50_ms": {k: float(np.median(v)) for k, v in T.items()},
           "p90_total_ms": float(np.percentile(T["total"], 90)), "keyframes": n,
           "keyframes_per_s": 1000.0 / float(np.median(T["total"])),
           "path": "dv_frame_upload, dv_sp_describe, dv_sp_detect, dv_mix_describe, dv_bank_append + dv_bank_search, "
                   "dv_lg_match - host buffers in and out, one keyframe per call (keyframe.cpp:74-81, :583-632)",
           "reference_readme_total_ms_rtx2070s_752x480": 12.5}
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="frames per rank per round (one batched engine call)")
    ap.add_argument("--rounds-per-step", type=int, default=4, help="keyframe rounds per step")
    ap.add_argument("--config", default="euroc_full", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-preflight", action="store_true")
    ap.add_argument("--async-match", action="store_true",
                    help="dv_batch_match_begin / _end: collect a round's matches after the NEXT round's extraction was queued "
                         "(measured equal to the synchronous call on one GPU: 5584 vs 5585 frames/s)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from d_vins_b200 import capi, sharding
    from oracle import synth              # synthetic inputs only (oracle/synth.py); the product never calls the oracle

    H, W, N_VIO, MAXV, BANK, mode = cfg["H"], cfg["W"], cfg["n_vio"], cfg["max_vio"], cfg["bank"], cfg["mode"]
    b, rps = args.batch, max(1, args.rounds_per_step)
    warm_steps = max(args.warmup, 3)
    wpath = make_weights() if rank == 0 else None
    if world > 1:
        dist.barrier()
        wpath = make_weights()

    parity_multi = None
    if world > 1 and not args.no_preflight:
        parity_multi = multi_gpu_preflight(dist, rank, world, local, wpath)

    total_rounds = (warm_steps + 3 * args.steps + 2 + 3) * rps + 4
    units_per_round = b // 2 if mode == "pair" else b
    eng = capi.Engine(device=local, height=H, width=W, max_batch=b, max_vio=MAXV, weights_path=wpath,
                      bank_capacity=BANK + total_rounds * b * world + 64,
                      store_capacity=min(total_rounds * b, 4096), world_size=world, rank=rank)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0])

    # ---------------- synthetic inputs (seeded): `nblocks` distinct blocks of b frames in pinned host memory.  Round R
    # of rank r shows content block (R * world + r) % nblocks; nblocks is NOT a multiple of world, so the keyframe that
    # saw the same place before - the one the kNN retrieves - was extracted by ANOTHER rank (remote LightGlue operand).
    npool = 3 if world <= 2 else 2
    nblocks = npool * world - 1 if world > 1 else npool
    st = synth.Stream(H, W, period=600, margin=400)
    pool = []
    for k in range(nblocks):
        t = torch.empty((b, H, W), dtype=torch.uint8).pin_memory()
        for i in range(b):
            if mode == "pair":      # frames 2j, 2j+1 of a block = a scene and the same scene translated by (6, 9) px
                y0, x0 = st.offset((k * b + i) // 2 * 7)
                t[i] = torch.from_numpy(synth.frame_from_canvas(st.canvas, min(y0 + 6 * (i & 1), 400), min(x0 + 9 * (i & 1), 400),
                                                                H, W, synth.BASE_SEED + 5000 + k * b + i))
            else:
                t[i] = torch.from_numpy(st.frame(k * b + i))
        pool.append(t)
    vio = np.zeros((b, MAXV, 2), np.float32)
    nv = np.full((b,), N_VIO, np.int32)
    if N_VIO:
        for i in range(b):
            vio[i, :N_VIO] = synth.vio_points(N_VIO, H, W, synth.BASE_SEED + 3 + i)
    bank = None
    if BANK:
        bank, _ = synth.make_bank(BANK, seed=synth.BASE_SEED + 9)
        eng.bank_import(bank)

    round_no = [0]
    h2d = [0]
    d2h = [0]
    prefetched = [False]
    partner = {"knn": 0, "fallback": 0, "remote": 0, "not_resident": 0}

    def block_of(R):
        return pool[(R * world + rank) % nblocks]

    pending = [False]

    def collect_match():
        """Results of the LightGlue pass queued by the previous round (dv_batch_match_begin / _end)."""
        if not pending[0]:
            return None
        res = eng.batch_match_end()
        pending[0] = False
        partner["not_resident"] += int((eng.last_match_status < 0).sum())
        d2h[0] += sum(m.nbytes + s.nbytes for m, s in res) + 4 * b
        return res

    def one_round(upload: bool):
        """upload=True: the round's frames come from pinned host memory.  Uploads are asynchronous and double-buffered
        in the engine, so round R+1's frames are queued right after round R's extraction and travel while round R is
        being searched and matched; every round still costs exactly one upload inside the timed region."""
        R = round_no[0]
        ids = sharding.round_frame_ids(R, rank, world, b)                  # contiguous block per rank per round
        if mode == "global":
            collect_match()
        if upload and not prefetched[0]:
            eng.batch_upload_ptr(b, block_of(R).data_ptr(), H * W, W)
            h2d[0] += b * H * W
        if mode == "global":
            eng.batch_describe_global(b)
        else:
            eng.batch_extract(vio, nv, ids)
            h2d[0] += vio.nbytes + nv.nbytes
            d2h[0] += 4 * b
            collect_match()                # the previous round's LightGlue results (complete: extraction synchronised)
        if upload:
            eng.batch_upload_ptr(b, block_of(R + 1).data_ptr(), H * W, W)   # next round's frames
            h2d[0] += b * H * W
            prefetched[0] = True
        res = None
        if mode in ("full", "global"):
            eng.batch_commit(b)
            D, I = eng.batch_search(None, b)           # keyframe.cpp:274-282 window, applied by the engine
            d2h[0] += D.nbytes + I.nbytes
        if mode == "full":
            # the reference matches against the keyframe detectLoop returns from the kNN result (pose_graph.cpp:451-509);
            # here every frame is matched (worst case) against its top-1 retrieved keyframe when that keyframe's local
            # features are resident on some rank, else against this rank's previous-round keyframe
            old = sharding.previous_round_ids(ids, world, b) if R > 0 else ids.copy()
            cand = I[:, 0] - BANK
            owner = eng.store_lookup_many(np.maximum(cand, -1))
            use = (cand >= 0) & (owner >= 0)
            old[use] = cand[use]
            partner["knn"] += int(use.sum())
            partner["remote"] += int((use & (owner != rank)).sum())
            partner["fallback"] += int((~use).sum())
            if not args.async_match:
                res = eng.batch_match(ids, old)
                partner["not_resident"] += int((eng.last_match_status < 0).sum())
                d2h[0] += sum(m.nbytes + s.nbytes for m, s in res) + 4 * b
            else:
                # queued, not awaited: the next round's upload + extraction are enqueued right behind the match and the
                # results are collected after that extraction (collect_match below) - the GPU never idles between rounds
                eng.batch_match_begin(ids, old)
                pending[0] = True
        elif mode == "pair":
            res = eng.batch_match_sp(ids[0::2], ids[1::2])
            d2h[0] += sum(m.nbytes + s.nbytes for m, s in res) + 4 * (b // 2)
        round_no[0] += 1
        return res

    def one_step(upload: bool):
        for _ in range(rps):
            one_round(upload)

    def barrier():
        collect_match()
        eng.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- warm-up (also fills the feature store / bank so the timed kNN retrieves real keyframes)
    eng.batch_upload_ptr(b, block_of(0).data_ptr(), H * W, W)
    for _ in range(warm_steps):
        one_step(upload=False)

    # ---------------- (1) device-resident: frames already in HBM
    probe_id = 1 if mode == "global" else 0           # 0: conv1a+conv1b implicit GEMM, 1: kNN bank scan
    clocks = ClockSampler(local)
    for k in partner:
        partner[k] = 0
    eng.stats_reset()
    eng.probe_select(probe_id)
    eng.probe_enable(True)
    eng.probe_read(reset=True)
    barrier()
    if rank == 0:
        clocks.start()
    eng.timer_start()
    for _ in range(args.steps):
        one_step(upload=False)
    collect_match()
    ms_dev = eng.timer_stop()
    barrier()
    probe_ms, probe_n = eng.probe_read(reset=True)
    eng.probe_enable(False)
    _, launches = eng.stats_read()
    ms_dev = max_over_ranks(ms_dev)
    partner_dev = dict(partner)
    bank_rows_mid = eng.bank_size()

    # ---------------- (2) end to end through the C ABI: pinned host frames in, results out, every step
    one_step(upload=True)
    h2d[0] = d2h[0] = 0
    for k in partner:
        partner[k] = 0
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        one_step(upload=True)
    collect_match()
    ms_e2e = eng.timer_stop()
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    partner_e2e = dict(partner)          # new content every round: at N > 1 the kNN-retrieved keyframe lives on a peer
    h2d_step, d2h_step = h2d[0] // args.steps, d2h[0] // args.steps
    # ---------------- (3) the device-resident loop once more: separates the cost of the uploads from clock / power
    # drift between the two timed regions (the e2e loop runs on a GPU that has been at full load for longer)
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        one_step(upload=False)
    collect_match()
    ms_dev2 = max_over_ranks(eng.timer_stop())
    barrier()
    clk = clocks.stop() if rank == 0 else None

    # ---------------- per-stage shares (separate pass; event records perturb the timed loops)
    eng.stats_enable(True)
    eng.stats_reset()
    for _ in range(3):
        one_round(upload=False)
    collect_match()
    stage_ms, _ = eng.stats_read()
    eng.stats_enable(False)
    stage_ms = {k: v / 3.0 for k, v in stage_ms.items()}

    # ---------------- p50 LightGlue latency (B=1, device-resident features)
    p50 = None
    if mode in ("full", "pair"):
        lat = []
        ids_last = sharding.round_frame_ids(round_no[0] - 1, rank, world, b)
        q1, o1 = ids_last[:1], ids_last[1:2]
        for i in range(60):
            eng.timer_start()
            if mode == "full":
                eng.batch_match(q1, o1)
            else:
                eng.batch_match_sp(q1, o1)
            t = eng.timer_stop()
            if i >= 10:
                lat.append(t)
        p50 = float(np.median(lat))

    units_total = args.steps * rps * units_per_round * world
    value = units_total / (ms_dev / 1000.0)
    e2e = units_total / (ms_e2e / 1000.0)
    if rank != 0:
        eng.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    tf_peak, tf_burst, hbm_peak, peak_src = _peaks()
    spg, lgg = sp_gflop(H, W), (lg_gflop(N_VIO, 512 + N_VIO) if mode == "full" else lg_gflop(512, 512))
    unit_gflop = {"full": spg + MIX_GFLOP + lgg, "pair": 2 * spg + lgg, "global": MIX_GFLOP}[mode]
    k_ms = probe_ms / max(probe_n, 1)
    if probe_id == 0:
        flop = 2.0 * H * W * 64 * (576 + 9) * b
        achieved = flop / (k_ms * 1e-3) / 1e12 if probe_n else None
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        # (r02, 480x752, 64 frames/launch: 23.3 MB read = the u8 frames, 686.6 MB written = the pooled fp16 output; the
        # algorithmic bytes are 23.1 + 739.2 MB, part of the output was still in L2 at kernel end), scaled by batch / pixels
        traffic = (23.271e6 + 686.55e6) * b / 64.0 * (H * W) / (480.0 * 752.0)
        roofline = {"kernel": "conv3x3_halo64_kernel<2> (SuperPoint conv1a 1->64 on the tensor cores inside conv1b 3x3 "
                              "64->64 + ReLU + 2x2 max-pool, halo-tile implicit GEMM, %d frames/launch)" % b,
                    "bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": (achieved / tf_peak) if achieved else None, "traffic": traffic,
                    "traffic_source": "ncu --set full, profiles/r02_kernels_ncu_selected.txt (conv3x3_halo64_kernel<2>, 64 "
                                      "frames; scaled by batch/64 and pixel count)",
                    "peak_source": peak_src, "frac_of_burst_peak": (achieved / tf_burst) if achieved else None,
                    "avg_launch_ms": k_ms, "launches_timed": probe_n, "flop_per_launch": flop}
    else:
        # kNN scan: one pass over the searched bank prefix per 8 queries; algorithmic bytes = rows * 512 * 4 per pass
        rows = 0.5 * (BANK + bank_rows_mid)
        byts = rows * 2048.0
        achieved = byts / (k_ms * 1e-3) / 1e9 if probe_n else None
        roofline = {"kernel": "k_knn_scan<3> (cosine kNN: one coalesced pass over the f32 bank per 8 queries, warp-shuffle "
                              "dot products, per-block top-3)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None, "traffic": None,
                    "peak_source": peak_src, "avg_launch_ms": k_ms, "launches_timed": probe_n,
                    "bytes_per_launch": byts,
                    "note": "launches of %d-row scans are latency-bound (a 20 MB bank is L2-resident after the first "
                            "pass); see profiles/ for the marginal rate at 200k rows" % int(rows)}
    roofline["whole_unit_tflops"] = unit_gflop * value / 1e3
    # the north-star's bar is stated on the dominant STAGE: all of its kernels together, from the per-stage event pass
    stages = {}
    for nm, key, gf in (("superpoint_convs", "sp_convs", spg), ("mixvpr", "mixvpr", MIX_GFLOP),
                        ("lightglue", "lightglue", lgg)):
        ms = stage_ms.get(key, 0.0)
        if ms > 0:
            n_units = b if key != "lightglue" else (b if mode == "full" else b // 2)
            tf = gf * 1e9 * n_units / (ms * 1e-3) / 1e12
            stages[nm] = {"ms_per_round": ms, "gflop_per_unit": gf, "achieved_tflops": tf, "frac": tf / tf_peak,
                          "frac_of_burst_peak": tf / tf_burst}
    dom = "mixvpr" if mode == "global" else "superpoint_convs"
    if dom in stages:
        roofline["dominant_stage"] = dict(name=dom, **stages[dom])
    roofline["stages"] = stages
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import weights
        W_all = weights.load_weights(wpath)
        frames_c = [pool[0][i].numpy() for i in range(min(b, 4))]
        nsample = {"full": 24, "pair": 16, "global": 120}[mode]
        fps, cores, dt = cpu_pipe(cfg, W_all, frames_c, vio[0, :N_VIO], bank if bank is not None else np.zeros((3, 512), np.float32),
                                  nsample, 1)
        cpu = {"value": fps, "unit": "pairs/s" if mode == "pair" else UNIT, "cores": cores, "kind": "port",
               "sample": "%d units of the same workload through the PyTorch fp32 CPU oracle (%.1f s)" % (nsample, dt)}
    lat_b1 = None
    if not args.no_latency and world == 1 and mode == "full":
        eng.sync()
        lat_b1 = latency_b1(cfg, wpath, local, [pool[0][i].numpy() for i in range(min(b, 8))], vio[0, :N_VIO],
                            bank if bank is not None else np.zeros((0, 512), np.float32))
    line = {
        "metric": METRIC if mode == "full" else METRIC + " [config %s]" % args.config, "value": value,
        "unit": "pairs/s" if mode == "pair" else UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm_steps,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config, "frames_per_rank_per_round": b, "rounds_per_step": rps,
                   "units_per_rank_per_step": units_per_round * rps, "global_units_per_step": units_per_round * rps * world,
                   "units_in_timed_region_per_rank": args.steps * rps * units_per_round,
                   "parallelism": "frame-sharded x%d, one NCCL all-gather of [b,516] per round" % world,
                   "l2": "per-round working set (>= %.1f GB of activations) far exceeds the 126 MB L2; no explicit flush" % (0.16 * b),
                   "weights": "seeded synthetic (oracle/weights.py)",
                   "lg_partner": "top-1 of the kNN result when resident on any rank, else previous round (counts below)",
                   "e2e_input": "pinned host frames, asynchronous double-buffered upload queued one round ahead (one "
                                "upload per round inside the timed region)",
                   "mixvpr_aggregator_parity": "unpinned (amaralibey/MixVPR not in the image; structure-checked only)"},
        "e2e": {"value": e2e, "unit": "pairs/s" if mode == "pair" else UNIT, "h2d_bytes_per_step": int(h2d_step),
                "d2h_bytes_per_step": int(d2h_step), "ms_per_step": ms_e2e / args.steps,
                "device_resident_rerun_after": units_total / (ms_dev2 / 1000.0)},
        "gpu_launches": int(launches),
        "p50_match_ms": p50,
        "lg_partner_counts_rank0": partner_dev,
        "lg_partner_counts_e2e_rank0": partner_e2e,
        "bank_rows_end": eng.bank_size(),
        "stage_ms_per_round": stage_ms,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "latency_b1": lat_b1,
        "parity_multi": parity_multi,
        "clocks": clk,
    }
    if mode == "full":
        # the reference only runs LightGlue when detectLoop fires; estimate for "LightGlue on 10 % of the
        # frames" from the per-stage event pass (device-resident round time minus 90 % of its LightGlue share)
        line["lg_on_10pct_frames_estimate"] = (units_per_round * world) / max(
            1e-9, (ms_dev / (args.steps * rps) - 0.9 * stage_ms.get("lightglue", 0.0)) / 1000.0)
    emit(line)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
