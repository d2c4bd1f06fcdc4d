"""GPU diagnostic (not a test): how close are the engine's DISCRETE outputs to the fp32 oracle and to the
quantisation-aware oracle on the three BASELINE frames?  Writes gpurun_out/parity_diag.json; the numbers pin the
thresholds asserted in tests/test_parity_exact_gpu.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d_vins_b200 import capi                                      # noqa: E402
from oracle import lightglue as olg, quant, superpoint as osp, synth, weights   # noqa: E402
import bench                                                      # noqa: E402


def kp_stats(o, r):
    so = [tuple(k) for k in o["kpts"]]; sr = [tuple(k) for k in r["kpts"]]
    common = len(set(so) & set(sr))
    same_pos = sum(a == b for a, b in zip(so, sr))
    return {"n_oracle": len(so), "n_engine": len(sr), "set_common": common, "set_equal": set(so) == set(sr),
            "order_equal": so == sr, "same_position": same_pos,
            "score_map_maxabs": float(np.abs(o["score_map"] - r["score_map_dbg"]).max())}


def pair_stats(mo, mg):
    a = {tuple(p) for p in mo}; b = {tuple(p) for p in mg}
    return {"n_oracle": len(a), "n_engine": len(b), "common": len(a & b), "equal": bool(np.array_equal(mo, mg))}


def main():
    wpath = bench.make_weights()
    W = weights.load_weights(wpath)
    ws, wl = weights.sub(W, "sp."), weights.sub(W, "lg.")
    out = {}
    frames = {"euroc": (480, 752, synth.make_frame(480, 752, synth.BASE_SEED)),
              "kitti": (376, 1241, synth.make_frame(376, 1241, synth.BASE_SEED + 5))}
    a, b = synth.make_pair(shift=(8, 16))
    frames["pair_a"] = (480, 752, a); frames["pair_b"] = (480, 752, b)
    for s in range(6):     # a few more EuRoC-shaped frames for a rate
        frames["euroc_s%d" % s] = (480, 752, synth.make_frame(480, 752, synth.BASE_SEED + 100 + s))
    engs = {}
    feats = {}
    for name, (H, Wd, img) in frames.items():
        if (H, Wd) not in engs:
            engs[(H, Wd)] = capi.Engine(height=H, width=Wd, weights_path=wpath)
        e = engs[(H, Wd)]
        e.frame_upload(img)
        r = e.sp_detect()
        r["score_map_dbg"] = e.dbg_read("score_map").reshape(H // 8 * 8, Wd // 8 * 8)
        o32 = osp.superpoint(ws, img)
        keepq = {}
        oq = quant.superpoint_q(ws, img, keep=keepq)
        out[name] = {"vs_fp32": kp_stats(o32, r), "vs_quant": kp_stats(oq, r)}
        if name in ("euroc", "kitti"):
            # per-layer fidelity of the quantisation-aware oracle: fraction of fp16 activations that are bit-identical
            layers = {}
            for nm in ("conv1b_pool", "conv2a", "conv2b_pool", "conv3a", "conv3b_pool", "conv4a", "conv4b"):
                ref = keepq[nm][0].permute(1, 2, 0).numpy()
                hh, ww, cc = ref.shape
                got = e.dbg_read(nm + "_blocked").reshape(cc // 8, hh, ww, 8).transpose(1, 2, 0, 3).reshape(hh, ww, cc)
                d = np.abs(got - ref)
                layers[nm] = {"equal_frac": float((d == 0).mean()), "maxabs": float(d.max()), "max_ref": float(np.abs(ref).max())}
            pd = e.dbg_read("convPD").reshape(H // 8, Wd // 8, 512)
            refp = np.concatenate([keepq["convPa"][0].permute(1, 2, 0).numpy(), keepq["convDa"][0].permute(1, 2, 0).numpy()], 2)
            d = np.abs(pd - refp)
            layers["convPa|Da"] = {"equal_frac": float((d == 0).mean()), "maxabs": float(d.max()), "max_ref": float(np.abs(refp).max())}
            lg_ = e.dbg_read("logits").reshape(H // 8, Wd // 8, 80)[:, :, :65]
            refl = keepq["logits"][0].permute(1, 2, 0).numpy()
            layers["logits"] = {"maxabs": float(np.abs(lg_ - refl).max()), "max_ref": float(np.abs(refl).max()),
                                "vs_fp32_maxabs": None}
            out[name]["layers_vs_quant"] = layers
        comm = {tuple(k): i for i, k in enumerate(oq["kpts"])}
        idx = [(comm[tuple(k)], j) for j, k in enumerate(r["kpts"]) if tuple(k) in comm]
        io, ig = np.array(idx).T
        out[name]["vs_quant"]["desc_maxabs"] = float(np.abs(oq["desc"][io] - r["desc"][ig]).max())
        out[name]["vs_quant"]["score_maxrel"] = float((np.abs(oq["scores"][io] - r["scores"][ig]) / oq["scores"][io]).max())
        feats[name] = (r, o32, oq)
        print(name, json.dumps(out[name]), flush=True)
    # LightGlue on the config-2 pair: engine features -> engine LG vs both oracles on the SAME (engine) features
    e = engs[(480, 752)]
    ra, rb = feats["pair_a"][0], feats["pair_b"][0]
    mg, sg = e.lg_match(ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752)
    m32, _ = olg.lightglue(wl, ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752)
    keep = {}
    mq, sq = quant.lightglue_q(wl, ra["kpts"], rb["kpts"], ra["desc"], rb["desc"], 480, 752, 480, 752, keep)
    Lg = e.dbg_read("lg_L").reshape(1024, 1024)[:len(ra["kpts"]), :len(rb["kpts"])]
    out["lg_pair_engine_feats"] = {"vs_fp32": pair_stats(m32, mg), "vs_quant": pair_stats(mq, mg),
                                   "L_maxabs_vs_quant": float(np.abs(Lg - keep["L"]).max()),
                                   "L_maxabs_rowmax_vs_quant": float(np.abs(Lg.max(1) - keep["L"].max(1)).max())}
    # the EuRoC shape: 150 window points vs 662
    vio = synth.vio_points(150, 480, 752, synth.BASE_SEED + 3)
    e.frame_upload(a); dre = e.sp_describe(vio)
    kp_all = np.concatenate([rb["kpts"].astype(np.float32), vio]); e.frame_upload(b); dre_b = e.sp_describe(vio)
    de_all = np.concatenate([rb["desc"], dre_b])
    mg, sg = e.lg_match(vio, kp_all, dre, de_all, 480, 752, 480, 752)
    m32, _ = olg.lightglue(wl, vio, kp_all, dre, de_all, 480, 752, 480, 752)
    mq, sq = quant.lightglue_q(wl, vio, kp_all, dre, de_all, 480, 752, 480, 752)
    out["lg_150x662_engine_feats"] = {"vs_fp32": pair_stats(m32, mg), "vs_quant": pair_stats(mq, mg)}
    # synthetic-descriptor pairs (as tests/test_lightglue_gpu.py)
    for (M, N) in ((300, 400), (512, 512), (37, 1000)):
        rng = np.random.default_rng(M + N)
        d1 = rng.standard_normal((N, 256)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
        perm = rng.permutation(N)[:M]
        d0 = d1[perm] + 0.03 * rng.standard_normal((M, 256)).astype(np.float32); d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
        k1 = np.stack([rng.uniform(8, 744, N), rng.uniform(8, 472, N)], 1).astype(np.float32)
        k0 = k1[perm] + rng.normal(0, 1, (M, 2)).astype(np.float32)
        mg, sg = e.lg_match(k0, k1, d0, d1, 480, 752, 480, 752)
        m32, _ = olg.lightglue(wl, k0, k1, d0, d1, 480, 752, 480, 752)
        mq, sq = quant.lightglue_q(wl, k0, k1, d0, d1, 480, 752, 480, 752)
        out["lg_synth_%dx%d" % (M, N)] = {"vs_fp32": pair_stats(m32, mg), "vs_quant": pair_stats(mq, mg)}
    for k in out:
        if k.startswith("lg_"):
            print(k, json.dumps(out[k]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_diag.json"), "w") as f:
        json.dump(out, f, indent=1)
    for e in engs.values():
        e.close()


if __name__ == "__main__":
    main()
