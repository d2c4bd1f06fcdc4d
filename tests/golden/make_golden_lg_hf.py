"""Pin the LightGlue oracle against an INDEPENDENT published implementation of cvg/LightGlue:
`transformers.models.lightglue` (HF transformers 5.5, in the image), run with the same seeded synthetic weights.

The reference tree holds no LightGlue arithmetic (it deserialises a TensorRT engine exported from
fabio-sim/LightGlue-ONNX v0.1.3 = cvg/LightGlue weights v0.1_arxiv, README.md:74-85), and ships no golden vectors.
The HF port is the closest runnable statement of the same published network; its checkpoint conversion
(convert_lightglue_to_hf.py upstream) defines how the upstream state_dict keys map onto it, restated in
`to_hf_state_dict` below:
    posenc.Wr                          -> positional_encoder.projector
    transformers.i.self_attn.Wqkv      -> q/k/v_proj, rows de-interleaved: upstream row = head*192 + d*3 + {q,k,v}
    transformers.i.self_attn.out_proj  -> self_attention.o_proj
    transformers.i.self_attn.ffn.{0,1,3} -> self_mlp.{fc1,layer_norm,fc2}
    transformers.i.cross_attn.to_qk    -> cross_attention.q_proj AND k_proj (shared), to_v -> v_proj, to_out -> o_proj
    transformers.i.cross_attn.ffn.*    -> cross_mlp.*
    log_assignment.i.{final_proj,matchability} -> match_assignment_layers.i.{final_projection,matchability}
Early exit and point pruning are off (depth_confidence = width_confidence = -1), as in the ONNX export D_VINS uses.

Run where transformers is importable:  python tests/golden/make_golden_lg_hf.py
Output (committed):  tests/golden/lg_hf.npz  - inputs, HF matches / scores / final descriptors for three cases
(M == N, ragged M < N through HF's padding mask, D_VINS-shaped 40-window-points vs many).
Nothing here is imported by the product.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import weights, synth, superpoint as osp, lightglue as olg          # noqa: E402

H, W = 160, 224          # even sizes: HF's float halves == D_VINS's integer halves (deep_net.cpp:839-841)


def to_hf_state_dict(w):
    sd = {}
    t = lambda k: torch.from_numpy(np.ascontiguousarray(w[k]))
    sd["positional_encoder.projector.weight"] = t("posenc.Wr.weight")
    for i in range(olg.N_LAYERS):
        p = "transformers.%d." % i
        q = "transformer_layers.%d." % i
        wq, bq = t(p + "self_attn.Wqkv.weight"), t(p + "self_attn.Wqkv.bias")
        wq = wq.reshape(olg.HEADS, 64, 3, 256)
        bq = bq.reshape(olg.HEADS, 64, 3)
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            sd[q + "self_attention.%s.weight" % nm] = wq[:, :, j].reshape(256, 256).contiguous()
            sd[q + "self_attention.%s.bias" % nm] = bq[:, :, j].reshape(256).contiguous()
        sd[q + "self_attention.o_proj.weight"] = t(p + "self_attn.out_proj.weight")
        sd[q + "self_attention.o_proj.bias"] = t(p + "self_attn.out_proj.bias")
        for nm in ("q_proj", "k_proj"):
            sd[q + "cross_attention.%s.weight" % nm] = t(p + "cross_attn.to_qk.weight")
            sd[q + "cross_attention.%s.bias" % nm] = t(p + "cross_attn.to_qk.bias")
        sd[q + "cross_attention.v_proj.weight"] = t(p + "cross_attn.to_v.weight")
        sd[q + "cross_attention.v_proj.bias"] = t(p + "cross_attn.to_v.bias")
        sd[q + "cross_attention.o_proj.weight"] = t(p + "cross_attn.to_out.weight")
        sd[q + "cross_attention.o_proj.bias"] = t(p + "cross_attn.to_out.bias")
        for a, b in (("self_attn", "self_mlp"), ("cross_attn", "cross_mlp")):
            for src, dst in (("ffn.0", "fc1"), ("ffn.1", "layer_norm"), ("ffn.3", "fc2")):
                sd[q + "%s.%s.weight" % (b, dst)] = t(p + "%s.%s.weight" % (a, src))
                sd[q + "%s.%s.bias" % (b, dst)] = t(p + "%s.%s.bias" % (a, src))
        la = "log_assignment.%d." % i
        if la + "final_proj.weight" in w:
            ma = "match_assignment_layers.%d." % i
            sd[ma + "final_projection.weight"] = t(la + "final_proj.weight")
            sd[ma + "final_projection.bias"] = t(la + "final_proj.bias")
            sd[ma + "matchability.weight"] = t(la + "matchability.weight")
            sd[ma + "matchability.bias"] = t(la + "matchability.bias")
    return sd


def hf_model(w):
    from transformers import LightGlueConfig, LightGlueForKeypointMatching
    cfg = LightGlueConfig(depth_confidence=-1.0, width_confidence=-1.0, filter_threshold=olg.FILTER_THRESHOLD)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    m = LightGlueForKeypointMatching(cfg).eval()
    missing, unexpected = m.load_state_dict(to_hf_state_dict(w), strict=False)
    assert not unexpected, unexpected
    # only the SuperPoint detector, the unused assignment layers 0..7 and token_confidence stay at their init
    bad = [k for k in missing if not (k.startswith("keypoint_detector.") or k.startswith("token_confidence.")
                                      or (k.startswith("match_assignment_layers.") and not k.startswith("match_assignment_layers.8.")))]
    assert not bad, bad
    return m


def run_hf(m, k0, k1, d0, d1, h, w):
    """HF wants both images padded to one length with a mask; returns (matches0 [M], scores0 [M], x0, x1)."""
    M, N = len(k0), len(k1)
    n = max(M, N)
    kp = torch.zeros(1, 2, n, 2); de = torch.zeros(1, 2, n, 256); mask = torch.zeros(1, 2, n, dtype=torch.int64)
    kp[0, 0, :M] = torch.from_numpy(k0); kp[0, 1, :N] = torch.from_numpy(k1)
    de[0, 0, :M] = torch.from_numpy(d0); de[0, 1, :N] = torch.from_numpy(d1)
    mask[0, 0, :M] = 1; mask[0, 1, :N] = 1
    with torch.no_grad():
        matches, scores, _, hs, _ = m._match_image_pair(kp, de, h, w, mask=mask, output_hidden_states=True)
    x = hs[-3]                                   # descriptors after the last cross block [2, n, 256]
    return (matches.reshape(2, n)[0, :M].numpy().astype(np.int32), scores.reshape(2, n)[0, :M].numpy().astype(np.float32),
            x[0, :M].numpy().astype(np.float32), x[1, :N].numpy().astype(np.float32))


def cases():
    W_all = weights.synth_all()
    ws = weights.sub(W_all, "sp.")
    a, b = synth.make_pair(H, W, shift=(8, 8))
    ra, rb = osp.superpoint(ws, a), osp.superpoint(ws, b)
    ka, kb = ra["kpts"].astype(np.float32), rb["kpts"].astype(np.float32)
    da, db = ra["desc"], rb["desc"]
    n = min(len(ka), len(kb), 96)
    out = {"equal": (ka[:n], kb[:n], da[:n], db[:n]),
           "ragged": (ka[:57], kb[:n], da[:57], db[:n])}
    vio = synth.vio_points(40, H, W, 5, min_dist=12)
    dv = osp.superpoint_recover(ws, a, vio, feat=ra["feat"])
    out["window"] = (vio.astype(np.float32), np.concatenate([kb, vio]).astype(np.float32)[:128],
                     dv, np.concatenate([db, osp.superpoint_recover(ws, b, vio, feat=rb["feat"])])[:128])
    return weights.sub(W_all, "lg."), out


def main():
    wl, cs = cases()
    m = hf_model(wl)
    save = {"h": H, "w": W}
    for tag, (k0, k1, d0, d1) in cs.items():
        m0, s0, x0, x1 = run_hf(m, k0, k1, d0, d1, H, W)
        keep = {}
        mo, so = olg.lightglue(wl, k0, k1, d0, d1, H, W, H, W, keep=keep)
        valid = np.nonzero(m0 >= 0)[0]
        pairs = np.stack([valid, m0[valid]], 1).astype(np.int32)
        print("%-7s M=%3d N=%3d  HF matches %3d  oracle matches %3d  identical=%s  max|ds|=%.2e  max|dx|=%.2e" % (
            tag, len(k0), len(k1), len(pairs), len(mo), np.array_equal(pairs, mo),
            np.abs(s0[valid] - so).max() if len(pairs) == len(mo) and len(mo) else -1,
            max(np.abs(x0 - keep["x0_8"]).max(), np.abs(x1 - keep["x1_8"]).max())))
        save.update({tag + "_k0": k0, tag + "_k1": k1, tag + "_d0": d0.astype(np.float32), tag + "_d1": d1.astype(np.float32),
                     tag + "_matches": pairs, tag + "_mscores": s0[valid], tag + "_x0": x0, tag + "_x1": x1})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lg_hf.npz"), **save)


if __name__ == "__main__":
    main()
