"""Convert the real checkpoints D_VINS uses into the engine's DVWGT001 weight file.

    python -m d_vins_b200.convert_weights --superpoint superpoint_v1.pth --lightglue superpoint_lightglue.pth \
        --mixvpr "resnet50_MixVPR_512_channels(256)_rows(2).ckpt" -o dvins.dvw

Sources (reference README.md:30-85): SuperPoint `superpoint_v1.pth` (keys conv1a..convDb, export/superpoint.py:126-141),
cvg/LightGlue `superpoint_lightglue.pth` release v0.1_arxiv (keys posenc.Wr, transformers.{i}.self_attn / cross_attn,
log_assignment.{i}), amaralibey/MixVPR checkpoint (Lightning: `state_dict` with backbone.model.* / aggregator.*).
Every tensor is stored as float32 under the prefixes `sp.`, `lg.`, `mix.` with the upstream key names - exactly what the
engine (csrc/sp.cu, lg.cu, mix.cu) and the CPU oracle look up; BatchNorm folding, Wqkv de-interleaving and fp16 rounding
happen inside the engine at load time.  The checkpoints cannot be fetched offline, so the shipped tests run on seeded
synthetic weights with the same key set; this tool validates the key set and shapes it writes against that contract.

File format "DVWGT001" (little endian): 8s magic | u32 n | n x { u32 name_len, name, u32 ndim, u32 dims[ndim],
u64 offset, u64 nbytes } | 64-byte aligned raw float32 data   (reader: csrc/weights.cpp).
"""
from __future__ import annotations

import argparse
import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"DVWGT001"
LG_LAYERS = 9
RESNET_LAYERS = [(64, 3), (128, 4), (256, 6)]      # torchvision ResNet-50 [:layer3]: (planes, blocks)


def save_dvw(path: str, tensors: "OrderedDict[str, np.ndarray]") -> None:
    names = list(tensors)
    arrs = [np.ascontiguousarray(np.asarray(tensors[k], dtype=np.float32)) for k in names]
    hdr = len(MAGIC) + 4 + sum(4 + len(k.encode()) + 4 + 4 * a.ndim + 16 for k, a in zip(names, arrs))
    off = (hdr + 63) // 64 * 64
    offsets = []
    for a in arrs:
        offsets.append(off)
        off = (off + a.nbytes + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(names)))
        for k, a, o in zip(names, arrs, offsets):
            kb = k.encode()
            f.write(struct.pack("<I", len(kb)) + kb + struct.pack("<I", a.ndim))
            f.write(struct.pack("<%dI" % a.ndim, *a.shape) + struct.pack("<QQ", o, a.nbytes))
        for a, o in zip(arrs, offsets):
            f.seek(o)
            f.write(a.tobytes())
        f.truncate(off)


def expected_keys() -> "OrderedDict[str, tuple]":
    """The key -> shape contract of the engine's loaders."""
    e: "OrderedDict[str, tuple]" = OrderedDict()
    for n, ci, co in (("conv1a", 1, 64), ("conv1b", 64, 64), ("conv2a", 64, 64), ("conv2b", 64, 64), ("conv3a", 64, 128),
                      ("conv3b", 128, 128), ("conv4a", 128, 128), ("conv4b", 128, 128), ("convPa", 128, 256),
                      ("convDa", 128, 256)):
        e["sp.%s.weight" % n] = (co, ci, 3, 3); e["sp.%s.bias" % n] = (co,)
    e["sp.convPb.weight"] = (65, 256, 1, 1); e["sp.convPb.bias"] = (65,)
    e["sp.convDb.weight"] = (256, 256, 1, 1); e["sp.convDb.bias"] = (256,)
    e["lg.posenc.Wr.weight"] = (32, 2)
    for i in range(LG_LAYERS):
        p = "lg.transformers.%d." % i
        for n, co, ci in (("self_attn.Wqkv", 768, 256), ("self_attn.out_proj", 256, 256), ("self_attn.ffn.0", 512, 512),
                          ("self_attn.ffn.3", 256, 512), ("cross_attn.to_qk", 256, 256), ("cross_attn.to_v", 256, 256),
                          ("cross_attn.to_out", 256, 256), ("cross_attn.ffn.0", 512, 512), ("cross_attn.ffn.3", 256, 512)):
            e[p + n + ".weight"] = (co, ci); e[p + n + ".bias"] = (co,)
        for blk in ("self_attn", "cross_attn"):
            e[p + blk + ".ffn.1.weight"] = (512,); e[p + blk + ".ffn.1.bias"] = (512,)
    p = "lg.log_assignment.%d." % (LG_LAYERS - 1)
    e[p + "final_proj.weight"] = (256, 256); e[p + "final_proj.bias"] = (256,)
    e[p + "matchability.weight"] = (1, 256); e[p + "matchability.bias"] = (1,)

    def bn(name, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            e["%s.%s" % (name, s)] = (c,)
    pre = "mix.backbone.model."
    e[pre + "conv1.weight"] = (64, 3, 7, 7); bn(pre + "bn1", 64)
    inpl = 64
    for li, (planes, blocks) in enumerate(RESNET_LAYERS, start=1):
        for b in range(blocks):
            q = "%slayer%d.%d." % (pre, li, b)
            e[q + "conv1.weight"] = (planes, inpl, 1, 1); bn(q + "bn1", planes)
            e[q + "conv2.weight"] = (planes, planes, 3, 3); bn(q + "bn2", planes)
            e[q + "conv3.weight"] = (planes * 4, planes, 1, 1); bn(q + "bn3", planes * 4)
            if b == 0:
                e[q + "downsample.0.weight"] = (planes * 4, inpl, 1, 1); bn(q + "downsample.1", planes * 4)
            inpl = planes * 4
    for i in range(4):
        q = "mix.aggregator.mix.%d.mix." % i
        e[q + "0.weight"] = (400,); e[q + "0.bias"] = (400,)
        e[q + "1.weight"] = (400, 400); e[q + "1.bias"] = (400,)
        e[q + "3.weight"] = (400, 400); e[q + "3.bias"] = (400,)
    e["mix.aggregator.channel_proj.weight"] = (256, 1024); e["mix.aggregator.channel_proj.bias"] = (256,)
    e["mix.aggregator.row_proj.weight"] = (2, 400); e["mix.aggregator.row_proj.bias"] = (2,)
    return e


def _state_dict(path):
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(ck, dict) and "state_dict" in ck and isinstance(ck["state_dict"], dict):
        ck = ck["state_dict"]            # PyTorch Lightning checkpoint (MixVPR)
    return {k: v.detach().cpu().float().numpy() for k, v in ck.items() if hasattr(v, "detach")}


def convert(superpoint=None, lightglue=None, mixvpr=None, state_dicts=None) -> "OrderedDict[str, np.ndarray]":
    """Paths to the three checkpoints (or, for tests, `state_dicts` = {"sp": {...}, "lg": {...}, "mix": {...}} of numpy
    arrays under the upstream key names).  Returns the prefixed tensor dict, validated against expected_keys()."""
    sds = dict(state_dicts or {})
    for pre, path in (("sp", superpoint), ("lg", lightglue), ("mix", mixvpr)):
        if path:
            sds[pre] = _state_dict(path)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    want = expected_keys()
    for key, shape in want.items():
        pre, name = key.split(".", 1)
        if pre not in sds:
            continue
        sd = sds[pre]
        cand = [name, "module." + name, "model." + name]
        src = next((c for c in cand if c in sd), None)
        if src is None:
            raise KeyError("checkpoint for '%s' lacks tensor %s" % (pre, name))
        a = np.asarray(sd[src], np.float32)
        if a.size != int(np.prod(shape)):
            raise ValueError("%s: shape %s, expected %s" % (key, a.shape, shape))
        out[key] = a.reshape(shape)
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--superpoint"); ap.add_argument("--lightglue"); ap.add_argument("--mixvpr")
    ap.add_argument("-o", "--output", required=True)
    a = ap.parse_args()
    t = convert(a.superpoint, a.lightglue, a.mixvpr)
    save_dvw(a.output, t)
    print("wrote %s: %d tensors, %.1f M parameters" % (a.output, len(t), sum(v.size for v in t.values()) / 1e6))


if __name__ == "__main__":
    main()
