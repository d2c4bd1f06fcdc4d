"""Weight container + seeded synthetic weights (oracle side).

Real checkpoints (superpoint_v1.pth, superpoint_lightglue.pth, the MixVPR .ckpt) are not
obtainable offline (reference .gitignore:24-31 excludes them), so parity runs on *seeded
synthetic* weights whose tensor names mirror the upstream ``state_dict`` keys
(export/superpoint.py:126-141; SURVEY.md §8(c) last row) - a real checkpoint converted with
``save_weights`` drops straight in.

File format "DVWGT001" (little endian, all tensors float32):
    8s magic | u32 n | n x { u32 name_len, name, u32 ndim, u32 dims[ndim], u64 offset, u64 nbytes }
    | 64-byte aligned raw data.
The CUDA engine's loader (d_vins_b200/csrc/weights.cpp) reads the same format.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"DVWGT001"


def save_weights(path: str, tensors: "OrderedDict[str, np.ndarray]") -> None:
    names = list(tensors.keys())
    arrs = [np.ascontiguousarray(np.asarray(tensors[k], dtype=np.float32)) for k in names]
    # header size
    hdr = len(MAGIC) + 4
    for k, a in zip(names, arrs):
        hdr += 4 + len(k.encode()) + 4 + 4 * a.ndim + 8 + 8
    off = (hdr + 63) // 64 * 64
    offsets = []
    for a in arrs:
        offsets.append(off)
        off = (off + a.nbytes + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(names)))
        for k, a, o in zip(names, arrs, offsets):
            kb = k.encode()
            f.write(struct.pack("<I", len(kb)))
            f.write(kb)
            f.write(struct.pack("<I", a.ndim))
            for d in a.shape:
                f.write(struct.pack("<I", d))
            f.write(struct.pack("<QQ", o, a.nbytes))
        for a, o in zip(arrs, offsets):
            f.seek(o)
            f.write(a.tobytes())
        f.truncate(off)


def load_weights(path: str) -> "OrderedDict[str, np.ndarray]":
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == MAGIC, "bad weight file magic"
    p = 8
    (n,) = struct.unpack_from("<I", buf, p)
    p += 4
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, p)
        p += 4
        name = buf[p:p + ln].decode()
        p += ln
        (nd,) = struct.unpack_from("<I", buf, p)
        p += 4
        dims = struct.unpack_from("<%dI" % nd, buf, p)
        p += 4 * nd
        off, nb = struct.unpack_from("<QQ", buf, p)
        p += 16
        out[name] = np.frombuffer(buf, dtype=np.float32, count=nb // 4, offset=off).reshape(dims).copy()
    return out


# ----------------------------------------------------------------------------------------
# synthetic weights
# ----------------------------------------------------------------------------------------

def _he(rng, cout, cin, k, gain=2.0):
    fan_in = cin * k * k
    return (rng.standard_normal((cout, cin, k, k)) * np.sqrt(gain / fan_in)).astype(np.float32)


def synth_superpoint(seed: int = 20240701, calibrate: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Seeded SuperPoint weights, keys as export/superpoint.py:126-141.

    He-normal convs (activations stay O(1) through the VGG stack); the detector head is scaled so
    the 65-way logits have std ~ 3 and the dustbin (channel 64) carries a positive bias: a peaky,
    non-degenerate score map (SURVEY.md §7 "hard parts": a flat score map is tie-dominated).
    """
    rng = np.random.default_rng(seed)
    w: "OrderedDict[str, np.ndarray]" = OrderedDict()
    chans = [("conv1a", 1, 64), ("conv1b", 64, 64), ("conv2a", 64, 64), ("conv2b", 64, 64),
             ("conv3a", 64, 128), ("conv3b", 128, 128), ("conv4a", 128, 128), ("conv4b", 128, 128),
             ("convPa", 128, 256), ("convDa", 128, 256)]
    for name, cin, cout in chans:
        w[name + ".weight"] = _he(rng, cout, cin, 3)
        w[name + ".bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)
    # conv1a sees a [0,1] image with a big DC term: centre its filters so edges dominate
    w["conv1a.weight"] = (w["conv1a.weight"] - w["conv1a.weight"].mean(axis=(1, 2, 3), keepdims=True)) * 4.0
    w["convPb.weight"] = _he(rng, 65, 256, 1, gain=1.0) * 3.0
    b = (rng.standard_normal(65) * 0.1).astype(np.float32)
    b[64] = 3.0
    w["convPb.bias"] = b
    w["convDb.weight"] = _he(rng, 256, 256, 1, gain=1.0)
    w["convDb.bias"] = (rng.standard_normal(256) * 0.05).astype(np.float32)
    if calibrate:
        # Random ReLU features share a large common-mode component, which makes every descriptor look alike
        # (mean cosine 0.88).  Centre the descriptor head on a seeded calibration frame so descriptors are
        # discriminative (deterministic in `seed`; uses the oracle's own forward pass).
        from . import superpoint as _sp
        from . import synth as _synth
        keep = {}
        _sp.superpoint(w, _synth.make_frame(480, 752, seed + 99), keep=keep)
        m = keep["convDa"][0].mean(dim=(1, 2)).numpy()
        w["convDb.bias"] = (-(w["convDb.weight"][:, :, 0, 0] @ m)).astype(np.float32)
    return w


def synth_lightglue(seed: int = 20240702, n_layers: int = 9) -> "OrderedDict[str, np.ndarray]":
    """Seeded LightGlue (SuperPoint variant) weights; key names follow cvg/LightGlue's state_dict
    (SURVEY.md §8(c)): posenc.Wr, transformers.{i}.self_attn.*, transformers.{i}.cross_attn.*,
    log_assignment.{i}.*  (only i = n_layers-1 is used by the ONNX export)."""
    rng = np.random.default_rng(seed)
    w: "OrderedDict[str, np.ndarray]" = OrderedDict()
    d = 256

    def lin(name, cout, cin, scale=1.0, bias_std=0.02):
        w[name + ".weight"] = (rng.standard_normal((cout, cin)) * scale / np.sqrt(cin)).astype(np.float32)
        w[name + ".bias"] = (rng.standard_normal(cout) * bias_std).astype(np.float32)

    w["posenc.Wr.weight"] = (rng.standard_normal((32, 2)) * 2.0).astype(np.float32)
    for i in range(n_layers):
        p = "transformers.%d." % i
        lin(p + "self_attn.Wqkv", 3 * d, d, scale=2.0)
        lin(p + "self_attn.out_proj", d, d)
        lin(p + "self_attn.ffn.0", 2 * d, 2 * d)
        w[p + "self_attn.ffn.1.weight"] = (1.0 + 0.1 * rng.standard_normal(2 * d)).astype(np.float32)
        w[p + "self_attn.ffn.1.bias"] = (0.05 * rng.standard_normal(2 * d)).astype(np.float32)
        lin(p + "self_attn.ffn.3", d, 2 * d, scale=0.5)
        lin(p + "cross_attn.to_qk", d, d, scale=2.0)
        lin(p + "cross_attn.to_v", d, d)
        lin(p + "cross_attn.to_out", d, d)
        lin(p + "cross_attn.ffn.0", 2 * d, 2 * d)
        w[p + "cross_attn.ffn.1.weight"] = (1.0 + 0.1 * rng.standard_normal(2 * d)).astype(np.float32)
        w[p + "cross_attn.ffn.1.bias"] = (0.05 * rng.standard_normal(2 * d)).astype(np.float32)
        lin(p + "cross_attn.ffn.3", d, 2 * d, scale=0.5)
    p = "log_assignment.%d." % (n_layers - 1)
    lin(p + "final_proj", d, d, scale=4.0)
    w[p + "matchability.weight"] = (rng.standard_normal((1, d)) * 0.5 / np.sqrt(d)).astype(np.float32)
    w[p + "matchability.bias"] = np.array([4.0], dtype=np.float32)
    return w


# torchvision ResNet-50 v1.5 [:layer3] structure: (planes, blocks, stride)
RESNET_LAYERS = [(64, 3, 1), (128, 4, 2), (256, 6, 2)]


def synth_mixvpr(seed: int = 20240703) -> "OrderedDict[str, np.ndarray]":
    """Seeded MixVPR weights: ``backbone.model.*`` = torchvision ResNet-50 cropped before layer4,
    ``aggregator.*`` = MixVPR(in_channels=1024,in_h=20,in_w=20,out_channels=256,mix_depth=4,
    mlp_ratio=1,out_rows=2) (reference README.md:35-45)."""
    rng = np.random.default_rng(seed)
    w: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def bn(name, c, gamma=1.0):
        w[name + ".weight"] = (gamma * (1.0 + 0.1 * rng.standard_normal(c))).astype(np.float32)
        w[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".running_var"] = (1.0 + 0.2 * rng.uniform(-1, 1, c)).astype(np.float32)

    pre = "backbone.model."
    w[pre + "conv1.weight"] = _he(rng, 64, 3, 7)
    bn(pre + "bn1", 64)
    inpl = 64
    for li, (planes, blocks, stride) in enumerate(RESNET_LAYERS, start=1):
        for b in range(blocks):
            p = "%slayer%d.%d." % (pre, li, b)
            w[p + "conv1.weight"] = _he(rng, planes, inpl, 1)
            bn(p + "bn1", planes)
            w[p + "conv2.weight"] = _he(rng, planes, planes, 3)
            bn(p + "bn2", planes)
            w[p + "conv3.weight"] = _he(rng, planes * 4, planes, 1)
            bn(p + "bn3", planes * 4, gamma=0.4)
            if b == 0:
                w[p + "downsample.0.weight"] = _he(rng, planes * 4, inpl, 1, gain=1.0)
                bn(p + "downsample.1", planes * 4)
            inpl = planes * 4
    hw = 400
    for i in range(4):
        p = "aggregator.mix.%d.mix." % i
        w[p + "0.weight"] = (1.0 + 0.1 * rng.standard_normal(hw)).astype(np.float32)
        w[p + "0.bias"] = (0.05 * rng.standard_normal(hw)).astype(np.float32)
        w[p + "1.weight"] = (rng.standard_normal((hw, hw)) * np.sqrt(2.0 / hw)).astype(np.float32)
        w[p + "1.bias"] = (0.02 * rng.standard_normal(hw)).astype(np.float32)
        w[p + "3.weight"] = (rng.standard_normal((hw, hw)) * 0.5 / np.sqrt(hw)).astype(np.float32)
        w[p + "3.bias"] = (0.02 * rng.standard_normal(hw)).astype(np.float32)
    w["aggregator.channel_proj.weight"] = (rng.standard_normal((256, 1024)) / np.sqrt(1024)).astype(np.float32)
    w["aggregator.channel_proj.bias"] = (0.02 * rng.standard_normal(256)).astype(np.float32)
    w["aggregator.row_proj.weight"] = (rng.standard_normal((2, hw)) / np.sqrt(hw)).astype(np.float32)
    w["aggregator.row_proj.bias"] = (0.02 * rng.standard_normal(2)).astype(np.float32)
    return w


def synth_all(seed: int = 20240701) -> "OrderedDict[str, np.ndarray]":
    """One file with all three nets, prefixed ``sp.``, ``lg.``, ``mix.``."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for pre, d in (("sp.", synth_superpoint(seed)), ("lg.", synth_lightglue(seed + 1)),
                   ("mix.", synth_mixvpr(seed + 2))):
        for k, v in d.items():
            out[pre + k] = v
    return out


def sub(weights, prefix: str) -> "OrderedDict[str, np.ndarray]":
    return OrderedDict((k[len(prefix):], v) for k, v in weights.items() if k.startswith(prefix))
