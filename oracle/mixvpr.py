"""CPU fp32 restatement of the MixVPR global descriptor as D_VINS runs it (TEST INFRASTRUCTURE).

In-repo part (pre-processing) follows:
  * deep_net.cpp:1254-1323   MixVPRImpl::mix_extractor (gray->BGR on CPU :1259-1262; 320x320; mean/std :1298-1300)
  * deep_net.cpp:1211-1235   AffineMatrix::compute (anisotropic scale, cv::invertAffineTransform)
  * preprocess_kernel.cu:348-435  warp_affine_bilinear_and_normalize_plane_kernel_mix
The network arithmetic is NOT in the reference tree (engine `mix_512.engine` built from amaralibey/MixVPR,
README.md:34-67): VPRModel(resnet50, layers_to_crop=[4], MixVPR(in_channels=1024,in_h=20,in_w=20,
out_channels=256, mix_depth=4, mlp_ratio=1, out_rows=2)).  Restated from the published architecture
(SURVEY.md Appendix A.3); the ResNet trunk is cross-checked against torchvision in tests.  Parity unpinned.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .weights import RESNET_LAYERS

MIX_HW = 320
# deep_net.cpp:1298-1300 - BGR-ordered constants applied to RGB-ordered planes (reference quirk (1): replicate)
MEAN = np.array([0.406, 0.456, 0.485], dtype=np.float32)
STD = np.array([0.225, 0.224, 0.229], dtype=np.float32)


def affine_d2i(src_w: int, src_h: int, dst_w: int = MIX_HW, dst_h: int = MIX_HW) -> np.ndarray:
    """deep_net.cpp:1215-1229: i2d = diag(scale_x, scale_y) in f32; d2i via cv::invertAffineTransform, whose
    CV_32F branch evaluates in double and rounds to f32 (OpenCV imgwarp.cpp)."""
    sx = np.float32(dst_w) / np.float32(src_w)
    sy = np.float32(dst_h) / np.float32(src_h)
    m = np.array([sx, 0, 0, 0, sy, 0], dtype=np.float32)
    D = float(m[0]) * float(m[4]) - float(m[1]) * float(m[3])
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = float(m[4]) * D, float(m[0]) * D
    A12, A21 = -float(m[1]) * D, -float(m[3]) * D
    b1 = -A11 * float(m[2]) - A12 * float(m[5])
    b2 = -A21 * float(m[2]) - A22 * float(m[5])
    return np.array([A11, A12, b1, A21, A22, b2], dtype=np.float32)


def preprocess_mix(img_u8: np.ndarray) -> np.ndarray:
    """u8 HxW gray (or HxWx3 BGR) -> f32 [3,320,320] planes, exactly as preprocess_kernel.cu:348-435
    (non-centred inverse affine, bilinear with const 114 outside, floorf(v+0.5f), Invert swap, (x/255-mean)/std)."""
    if img_u8.ndim == 2:
        img = np.repeat(img_u8[:, :, None], 3, axis=2)      # cv::cvtColor GRAY2BGR (deep_net.cpp:1261)
    else:
        img = img_u8
    H, W = img.shape[:2]
    m = affine_d2i(W, H)
    dx = np.arange(MIX_HW, dtype=np.float32)[None, :].repeat(MIX_HW, 0)
    dy = np.arange(MIX_HW, dtype=np.float32)[:, None].repeat(MIX_HW, 1)
    sx = (m[0] * dx + m[1] * dy + m[2]).astype(np.float32)
    sy = (m[3] * dx + m[4] * dy + m[5]).astype(np.float32)
    oob = (sx <= -1) | (sx >= W) | (sy <= -1) | (sy >= H)
    yl = np.floor(sy).astype(np.int64); xl = np.floor(sx).astype(np.int64)
    yh, xh = yl + 1, xl + 1
    ly = (sy - yl.astype(np.float32)).astype(np.float32); lx = (sx - xl.astype(np.float32)).astype(np.float32)
    hy = np.float32(1) - ly; hx = np.float32(1) - lx
    w1, w2, w3, w4 = hy * hx, hy * lx, ly * hx, ly * lx
    imgf = img.astype(np.float32)
    const = np.float32(114)

    def fetch(y, x, ok):
        yy = np.clip(y, 0, H - 1); xx = np.clip(x, 0, W - 1)
        v = imgf[yy, xx]
        return np.where(ok[..., None], v, const)
    v1 = fetch(yl, xl, (yl >= 0) & (xl >= 0))
    v2 = fetch(yl, xh, (yl >= 0) & (xh < W))
    v3 = fetch(yh, xl, (yh < H) & (xl >= 0))
    v4 = fetch(yh, xh, (yh < H) & (xh < W))
    c = np.floor(w1[..., None] * v1 + w2[..., None] * v2 + w3[..., None] * v3 + w4[..., None] * v4
                 + np.float32(0.5)).astype(np.float32)
    c = np.where(oob[..., None], const, c)
    c = c[..., ::-1]                                          # Invert: c0 <-> c2
    a = np.float32(1.0) / np.float32(255.0)
    out = (c * a - MEAN[None, None, :]) / STD[None, None, :]
    return np.ascontiguousarray(out.transpose(2, 0, 1).astype(np.float32))


def _t(w, k):
    return torch.from_numpy(np.ascontiguousarray(w[k]))


def _bn(w, name, x):
    return F.batch_norm(x, _t(w, name + ".running_mean"), _t(w, name + ".running_var"),
                        _t(w, name + ".weight"), _t(w, name + ".bias"), training=False, eps=1e-5)


def backbone(w, x: torch.Tensor, keep=None) -> torch.Tensor:
    """torchvision ResNet-50 v1.5 conv1..layer3 (stride on the 3x3), eval-mode BN.  [1,3,320,320] -> [1,1024,20,20]."""
    p = "backbone.model."
    x = F.relu(_bn(w, p + "bn1", F.conv2d(x, _t(w, p + "conv1.weight"), None, stride=2, padding=3)))
    if keep is not None:
        keep["stem"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    if keep is not None:
        keep["pool"] = x
    for li, (planes, blocks, stride) in enumerate(RESNET_LAYERS, start=1):
        for b in range(blocks):
            q = "%slayer%d.%d." % (p, li, b)
            s = stride if b == 0 else 1
            idt = x
            o = F.relu(_bn(w, q + "bn1", F.conv2d(x, _t(w, q + "conv1.weight"))))
            o = F.relu(_bn(w, q + "bn2", F.conv2d(o, _t(w, q + "conv2.weight"), None, stride=s, padding=1)))
            o = _bn(w, q + "bn3", F.conv2d(o, _t(w, q + "conv3.weight")))
            if b == 0:
                idt = _bn(w, q + "downsample.1", F.conv2d(x, _t(w, q + "downsample.0.weight"), None, stride=s))
            x = F.relu(o + idt)
            if keep is not None:
                keep["layer%d.%d" % (li, b)] = x
    return x


def aggregator(w, feat: torch.Tensor, keep=None) -> torch.Tensor:
    """MixVPR aggregator: flatten(2) -> 4x FeatureMixer (x + W2 relu(W1 LN(x))) -> channel_proj -> row_proj -> L2."""
    x = feat.flatten(2)[0]                                     # [1024,400]
    for i in range(4):
        p = "aggregator.mix.%d.mix." % i
        h = F.layer_norm(x, (x.shape[-1],), _t(w, p + "0.weight"), _t(w, p + "0.bias"), eps=1e-5)
        h = F.relu(F.linear(h, _t(w, p + "1.weight"), _t(w, p + "1.bias")))
        x = x + F.linear(h, _t(w, p + "3.weight"), _t(w, p + "3.bias"))
        if keep is not None:
            keep["mix%d" % i] = x
    y = F.linear(x.t(), _t(w, "aggregator.channel_proj.weight"), _t(w, "aggregator.channel_proj.bias"))   # [400,256]
    z = F.linear(y.t(), _t(w, "aggregator.row_proj.weight"), _t(w, "aggregator.row_proj.bias"))           # [256,2]
    if keep is not None:
        keep["chan"] = y
        keep["rowp"] = z
    return F.normalize(z.flatten()[None], p=2, dim=-1)[0]


def mixvpr(w, img_u8: np.ndarray, keep=None) -> np.ndarray:
    """a4: mix_extractor(img) deep_net.cpp:1254-1323 -> 512 f32 unit vector."""
    x = torch.from_numpy(preprocess_mix(img_u8))[None]
    with torch.no_grad():
        f = backbone(w, x, keep)
        d = aggregator(w, f, keep)
    return d.numpy().astype(np.float32)
