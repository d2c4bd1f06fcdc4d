"""CPU fp32 restatement of SuperPoint / SP-recover as D_VINS runs them (TEST INFRASTRUCTURE).

Follows, line by line:
  * export/superpoint.py:52-69   simple_nms
  * export/superpoint.py:73-80   top_k_keypoints
  * export/superpoint.py:83-98   sample_descriptors
  * export/superpoint.py:153-224 SuperPoint.forward
  * export/ultrapoint.py:101-127 UltraPoint.forward (SP_RE: describe at caller's float keypoints)
  * loop_fusion/src/deep_net/tensorrt_tools/preprocess_kernel.cu:193-346 (u8 -> f32 plane)
  * loop_fusion/src/deep_net/deep_net.cpp:578 (alpha = 1/255.f, beta 0, "Invert"), :633-662 (post)

Pinned against the reference's own export modules by tests/golden/make_golden.py.
Tie rule (PyTorch leaves topk ties unspecified): score descending, then linear index ascending.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NMS_RADIUS = 4          # export/superpoint.py:112
DET_THRESH = 0.0005     # :114
BORDER = 4              # :115
MAX_KPTS = 512          # README.md:128-161 (export with max_num_keypoints=512)


def preprocess_gray(img_u8: np.ndarray) -> np.ndarray:
    """u8 HxW (or HxWx3 BGR) -> f32 HxW in [0,1].

    preprocess_kernel.cu:193-346 with width_adj==cols, height_adj==rows (SURVEY.md §8
    precondition): the affine is the identity, bilinear weights are exactly (1,0,0,0), the
    ``floorf(v+0.5f)`` re-quantisation is a no-op, so the result is ``u8 * (1/255.f)``
    (deep_net.cpp:578: a *multiply* by the rounded reciprocal).  3-channel input: BGR->RGB swap
    ("Invert") then gray = 0.299*c0 + 0.587*c1 + 0.114*c2 in fp64 (preprocess_kernel.cu:257-260, :280)."""
    a = np.float32(1.0) / np.float32(255.0)
    if img_u8.ndim == 3 and img_u8.shape[2] == 3:
        c = img_u8.astype(np.float32)
        b, g, r = c[..., 0], c[..., 1], c[..., 2]
        c0, c1, c2 = r * a, g * a, b * a        # Invert: c0<->c2 then alpha/beta per channel
        # preprocess_kernel.cu:280: DOUBLE literals -> the mix is evaluated in fp64, rounded to fp32 once (pinned against
        # the compiled reference kernel on the GPU: tests/test_refpre_gpu.py)
        return (0.299 * c0.astype(np.float64) + 0.587 * c1.astype(np.float64)
                + 0.114 * c2.astype(np.float64)).astype(np.float32)
    return img_u8.astype(np.float32) * a


def _t(w, k):
    return torch.from_numpy(np.ascontiguousarray(w[k]))


def encoder(w, x: torch.Tensor, keep=None) -> torch.Tensor:
    """export/superpoint.py:159-169 - VGG encoder, 1x1xHxW -> 1x128x(H/8)x(W/8)."""
    def cr(x, n):
        y = F.relu(F.conv2d(x, _t(w, n + ".weight"), _t(w, n + ".bias"), padding=1))
        if keep is not None:
            keep[n] = y
        return y
    x = cr(x, "conv1a"); x = cr(x, "conv1b"); x = F.max_pool2d(x, 2, 2)
    x = cr(x, "conv2a"); x = cr(x, "conv2b"); x = F.max_pool2d(x, 2, 2)
    x = cr(x, "conv3a"); x = cr(x, "conv3b"); x = F.max_pool2d(x, 2, 2)
    x = cr(x, "conv4a"); x = cr(x, "conv4b")
    return x


def score_map(w, feat: torch.Tensor, keep=None) -> torch.Tensor:
    """export/superpoint.py:172-178 - detector head + softmax-65 + depth-to-space -> [H8*8, W8*8]."""
    cPa = F.relu(F.conv2d(feat, _t(w, "convPa.weight"), _t(w, "convPa.bias"), padding=1))
    logits = F.conv2d(cPa, _t(w, "convPb.weight"), _t(w, "convPb.bias"))
    if keep is not None:
        keep["convPa"] = cPa
        keep["logits"] = logits
    s = F.softmax(logits, 1)[:, :-1]
    _, _, h, wd = s.shape
    s = s.permute(0, 2, 3, 1).reshape(1, h, wd, 8, 8)
    s = s.permute(0, 1, 3, 2, 4).reshape(1, h * 8, wd * 8)
    return s[0]


def simple_nms(scores: torch.Tensor, r: int = NMS_RADIUS) -> torch.Tensor:
    """export/superpoint.py:52-69 (exact float equality; -inf implicit padding)."""
    def mp(x):
        return F.max_pool2d(x[None, None], kernel_size=2 * r + 1, stride=1, padding=r)[0, 0]
    zeros = torch.zeros_like(scores)
    max_mask = scores == mp(scores)
    for _ in range(2):
        supp_mask = mp(max_mask.float()) > 0
        supp_scores = torch.where(supp_mask, zeros, scores)
        new_max_mask = supp_scores == mp(supp_scores)
        max_mask = max_mask | (new_max_mask & (~supp_mask))
    return torch.where(max_mask, scores, zeros)


def select_keypoints(nms: torch.Tensor, k: int = MAX_KPTS, border: int = BORDER,
                     thresh: float = DET_THRESH):
    """export/superpoint.py:183-207: border -> -1, strict threshold, row-major candidates, top-k.

    Returns (kpts_xy int64 [N,2], scores f32 [N], lin_index int64 [N]).  If candidates <= k they stay in
    row-major order *unsorted* (export/superpoint.py:76-77), else score-descending with the oracle's
    tie rule (score desc, linear index asc)."""
    s = nms.clone()
    s[:border] = -1
    s[:, :border] = -1
    s[-border:] = -1
    s[:, -border:] = -1
    H, W = s.shape
    ys, xs = torch.where(s > thresh)
    sc = s[ys, xs]
    lin = ys * W + xs
    if k < lin.numel():
        sc_np = sc.numpy()
        lin_np = lin.numpy()
        order = np.lexsort((lin_np, -sc_np.astype(np.float64)))[:k]     # primary: -score, secondary: index
        order = torch.from_numpy(order)
        sc, lin, ys, xs = sc[order], lin[order], ys[order], xs[order]
    kp = torch.stack((xs, ys), dim=-1)
    return kp, sc, lin


def dense_descriptors(w, feat: torch.Tensor, keep=None) -> torch.Tensor:
    """export/superpoint.py:212-214: descriptor head + channel L2-normalise -> 1x256xhxw."""
    cDa = F.relu(F.conv2d(feat, _t(w, "convDa.weight"), _t(w, "convDa.bias"), padding=1))
    d = F.conv2d(cDa, _t(w, "convDb.weight"), _t(w, "convDb.bias"))
    if keep is not None:
        keep["convDa"] = cDa
        keep["convDb"] = d
    return F.normalize(d, p=2, dim=1)


def sample_descriptors(kpts_xy: torch.Tensor, desc: torch.Tensor, s: int = 8) -> torch.Tensor:
    """export/superpoint.py:83-98 / export/ultrapoint.py:39-54 -> [N,256] (unit rows)."""
    b, c, h, wd = desc.shape
    k = kpts_xy.to(torch.float32) - s / 2 + 0.5
    kx = torch.div(k[..., 0], (wd * s - s / 2 - 0.5))
    ky = torch.div(k[..., 1], (h * s - s / 2 - 0.5))
    g = torch.stack((kx, ky), dim=-1) * 2 - 1
    d = F.grid_sample(desc, g.view(b, 1, -1, 2), mode="bilinear", align_corners=True)
    d = F.normalize(d.reshape(b, c, -1), p=2, dim=1)
    return d[0].t().contiguous()


def normalize_kpts(kpts_xy: np.ndarray, width: int, height: int) -> np.ndarray:
    """deep_net.cpp:633-659 / :839-841 + preprocess_kernel.cu:52-65: ``(kp - [W/2, H/2]) / max(W/2, H/2)``
    with *integer* halves (quirk (2))."""
    sw = np.float32(width // 2)
    sh = np.float32(height // 2)
    sc = np.float32(max(width // 2, height // 2))
    k = kpts_xy.astype(np.float32).copy()
    k[:, 0] = (k[:, 0] - sw) / sc
    k[:, 1] = (k[:, 1] - sh) / sc
    return k


def superpoint(w, img_u8: np.ndarray, max_kpts: int = MAX_KPTS, keep=None):
    """a1: EstimatorImpl::sp_extractor(img) deep_net.cpp:527-688.

    Returns dict: kpts [N,2] int32 (x,y), scores [N] f32, desc [N,256] f32, kpts_norm [N,2] f32,
    plus 'score_map' and 'nms' (debug)."""
    H, W = img_u8.shape[:2]
    x = torch.from_numpy(preprocess_gray(img_u8))[None, None]
    with torch.no_grad():
        feat = encoder(w, x, keep)
        smap = score_map(w, feat, keep)
        nms = simple_nms(smap)
        kp, sc, lin = select_keypoints(nms, max_kpts)
        dmap = dense_descriptors(w, feat, keep)
        desc = sample_descriptors(kp[None], dmap)
    kp_np = kp.numpy().astype(np.int32)
    return {
        "kpts": kp_np, "scores": sc.numpy().astype(np.float32), "desc": desc.numpy(),
        "kpts_norm": normalize_kpts(kp_np, W, H), "score_map": smap.numpy(), "nms": nms.numpy(),
        "feat": feat, "dmap": dmap,
    }


def superpoint_recover(w, img_u8: np.ndarray, kpts_xy: np.ndarray, feat=None):
    """a2: sp_extractor(img, kpts) deep_net.cpp:690-812; export/ultrapoint.py:101-127.
    ``feat`` may be shared with superpoint() (same image => same encoder output)."""
    with torch.no_grad():
        if feat is None:
            x = torch.from_numpy(preprocess_gray(img_u8))[None, None]
            feat = encoder(w, x)
        dmap = dense_descriptors(w, feat)
        desc = sample_descriptors(torch.from_numpy(kpts_xy.astype(np.float32))[None], dmap)
    return desc.numpy()
