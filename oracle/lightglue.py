"""CPU fp32 restatement of LightGlue (SuperPoint variant) as D_VINS runs it (TEST INFRASTRUCTURE).

The arithmetic is NOT in the reference tree: D_VINS deserialises `superpoint_lightglue_10_1024.engine`
built from fabio-sim/LightGlue-ONNX v0.1.3 (README.md:74-85), itself an export of cvg/LightGlue
(weights release v0.1_arxiv).  This file restates that published architecture:
9 layers, d=256, 4 heads x 64, learnable-Fourier rotary encoding on self-attention only, bidirectional
cross attention with shared weights, sigmoid-log-double-softmax assignment, filter_threshold 0.1, no early
exit / pruning.  Parity is anchored on the reference call site:
  * deep_net.cpp:814-1000  EstimatorImpl::lg_matcher (bindings kpts0,kpts1,desc0,desc1 -> matches0,mscores0)
  * deep_net.cpp:839-841, :874-880 + preprocess_kernel.cu:52-65  caller-side keypoint normalisation
  * keyframe.cpp:623-654  consumption order (pairs ascending in i0, each i0 at most once)
Parity unpinned (no golden vectors exist upstream; see oracle/__init__.py).
Tie rule: argmax -> lowest index.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

N_LAYERS = 9
HEADS = 4
FILTER_THRESHOLD = 0.1


def _t(w, k):
    return torch.from_numpy(np.ascontiguousarray(w[k]))


def _lin(w, name, x):
    return F.linear(x, _t(w, name + ".weight"), _t(w, name + ".bias"))


def normalize_kpts(kpts_xy: np.ndarray, width: int, height: int) -> np.ndarray:
    """deep_net.cpp:839-841, :874-880: shift = (W/2, H/2) with INTEGER division, scale = max of the two."""
    sw = np.float32(width // 2)
    sh = np.float32(height // 2)
    sc = np.float32(max(width // 2, height // 2))
    k = kpts_xy.astype(np.float32).copy()
    k[:, 0] = (k[:, 0] - sw) / sc
    k[:, 1] = (k[:, 1] - sh) / sc
    return k


def posenc(w, kpts_norm: torch.Tensor):
    """LearnableFourierPositionalEncoding: P = kp Wr^T (no bias); cos/sin each repeat_interleave'd x2 -> [n,64]."""
    proj = F.linear(kpts_norm, _t(w, "posenc.Wr.weight"))
    return torch.cos(proj).repeat_interleave(2, dim=-1), torch.sin(proj).repeat_interleave(2, dim=-1)


def _rotate_half(t):
    t2 = t.unflatten(-1, (-1, 2))
    x1, x2 = t2.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(start_dim=-2)


def _rope(cs, t):   # t [h,n,64]
    return t * cs[0][None] + _rotate_half(t) * cs[1][None]


def _ffn(w, p, x, msg):
    h = _lin(w, p + "ffn.0", torch.cat([x, msg], -1))
    h = F.layer_norm(h, (h.shape[-1],), _t(w, p + "ffn.1.weight"), _t(w, p + "ffn.1.bias"), eps=1e-5)
    h = F.gelu(h)       # erf GELU
    return _lin(w, p + "ffn.3", h)


def self_block(w, i, x, enc):
    p = "transformers.%d.self_attn." % i
    n = x.shape[0]
    qkv = _lin(w, p + "Wqkv", x).unflatten(-1, (HEADS, -1, 3)).transpose(0, 1)   # [h,n,64,3]
    q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
    q, k = _rope(enc, q), _rope(enc, k)
    att = F.softmax(q @ k.transpose(-1, -2) * (q.shape[-1] ** -0.5), dim=-1)
    ctx = (att @ v).transpose(0, 1).reshape(n, -1)
    msg = _lin(w, p + "out_proj", ctx)
    return x + _ffn(w, p, x, msg)


def cross_block(w, i, x0, x1):
    p = "transformers.%d.cross_attn." % i
    def heads(t):
        return t.unflatten(-1, (HEADS, -1)).transpose(0, 1)
    qk0, qk1 = heads(_lin(w, p + "to_qk", x0)), heads(_lin(w, p + "to_qk", x1))
    v0, v1 = heads(_lin(w, p + "to_v", x0)), heads(_lin(w, p + "to_v", x1))
    s = qk0.shape[-1] ** -0.25
    sim = (qk0 * s) @ (qk1 * s).transpose(-1, -2)
    m0 = F.softmax(sim, dim=-1) @ v1
    m1 = F.softmax(sim.transpose(-1, -2), dim=-1) @ v0
    m0 = _lin(w, p + "to_out", m0.transpose(0, 1).reshape(x0.shape[0], -1))
    m1 = _lin(w, p + "to_out", m1.transpose(0, 1).reshape(x1.shape[0], -1))
    return x0 + _ffn(w, p, x0, m0), x1 + _ffn(w, p, x1, m1)


def log_assignment(w, x0, x1):
    """MatchAssignment + sigmoid_log_double_softmax (dustbin row/col excluded: never argmax'd)."""
    p = "log_assignment.%d." % (N_LAYERS - 1)
    md0, md1 = _lin(w, p + "final_proj", x0), _lin(w, p + "final_proj", x1)
    d = md0.shape[-1]
    md0, md1 = md0 / d ** 0.25, md1 / d ** 0.25
    sim = md0 @ md1.t()
    z0, z1 = _lin(w, p + "matchability", x0), _lin(w, p + "matchability", x1)
    cert = F.logsigmoid(z0) + F.logsigmoid(z1).t()
    return F.log_softmax(sim, 1) + F.log_softmax(sim.t().contiguous(), 1).t() + cert, sim, z0, z1


def filter_matches(L: np.ndarray, th: float = FILTER_THRESHOLD):
    """filter_matches + LightGlue-ONNX output packing: pairs [i, m0[i]] for valid i ascending; mscores."""
    m, n = L.shape
    if m == 0 or n == 0:
        return np.zeros((0, 2), np.int32), np.zeros((0,), np.float32)
    m0 = L.argmax(1)                     # numpy argmax: first max = lowest index
    m1 = L.argmax(0)
    mx = L[np.arange(m), m0]
    mutual = m1[m0] == np.arange(m)
    ms = np.exp(mx.astype(np.float32))
    valid = mutual & (ms > np.float32(th))
    idx = np.nonzero(valid)[0]
    return np.stack([idx, m0[idx]], 1).astype(np.int32), ms[idx].astype(np.float32)


def lightglue(w, kpts0, kpts1, desc0, desc1, h0, w0, h1, w1, keep=None):
    """a3: lg_matcher(kpts0,kpts1,desc0,desc1,h0,w0,h1,w1) deep_net.cpp:814-1000.
    kpts in pixels [M,2],[N,2]; desc [M,256],[N,256].  Returns matches [K,2] int32 (i0 asc), mscores [K]."""
    k0 = torch.from_numpy(normalize_kpts(np.asarray(kpts0, np.float32), w0, h0))
    k1 = torch.from_numpy(normalize_kpts(np.asarray(kpts1, np.float32), w1, h1))
    x0 = torch.from_numpy(np.ascontiguousarray(desc0, dtype=np.float32))
    x1 = torch.from_numpy(np.ascontiguousarray(desc1, dtype=np.float32))
    with torch.no_grad():
        e0, e1 = posenc(w, k0), posenc(w, k1)
        for i in range(N_LAYERS):
            x0 = self_block(w, i, x0, e0)
            x1 = self_block(w, i, x1, e1)
            x0, x1 = cross_block(w, i, x0, x1)
            if keep is not None:
                keep["x0_%d" % i] = x0.numpy().copy()
                keep["x1_%d" % i] = x1.numpy().copy()
        L, sim, z0, z1 = log_assignment(w, x0, x1)
    Ln = L.numpy()
    if keep is not None:
        keep["L"] = Ln
        keep["sim"] = sim.numpy()
    matches, ms = filter_matches(Ln)
    return matches, ms
