"""Quantisation-aware CPU oracle (TEST INFRASTRUCTURE): the SAME arithmetic as oracle/superpoint.py and
oracle/lightglue.py, evaluated with the operand precisions the B200 engine uses - fp16-rounded weights, fp16-rounded
activations at every point where the engine stores an activation as fp16, fp32 accumulation, fp32 bias / residual /
softmax / LayerNorm - so that ONLY the summation order inside a dot product differs between this oracle and the CUDA
path.  Purpose (SURVEY.md §7 "Hard parts"): the discrete outputs the north star wants bit-exact
(keypoint indices and order, NMS survivors, match pairs) are discontinuous functions of the score / assignment floats;
against the plain fp32 oracle they can only be compared up to the fp16 tolerance, against this oracle they are compared
with np.array_equal.

Rounding points mirror the engine (d_vins_b200/csrc):
  SuperPoint (sp.cu, conv_halo*.cu): frame as u8/256 (exact in fp16) with 256/255 folded into conv1a's fp16 weights
    (conv_halo.cu:107-113); every conv: fp16 weights, fp16 input, fp32 accumulate + fp32 bias, ReLU, (2x2 max-pool),
    store fp16; convPb / convDb outputs stay fp32; softmax / NMS / top-k / sampling in fp32 (identical code paths to
    oracle/superpoint.py, which follows export/superpoint.py:52-224).
  LightGlue (lg.cu, gemm_wres.cu, lg_attn.cu): fp32 residual stream x32 with an fp16 copy feeding every GEMM; qkv
    stored fp16 after the rotary (rotary table itself stored as fp16 cos/sin); attention probabilities rounded to fp16
    before P.V while the row sum uses the unrounded values; ctx / msg / FFN hidden / GELU output stored fp16;
    final_proj weights pre-scaled by 1/4 and rounded to fp16, md stored fp16; similarity, log-softmax, matchability fp32.
    out_proj / to_out are composed with the first FFN linear offline (W0b Wout in float64, then ONE fp16 rounding), so the
    attention context feeds the FFN directly (lg.cu fold_out_proj).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import lightglue as olg
from . import superpoint as osp


def h16(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to fp16, back to fp32"""
    return t.to(torch.float16).to(torch.float32)


def _w16(w, k):
    return h16(torch.from_numpy(np.ascontiguousarray(w[k])))


def _b(w, k):
    return torch.from_numpy(np.ascontiguousarray(w[k]))


# ------------------------------------------------------------------------------------------------ SuperPoint
def encoder_q(w, img_u8: np.ndarray, keep=None) -> torch.Tensor:
    assert img_u8.ndim == 2, "quantisation-aware path: 1-channel frames (conv1a on the tensor cores)"
    x = torch.from_numpy(img_u8.astype(np.float32) * np.float32(1.0 / 256.0))[None, None]      # exact in fp16
    w1a = h16(_b(w, "conv1a.weight") * (np.float32(256.0) / np.float32(255.0)))

    def cr(x, wt, name, pool=False):
        y = F.relu(F.conv2d(x, wt, _b(w, name + ".bias"), padding=1))
        if pool:
            y = F.max_pool2d(y, 2, 2)
        y = h16(y)
        if keep is not None:
            keep[name + ("_pool" if pool else "")] = y
        return y
    x = cr(x, w1a, "conv1a")
    x = cr(x, _w16(w, "conv1b.weight"), "conv1b", pool=True)
    x = cr(x, _w16(w, "conv2a.weight"), "conv2a")
    x = cr(x, _w16(w, "conv2b.weight"), "conv2b", pool=True)
    x = cr(x, _w16(w, "conv3a.weight"), "conv3a")
    x = cr(x, _w16(w, "conv3b.weight"), "conv3b", pool=True)
    x = cr(x, _w16(w, "conv4a.weight"), "conv4a")
    x = cr(x, _w16(w, "conv4b.weight"), "conv4b")
    return x


def heads_q(w, feat: torch.Tensor, keep=None):
    cPa = h16(F.relu(F.conv2d(feat, _w16(w, "convPa.weight"), _b(w, "convPa.bias"), padding=1)))
    logits = F.conv2d(cPa, _w16(w, "convPb.weight"), _b(w, "convPb.bias"))
    cDa = h16(F.relu(F.conv2d(feat, _w16(w, "convDa.weight"), _b(w, "convDa.bias"), padding=1)))
    d = F.conv2d(cDa, _w16(w, "convDb.weight"), _b(w, "convDb.bias"))
    if keep is not None:
        keep["convPa"] = cPa
        keep["convDa"] = cDa
        keep["logits"] = logits
        keep["convDb"] = d
    s = F.softmax(logits, 1)[:, :-1]
    _, _, h, wd = s.shape
    s = s.permute(0, 2, 3, 1).reshape(1, h, wd, 8, 8).permute(0, 1, 3, 2, 4).reshape(1, h * 8, wd * 8)
    return s[0], F.normalize(d, p=2, dim=1)


def superpoint_q(w, img_u8: np.ndarray, max_kpts: int = osp.MAX_KPTS, keep=None):
    """oracle.superpoint.superpoint with the engine's operand precisions (see module docstring)."""
    H, W = img_u8.shape[:2]
    with torch.no_grad():
        feat = encoder_q(w, img_u8, keep)
        smap, dmap = heads_q(w, feat, keep)
        nms = osp.simple_nms(smap)
        kp, sc, lin = osp.select_keypoints(nms, max_kpts)
        desc = osp.sample_descriptors(kp[None], dmap)
    kp_np = kp.numpy().astype(np.int32)
    return {"kpts": kp_np, "scores": sc.numpy().astype(np.float32), "desc": desc.numpy(),
            "kpts_norm": osp.normalize_kpts(kp_np, W, H), "score_map": smap.numpy(), "nms": nms.numpy(),
            "feat": feat, "dmap": dmap}


def superpoint_recover_q(w, img_u8, kpts_xy, dmap=None):
    with torch.no_grad():
        if dmap is None:
            _, dmap = heads_q(w, encoder_q(w, img_u8))
        desc = osp.sample_descriptors(torch.from_numpy(kpts_xy.astype(np.float32))[None], dmap)
    return desc.numpy()


# ------------------------------------------------------------------------------------------------ LightGlue
def _lin16(w, name, x16, scale=None):
    W = _b(w, name + ".weight"); b = _b(w, name + ".bias")
    if scale is not None:
        W = W * scale; b = b * scale
    return F.linear(x16, h16(W), b)


def _attend(q16, k16, v16, scale):
    """q16 [h,nq,64], k16/v16 [h,nk,64] (fp16-valued): base-2 softmax as lg_attn.cu - P rounded to fp16 for P.V, the row
    sum from the unrounded P, output rounded to fp16."""
    s = q16 @ k16.transpose(-1, -2)
    sl2 = np.float32(scale * 1.4426950408889634)
    m = s.max(dim=-1, keepdim=True).values
    p = torch.exp2(s * sl2 - m * sl2)
    o = h16(p) @ v16
    return h16(o / p.sum(dim=-1, keepdim=True))


def _folded_ffn0(w, p, out_name):
    """[W0a | W0b Wout], b0 + W0b bout in float64 (lg.cu fold_out_proj), then the engine's fp16 weight rounding."""
    W0 = w[p + "ffn.0.weight"].astype(np.float64); b0 = w[p + "ffn.0.bias"].astype(np.float64)
    Wo = w[p + out_name + ".weight"].astype(np.float64); bo = w[p + out_name + ".bias"].astype(np.float64)
    Wf = np.concatenate([W0[:, :256], W0[:, 256:] @ Wo], 1).astype(np.float32)
    bf = (b0 + W0[:, 256:] @ bo).astype(np.float32)
    return h16(torch.from_numpy(Wf)), torch.from_numpy(bf)


def _ffn_q(w, p, x32, x16, msg16, folded=None):
    if folded is not None:          # msg16 is the attention context; out_proj lives inside the folded weights
        h = h16(F.linear(torch.cat([x16, msg16], -1), folded[0], folded[1]))
    else:
        h = h16(_lin16(w, p + "ffn.0", torch.cat([x16, msg16], -1)))
    h = F.layer_norm(h, (h.shape[-1],), _b(w, p + "ffn.1.weight"), _b(w, p + "ffn.1.bias"), eps=1e-5)
    g = h16(F.gelu(h))
    x32 = (F.linear(g, _w16(w, p + "ffn.3.weight")) + _b(w, p + "ffn.3.bias")) + x32
    return x32, h16(x32)


def _heads(t):
    return t.unflatten(-1, (olg.HEADS, -1)).transpose(0, 1)


def self_block_q(w, i, x32, x16, enc16, fold_out=True):
    p = "transformers.%d.self_attn." % i
    n = x16.shape[0]
    qkv = _lin16(w, p + "Wqkv", x16).unflatten(-1, (olg.HEADS, -1, 3)).transpose(0, 1)   # [h,n,64,3] fp32
    q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]

    def rope(t):   # gemm_wres.cu epilogue: x0' = x0 c - x1 s ; x1' = x1 c + x0 s with the fp16 table, in fp32
        c, s = enc16
        t2 = t.unflatten(-1, (-1, 2))
        x0, x1 = t2[..., 0], t2[..., 1]
        return torch.stack((x0 * c[None] - x1 * s[None], x1 * c[None] + x0 * s[None]), dim=-1).flatten(start_dim=-2)
    q16, k16, v16 = h16(rope(q)), h16(rope(k)), h16(v)
    ctx16 = _attend(q16, k16, v16, 0.125).transpose(0, 1).reshape(n, -1)
    if fold_out:
        return _ffn_q(w, p, x32, x16, ctx16, _folded_ffn0(w, p, "out_proj"))
    msg16 = h16(_lin16(w, p + "out_proj", ctx16))
    return _ffn_q(w, p, x32, x16, msg16)


def cross_block_q(w, i, x0, x1, fold_out=True):
    p = "transformers.%d.cross_attn." % i
    (x0_32, x0_16), (x1_32, x1_16) = x0, x1
    qk0, qk1 = _heads(h16(_lin16(w, p + "to_qk", x0_16))), _heads(h16(_lin16(w, p + "to_qk", x1_16)))
    v0, v1 = _heads(h16(_lin16(w, p + "to_v", x0_16))), _heads(h16(_lin16(w, p + "to_v", x1_16)))
    m0 = _attend(qk0, qk1, v1, 0.125).transpose(0, 1).reshape(x0_16.shape[0], -1)
    m1 = _attend(qk1, qk0, v0, 0.125).transpose(0, 1).reshape(x1_16.shape[0], -1)
    if fold_out:
        fw = _folded_ffn0(w, p, "to_out")
        return _ffn_q(w, p, x0_32, x0_16, m0, fw), _ffn_q(w, p, x1_32, x1_16, m1, fw)
    m0 = h16(_lin16(w, p + "to_out", m0))
    m1 = h16(_lin16(w, p + "to_out", m1))
    return _ffn_q(w, p, x0_32, x0_16, m0), _ffn_q(w, p, x1_32, x1_16, m1)


def lightglue_q(w, kpts0, kpts1, desc0, desc1, h0, w0, h1, w1, keep=None, fold_out=True):
    """oracle.lightglue.lightglue with the engine's operand precisions (see module docstring).  fold_out mirrors the
    engine default (DV_LG_FOLD_OUT): out_proj / to_out composed with the FFN's first linear before the fp16 rounding."""
    k0 = torch.from_numpy(olg.normalize_kpts(np.asarray(kpts0, np.float32), w0, h0))
    k1 = torch.from_numpy(olg.normalize_kpts(np.asarray(kpts1, np.float32), w1, h1))
    x0 = torch.from_numpy(np.ascontiguousarray(desc0, dtype=np.float32))
    x1 = torch.from_numpy(np.ascontiguousarray(desc1, dtype=np.float32))
    with torch.no_grad():
        def enc16(k):
            proj = F.linear(k, _b(w, "posenc.Wr.weight"))
            return h16(torch.cos(proj)), h16(torch.sin(proj))      # [n,32] (cos_j, sin_j), rope16 in lg.cu
        e0, e1 = enc16(k0), enc16(k1)
        x0, x1 = (x0, h16(x0)), (x1, h16(x1))
        for i in range(olg.N_LAYERS):
            x0 = self_block_q(w, i, x0[0], x0[1], e0, fold_out)
            x1 = self_block_q(w, i, x1[0], x1[1], e1, fold_out)
            x0, x1 = cross_block_q(w, i, x0, x1, fold_out)
        p = "log_assignment.%d." % (olg.N_LAYERS - 1)
        md0 = h16(_lin16(w, p + "final_proj", x0[1], scale=0.25))
        md1 = h16(_lin16(w, p + "final_proj", x1[1], scale=0.25))
        sim = md0 @ md1.t()
        z0 = F.linear(x0[0], _b(w, p + "matchability.weight"), _b(w, p + "matchability.bias"))
        z1 = F.linear(x1[0], _b(w, p + "matchability.weight"), _b(w, p + "matchability.bias"))
        rl = torch.logsumexp(sim, 1, keepdim=True)
        cl = torch.logsumexp(sim, 0, keepdim=True)
        L = ((sim - rl) + (sim - cl)) + (F.logsigmoid(z0) + F.logsigmoid(z1).t())
    Ln = L.numpy()
    if keep is not None:
        keep["L"] = Ln
        keep["sim"] = sim.numpy()
    return olg.filter_matches(Ln)
