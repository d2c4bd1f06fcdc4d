"""Generate golden vectors by running the REFERENCE's own Python modules (export/superpoint.py,
export/ultrapoint.py) in the build container, on seeded synthetic weights + frames.

Run here (needs /root/reference; it does not exist on the GPU box):   python tests/golden/make_golden.py
Outputs (committed):  tests/golden/sp_euroc.npz, tests/golden/sp_kitti.npz
The hard-coded ``torch.load('/home/sy/...')`` in the reference constructors (superpoint.py:145,
ultrapoint.py:93) is bypassed by patching torch.load to return the synthetic state_dict - the reference
code itself is executed unmodified.  Nothing here is imported by the product.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import weights, synth          # noqa: E402

REF = "/root/reference/export"


def _load_module(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(w, img_u8, vio_xy):
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    real_load = torch.load
    torch.load = lambda *a, **k: sd
    try:
        sp_mod = _load_module("superpoint")
        up_mod = _load_module("ultrapoint")
        sp = sp_mod.SuperPoint(max_num_keypoints=512).eval()
        up = up_mod.UltraPoint().eval()
    finally:
        torch.load = real_load
    # deep_net.cpp:578: the engine input is u8 * (1/255.f)
    x = torch.from_numpy(img_u8.astype(np.float32) * (np.float32(1) / np.float32(255)))[None, None]
    with torch.no_grad():
        kp, sc, de = sp(x)
        _, de_r = up(x, torch.from_numpy(vio_xy)[None])
    return kp[0].numpy(), sc[0].numpy(), de[0].numpy(), de_r[0].numpy()


def main():
    w = weights.synth_superpoint(synth.BASE_SEED)
    for tag, (h, wd) in (("euroc", (480, 752)), ("kitti", (376, 1241))):
        img = synth.make_frame(h, wd, synth.BASE_SEED + (0 if tag == "euroc" else 5))
        vio = synth.vio_points(150 if tag == "euroc" else 200, h, wd, synth.BASE_SEED + 3)
        kp, sc, de, de_r = run_reference(w, img, vio)
        # the reference's topk tie order is unspecified; record whether ties exist among the selected scores
        ties = int(len(sc) - len(np.unique(sc)))
        out = os.path.join(ROOT, "tests", "golden", "sp_%s.npz" % tag)
        np.savez_compressed(out, h=h, w=wd, kpts=kp.astype(np.int32), scores=sc.astype(np.float32),
                            desc_head=de[:48].astype(np.float32), desc_rowsum=de.sum(1).astype(np.float32),
                            vio=vio, desc_r_head=de_r[:48].astype(np.float32),
                            desc_r_rowsum=de_r.sum(1).astype(np.float32), ties=ties)
        print(tag, "kpts", kp.shape, "ties", ties, "->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
