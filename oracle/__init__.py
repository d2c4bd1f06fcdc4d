"""CPU oracle for the D_VINS loop-closure perception path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
The product path (``d_vins_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Every function cites the reference file:line (relative to the upstream
kajo-kurisu/D_VINS tree) whose arithmetic it restates.

Parity pinning status:
  * SuperPoint / SP_RE  : PINNED against the reference's own ``export/superpoint.py``
                          and ``export/ultrapoint.py`` modules imported in the build
                          container (``tests/golden/make_golden.py``); outputs committed as
                          ``tests/golden/sp_*.npz``.
  * pre-processing      : restated from ``preprocess_kernel.cu`` (CUDA source, cannot run
                          here) -> parity unpinned.
  * LightGlue / MixVPR  : arithmetic lives in un-vendored third-party projects
                          (cvg/LightGlue v0.1_arxiv via fabio-sim/LightGlue-ONNX v0.1.3;
                          amaralibey/MixVPR; torchvision ResNet-50).  The ResNet-50 trunk is
                          pinned against torchvision (installed here); the rest is restated from
                          the published architectures -> parity unpinned.
  * kNN                 : faiss 1.7.2 IndexFlatIP (not installable here) -> restated, parity
                          unpinned (exact inner-product top-k is unambiguous up to tie order).
"""
