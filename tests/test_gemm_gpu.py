"""tcgen05 GEMM / implicit-GEMM conv3x3 vs an fp32 reference of the same op (through the C ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from d_vins_b200 import capi
    e = capi.Engine(height=64, width=64)      # no weights: stage-level entry points only
    yield e
    e.close()


def _q(a):      # operands are rounded to fp16 on the device; compare against the same rounding
    return a.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 256), (300, 80, 256), (1000, 400, 400),
                                   (5640, 256, 1152), (77, 768, 256), (4096, 64, 576)])
def test_gemm(eng, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    D = eng.dbg_gemm(A, B, bias, relu=True)
    ref = np.maximum(_q(A).astype(np.float64) @ _q(B).astype(np.float64).T + bias, 0)
    err = np.abs(D - ref).max()
    assert err < 2e-3, err     # fp32 accumulation of exactly-representable fp16 products: order-of-summation only


@pytest.mark.parametrize("n,h,w,cin,cout,pool", [(1, 16, 16, 64, 64, False), (2, 24, 40, 64, 64, True),
                                                 (1, 30, 47, 128, 128, False), (1, 60, 94, 128, 256, False),
                                                 (2, 17, 23, 64, 128, True), (1, 120, 188, 64, 64, True)])
def test_conv3x3(eng, n, h, w, cin, cout, pool):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(h * w + cin)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
    bias = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    y = eng.dbg_conv3x3(x, wt, bias, relu=True, pool=pool)
    xt = torch.from_numpy(_q(x)).permute(0, 3, 1, 2).double()
    r = F.relu(F.conv2d(xt, torch.from_numpy(_q(wt)).double(), torch.from_numpy(bias).double(), padding=1))
    if pool:
        r = F.max_pool2d(r, 2, 2)
    ref = r.permute(0, 2, 3, 1).numpy()
    assert y.shape == ref.shape
    err = np.abs(y - ref).max()
    assert err < 4e-3, err     # output stored as fp16 (rel 2^-11 of values up to ~4)


@pytest.mark.parametrize("n,h,w,pool,blocked", [(1, 16, 8, False, True), (1, 32, 24, True, True), (2, 48, 40, False, False),
                                               (1, 37, 29, True, False), (3, 120, 188, True, True), (1, 480, 752, True, True)])
def test_conv3x3_halo64(eng, n, h, w, pool, blocked):
    """Weights-stationary halo-tile kernel (conv_halo.cu): one TMA halo box per 16x8 tile, nine shifted no-swizzle views."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(h * w + n)
    x = rng.standard_normal((n, h, w, 64)).astype(np.float32)
    wt = (rng.standard_normal((64, 64, 3, 3)) * np.sqrt(2.0 / 576)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(64)).astype(np.float32)
    y = eng.dbg_conv3x3_halo64(x, wt, bias, relu=True, pool=pool, out_blocked=blocked)
    xt = torch.from_numpy(_q(x)).permute(0, 3, 1, 2).double()
    r = F.relu(F.conv2d(xt, torch.from_numpy(_q(wt)).double(), torch.from_numpy(bias).double(), padding=1))
    if pool:
        r = F.max_pool2d(r, 2, 2)
    ref = r.permute(0, 2, 3, 1).numpy()
    assert y.shape == ref.shape
    err = np.abs(y - ref).max()
    assert err < 4e-3, err


@pytest.mark.parametrize("n,h,w,cin,cout,pool,blocked", [
    (1, 32, 8, 128, 128, False, True), (1, 32, 8, 64, 128, False, False), (2, 40, 24, 128, 128, True, True),
    (1, 37, 29, 128, 128, True, False), (1, 37, 29, 64, 128, False, True), (2, 60, 94, 128, 512, False, False),
    (3, 120, 188, 128, 128, True, True), (2, 47, 155, 128, 256, False, True),
    # CTA-pair kernel with an ODD tile count (the odd CTA of the last pair has no tile), one / three images
    (1, 16, 40, 128, 128, False, False), (3, 32, 24, 64, 128, True, True), (3, 32, 24, 128, 256, False, False)])
def test_conv3x3_halo128(eng, n, h, w, cin, cout, pool, blocked):
    """256-pixel halo-tile kernel (conv_halo128.cu): halo fetched once per 32x8 tile, weights streamed, two M=128 row
    blocks per weight block; ragged tiles, odd sizes (pool floors), several 128-column output chunks."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(h * w + n + cin + cout)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
    bias = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    y = eng.dbg_conv3x3_halo128(x, wt, bias, relu=True, pool=pool, out_blocked=blocked)
    xt = torch.from_numpy(_q(x)).permute(0, 3, 1, 2).double()
    r = F.relu(F.conv2d(xt, torch.from_numpy(_q(wt)).double(), torch.from_numpy(bias).double(), padding=1))
    if pool:
        r = F.max_pool2d(r, 2, 2)
    ref = r.permute(0, 2, 3, 1).numpy()
    assert y.shape == ref.shape
    err = np.abs(y - ref).max()
    assert err < 4e-3, err


@pytest.mark.parametrize("M,N,K", [(1000, 256, 512), (300, 512, 256), (77, 768, 256), (5000, 192, 64),
                                   (129, 136, 128), (4099, 1024, 320),
                                   # weights-resident kernel (gemm_wres.cu): 256-column slabs (K <= 256), 128-column
                                   # slabs (K = 512), several row tiles per CTA, ragged last tile
                                   (2000, 768, 256), (20001, 256, 256), (1500, 512, 512), (2085, 1024, 128),
                                   # CTA-pair kernel (gemm_pair.cu, cta_group::2): 256-row super-tiles, ragged M that ends
                                   # in the first / second CTA's half, one and several slabs, K = 512 (three A stages)
                                   (4000, 512, 512), (5249, 256, 512), (33010, 768, 256), (2048, 256, 64)])
@pytest.mark.parametrize("mode", ["o16", "o32", "res32_o32_o16", "res16_relu_o16", "res32_o16"])
def test_gemm_fused_epilogues(eng, M, N, K, mode):
    """Every epilogue operand of the staged (TMA in / TMA out) kernel: fp16 / fp32 outputs, in-place fp32 residual,
    fp16 residual + ReLU; ragged M, N not a multiple of the 128 / 64 / 32-column boxes."""
    rng = np.random.default_rng(M * 7 + N * 3 + K + len(mode))
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32) if "res" in mode else None
    r16 = "res16" in mode
    relu = "relu" in mode
    D32, D16 = eng.dbg_gemm_ex(A, B, bias, res=res, res_is_f16=r16, relu=relu, want32="o32" in mode,
                               want16="o16" in mode)
    ref = _q(A).astype(np.float64) @ _q(B).astype(np.float64).T + bias
    if res is not None:
        ref = ref + (_q(res) if r16 else res)
    if relu:
        ref = np.maximum(ref, 0)
    if D32 is not None:
        assert np.abs(D32 - ref).max() < 2e-3
    if D16 is not None:
        # one fp16 rounding of the fp32 result
        assert np.abs(D16 - ref).max() < 2e-3 + np.abs(ref).max() * 2.0 ** -10
