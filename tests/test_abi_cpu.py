"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/dvins_perception.h declares,
and refuses to run without a GPU (no CPU fallback).  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dvins_perception.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from d_vins_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_config_default_and_struct_size():
    from d_vins_b200 import capi
    cfg = capi.default_config()
    assert cfg.struct_size == ctypes.sizeof(capi.DvConfig)
    assert (cfg.height, cfg.width, cfg.max_kpts, cfg.knn_k, cfg.exclude_recent) == (480, 752, 512, 3, 50)
    assert abs(cfg.det_thresh - 0.0005) < 1e-8 and abs(cfg.lg_filter_thresh - 0.1) < 1e-7


def test_no_cpu_fallback():
    import torch
    from d_vins_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.DvError) as ei:
        capi.Engine()
    assert ei.value.status == 7          # DV_ERR_NOGPU


def test_bad_config_rejected_before_touching_cuda():
    from d_vins_b200 import capi
    cfg = capi.default_config()
    cfg.struct_size = 4
    h = ctypes.c_void_p()
    assert capi._lib.dv_create(ctypes.byref(cfg), ctypes.byref(h)) == 1     # DV_ERR_INVALID
    assert b"size mismatch" in capi._lib.dv_last_error()


def test_logger_levels_and_sink():
    """Levelled logger (ilogger.hpp:24-29): failures are logged at level 1 through the installed sink; level 0
    silences them."""
    from d_vins_b200 import capi
    seen = []
    capi.log_set_sink(lambda lvl, msg: seen.append((lvl, msg)))
    try:
        capi.log_set_level(2)
        cfg = capi.default_config()
        cfg.struct_size = 4
        h = ctypes.c_void_p()
        assert capi._lib.dv_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
        assert seen and seen[-1][0] == 1 and "size mismatch" in seen[-1][1]
        n = len(seen)
        capi.log_set_level(0)
        assert capi._lib.dv_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
        assert len(seen) == n and capi._lib.dv_log_get_level() == 0
    finally:
        capi.log_set_level(2)
        capi.log_set_sink(None)


def test_product_never_imports_oracle():
    """The product package must not import, call or link anything under oracle/."""
    pkg = os.path.join(ROOT, "d_vins_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/dvins_perception.h must compile as C99 (no C++ / torch / CUDA types)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "dvins_perception.h"\nint main(void) { dv_config c; dv_config_default(&c); return (int)c.struct_size == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_shim_compiles_against_the_header(tmp_path):
    """The reference-facing C++ facade (same class / member names as deep_net.h) builds with g++ alone - no CUDA, no
    torch - and links against the C-ABI library."""
    import subprocess
    from d_vins_b200 import build
    lib = build.build()
    obj = tmp_path / "shim.o"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-c", os.path.join(ROOT, "d_vins_b200", "csrc", "shim", "deep_net_shim.cpp"),
                        "-I", os.path.join(ROOT, "include"), "-o", str(obj)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = build.build_shim_demo()
    assert os.path.exists(exe) and os.path.exists(lib)


def test_shim_opencv_overloads_and_stream_driver_compile(tmp_path):
    """The cv::Mat / cv::Point2f overloads (DV_SHIM_WITH_OPENCV) compile against a minimal stand-in for
    <opencv2/core.hpp> (tests/cpp/opencv_stub; no OpenCV C++ headers in this image), and the ROS-free stream driver
    (csrc/shim/loop_closure.cpp) builds and links against the C-ABI library with g++ alone."""
    import subprocess
    from d_vins_b200 import build
    src = tmp_path / "cvuse.cpp"
    src.write_text('''
#include "deep_net_shim.h"
int use(Estimator_net::Estimator* e, MixVPR_net::MixVPR* m, const cv::Mat& img, std::vector<cv::Point2f>& pts) {
  e->sp_extractor(img); e->sp_extractor(img, pts); e->lg_matcher(); m->mix_extractor(img); m->test_in_dataset("/tmp");
  return (int)e->sp_kpts.size() + (int)m->sim_map.size() + e->image.rows;
}
''')
    inc = ["-I", os.path.join(ROOT, "tests", "cpp", "opencv_stub"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "d_vins_b200", "csrc", "shim")]
    for unit in (str(src), os.path.join(ROOT, "d_vins_b200", "csrc", "shim", "deep_net_shim.cpp")):
        r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-DDV_SHIM_WITH_OPENCV", "-c", unit, "-o",
                            str(tmp_path / "o.o")] + inc, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    assert os.path.exists(build.build_stream_demo())


def test_checkpoint_converter_contract(tmp_path):
    """d_vins_b200.convert_weights maps upstream state_dict keys to the DVWGT001 file the engine loads: its key / shape
    contract equals the synthetic weight set every test runs on, a torch-saved state_dict (plain and Lightning-style)
    converts to a byte-identical tensor set, and a missing tensor is an error."""
    import numpy as np
    import torch
    from d_vins_b200 import convert_weights as cw
    from oracle import weights
    W = weights.synth_all()
    want = cw.expected_keys()
    assert set(want) == set(W) and all(tuple(W[k].shape) == want[k] for k in want)
    sp = {k: torch.from_numpy(v) for k, v in weights.sub(W, "sp.").items()}
    lg = {k: torch.from_numpy(v) for k, v in weights.sub(W, "lg.").items()}
    lg["transformers.0.token_confidence.0.weight"] = torch.zeros(1, 256)        # unused upstream tensors are ignored
    mix = {"state_dict": {k: torch.from_numpy(v) for k, v in weights.sub(W, "mix.").items()}}
    torch.save(sp, tmp_path / "sp.pth"); torch.save(lg, tmp_path / "lg.pth"); torch.save(mix, tmp_path / "mix.ckpt")
    out = cw.convert(str(tmp_path / "sp.pth"), str(tmp_path / "lg.pth"), str(tmp_path / "mix.ckpt"))
    cw.save_dvw(str(tmp_path / "w.dvw"), out)
    back = weights.load_weights(str(tmp_path / "w.dvw"))
    assert list(back) == list(want) and all(np.array_equal(back[k], W[k]) for k in W)
    del sp["convDb.bias"]
    torch.save(sp, tmp_path / "bad.pth")
    with pytest.raises(KeyError):
        cw.convert(superpoint=str(tmp_path / "bad.pth"))


def test_every_environment_switch_is_documented():
    """DESIGN.md §4's A/B switch table must name every DV_* variable the library reads (getenv in csrc/)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for dirpath, _, files in os.walk(os.path.join(root, "d_vins_b200", "csrc")):
        for f in files:
            if f.endswith((".cu", ".cpp", ".cuh", ".h")):
                names |= set(re.findall(r'getenv\("(DV_[A-Z0-9_]+)"\)', open(os.path.join(dirpath, f)).read()))
    design = open(os.path.join(root, "DESIGN.md")).read()
    missing = sorted(n for n in names if n not in design)
    assert len(names) > 20 and not missing, missing
