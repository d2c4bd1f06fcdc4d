"""ctypes binding of include/dvins_perception.h (Python host side; no torch types cross the boundary).

The library is REQUIRED: importing this module raises if libdvins_b200.so is missing, and
``Engine(...)`` raises ``DvError`` (DV_ERR_NOGPU) when no sm_100 GPU is visible - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdvins_b200.so")

DESC_DIM = 256
GLOBAL_DIM = 512

STATUS = {0: "DV_OK", 1: "DV_ERR_INVALID", 2: "DV_ERR_CUDA", 3: "DV_ERR_UNSUPPORTED", 4: "DV_ERR_WEIGHTS",
          5: "DV_ERR_CAPACITY", 6: "DV_ERR_COMM", 7: "DV_ERR_NOGPU"}


class DvConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("max_batch", C.c_int32), ("max_kpts", C.c_int32), ("nms_radius", C.c_int32),
                ("det_thresh", C.c_float), ("border", C.c_int32), ("max_vio", C.c_int32), ("knn_k", C.c_int32),
                ("exclude_recent", C.c_int32), ("lg_filter_thresh", C.c_float), ("lg_max_kpts", C.c_int32),
                ("bank_capacity", C.c_int64), ("store_capacity", C.c_int32), ("world_size", C.c_int32),
                ("rank", C.c_int32), ("weights_path", C.c_char_p)]


class DvLoopParams(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("min_loop_num", C.c_int32), ("ransac_iters", C.c_int32),
                ("min_frame_index", C.c_int32), ("pnp_inflation", C.c_double), ("max_theta_diff", C.c_double),
                ("max_pose_diff", C.c_double), ("loop_top_thres", C.c_double), ("loop_back_thres", C.c_double),
                ("qic", C.c_double * 9), ("tic", C.c_double * 3), ("seed", C.c_uint64)]


class DvLoopResult(C.Structure):
    _fields_ = [("has_loop", C.c_int32), ("n_inliers", C.c_int32), ("pnp_t_old", C.c_double * 3),
                ("pnp_r_old", C.c_double * 9), ("relative_t", C.c_double * 3), ("relative_q", C.c_double * 4),
                ("relative_yaw", C.c_double)]


class DvError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS.get(status, status), msg))
        self.status = status


def load_library():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libdvins_b200.so not built (run `python -m d_vins_b200.build`); no fallback exists")
    return C.CDLL(LIB_PATH)


_lib = load_library()
_lib.dv_last_error.restype = C.c_char_p
_lib.dv_version.restype = C.c_char_p
_lib.dv_config_default.restype = None
LOG_SINK = C.CFUNCTYPE(None, C.c_int32, C.c_char_p, C.c_void_p)
_lib.dv_log_set_level.restype = None
_lib.dv_log_set_level.argtypes = [C.c_int32]
_lib.dv_log_get_level.restype = C.c_int32
_lib.dv_log_set_sink.restype = None
_lib.dv_log_set_sink.argtypes = [LOG_SINK, C.c_void_p]
_sink_keepalive = [None]


def log_set_level(level: int) -> None:
    """0 off, 1 error, 2 warning (default), 3 info, 4 debug (ilogger.hpp:24-29 levels, minus FATAL's abort)."""
    _lib.dv_log_set_level(int(level))


def log_set_sink(fn) -> None:
    """fn(level, message) receives every log line instead of stderr; None restores stderr."""
    if fn is None:
        _lib.dv_log_set_sink(C.cast(None, LOG_SINK), None)
        _sink_keepalive[0] = None
        return
    cb = LOG_SINK(lambda lvl, msg, user: fn(int(lvl), (msg or b"").decode(errors="replace")))
    _sink_keepalive[0] = cb
    _lib.dv_log_set_sink(cb, None)


def _ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def _chk(rc):
    if rc != 0:
        raise DvError(rc, (_lib.dv_last_error() or b"").decode())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def default_config(**kw) -> DvConfig:
    cfg = DvConfig()
    _lib.dv_config_default(C.byref(cfg))
    for k, v in kw.items():
        if k == "weights_path" and v is not None:
            v = os.fsencode(v)
        setattr(cfg, k, v)
    return cfg


def loop_params(qic=None, tic=None, **kw) -> DvLoopParams:
    p = DvLoopParams()
    _lib.dv_loop_params_default(C.byref(p))
    if qic is not None:
        p.qic = (C.c_double * 9)(*np.asarray(qic, np.float64).reshape(9))
    if tic is not None:
        p.tic = (C.c_double * 3)(*np.asarray(tic, np.float64).reshape(3))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def detect_loop(params: DvLoopParams, top_sim, top_sim_index, frame_index) -> int:
    """PoseGraph::detectLoop (pose_graph.cpp:451-509) on one keyframe's kNN result."""
    d = np.ascontiguousarray(top_sim, np.float32); i = np.ascontiguousarray(top_sim_index, np.int64)
    _lib.dv_detect_loop.restype = C.c_int64
    return int(_lib.dv_detect_loop(C.byref(params), _ptr(d, C.c_float), _ptr(i, C.c_int64), int(d.shape[0]),
                                   C.c_int64(int(frame_index))))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(_lib.dv_comm_unique_id(buf))
    return buf.raw


class Engine:
    """Owns one dv_engine.  Methods mirror the C entry points 1:1 with numpy arrays."""

    def __init__(self, **kw):
        self.cfg = default_config(**kw)
        self._h = C.c_void_p()
        _chk(_lib.dv_create(C.byref(self.cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.dv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- per-keyframe API
    def frame_upload(self, img: np.ndarray):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        _chk(_lib.dv_frame_upload(self._h, _ptr(img, C.c_uint8), h, w, w * ch, ch))

    def sp_detect(self):
        k = self.cfg.max_kpts
        kp = np.zeros((k, 2), np.int32); sc = np.zeros((k,), np.float32)
        de = np.zeros((k, DESC_DIM), np.float32); kn = np.zeros((k, 2), np.float32)
        n = C.c_int32(0)
        _chk(_lib.dv_sp_detect(self._h, _ptr(kp, C.c_int32), _ptr(sc, C.c_float), _ptr(de, C.c_float),
                               _ptr(kn, C.c_float), C.byref(n)))
        n = n.value
        return {"kpts": kp[:n], "scores": sc[:n], "desc": de[:n], "kpts_norm": kn[:n]}

    def sp_describe(self, kpts_xy: np.ndarray):
        k = _f32(kpts_xy)
        n = k.shape[0]
        de = np.zeros((n, DESC_DIM), np.float32)
        _chk(_lib.dv_sp_describe(self._h, _ptr(k, C.c_float), n, _ptr(de, C.c_float)))
        return de

    def mix_describe(self):
        d = np.zeros((GLOBAL_DIM,), np.float32)
        _chk(_lib.dv_mix_describe(self._h, _ptr(d, C.c_float)))
        return d

    def bank_append(self, des):
        d = _f32(des); row = C.c_int64(-1)
        _chk(_lib.dv_bank_append(self._h, _ptr(d, C.c_float), C.byref(row)))
        return row.value

    def bank_size(self):
        r = C.c_int64(0)
        _chk(_lib.dv_bank_size(self._h, C.byref(r)))
        return r.value

    def bank_import(self, bank):
        b = _f32(bank)
        _chk(_lib.dv_bank_import(self._h, _ptr(b, C.c_float), C.c_int64(b.shape[0])))

    def bank_export(self):
        n = self.bank_size()
        out = np.zeros((max(n, 1), GLOBAL_DIM), np.float32); rows = C.c_int64(0)
        _chk(_lib.dv_bank_export(self._h, _ptr(out, C.c_float), C.c_int64(out.shape[0]), C.byref(rows)))
        return out[:rows.value]

    def bank_search(self, q, nb_limit, k=None):
        k = k or self.cfg.knn_k
        q = _f32(q); D = np.zeros((k,), np.float32); I = np.zeros((k,), np.int64)
        _chk(_lib.dv_bank_search(self._h, _ptr(q, C.c_float), C.c_int64(nb_limit), k, _ptr(D, C.c_float),
                                 _ptr(I, C.c_int64)))
        return D, I

    def lg_match(self, kpts0, kpts1, desc0, desc1, h0, w0, h1, w1, want_mkpts=False):
        k0, k1, d0, d1 = _f32(kpts0), _f32(kpts1), _f32(desc0), _f32(desc1)
        m, n = k0.shape[0], k1.shape[0]
        cap = max(1, min(m, n))
        ma = np.zeros((cap, 2), np.int32); ms = np.zeros((cap,), np.float32)
        mk0 = np.zeros((cap, 2), np.float32) if want_mkpts else None
        mk1 = np.zeros((cap, 2), np.float32) if want_mkpts else None
        ko = C.c_int32(0)
        _chk(_lib.dv_lg_match(self._h, _ptr(k0, C.c_float), m, _ptr(k1, C.c_float), n, _ptr(d0, C.c_float),
                              _ptr(d1, C.c_float), h0, w0, h1, w1, _ptr(ma, C.c_int32), _ptr(ms, C.c_float),
                              _ptr(mk0, C.c_float), _ptr(mk1, C.c_float), C.byref(ko)))
        k = ko.value
        if want_mkpts:
            return ma[:k], ms[:k], mk0[:k], mk1[:k]
        return ma[:k], ms[:k]

    # ---------------------------------------------------------------- batched API
    def batch_upload(self, imgs: np.ndarray):
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
        b, h, w = imgs.shape[:3]
        _chk(_lib.dv_batch_upload(self._h, b, _ptr(imgs, C.c_uint8), C.c_int64(h * w), w))

    def batch_upload_ptr(self, b, ptr, frame_stride, stride):
        """Raw-pointer variant (e.g. a pinned torch tensor's data_ptr()) - avoids numpy copies in timed loops."""
        _chk(_lib.dv_batch_upload(self._h, b, C.cast(ptr, C.POINTER(C.c_uint8)), C.c_int64(frame_stride), stride))

    def batch_extract(self, vio_xy, n_vio, frame_ids):
        v = _f32(vio_xy); nv = np.ascontiguousarray(n_vio, dtype=np.int32)
        ids = np.ascontiguousarray(frame_ids, dtype=np.int64)
        b = ids.shape[0]
        assert v.shape == (b, self.cfg.max_vio, 2), v.shape
        _chk(_lib.dv_batch_extract(self._h, b, _ptr(v, C.c_float), _ptr(nv, C.c_int32), _ptr(ids, C.c_int64)))

    def batch_commit(self, b):
        r = C.c_int64(-1)
        _chk(_lib.dv_batch_commit(self._h, b, C.byref(r)))
        return r.value

    def batch_search(self, nb_limit=None, b=None):
        """nb_limit [b] explicit windows, or None (+ b): the engine applies keyframe.cpp:274-282 with cfg.exclude_recent
        to the bank rows the last batch_commit assigned."""
        nb = None if nb_limit is None else np.ascontiguousarray(nb_limit, dtype=np.int64)
        b = nb.shape[0] if nb is not None else int(b)
        k = self.cfg.knn_k
        D = np.zeros((b, k), np.float32); I = np.zeros((b, k), np.int64)
        _chk(_lib.dv_batch_search(self._h, b, _ptr(nb, C.c_int64), _ptr(D, C.c_float), _ptr(I, C.c_int64)))
        return D, I

    def batch_match(self, query_ids, old_ids):
        q = np.ascontiguousarray(query_ids, dtype=np.int64); o = np.ascontiguousarray(old_ids, dtype=np.int64)
        b = q.shape[0]; cap = self.cfg.max_vio
        ma = np.zeros((b, cap, 2), np.int32); ms = np.zeros((b, cap), np.float32); ko = np.zeros((b,), np.int32)
        _chk(_lib.dv_batch_match(self._h, b, _ptr(q, C.c_int64), _ptr(o, C.c_int64), _ptr(ma, C.c_int32),
                                 _ptr(ms, C.c_float), _ptr(ko, C.c_int32)))
        self.last_match_status = ko.copy()      # -1: a keyframe of the pair is not resident on any rank
        return [(ma[i, :max(ko[i], 0)], ms[i, :max(ko[i], 0)]) for i in range(b)]

    def batch_match_ex(self, query_ids, old_ids, query_part=0, old_part=2, cap=None):
        """query_part / old_part: 0 window points, 1 SuperPoint points, 2 all (dvins_perception.h DV_PART_*)."""
        q = np.ascontiguousarray(query_ids, dtype=np.int64); o = np.ascontiguousarray(old_ids, dtype=np.int64)
        b = q.shape[0]; cap = cap or (self.cfg.max_kpts + self.cfg.max_vio)
        ma = np.zeros((b, cap, 2), np.int32); ms = np.zeros((b, cap), np.float32); ko = np.zeros((b,), np.int32)
        _chk(_lib.dv_batch_match_ex(self._h, b, _ptr(q, C.c_int64), _ptr(o, C.c_int64), int(query_part), int(old_part),
                                    int(cap), _ptr(ma, C.c_int32), _ptr(ms, C.c_float), _ptr(ko, C.c_int32)))
        self.last_match_status = ko.copy()
        return [(ma[i, :max(ko[i], 0)], ms[i, :max(ko[i], 0)]) for i in range(b)]

    def batch_match_begin(self, query_ids, old_ids, query_part=0, old_part=2, cap=None):
        """Queue the LightGlue pass of a round and return; collect with batch_match_end() - meanwhile the next round may be
        uploaded and extracted (its kernels queue behind the match)."""
        q = np.ascontiguousarray(query_ids, dtype=np.int64); o = np.ascontiguousarray(old_ids, dtype=np.int64)
        b = q.shape[0]; cap = cap or self.cfg.max_vio
        self._mp = (b, int(cap), q, o)
        _chk(_lib.dv_batch_match_begin(self._h, b, _ptr(q, C.c_int64), _ptr(o, C.c_int64), int(query_part), int(old_part), int(cap)))

    def batch_match_end(self):
        b, cap, _, _ = self._mp
        ma = np.zeros((b, cap, 2), np.int32); ms = np.zeros((b, cap), np.float32); ko = np.zeros((b,), np.int32)
        _chk(_lib.dv_batch_match_end(self._h, _ptr(ma, C.c_int32), _ptr(ms, C.c_float), _ptr(ko, C.c_int32)))
        self._mp = None
        self.last_match_status = ko.copy()
        return [(ma[i, :max(ko[i], 0)], ms[i, :max(ko[i], 0)]) for i in range(b)]

    def batch_match_sp(self, query_ids, old_ids):
        """SuperPoint-vs-SuperPoint pair match (BASELINE config 2)."""
        return self.batch_match_ex(query_ids, old_ids, 1, 1, self.cfg.max_kpts)

    def batch_describe_global(self, b):
        _chk(_lib.dv_batch_describe_global(self._h, int(b)))

    def store_lookup_many(self, frame_ids):
        ids = np.ascontiguousarray(frame_ids, dtype=np.int64)
        out = np.full((ids.shape[0],), -1, np.int32)
        _chk(_lib.dv_store_lookup_many(self._h, int(ids.shape[0]), _ptr(ids, C.c_int64), _ptr(out, C.c_int32)))
        return out

    def verify_loop(self, pts3d, pts2d_norm, vio_R, vio_T, params: DvLoopParams):
        """Batched KeyFrame::PnPRANSAC + findConnection acceptance.  pts3d: list of [n_i,3], pts2d_norm: list of
        [n_i,2], vio_R [b,3,3], vio_T [b,3].  -> list of dicts (status, has_loop, n_inliers, poses)."""
        b = len(pts3d)
        n = np.array([len(x) for x in pts3d], np.int32)
        cap = max(1, int(n.max()))
        X = np.zeros((b, cap, 3), np.float64); U = np.zeros((b, cap, 2), np.float64)
        for i in range(b):
            X[i, :n[i]] = pts3d[i]; U[i, :n[i]] = pts2d_norm[i]
        R = np.ascontiguousarray(vio_R, np.float64).reshape(b, 9); T = np.ascontiguousarray(vio_T, np.float64).reshape(b, 3)
        st = np.zeros((b, cap), np.uint8)
        out = (DvLoopResult * b)()
        _chk(_lib.dv_verify_loop(self._h, b, _ptr(n, C.c_int32), cap, _ptr(X, C.c_double), _ptr(U, C.c_double),
                                 _ptr(R, C.c_double), _ptr(T, C.c_double), C.byref(params), _ptr(st, C.c_uint8), out))
        res = []
        for i in range(b):
            o = out[i]
            res.append(dict(has_loop=bool(o.has_loop), n_inliers=int(o.n_inliers), status=st[i, :n[i]].copy(),
                            pnp_T_old=np.array(o.pnp_t_old), pnp_R_old=np.array(o.pnp_r_old).reshape(3, 3),
                            relative_t=np.array(o.relative_t), relative_q=np.array(o.relative_q),
                            relative_yaw=float(o.relative_yaw)))
        return res

    def store_lookup(self, frame_id):
        """-> (owner_rank or -1, n_total, n_sp)"""
        r = C.c_int32(-1); n = C.c_int32(0); nsp = C.c_int32(0)
        _chk(_lib.dv_store_lookup(self._h, C.c_int64(frame_id), C.byref(r), C.byref(n), C.byref(nsp)))
        return r.value, n.value, nsp.value

    def store_sync(self):
        _chk(_lib.dv_store_sync(self._h))

    def store_read(self, frame_id):
        cap = self.cfg.max_kpts + self.cfg.max_vio
        kp = np.zeros((cap, 2), np.float32); de = np.zeros((cap, DESC_DIM), np.float32)
        n = C.c_int32(0); nsp = C.c_int32(0)
        _chk(_lib.dv_store_read(self._h, C.c_int64(frame_id), _ptr(kp, C.c_float), _ptr(de, C.c_float),
                                C.byref(n), C.byref(nsp)))
        return kp[:n.value], de[:n.value], nsp.value

    def store_put(self, frame_id, kpts_xy, desc, n_sp):
        k, d = _f32(kpts_xy), _f32(desc)
        assert k.shape[0] == d.shape[0]
        _chk(_lib.dv_store_put(self._h, C.c_int64(frame_id), _ptr(k, C.c_float), _ptr(d, C.c_float), int(k.shape[0]),
                               int(n_sp)))

    # ---------------------------------------------------------------- session persistence
    # File "DVSESS01" (little endian): 8s magic | i64 bank_rows | i64 n_keyframes | bank f32 [rows,512] |
    # per keyframe: i64 frame_id, i32 n_total, i32 n_sp, kpts f32 [n_total,2], desc f32 [n_total,256].
    # Replaces the reference's per-keyframe text dumps (pose_graph.cpp:1042-1069), whose load path (:1156-1194) never
    # restores the deep features.
    def save_session(self, path: str, frame_ids):
        import struct
        bank = self.bank_export()
        with open(path, "wb") as f:
            f.write(b"DVSESS01")
            f.write(struct.pack("<qq", bank.shape[0] if self.bank_size() else 0, len(frame_ids)))
            if self.bank_size():
                f.write(np.ascontiguousarray(bank, np.float32).tobytes())
            for fid in frame_ids:
                kp, de, nsp = self.store_read(int(fid))
                f.write(struct.pack("<qii", int(fid), kp.shape[0], nsp))
                f.write(np.ascontiguousarray(kp, np.float32).tobytes())
                f.write(np.ascontiguousarray(de, np.float32).tobytes())

    def load_session(self, path: str):
        import struct
        with open(path, "rb") as f:
            if f.read(8) != b"DVSESS01":
                raise ValueError("not a DVSESS01 file")
            rows, nkf = struct.unpack("<qq", f.read(16))
            bank = np.frombuffer(f.read(rows * GLOBAL_DIM * 4), np.float32).reshape(rows, GLOBAL_DIM)
            self.bank_import(bank)
            ids = []
            for _ in range(nkf):
                fid, n, nsp = struct.unpack("<qii", f.read(16))
                kp = np.frombuffer(f.read(n * 2 * 4), np.float32).reshape(n, 2)
                de = np.frombuffer(f.read(n * DESC_DIM * 4), np.float32).reshape(n, DESC_DIM)
                self.store_put(fid, kp, de, nsp)
                ids.append(fid)
        return ids

    def batch_read_global(self, i):
        d = np.zeros((GLOBAL_DIM,), np.float32)
        _chk(_lib.dv_batch_read_global(self._h, i, _ptr(d, C.c_float)))
        return d

    def comm_init(self, uid: bytes):
        _chk(_lib.dv_comm_init(self._h, C.c_char_p(uid)))

    # ---------------------------------------------------------------- measurement
    def timer_start(self):
        _chk(_lib.dv_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        _chk(_lib.dv_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def sync(self):
        _chk(_lib.dv_sync(self._h))

    def stats_enable(self, on=True):
        _chk(_lib.dv_stats_enable(self._h, 1 if on else 0))

    def stats_reset(self):
        _chk(_lib.dv_stats_reset(self._h))

    def stats_read(self):
        ms = (C.c_double * 6)(); n = C.c_int64(0)
        _chk(_lib.dv_stats_read(self._h, ms, C.byref(n)))
        names = ["sp_convs", "sp_post", "mixvpr", "knn", "lightglue", "copies"]
        return dict(zip(names, list(ms))), n.value

    def probe_enable(self, on=True):
        _chk(_lib.dv_probe_enable(self._h, 1 if on else 0))

    def probe_select(self, which):
        _chk(_lib.dv_probe_select(self._h, int(which)))

    def probe_read(self, reset=True):
        ms = C.c_double(0); n = C.c_int64(0)
        _chk(_lib.dv_probe_read(self._h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    # ---------------------------------------------------------------- stage-level
    def dbg_gemm(self, A, B, bias=None, relu=False):
        A, B = _f32(A), _f32(B)
        M, K = A.shape; N = B.shape[0]
        D = np.zeros((M, N), np.float32)
        bp = _f32(bias) if bias is not None else None
        _chk(_lib.dv_dbg_gemm(self._h, _ptr(A, C.c_float), _ptr(B, C.c_float), _ptr(bp, C.c_float), M, N, K,
                              int(relu), _ptr(D, C.c_float)))
        return D

    def dbg_gemm_ex(self, A, B, bias=None, res=None, res_is_f16=False, relu=False, want32=True, want16=False):
        A, B = _f32(A), _f32(B)
        M, K = A.shape; N = B.shape[0]
        D32 = np.zeros((M, N), np.float32) if want32 else None
        D16 = np.zeros((M, N), np.float32) if want16 else None
        bp = _f32(bias) if bias is not None else None
        rp = _f32(res) if res is not None else None
        _chk(_lib.dv_dbg_gemm_ex(self._h, _ptr(A, C.c_float), _ptr(B, C.c_float), _ptr(bp, C.c_float),
                                 _ptr(rp, C.c_float), int(res_is_f16), M, N, K, int(relu), _ptr(D32, C.c_float),
                                 _ptr(D16, C.c_float)))
        return D32, D16

    def dbg_conv3x3(self, x_nhwc, w_oihw, bias=None, relu=False, pool=False):
        x, w = _f32(x_nhwc), _f32(w_oihw)
        n, h, wd, cin = x.shape; cout = w.shape[0]
        ho, wo = (h // 2, wd // 2) if pool else (h, wd)
        y = np.zeros((n, ho, wo, cout), np.float32)
        bp = _f32(bias) if bias is not None else None
        _chk(_lib.dv_dbg_conv3x3(self._h, _ptr(x, C.c_float), _ptr(w, C.c_float), _ptr(bp, C.c_float), n, h, wd, cin,
                                 cout, int(relu), int(pool), _ptr(y, C.c_float)))
        return y

    def dbg_conv3x3_halo64(self, x_nhwc, w_oihw, bias, relu=True, pool=False, out_blocked=True):
        x, w, bp = _f32(x_nhwc), _f32(w_oihw), _f32(bias)
        n, h, wd, cin = x.shape
        assert cin == 64 and w.shape[:2] == (64, 64)
        ho, wo = (h // 2, wd // 2) if pool else (h, wd)
        y = np.zeros((n, ho, wo, 64), np.float32)
        _chk(_lib.dv_dbg_conv3x3_halo64(self._h, _ptr(x, C.c_float), _ptr(w, C.c_float), _ptr(bp, C.c_float), n, h, wd,
                                        int(relu), int(pool), int(out_blocked), _ptr(y, C.c_float)))
        return y

    def dbg_conv3x3_halo128(self, x_nhwc, w_oihw, bias, relu=True, pool=False, out_blocked=True):
        x, w, bp = _f32(x_nhwc), _f32(w_oihw), _f32(bias)
        n, h, wd, cin = x.shape
        cout = w.shape[0]
        ho, wo = (h // 2, wd // 2) if pool else (h, wd)
        y = np.zeros((n, ho, wo, cout), np.float32)
        _chk(_lib.dv_dbg_conv3x3_halo128(self._h, _ptr(x, C.c_float), _ptr(w, C.c_float), _ptr(bp, C.c_float), n, h, wd,
                                         cin, cout, int(relu), int(pool), int(out_blocked), _ptr(y, C.c_float)))
        return y

    def dbg_nms_select(self, score_map):
        s = _f32(score_map); h8, w8 = s.shape
        nms = np.zeros_like(s); k = self.cfg.max_kpts
        kp = np.zeros((k, 2), np.int32); sc = np.zeros((k,), np.float32); n = C.c_int32(0)
        _chk(_lib.dv_dbg_nms_select(self._h, _ptr(s, C.c_float), h8, w8, _ptr(nms, C.c_float), _ptr(kp, C.c_int32),
                                    _ptr(sc, C.c_float), C.byref(n)))
        return nms, kp[:n.value], sc[:n.value]

    def dbg_match_extract(self, L):
        L = _f32(L); m, n = L.shape
        cap = max(1, min(m, n))
        ma = np.zeros((cap, 2), np.int32); ms = np.zeros((cap,), np.float32); ko = C.c_int32(0)
        _chk(_lib.dv_dbg_match_extract(self._h, _ptr(L, C.c_float), m, n, _ptr(ma, C.c_int32), _ptr(ms, C.c_float),
                                       C.byref(ko)))
        return ma[:ko.value], ms[:ko.value]

    def dbg_read(self, name: str) -> np.ndarray:
        cnt = C.c_int64(0)
        _chk(_lib.dv_dbg_read(self._h, name.encode(), None, C.c_int64(0), C.byref(cnt)))
        out = np.zeros((cnt.value,), np.float32)
        _chk(_lib.dv_dbg_read(self._h, name.encode(), _ptr(out, C.c_float), C.c_int64(cnt.value), C.byref(cnt)))
        return out
