"""ORACLE — TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from msckf_stereo_c_b200 import abi  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.Config)]
        for name in ("orc_destroy", "orc_backend", "orc_reset"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_imu.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_stereo.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_backend_features.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        L.orc_get_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_get_n_published.argtypes = [C.c_void_p]
        L.orc_get_tracking_info.argtypes = [C.c_void_p, C.POINTER(abi.TrackingInfo)]
        L.orc_get_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_get_state.argtypes = [C.c_void_p, C.POINTER(abi.State)]
        L.orc_get_cam_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_cov.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_pyr_down.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_detect.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_klt.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_undistort.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_distort.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_rodrigues.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_qr_thin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ldlt_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_nullspace_update.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_update_math.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                      C.c_void_p]
        L.orc_keep_last_update.argtypes = [C.c_void_p, C.c_int]
        L.orc_last_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_get_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_last_update_cam_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_chi2.argtypes = [C.POINTER(abi.Config), C.c_int]
        L.orc_chi2.restype = C.c_double
        _LIB = L
    return _LIB


GRID_DT = np.dtype([("id", "<u8"), ("response", "<f4"), ("lifetime", "<i4"), ("cam0", "<f4", 2), ("cam1", "<f4", 2),
                    ("cell", "<i4"), ("pad", "<i4")])
FEAT_DT = np.dtype([("id", "<u4"), ("pad", "<u4"), ("u0", "<f8"), ("v0", "<f8"), ("u1", "<f8"), ("v1", "<f8")])
CAM_DT = np.dtype([("id", "<i8"), ("time", "<f8"), ("orientation", "<f8", 4), ("position", "<f8", 3)])


class Oracle:
    """System re-host (system.cpp:12-54) over the CPU restatement."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().orc_create(C.byref(cfg))

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.orc_destroy(self.h)
            self.h = None

    def imu(self, t, w, a):
        w = np.ascontiguousarray(w, np.float64)
        a = np.ascontiguousarray(a, np.float64)
        lib().orc_imu(self.h, t, w.ctypes.data, a.ctypes.data)

    def stereo(self, t, im0, im1):
        im0 = np.ascontiguousarray(im0, np.uint8)
        im1 = np.ascontiguousarray(im1, np.uint8)
        lib().orc_stereo(self.h, t, im0.ctypes.data, im1.ctypes.data)

    def backend(self):
        lib().orc_backend(self.h)

    def backend_features(self, t, feats):
        feats = np.ascontiguousarray(feats, FEAT_DT)
        lib().orc_backend_features(self.h, t, feats.ctypes.data, len(feats))

    def features(self):
        t = C.c_double()
        n = lib().orc_get_features(self.h, None, 0, C.byref(t))
        out = np.zeros(n, FEAT_DT)
        lib().orc_get_features(self.h, out.ctypes.data, n, C.byref(t))
        return t.value, out, lib().orc_get_n_published(self.h)

    def tracking_info(self):
        ti = abi.TrackingInfo()
        lib().orc_get_tracking_info(self.h, C.byref(ti))
        return ti

    def grid(self):
        n = lib().orc_get_grid(self.h, None, 0)
        out = np.zeros(n, GRID_DT)
        lib().orc_get_grid(self.h, out.ctypes.data, n)
        return out

    def pyramid(self, cam, level):
        r, c = C.c_int(), C.c_int()
        buf = np.zeros(self.cfg.img_rows * self.cfg.img_cols, np.uint8)
        rc = lib().orc_get_pyramid(self.h, cam, level, buf.ctypes.data, buf.size, C.byref(r), C.byref(c))
        assert rc == 0
        return buf[: r.value * c.value].reshape(r.value, c.value).copy()

    def state(self):
        s = abi.State()
        lib().orc_get_state(self.h, C.byref(s))
        return s

    def cam_states(self):
        n = lib().orc_get_cam_states(self.h, None, 0)
        out = np.zeros(n, CAM_DT)
        lib().orc_get_cam_states(self.h, out.ctypes.data, n)
        return out

    def feature_map(self):
        """[(id, is_initialized, position, n_observations)] of map_server, ascending id."""
        n = lib().orc_get_map(self.h, None, None, None, None, 0)
        ids = np.zeros(n, np.int64)
        init = np.zeros(n, np.int32)
        pos = np.zeros((n, 3))
        nobs = np.zeros(n, np.int32)
        lib().orc_get_map(self.h, ids.ctypes.data, init.ctypes.data, pos.ctypes.data, nobs.ctypes.data, n)
        return ids, init, pos, nobs

    def keep_last_update(self, on=True):
        lib().orc_keep_last_update(self.h, 1 if on else 0)

    def last_update(self):
        """(H, r, P-) of the latest measurementUpdate (needs keep_last_update())."""
        n = C.c_int()
        m = lib().orc_last_update(self.h, None, None, None, 0, C.byref(n))
        H = np.zeros((m, n.value))
        r = np.zeros(m)
        P = np.zeros((n.value, n.value))
        lib().orc_last_update(self.h, H.ctypes.data, r.ctypes.data, P.ctypes.data, max(H.size, P.size), C.byref(n))
        return H, r, P

    def last_update_cam_ids(self):
        """Camera-state ids behind the column groups 21 + 6 i of last_update()'s H."""
        ids = np.zeros(64, np.int64)
        n = lib().orc_last_update_cam_ids(self.h, ids.ctypes.data, 64)
        return ids[:n].copy()

    def cov(self):
        n = lib().orc_get_cov(self.h, None, 0)
        out = np.zeros((n, n))
        lib().orc_get_cov(self.h, out.ctypes.data, n * n)
        return out


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(((img.shape[0] + 1) // 2, (img.shape[1] + 1) // 2), np.uint8)
    lib().orc_pyr_down(img.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data)
    return out


def detect(cfg, img, occupied=None, want_scores=False):
    img = np.ascontiguousarray(img, np.uint8)
    occ = np.ascontiguousarray(occupied if occupied is not None else np.zeros((0, 2)), np.float32)
    cap = cfg.det_rows * cfg.det_cols
    xy = np.zeros((cap, 2), np.float32)
    resp = np.zeros(cap)
    sm = np.zeros(img.shape, np.uint8) if want_scores else None
    n = lib().orc_detect(C.byref(cfg), img.ctypes.data, occ.ctypes.data, len(occ), xy.ctypes.data, resp.ctypes.data, cap,
                         sm.ctypes.data if want_scores else None)
    return (xy[:n], resp[:n], sm) if want_scores else (xy[:n], resp[:n])


def klt(cfg, a, b, pts_a, pts_b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    pa = np.ascontiguousarray(pts_a, np.float32)
    pb = np.array(pts_b, np.float32, copy=True)
    st = np.zeros(len(pa), np.uint8)
    lib().orc_klt(C.byref(cfg), a.ctypes.data, b.ctypes.data, pa.ctypes.data, pb.ctypes.data, st.ctypes.data, len(pa))
    return pb, st


def undistort(pts, K, model, D, R=None, Kn=(1, 1, 0, 0)):
    pts = np.ascontiguousarray(pts, np.float32)
    K = np.ascontiguousarray(K, np.float64)
    D = np.ascontiguousarray(D, np.float64)
    R = np.ascontiguousarray(np.eye(3) if R is None else R, np.float64)
    Kn = np.ascontiguousarray(Kn, np.float64)
    out = np.zeros_like(pts)
    lib().orc_undistort(pts.ctypes.data, len(pts), K.ctypes.data, model, D.ctypes.data, R.ctypes.data, Kn.ctypes.data, out.ctypes.data)
    return out


def distort(pts, K, model, D):
    pts = np.ascontiguousarray(pts, np.float32)
    K = np.ascontiguousarray(K, np.float64)
    D = np.ascontiguousarray(D, np.float64)
    out = np.zeros_like(pts)
    lib().orc_distort(pts.ctypes.data, len(pts), K.ctypes.data, model, D.ctypes.data, out.ctypes.data)
    return out


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def rodrigues(v):
    v = _f64(v)
    R = np.zeros((3, 3))
    lib().orc_rodrigues(v.ctypes.data, R.ctypes.data)
    return R


def qr_thin(A, b):
    A, b = _f64(A), _f64(b)
    m, n = A.shape
    R = np.zeros((n, n))
    qtb = np.zeros(n)
    lib().orc_qr_thin(A.ctypes.data, m, n, b.ctypes.data, R.ctypes.data, qtb.ctypes.data)
    return R, qtb


def ldlt_solve(S, B):
    S, B = _f64(S), _f64(B)
    B2 = B.reshape(len(S), -1)
    X = np.zeros_like(B2)
    lib().orc_ldlt_solve(S.ctypes.data, len(S), B2.ctypes.data, B2.shape[1], X.ctypes.data)
    return X.reshape(B.shape)


def nullspace_update(Hx, Hf, r, P, obs_noise, basis=None):
    Hx, Hf, r, P = _f64(Hx), _f64(Hf), _f64(r), _f64(P)
    rows, n = Hx.shape
    dx = np.zeros(n)
    Pn = np.zeros((n, n))
    g = C.c_double()
    b = _f64(basis) if basis is not None else None
    lib().orc_nullspace_update(n, rows, Hx.ctypes.data, Hf.ctypes.data, r.ctypes.data, P.ctypes.data, obs_noise,
                               b.ctypes.data if b is not None else None, dx.ctypes.data, Pn.ctypes.data, C.byref(g))
    return dx, Pn, g.value


def update_math(H, r, P, obs_noise):
    H, r, P = _f64(H), _f64(r), _f64(P)
    m, n = H.shape
    dx = np.zeros(n)
    Pn = np.zeros((n, n))
    lib().orc_update_math(n, m, H.ctypes.data, r.ctypes.data, P.ctypes.data, obs_noise, dx.ctypes.data, Pn.ctypes.data)
    return dx, Pn


def triangulate(cfg, cam_q, cam_p, mask, obs):
    """Feature::checkMotion + initializePosition (feature.hpp:257-450); same layout as Engine.op_triangulate."""
    cam_q, cam_p, obs = _f64(cam_q), _f64(cam_p), _f64(obs)
    mask = np.ascontiguousarray(mask, np.uint32)
    n_cam, n_feat = cam_q.shape[0], mask.shape[0]
    pos = np.zeros((n_feat, 3))
    ok = np.zeros(n_feat, np.int32)
    f = lib().orc_triangulate
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f.restype = None
    f(C.byref(cfg), n_cam, cam_q.ctypes.data, cam_p.ctypes.data, n_feat, mask.ctypes.data, obs.ctypes.data,
      pos.ctypes.data, ok.ctypes.data)
    return pos, ok


def chi2(cfg, dof):
    return lib().orc_chi2(C.byref(cfg), dof)
