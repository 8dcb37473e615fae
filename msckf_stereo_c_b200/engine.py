"""Host-side mirror of the reference's System / ImageProcessor / MsckfVio callback API over
the C ABI of include/msckf_b200.h (libmsckf_b200.so, hand-written sm_100a kernels).

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but
creating an engine without the built library or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSKF_B200_LIB", os.path.join(_HERE, "libmsckf_b200.so"))  # override: A/B kernel experiments
_LIB = None

GRID_DT = np.dtype([("id", "<u8"), ("response", "<f4"), ("lifetime", "<i4"), ("cam0", "<f4", 2), ("cam1", "<f4", 2),
                    ("cell", "<i4"), ("pad", "<i4")])
FEAT_DT = np.dtype([("id", "<u4"), ("pad", "<u4"), ("u0", "<f8"), ("v0", "<f8"), ("u1", "<f8"), ("v1", "<f8")])
CAM_DT = np.dtype([("id", "<i8"), ("time", "<f8"), ("orientation", "<f8", 4), ("position", "<f8", 3)])

# every symbol include/msckf_b200.h declares
ABI_SYMBOLS = [
    "mskf_default_config", "mskf_create", "mskf_destroy", "mskf_last_error", "mskf_set_cuda_stream",
    "mskf_push_imu", "mskf_push_stereo", "mskf_push_stereo_device", "mskf_frontend_step", "mskf_backend_step",
    "mskf_step", "mskf_sync", "mskf_backend_step_features", "mskf_get_features", "mskf_get_tracking_info",
    "mskf_get_grid", "mskf_get_pyramid", "mskf_get_state", "mskf_get_cam_states", "mskf_get_covariance",
    "mskf_reset", "mskf_op_pyramid", "mskf_op_detect", "mskf_op_klt",
    "mskf_launch_count", "mskf_get_n_published", "mskf_get_poses", "mskf_profile_enable", "mskf_profile_read",
    "mskf_debug_detect_scores", "mskf_debug_get_map", "mskf_debug_last_gram", "mskf_op_ekf_update", "mskf_push_imu_batch",
    "mskf_push_stereo_batch", "mskf_push_stereo_device_batch", "mskf_get_work", "mskf_get_poses_prev", "mskf_join",
    "mskf_set_overlap", "mskf_debug_update_dims", "mskf_op_triangulate", "mskf_push_imu_to",
    "mskf_get_features_head", "mskf_wait_uploads", "mskf_host_alloc", "mskf_host_alloc_wc", "mskf_host_free",
]


class EngineError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        P, I, D = C.c_void_p, C.c_int, C.c_double
        L.mskf_default_config.argtypes = [C.POINTER(abi.Config), C.c_char_p]
        L.mskf_create.argtypes = [C.POINTER(abi.Config), I, I, C.POINTER(P)]
        L.mskf_destroy.argtypes = [P]
        L.mskf_destroy.restype = None
        L.mskf_last_error.argtypes = [P]
        L.mskf_last_error.restype = C.c_char_p
        L.mskf_set_cuda_stream.argtypes = [P, P]
        L.mskf_launch_count.argtypes = [P]
        L.mskf_launch_count.restype = C.c_longlong
        L.mskf_push_imu.argtypes = [P, I, D, P, P]
        L.mskf_push_imu_to.argtypes = [P, I, I, D, P, P]
        L.mskf_get_features_head.argtypes = [P, I, P, I, C.POINTER(I), C.POINTER(C.c_longlong), C.POINTER(D)]
        L.mskf_push_stereo.argtypes = [P, I, D, P, P, I, I, I]
        L.mskf_push_stereo_device.argtypes = [P, I, D, P, P]
        for n in ("mskf_frontend_step", "mskf_backend_step", "mskf_step", "mskf_sync"):
            getattr(L, n).argtypes = [P]
        L.mskf_backend_step_features.argtypes = [P, I, D, P, I]
        L.mskf_get_features.argtypes = [P, I, P, I, C.POINTER(I), C.POINTER(D)]
        L.mskf_get_n_published.argtypes = [P, I, C.POINTER(I)]
        L.mskf_get_tracking_info.argtypes = [P, I, C.POINTER(abi.TrackingInfo)]
        L.mskf_get_grid.argtypes = [P, I, P, I, C.POINTER(I)]
        L.mskf_get_pyramid.argtypes = [P, I, I, I, P, I, C.POINTER(I), C.POINTER(I)]
        L.mskf_get_state.argtypes = [P, I, C.POINTER(abi.State)]
        L.mskf_get_cam_states.argtypes = [P, I, P, I, C.POINTER(I)]
        L.mskf_get_covariance.argtypes = [P, I, P, I, C.POINTER(I)]
        L.mskf_reset.argtypes = [P, I]
        L.mskf_op_pyramid.argtypes = [P, P, I, I, I, I, P]
        L.mskf_op_detect.argtypes = [P, P, I, I, P, I, P, P, I, C.POINTER(I)]
        L.mskf_op_klt.argtypes = [P, P, P, I, I, P, P, P, I]
        L.mskf_debug_detect_scores.argtypes = [P, P, I, I, P, P, I, C.POINTER(I), P]
        L.mskf_get_poses.argtypes = [P, P, I]
        L.mskf_get_poses_prev.argtypes = [P, P, I]
        L.mskf_join.argtypes = [P]
        L.mskf_debug_update_dims.argtypes = [P, P]
        L.mskf_debug_last_gram.argtypes = [P, I, P, I, C.POINTER(I), C.POINTER(I), P, C.POINTER(I)]
        L.mskf_set_overlap.argtypes = [P, I]
        L.mskf_push_imu_batch.argtypes = [P, I, I, P]
        L.mskf_push_stereo_batch.argtypes = [P, P, P, P, C.c_size_t]
        L.mskf_push_stereo_device_batch.argtypes = [P, P, P, P, C.c_size_t]
        L.mskf_get_work.argtypes = [P, I, C.POINTER(D)]
        L.mskf_op_ekf_update.argtypes = [P, I, I, P, P, P, P, P]
        L.mskf_op_triangulate.argtypes = [P, I, P, P, I, P, P, P, P]
        L.mskf_debug_get_map.argtypes = [P, I, P, P, P, P, I, C.POINTER(I)]
        L.mskf_profile_enable.argtypes = [P, I]
        L.mskf_profile_read.argtypes = [P, I, C.POINTER(C.c_char_p), C.POINTER(D), C.POINTER(C.c_longlong)]
        _LIB = L
    return _LIB


def default_config(preset="ref"):
    cfg = abi.Config()
    if lib().mskf_default_config(C.byref(cfg), preset.encode()) != 0:
        raise ValueError(f"unknown preset {preset!r}")
    return cfg


class Engine:
    """`n_streams` independent System instances (system.cpp:12-54) advanced together on one GPU."""

    def __init__(self, cfg, n_streams=1, device=0, cuda_stream=None):
        self.cfg, self.n_streams = cfg, n_streams
        self.h = C.c_void_p()
        rc = lib().mskf_create(C.byref(cfg), n_streams, device, C.byref(self.h))
        if rc != 0:
            msg = lib().mskf_last_error(self.h).decode() if self.h else "no CUDA device"
            if self.h:
                lib().mskf_destroy(self.h)
                self.h = None
            raise EngineError(f"mskf_create failed ({rc}): {msg}")
        if cuda_stream is not None:
            self._ck(lib().mskf_set_cuda_stream(self.h, C.c_void_p(cuda_stream)))

    def close(self):
        if getattr(self, "h", None) and _LIB is not None:  # _LIB is None again at interpreter shutdown
            _LIB.mskf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: ctypes may already be torn down
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"msckf_b200 error {rc}: {lib().mskf_last_error(self.h).decode()}")

    # ---- System::imu_callback / stereo_callback / backend_callback ---------------------
    def imu_callback(self, t, w, a, stream=0):
        w = np.ascontiguousarray(w, np.float64)
        a = np.ascontiguousarray(a, np.float64)
        self._ck(lib().mskf_push_imu(self.h, stream, t, w.ctypes.data, a.ctypes.data))

    def imu_callback_to(self, halves, t, w, a, stream=0):
        """halves: 1 = ImageProcessor::imuCallback only, 2 = MsckfVio::imuCallback only, 3 = both."""
        w = np.ascontiguousarray(w, np.float64)
        a = np.ascontiguousarray(a, np.float64)
        self._ck(lib().mskf_push_imu_to(self.h, stream, halves, t, w.ctypes.data, a.ctypes.data))

    def features_head(self, stream=0):
        """(t, entries that can hold measurements, length of the reference's never-cleared vector)."""
        n, tot, t = C.c_int(), C.c_longlong(), C.c_double()
        self._ck(lib().mskf_get_features_head(self.h, stream, None, 0, C.byref(n), C.byref(tot), C.byref(t)))
        out = np.zeros(n.value, FEAT_DT)
        if n.value:
            self._ck(lib().mskf_get_features_head(self.h, stream, out.ctypes.data, n.value, C.byref(n), C.byref(tot), C.byref(t)))
        return t.value, out, tot.value

    def push_stereo(self, t, cam0, cam1, stream=0):
        cam0 = np.ascontiguousarray(cam0, np.uint8)
        cam1 = np.ascontiguousarray(cam1, np.uint8)
        r, c = cam0.shape
        self._ck(lib().mskf_push_stereo(self.h, stream, t, cam0.ctypes.data, cam1.ctypes.data, r, c, c))

    def push_stereo_ptr(self, t, p0, p1, stream=0, device=False):
        if device:
            self._ck(lib().mskf_push_stereo_device(self.h, stream, t, C.c_void_p(p0), C.c_void_p(p1)))
        else:
            self._ck(lib().mskf_push_stereo(self.h, stream, t, C.c_void_p(p0), C.c_void_p(p1), self.cfg.img_rows,
                                            self.cfg.img_cols, self.cfg.img_cols))

    def stereo_callback(self, t, cam0, cam1, stream=0):
        self.push_stereo(t, cam0, cam1, stream)
        self.frontend_step()

    def frontend_step(self):
        self._ck(lib().mskf_frontend_step(self.h))

    def backend_callback(self):
        self._ck(lib().mskf_backend_step(self.h))

    def step(self):
        self._ck(lib().mskf_step(self.h))

    def sync(self):
        self._ck(lib().mskf_sync(self.h))

    def backend_features(self, t, feats, stream=0):
        feats = np.ascontiguousarray(feats, FEAT_DT)
        self._ck(lib().mskf_backend_step_features(self.h, stream, t, feats.ctypes.data, len(feats)))

    # sink protocol shared with the oracle (synth.feed)
    def imu(self, t, w, a):
        self.imu_callback(t, w, a, 0)

    def stereo(self, t, im0, im1):
        self.stereo_callback(t, im0, im1, 0)

    def backend(self):
        self.backend_callback()

    # ---- instrumentation ----------------------------------------------------------------
    def profile_enable(self, on=True):
        self._ck(lib().mskf_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """{kernel class: (total ms, launches, algorithmic work)} since profile_enable(True); work is
        bytes for the front-end classes and flops for the EKF classes."""
        out, tag = {}, 0
        while True:
            name, ms, n, w = C.c_char_p(), C.c_double(), C.c_longlong(), C.c_double()
            rc = lib().mskf_profile_read(self.h, tag, C.byref(name), C.byref(ms), C.byref(n))
            if rc == 1:
                break
            self._ck(rc)
            self._ck(lib().mskf_get_work(self.h, tag, C.byref(w)))
            out[name.value.decode()] = (ms.value, n.value, w.value)
            tag += 1
        return out

    # ---- fleet variants ----------------------------------------------------------------
    def push_imu_batch(self, samples, stream=-1):
        samples = np.ascontiguousarray(samples, np.float64)
        n = samples.shape[-2]
        self._ck(lib().mskf_push_imu_batch(self.h, stream, n, samples.ctypes.data))

    def push_stereo_batch(self, t, p0, p1, stream_stride, device=False):
        t = np.ascontiguousarray(t, np.float64)
        f = lib().mskf_push_stereo_device_batch if device else lib().mskf_push_stereo_batch
        self._ck(f(self.h, t.ctypes.data, C.c_void_p(p0), C.c_void_p(p1), stream_stride))

    def feature_map(self, stream=0):
        n = C.c_int()
        cap = 8192
        ids = np.zeros(cap, np.int64)
        init = np.zeros(cap, np.int32)
        pos = np.zeros((cap, 3))
        nobs = np.zeros(cap, np.int32)
        self._ck(lib().mskf_debug_get_map(self.h, stream, ids.ctypes.data, init.ctypes.data, pos.ctypes.data,
                                          nobs.ctypes.data, cap, C.byref(n)))
        k = n.value
        return ids[:k], init[:k], pos[:k], nobs[:k]

    def poses(self, prev=False):
        """T_b_w of every stream after the latest back-end step (prev=True: the step before it, which
        is on the host already while the latest one is still running)."""
        out = np.zeros((self.n_streams, 4, 4))
        f = lib().mskf_get_poses_prev if prev else lib().mskf_get_poses
        self._ck(f(self.h, out.ctypes.data, self.n_streams))
        return out

    def last_gram(self, stream=0):
        """(G, m, k, cam_ids) of the latest measurementUpdate of a stream: G = [H r]^T [H r] over the k active
        camera columns, or G = None when m <= k (mskf_debug_last_gram)."""
        m, k, valid = C.c_int(), C.c_int(), C.c_int()
        ids = np.zeros(32, np.int64)
        G = np.zeros((6 * 32 + 1) ** 2)
        self._ck(lib().mskf_debug_last_gram(self.h, stream, G.ctypes.data, G.size, C.byref(m), C.byref(k), ids.ctypes.data,
                                            C.byref(valid)))
        kw = k.value + 1
        return (G[:kw * kw].reshape(kw, kw).copy() if valid.value else None), m.value, k.value, ids[:k.value // 6].copy()

    def update_dims(self):
        out = np.zeros((self.n_streams, 2, 3), np.int32)
        self._ck(lib().mskf_debug_update_dims(self.h, out.ctypes.data))
        return out

    def join(self):
        self._ck(lib().mskf_join(self.h))

    def set_overlap(self, on=True):
        self._ck(lib().mskf_set_overlap(self.h, 1 if on else 0))

    # ---- outputs -----------------------------------------------------------------------
    def launch_count(self):
        return int(lib().mskf_launch_count(self.h))

    def features(self, stream=0):
        n, t = C.c_int(), C.c_double()
        self._ck(lib().mskf_get_features(self.h, stream, None, 0, C.byref(n), C.byref(t)))
        out = np.zeros(n.value, FEAT_DT)
        self._ck(lib().mskf_get_features(self.h, stream, out.ctypes.data, n.value, C.byref(n), C.byref(t)))
        npub = C.c_int()
        self._ck(lib().mskf_get_n_published(self.h, stream, C.byref(npub)))
        return t.value, out, npub.value

    def tracking_info(self, stream=0):
        ti = abi.TrackingInfo()
        self._ck(lib().mskf_get_tracking_info(self.h, stream, C.byref(ti)))
        return ti

    def grid(self, stream=0):
        n = C.c_int()
        self._ck(lib().mskf_get_grid(self.h, stream, None, 0, C.byref(n)))
        out = np.zeros(n.value, GRID_DT)
        if n.value:
            self._ck(lib().mskf_get_grid(self.h, stream, out.ctypes.data, n.value, C.byref(n)))
        return out

    def pyramid(self, cam, level, stream=0):
        r, c = C.c_int(), C.c_int()
        self._ck(lib().mskf_get_pyramid(self.h, stream, cam, level, None, 0, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.uint8)
        self._ck(lib().mskf_get_pyramid(self.h, stream, cam, level, out.ctypes.data, out.size, C.byref(r), C.byref(c)))
        return out

    def state(self, stream=0):
        s = abi.State()
        self._ck(lib().mskf_get_state(self.h, stream, C.byref(s)))
        return s

    def cam_states(self, stream=0):
        n = C.c_int()
        self._ck(lib().mskf_get_cam_states(self.h, stream, None, 0, C.byref(n)))
        out = np.zeros(n.value, CAM_DT)
        if n.value:
            self._ck(lib().mskf_get_cam_states(self.h, stream, out.ctypes.data, n.value, C.byref(n)))
        return out

    def cov(self, stream=0):
        d = C.c_int()
        self._ck(lib().mskf_get_covariance(self.h, stream, None, 0, C.byref(d)))
        out = np.zeros((d.value, d.value))
        self._ck(lib().mskf_get_covariance(self.h, stream, out.ctypes.data, out.size, C.byref(d)))
        return out

    def reset(self, stream=0):
        self._ck(lib().mskf_reset(self.h, stream))

    # ---- stand-alone operators ---------------------------------------------------------
    def op_pyramid(self, imgs, levels):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        n, r, c = imgs.shape
        sizes, rr, cc = [], r, c
        for _ in range(1, levels):
            rr, cc = (rr + 1) // 2, (cc + 1) // 2
            sizes.append((rr, cc))
        per = sum(a * b for a, b in sizes)
        out = np.empty((n, per), np.uint8)
        self._ck(lib().mskf_op_pyramid(self.h, imgs.ctypes.data, n, r, c, levels, out.ctypes.data))
        res = []
        for i in range(n):
            o, lv = 0, []
            for a, b in sizes:
                lv.append(out[i, o:o + a * b].reshape(a, b))
                o += a * b
            res.append(lv)
        return res

    def op_detect(self, img, occupied=None):
        img = np.ascontiguousarray(img, np.uint8)
        occ = np.ascontiguousarray(occupied if occupied is not None else np.zeros((0, 2)), np.float32)
        cap = self.cfg.det_rows * self.cfg.det_cols
        xy = np.zeros((cap, 2), np.float32)
        resp = np.zeros(cap)
        n = C.c_int()
        self._ck(lib().mskf_op_detect(self.h, img.ctypes.data, img.shape[0], img.shape[1], occ.ctypes.data, len(occ),
                                      xy.ctypes.data, resp.ctypes.data, cap, C.byref(n)))
        return xy[:n.value], resp[:n.value]

    def op_ekf_update(self, H, r, P):
        H = np.ascontiguousarray(H, np.float64)
        r = np.ascontiguousarray(r, np.float64)
        P = np.ascontiguousarray(P, np.float64)
        m, n = H.shape
        dx = np.zeros(n)
        Pn = np.zeros((n, n))
        self._ck(lib().mskf_op_ekf_update(self.h, (n - 21) // 6, m, H.ctypes.data, r.ctypes.data, P.ctypes.data,
                                          dx.ctypes.data, Pn.ctypes.data))
        return dx, Pn

    def op_triangulate(self, cam_q, cam_p, mask, obs):
        """Feature::checkMotion + initializePosition on n_cam camera states (ascending id) and
        obs[n_feat][n_cam][4]; mask bit c of feature f = camera state c observes f."""
        cam_q = np.ascontiguousarray(cam_q, np.float64)
        cam_p = np.ascontiguousarray(cam_p, np.float64)
        mask = np.ascontiguousarray(mask, np.uint32)
        obs = np.ascontiguousarray(obs, np.float64)
        n_cam, n_feat = cam_q.shape[0], mask.shape[0]
        assert obs.shape == (n_feat, n_cam, 4)
        pos = np.zeros((n_feat, 3))
        ok = np.zeros(n_feat, np.int32)
        self._ck(lib().mskf_op_triangulate(self.h, n_cam, cam_q.ctypes.data, cam_p.ctypes.data, n_feat, mask.ctypes.data,
                                           obs.ctypes.data, pos.ctypes.data, ok.ctypes.data))
        return pos, ok

    def debug_detect_scores(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.cfg.det_rows * self.cfg.det_cols
        xy = np.zeros((cap, 2), np.float32)
        resp = np.zeros(cap)
        sm = np.zeros(img.shape, np.uint8)
        n = C.c_int()
        self._ck(lib().mskf_debug_detect_scores(self.h, img.ctypes.data, img.shape[0], img.shape[1], xy.ctypes.data,
                                                resp.ctypes.data, cap, C.byref(n), sm.ctypes.data))
        return xy[:n.value], resp[:n.value], sm

    def op_klt(self, a, b, pts_a, pts_b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        pa = np.ascontiguousarray(pts_a, np.float32)
        pb = np.array(pts_b, np.float32, copy=True)
        st = np.zeros(len(pa), np.uint8)
        self._ck(lib().mskf_op_klt(self.h, a.ctypes.data, b.ctypes.data, a.shape[0], a.shape[1], pa.ctypes.data,
                                   pb.ctypes.data, st.ctypes.data, len(pa)))
        return pb, st
