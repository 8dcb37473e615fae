"""Seeded synthetic EuRoC-style stereo+IMU streams (host build of synth/scene.h).

Makes the inputs of BASELINE.json config 1 ("60 s textured room, 752x480 stereo @20 Hz,
IMU @200 Hz").  Input generation only: nothing here is on the hot path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    d = os.path.join(_HERE, "synth")
    so = os.path.join(d, "libmskf_synth.so")
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(d, "libmskf_synth_cuda.so")):
        subprocess.check_call(["make", "-s", "-C", d])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.synth_create.restype = C.c_void_p
        L.synth_create.argtypes = [C.POINTER(abi.Config), C.c_uint32, C.c_double]
        L.synth_destroy.argtypes = [C.c_void_p]
        L.synth_get_pose.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.synth_get_imu.argtypes = [C.c_void_p, C.c_double, C.c_uint32, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.synth_render.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int]
        L.synth_rays.restype = C.c_void_p
        L.synth_rays.argtypes = [C.c_void_p, C.c_int]
        L.synth_traj.restype = C.c_void_p
        L.synth_traj.argtypes = [C.c_void_p]
        L.synth_cam.restype = C.c_void_p
        L.synth_cam.argtypes = [C.c_void_p, C.c_int]
        L.synth_default_config.argtypes = [C.POINTER(abi.Config), C.c_char_p]
        L.synth_make_trajs.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.synth_imu_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.synth_traj_pose.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


_CUDA_LIB = None


def cuda_lib():
    """The device renderer (libmskf_synth_cuda.so, built from synth/synth_cuda.cu)."""
    global _CUDA_LIB
    if _CUDA_LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "synth", "libmskf_synth_cuda.so"))
        P, I = C.c_void_p, C.c_int
        L.mskf_synth_render_device.argtypes = [P, P, P, P, P, P, I, I, I, P]
        _CUDA_LIB = L
    return _CUDA_LIB


def default_config(preset="ref"):
    cfg = abi.Config()
    rc = lib().synth_default_config(C.byref(cfg), preset.encode())
    if rc != 0:
        raise ValueError(f"unknown preset {preset!r}")
    return cfg


class Stream:
    """One synthetic stereo+IMU stream: images at `frame_rate`, IMU at `imu_rate`."""

    def __init__(self, cfg, seed=0, t0=1000.0, frame_rate=20.0, imu_rate=200.0, imu_noise=True, threads=8):
        self.cfg, self.seed, self.t0 = cfg, seed, t0
        self.frame_dt, self.imu_dt = 1.0 / frame_rate, 1.0 / imu_rate
        self.threads = threads
        self.h = lib().synth_create(C.byref(cfg), seed, t0)
        # discrete-time sigma of the continuous noise densities the filter is configured with
        self.ng = cfg.noise_gyro * np.sqrt(imu_rate) if imu_noise else 0.0
        self.na = cfg.noise_acc * np.sqrt(imu_rate) if imu_noise else 0.0

    def __del__(self):
        try:
            if getattr(self, "h", None) and _LIB is not None:
                _LIB.synth_destroy(self.h)
                self.h = None
        except Exception:  # interpreter shutdown
            pass

    def frame_time(self, k):
        # images lag the IMU clock start by 1/4 IMU period so stamps never coincide exactly
        return self.t0 + k * self.frame_dt + 0.25 * self.imu_dt

    def imu_time(self, j):
        return self.t0 + j * self.imu_dt

    def imu(self, j):
        w = np.zeros(3)
        a = np.zeros(3)
        lib().synth_get_imu(self.h, self.imu_time(j), j, self.ng, self.na, w.ctypes.data, a.ctypes.data)
        return self.imu_time(j), w, a

    def pose(self, t):
        R = np.zeros(9)
        p = np.zeros(3)
        lib().synth_get_pose(self.h, t, R.ctypes.data, p.ctypes.data)
        return R.reshape(3, 3), p

    def render(self, k):
        t = self.frame_time(k)
        out = []
        for cam in (0, 1):
            img = np.empty((self.cfg.img_rows, self.cfg.img_cols), np.uint8)
            lib().synth_render(self.h, t, cam, img.ctypes.data, self.threads)
            out.append(img)
        return t, out[0], out[1]

    def rays(self, cam):
        n = self.cfg.img_rows * self.cfg.img_cols * 8
        p = lib().synth_rays(self.h, cam)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy()


def feed(stream, n_frames, sink):
    """EuRoC runner feed order (apps/run_euroc_single_thread.cpp:209-254): before image k,
    push IMU rows until one has t > t_img (the overshoot row is pushed too)."""
    j = 0
    for k in range(n_frames):
        t_img, im0, im1 = stream.render(k)
        while True:
            t, w, a = stream.imu(j)
            j += 1
            sink.imu(t, w, a)
            if not (t <= t_img):
                break
        sink.stereo(t_img, im0, im1)
        sink.backend()
        yield k, t_img


class Fleet:
    """`n` synthetic streams sharing one calibration (BASELINE.json config 4: independent
    stereo+IMU streams, seed = global stream index).  IMU rows come from the host generator; the
    images can be rendered on the device (scene.h compiled by nvcc into the engine library) so a
    256-stream batch does not wait for the CPU renderer.  Input generation only."""

    def __init__(self, cfg, seeds, t0=1000.0, frame_rate=20.0, imu_rate=200.0, imu_noise=True):
        self.cfg, self.n = cfg, len(seeds)
        self.frame_dt, self.imu_dt, self.t0 = 1.0 / frame_rate, 1.0 / imu_rate, t0
        self.ng = cfg.noise_gyro * np.sqrt(imu_rate) if imu_noise else 0.0
        self.na = cfg.noise_acc * np.sqrt(imu_rate) if imu_noise else 0.0
        L = lib()
        self.tsz = L.synth_traj_size()
        self.trajs = np.zeros(self.n * self.tsz, np.uint8)
        sd = np.ascontiguousarray(seeds, np.uint32)
        L.synth_make_trajs(sd.ctypes.data, self.n, t0, self.trajs.ctypes.data)
        self.proto = Stream(cfg, seed=int(seeds[0]), t0=t0, frame_rate=frame_rate, imu_rate=imu_rate, imu_noise=imu_noise)
        self._dev = None
        self._j = 0

    def frame_time(self, k):
        return self.t0 + k * self.frame_dt + 0.25 * self.imu_dt

    def imu_rows_for_frame(self, k):
        """IMU rows the EuRoC feed order pushes before image k (rows up to and including the first
        one with t > t_img): array [n][rows][7]."""
        t_img = self.frame_time(k)
        j1 = self._j
        while True:  # same stamps for every stream
            t = self.t0 + j1 * self.imu_dt
            j1 += 1
            if not (t <= t_img):
                break
        out = np.zeros((self.n, j1 - self._j, 7))
        lib().synth_imu_block(self.trajs.ctypes.data, self.n, self._j, j1, self.imu_dt, self.ng, self.na, out.ctypes.data)
        self._j = j1
        return out

    def pose(self, i, t):
        R = np.zeros(9)
        p = np.zeros(3)
        lib().synth_traj_pose(self.trajs[i * self.tsz:].ctypes.data, t, R.ctypes.data, p.ctypes.data)
        return R.reshape(3, 3), p

    def render_device(self, k, out, cuda_stream=0):
        """Render frame k of every stream into the CUDA uint8 tensor `out` [n][2][rows*cols]."""
        import torch

        if self._dev is None:
            L = lib()
            csz = L.synth_cam_size()
            cams = np.concatenate([np.ctypeslib.as_array(C.cast(L.synth_cam(self.proto.h, c), C.POINTER(C.c_uint8)), shape=(csz,)).copy()
                                   for c in (0, 1)])
            dev = out.device
            self._dev = dict(traj=torch.from_numpy(self.trajs.copy()).to(dev), cams=torch.from_numpy(cams).to(dev),
                             r0=torch.from_numpy(self.proto.rays(0)).to(dev), r1=torch.from_numpy(self.proto.rays(1)).to(dev),
                             t=torch.zeros(self.n, dtype=torch.float64, device=dev))
        d = self._dev
        # the time stamps must be written on the stream the render kernel runs on
        ctx = torch.cuda.stream(torch.cuda.ExternalStream(cuda_stream)) if cuda_stream else torch.cuda.stream(torch.cuda.current_stream())
        with ctx:
            d["t"].fill_(self.frame_time(k))
        rc = cuda_lib().mskf_synth_render_device(d["traj"].data_ptr(), d["cams"].data_ptr(), d["r0"].data_ptr(), d["r1"].data_ptr(),
                                                   d["t"].data_ptr(), out.data_ptr(), self.n, self.cfg.img_rows, self.cfg.img_cols,
                                                   C.c_void_p(cuda_stream))
        if rc != 0:
            raise RuntimeError("device render failed")
