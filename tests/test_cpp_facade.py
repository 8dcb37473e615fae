"""The C++ host façade (include/msckf_b200.hpp) that mirrors the reference's ImageProcessor /
MsckfVio / System API over the C ABI.  CPU: it compiles against the C header alone and fails
loudly without a GPU.  GPU: the C++ feed loop reproduces the Python-driven engine bit for bit."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    return os.path.join(ROOT, "examples", "run_stream")


def _dump(path, synth, cfg, seed, n_frames):
    s = synth.Stream(cfg, seed=seed)
    rec = []
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", n_frames, cfg.img_rows, cfg.img_cols))
        j = 0
        for k in range(n_frames):
            t_img, im0, im1 = s.render(k)
            rows = []
            while True:
                t, w, a = s.imu(j)
                j += 1
                rows.append((t, w, a))
                if not (t <= t_img):
                    break
            f.write(struct.pack("<i", len(rows)))
            for t, w, a in rows:
                f.write(struct.pack("<7d", t, *w, *a))
            f.write(struct.pack("<d", t_img))
            f.write(im0.tobytes())
            f.write(im1.tobytes())
            rec.append((rows, t_img, im0, im1))
    return rec


def test_facade_compiles_and_has_no_cpu_fallback(tmp_path, synth):
    import torch

    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = synth.default_config("ref")
    path = str(tmp_path / "dump.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", 0, cfg.img_rows, cfg.img_cols))
    p = subprocess.run([exe, path], capture_output=True, text=True)
    assert p.returncode == 1 and "mskf_create failed" in p.stderr


@pytest.mark.gpu
def test_cpp_feed_loop_equals_python_engine(tmp_path, synth):
    from msckf_stereo_c_b200 import engine

    exe = _build()
    cfg = synth.default_config("ref")
    path = str(tmp_path / "dump.bin")
    rec = _dump(path, synth, cfg, 4, 50)
    out = subprocess.run([exe, path, "ref"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(rec)
    e = engine.Engine(cfg, 1)
    for line, (rows, t_img, im0, im1) in zip(out, rec):
        for t, w, a in rows:
            e.imu_callback(t, w, a)
        e.stereo_callback(t_img, im0, im1)
        e.backend_callback()
        st = e.state()
        v = line.split()
        got = np.array([float(x) for x in v[1:8]])
        want = np.array(list(st.position[:]) + list(st.orientation[:]))
        assert np.array_equal(got, want)
        assert int(v[8]) == st.n_cam_states and int(v[9]) == len(e.features()[1])
    assert e.state().n_cam_states > 10
    e.close()


@pytest.mark.gpu
def test_cpp_standalone_msckfvio_equals_system(tmp_path, synth):
    """MsckfVio driven alone (imuCallback + featureCallback(msg), msckf_vio.cpp:190-207, 306-375) initialises
    gravity from its own IMU buffer and ends in the same state as the one inside System."""
    exe = _build()
    cfg = synth.default_config("ref")
    path = str(tmp_path / "dump.bin")
    _dump(path, synth, cfg, 6, 45)
    a = subprocess.run([exe, path, "ref"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    b = subprocess.run([exe, path, "ref", "vio"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(a) == len(b) == 45
    assert a == b
    assert int(a[-1].split()[8]) > 10


@pytest.mark.gpu
def test_features_head_equals_full_message(synth):
    """mskf_get_features_head returns the part of the never-cleared message that can hold measurements;
    the rest of the reference's vector is value-initialised."""
    from msckf_stereo_c_b200 import engine

    cfg = synth.default_config("ref")
    s = synth.Stream(cfg, seed=2)
    e = engine.Engine(cfg, 1)
    for k, t in synth.feed(s, 12, e):
        t_full, full, n_pub = e.features()
        t_head, head, total = e.features_head()
        assert t_full == t_head and total == len(full) and len(head) <= total and n_pub <= len(head)
        assert full[:len(head)].tobytes() == head.tobytes()
        assert not full[len(head):].tobytes().strip(b"\0")
    assert total > len(head)
    e.close()
