"""SURVEY section 8f "next" rows 1-2: EuRoC mav0 reader + feed protocol (examples/run_euroc.cpp over
the C++ façade), TUM trajectory output, ATE evaluation."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    return os.path.join(ROOT, "examples", "run_euroc")


def test_png_and_pgm_decoders_match_opencv(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    # smooth + noisy content so that the encoder uses all five PNG filter types
    img = (np.add.outer(np.arange(97), np.arange(131)) % 256).astype(np.uint8)
    img[20:60] = rng.integers(0, 256, (40, 131), dtype=np.uint8)
    exe = _exe()
    for name in ("a.png", "a.pgm"):
        path = str(tmp_path / name)
        assert cv2.imwrite(path, img)
        out = str(tmp_path / (name + ".raw"))
        dims = subprocess.check_output([exe, "--dump-image", path, out], text=True).split()
        assert [int(x) for x in dims] == list(img.shape)
        assert np.array_equal(np.fromfile(out, np.uint8).reshape(img.shape), cv2.imread(path, 0))


def test_ate_recovers_a_rigid_offset():
    from msckf_stereo_c_b200 import euroc

    t = np.linspace(0, 10, 200)
    gt = np.stack([t, np.sin(t), np.cos(0.5 * t), 0.1 * t, 0 * t, 0 * t, 0 * t, 1 + 0 * t], 1)
    c, s = np.cos(0.7), np.sin(0.7)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    est = gt.copy()
    est[:, 1:4] = gt[:, 1:4] @ R.T + [3.0, -2.0, 0.5]
    r = euroc.ate(est, gt)
    assert r["pairs"] == 200 and r["rmse"] < 1e-12  # a rigid transform is aligned away
    est[:, 1] += 0.01 * np.sin(7 * t)
    assert 1e-3 < euroc.ate(est, gt)["rmse"] < 1e-2


@pytest.mark.gpu
def test_run_euroc_on_synthetic_mav0(tmp_path, synth):
    """mav0 on disk -> C++ runner -> pose_out.txt; the same files driven through the Python engine give
    the same trajectory, and the ATE against the ground truth is small."""
    import cv2

    from msckf_stereo_c_b200 import engine, euroc

    cfg = synth.default_config("ref")
    s = synth.Stream(cfg, seed=6)
    nf = 80
    d = euroc.write_mav0(s, nf, str(tmp_path / "mav0"))
    out = str(tmp_path / "pose_out.txt")
    subprocess.run([_exe(), d, "ref", out, "9"], check=True, capture_output=True)
    est = euroc.read_tum(out)
    # the same feed through ctypes, reading the files back exactly like the runner
    e = engine.Engine(cfg, 1)
    cam = [[l.strip().split(",") for l in open(os.path.join(d, f"cam{c}", "data.csv"), newline="").read().split("\r\n")[1:] if l] for c in (0, 1)]
    imu = [l.split(",") for l in open(os.path.join(d, "imu0", "data.csv")).read().splitlines()[1:]]
    stamp = lambda sns: (int(sns[:-9]) * 1e9 + int(sns[-9:])) * 1e-9
    j, poses = 0, []
    for k in range(nf):
        t_img = stamp(cam[0][k][0])
        while True:
            row = imu[j]
            j += 1
            t = stamp(row[0])
            e.imu_callback(t, [float(np.float32(v)) for v in row[1:4]], [float(np.float32(v)) for v in row[4:7]])
            if not (t <= t_img):
                break
        im = [cv2.imread(os.path.join(d, f"cam{c}", "data", cam[c][k][1]), 0) for c in (0, 1)]
        e.stereo_callback(t_img, im[0], im[1])
        e.backend_callback()
        st = e.state()
        if st.is_gravity_set:
            T = np.array(st.T_b_w[:]).reshape(4, 4)
            poses.append([t_img, T[0, 3], T[1, 3], T[2, 3]])
    poses = np.array(poses)
    assert len(poses) == len(est)
    assert np.abs(est[:, :4] - poses).max() < 2e-9  # printed with 9 decimals
    r = euroc.ate(est, euroc.read_tum(os.path.join(d, "groundtruth_tum.txt")))
    assert r["pairs"] == len(est) and r["rmse"] < 0.1
    e.close()


@pytest.mark.gpu
def test_run_euroc_fleet_equals_single_stream_runs(tmp_path, synth):
    """several mav0 directories on disk -> decoder threads -> page-locked ring -> one
    asynchronous upload per step -> one engine handle.  Every stream's trajectory file is byte-identical to
    what the single-stream runner writes for the same directory (a stream's result does not depend on its
    neighbours, on the ring depth or on the number of decoder threads)."""
    import json

    from msckf_stereo_c_b200 import euroc

    exe = _exe()
    fleet = os.path.join(ROOT, "examples", "run_euroc_fleet")
    cfg = synth.default_config("ref")
    nf = 60
    dirs, single = [], []
    for i, (seed, fmt) in enumerate(((6, "png"), (11, "pgm"))):
        d = euroc.write_mav0(synth.Stream(cfg, seed=seed), nf, str(tmp_path / f"mav0_{i}"), image_format=fmt)
        out = str(tmp_path / f"single_{i}.txt")
        subprocess.run([exe, d, "ref", out, "9"], check=True, capture_output=True)
        dirs.append(d)
        single.append(open(out).read())
        assert len(single[-1].splitlines()) > 10
    for threads, ring in ((3, 2), (8, 4)):
        od = tmp_path / f"fleet_{threads}_{ring}"
        od.mkdir()
        r = subprocess.run([fleet, "--threads", str(threads), "--ring", str(ring), "--repeat", "2", "--preset", "ref", "--decimals", "9",
                            "--out", str(od)] + dirs, check=True, capture_output=True, text=True)
        rep = json.loads(r.stdout.strip().splitlines()[-1])
        assert rep["streams"] == 4 and rep["steps"] == nf and rep["disk_to_pose_frames_per_s"] > 0
        for s in range(4):
            assert open(od / f"pose_{s}.txt").read() == single[s // 2], (threads, ring, s)
