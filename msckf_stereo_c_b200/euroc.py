"""EuRoC "mav0" files for the synthetic streams, and the trajectory evaluation the reference's README
uses (TUM rgbd_benchmark_tools absolute trajectory error: README.md:53-88).

  write_mav0   BASELINE.json config 1: cam{0,1}/data.csv (CRLF line ends, the runner strips the last
               character of the file-name field: run_euroc_single_thread.cpp:168), cam{0,1}/data/*.png,
               imu0/data.csv, and the ground truth as a TUM file (time tx ty tz qx qy qz qw)
  read_tum / ate   association by time stamp, Horn alignment (rotation + translation), RMSE / mean /
               median / std / min / max of the translational error
Input / evaluation tooling: nothing here is on the hot path."""
import os

import numpy as np


def _quat_hamilton(R):
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w, x, y, z = (R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w, x, y, z = (R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w, x, y, z = (R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s
    return x, y, z, w


def write_mav0(stream, n_frames, out_dir, image_format="png"):
    """Writes the first `n_frames` frames of a synth.Stream as an EuRoC mav0 directory.  Time stamps are
    integer nanoseconds (10-digit seconds as in EuRoC; std::stoi needs seconds < 2^31)."""
    import cv2

    for cam in (0, 1):
        os.makedirs(os.path.join(out_dir, f"cam{cam}", "data"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "imu0"), exist_ok=True)
    base_ns = 1403636579 * 10 ** 9  # a EuRoC-like epoch
    to_ns = lambda t: base_ns + int(round((t - stream.t0) * 1e9))
    rows = [[], []]
    gt = []
    j = 0
    imu_lines = ["#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]"]
    for k in range(n_frames):
        t_img, im0, im1 = stream.render(k)
        ns = to_ns(t_img)
        for cam, im in ((0, im0), (1, im1)):
            name = f"{ns}.{image_format}"
            path = os.path.join(out_dir, f"cam{cam}", "data", name)
            if image_format == "png":
                cv2.imwrite(path, im)
            else:
                with open(path, "wb") as f:
                    f.write(b"P5\n%d %d\n255\n" % (im.shape[1], im.shape[0]) + im.tobytes())
            rows[cam].append(f"{ns},{name}")
        while True:
            t, w, a = stream.imu(j)
            j += 1
            imu_lines.append(f"{to_ns(t)}," + ",".join(repr(float(np.float32(v))) for v in list(w) + list(a)))
            if not (t <= t_img):
                break
        R, p = stream.pose(t_img)
        gt.append((ns * 1e-9, *p, *_quat_hamilton(R)))
    for cam in (0, 1):
        with open(os.path.join(out_dir, f"cam{cam}", "data.csv"), "w", newline="") as f:
            f.write("#timestamp [ns],filename\r\n" + "".join(r + "\r\n" for r in rows[cam]))
    with open(os.path.join(out_dir, "imu0", "data.csv"), "w", newline="") as f:
        f.write("\n".join(imu_lines) + "\n")
    with open(os.path.join(out_dir, "groundtruth_tum.txt"), "w") as f:
        for r in gt:
            f.write(" ".join(f"{v:.9f}" for v in r) + "\n")
    return out_dir


def read_tum(path):
    rows = [[float(x) for x in line.split()] for line in open(path) if line.strip() and not line.startswith("#")]
    return np.array(rows).reshape(-1, 8)


def ate(est, gt, max_dt=0.02):
    """Absolute trajectory error the way TUM's evaluate_ate.py computes it: nearest-stamp association
    within `max_dt`, Horn alignment of the estimate onto the ground truth, translational error stats."""
    est, gt = np.asarray(est), np.asarray(gt)
    idx = np.searchsorted(gt[:, 0], est[:, 0])
    idx = np.clip(idx, 1, len(gt) - 1)
    left = np.abs(gt[idx - 1, 0] - est[:, 0]) < np.abs(gt[idx, 0] - est[:, 0])
    idx = np.where(left, idx - 1, idx)
    ok = np.abs(gt[idx, 0] - est[:, 0]) <= max_dt
    a, b = est[ok, 1:4], gt[idx[ok], 1:4]
    if len(a) < 3:
        raise ValueError("not enough matching poses")
    ma, mb = a.mean(0), b.mean(0)
    U, _, Vt = np.linalg.svd((a - ma).T @ (b - mb))
    S = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ S @ U.T
    err = np.linalg.norm((R @ (a - ma).T).T + mb - b, axis=1)
    return {"pairs": int(len(err)), "rmse": float(np.sqrt((err ** 2).mean())), "mean": float(err.mean()),
            "median": float(np.median(err)), "std": float(err.std()), "min": float(err.min()), "max": float(err.max())}
