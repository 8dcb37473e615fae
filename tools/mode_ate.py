#!/usr/bin/env python
"""ATE of the GPU engine on synthetic streams under the reference's defects and with them fixed
(SURVEY 8f rank 4): compat (message never cleared, F4; previous-image alias, F6), each fix alone, both,
and both + two-point RANSAC.  One handle drives all seeds of a mode as independent streams.

    python tools/mode_ate.py [frames=600] [seeds=4] [preset=ref]      (GPU)
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from msckf_stereo_c_b200 import engine, synth, abi


def ate(est, gt):
    ma, mb = est.mean(0), gt.mean(0)
    U, _, Vt = np.linalg.svd((est - ma).T @ (gt - mb))
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1, 1, d]) @ U.T
    return float(np.sqrt((np.linalg.norm((R @ (est - ma).T).T + mb - gt, axis=1) ** 2).mean()))


def run(cfg, seeds, frames):
    import torch

    fl = synth.Fleet(cfg, seeds)
    n = len(seeds)
    e = engine.Engine(cfg, n)
    img = torch.empty((n, 2, cfg.img_rows * cfg.img_cols), dtype=torch.uint8, device="cuda")
    est = [[] for _ in seeds]
    gt = [[] for _ in seeds]
    for k in range(frames):
        e.push_imu_batch(fl.imu_rows_for_frame(k))
        fl.render_device(k, img)
        torch.cuda.synchronize()
        t = fl.frame_time(k)
        e.push_stereo_batch(np.full(n, t), img.data_ptr(), img.data_ptr() + cfg.img_rows * cfg.img_cols, 2 * cfg.img_rows * cfg.img_cols, device=True)
        e.step()
        e.sync()
        for i in range(n):
            s = e.state(i)
            if s.n_cam_states:
                est[i].append(np.array(s.position[:]))
                gt[i].append(fl.pose(i, t)[1])
    resets = sum(e.state(i).n_resets for i in range(n))
    e.close()
    return [ate(np.array(a), np.array(g)) for a, g in zip(est, gt)], resets


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    preset = sys.argv[3] if len(sys.argv) > 3 else "ref"
    seeds = list(range(n_seeds))
    modes = [("compat (reference behaviour)", dict()),
             ("F4 fixed (fresh message per frame)", dict(compat_stale_features=0)),
             ("F6 fixed (previous image kept)", dict(fix_prev_image_alias=1)),
             ("F4 + F6 fixed", dict(compat_stale_features=0, fix_prev_image_alias=1)),
             ("F4 + F6 fixed + twoPointRansac", dict(compat_stale_features=0, fix_prev_image_alias=1, use_ransac=1))]
    out = []
    for name, kw in modes:
        cfg = abi.Config.from_buffer_copy(bytes(synth.default_config(preset)))
        for k, v in kw.items():
            setattr(cfg, k, v)
        a, resets = run(cfg, seeds, frames)
        row = {"mode": name, "ate_rmse_m": [round(x, 4) for x in a], "mean": round(float(np.mean(a)), 4), "resets": resets}
        print(json.dumps(row))
        out.append(row)
    print(json.dumps({"frames": frames, "seeds": seeds, "preset": preset}))


if __name__ == "__main__":
    main()
