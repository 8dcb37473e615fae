"""Micro-benchmark of the update chain on one stream: per-kernel time for a dense m x k system.
usage: python tools/upd_micro.py [m n_cam]   (GPU)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from msckf_stereo_c_b200 import engine


def main():
    cfg = engine.default_config("bench")
    e = engine.Engine(cfg, 1)
    rng = np.random.default_rng(0)
    sizes = [(100, 29), (200, 29), (320, 29), (480, 29), (800, 29), (1300, 2), (320, 15), (800, 15)]
    if len(sys.argv) > 2:
        sizes = [(int(sys.argv[1]), int(sys.argv[2]))]
    for (m, nc) in sizes:
        n = 21 + 6 * nc
        H = np.zeros((m, n))
        H[:, 21:] = rng.standard_normal((m, 6 * nc))
        r = rng.standard_normal(m) * 0.01
        A = rng.standard_normal((n, n))
        P = A @ A.T * 1e-3 + np.eye(n) * 1e-2
        e.op_ekf_update(H, r, P)
        e.profile_enable(True)
        for _ in range(5):
            e.op_ekf_update(H, r, P)
        e.sync()
        pr = e.profile_read()
        e.profile_enable(False)
        row = " ".join(f"{k}={v[0] / max(v[1], 1) * 1e3:.0f}us" for k, v in pr.items() if v[1] > 0 and k.startswith("be_"))
        print(f"m={m} k={6 * nc}: {row}")
    e.close()


if __name__ == "__main__":
    main()
