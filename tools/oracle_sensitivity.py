"""The reference algorithm's own noise floor: two CPU oracles on the same synthetic stream, the second
with every injected measurement moved by one ulp (u0 only).  Prints, per frame, the relative deviation of
the covariance, the absolute deviation of the IMU position and of the triangulated feature positions.
Feature::initializePosition accepts an LM step iff new_cost < total_cost (feature.hpp:417); for the last
steps (|delta| ~ 1e-9) that comparison is decided by rounding, so a one-ulp change of any input re-rolls
it.  Usage: python tools/oracle_sensitivity.py [preset] [frames] [seed]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from msckf_stereo_c_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else "bench"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 62
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = synth.default_config(preset)
s = synth.Stream(cfg, seed=seed)
a, b = ob.Oracle(cfg), ob.Oracle(cfg)


class Both:
    def imu(self, t, w, acc):
        a.imu(t, w, acc)
        b.imu(t, w, acc)

    def stereo(self, t, i0, i1):
        a.stereo(t, i0, i1)

    def backend(self):
        a.backend()
        t, f, _ = a.features()
        g = f.copy()
        g["u0"] = np.nextafter(g["u0"], np.inf)
        b.backend_features(t, g)


worst = [0.0, 0.0, 0.0]
for k, t in synth.feed(s, nf, Both()):
    sa, sb = a.state(), b.state()
    if not sa.n_cam_states or sa.n_cam_states != sb.n_cam_states:
        continue
    Pa, Pb = a.cov(), b.cov()
    dP = np.abs(Pa - Pb).max() / np.abs(Pa).max()
    dp = np.abs(np.array(sa.position[:]) - np.array(sb.position[:])).max()
    ia, na, pa, _ = a.feature_map()
    ib, nb, pb, _ = b.feature_map()
    dpos = 0.0
    if np.array_equal(ia, ib) and np.array_equal(na, nb) and na.any():
        dpos = np.abs(pa - pb)[na == 1].max()
    worst = [max(worst[0], dP), max(worst[1], dp), max(worst[2], dpos)]
    print(f"{k} upd {sa.n_updates}/{sb.n_updates} dP {dP:.2e} dp {dp:.2e} dpos {dpos:.2e}")
print("WORST dP %.2e dp %.2e dpos %.2e" % tuple(worst))
