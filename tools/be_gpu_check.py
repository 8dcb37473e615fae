"""Back-end bring-up: feed the oracle's CameraMeasurement to the CUDA EKF (split-phase ABI) and
compare state / covariance after every frame."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from msckf_stereo_c_b200 import synth, engine
from oracle import binding as ob
preset = sys.argv[1] if len(sys.argv) > 1 else "ref"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 60
stale = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = synth.default_config(preset)
cfg.compat_stale_features = stale
s = synth.Stream(cfg, seed=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
e = engine.Engine(cfg, 1)
o = ob.Oracle(cfg)
class Both:
    def imu(self, t, w, a): o.imu(t, w, a); e.imu_callback(t, w, a)
    def stereo(self, t, i0, i1): o.stereo(t, i0, i1)
    def backend(self):
        o.backend()
        t, f, n = o.features()
        e.backend_features(t, f)
worst = 0
for k, t in synth.feed(s, nf, Both()):
    so, sg = o.state(), e.state()
    msg = f"{k} N {so.n_cam_states}/{sg.n_cam_states} map {so.n_map_features}/{sg.n_map_features} upd {so.n_updates}/{sg.n_updates} rst {so.n_resets}/{sg.n_resets} grav {so.is_gravity_set}/{sg.is_gravity_set}"
    if so.n_cam_states == sg.n_cam_states and so.n_cam_states > 0:
        Po, Pg = o.cov(), e.cov()
        dP = np.abs(Po - Pg).max() / np.abs(Po).max()
        dq = np.abs(np.array(so.orientation[:]) - np.array(sg.orientation[:])).max()
        dp = np.abs(np.array(so.position[:]) - np.array(sg.position[:])).max()
        dv = np.abs(np.array(so.velocity[:]) - np.array(sg.velocity[:])).max()
        dbg = np.abs(np.array(so.gyro_bias[:]) - np.array(sg.gyro_bias[:])).max()
        co, cg = o.cam_states(), e.cam_states()
        dc = max(np.abs(co["position"] - cg["position"]).max(), np.abs(co["orientation"] - cg["orientation"]).max()) if np.array_equal(co["id"], cg["id"]) else -1
        worst = max(worst, dP, dq, dp, dv)
        msg += f" dP {dP:.2e} dq {dq:.2e} dp {dp:.2e} dv {dv:.2e} dbg {dbg:.2e} dcam {dc:.2e} tr {so.tracking_rate:.3f}/{sg.tracking_rate:.3f}"
    if so.n_cam_states:
        io, no, po, oo = o.feature_map(); ig, ng, pg, og = e.feature_map()
        if np.array_equal(io, ig) and np.array_equal(no, ng) and np.array_equal(oo, og):
            both = no > 0
            d = np.abs(po - pg).max(1) * both
            top = np.argsort(-d)[:3]
            msg += " top " + " ".join(f"{io[i]}:{d[i]:.1e}" for i in top)
            msg += f" map ok ninit {both.sum()} dpos {np.abs(po - pg)[both].max() if both.any() else 0:.2e} rel {(np.abs(po - pg)[both] / np.abs(po[both]).max()).max() if both.any() else 0:.2e}"
        else:
            msg += f" MAP DIFF ids {np.array_equal(io, ig)} init {np.array_equal(no, ng)} nobs {np.array_equal(oo, og)}"
    print(msg)
print("BACKEND WORST", worst)
