import sys, time, numpy as np
sys.path.insert(0, '.')
from msckf_stereo_c_b200 import synth, engine
from oracle import binding as ob
cfg = synth.default_config(sys.argv[1] if len(sys.argv) > 1 else "ref")
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 25
s = synth.Stream(cfg, seed=0)
e = engine.Engine(cfg, 1)
o = ob.Oracle(cfg)
# operators
t, a, b = s.render(40); t2, a2, b2 = s.render(41)
pg = e.op_pyramid(np.stack([a, b]), cfg.pyramid_levels)
ref = a
for l in range(cfg.pyramid_levels - 1):
    ref = ob.pyr_down(ref); print("pyr level", l + 1, "equal", np.array_equal(ref, pg[0][l]))
xy_o, r_o = ob.detect(cfg, a); xy_g, r_g = e.op_detect(a)
print("detect n", len(xy_o), len(xy_g), "equal", np.array_equal(xy_o, xy_g) and np.array_equal(r_o, r_g))
pb_o, st_o = ob.klt(cfg, a, a2, xy_o, xy_o); pb_g, st_g = e.op_klt(a, a2, xy_o, xy_o)
print("klt status agree", (st_o == st_g).mean(), "tracked", st_o.sum(), "max |d|", np.abs(pb_o - pb_g)[st_o > 0].max(), "bit-equal", np.array_equal(pb_o, pb_g))
# pipeline
class Both:
    def imu(self, t, w, a): o.imu(t, w, a); e.imu_callback(t, w, a)
    def stereo(self, t, i0, i1): o.stereo(t, i0, i1); e.stereo_callback(t, i0, i1)
    def backend(self): pass
bad = 0
for k, t in synth.feed(s, nf, Both()):
    go, gg = o.grid(), e.grid()
    to, fo, no = o.features(); tg, fg, ng = e.features()
    io, ig = o.tracking_info(), e.tracking_info()
    same_grid = len(go) == len(gg) and all(np.array_equal(go[f], gg[f]) for f in ("id", "lifetime", "cam0", "cam1", "cell", "response"))
    same_msg = len(fo) == len(fg) and fo.tobytes() == fg.tobytes() and no == ng
    same_info = (io.before_tracking, io.after_tracking, io.after_matching, io.after_ransac) == (ig.before_tracking, ig.after_tracking, ig.after_matching, ig.after_ransac)
    p_ok = all(np.array_equal(o.pyramid(c, l), e.pyramid(c, l)) for c in (0, 1) for l in range(cfg.pyramid_levels))
    if not (same_grid and same_msg and same_info and p_ok): bad += 1
    print(k, "n", len(go), len(gg), "grid", same_grid, "msg", same_msg, len(fo), len(fg), "info", same_info, "pyr", p_ok)
print("FRONTEND PARITY", "OK" if bad == 0 else f"FAIL {bad}")
