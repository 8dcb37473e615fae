import sys, numpy as np
sys.path.insert(0, '.')
from msckf_stereo_c_b200 import synth, engine
from oracle import binding as ob
cfg = synth.default_config("ref")
s = synth.Stream(cfg, seed=0)
e = engine.Engine(cfg, 1)
t, a, b = s.render(40)
xy_o, r_o, sm = ob.detect(cfg, a, want_scores=True); xy_g, r_g, smg = e.debug_detect_scores(a)
print("score maps equal", np.array_equal(sm, smg), (sm != smg).sum())
ys, xs = np.nonzero(sm != smg)
for y, x in list(zip(ys, xs))[:20]: print((x, y), sm[y, x], smg[y, x])
print(sm[9:16, 21:29]); print(smg[9:16, 21:29])
