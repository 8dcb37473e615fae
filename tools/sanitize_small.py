"""Small end-to-end run for compute-sanitizer (memcheck): 2 streams, ref preset, through pruning."""
import sys, numpy as np
sys.path.insert(0, '.')
from msckf_stereo_c_b200 import synth, engine
cfg = synth.default_config("ref")
ss = [synth.Stream(cfg, seed=i) for i in range(2)]
e = engine.Engine(cfg, 2)
js = [0, 0]
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 44):
    for i, s in enumerate(ss):
        t_img, a, b = s.render(k)
        while True:
            t, w, acc = s.imu(js[i]); js[i] += 1
            e.imu_callback(t, w, acc, stream=i)
            if not (t <= t_img): break
        e.push_stereo(t_img, a, b, stream=i)
    e.step()
e.sync()
st = e.state(1)
print("ok", st.n_cam_states, st.n_updates, np.isfinite(e.cov(1)).all())
