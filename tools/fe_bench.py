"""Front-end only timing (kernel bring-up tool; bench.py is the contract benchmark).
usage: python tools/fe_bench.py [preset=bench] [streams=64] [frames=12] [blur_sigma=0]
blur_sigma > 0 low-pass filters the synthetic frames: sigma 2.0 brings the FAST corner density from the synthetic
texture's 26 % of the pixels down to the 2-5 % of real imagery, the regime where the compacted stages of the
detector (arc score, NMS, Shi-Tomasi) are nearly free."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
import torch
from msckf_stereo_c_b200 import synth, engine
preset = sys.argv[1] if len(sys.argv) > 1 else "bench"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
cfg = synth.default_config(preset)
streams = [synth.Stream(cfg, seed=i) for i in range(min(S, 8))]
frames = []
for k in range(30, 30 + nf):
    per = [s.render(k) for s in streams]
    if sigma > 0:
        from scipy.ndimage import gaussian_filter
        blur = lambda im: np.clip(np.rint(gaussian_filter(im.astype(np.float32), sigma)), 0, 255).astype(np.uint8)
        per = [(t, blur(a), blur(b)) for (t, a, b) in per]
    a = np.stack([per[i % len(per)][1] for i in range(S)]); b = np.stack([per[i % len(per)][2] for i in range(S)])
    frames.append((per[0][0], torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()))
e = engine.Engine(cfg, S)
torch.cuda.synchronize()
def step(k):
    t, a, b = frames[k]
    sz = cfg.img_rows * cfg.img_cols
    for i in range(S):
        e.push_stereo_ptr(t, a.data_ptr() + i * sz, b.data_ptr() + i * sz, stream=i, device=True)
    e.frontend_step()
for k in range(4): step(k)
e.sync()
t0 = time.time()
for k in range(4, nf): step(k)
e.sync()
dt = (time.time() - t0) / (nf - 4)
print(f"preset {preset} S {S}: {dt*1e3:.3f} ms/step  {S/dt:.0f} stereo frames/s (front end only)  n_feat {len(e.grid(0))}")
e.profile_enable(True)
for k in range(4, nf): step(k)
e.sync()
pr = e.profile_read()
e.profile_enable(False)
print("  " + " ".join(f"{k}={v[0] / max(v[1], 1) * 1e3:.0f}us" for k, v in pr.items() if v[1] > 0))
