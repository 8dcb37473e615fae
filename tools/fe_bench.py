"""Front-end only timing (kernel bring-up tool; bench.py is the contract benchmark)."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
import torch
from msckf_stereo_c_b200 import synth, engine
preset = sys.argv[1] if len(sys.argv) > 1 else "bench"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 12
cfg = synth.default_config(preset)
streams = [synth.Stream(cfg, seed=i) for i in range(min(S, 8))]
frames = []
for k in range(30, 30 + nf):
    per = [s.render(k) for s in streams]
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
