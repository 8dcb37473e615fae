"""Experiment: H handles x (256/H) streams on one GPU, steps interleaved, wall-clock between syncs."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from msckf_stereo_c_b200 import synth, engine
H = int(sys.argv[1]); S_total = int(sys.argv[2]) if len(sys.argv) > 2 else 256; K = 30; PRIME = 72
cfg = synth.default_config("bench")
S = S_total // H
img = cfg.img_rows * cfg.img_cols
fleets = [synth.Fleet(cfg, list(range(h * S, (h + 1) * S))) for h in range(H)]
streams = [torch.cuda.Stream() for _ in range(H)]
engs = [engine.Engine(cfg, S, cuda_stream=streams[h].cuda_stream) for h in range(H)]
scratch = [torch.empty((S, 2, img), dtype=torch.uint8, device="cuda") for _ in range(H)]
tvec = np.zeros(S)
def step(h, k, buf):
    engs[h].push_imu_batch(fleets[h].imu_rows_for_frame(k))
    tvec[:] = fleets[h].frame_time(k)
    engs[h].push_stereo_batch(tvec, buf.data_ptr(), buf.data_ptr() + img, 2 * img, device=True)
    engs[h].step()
for k in range(PRIME):
    for h in range(H):
        fleets[h].render_device(k, scratch[h], streams[h].cuda_stream)
        step(h, k, scratch[h])
frames = [torch.empty((K, S, 2, img), dtype=torch.uint8, device="cuda") for _ in range(H)]
for h in range(H):
    for i in range(K):
        fleets[h].render_device(PRIME + i, frames[h][i], streams[h].cuda_stream)
for e in engs: e.sync()
torch.cuda.synchronize()
for i in range(5):
    for h in range(H): step(h, PRIME + i, frames[h][i])
for e in engs: e.sync()
t0 = time.perf_counter()
for i in range(5, K):
    for h in range(H): step(h, PRIME + i, frames[h][i])
for e in engs: e.sync()
dt = (time.perf_counter() - t0) / (K - 5)
print(f"handles {H} x {S} streams: {dt*1e3:.3f} ms/step  {S_total/dt:.0f} frames/s")
st = engs[0].state(0); d = engs[0].update_dims()
print("state0: N", st.n_cam_states, "upd", st.n_updates, "feat", len(engs[0].grid(0)), "launches", engs[0].launch_count(), "lost m max", d[:, 0, 0].max(), "prune m max", d[:, 1, 0].max())
