"""Bring-up tool: run a fleet and report the first non-finite pose (stream, frame)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from msckf_stereo_c_b200 import synth, engine
S = int(sys.argv[1]); nf = int(sys.argv[2]); overlap = int(sys.argv[3])
cfg = synth.default_config("bench")
fleet = synth.Fleet(cfg, list(range(S)))
img = cfg.img_rows * cfg.img_cols
stream = torch.cuda.Stream()
e = engine.Engine(cfg, S, cuda_stream=stream.cuda_stream)
e.set_overlap(bool(overlap))
scratch = torch.empty((S, 2, img), dtype=torch.uint8, device="cuda")
tvec = np.zeros(S)
bad = None
for k in range(nf):
    e.push_imu_batch(fleet.imu_rows_for_frame(k))
    fleet.render_device(k, scratch, stream.cuda_stream)
    tvec[:] = fleet.frame_time(k)
    e.push_stereo_batch(tvec, scratch.data_ptr(), scratch.data_ptr() + img, 2 * img, device=True)
    e.step()
    e.sync()
    p = e.poses()
    fin = np.isfinite(p).all(axis=(1, 2))
    if k >= nf - 6:
        d = e.update_dims()
        m0, k0, n0 = d[:, 0, 0], d[:, 0, 1], d[:, 0, 2]
        m1, k1, n1 = d[:, 1, 0], d[:, 1, 1], d[:, 1, 2]
        q = lambda a: [int(x) for x in np.percentile(a, [0, 25, 50, 75, 90, 99, 100])]
        print(k, "lost m pct", q(m0), "k", q(k0), "nlist", q(n0), "| prune m", q(m1), "streams pruning", int((m1 > 0).sum()), "nlist", q(n1[n1 > 0]) if (n1 > 0).any() else [])
    if not fin.all():
        bad = (k, np.nonzero(~fin)[0][:8])
        break
print("overlap", overlap, "first non-finite:", bad)
if bad:
    s0 = int(bad[1][0]); st = e.state(s0)
    print("stream", s0, "N", st.n_cam_states, "upd", st.n_updates, "resets", st.n_resets, "pos", st.position[:], "map", st.n_map_features)
