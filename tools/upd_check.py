"""Operator parity on the filter's own updates: runs the CPU oracle on a synthetic stream and feeds every
(H, r, P-) it hands to measurementUpdate (msckf_vio.cpp:778-907) through mskf_op_ekf_update, printing the
size of the system and the deviation of the posterior.  This is the check that exposed a QR sweep that
stopped a few columns early (3e-7 at one update of the ref preset).

    python tools/upd_check.py [preset=bench] [seed=1] [frames=56]      (GPU)
"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msckf_stereo_c_b200 import engine as eng, synth, abi
from oracle import binding as ob
cfg = abi.Config.from_buffer_copy(bytes(synth.default_config(sys.argv[1] if len(sys.argv) > 1 else "bench")))
cfg.compat_stale_features = 0
s = synth.Stream(cfg, seed=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
o = ob.Oracle(cfg); o.keep_last_update()
e = eng.Engine(cfg, 1)
class Sink:
    def imu(self, t, w, a): o.imu(t, w, a)
    def stereo(self, t, a, b): o.stereo(t, a, b)
    def backend(self): o.backend()
last = 0
for k, t in synth.feed(s, int(sys.argv[3]) if len(sys.argv) > 3 else 56, Sink()):
    st = o.state()
    if st.n_updates == last: continue
    last = st.n_updates
    H, r, P = o.last_update()
    dx_o, P_o = ob.update_math(H, r, P, cfg.noise_feature ** 2)
    dx_g, P_g = e.op_ekf_update(H, r, P)
    nz = np.abs(H).sum(0) > 0
    print(k, 'm', H.shape[0], 'k', int(nz[21:].sum()), 'dP %.2e' % (np.abs(P_g - P_o).max() / np.abs(P_o).max()), 'dx %.2e' % np.abs(dx_g - dx_o).max(), flush=True)
