#!/usr/bin/env python
"""ATE of a TUM trajectory file against a ground-truth TUM file, printed like the reference README
(README.md:57-87):  python tools/ate.py pose_out.txt groundtruth_tum.txt"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from msckf_stereo_c_b200 import euroc
r = euroc.ate(euroc.read_tum(sys.argv[1]), euroc.read_tum(sys.argv[2]))
print(f"compared_pose_pairs {r['pairs']} pairs")
for k in ("rmse", "mean", "median", "std", "min", "max"):
    print(f"absolute_translational_error.{k} {r[k]:.6f} m")
