"""Regenerates the committed golden vectors under tests/golden/.

The reference (mfkiwl/msckf_stereo_c) ships no tests and its pixel / point primitives live
in the un-vendored, unpinned vikit_cg, so there are no reference-side golden
vectors to pin.  These fixtures pin the SPEC instead: each oracle primitive adopts the
semantics of the OpenCV call that the reference's commented-out code names at that call site
(image_processor.cpp:217-227 buildOpticalFlowPyramid/pyrDown, :130 FastFeatureDetector,
:399-408 calcOpticalFlowPyrLK, :809-816 undistortPoints / fisheye::undistortPoints,
:837-844 projectPoints / fisheye::distortPoints), and the vectors below are the outputs of
those OpenCV calls (python cv2 4.13, present in the build container only).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from msckf_stereo_c_b200 import synth  # noqa: E402


def main():
    cfg = synth.default_config("ref")
    s = synth.Stream(cfg, seed=3)
    _, a, _ = s.render(30)
    _, a2, _ = s.render(31)
    # --- pyramid + FAST on a small odd-sized crop (odd sizes exercise the (n+1)/2 rule and
    # BORDER_REFLECT_101) ------------------------------------------------------------------
    crop = np.ascontiguousarray(a[100:223, 200:357])  # 123 x 157
    lv = [crop]
    for _ in range(3):
        lv.append(cv2.pyrDown(lv[-1]))
    f_all = cv2.FastFeatureDetector_create(threshold=10, nonmaxSuppression=False).detect(crop)
    f_nms = cv2.FastFeatureDetector_create(threshold=10, nonmaxSuppression=True).detect(crop)
    all_xy = np.array([[int(k.pt[0]), int(k.pt[1])] for k in f_all], np.int32)
    nms = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in f_nms], np.int32)
    np.savez_compressed(os.path.join(HERE, "cv2_pyr_fast.npz"), img=crop, l1=lv[1], l2=lv[2], l3=lv[3],
                        fast_all_xy=all_xy, fast_nms_xyr=nms)
    # --- point maps ---------------------------------------------------------------------------
    rng = np.random.default_rng(7)
    pts = np.stack([rng.uniform(0, 751, 64), rng.uniform(0, 479, 64)], 1).astype(np.float32)
    out = {"pts": pts}
    for cam, (K, D) in enumerate(((cfg.cam0_intrinsics, cfg.cam0_distortion), (cfg.cam1_intrinsics, cfg.cam1_distortion))):
        K = np.array(K[:])
        D = np.array(D[:])
        Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]])
        und = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, D).reshape(-1, 2)
        R = cv2.Rodrigues(np.array([0.01, -0.02, 0.005]))[0]
        und_R = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, D, R=R).reshape(-1, 2)
        xyz = np.concatenate([und, np.ones((len(und), 1), np.float32)], 1).astype(np.float64)
        dist = cv2.projectPoints(xyz, np.zeros(3), np.zeros(3), Km, D)[0].reshape(-1, 2)
        Df = D * 0.1  # a mild equidistant model (EuRoC itself is radtan)
        fund = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2).astype(np.float64), Km, Df).reshape(-1, 2)
        fdist = cv2.fisheye.distortPoints(fund.reshape(-1, 1, 2), Km, Df).reshape(-1, 2)
        out.update({f"K{cam}": K, f"D{cam}": D, f"R{cam}": R, f"und{cam}": und, f"undR{cam}": und_R, f"dist{cam}": dist,
                    f"Df{cam}": Df, f"fund{cam}": fund, f"fdist{cam}": fdist})
    out["rodrigues_v"] = np.array([[0.3, -0.2, 0.1], [1e-14, 0, 0], [0, 2.5, -1.0]])
    out["rodrigues_R"] = np.stack([cv2.Rodrigues(v)[0] for v in out["rodrigues_v"]])
    np.savez_compressed(os.path.join(HERE, "cv2_points.npz"), **out)
    # --- KLT sanity vector (float LK of OpenCV vs the SPEC's fixed-point LK: close, not equal)
    A = np.ascontiguousarray(a[40:296, 120:504])  # 256 x 384
    B = np.ascontiguousarray(a2[40:296, 120:504])
    p0 = cv2.goodFeaturesToTrack(A, 120, 0.01, 8).reshape(-1, 2).astype(np.float32)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(A, B, p0.reshape(-1, 1, 2), p0.reshape(-1, 1, 2).copy(), winSize=(15, 15), maxLevel=3,
                                         criteria=(cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                         flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    np.savez_compressed(os.path.join(HERE, "cv2_klt.npz"), a=A, b=B, p0=p0, p1=p1.reshape(-1, 2), st=st.ravel())
    # --- KLT under real motion: temporal pair (frames 70 -> 72, flow ~5 px) and stereo pair (cam0 -> cam1,
    # initial guess -20 px, flow ~13 px) on a 256 x 384 crop; cv2 with a tight termination so that the
    # vectors are OpenCV's fixed points, not its 0.01 px stopping noise
    _, m0, m1 = s.render(70)
    _, m2, _ = s.render(72)
    sl = (slice(100, 356), slice(180, 564))
    A, Bt, Bs = (np.ascontiguousarray(x[sl]) for x in (m0, m2, m1))
    p0 = cv2.goodFeaturesToTrack(A, 150, 0.01, 10).reshape(-1, 2).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 100, 1e-4)
    out = {"a": A, "b_temporal": Bt, "b_stereo": Bs, "p0": p0}
    for name, B, shift in (("temporal", Bt, (0.0, 0.0)), ("stereo", Bs, (-20.0, 0.0))):
        g0 = (p0 + np.array(shift, np.float32)).reshape(-1, 1, 2)
        p1, st, _ = cv2.calcOpticalFlowPyrLK(A, B, p0.reshape(-1, 1, 2), g0.copy(), winSize=(15, 15), maxLevel=3, criteria=crit,
                                             flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        out[f"p1_{name}"] = p1.reshape(-1, 2)
        out[f"st_{name}"] = st.ravel()
        out[f"guess_{name}"] = g0.reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "cv2_klt_motion.npz"), **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
