"""CPU: the oracle's SPEC primitives against the committed OpenCV golden vectors
(tests/golden/make_golden.py explains why OpenCV is the pin: SURVEY F1/F2 + section 8c)."""
import os

import numpy as np

from conftest import copy_cfg


def test_pyr_down_matches_cv2_bit_exact(ob, golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_pyr_fast.npz"))
    img = g["img"]
    for k in ("l1", "l2", "l3"):
        img = ob.pyr_down(img)
        assert img.shape == g[k].shape
        assert np.array_equal(img, g[k]), k


def test_pyr_down_degenerate_sizes(ob):
    rng = np.random.default_rng(0)
    for shape in ((1, 1), (1, 7), (2, 2), (3, 5), (5, 1)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        out = ob.pyr_down(img)
        assert out.shape == ((shape[0] + 1) // 2, (shape[1] + 1) // 2)
    flat = np.full((9, 11), 200, np.uint8)
    assert np.all(ob.pyr_down(flat) == 200)  # unit DC gain, rounding included


def test_fast_scores_match_cv2(ob, synth, golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_pyr_fast.npz"))
    img = g["img"]
    cfg = copy_cfg(synth.default_config("ref"), img_rows=img.shape[0], img_cols=img.shape[1], det_rows=8, det_cols=10)
    _, _, sm = ob.detect(cfg, img, want_scores=True)
    got = np.zeros(img.shape, bool)
    got[g["fast_all_xy"][:, 1], g["fast_all_xy"][:, 0]] = True
    assert np.array_equal(sm > 0, got)  # FAST-9/16 segment test, threshold 10
    for x, y, r in g["fast_nms_xyr"]:  # cv2 response == score where cv2 keeps the corner
        assert sm[y, x] == r
    # strict 3x3 NMS of the score map == cv2's nonmaxSuppression set
    s = np.pad(sm.astype(int), 1)
    H, W = sm.shape
    ismax = sm > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                ismax &= sm.astype(int) > s[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    want = np.zeros(img.shape, bool)
    want[g["fast_nms_xyr"][:, 1], g["fast_nms_xyr"][:, 0]] = True
    assert np.array_equal(ismax, want)


def test_detector_grid_and_occupancy(ob, synth):
    cfg = synth.default_config("ref")
    s = synth.Stream(cfg, seed=5)
    _, img, _ = s.render(30)
    xy, resp = ob.detect(cfg, img)
    ch, cw = cfg.img_rows // cfg.det_rows + 1, cfg.img_cols // cfg.det_cols + 1
    cells = (xy[:, 1].astype(int) // ch) * cfg.det_cols + xy[:, 0].astype(int) // cw
    assert len(np.unique(cells)) == len(cells)  # one corner per fine cell
    assert np.all(resp > cfg.detection_threshold)
    assert np.all(np.diff(cells) > 0)  # cell-major output order
    # occupied cells produce nothing (CornerDetector::set_grid_position, image_processor.cpp:647)
    xy2, _ = ob.detect(cfg, img, occupied=xy[:50])
    cells2 = (xy2[:, 1].astype(int) // ch) * cfg.det_cols + xy2[:, 0].astype(int) // cw
    assert not set(cells[:50]) & set(cells2)
    assert set(cells[50:]) == set(cells2)
    # a flat image has no corners
    assert len(ob.detect(cfg, np.full_like(img, 77))[0]) == 0


def test_point_maps_match_cv2(ob, golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_points.npz"))
    pts = g["pts"]
    for cam in (0, 1):
        K, D = g[f"K{cam}"], g[f"D{cam}"]
        und = ob.undistort(pts, K, 0, D)
        assert np.abs(und - g[f"und{cam}"]).max() <= 1e-6
        assert np.abs(ob.undistort(pts, K, 0, D, R=g[f"R{cam}"]) - g[f"undR{cam}"]).max() <= 1e-6
        assert np.abs(ob.distort(g[f"und{cam}"], K, 0, D) - g[f"dist{cam}"]).max() <= 1e-3  # pixels (float32 in/out)
        assert np.abs(ob.undistort(pts, K, 1, g[f"Df{cam}"]) - g[f"fund{cam}"]).max() <= 1e-6
        assert np.abs(ob.distort(g[f"fund{cam}"].astype(np.float32), K, 1, g[f"Df{cam}"]) - g[f"fdist{cam}"]).max() <= 1e-3
    for v, R in zip(g["rodrigues_v"], g["rodrigues_R"]):
        assert np.abs(ob.rodrigues(v) - R).max() <= 1e-12


def test_klt_close_to_cv2(ob, synth, golden_dir):
    """The SPEC's LK is fixed point (so CPU and GPU agree bit for bit); OpenCV's is float.
    They must agree to well under a pixel on every tracked point."""
    g = np.load(os.path.join(golden_dir, "cv2_klt.npz"))
    a, b = g["a"], g["b"]
    cfg = copy_cfg(synth.default_config("ref"), img_rows=a.shape[0], img_cols=a.shape[1])
    pb, st = ob.klt(cfg, a, b, g["p0"], g["p0"])
    both = (st > 0) & (g["st"] > 0)
    assert (st == g["st"]).mean() >= 0.97
    d = np.abs(pb - g["p1"])[both]
    assert np.median(d) < 5e-3 and d.max() < 0.1


def test_klt_identity_and_failures(ob, synth):
    cfg = synth.default_config("ref")
    s = synth.Stream(cfg, seed=1)
    _, a, _ = s.render(30)
    xy, _ = ob.detect(cfg, a)
    pb, st = ob.klt(cfg, a, a, xy[:100], xy[:100])
    assert st.all() and np.abs(pb - xy[:100]).max() < 1e-2  # same image: stays put
    flat = np.full_like(a, 90)
    _, st = ob.klt(cfg, flat, flat, xy[:10], xy[:10])
    assert not st.any()  # no texture -> min-eigenvalue test fails
    far = np.array([[5000.0, 100.0]], np.float32)
    _, st = ob.klt(cfg, a, a, xy[:1], far)
    assert st[0] == 0  # initial guess outside the image
    pb, st = ob.klt(cfg, a, a, xy[:0], xy[:0])
    assert len(pb) == 0 and len(st) == 0
