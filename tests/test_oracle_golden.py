"""CPU: the oracle's SPEC primitives against the committed OpenCV golden vectors
(tests/golden/make_golden.py explains why OpenCV is the pin: SURVEY F1/F2 + section 8c)."""
import os

import numpy as np
import pytest

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


# ---- KLT under real motion: what separates the SPEC from OpenCV, with numbers (SPEC.md section 3) -------------
def _bil(I, x, y):
    x0, y0 = np.floor(x).astype(int), np.floor(y).astype(int)
    ax, ay = x - x0, y - y0
    return (1 - ax) * (1 - ay) * I[y0, x0] + ax * (1 - ay) * I[y0, x0 + 1] + (1 - ax) * ay * I[y0 + 1, x0] + ax * ay * I[y0 + 1, x0 + 1]


def _float_lk(A, B, p, q, win, gradient):
    """Independent float64 Lucas-Kanade on level 0, started at q, iterated to its fixed point.
    gradient "template": central differences of the interpolated template (the SPEC, unnormalised: x2);
    gradient "scharr": OpenCV's 3-10-3 Scharr derivative of the image, bilinearly interpolated."""
    h = win // 2
    ii, jj = np.meshgrid(np.arange(-h, h + 1), np.arange(-h, h + 1))
    x, y = p[0] + ii, p[1] + jj
    T = _bil(A, x, y)
    if gradient == "template":
        Ix, Iy, step = _bil(A, x + 1, y) - _bil(A, x - 1, y), _bil(A, x, y + 1) - _bil(A, x, y - 1), 2.0
    else:
        P = np.pad(A, 1, mode="reflect")
        sm_v = 3 * P[:-2, :] + 10 * P[1:-1, :] + 3 * P[2:, :]   # vertical smoothing, full width
        sm_h = 3 * P[:, :-2] + 10 * P[:, 1:-1] + 3 * P[:, 2:]   # horizontal smoothing, full height
        gx = (sm_v[:, 2:] - sm_v[:, :-2]) / 32.0
        gy = (sm_h[2:, :] - sm_h[:-2, :]) / 32.0
        Ix, Iy, step = _bil(gx, x, y), _bil(gy, x, y), 1.0
    G = np.array([[np.sum(Ix * Ix), np.sum(Ix * Iy)], [np.sum(Ix * Iy), np.sum(Iy * Iy)]])
    q = np.array(q, float)
    for _ in range(300):
        d = _bil(B, q[0] + ii, q[1] + jj) - T
        delta = -step * np.linalg.solve(G, np.array([np.sum(d * Ix), np.sum(d * Iy)]))
        q += delta
        if delta @ delta < 1e-14:
            break
    return q


@pytest.mark.parametrize("pair", ["temporal", "stereo"])
def test_klt_under_motion_vs_cv2_and_float_reference(ob, synth, golden_dir, pair):
    """Flow of ~5 px (temporal) / ~13 px from a -20 px guess (stereo), tight termination on every side so that
    fixed points are compared, interior well-textured points.  (1) The SPEC's fixed-point LK is within 1e-3 px
    of a float64 LK that uses the same gradient: fixed point is not the limitation.  (2) It is ~1e-2 px from
    OpenCV, and (3) a float64 LK with OpenCV's Scharr gradient is within 1e-3 px of OpenCV: the gradient
    operator is the whole difference."""
    g = np.load(os.path.join(golden_dir, "cv2_klt_motion.npz"))
    a, b = g["a"], g[f"b_{pair}"]
    cfg = copy_cfg(synth.default_config("ref"), img_rows=a.shape[0], img_cols=a.shape[1], klt_eps=1e-4, klt_max_iters=100)
    p0, p_cv, st_cv = g["p0"], g[f"p1_{pair}"], g[f"st_{pair}"]
    p_or, st_or = ob.klt(cfg, a, b, p0, g[f"guess_{pair}"])
    af, bf = a.astype(float), b.astype(float)
    H, W = a.shape

    def interior(p, m=30):
        return m <= p[0] <= W - 1 - m and m <= p[1] <= H - 1 - m

    # status: OpenCV keeps tracking while part of the window is inside the (padded) image, the SPEC drops a
    # point as soon as it leaves [0, cols - 1] x [0, rows - 1]; away from the border they agree
    inner = np.array([interior(p) and interior(gq) and interior(qc) for p, gq, qc in zip(p0, g[f"guess_{pair}"], p_cv)])
    assert inner.sum() >= 60 and (st_or == st_cv)[inner].mean() >= 0.97

    d_cv, d_float, d_scharr = [], [], []
    for p, q, qc, ok, okc in zip(p0, p_or, p_cv, st_or, st_cv):
        if not (ok and okc and interior(p) and interior(q) and interior(qc)) or np.abs(q - qc).max() > 0.5:
            continue  # (a gross mismatch is a different local minimum, not arithmetic)
        d_cv.append(np.abs(q - qc).max())
        d_float.append(np.abs(_float_lk(af, bf, p.astype(float), q, 15, "template") - q).max())
        d_scharr.append(np.abs(_float_lk(af, bf, p.astype(float), qc, 15, "scharr") - qc).max())
    d_cv, d_float, d_scharr = map(np.array, (d_cv, d_float, d_scharr))
    assert len(d_cv) >= 60
    # (1) fixed point vs float64, same algorithm: measured median 1.1e-4, max 1.1e-3
    assert np.median(d_float) <= 3e-4 and np.quantile(d_float, 0.97) <= 1e-3 and d_float.max() <= 2e-3
    # (2) SPEC vs OpenCV: measured median 1.0e-2, max 6e-2
    assert 2e-3 <= np.median(d_cv) <= 2e-2 and d_cv.max() <= 0.15
    # (3) float64 + Scharr vs OpenCV: measured median 1.9e-4, max 9e-4
    assert np.median(d_scharr) <= 5e-4 and np.quantile(d_scharr, 0.97) <= 2e-3
