"""GPU: the CUDA front end (through the C ABI) against the CPU oracle on identical seeded
inputs.  Bars (BASELINE.json north_star): pyramids and FAST keypoint sets bit-exact; KLT
within 1e-3 px with >= 99.5 % status agreement (the fixed-point SPEC makes it bit-exact in
practice, which is what is asserted)."""
import numpy as np
import pytest

from conftest import copy_cfg

pytestmark = pytest.mark.gpu

KLT_TOL_PX = 1e-3
KLT_STATUS_AGREEMENT = 0.995


@pytest.fixture(scope="module")
def eng():
    from msckf_stereo_c_b200 import engine

    return engine


def _frames(synth, cfg, seed, ks):
    s = synth.Stream(cfg, seed=seed)
    return [s.render(k) for k in ks]


# (200, 360): pitch not a multiple of 16 -> the one-launch pyramid tail stages its chunks with word loads instead of
# bulk copies, and level 3 is 45 wide (odd); (480, 752) and (1024, 1280) take the bulk-staged tail from level 2 / 3
@pytest.mark.parametrize("shape,levels", [((480, 752), 4), ((123, 157), 4), ((1024, 1280), 6), ((65, 33), 3), ((200, 360), 4)])
def test_pyramid_bit_exact(eng, ob, synth, shape, levels):
    rng = np.random.default_rng(shape[0])
    imgs = rng.integers(0, 256, (3,) + shape, dtype=np.uint8)
    cfg = synth.default_config("ref")
    e = eng.Engine(cfg, 1)
    got = e.op_pyramid(imgs, levels)
    for i in range(3):
        ref = imgs[i]
        for l in range(levels - 1):
            ref = ob.pyr_down(ref)
            assert np.array_equal(ref, got[i][l]), (i, l)
    e.close()


def test_pyramid_golden_cv2(eng, golden_dir, synth):
    import os

    g = np.load(os.path.join(golden_dir, "cv2_pyr_fast.npz"))
    e = eng.Engine(synth.default_config("ref"), 1)
    got = e.op_pyramid(g["img"][None], 4)[0]
    for l, k in enumerate(("l1", "l2", "l3")):
        assert np.array_equal(got[l], g[k])
    e.close()


@pytest.mark.parametrize("preset,seed", [("ref", 0), ("ref", 11), ("stress", 2)])
def test_detect_bit_exact(eng, ob, synth, preset, seed):
    cfg = synth.default_config(preset)
    (_, a, _), = _frames(synth, cfg, seed, [35])
    e = eng.Engine(cfg, 1)
    xy_o, r_o, sm_o = ob.detect(cfg, a, want_scores=True)
    xy_g, r_g, sm_g = e.debug_detect_scores(a)
    assert np.array_equal(sm_o, sm_g)  # per-pixel FAST scores
    assert len(xy_o) > 100
    assert np.array_equal(xy_o, xy_g) and np.array_equal(r_o, r_g)  # keypoint set, order and responses
    occ = xy_o[::3]
    xy_o2, r_o2 = ob.detect(cfg, a, occupied=occ)
    xy_g2, r_g2 = e.op_detect(a, occupied=occ)
    assert np.array_equal(xy_o2, xy_g2) and np.array_equal(r_o2, r_g2)
    flat = np.full_like(a, 128)
    assert len(e.op_detect(flat)[0]) == 0
    e.close()


def test_detect_bit_exact_realistic_corner_density(eng, ob, synth):
    """The synthetic texture is corner-dense (26 % of the pixels are FAST corners); real imagery has 2-5 %.  A
    low-pass filtered frame puts the detector in that regime (short candidate lists, mostly-empty tiles): the
    score map, the keypoint set, its order and the responses must still be bit-exact."""
    from scipy.ndimage import gaussian_filter

    cfg = synth.default_config("bench")
    (_, a, _), = _frames(synth, cfg, 3, [40])
    a = np.clip(np.rint(gaussian_filter(a.astype(np.float32), 2.0)), 0, 255).astype(np.uint8)
    e = eng.Engine(cfg, 1)
    xy_o, r_o, sm_o = ob.detect(cfg, a, want_scores=True)
    density = (sm_o > 0).mean()
    assert 0.01 <= density <= 0.06, density
    xy_g, r_g, sm_g = e.debug_detect_scores(a)
    assert np.array_equal(sm_o, sm_g)
    assert len(xy_o) > 50
    assert np.array_equal(xy_o, xy_g) and np.array_equal(r_o, r_g)
    e.close()


@pytest.mark.parametrize("preset", ["ref", "bench"])
def test_klt_parity(eng, ob, synth, preset):
    cfg = synth.default_config(preset)
    (_, a, b), (_, a2, _) = _frames(synth, cfg, 4, [40, 41])
    xy, _ = ob.detect(cfg, a)
    e = eng.Engine(cfg, 1)
    for img_b, guess in ((a2, xy), (b, xy + np.float32([-20, 0])), (a2, xy + np.float32([3.3, -2.7]))):
        pb_o, st_o = ob.klt(cfg, a, img_b, xy, guess)
        pb_g, st_g = e.op_klt(a, img_b, xy, guess)
        assert (st_o == st_g).mean() >= KLT_STATUS_AGREEMENT
        both = (st_o > 0) & (st_g > 0)
        assert both.sum() > 50
        assert np.abs(pb_o - pb_g)[both].max() <= KLT_TOL_PX
        assert np.array_equal(pb_o[both], pb_g[both]) and np.array_equal(st_o, st_g)  # fixed point: bit-exact
    # failures: flat template, guess outside the image
    flat = np.full_like(a, 90)
    _, st = e.op_klt(flat, flat, xy[:8], xy[:8])
    assert not st.any()
    _, st = e.op_klt(a, a, xy[:1], np.float32([[9000, 10]]))
    assert st[0] == 0
    e.close()


def _run_pipeline(eng, ob, synth, cfg, seed, n_frames):
    s = synth.Stream(cfg, seed=seed)
    e = eng.Engine(cfg, 1)
    o = ob.Oracle(cfg)

    class Both:
        def imu(self, t, w, a):
            o.imu(t, w, a)
            e.imu_callback(t, w, a)

        def stereo(self, t, i0, i1):
            o.stereo(t, i0, i1)
            e.stereo_callback(t, i0, i1)

        def backend(self):
            pass

    for k, t in synth.feed(s, n_frames, Both()):
        go, gg = o.grid(), e.grid()
        assert len(go) == len(gg), k
        for f in ("id", "lifetime", "cam0", "cam1", "cell", "response"):
            assert np.array_equal(go[f], gg[f]), (k, f)
        (to, fo, no), (tg, fg, ng) = o.features(), e.features()
        assert to == tg and no == ng and fo.tobytes() == fg.tobytes(), k  # CameraMeasurement incl. the stale tail (F4)
        io, ig = o.tracking_info(), e.tracking_info()
        assert (io.before_tracking, io.after_tracking, io.after_matching, io.after_ransac) == \
               (ig.before_tracking, ig.after_tracking, ig.after_matching, ig.after_ransac), k
        for cam in (0, 1):
            for l in range(cfg.pyramid_levels):
                assert np.array_equal(o.pyramid(cam, l), e.pyramid(cam, l)), (k, cam, l)
    e.close()


def test_frontend_pipeline_ref(eng, ob, synth):
    """stereoCallback frame by frame (image_processor.cpp:139-203), preset `ref`, through the
    static start and the onset of motion."""
    _run_pipeline(eng, ob, synth, synth.default_config("ref"), 0, 45)


def test_frontend_pipeline_bench_and_fixed_modes(eng, ob, synth):
    _run_pipeline(eng, ob, synth, synth.default_config("bench"), 1, 14)
    cfg = copy_cfg(synth.default_config("ref"), compat_stale_features=0, fix_prev_image_alias=1)
    _run_pipeline(eng, ob, synth, cfg, 2, 40)


def test_batched_streams_equal_single_stream(eng, synth):
    """Streams are independent: a 3-stream batch must reproduce each 1-stream run bit for bit
    (this is also the multi-GPU correctness argument: sharding = choosing which streams)."""
    cfg = synth.default_config("ref")
    seeds = [0, 7, 9]
    streams = [synth.Stream(cfg, seed=s) for s in seeds]
    nf = 34
    frames = [[st.render(k) for k in range(nf)] for st in streams]
    singles = []
    for i in range(3):
        e = eng.Engine(cfg, 1)
        out = []
        for k in range(nf):
            t, a, b = frames[i][k]
            e.stereo_callback(t, a, b)
            out.append((e.grid().tobytes(), e.features()[1].tobytes()))
        singles.append(out)
        e.close()
    e = eng.Engine(cfg, 3)
    for k in range(nf):
        for i in range(3):
            t, a, b = frames[i][k]
            e.push_stereo(t, a, b, stream=i)
        e.frontend_step()
        for i in range(3):
            assert (e.grid(i).tobytes(), e.features(i)[1].tobytes()) == singles[i][k], (k, i)
    assert e.launch_count() > 0
    e.close()


def test_bad_arguments(eng, synth):
    cfg = synth.default_config("ref")
    e = eng.Engine(cfg, 1)
    with pytest.raises(eng.EngineError):
        e.push_stereo(0.0, np.zeros((10, 10), np.uint8), np.zeros((10, 10), np.uint8))
    with pytest.raises(eng.EngineError):
        e.imu_callback(0.0, [0, 0, 0], [0, 0, 0], stream=5)
    e.close()
    with pytest.raises(eng.EngineError):
        eng.Engine(copy_cfg(cfg, klt_win=16), 1)


def test_frontend_pipeline_stress_preset(eng, ob, synth):
    """BASELINE.json config 5 geometry: 1280x1024, 6 pyramid levels, 8x10 grid (~1000 features)."""
    _run_pipeline(eng, ob, synth, synth.default_config("stress"), 3, 6)


@pytest.mark.parametrize("fix_alias", [0, 1])
def test_frontend_pipeline_with_two_point_ransac(eng, ob, synth, fix_alias):
    """twoPointRansac (image_processor.cpp:911-1135) is dead code in the reference (:482-493); with
    use_ransac the engine and the oracle run it as the commented-out calls read, with a shared
    counter-based sampler in place of the unseeded cg::uniform_integer: identical grids and counters,
    through the static start (degenerate-motion branch) and the moving part (RANSAC branch)."""
    cfg = copy_cfg(synth.default_config("ref"), use_ransac=1, fix_prev_image_alias=fix_alias)
    _run_pipeline(eng, ob, synth, cfg, 0, 64)
    # the check must have had something to reject: the oracle drops features at the RANSAC stage
    s = synth.Stream(cfg, seed=0)
    o = ob.Oracle(cfg)
    dropped = 0
    for k in range(64):
        t, a, b = s.render(k)
        o.stereo(t, a, b)
        ti = o.tracking_info()
        dropped += ti.after_matching - ti.after_ransac
    assert dropped > 0


def _equidistant(cfg):
    """Both cameras on the equidistant (fisheye) model, image_processor.cpp:811-813, 838-841."""
    c = copy_cfg(cfg, cam0_model=1, cam1_model=1)
    for i, v in enumerate([-0.0135, 0.021, -0.03, 0.012]):
        c.cam0_distortion[i] = v
    for i, v in enumerate([-0.0121, 0.018, -0.027, 0.011]):
        c.cam1_distortion[i] = v
    return c


@pytest.mark.gpu
def test_frontend_pipeline_equidistant_model(eng, ob, synth):
    """The whole front end on equidistant cameras (rendered with the same model): stereo guess through
    undistort_points_fisheye / distort_points_fisheye, epipolar gate, published normalized coordinates;
    also with the two-point RANSAC on, which undistorts both frames' points."""
    _run_pipeline(eng, ob, synth, _equidistant(synth.default_config("ref")), 3, 40)
    _run_pipeline(eng, ob, synth, copy_cfg(_equidistant(synth.default_config("ref")), use_ransac=1), 4, 30)
