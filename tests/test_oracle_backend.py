"""CPU: the oracle's fp64 linear algebra against numpy, the null-space / QR invariance
property the EKF parity argument rests on, and a filter sanity run."""
import numpy as np

from conftest import copy_cfg


def _spd(rng, n, scale=1e-2):
    A = rng.standard_normal((n, n))
    return scale * (A @ A.T / n + 0.1 * np.eye(n))


def test_householder_qr_vs_numpy(ob):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((60, 17))
    b = rng.standard_normal(60)
    R, qtb = ob.qr_thin(A, b)
    Qn, Rn = np.linalg.qr(A)
    sgn = np.sign(np.diag(R)) * np.sign(np.diag(Rn))
    assert np.abs(R - sgn[:, None] * Rn).max() < 1e-12
    assert np.abs(qtb - sgn * (Qn.T @ b)).max() < 1e-12
    assert np.abs(R.T @ R - A.T @ A).max() < 1e-11


def test_ldlt_solve_vs_numpy(ob):
    rng = np.random.default_rng(2)
    S = _spd(rng, 40, 1.0)
    B = rng.standard_normal((40, 7))
    X = ob.ldlt_solve(S, B)
    assert np.abs(X - np.linalg.solve(S, B)).max() < 1e-11


def test_update_math_vs_numpy(ob):
    rng = np.random.default_rng(3)
    n = 33
    P = _spd(rng, n)
    for m in (10, 90):  # m <= n: H used as is; m > n: QR-compressed (msckf_vio.cpp:795-810)
        H = rng.standard_normal((m, n))
        H[:, :21] = 0.0  # featureJacobian never touches the IMU columns (msckf_vio.cpp:709-712)
        r = rng.standard_normal(m) * 1e-2
        dx, Pn = ob.update_math(H, r, P, 0.035 ** 2)
        S = H @ P @ H.T + 0.035 ** 2 * np.eye(m)
        K = np.linalg.solve(S, H @ P).T
        Pref = (np.eye(n) - K @ H) @ P
        Pref = 0.5 * (Pref + Pref.T)
        assert np.abs(dx - K @ r).max() < 1e-12 * max(1.0, np.abs(K @ r).max())
        assert np.abs(Pn - Pref).max() / np.abs(Pref).max() < 1e-11


def test_posterior_independent_of_nullspace_basis(ob):
    """The reference projects with the last 4M-3 columns of U from svd_fulluv
    (msckf_vio.cpp:757-766); the oracle and the CUDA engine use Householder reflectors.
    gamma, delta_x and the posterior covariance must not depend on that choice."""
    rng = np.random.default_rng(4)
    M, N = 6, 8
    n = 21 + 6 * N
    rows = 4 * M
    Hx = np.zeros((rows, n))
    for i in range(M):
        Hx[4 * i:4 * i + 4, 21 + 6 * i:27 + 6 * i] = rng.standard_normal((4, 6))
    Hf = rng.standard_normal((rows, 3))
    r = rng.standard_normal(rows) * 1e-2
    P = _spd(rng, n)
    U, _, _ = np.linalg.svd(Hf, full_matrices=True)
    A = U[:, 3:]
    dx_q, P_q, g_q = ob.nullspace_update(Hx, Hf, r, P, 0.035 ** 2)
    dx_s, P_s, g_s = ob.nullspace_update(Hx, Hf, r, P, 0.035 ** 2, basis=A)
    assert abs(g_q - g_s) <= 1e-12 * abs(g_s)
    assert np.abs(dx_q - dx_s).max() <= 1e-12 * np.abs(dx_s).max()
    assert np.abs(P_q - P_s).max() <= 1e-12 * np.abs(P_s).max()


def test_chi2_table(ob, synth):
    from scipy.stats import chi2

    cfg = synth.default_config("ref")
    for mode, q in ((0, 0.05), (1, 0.95)):
        c = copy_cfg(cfg, chi2_mode=mode)
        for dof in (1, 2, 10, 57, 99):
            assert abs(ob.chi2(c, dof) - chi2.ppf(q, dof)) < 1e-9 * chi2.ppf(q, dof) + 1e-12
    assert ob.chi2(cfg, 0) == 0.0 and ob.chi2(cfg, 100) == 0.0  # std::map operator[] off the table


def test_filter_tracks_ground_truth(ob, synth):
    """Full oracle pipeline on the seeded synthetic mav0-style stream: gravity init after 200
    static IMU rows (msckf_vio.cpp:198), covariance symmetric PSD, bounded drift."""
    cfg = synth.default_config("ref")
    s = synth.Stream(cfg, seed=0)
    o = ob.Oracle(cfg)
    est, gt = [], []
    p0 = s.pose(s.frame_time(0))[1]

    class Sink:
        def imu(self, t, w, a):
            o.imu(t, w, a)

        def stereo(self, t, a, b):
            o.stereo(t, a, b)

        def backend(self):
            o.backend()

    for k, t in synth.feed(s, 70, Sink()):
        st = o.state()
        assert st.is_gravity_set == (1 if k >= 20 else 0) or k in (19, 20)
        if st.n_cam_states:
            est.append(np.array(st.position[:]))
            gt.append(s.pose(t)[1] - p0)
    st = o.state()
    assert st.n_cam_states <= cfg.max_cam_state_size and st.cov_dim == 21 + 6 * st.n_cam_states
    assert st.n_updates > 5 and st.n_resets == 0
    P = o.cov()
    assert np.abs(P - P.T).max() == 0.0
    assert np.linalg.eigvalsh(P).min() > -1e-12
    est, gt = np.array(est), np.array(gt)
    assert abs(np.linalg.norm(est[-1]) - np.linalg.norm(gt[-1])) < 0.1
    assert abs(np.linalg.norm(np.array(st.gravity[:])) - 9.81) < 0.05
