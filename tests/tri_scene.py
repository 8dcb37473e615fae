"""Shared by the CPU and GPU triangulation tests: a seeded multi-view stereo scene in the
conventions of Feature::initializePosition (feature.hpp:289-450): camera-state orientation is the
JPL quaternion [x y z w] of R_w_c (world -> cam0), position is the cam0 centre in the world,
T_cn_cnm1 takes cam0 coordinates to cam1 coordinates."""
import numpy as np


def _rot_to_quat_jpl(R):
    """JPL [x y z w] with R = (2w^2-1) I - 2w [q]x + 2 q q^T (oracle/kin.h quat_to_rot)."""
    tr = np.trace(R)
    w = np.sqrt(max(1 + tr, 1e-12)) / 2
    return np.array([(R[1, 2] - R[2, 1]) / (4 * w), (R[2, 0] - R[0, 2]) / (4 * w), (R[0, 1] - R[1, 0]) / (4 * w), w])


def _rodrigues(v):
    th = np.linalg.norm(v)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    K = K / th
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def make_scene(cfg, n_cam, n_feat, seed, noise=2e-3, outlier_frac=0.1):
    rng = np.random.default_rng(seed)
    T01 = np.array(cfg.T_cn_cnm1[:]).reshape(4, 4)
    cam_q, cam_p, R_wc = [], [], []
    p = np.zeros(3)
    for c in range(n_cam):
        p = p + rng.normal(0, 0.12, 3) + np.array([0.15, 0.0, 0.02])
        R = _rodrigues(rng.normal(0, 0.08, 3))  # R_w_c
        q = _rot_to_quat_jpl(R)
        cam_q.append(q / np.linalg.norm(q))
        cam_p.append(p.copy())
        R_wc.append(R)
    pts = np.column_stack([rng.uniform(-3, 5, n_feat), rng.uniform(-2, 2, n_feat), rng.uniform(2.5, 9, n_feat)])
    obs = np.zeros((n_feat, n_cam, 4))
    mask = np.zeros(n_feat, np.uint32)
    for f in range(n_feat):
        M = int(rng.integers(1, n_cam + 1)) if f % 7 else n_cam
        first = int(rng.integers(0, n_cam - M + 1))
        for c in range(first, first + M):
            if M > 4 and rng.random() < 0.1:
                continue  # gaps: a feature need not be seen by consecutive camera states
            pc0 = R_wc[c] @ (pts[f] - cam_p[c])
            pc1 = T01[:3, :3] @ pc0 + T01[:3, 3]
            z = np.array([pc0[0] / pc0[2], pc0[1] / pc0[2], pc1[0] / pc1[2], pc1[1] / pc1[2]])
            z += rng.normal(0, noise, 4)
            if rng.random() < outlier_frac:
                z += rng.normal(0, 0.05, 4)  # beyond the Huber radius (0.01)
            obs[f, c] = z
            mask[f] |= np.uint32(1 << c)
        if mask[f] == 0:
            mask[f] = np.uint32(1 << first)
            obs[f, first] = [0.01, 0.02, -0.01, 0.02]
    return np.array(cam_q), np.array(cam_p), mask, obs, pts
