"""Pins for the BA restatements (g2o / Ceres are third-party code outside the reference tree, so no reference-run output
exists): (1) the reference's own synthetic fixture test/test_local_ba.cpp:9-98 with EXACTLY its cv::RNG noise
(tests/golden/local_ba_cvrng.npz, tools/make_local_ba_fixture.py); (2) the optimum of both BA flavours cross-checked with
scipy.optimize.least_squares on an independent numpy statement of the two cost functions (written from
include/ygz/G2oTypes.h:84-91 and include/ygz/Ceres/CeresReprojectionError.h:33-69, sharing no code with oracle/ba.cpp)."""
from pathlib import Path

import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from ygz_slam_b200 import se3

G = np.load(Path(__file__).parent / "golden" / "local_ba_cvrng.npz")
FX, FY, CX, CY = (float(np.float32(v)) for v in (520.9, 521.0, 325.1, 249.7))   # PinholeCamera keeps floats (Camera.h:14-22)


def _g2o(v):  # [upsilon; omega] -> vertex order [omega; upsilon]
    v = np.asarray(v)
    return np.concatenate([v[..., 3:], v[..., :3]], -1)


def test_fixture_is_the_reference_scene():
    """Shape and statistics of the cv::RNG fixture: 8 key-frames x 16 points, every point seen by every key-frame."""
    assert G["px"].shape == (128, 2) and G["logs_noisy"].shape == (8, 6) and G["pts_noisy"].shape == (16, 3)
    assert np.array_equal(G["logs_noisy"][0], G["logs_true"][0])          # key-frame 0 gets no noise (test_local_ba.cpp:59)
    assert abs(np.std(G["unit_normals"]) - 1.0) < 0.1
    # the pixels are the projections of the TRUE scene plus the unit-normal draws, u first
    T = [se3.se3_exp(v) for v in G["logs_true"]]
    k, j = int(G["kf_idx"][37]), int(G["pt_idx"][37])
    pc = T[k][:, :3] @ G["pts_true"][j] + T[k][:, 3]
    clean = np.array([FX * pc[0] / pc[2] + CX, FY * pc[1] / pc[2] + CY])
    assert np.abs(G["px"][37] - clean).max() < 5.0


def test_g2o_flavour_on_the_reference_fixture(oracle):
    fixed = np.zeros(8, np.uint8)
    fixed[0] = 1
    P, X, outl, st = oracle.local_ba(_g2o(G["logs_noisy"]), fixed, G["pts_noisy"], G["kf_idx"], G["pt_idx"], G["px"])
    assert st["iters"] >= 5 and st["chi2_final"] < 1e-2 * st["chi2_initial"]
    assert st["chi2_final"] < 2.5 * len(G["px"])                          # pixel noise sigma = 1: ~2 per edge at the optimum
    assert np.abs(P[:, :3] - G["logs_true"][:, 3:]).max() < 0.03          # rotations come back to the truth (scale is gauge)
    assert outl.sum() <= 0.1 * len(G["px"])


def _unpack(x, pose0, n_kf, n_pt):
    poses = np.vstack([pose0[None], x[:6 * (n_kf - 1)].reshape(n_kf - 1, 6)])
    return poses, x[6 * (n_kf - 1):].reshape(n_pt, 3)


def _scipy_g2o(poses0, pts0, kf, pt, px):
    """min sum |px - (f X/Z + c)|^2 over poses [upsilon; omega] of T_cw (key-frame 0 fixed) and points; any smooth pose
    chart has the same minimiser, so the se3 exponential of ygz_slam_b200.se3 (plain numpy) is used."""
    n_kf, n_pt = len(poses0), len(pts0)

    def res(x):
        poses, pts = _unpack(x, poses0[0], n_kf, n_pt)
        T = np.stack([se3.se3_exp(v) for v in poses])
        pc = np.einsum("nij,nj->ni", T[kf][:, :, :3], pts[pt]) + T[kf][:, :, 3]
        return np.concatenate([px[:, 0] - (FX * pc[:, 0] / pc[:, 2] + CX), px[:, 1] - (FY * pc[:, 1] / pc[:, 2] + CY)])

    x0 = np.concatenate([poses0[1:].ravel(), pts0.ravel()])
    r = least_squares(res, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    return _unpack(r.x, poses0[0], n_kf, n_pt) + (float(r.fun @ r.fun),)


def _scipy_ceres(poses0, pts0, kf, pt, px):
    """min sum |(px - c)/f - p/p.z|^2, p = R(angle-axis) X + t, poses [t; angle-axis] (CeresReprojectionError.h:33-69)."""
    n_kf, n_pt = len(poses0), len(pts0)
    obs = np.stack([(px[:, 0] - CX) / FX, (px[:, 1] - CY) / FY], 1)

    def res(x):
        poses, pts = _unpack(x, poses0[0], n_kf, n_pt)
        R = Rotation.from_rotvec(poses[:, 3:]).as_matrix()
        p = np.einsum("nij,nj->ni", R[kf], pts[pt]) + poses[kf, :3]
        return np.concatenate([obs[:, 0] - p[:, 0] / p[:, 2], obs[:, 1] - p[:, 1] / p[:, 2]])

    x0 = np.concatenate([poses0[1:].ravel(), pts0.ravel()])
    r = least_squares(res, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    return _unpack(r.x, poses0[0], n_kf, n_pt) + (float(r.fun @ r.fun),)


def test_g2o_flavour_optimum_matches_scipy(oracle):
    """No robust kernel (Huber changes the cost): the Levenberg restatement and scipy's MINPACK reach the same minimum."""
    kf, pt, px = G["kf_idx"], G["pt_idx"], G["px"]
    fixed = np.zeros(8, np.uint8)
    fixed[0] = 1
    P, X, _, st = oracle.local_ba(_g2o(G["logs_noisy"]), fixed, G["pts_noisy"], kf, pt, px, max_iters=200, huber=0.0)
    Ps, Xs, cost = _scipy_g2o(G["logs_noisy"], G["pts_noisy"], kf, pt, px)
    assert abs(st["chi2_final"] - cost) < 1e-6 * cost
    # monocular BA with one fixed key-frame leaves the global scale free (a flat direction of the cost): rotations are
    # compared directly, translations and points after aligning the one scale factor
    sc = float((X * Xs).sum() / (Xs * Xs).sum())
    for k in range(8):
        To, Ts = se3.se3_exp(np.r_[P[k, 3:], P[k, :3]]), se3.se3_exp(Ps[k])
        assert np.abs(To[:, :3] - Ts[:, :3]).max() < 1e-5, k
        assert np.abs(To[:, 3] - sc * Ts[:, 3]).max() < 1e-4, k
    assert np.abs(X - sc * Xs).max() < 1e-4


def test_ceres_flavour_optimum_matches_scipy(oracle):
    kf, pt, px = G["kf_idx"], G["pt_idx"], G["px"]
    t_aa = []
    for v in G["logs_noisy"]:
        T = se3.se3_exp(v)
        t_aa.append(np.r_[T[:, 3], se3.so3_log(T[:, :3])])
    t_aa = np.array(t_aa)
    fixed = np.zeros(8, np.uint8)
    fixed[0] = 1
    P, X, st = oracle.local_ba_ceres(t_aa, fixed, G["pts_noisy"], kf, pt, px, max_iters=200)
    Ps, Xs, cost = _scipy_ceres(t_aa, G["pts_noisy"], kf, pt, px)
    assert abs(2 * st["cost_final"] - cost) < 1e-5 * cost                 # Ceres reports 1/2 sum r^2
    sc = float((X * Xs).sum() / (Xs * Xs).sum())                          # free global scale, see above
    assert np.abs(P[1:, 3:] - Ps[1:, 3:]).max() < 2e-3                    # rotations (function tolerance 1e-6 stops Ceres early)
    assert np.abs(P[1:, :3] - sc * Ps[1:, :3]).max() < 5e-3 and np.abs(X - sc * Xs).max() < 5e-3
