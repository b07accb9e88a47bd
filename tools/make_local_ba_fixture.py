#!/usr/bin/env python3
"""Generates tests/golden/local_ba_cvrng.npz: the synthetic scene of the reference's test/test_local_ba.cpp:9-98 with
EXACTLY its noise -- the test draws from a default-constructed cv::RNG (state 0xffffffff) with rng.gaussian(sigma), which
OpenCV's Python binding reproduces: cv2.setRNGSeed(0) resets theRNG() to that default state (RNG(0) -> 0xffffffff) and
cv2.randn fills from the same ziggurat stream (randn_0_1_32f, value = float normal * sigma), one draw per element.
Draw order of the test: for key-frames 1..7 six pose components (sigma 0.1); then per map point three coordinates
(sigma 0.1) followed, per key-frame 0..7, by the pixel noise Vector2d(rng.gaussian(1), rng.gaussian(1)).  The evaluation
order of those two constructor arguments is unspecified in C++; this fixture takes u first (documented choice).
Run once here (cv2 4.13.0): python tools/make_local_ba_fixture.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ygz_slam_b200 import se3  # noqa: E402

FX, FY, CX, CY = (np.float32(v).astype(np.float64) for v in (520.9, 521.0, 325.1, 249.7))   # PinholeCamera stores floats (Camera.h:14-22)

rot = [(0, 0, 0), (0.1, 0, 0), (0, 0.1, 0), (0, 0, 0.1), (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0)]
tr = [(0, 0, 0)] * 4 + [(0.1, 0, 0), (0, 0.1, 0), (0, 0, 0.1), (0.1, 0.1, 0.1)]
# SE3(SO3::exp(w), t): rotation from the so3 exponential, translation given directly (not the se3 exponential's V * upsilon)
poses_true = []
for w, t in zip(rot, tr):
    T = se3.se3_exp(np.r_[np.zeros(3), w])
    T[:, 3] = t
    poses_true.append(T)
pts_true = np.array([[x, y, z] for z in (2, 3, 4, 5) for (x, y) in ((0, 0), (0, 1), (1, 0), (1, 1))], np.float64)

cv2.setRNGSeed(0)            # theRNG() = RNG(0) -> state 0xffffffff, the default-constructed cv::RNG of the test
draws = np.zeros((7 * 6 + 16 * (3 + 8 * 2), 1), np.float64)
cv2.randn(draws, 0.0, 1.0)   # unit normals in stream order; gaussian(sigma) = normal * sigma
d = iter(draws.ravel())

logs_true = np.array([se3.se3_log(T) for T in poses_true])   # Sophus order [upsilon; omega]
logs_noisy = logs_true.copy()
for i in range(1, 8):
    for j in range(6):
        logs_noisy[i, j] = logs_true[i, j] + next(d) * 0.1
pts_noisy = pts_true.copy()
kf_idx, pt_idx, px = [], [], []
for i in range(16):
    for j in range(3):
        pts_noisy[i, j] += next(d) * 0.1
    for k in range(8):
        pc = poses_true[k][:, :3] @ pts_true[i] + poses_true[k][:, 3]
        u, v = FX * pc[0] / pc[2] + CX, FY * pc[1] / pc[2] + CY     # PinholeCamera::World2Pixel
        nu = next(d) * 1.0
        nv = next(d) * 1.0
        kf_idx.append(k)
        pt_idx.append(i)
        px.append([u + nu, v + nv])
assert next(d, None) is None
np.savez(ROOT / "tests" / "golden" / "local_ba_cvrng.npz", cv2_version=np.array(cv2.__version__), unit_normals=draws.ravel(),
         logs_true=logs_true, logs_noisy=logs_noisy, pts_true=pts_true, pts_noisy=pts_noisy, kf_idx=np.array(kf_idx, np.int32),
         pt_idx=np.array(pt_idx, np.int32), px=np.array(px, np.float64))
print("first unit normals of the default cv::RNG:", draws.ravel()[:4])
