#!/usr/bin/env python3
"""Generates tests/golden/initializer_cvrng.npz: the two synthetic scenes of the reference's test/test_initializer.cpp:10-94
(12 landmarks on the plane z = 2 for the homography case, 12 landmarks on z = 2, 3, 4 for the fundamental-matrix case, cameras at
the origin and at t = (1, 0, 0), pixel noise sigma 2) with EXACTLY its noise: the test draws rng.gaussian(2.0) from a
default-constructed cv::RNG, which cv2 reproduces (cv2.setRNGSeed(0) -> state 0xffffffff, cv2.randn = the same ziggurat stream;
see tools/make_local_ba_fixture.py).  Draw order of the test, per landmark i: px1H, px2H, px1F, px2F, each
Vector2d(rng.gaussian(2), rng.gaussian(2)); the evaluation order of the two constructor arguments is unspecified in C++ -- this
fixture takes u first (documented choice, as in the local-BA fixture).  Run once here (cv2 4.13.0)."""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
FX, FY, CX, CY = (np.float32(v).astype(np.float64) for v in (520.9, 521.0, 325.1, 249.7))   # PinholeCamera stores floats

landmarks_H = np.array([[x, y, 2] for x in (-1, 0, 1) for y in (-1, 0, 1, 2)], np.float64)
landmarks_F = np.array([[x, y, z] for z in (2, 3, 4) for (x, y) in ((-1, -1), (-1, 1), (1, -1), (1, 1))], np.float64)
t2 = np.array([1.0, 0.0, 0.0])   # pose2 = SE3(SO3::exp(0), (1, 0, 0)): p_c2 = p_w + t2


def pixel(p):
    return np.array([FX * p[0] / p[2] + CX, FY * p[1] / p[2] + CY])


cv2.setRNGSeed(0)
draws = np.zeros((12 * 8, 1), np.float64)
cv2.randn(draws, 0.0, 1.0)
d = iter(draws.ravel())
px = {k: np.zeros((12, 2)) for k in ("px1H", "px2H", "px1F", "px2F")}
for i in range(12):
    for key, lm, t in (("px1H", landmarks_H, 0 * t2), ("px2H", landmarks_H, t2), ("px1F", landmarks_F, 0 * t2), ("px2F", landmarks_F, t2)):
        clean = pixel(lm[i] + t)
        nu = next(d) * 2.0
        nv = next(d) * 2.0
        px[key][i] = clean + [nu, nv]
assert next(d, None) is None
np.savez(ROOT / "tests" / "golden" / "initializer_cvrng.npz", cv2_version=np.array(cv2.__version__), unit_normals=draws.ravel(),
         landmarks_H=landmarks_H, landmarks_F=landmarks_F, t2=t2, **px)
print("first unit normals:", draws.ravel()[:4], "px1F[0]", px["px1F"][0])
