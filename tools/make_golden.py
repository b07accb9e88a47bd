#!/usr/bin/env python3
"""Generates tests/golden/cv2_pins.npz: outputs of OpenCV (cv2, the reference's own dependency for this arithmetic) on
small seeded inputs.  tests/test_golden.py checks the CPU oracle against them WITHOUT importing cv2, so the pin also
holds where cv2 is not installed.  Run once here (cv2 4.13.0): python tools/make_golden.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ygz_slam_b200 import synth  # noqa: E402

out = {"cv2_version": np.array(cv2.__version__)}
rng = np.random.default_rng(20240924)

# cvtColor(BGR2GRAY)
bgr = rng.integers(0, 256, (24, 40, 3), dtype=np.uint8)
out["bgr"] = bgr
out["bgr_gray"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)

# pyrDown on even / odd / tiny shapes
for i, shape in enumerate([(48, 64), (15, 21), (7, 9), (4, 5), (33, 34)]):
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    out[f"pyr_in_{i}"] = img
    out[f"pyr_out_{i}"] = cv2.pyrDown(img)

# fastAtan2 (degrees) on integer moments
yx = rng.integers(-40000, 40000, (3000, 2)).astype(np.float32)
yx[:8] = [[0, 0], [1, 1], [-3, 2], [0, 5], [5, 0], [0, -5], [-5, 0], [7, -7]]
out["atan2_yx"] = yx
out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)

# BFMatcher(NORM_HAMMING, crossCheck=True) with ties: descriptors drawn from a small pool of bytes
A = rng.integers(0, 4, (180, 32), dtype=np.uint8) * 85
B = rng.integers(0, 4, (150, 32), dtype=np.uint8) * 85
B[:20] = A[40:60]
B[20:25] = A[40:45]        # duplicates: two train rows tie on the same queries
A[100:105] = A[0:5]        # two queries tie on one train row
m = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(A, B)
idx = np.full(len(A), -1, np.int32)
dist = np.full(len(A), -1, np.int32)
for x in m:
    idx[x.queryIdx] = x.trainIdx
    dist[x.queryIdx] = int(x.distance)
out["bf_A"], out["bf_B"], out["bf_idx"], out["bf_dist"] = A, B, idx, dist

# calcOpticalFlowPyrLK with the reference's parameters (Tracker.cpp:92-98) on a 200 x 160 crop pair
tex = synth.texture(0x59475A00, 512)
g1 = tex[100:260, 120:320].copy()
g2 = tex[102:262, 117:317].copy()
pts = np.stack(np.meshgrid(np.arange(15, 190, 12), np.arange(15, 150, 12)), -1).reshape(-1, 2).astype(np.float32)
pts = np.concatenate([pts, np.array([[2.0, 3.0], [198.5, 80.0], [-10.0, 20.0]], np.float32)])
init = pts + np.float32(1.0)
nxt, st, err = cv2.calcOpticalFlowPyrLK(g1, g2, pts.copy(), init.copy(), winSize=(21, 21), maxLevel=4,
                                        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.001),
                                        flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
out["lk_g1"], out["lk_g2"], out["lk_pts"], out["lk_init"] = g1, g2, pts, init
out["lk_next"], out["lk_status"], out["lk_err"] = nxt.reshape(-1, 2), st.reshape(-1), err.reshape(-1)

dst = ROOT / "tests" / "golden" / "cv2_pins.npz"
dst.parent.mkdir(exist_ok=True)
np.savez_compressed(dst, **out)
print("wrote", dst, dst.stat().st_size, "bytes; cv2", cv2.__version__)
