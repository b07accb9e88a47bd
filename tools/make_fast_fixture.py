#!/usr/bin/env python3
"""Generates tests/golden/cv2_fast9.npz: OpenCV's FAST 9/16 WITH non-maximum suppression on a synthetic frame (positions and
responses).  uzh-rpg `fast` (the reference's FAST-10, call sites FeatureDetector.cpp:365-381) is absent and OpenCV ships no arc-10
detector, but both descend from Rosten's generated code and share everything except the arc length: the 16-pixel ring, the strict
comparisons, the score ("the largest threshold for which the pixel is still a corner" = max over arcs of the smallest |difference|
on the arc, minus 1) and the 3 x 3 suppression (a corner survives iff its score is strictly greater than the score of every
neighbouring corner).  The oracle's detector takes the arc length as a parameter, its suppression takes any score list, so with
arc = 9 the whole chain must reproduce OpenCV: tests/test_golden.py (no cv2) and tests/test_oracle_cv2.py (live).
Run once here (cv2 4.13.0): python tools/make_fast_fixture.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from ygz_slam_b200 import synth  # noqa: E402

from fast_closed_form import score_closed_form  # noqa: E402,F401  (cv2-free helper shared with the tests)


def cv2_fast9_nms(img, threshold=15):
    f = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = f.detect(img)
    out = np.array(sorted((int(k.pt[1]), int(k.pt[0]), int(k.response)) for k in kps), np.int32)   # raster order: (y, x, score)
    return out


if __name__ == "__main__":
    frame = 1
    img = synth.stream_frame(frame)[0]
    kp = cv2_fast9_nms(img)
    np.savez_compressed(ROOT / "tests" / "golden" / "cv2_fast9.npz", cv2_version=np.array(cv2.__version__), frame=np.array(frame),
                        threshold=np.array(15), yx_score=kp.astype(np.int16))
    print(len(kp), "corners after suppression")
