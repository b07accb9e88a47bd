#!/usr/bin/env python3
"""Generates tests/golden/cv2_orb.npz: OpenCV's OWN ORB orientation and descriptor on a synthetic frame, the third-party pin of
FeatureDetector::IC_Angle (FeatureDetector.cpp:509-537, incl. the `_umax` table the reference reads uninitialised, :304-322) and
FeatureDetector::ComputeOrbDescriptor (:539-578, the 256 x 4 `bit_pattern_31_` table of :37-295).  Both are copies of OpenCV's
orb.cpp (via ORB-SLAM), so cv2.ORB is the authority:

  * angle: cv2.ORB.detect computes ICAngles on the un-blurred level image; at octave 0 its key-points sit on integer pixels, so
    the oracle's angle at the same pixel must equal KeyPoint.angle bit for bit (float32);
  * descriptor: cv2.ORB.compute honours the angle of a provided key-point but first blurs its (bordered, in-place) pyramid
    level with a 7x7 sigma-2 Gaussian -- measured here to be the FLOAT separable filter (cv2.sepFilter2D with
    getGaussianKernel(7, 2)), not the fixed-point path cv2.GaussianBlur takes for a plain 8-bit image (those differ by +-1 in
    some pixels).  The reference does not blur, so the oracle is run on the blurred image J and OpenCV on the original I, with the
    oracle's angle on J handed to OpenCV: equal descriptors pin the pattern table, the rotation (float cos / sin), the cvRound taps
    and the bit order.

tests/test_golden.py re-creates J with numpy (double-precision separable convolution, reflect-101) and needs no cv2.
Run once here (cv2 4.13.0): python tools/make_orb_fixture.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle  # noqa: E402
from ygz_slam_b200 import synth  # noqa: E402


def orb_blur(img):
    g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    return cv2.sepFilter2D(img, cv2.CV_8U, g, g, borderType=cv2.BORDER_REFLECT_101)


def cv2_orb_pins(ora, img, n_features):
    """(px, py, OpenCV's angle on img, the oracle's angle on the blurred image, OpenCV's descriptor for that angle)."""
    orb = cv2.ORB_create(nfeatures=n_features, nlevels=1, edgeThreshold=31, patchSize=31, fastThreshold=20)
    kps = orb.detect(img, None)
    px = np.array([k.pt[0] for k in kps])
    py = np.array([k.pt[1] for k in kps])
    assert np.all(px == np.round(px)) and np.all(py == np.round(py)) and all(k.octave == 0 for k in kps)
    cv_angle = np.array([k.angle for k in kps], np.float32)
    h, w = img.shape
    J = orb_blur(img)
    ang_j, _ = ora.describe(ora.build_pyramid(J, 3), w, h, 3, px, py, np.zeros(len(kps), np.int32))
    given = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 0.0, 0) for x, y, a in zip(px, py, ang_j)]
    kept, cv_desc = orb.compute(img, given)
    assert len(kept) == len(given) and all(a.pt == b.pt and a.angle == b.angle for a, b in zip(given, kept))
    return px.astype(np.int16), py.astype(np.int16), cv_angle, ang_j, cv_desc


if __name__ == "__main__":
    ora = Oracle()
    img = synth.stream_frame(1)[0]
    px, py, cv_angle, ang_j, cv_desc = cv2_orb_pins(ora, img, 500)
    np.savez_compressed(ROOT / "tests" / "golden" / "cv2_orb.npz", cv2_version=np.array(cv2.__version__), frame=np.array(1), px=px, py=py,
                        cv_angle=cv_angle, angle_on_blurred=ang_j, cv_desc=cv_desc,
                        blur_kernel=cv2.getGaussianKernel(7, 2, cv2.CV_64F).ravel())
    print(len(px), "key-points")
