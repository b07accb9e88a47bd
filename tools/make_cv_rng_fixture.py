#!/usr/bin/env python3
"""Generates tests/golden/cv_rng_sets.npz: the RANSAC minimal sets of Initializer::TryInitialize (Initializer.cpp:25-49) drawn
with OpenCV's OWN generator.  The reference draws `rng.uniform(0, avail.size())` from a default-constructed cv::RNG; cv2 exposes
that generator as theRNG(): cv2.setRNGSeed(0) resets it to RNG(0) -> state 0xffffffff (the default), and cv2.randu on a 1 x 1
CV_32S array with the range [0, b) consumes exactly one RNG::next() and returns next() % b for every b that is not a power of
two (OpenCV's integer randu takes a bit-mask path for powers of two, so the list sizes below avoid b in {2^k}).
tests/test_golden.py checks oracle/initializer.cpp's restated multiply-with-carry generator against these sets without cv2.
Run once here (cv2 4.13.0): python tools/make_cv_rng_fixture.py"""
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def sets_from_cv2(n, iters):
    cv2.setRNGSeed(0)
    one = np.zeros((1, 1), np.int32)
    sets = np.zeros((iters, 8), np.int32)
    for it in range(iters):
        avail = list(range(n))
        for j in range(8):
            b = len(avail)
            assert b & (b - 1), "a power-of-two range takes randu's bit-mask path"
            cv2.randu(one, 0, b)
            r = int(one[0, 0])
            sets[it, j] = avail[r]            # draw without replacement: swap with the last, pop (Initializer.cpp:40-46)
            avail[r] = avail[-1]
            avail.pop()
    return sets


if __name__ == "__main__":
    out = {"cv2_version": np.array(cv2.__version__)}
    for n in (12, 57, 300, 1500):             # 12: the reference's own test scene (test_initializer.cpp) -- 12..5 contains 8
        if any((b & (b - 1)) == 0 for b in range(n - 7, n + 1)):
            continue
        out[f"sets_{n}"] = sets_from_cv2(n, 200)   # 200 = the reference's max_iter (Initializer.cpp:22)
    # raw uniform draws for a few moduli: next() % b of the first 64 numbers of the default stream
    for b in (3, 57, 1500, 999983, 2147483647):
        cv2.setRNGSeed(0)
        a = np.zeros((1, 64), np.int32)
        cv2.randu(a, 0, b)
        out[f"uniform_{b}"] = a[0].copy()
    np.savez_compressed(ROOT / "tests" / "golden" / "cv_rng_sets.npz", **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})
