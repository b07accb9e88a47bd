"""CPU suite: the oracle against the committed golden vectors of tests/golden/cv2_pins.npz (OpenCV outputs generated once
by tools/make_golden.py).  Unlike tests/test_oracle_cv2.py this needs no cv2 at run time, so the pin of the
OpenCV-owned arithmetic (cvtColor, pyrDown, fastAtan2, BFMatcher cross-check, calcOpticalFlowPyrLK) always runs."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "cv2_pins.npz"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_bgr2gray_golden(oracle, gold):
    assert np.array_equal(oracle.bgr2gray(gold["bgr"]), gold["bgr_gray"])


def test_pyrdown_golden(oracle, gold):
    for i in range(5):
        assert np.array_equal(oracle.pyrdown(gold[f"pyr_in_{i}"]), gold[f"pyr_out_{i}"]), i


def test_fast_atan2_golden(oracle, gold):
    got = np.array([oracle.fast_atan2(y, x) for y, x in gold["atan2_yx"]], np.float32)
    assert np.array_equal(got.view(np.uint32), gold["atan2_deg"].view(np.uint32))


def test_bf_match_cross_check_golden(oracle, gold):
    idx, dist = oracle.match_bf(gold["bf_A"], gold["bf_B"], True)
    assert np.array_equal(idx, gold["bf_idx"]) and np.array_equal(dist, gold["bf_dist"])
    assert (gold["bf_idx"] >= 0).sum() > 20 and (gold["bf_idx"] < 0).sum() > 20   # both outcomes occur, ties included


def test_klt_golden(oracle, gold):
    got, st, err = oracle.klt(gold["lk_g1"], gold["lk_g2"], gold["lk_pts"], gold["lk_init"])
    wst = gold["lk_status"].astype(bool)
    st = st.astype(bool)
    assert (st != wst).sum() <= 1
    both = st & wst
    d = np.abs(got[both] - gold["lk_next"][both]).max(1)
    # OpenCV's SIMD build accumulates the f32 normal equations in another order: tolerance, not bit-exactness
    assert np.percentile(d, 99) < 0.01 and np.median(d) < 1e-3
    assert np.abs(err[both] - gold["lk_err"][both]).max() < 0.05
    assert both.sum() > 0.9 * len(wst) - 3


# ---- cv::RNG (the RANSAC minimal sets of Initializer::TryInitialize), tests/golden/cv_rng_sets.npz from tools/make_cv_rng_fixture.py
RNG_GOLD = Path(__file__).resolve().parent / "golden" / "cv_rng_sets.npz"


def _mwc_stream(count, state=0xFFFFFFFF):
    """cv::RNG::next() written from OpenCV's operations.hpp: state = (unsigned)state * CV_RNG_COEFF + (state >> 32)."""
    out = np.zeros(count, np.int64)
    for i in range(count):
        state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
        out[i] = state & 0xFFFFFFFF
    return out


def test_cv_rng_uniform_golden():
    """OpenCV's own draws (cv2.randu on theRNG() reset to the default state) are next() % b of the multiply-with-carry stream:
    this pins the coefficient, the default state and RNG::uniform(int, int) of the restatement every other test relies on."""
    g = np.load(RNG_GOLD)
    stream = _mwc_stream(64)
    for key in g.files:
        if key.startswith("uniform_"):
            b = int(key.split("_")[1])
            assert np.array_equal(g[key].astype(np.int64), stream % b), key


def test_initializer_sets_golden(oracle):
    """The oracle's minimal sets (a fresh default cv::RNG, draws without replacement) equal the sets drawn with OpenCV's generator."""
    g = np.load(RNG_GOLD)
    keys = [k for k in g.files if k.startswith("sets_")]
    assert len(keys) >= 3
    for key in keys:
        n = int(key.split("_")[1])
        assert np.array_equal(oracle.initializer_sets(n, 200), g[key]), key


# ---- OpenCV's own ORB orientation / descriptor (tests/golden/cv2_orb.npz from tools/make_orb_fixture.py)
ORB_GOLD = Path(__file__).resolve().parent / "golden" / "cv2_orb.npz"


def _orb_blur_numpy(img, kernel):
    """7 x 7 separable Gaussian, BORDER_REFLECT_101, in double precision, rounded to uint8 (what cv2.ORB applies before it
    computes descriptors; see tools/make_orb_fixture.py)."""
    r = len(kernel) // 2
    pad = np.pad(img.astype(np.float64), r, mode="reflect")
    rows = sum(kernel[k] * pad[:, k:k + img.shape[1]] for k in range(2 * r + 1))
    out = sum(kernel[k] * rows[k:k + img.shape[0], :] for k in range(2 * r + 1))
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def test_ic_angle_and_orb_descriptor_golden(oracle):
    """IC_Angle (with the canonical umax table) and ComputeOrbDescriptor equal OpenCV's ORB bit for bit on 500 key-points."""
    from ygz_slam_b200 import synth
    g = np.load(ORB_GOLD)
    img = synth.stream_frame(int(g["frame"]))[0]
    h, w = img.shape
    px, py = g["px"].astype(np.float64), g["py"].astype(np.float64)
    lvl = np.zeros(len(px), np.int32)
    ang, _ = oracle.describe(oracle.build_pyramid(img, 3), w, h, 3, px, py, lvl)
    assert np.array_equal(ang.view(np.uint32), g["cv_angle"].view(np.uint32))           # ICAngles of cv2.ORB.detect
    J = _orb_blur_numpy(img, g["blur_kernel"])
    ang_j, desc_j = oracle.describe(oracle.build_pyramid(J, 3), w, h, 3, px, py, lvl)
    assert np.array_equal(ang_j.view(np.uint32), g["angle_on_blurred"].view(np.uint32))  # (the blurred image is reproduced)
    assert np.array_equal(desc_j, g["cv_desc"])                                          # computeOrbDescriptors of cv2.ORB.compute
    assert len(px) == 500 and len(np.unique(desc_j, axis=0)) > 450


# ---- OpenCV's FAST 9/16 with non-maximum suppression (tests/golden/cv2_fast9.npz from tools/make_fast_fixture.py)
def _fast9_chain(oracle, img, threshold):
    """oracle detector (arc = 9) -> closed-form score -> oracle 3 x 3 suppression: rows of (y, x, score) in raster order."""
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from fast_closed_form import score_closed_form
    xy = oracle.fast_detect(img, threshold, arc=9)
    sc = score_closed_form(img, xy, 9)
    keep = oracle.fast_nonmax(xy, sc)
    return xy, sc, np.array(sorted((int(xy[i, 1]), int(xy[i, 0]), int(sc[i])) for i in keep), np.int32)


def test_fast_score_and_nonmax_golden(oracle):
    """Score definition and suppression rule of the oracle's FAST equal OpenCV's (arc 9, the only arc both implement); the
    arc-10 score the product is checked against is the same closed form with arc = 10."""
    from ygz_slam_b200 import synth
    g = np.load(Path(__file__).resolve().parent / "golden" / "cv2_fast9.npz")
    img = synth.stream_frame(int(g["frame"]))[0]
    _, _, got = _fast9_chain(oracle, img, int(g["threshold"]))
    assert np.array_equal(got, g["yx_score"].astype(np.int32)) and len(got) > 2000
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from fast_closed_form import score_closed_form
    xy10 = oracle.fast_detect(img, 15)
    assert np.array_equal(score_closed_form(img, xy10, 10), oracle.fast_score(img, xy10))
