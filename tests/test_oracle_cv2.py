"""CPU suite: pins the oracle's OpenCV-owned arithmetic against cv2 (the reference's own dependency,
importable here) and its FAST restatement structurally.  No GPU."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def test_bgr2gray_matches_cv2(oracle):
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    assert np.array_equal(oracle.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("shape", [(480, 640), (15, 20), (15, 21), (8, 10), (4, 5), (30, 40), (7, 9)])
def test_pyrdown_matches_cv2(oracle, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(oracle.pyrdown(img), cv2.pyrDown(img))


def test_pyramid_8_levels_matches_cv2(oracle, synth_frames):
    g = synth_frames[0][0]
    pyr = oracle.build_pyramid(g, 8)
    lv = g
    for L in range(1, 8):
        lv = cv2.pyrDown(lv)
        assert np.array_equal(oracle.level_view(pyr, 640, 480, 8, L), lv)
    assert lv.shape == (4, 5)


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.default_rng(1)
    ys = rng.integers(-40000, 40000, 5000).astype(np.float32)
    xs = rng.integers(-40000, 40000, 5000).astype(np.float32)
    for y, x in zip(ys, xs):
        assert oracle.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))
    assert oracle.fast_atan2(0.0, 0.0) == np.float32(cv2.fastAtan2(0.0, 0.0))


def test_fast_ring_geometry_against_cv2_fast9(oracle, synth_frames):
    """cv2 only ships FAST 9/16; with the arc length switched to 9 the oracle's segment test must
    give the same corner set, which pins the ring offsets, strictness and the scan range."""
    g = synth_frames[0][0]
    f = cv2.FastFeatureDetector_create(threshold=15, nonmaxSuppression=False, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    want = sorted((int(k.pt[1]), int(k.pt[0])) for k in f.detect(g))
    got = sorted((int(y), int(x)) for x, y in oracle.fast_detect(g, 15, arc=9))
    assert got == want and len(got) > 1000


def test_fast10_score_closed_form_and_raster_order(oracle, synth_frames):
    g = synth_frames[0][0]
    xy = oracle.fast_detect(g, 15)
    assert len(xy) > 1000
    lin = xy[:, 1].astype(np.int64) * 640 + xy[:, 0]
    assert np.all(np.diff(lin) > 0)  # raster order
    sc = oracle.fast_score(g, xy)
    ring = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
            (-3, 0), (-3, 1), (-2, 2), (-1, 3)]
    gi = g.astype(np.int32)
    for i in range(0, len(xy), 7):
        x, y = xy[i]
        d = [gi[y + dy, x + dx] - gi[y, x] for dx, dy in ring]
        best = max(max(min(d[(s + k) % 16] for k in range(10)), min(-d[(s + k) % 16] for k in range(10)))
                   for s in range(16))
        assert best - 1 == sc[i]
    assert sc.min() >= 15


def test_fast_nonmax_equals_dense_8_neighbour_rule(oracle, synth_frames):
    g = synth_frames[1][0]
    xy = oracle.fast_detect(g, 15)
    sc = oracle.fast_score(g, xy)
    nm = oracle.fast_nonmax(xy, sc)
    S = np.zeros((480, 640), np.int32)
    S[xy[:, 1], xy[:, 0]] = sc
    keep = []
    for i, (x, y) in enumerate(xy):
        nb = S[y - 1:y + 2, x - 1:x + 2].copy()
        nb[1, 1] = -1
        if not (nb >= sc[i]).any():
            keep.append(i)
    assert np.array_equal(np.array(keep, np.int32), nm)


def test_bf_match_matches_cv2(oracle, synth_frames):
    f1 = oracle.detect(oracle.build_pyramid(synth_frames[0][0], 3))
    f2 = oracle.detect(oracle.build_pyramid(synth_frames[1][0], 3))
    assert f1["n"] > 500 and f2["n"] > 500
    for cross in (True, False):
        idx, dist = oracle.match_bf(f1["desc"], f2["desc"], cross)
        ms = cv2.BFMatcher(cv2.NORM_HAMMING, cross).match(f1["desc"], f2["desc"])
        cvidx = -np.ones(f1["n"], np.int32)
        cvd = -np.ones(f1["n"], np.int32)
        for m in ms:
            cvidx[m.queryIdx] = m.trainIdx
            cvd[m.queryIdx] = int(m.distance)
        assert np.array_equal(idx, cvidx) and np.array_equal(dist, cvd)
    # ties: duplicated rows must resolve like OpenCV (first minimum; the later of two tied queries is dropped)
    A = np.repeat(f1["desc"][:50], 2, axis=0)
    B = np.repeat(f2["desc"][:60], 2, axis=0)
    idx, dist = oracle.match_bf(A, B, True)
    ms = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(A, B)
    cvidx = -np.ones(len(A), np.int32)
    for m in ms:
        cvidx[m.queryIdx] = m.trainIdx
    assert np.array_equal(idx, cvidx)


def test_descriptor_distance_is_popcount(oracle):
    rng = np.random.default_rng(3)
    A = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    d, keep, n = oracle.check_descriptors(A, B, np.arange(64), np.arange(64))
    want = np.unpackbits(A ^ B, axis=1).sum(1)
    assert np.array_equal(d, want)
    best = min(max(want.min(), 30), 100)
    assert np.array_equal(keep, want < 3.0 * best) and n == keep.sum()


def test_detect_semantics(oracle, synth_frames):
    """Grid selection invariants of FeatureDetector::Detect on a synthetic frame."""
    g = synth_frames[0][0]
    pyr = oracle.build_pyramid(g, 3)
    f = oracle.detect(pyr)
    assert 800 < f["n"] <= 3072
    # one feature per cell, emitted in cell order; pixel = level coordinate * 2^level
    assert np.all(np.diff(f["cell"]) > 0)
    gx = (f["px"] // 10).astype(int)
    gy = (f["py"] // 10).astype(int)
    assert np.array_equal(gy * 64 + gx, f["cell"])
    assert np.all(f["px"] % (1 << f["level"]) == 0)
    # the double-scaled InFrame (hazard 3): level-L coordinates must be >= 20 * 2^L
    lx = f["px"] / (1 << f["level"])
    assert np.all(lx >= 20 * (1 << f["level"]))
    assert set(np.unique(f["level"])) <= {0, 1, 2}
    # levels >= 3 can never contribute: an 8-level pyramid yields the same features
    f8 = oracle.detect(oracle.build_pyramid(g, 8), n_levels=8)
    assert f8["n"] == f["n"] and np.array_equal(f8["desc"], f["desc"])
    # occupied cells are skipped (SetExistingFeatures)
    occ = np.zeros(3072, np.uint8)
    occ[f["cell"][::2]] = 1
    f2 = oracle.detect(pyr, occupied=occ)
    assert np.array_equal(f2["cell"], f["cell"][1::2])
    # describe() on the detected pixels reproduces angle + descriptor
    ang, desc = oracle.describe(pyr, 640, 480, 3, f["px"], f["py"], f["level"])
    assert np.array_equal(ang, f["angle"]) and np.array_equal(desc, f["desc"])


def test_fast_nonmax_rule_is_switchable(oracle, synth_frames):
    """FAST-10 / fast_nonmax_3x3 stay PARITY UNPINNED (uzh-rpg/fast is not in the reference tree).  The one rule restated from
    memory -- a corner dies when a neighbour scores ">=" -- is a flag of the oracle: with ">" the survivors are a superset
    (ties keep both corners), everything else is unchanged."""
    g = synth_frames[0][0]
    xy = oracle.fast_detect(g)
    sc = oracle.fast_score(g, xy)
    keep_ge = oracle.fast_nonmax(xy, sc)
    try:
        oracle.lib.ora_set_fast_nonmax_strict(1)
        keep_gt = oracle.fast_nonmax(xy, sc)
    finally:
        oracle.lib.ora_set_fast_nonmax_strict(0)
    assert set(keep_ge.tolist()) <= set(keep_gt.tolist()) and len(keep_gt) >= len(keep_ge)
    assert np.array_equal(oracle.fast_nonmax(xy, sc), keep_ge)


@pytest.mark.parametrize("n", [57, 100, 777, 1500])
def test_initializer_sets_match_live_cv_rng(oracle, n):
    """cv::RNG pinned against OpenCV itself: cv2.setRNGSeed(0) resets theRNG() to the default-constructed state of the
    reference's `cv::RNG rng` (Initializer.cpp:25), cv2.randu on one CV_32S element with range [0, b) is one rng.uniform(0, b)
    (b not a power of two: those take randu's bit-mask path); the swap-with-last draw of Initializer.cpp:33-49 on top."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from make_cv_rng_fixture import sets_from_cv2
    assert np.array_equal(oracle.initializer_sets(n, 50), sets_from_cv2(n, 50))


@pytest.mark.parametrize("frame", [1, 2, 7])
def test_ic_angle_and_orb_descriptor_match_live_cv2_orb(oracle, frame):
    """FeatureDetector::IC_Angle / ComputeOrbDescriptor are OpenCV's ORB code (via ORB-SLAM): angle and descriptor of 2000
    key-points per frame equal cv2.ORB's bit for bit (tools/make_orb_fixture.py explains the blur and the provided angle)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from make_orb_fixture import cv2_orb_pins, orb_blur
    from ygz_slam_b200 import synth
    img = synth.stream_frame(frame)[0]
    px, py, cv_angle, ang_j, cv_desc = cv2_orb_pins(oracle, img, 2000)
    assert len(px) > 1000
    h, w = img.shape
    lvl = np.zeros(len(px), np.int32)
    ang, _ = oracle.describe(oracle.build_pyramid(img, 3), w, h, 3, px.astype(np.float64), py.astype(np.float64), lvl)
    assert np.array_equal(ang.view(np.uint32), cv_angle.view(np.uint32))
    _, desc_j = oracle.describe(oracle.build_pyramid(orb_blur(img), 3), w, h, 3, px.astype(np.float64), py.astype(np.float64), lvl)
    assert np.array_equal(desc_j, cv_desc)


@pytest.mark.parametrize("frame", [2, 5, 9])
def test_fast_score_and_nonmax_match_live_cv2_fast9(oracle, frame):
    """Detector (arc 9) -> closed-form score -> 3 x 3 suppression of the oracle == cv2 FAST 9/16 with nonmaxSuppression:
    same corners, same responses (tools/make_fast_fixture.py)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from make_fast_fixture import cv2_fast9_nms, score_closed_form
    from ygz_slam_b200 import synth
    img = synth.stream_frame(frame)[0]
    xy = oracle.fast_detect(img, 15, arc=9)
    sc = score_closed_form(img, xy, 9)
    keep = oracle.fast_nonmax(xy, sc)
    got = np.array(sorted((int(xy[i, 1]), int(xy[i, 0]), int(sc[i])) for i in keep), np.int32)
    assert np.array_equal(got, cv2_fast9_nms(img)) and len(got) > 2000


def test_shi_tomasi_agrees_with_cv2_min_eigenvalue(oracle, synth_frames):
    """FeatureDetector::ShiTomasiScore (:467-507) is the smaller eigenvalue of the 8 x 8 gradient matrix built from central
    differences over x in [u-4, u+4), y in [v-4, v+4), divided by 2 * 64: cv2.cornerMinEigenVal(blockSize = 8, ksize = 1) is the
    same quantity (same window anchor, Sobel aperture 1 = central difference) scaled by 1 / (8 * 255)^2, in float32 with another
    summation order -- so structure (window, anchor, formula) must agree to float precision; the 1-ulp pin is the numpy re-derivation."""
    g = synth_frames[2][0]
    ev = cv2.cornerMinEigenVal(g, blockSize=8, ksize=1, borderType=cv2.BORDER_REFLECT_101).astype(np.float64) * 2040.0 ** 2 / 128.0
    rng = np.random.default_rng(0)
    us, vs = rng.integers(20, 620, 600), rng.integers(20, 460, 600)
    got = np.array([oracle.shi_tomasi(g, u, v) for u, v in zip(us, vs)])
    want = ev[vs, us]
    assert np.all(np.abs(got - want) <= 2e-4 * np.maximum(1.0, np.abs(want)))
    assert got.max() > 100 and (got > 1).sum() > 100


# ---- two-view geometry of the Initializer against OpenCV's own decompositions ------------------------------------------
def _two_view_scenes():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import test_initializer as ti
    return ti


_KF = np.array([[np.float32(520.9), 0, np.float32(325.1)], [0, np.float32(521.0), np.float32(249.7)], [0, 0, 1.0]], np.float64)


@pytest.mark.parametrize("seed", [0, 3])
def test_decompose_e_and_triangulate_match_cv2(oracle, seed):
    """Initializer::DecomposeE (Initializer.cpp:643-675) and ::Triangulate (:615-641) are textbook constructions OpenCV also
    ships: the four (R, t) candidates of ReconstructF equal cv2.decomposeEssentialMat of E = K^T F K, the triangulated points equal
    cv2.triangulatePoints (the same DLT) for the selected pose."""
    ti = _two_view_scenes()
    p1, p2, R, t, n_out = ti.two_view(seed, noise=0.2)
    r = oracle.initializer_ransac(p1, p2, oracle.initializer_sets(len(p1), 200))
    q = oracle.initializer_reconstruct(p1, p2, 0, r["F21"], r["inliers_F"])
    assert q["ok"]
    R1, R2, tc = cv2.decomposeEssentialMat(_KF.T @ r["F21"] @ _KF)
    seen = set()
    for c in q["candidates"][:4]:
        Rc, tcand = c[:9].reshape(3, 3), c[9:]
        which = int(np.abs(Rc - R2).max() < np.abs(Rc - R1).max())
        sign = int(np.abs(tcand + tc.ravel()).max() < np.abs(tcand - tc.ravel()).max())
        assert np.abs(Rc - (R1, R2)[which]).max() < 1e-12 and np.abs(tcand - (1 - 2 * sign) * tc.ravel()).max() < 1e-12
        seen.add((which, sign))
    assert len(seen) == 4                                   # all four combinations, each once
    tri = q["triangulated"]
    P1 = _KF @ np.c_[np.eye(3), np.zeros(3)]
    P2 = _KF @ np.c_[q["R21"], q["t21"]]
    X = cv2.triangulatePoints(P1, P2, p1[tri].T.copy(), p2[tri].T.copy())
    X = (X[:3] / X[3]).T
    assert tri.sum() > 200 and np.abs(X - q["p3d"][tri]).max() < 1e-9 * np.abs(X).max()


def test_reconstruct_h_candidates_contain_cv2_decomposition(oracle):
    """ReconstructH (Initializer.cpp:330-513) enumerates Faugeras' eight (R, t, n) hypotheses; OpenCV's decomposeHomographyMat
    (Malis & Vargas, an analytically different route) returns the four physically distinct ones: each must be among the eight
    (rotation to 1e-6 -- the reference builds its rotations from float cos / sin --, translation direction up to sign)."""
    ti = _two_view_scenes()
    for seed, noise in ((0, 0.0), (1, 0.3)):
        p1, p2, R, t, _ = ti.planar_scene(seed, noise=noise)
        r = oracle.initializer_ransac(p1, p2, oracle.initializer_sets(len(p1), 200))
        q = oracle.initializer_reconstruct(p1, p2, 1, r["H21"], r["inliers_H"])
        n, Rs, ts, _ = cv2.decomposeHomographyMat(r["H21"], _KF)
        assert n == 4
        for Rc, tc in zip(Rs, ts):
            tn = tc.ravel() / np.linalg.norm(tc)
            errs = [max(np.abs(c[:9].reshape(3, 3) - Rc).max(), min(np.abs(c[9:] - tn).max(), np.abs(c[9:] + tn).max())) for c in q["candidates"]]
            assert min(errs) < 1e-6
