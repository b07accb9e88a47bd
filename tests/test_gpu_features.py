"""GPU parity suite for the extract + match path (BASELINE configs C1/C2): every call goes through the
C ABI (ctypes) and is compared bit-exactly with the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from ygz_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _upload(ctx, imgs):
    fr = ctx.frames(len(imgs))
    fr.upload(np.stack(imgs))
    return fr


def test_pyramid_bit_exact(ctx8, oracle, synth_frames):
    imgs = [f[0] for f in synth_frames]
    fr = _upload(ctx8, imgs)
    for s, g in enumerate(imgs):
        pyr = oracle.build_pyramid(g, 8)
        for L in range(8):
            assert np.array_equal(fr.download_level(s, L), oracle.level_view(pyr, 640, 480, 8, L)), (s, L)
    fr.close()


def test_bgr_upload_matches_grey(ctx3, oracle):
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (2, 480, 640, 3), dtype=np.uint8)
    fr = ctx3.frames(2)
    fr.upload(bgr)
    for s in range(2):
        gray = oracle.bgr2gray(bgr[s])
        assert np.array_equal(fr.download_level(s, 0), gray)
        assert np.array_equal(fr.download_level(s, 1), oracle.pyrdown(gray))
    fr.close()


def test_fast_keypoint_indices_bit_exact(ctx8, oracle, synth_frames):
    """Corner list (raster order), bisection scores and fast_nonmax_3x3 indices, per level."""
    g = synth_frames[0][0]
    fr = _upload(ctx8, [g])
    pyr = oracle.build_pyramid(g, 8)
    for L in range(8):
        img = oracle.level_view(pyr, 640, 480, 8, L)
        xy = oracle.fast_detect(img, 15)
        sc = oracle.fast_score(img, xy)
        nm = oracle.fast_nonmax(xy, sc)
        gxy, gsc, gnm = fr.fast_debug(0, L)
        assert np.array_equal(gxy, xy), L
        assert np.array_equal(gsc, sc), L
        assert np.array_equal(gnm, nm), L
    fr.close()


def _assert_features_equal(got, want):
    assert got["n"] == want["n"]
    for k in ("px", "py", "level", "cell"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32))
    assert np.array_equal(got["angle"].view(np.uint32), want["angle"].view(np.uint32))
    assert np.array_equal(got["desc"], want["desc"])


@pytest.mark.parametrize("levels", [3, 8])
def test_detect_bit_exact(levels, ctx3, ctx8, oracle, synth_frames):
    ctx = ctx3 if levels == 3 else ctx8
    imgs = [f[0] for f in synth_frames]
    fr = _upload(ctx, imgs)
    got = fr.detect([0, 1, 2])
    stats = fr.detect_stats(3)
    for s, g in enumerate(imgs):
        pyr = oracle.build_pyramid(g, levels)
        want = oracle.detect(pyr, n_levels=levels)
        _assert_features_equal(got[s], want)
        for L in range(levels):
            img = oracle.level_view(pyr, 640, 480, levels, L)
            xy = oracle.fast_detect(img, 15)
            nm = oracle.fast_nonmax(xy, oracle.fast_score(img, xy))
            assert stats[s, L, 0] == len(xy) and stats[s, L, 1] == len(nm), (s, L)
    # slot indirection + occupied cells (overwrite_existing_features = false)
    occ = np.zeros((2, 3072), np.uint8)
    occ[0, got[2]["cell"][::3]] = 1
    occ[1, ::2] = 1
    got2 = fr.detect([2, 0], occupied=occ)
    for i, s in enumerate((2, 0)):
        want = oracle.detect(oracle.build_pyramid(imgs[s], levels), n_levels=levels, occupied=occ[i])
        _assert_features_equal(got2[i], want)
    fr.close()


def test_detect_edge_images(ctx3, oracle):
    """Flat image (no corners), saturated noise (many corners), checkerboard (score ties)."""
    rng = np.random.default_rng(9)
    flat = np.full((480, 640), 77, np.uint8)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    yy, xx = np.mgrid[0:480, 0:640]
    checker = ((((yy // 7) + (xx // 7)) % 2) * 200 + 20).astype(np.uint8)
    imgs = [flat, noise, checker]
    fr = _upload(ctx3, imgs)
    got = fr.detect([0, 1, 2])
    for s, g in enumerate(imgs):
        want = oracle.detect(oracle.build_pyramid(g, 3))
        _assert_features_equal(got[s], want)
    assert got[0]["n"] == 0
    fr.close()


def test_describe_matches_oracle(ctx3, oracle, synth_frames):
    """ComputeAngleAndDescriptor on caller pixels, incl. non-integer pixels and the L1 row-wrap zone."""
    g = synth_frames[0][0]
    fr = _upload(ctx3, [g, synth_frames[1][0]])
    rng = np.random.default_rng(2)
    n = 600
    level = rng.integers(0, 3, n).astype(np.uint8)
    px = rng.uniform(40, 600, n)
    py = rng.uniform(40, 440, n)
    # force some level-1/2 features next to the right/bottom border (hazard 4: taps wrap / leave the buffer)
    px[:40] = rng.uniform(600, 636, 40)
    py[40:80] = rng.uniform(450, 478, 40)
    level[:80] = rng.integers(1, 3, 80)
    pyr0 = oracle.build_pyramid(g, 3)
    pyr1 = oracle.build_pyramid(synth_frames[1][0], 3)
    ang, desc = fr.describe([0, 1], [0, 400, n], px, py, level)
    a0, d0 = oracle.describe(pyr0, 640, 480, 3, px[:400], py[:400], level[:400].astype(np.int32))
    a1, d1 = oracle.describe(pyr1, 640, 480, 3, px[400:], py[400:], level[400:].astype(np.int32))
    assert np.array_equal(ang.view(np.uint32), np.concatenate([a0, a1]).view(np.uint32))
    assert np.array_equal(desc, np.concatenate([d0, d1]))
    fr.close()


def test_match_bf_bit_exact(ctx3, oracle, synth_frames):
    f1 = oracle.detect(oracle.build_pyramid(synth_frames[0][0], 3))
    f2 = oracle.detect(oracle.build_pyramid(synth_frames[1][0], 3))
    for cross in (True, False):
        idx, dist = ctx3.match_bf(f1["desc"], f2["desc"], cross)
        widx, wdist = oracle.match_bf(f1["desc"], f2["desc"], cross)
        assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)
    # ragged / tiny / tie-heavy inputs
    rng = np.random.default_rng(4)
    for nA, nB in [(1, 1), (1, 700), (700, 1), (129, 257), (1000, 1000), (3072, 3072), (5, 0)]:
        A = rng.integers(0, 256, (nA, 32), dtype=np.uint8)
        B = rng.integers(0, 4, (nB, 32), dtype=np.uint8) if nB else np.zeros((0, 32), np.uint8)
        if nA > 4 and nB > 4:
            A[3] = A[1]
            B[4] = B[2]
        idx, dist = ctx3.match_bf(A, B, True)
        widx, wdist = oracle.match_bf(A, B, True)
        assert np.array_equal(idx, widx) and np.array_equal(dist, wdist), (nA, nB)
    # CheckFrameDescriptors distances
    ia = rng.integers(0, f1["n"], 500)
    ib = rng.integers(0, f2["n"], 500)
    d = ctx3.hamming_pairs(f1["desc"], f2["desc"], ia, ib)
    wd, _, _ = oracle.check_descriptors(f1["desc"], f2["desc"], ia, ib)
    assert np.array_equal(d, wd)


def test_match_frames_device_resident(ctx3, oracle, synth_frames):
    """test_orb_match shape: detect on two frames, cross-checked BF match of the device-resident descriptors."""
    imgs = [f[0] for f in synth_frames]
    fr = _upload(ctx3, imgs)
    feats = fr.detect([0, 1, 2])
    res = fr.match([0, 1, 2], [1, 2, 0], True)
    for (a, b), (idx, dist) in zip([(0, 1), (1, 2), (2, 0)], res):
        widx, wdist = oracle.match_bf(feats[a]["desc"], feats[b]["desc"], True)
        assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)
        keep, n_good = oracle.good_matches(idx, dist)
        assert n_good > 300  # the synthetic pair really overlaps
    fr.close()


@pytest.mark.parametrize("w,h,levels,cell", [(752, 480, 4, 10), (324, 246, 3, 8), (1281, 721, 5, 20), (640, 480, 1, 10), (640, 480, 2, 10),
                                              (128, 96, 3, 8), (200, 150, 8, 10)])
def test_other_geometries_bit_exact(oracle, w, h, levels, cell):
    """Image sizes other than 640x480: widths that are not multiples of 8 / 16 (tiled pyrDown fallback, word-load FAST
    staging), ragged FAST tiles, odd level sizes, other grid cell sizes -- pyramid, corner statistics, features and
    matches must still be bit-exact."""
    from ygz_slam_b200 import Context
    rng = np.random.default_rng(w * 7 + h)
    tex = synth.texture(0x59475A00, 2048)
    imgs = []
    for k in range(2):
        y0, x0 = 100 + 37 * k, 60 + 11 * k
        g = tex[y0:y0 + h, x0:x0 + w].astype(np.int16) + rng.integers(-3, 4, (h, w))
        imgs.append(np.clip(g, 0, 255).astype(np.uint8))
    ctx = Context(0, image_width=w, image_height=h, n_levels=levels, cell_size=cell)
    try:
        fr = ctx.frames(2)
        fr.upload(np.stack(imgs))
        got = fr.detect([0, 1])
        stats = fr.detect_stats(2)
        wants = []
        for s, g in enumerate(imgs):
            pyr = oracle.build_pyramid(g, levels)
            for L in range(levels):
                lv = oracle.level_view(pyr, w, h, levels, L)
                assert np.array_equal(fr.download_level(s, L), lv), (s, L)
                xy = oracle.fast_detect(lv, 15)
                nm = oracle.fast_nonmax(xy, oracle.fast_score(lv, xy))
                assert stats[s, L, 0] == len(xy) and stats[s, L, 1] == len(nm), (s, L)
            want = oracle.detect(pyr, w=w, h=h, n_levels=levels, cell=cell)
            _assert_features_equal(got[s], want)
            assert want["n"] > (50 if w >= 300 else 0)
            wants.append(want)
        idx, dist = fr.match([0], [1], True)[0]
        widx, wdist = oracle.match_bf(wants[0]["desc"], wants[1]["desc"], True)
        assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)
        fr.close()
    finally:
        ctx.close()
