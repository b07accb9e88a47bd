"""Initializer RANSAC: FindHomography / FindFundamental (Initializer.cpp:89-318, 670-853).

The oracle (oracle/initializer.cpp) is checked against numpy's LAPACK SVD (an independent solver of the same 8-point
systems: models agree up to sign / rounding, scores to float accuracy) and against ground-truth two-view geometry; the CUDA
path performs the oracle's operations in the same order without FMA contraction and must agree BIT FOR BIT (models, float
scores, winning iterations, inlier flags).
"""
from pathlib import Path

import numpy as np
import pytest

from oracle.pyoracle import Oracle

K = np.array([[520.9, 0, 325.1], [0, 521.0, 249.7], [0, 0, 1.0]])


def two_view(seed, n=300, planar=False, noise=0.5, outliers=0.1):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(3, 8, n)]
    if planar:
        X[:, 2] = 5.0 + 0.1 * X[:, 0]
    th = 0.05
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    t = np.array([0.3, 0.02, 0.05])

    def proj(Rm, tv):
        Y = (Rm @ X.T).T + tv
        return (K @ (Y / Y[:, 2:]).T).T[:, :2]

    p1, p2 = proj(np.eye(3), np.zeros(3)), proj(R, t) + rng.normal(0, noise, (n, 2))
    n_out = int(outliers * n)
    p2[:n_out] += rng.uniform(-40, 40, (n_out, 2))
    return p1, p2, R, t, n_out


@pytest.fixture(scope="module")
def ora():
    return Oracle()


def test_sets_are_the_mwc_sequence(ora):
    """cv::RNG restated: multiply-with-carry with CV_RNG_COEFF, default state, uniform(0, b) = next() % b; draws without
    replacement by swap-with-last (Initializer.cpp:33-49)."""
    state = 0xFFFFFFFF
    n, iters = 57, 30
    want = np.zeros((iters, 8), np.int64)
    for it in range(iters):
        avail = list(range(n))
        for j in range(8):
            state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
            r = (state & 0xFFFFFFFF) % len(avail)
            want[it, j] = avail[r]
            avail[r] = avail[-1]
            avail.pop()
    got = ora.initializer_sets(n, iters)
    assert np.array_equal(got, want)
    assert all(len(set(row)) == 8 for row in got)


def numpy_models(p1, p2, sets):
    """The same 8-point systems through numpy's SVD (LAPACK), written from Initializer.cpp:140-239, 730-762."""
    def normalize(px):
        mean = px.mean(0)
        d = px - mean
        s = (1.0 / np.abs(d).mean(0)).astype(np.float32).astype(np.float64)
        T = np.array([[s[0], 0, -mean[0] * s[0]], [0, s[1], -mean[1] * s[1]], [0, 0, 1]])
        return d * s, T
    n1, T1 = normalize(p1)
    n2, T2 = normalize(p2)
    out = []
    for s in sets:
        A = []
        for (u1, v1), (u2, v2) in zip(n1[s], n2[s]):
            A.append([0, 0, 0, -u1, -v1, -1, v2 * u1, v2 * v1, v2])
            A.append([u1, v1, 1, 0, 0, 0, -u2 * u1, -u2 * v1, -u2])
        Hn = np.linalg.svd(np.array(A))[2][-1].reshape(3, 3)
        H = np.linalg.inv(T2) @ Hn @ T1
        B = [[u2 * u1, u2 * v1, u2, v2 * u1, v2 * v1, v2, u1, v1, 1] for (u1, v1), (u2, v2) in zip(n1[s], n2[s])]
        Fp = np.linalg.svd(np.array(B))[2][-1].reshape(3, 3)
        U, S, Vt = np.linalg.svd(Fp)
        F = T2.T @ (U @ np.diag([S[0], S[1], 0]) @ Vt) @ T1
        out.append((H, F))
    return out


def test_oracle_models_vs_lapack(ora):
    p1, p2, _, _, _ = two_view(3, n=120)
    sets = ora.initializer_sets(len(p1), 25)
    r = ora.initializer_ransac(p1, p2, sets, models=True)
    for it, (H, F) in enumerate(numpy_models(p1, p2, sets)):
        for got, want in ((r["models"][it, :9].reshape(3, 3), H), (r["models"][it, 9:].reshape(3, 3), F)):
            sgn = np.sign((got * want).sum())
            assert np.abs(got - sgn * want).max() < 1e-9 * np.abs(want).max(), it
        assert np.linalg.svd(r["models"][it, 9:].reshape(3, 3))[1][2] < 1e-12      # rank 2


def test_oracle_recovers_the_geometry(ora):
    # general scene: the fundamental matrix wins (rh = sh / (sh + sf) <= 0.4, Initializer.cpp:66-78) and its inliers are the true ones
    p1, p2, R, t, n_out = two_view(0)
    r = ora.initializer_ransac(p1, p2, ora.initializer_sets(len(p1), 200))
    assert r["best_F"] >= 0 and r["best_H"] >= 0
    assert r["score_H"] / (r["score_H"] + r["score_F"]) < 0.4
    assert r["inliers_F"][n_out:].mean() > 0.9 and r["inliers_F"][:n_out].mean() < 0.3
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ft = np.linalg.inv(K).T @ tx @ R @ np.linalg.inv(K)
    x1, x2 = np.c_[p1, np.ones(len(p1))], np.c_[p2, np.ones(len(p2))]
    l2 = (r["F21"] @ x1.T).T
    d = np.abs((x2 * l2).sum(1)) / np.hypot(l2[:, 0], l2[:, 1])
    assert np.median(d[n_out:]) < 1.5                                            # epipolar distance of the true matches (px)
    assert np.abs((x2 * (Ft @ x1.T).T).sum(1)).max() < 1e3                        # (ground truth is consistent)
    # planar scene: the homography explains every true match.  (CheckHomography adds ONE term per point, CheckFundamental two, so
    # sh / (sh + sf) stays near 1/3 even here -- a property of the reference, Initializer.cpp:284-303 vs :798-850, kept.)
    p1, p2, _, _, n_out = two_view(1, planar=True)
    r = ora.initializer_ransac(p1, p2, ora.initializer_sets(len(p1), 200))
    assert 0.25 < r["score_H"] / (r["score_H"] + r["score_F"]) < 0.4
    assert r["inliers_H"][n_out:].mean() > 0.9
    y = (r["H21"] @ np.c_[p1, np.ones(len(p1))].T).T
    y = y[:, :2] / y[:, 2:]
    assert np.median(np.linalg.norm(y - p2, axis=1)[n_out:]) < 1.5


def test_oracle_without_a_scoring_model(ora):
    # gross mismatches only: every hypothesis may still score a little; a list where nothing scores reports best = -1
    rng = np.random.default_rng(5)
    p1 = rng.uniform(0, 600, (8, 2))
    p2 = rng.uniform(0, 600, (8, 2))
    r = ora.initializer_ransac(p1, p2, ora.initializer_sets(8, 5))
    assert r["best_F"] >= -1 and r["best_H"] >= -1
    if r["best_H"] < 0:
        assert not r["H21"].any() and not r["inliers_H"].any()


def planar_scene(seed, n=300, noise=0.0):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), np.zeros(n)]
    X[:, 2] = 5.0 + 0.3 * X[:, 0]
    th = 0.1
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    t = np.array([1.5, 0.2, 0.3])

    def proj(Rm, tv):
        Y = (Rm @ X.T).T + tv
        return (K @ (Y / Y[:, 2:]).T).T[:, :2]

    return proj(np.eye(3), np.zeros(3)), proj(R, t) + rng.normal(0, noise, (n, 2)), R, t, X


def test_oracle_reconstruct_f_recovers_pose_and_structure(ora):
    """ReconstructF / DecomposeE / CheckRT / Triangulate (Initializer.cpp:515-675, 855-963) on a general scene."""
    p1, p2, R, t, n_out = two_view(0, noise=0.2)
    r = ora.initializer_ransac(p1, p2, ora.initializer_sets(len(p1), 200))
    q = ora.initializer_reconstruct(p1, p2, 0, r["F21"], r["inliers_F"])
    assert q["ok"]
    assert np.linalg.norm(q["R21"] - R) < 0.01 and abs(np.linalg.det(q["R21"]) - 1) < 1e-12
    assert np.linalg.norm(q["t21"] - t / np.linalg.norm(t)) < 0.1 and abs(np.linalg.norm(q["t21"]) - 1) < 1e-12
    assert q["n_good"][:4].argmax() == int(np.flatnonzero(q["n_good"][:4] == q["n_good"][:4].max())[0])
    assert sorted(q["n_good"][:4])[-2] < 0.7 * q["n_good"][:4].max()          # an unambiguous winner among the four candidates
    tri = q["triangulated"]
    assert tri[n_out:].mean() > 0.9
    # the triangulated points reproject onto their pixels in both views (scale = |t| = 1)
    X1 = q["p3d"][tri]
    u1 = (K @ (X1 / X1[:, 2:]).T).T[:, :2]
    Y = (q["R21"] @ X1.T).T + q["t21"]
    u2 = (K @ (Y / Y[:, 2:]).T).T[:, :2]
    assert np.median(np.linalg.norm(u1 - p1[tri], axis=1)) < 1.0 and np.median(np.linalg.norm(u2 - p2[tri], axis=1)) < 1.0
    assert q["parallax"] > 1.0


def test_oracle_reconstruct_h_candidates(ora):
    """ReconstructH (Initializer.cpp:330-513): Faugeras' eight candidates contain the true motion exactly on a noise-free plane;
    a plane seen entirely in front of both cameras leaves the classical two-fold ambiguity, which the reference's
    `secondBestGood < 0.75 bestGood` test rejects."""
    p1, p2, R, t, _ = planar_scene(0)
    r = ora.initializer_ransac(p1, p2, ora.initializer_sets(len(p1), 200))
    q = ora.initializer_reconstruct(p1, p2, 1, r["H21"], r["inliers_H"])
    errs = [np.linalg.norm(c[:9].reshape(3, 3) - R) + np.linalg.norm(c[9:] - t / np.linalg.norm(t)) for c in q["candidates"]]
    assert min(errs) < 1e-6
    for c in q["candidates"]:
        assert abs(np.linalg.det(c[:9].reshape(3, 3)) - 1) < 1e-6 and abs(np.linalg.norm(c[9:]) - 1) < 1e-9   # (float cos / sin in the reference)
    assert sorted(q["n_good"])[-1] == len(p1) and sorted(q["n_good"])[-2] == len(p1) and not q["ok"]
    assert not q["triangulated"].any() and not q["R21"].any()


GOLDEN = Path(__file__).resolve().parent / "golden" / "initializer_cvrng.npz"


def try_initialize(run_ransac, run_reconstruct, p1, p2, sets):
    """Initializer::TryInitialize (Initializer.cpp:9-87) on top of the two entry points."""
    r = run_ransac(p1, p2, sets)
    use_h = bool(r["score_H"] / (r["score_H"] + r["score_F"]) > 0.4)
    q = run_reconstruct(p1, p2, use_h, r["H21"] if use_h else r["F21"], r["inliers_H"] if use_h else r["inliers_F"])
    return use_h, r, q


def test_reference_test_initializer_scene(ora):
    """The reference's own test, test/test_initializer.cpp:10-140, with EXACTLY its data: landmarks on three depth layers, second
    camera at t = (1, 0, 0), pixel noise sigma 2 drawn from the default cv::RNG (tools/make_initializer_fixture.py).  Like the
    test: TryInitialize on the F scene, then ba::TwoViewBACeres on the result, scale normalised by |t|."""
    fx = np.load(GOLDEN)
    p1, p2 = fx["px1F"], fx["px2F"]
    use_h, r, q = try_initialize(ora.initializer_ransac, ora.initializer_reconstruct, p1, p2, ora.initializer_sets(12, 200))
    assert not use_h and q["ok"] and q["triangulated"].all()                         # "Initialize succeeded", 12 inliers
    assert np.abs(q["R21"] - np.eye(3)).max() < 0.05 and np.linalg.norm(q["t21"] - fx["t2"]) < 0.15
    T2 = np.c_[q["R21"], q["t21"]]
    T2b, inl, pts, st, cnt = ora.two_view_ba(np.eye(4)[:3], T2, p1, p2, q["triangulated"], q["p3d"])
    scale = np.linalg.norm(T2b[:, 3])
    assert cnt >= 10 and np.linalg.norm(T2b[:, 3] / scale - fx["t2"]) < 0.1 and np.abs(T2b[:, :3] - np.eye(3)).max() < 0.05
    err = np.linalg.norm(pts[inl] / scale - fx["landmarks_F"][inl], axis=1)
    assert np.median(err) < 0.3                                                      # 2 px of noise on 12 points, depths 2 .. 4
    # the planar scene of the same test file: sh / (sh + sf) < 0.4 sends it to ReconstructF, which cannot decide on a plane
    use_h, r, q = try_initialize(ora.initializer_ransac, ora.initializer_reconstruct, fx["px1H"], fx["px2H"], ora.initializer_sets(12, 200))
    assert not use_h and not q["ok"]


# ---- CUDA path -------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_ransac_is_bit_exact(ora):
    from ygz_slam_b200 import Context
    cases = [two_view(0), two_view(1, planar=True), two_view(2, n=57, noise=0.2), two_view(4, n=8, outliers=0.0), two_view(6, n=1500, noise=1.0)]
    offsets = np.r_[0, np.cumsum([len(c[0]) for c in cases])].astype(np.int32)
    iters = 200
    sets = np.stack([ora.initializer_sets(len(c[0]), iters) for c in cases])
    ctx = Context(0)
    g = ctx.initializer_ransac(offsets, np.concatenate([c[0] for c in cases]), np.concatenate([c[1] for c in cases]), sets, models=True)
    for q, c in enumerate(cases):
        o = ora.initializer_ransac(c[0], c[1], sets[q], models=True)
        a, b = offsets[q], offsets[q + 1]
        assert np.array_equal(g["models"][q].view(np.uint64), o["models"].view(np.uint64)), q        # every hypothesis, bit for bit
        assert g["best_H"][q] == o["best_H"] and g["best_F"][q] == o["best_F"]
        assert g["score_H"][q].tobytes() == o["score_H"].tobytes() and g["score_F"][q].tobytes() == o["score_F"].tobytes()
        assert np.array_equal(g["H21"][q], o["H21"]) and np.array_equal(g["F21"][q], o["F21"])
        assert np.array_equal(g["inliers_H"][a:b], o["inliers_H"]) and np.array_equal(g["inliers_F"][a:b], o["inliers_F"])
    ctx.close()


@pytest.mark.gpu
def test_gpu_ransac_rejects_bad_input(ora):
    from ygz_slam_b200 import Context, YgzbError
    ctx = Context(0)
    p = np.random.default_rng(0).uniform(0, 400, (7, 2))
    with pytest.raises(YgzbError):
        ctx.initializer_ransac([0, 7], p, p, np.zeros((1, 3, 8), np.int32))                   # fewer than 8 pairs
    p = np.random.default_rng(0).uniform(0, 400, (20, 2))
    bad = ora.initializer_sets(20, 3)[None].copy()
    bad[0, 1, 2] = 20
    with pytest.raises(YgzbError):
        ctx.initializer_ransac([0, 20], p, p, bad)                                             # set index out of range
    ctx.close()


@pytest.mark.gpu
def test_gpu_reconstruct_is_bit_exact(ora):
    from ygz_slam_b200 import Context
    lists = []
    for seed, planar, use_h in ((0, False, 0), (2, False, 0), (1, True, 1), (0, False, 1), (7, False, 0)):
        p1, p2 = (planar_scene(seed, noise=0.2)[:2] if planar else two_view(seed, n=300 + 37 * seed, noise=0.2)[:2])
        r = ora.initializer_ransac(p1, p2, ora.initializer_sets(len(p1), 200))
        lists.append((p1, p2, use_h, r["H21"] if use_h else r["F21"], r["inliers_H"] if use_h else r["inliers_F"]))
    lists.append(planar_scene(3)[:2] + (1, np.eye(3), np.ones(300, bool)))                       # degenerate H: d1/d2 < 1.00001 -> false
    offsets = np.r_[0, np.cumsum([len(l[0]) for l in lists])].astype(np.int32)
    ctx = Context(0)
    g = ctx.initializer_reconstruct(offsets, np.concatenate([l[0] for l in lists]), np.concatenate([l[1] for l in lists]), [l[2] for l in lists],
                                    np.stack([l[3] for l in lists]), np.concatenate([l[4] for l in lists]))
    n_ok = 0
    for q, l in enumerate(lists):
        o = ora.initializer_reconstruct(*l)
        a, b = offsets[q], offsets[q + 1]
        assert bool(g["ok"][q]) == o["ok"], q
        assert np.array_equal(g["n_good"][q], o["n_good"]), q
        assert np.array_equal(g["candidates"][q].view(np.uint64), o["candidates"].view(np.uint64)), q
        assert np.array_equal(g["R21"][q], o["R21"]) and np.array_equal(g["t21"][q], o["t21"]), q
        assert np.array_equal(g["p3d"][a:b].view(np.uint64), o["p3d"].view(np.uint64)), q
        assert np.array_equal(g["triangulated"][a:b], o["triangulated"]), q
        assert abs(g["parallax"][q] - o["parallax"]) <= 1e-4 * max(1.0, abs(o["parallax"])), q      # acosf vs glibc acosf
        n_ok += int(o["ok"])
    assert n_ok >= 3
    ctx.close()


@pytest.mark.gpu
def test_gpu_initializer_properties_at_full_size():
    """3072 noise-free point pairs (a full grid of features), no oracle in the loop: every pair is an inlier of the recovered F,
    the epipolar residual is at rounding level, ReconstructF returns the true rotation and translation direction, and the
    triangulated structure is the true structure up to the scale |t|."""
    from ygz_slam_b200 import Context
    rng = np.random.default_rng(12)
    n = 3072
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(3, 8, n)]
    th = 0.08
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    t = np.array([0.4, -0.05, 0.1])
    Kf = np.array([[np.float32(520.9), 0, np.float32(325.1)], [0, np.float32(521.0), np.float32(249.7)], [0, 0, 1]], np.float64)

    def proj(Rm, tv):
        Y = (Rm @ X.T).T + tv
        return (Kf @ (Y / Y[:, 2:]).T).T[:, :2]

    p1, p2 = proj(np.eye(3), np.zeros(3)), proj(R, t)
    sets = Oracle().initializer_sets(n, 200)[None]                              # (only the cv::RNG index draw comes from the oracle)
    ctx = Context(0)
    g = ctx.initializer_ransac([0, n], p1, p2, sets)
    assert g["inliers_F"].all() and g["best_F"][0] >= 0
    F = g["F21"][0]
    x1, x2 = np.c_[p1, np.ones(n)], np.c_[p2, np.ones(n)]
    l2 = (F @ x1.T).T
    assert (np.abs((x2 * l2).sum(1)) / np.hypot(l2[:, 0], l2[:, 1])).max() < 1e-6
    q = ctx.initializer_reconstruct([0, n], p1, p2, [0], F[None], g["inliers_F"])
    assert q["ok"][0] and np.abs(q["R21"][0] - R).max() < 1e-7 and np.abs(q["t21"][0] - t / np.linalg.norm(t)).max() < 1e-6
    assert q["triangulated"].all()
    assert np.abs(q["p3d"] * np.linalg.norm(t) - X).max() < 1e-5
    ctx.close()


@pytest.mark.gpu
def test_gpu_reference_test_initializer_scene(ora):
    """test/test_initializer.cpp's two scenes through the device path: every output equals the oracle's bit for bit, and the
    two-view BA that follows in the reference's test lands on the same pose."""
    from ygz_slam_b200 import Context
    fx = np.load(GOLDEN)
    ctx = Context(0)
    sets = ora.initializer_sets(12, 200)
    for tag in ("F", "H"):
        p1, p2 = fx["px1" + tag], fx["px2" + tag]
        use_o, ro, qo = try_initialize(ora.initializer_ransac, ora.initializer_reconstruct, p1, p2, sets)

        def g_ransac(a, b, s_):
            g = ctx.initializer_ransac([0, len(a)], a, b, s_[None])
            return dict(H21=g["H21"][0], F21=g["F21"][0], score_H=g["score_H"][0], score_F=g["score_F"][0], inliers_H=g["inliers_H"], inliers_F=g["inliers_F"])

        def g_recon(a, b, uh, model, inl):
            g = ctx.initializer_reconstruct([0, len(a)], a, b, [int(uh)], model[None], inl)
            return dict(ok=bool(g["ok"][0]), R21=g["R21"][0], t21=g["t21"][0], p3d=g["p3d"], triangulated=g["triangulated"], n_good=g["n_good"][0])

        use_g, rg, qg = try_initialize(g_ransac, g_recon, p1, p2, sets)
        assert use_g == use_o and qg["ok"] == qo["ok"]
        assert np.array_equal(rg["F21"], ro["F21"]) and np.array_equal(rg["H21"], ro["H21"])
        assert rg["score_F"].tobytes() == ro["score_F"].tobytes() and rg["score_H"].tobytes() == ro["score_H"].tobytes()
        assert np.array_equal(qg["R21"], qo["R21"]) and np.array_equal(qg["t21"], qo["t21"]) and np.array_equal(qg["p3d"], qo["p3d"])
        assert np.array_equal(qg["n_good"], qo["n_good"]) and np.array_equal(qg["triangulated"], qo["triangulated"])
        if qo["ok"]:
            T2 = np.c_[qo["R21"], qo["t21"]]
            wT, winl, wpts, _, _ = ora.two_view_ba(np.eye(4)[:3], T2, p1, p2, qo["triangulated"], qo["p3d"])
            gT, ginl, gpts, _ = ctx.two_view_ba([0, 12], np.eye(4)[:3], T2, p1, p2, qg["triangulated"], qg["p3d"])
            assert np.abs(gT[0] - wT).max() < 1e-6 and np.array_equal(ginl, winl) and np.abs(gpts - wpts).max() < 1e-6
    ctx.close()
