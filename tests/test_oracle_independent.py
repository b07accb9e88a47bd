"""Independent numpy re-derivations of the reference-owned float / integer kernels, written from the reference TEXT
(src/Algorithm/FeatureDetector.cpp:467-578, src/Algorithm/CVUtils.cpp:186-318) and sharing no code with oracle/*.cpp:
a misreading that the oracle and the CUDA kernels have in common (they were written by the same author) would show up
here.  Float steps follow the C++ types statement by statement (float32 locals, double literals promote)."""
from pathlib import Path

import numpy as np
import pytest

from ygz_slam_b200 import synth

f32 = np.float32
ROOT = Path(__file__).resolve().parent.parent


def _pattern():
    """bit_pattern_31_ (FeatureDetector.cpp:37-295) as (256, 2, 2): pairs of (x, y) offsets."""
    txt = "\n".join(l for l in (ROOT / "ygz_slam_b200" / "csrc" / "orb_pattern.inc").read_text().splitlines() if not l.startswith("//"))
    vals = [int(t) for t in txt.replace("{", " ").replace("}", " ").replace(",", " ").split() if t.lstrip("-").isdigit()]
    assert len(vals) == 1024
    return np.array(vals, np.int64).reshape(512, 2)


def shi_tomasi(img, u, v):
    """FeatureDetector::ShiTomasiScore (:467-507)."""
    h, w = img.shape
    x_min, x_max, y_min, y_max = u - 4, u + 4, v - 4, v + 4
    if x_min < 1 or x_max >= w - 1 or y_min < 1 or y_max >= h - 1:
        return f32(0.0)
    dXX = dYY = dXY = f32(0.0)
    I = img.astype(np.int32)
    for y in range(y_min, y_max):
        for x in range(x_min, x_max):
            dx = f32(I[y, x + 1] - I[y, x - 1])
            dy = f32(I[y + 1, x] - I[y - 1, x])
            dXX = f32(dXX + dx * dx)
            dYY = f32(dYY + dy * dy)
            dXY = f32(dXY + dx * dy)
    dXX = f32(np.float64(dXX) / 128.0)      # dXX / (2.0 * box_area): double division, stored to float
    dYY = f32(np.float64(dYY) / 128.0)
    dXY = f32(np.float64(dXY) / 128.0)
    s = f32(dXX + dYY)
    disc = f32(f32(s * s) - f32(f32(4) * f32(f32(dXX * dYY) - f32(dXY * dXY))))
    # `sqrt` of a float under `using namespace std` (Common.h:17) is the f32 overload; 0.5 is a double literal
    return f32(0.5 * np.float64(f32(s - np.sqrt(disc))))


def fast_atan2(y, x):
    """cv::fastAtan2 (OpenCV), float32, degrees."""
    # static const float atan2_p1 = 0.9997878412794807f * (float)(180 / CV_PI): a float x float product
    p1, p3, p5, p7 = (f32(f32(c) * f32(57.29577951308232)) for c in (0.9997878412794807, -0.3258083974640975, 0.1555786518463281, -0.04432655554792128))
    ax, ay = f32(abs(x)), f32(abs(y))
    eps = f32(2.220446049250313e-16)
    if ax >= ay:
        c = f32(ay / f32(ax + eps))
        c2 = f32(c * c)
        a = f32(f32(f32(f32(f32(f32(p7 * c2) + p5) * c2) + p3) * c2 + p1) * c)
    else:
        c = f32(ax / f32(ay + eps))
        c2 = f32(c * c)
        a = f32(f32(90.0) - f32(f32(f32(f32(f32(f32(p7 * c2) + p5) * c2) + p3) * c2 + p1) * c))
    if x < 0:
        a = f32(f32(180.0) - a)
    if y < 0:
        a = f32(f32(360.0) - a)
    return a


UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]   # ORB-SLAM's table (the reference's own init reads past a vector)


def ic_angle(img, px, py):
    """FeatureDetector::IC_Angle (:509-537) at cvRound(pt) of the level image."""
    cx, cy = int(np.rint(px)), int(np.rint(py))
    I = img.astype(np.int64)
    m01 = m10 = 0
    for u in range(-15, 16):
        m10 += u * I[cy, cx + u]
    for v in range(1, 16):
        d = UMAX[v]
        v_sum = 0
        for u in range(-d, d + 1):
            plus, minus = I[cy + v, cx + u], I[cy - v, cx + u]
            v_sum += plus - minus
            m10 += u * (plus + minus)
        m01 += v * v_sum
    return fast_atan2(f32(m01), f32(m10))


def orb_descriptor(img, px, py, angle_deg, pattern):
    """FeatureDetector::ComputeOrbDescriptor (:539-578) on the level image, pixel already divided by 2^level."""
    ang = f32(f32(angle_deg) * f32(np.pi / 180.0))
    # (float)cos(angle): the double cosine of the float angle, rounded to float
    a, b = f32(np.cos(np.float64(ang))), f32(np.sin(np.float64(ang)))
    cx, cy = int(np.rint(px)), int(np.rint(py))
    out = np.zeros(32, np.uint8)
    for i in range(32):
        val = 0
        for k in range(8):
            t = []
            for idx in (16 * i + 2 * k, 16 * i + 2 * k + 1):
                x, y = f32(pattern[idx, 0]), f32(pattern[idx, 1])
                r = int(np.rint(f32(f32(x * b) + f32(y * a))))      # cvRound: round half to even, like np.rint
                c = int(np.rint(f32(f32(x * a) - f32(y * b))))
                t.append(int(img[cy + r, cx + c]))
            val |= int(t[0] < t[1]) << k
        out[i] = val
    return out


def align2d(img, ref_border, ref, n_iter, u0, v0):
    """cvutils::Align2D (CVUtils.cpp:186-318): returns (ok, u, v)."""
    h, w = img.shape
    rb = ref_border.astype(np.int32).reshape(10, 10)
    dx, dy = np.zeros(64, f32), np.zeros(64, f32)
    H = np.zeros((3, 3), f32)
    k = 0
    for y in range(8):
        for x in range(8):
            J = np.array([f32(0.5 * (rb[y + 1, x + 2] - rb[y + 1, x])), f32(0.5 * (rb[y + 2, x + 1] - rb[y, x + 1])), f32(1)], f32)
            dx[k], dy[k] = J[0], J[1]
            H = (H + np.outer(J, J).astype(f32)).astype(f32)
            k += 1
    Hinv = np.linalg.inv(H.astype(np.float64))     # Eigen's closed-form float inverse: compared with a tolerance below
    mean_diff = f32(0)
    u, v = f32(u0), f32(v0)
    chi2 = f32(0)
    converged = False
    I = img.astype(np.float32)
    rf = ref.astype(np.float32).reshape(-1)
    for _ in range(n_iter):
        chi2 = f32(0)
        u_r, v_r = int(np.floor(u)), int(np.floor(v))
        if u_r < 4 or v_r < 4 or u_r >= w - 4 or v_r >= h - 4:
            break
        sx, sy = f32(u - f32(u_r)), f32(v - f32(v_r))
        wTL = f32((1.0 - np.float64(sx)) * (1.0 - np.float64(sy)))
        wTR = f32(np.float64(sx) * (1.0 - np.float64(sy)))
        wBL = f32((1.0 - np.float64(sx)) * np.float64(sy))
        wBR = f32(sx * sy)
        Jres = np.zeros(3, f32)
        k = 0
        for y in range(8):
            for x in range(8):
                yy, xx = v_r + y - 4, u_r + x - 4
                sp = f32(f32(f32(f32(wTL * I[yy, xx]) + f32(wTR * I[yy, xx + 1])) + f32(wBL * I[yy + 1, xx])) + f32(wBR * I[yy + 1, xx + 1]))
                res = f32(f32(sp - rf[k]) + mean_diff)
                Jres[0] = f32(Jres[0] - f32(res * dx[k]))
                Jres[1] = f32(Jres[1] - f32(res * dy[k]))
                Jres[2] = f32(Jres[2] - res)
                chi2 = f32(chi2 + f32(res * res))
                k += 1
        upd = (Hinv @ Jres.astype(np.float64)).astype(f32)
        u, v, mean_diff = f32(u + upd[0]), f32(v + upd[1]), f32(mean_diff + upd[2])
        if f32(f32(upd[0] * upd[0]) + f32(upd[1] * upd[1])) < f32(0.03 * 0.03):
            converged = True
            break
    return bool(converged and chi2 < 20000), float(u), float(v)


@pytest.fixture(scope="module")
def scene(oracle):
    g = synth.stream_frame(2)[0]
    pyr = oracle.build_pyramid(g, 3)
    feats = oracle.detect(pyr)
    return g, pyr, feats


def test_shi_tomasi_score_rederived(oracle, scene):
    g, pyr, feats = scene
    rng = np.random.default_rng(3)
    sel = rng.choice(np.nonzero(feats["level"] == 0)[0], 40, replace=False)
    for i in sel:
        u, v = int(feats["px"][i]), int(feats["py"][i])
        want = shi_tomasi(g, u, v)
        got = f32(oracle.shi_tomasi(g, u, v))
        assert got == want and got == feats["score"][i], (u, v, got, want)
    assert shi_tomasi(g, 3, 200) == 0 and oracle.shi_tomasi(g, 3, 200) == 0   # border rule


def test_ic_angle_and_descriptor_rederived(oracle, scene):
    g, pyr, feats = scene
    pat = _pattern()
    rng = np.random.default_rng(4)
    for L in (0, 1):
        img = oracle.level_view(pyr, 640, 480, 3, L)
        cand = np.nonzero(feats["level"] == L)[0]
        # keep clear of the image end: the reference's unguarded taps there are a documented parity decision of its own
        cand = [i for i in cand if 40 <= feats["py"][i] / (1 << L) < img.shape[0] - 40 and 40 <= feats["px"][i] / (1 << L) < img.shape[1] - 40]
        for i in rng.choice(cand, 12, replace=False):
            px, py = feats["px"][i] / (1 << L), feats["py"][i] / (1 << L)
            ang = ic_angle(img, px, py)
            assert ang == feats["angle"][i], (i, ang, feats["angle"][i])
            d = orb_descriptor(img, px, py, feats["angle"][i], pat)
            assert np.array_equal(d, feats["desc"][i]), i


def test_align2d_rederived(oracle):
    g1 = synth.stream_frame(1)[0]
    g2 = synth.stream_frame(2)[0]
    rng = np.random.default_rng(6)
    n_ok = 0
    for _ in range(25):
        x, y = int(rng.integers(40, 600)), int(rng.integers(40, 440))
        rb = g1[y - 5:y + 5, x - 5:x + 5].copy()
        ref = rb[1:9, 1:9].copy()
        u0, v0 = x + rng.uniform(-1.5, 1.5), y + rng.uniform(-1.5, 1.5)
        ok, u, v = align2d(g2, rb, ref, 10, u0, v0)
        got_ok, gu, gv = oracle.align2d(g2, rb, ref, u0, v0, 10)
        # the only shared-misreading-proof difference: numpy's generic inverse vs Eigen's cofactor float inverse
        assert got_ok == ok and abs(gu - u) < 2e-3 and abs(gv - v) < 2e-3, (x, y, ok, got_ok, u, gu, v, gv)
        n_ok += ok
    assert n_ok >= 10
