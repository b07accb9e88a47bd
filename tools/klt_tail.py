#!/usr/bin/env python3
"""Error tail of the GPU KLT (ygzb_klt) against the oracle and cv2 on the synthetic frames: prints the distribution and the
worst points with their conditioning (run on the GPU box; tests/test_gpu_klt.py asserts the bounds it reports)."""
import sys
from pathlib import Path

import cv2
import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.pyoracle import Oracle  # noqa: E402
from ygz_slam_b200 import Context, synth  # noqa: E402

o = Oracle()
ctx = Context(0)
g1, g2, g3 = (synth.stream_frame(k)[0] for k in (1, 2, 3))
fr = ctx.frames(3)
fr.upload(np.stack([g1, g2, g3]))
f = o.detect(o.build_pyramid(g1, 3))
ref = np.stack([f["px"], f["py"]], 1).astype(np.float32)
n = len(ref)
init = (ref + np.float32(1.5)).astype(np.float32)
lam = cv2.cornerMinEigenVal(g1, 21, ksize=3)[ref[:, 1].astype(int), ref[:, 0].astype(int)]
got, gst, gerr = fr.klt([0, 0], [2, 1], [0, n, 2 * n], np.concatenate([ref, ref]), np.concatenate([init, init]))
for p, cur in enumerate((g3, g2)):
    s = slice(p * n, (p + 1) * n)
    want, wst, werr = o.klt(g1, cur, ref, init)
    both = wst.astype(bool) & gst[s].astype(bool)
    d = np.abs(got[s] - want).max(1)
    print(f"pair {p}: n {n} both {both.sum()} status differs {(wst.astype(bool) != gst[s].astype(bool)).sum()} "
          f"median {np.median(d[both]):.2e} p99 {np.percentile(d[both], 99):.2e} p99.9 {np.percentile(d[both], 99.9):.2e} max {d[both].max():.3e}")
    good = both & (lam > np.percentile(lam, 20))
    print(f"   well-conditioned (min-eig above the 20th percentile): n {good.sum()} max {d[good].max():.3e}")
    for i in np.argsort(-d * both)[:8]:
        print(f"   d={d[i]:.4f} px  min-eig={lam[i]:.5f}  err gpu/oracle={gerr[s][i]:.2f}/{werr[i]:.2f}  at ({ref[i, 0]:.0f},{ref[i, 1]:.0f})")
