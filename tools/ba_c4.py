#!/usr/bin/env python3
"""BASELINE config C4 (local BA, 10 key-frames x 2000 landmarks x ~8000 observations) alone: kernel time per LM trial, parity
against the oracle, and -- with YGZB_BA_DEBUG=1 -- the per-phase cycle counts the kernel collects (run on the GPU box)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ygz_slam_b200 import Context, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = Context(0)
sc = synth.ba_scene()
g2o = np.concatenate([sc["poses_noisy"][:, 3:], sc["poses_noisy"][:, :3]], 1)
fixed = np.zeros(10, np.uint8)
fixed[0] = 1
n_obs = len(sc["kf_idx"])
args = ([0, 10], [0, 2000], [0, n_obs], g2o, fixed, sc["pts_noisy"], sc["kf_idx"], sc["pt_idx"], sc["px"])
ctx.local_ba(*args)
ctx.profile(True)
t0 = time.perf_counter()
for _ in range(reps):
    P, X, outl, st = ctx.local_ba(*args)
t = (time.perf_counter() - t0) / reps
prof = ctx.profile_read()
k_ms = prof["local_ba"][0] / max(prof["local_ba"][1], 1)
print(f"C4: {n_obs} observations, iters {st[0]['iters']} trials {st[0]['lm_trials']} chi2 {st[0]['chi2_final']:.6f} "
      f"kernel {k_ms:.3f} ms = {k_ms / st[0]['lm_trials']:.4f} ms/trial, call {t * 1e3:.3f} ms")
if "--oracle" in sys.argv:
    from oracle.pyoracle import Oracle
    wP, wX, wout, wst = Oracle().local_ba(g2o, fixed, sc["pts_noisy"], sc["kf_idx"], sc["pt_idx"], sc["px"])
    print(f"    oracle: iters {wst['iters']} trials {wst['lm_trials']} chi2 {wst['chi2_final']:.6f} max landmark diff {np.abs(X - wX).max():.3e} "
          f"max pose diff {np.abs(P - wP).max():.3e}")
ctx.close()
