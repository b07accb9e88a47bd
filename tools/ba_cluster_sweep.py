#!/usr/bin/env python3
"""Kernel time of the local BA for one cluster size (YGZB_BA_CLUSTER in the environment): the 8 small problems of the
tracking loop and the C4 problem.  for c in 1 2 4 8 16; do YGZB_BA_CLUSTER=$c python tools/ba_cluster_sweep.py; done"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from ygz_slam_b200 import Context, synth  # noqa: E402

c = Context(0)
c.profile(True)
c.profile_read()
r = bench.run_vo(c, 8, 40, threads=1)
p = c.profile_read()
sc = synth.ba_scene()
g2o = np.concatenate([sc["poses_noisy"][:, 3:], sc["poses_noisy"][:, :3]], 1)
fixed = np.zeros(10, np.uint8)
fixed[0] = 1
n = len(sc["kf_idx"])
for _ in range(3):
    P, X, o, st = c.local_ba([0, 10], [0, 2000], [0, n], g2o, fixed, sc["pts_noisy"], sc["kf_idx"], sc["pt_idx"], sc["px"])
q = c.profile_read()
print("cluster", os.environ.get("YGZB_BA_CLUSTER", "auto"), "vo fps", round(r["tracked_frames_per_s"]), "vo ba ms/launch",
      round(p["local_ba"][0] / p["local_ba"][1], 3), "C4 ba ms", round(q["local_ba"][0] / q["local_ba"][1], 3), st[0]["iters"],
      st[0]["lm_trials"])
