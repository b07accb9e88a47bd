#!/usr/bin/env python3
"""One resident step of the C2 pipeline (no torch), for ncu captures:
   ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/prof python tools/profile_step.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ygz_slam_b200 import Context  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = Context(0, n_levels=8)
fr = ctx.frames(B)
fr.upload(bench.make_frames(B, 0))
slots = np.arange(B, dtype=np.int32)
for _ in range(reps):
    fr.build_pyramid(0, B)
    fr.detect(slots, fetch=False)
    fr.match(slots, (slots + 1) % B, True, fetch=False)
ctx.synchronize()
print("done", ctx.launch_count, "launches")
