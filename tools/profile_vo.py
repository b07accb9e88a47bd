#!/usr/bin/env python3
"""Where does a lock-step frame of the C5 tracking loop go?  python tools/profile_vo.py [streams] [frames]
Prints wall time per lock-step frame, the library's per-stage kernel time (CUDA events) and a cProfile of the host loop."""
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ygz_slam_b200 import Context, se3, synth, vo  # noqa: E402

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = Context(0)
data = [synth.shift_stream(s, n_frames) for s in range(n_streams)]
be = vo.GpuBackend(ctx, n_streams * vo.VisualOdometry.SLOTS_PER_STREAM)
V = vo.VisualOdometry(be, n_streams, kf_min_frames=5, kf_min_rot=0.03, kf_min_trans=0.03)
warm = 3
pr = cProfile.Profile()
for k in range(n_frames):
    if k == warm:
        ctx.synchronize()
        ctx.profile(True)
        ctx.profile_read()
        t0 = time.perf_counter()
        pr.enable()
    V.add_frames([data[s][0][k] for s in range(n_streams)], [data[s][1] for s in range(n_streams)], k)
ctx.synchronize()
pr.disable()
dt = time.perf_counter() - t0
prof = ctx.profile_read()
nf = n_frames - warm
print(f"{n_streams} streams: {1e3 * dt / nf:.3f} ms per lock-step frame, {n_streams * nf / dt:.1f} tracked frames/s")
tot = 0.0
for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if cnt:
        print(f"  kernel {k:16s} {ms / nf:8.3f} ms/frame  ({cnt / nf:.1f} launches/frame, {ms / cnt:.3f} ms each)")
        tot += ms
print(f"  kernels total {tot / nf:.3f} ms/frame (profiling adds event overhead to the wall time)")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
