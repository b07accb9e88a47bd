#!/usr/bin/env python3
"""Runs C3 sparse alignment and C4 local BA once (for ncu captures) and prints an e2e breakdown of the C2 step."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ygz_slam_b200 import Context, synth, se3  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ba"
ctx = Context(0, n_levels=8)
if mode == "ba":
    sc = synth.ba_scene()
    g2o = np.concatenate([sc["poses_noisy"][:, 3:], sc["poses_noisy"][:, :3]], 1)
    fixed = np.zeros(10, np.uint8); fixed[0] = 1
    n_obs = len(sc["kf_idx"])
    for _ in range(2):
        P, X, o, st = ctx.local_ba([0, 10], [0, 2000], [0, n_obs], g2o, fixed, sc["pts_noisy"], sc["kf_idx"], sc["pt_idx"], sc["px"])
    print(st)
elif mode == "sparse":
    from oracle.pyoracle import Oracle
    ora = Oracle()
    g1, d1, T1 = synth.stream_frame(1); g2, _, T2 = synth.stream_frame(4)
    fr = ctx.frames(2); fr.upload(np.stack([g1, g2]))
    f = ora.detect(ora.build_pyramid(g1, 8), n_levels=8)
    rng = np.random.default_rng(7); idx = rng.integers(0, f["n"], 2000)
    px = np.stack([f["px"][idx], f["py"][idx]], 1); depth = d1[px[:, 1].astype(int), px[:, 0].astype(int)]
    for _ in range(2):
        T, n, it = fr.sparse_align([0], [1], [0, 2000], px, depth, np.ones(2000, np.uint8), T1.reshape(1, 12), T1.reshape(1, 12), max_level=3)
    print(n, it)
else:  # e2e breakdown
    import torch
    B = 512
    fr = ctx.frames(B)
    frames = bench.make_frames(B, 0)
    pinned = torch.empty((B, 480, 640), dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = frames
    slots = np.arange(B, dtype=np.int32); nxt = (slots + 1) % B
    for rep in range(3):
        t0 = time.perf_counter(); fr.upload_raw(pinned.data_ptr(), B, 1, 640 * 480); ctx.synchronize()
        t1 = time.perf_counter(); off, _ = fr.detect_packed(slots)
        t2 = time.perf_counter(); q, _, _ = fr.match_packed(slots, nxt, True)
        t3 = time.perf_counter()
        fr.detect(slots, fetch=False); ctx.synchronize(); t4 = time.perf_counter()
        fr.match(slots, nxt, True, fetch=False); ctx.synchronize(); t5 = time.perf_counter()
        print(f"upload+pyr {1e3*(t1-t0):.2f} ms | detect+fetch {1e3*(t2-t1):.2f} | match+fetch {1e3*(t3-t2):.2f} | detect nofetch {1e3*(t4-t3):.2f} | match nofetch {1e3*(t5-t4):.2f}")
    # raw H2D bandwidth of the same buffer, 1-D copy
    d = torch.empty((B, 480, 640), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): d.copy_(pinned, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"torch pinned H2D {B*307200/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms)")
