"""The tracking loop of BASELINE config C5 (ygz_slam_b200/vo.py): on the CPU oracle it follows the synthetic
ground truth; on the GPU it reproduces the oracle's trajectory (end-to-end pose parity through sparse alignment,
direct projection, pose-only refinement, detection at keyframes and local BA)."""
import numpy as np
import pytest

from ygz_slam_b200 import se3, synth, vo
from oracle.vo_backend import OracleBackend


def _run(backend, n_streams, n_frames, step=2):
    V = vo.VisualOdometry(backend, n_streams, kf_min_frames=5, kf_min_rot=0.03, kf_min_trans=0.03)
    T0 = [None] * n_streams
    errs = np.zeros((n_streams, n_frames))
    for k in range(n_frames):
        frames = [synth.stream_frame(step * k, stream=s) for s in range(n_streams)]
        V.add_frames([f[0] for f in frames], [f[1] for f in frames], k)
        for s in range(n_streams):
            if T0[s] is None:
                T0[s] = frames[s][2]
            gt = se3.mul(frames[s][2], se3.inv(T0[s]))
            errs[s, k] = np.linalg.norm(se3.se3_log(se3.mul(V.streams[s].T_cw, se3.inv(gt))))
    return V, errs


def test_vo_on_oracle_follows_ground_truth(oracle):
    V, errs = _run(OracleBackend(oracle), 1, 24)
    st = V.streams[0]
    assert not st.lost
    assert st.stats["keyframes"] >= 3 and st.stats["ba"] >= 2
    assert errs.max() < 3e-3                       # metres / radians on a 2 m scene, pixel noise sigma = 2 grey levels
    assert st.stats["projected"] > 0.9 * st.stats["candidates"]


@pytest.mark.gpu
def test_vo_gpu_matches_oracle_trajectory(ctx3, oracle):
    n_streams, n_frames = 2, 20
    Vo, _ = _run(OracleBackend(oracle), n_streams, n_frames)
    be = vo.GpuBackend(ctx3, n_streams * vo.VisualOdometry.SLOTS_PER_STREAM)
    Vg, errs = _run(be, n_streams, n_frames)
    for s in range(n_streams):
        assert not Vg.streams[s].lost
        sg, so = Vg.streams[s].stats, Vo.streams[s].stats
        assert sg["frames"] == so["frames"] and sg["keyframes"] == so["keyframes"] and sg["ba"] == so["ba"]
        # the alignment pose agrees to ~1e-9, so a borderline patch may converge on one side only: allow 0.1 %
        for key in ("candidates", "projected", "inliers"):
            assert abs(sg[key] - so[key]) <= 1e-3 * so[key], key
        for Tg, Tw in zip(Vg.streams[s].trajectory, Vo.streams[s].trajectory):
            assert np.linalg.norm(se3.se3_log(se3.mul(Tg, se3.inv(Tw)))) < 1e-4   # BASELINE: pose error < 1e-4 vs reference
    assert errs.max() < 3e-3
    be.fr.close()


@pytest.mark.gpu
def test_native_driver_matches_python_loop(ctx3):
    """host/vo_driver.cpp against vo.VisualOdometry on the same GPU backend -- the per-stage C++ loop (one blocking C-ABI call
    per stage) and the device-resident engine (ygzb_tracker_*: local map, candidate projection, key-frame insertion and BA
    assembly on the device; one and several frames per stream in flight): same key-frames and BAs, every pose within the
    stated 1e-4 (the loops differ in the rounding of SE3 products and in the summation order of the BA, which can flip a
    borderline candidate at the 20 px border or a pose-only inlier on the threshold; measured 1.5e-5)."""
    from ygz_slam_b200 import vo_native
    n_streams, n_frames = 3, 26
    data = [synth.shift_stream(s, n_frames) for s in range(n_streams)]
    be = vo.GpuBackend(ctx3, n_streams * vo.VisualOdometry.SLOTS_PER_STREAM)
    V = vo.VisualOdometry(be, n_streams, kf_min_frames=5, kf_min_rot=0.03, kf_min_trans=0.03)
    for k in range(n_frames):
        V.add_frames([data[s][0][k] for s in range(n_streams)], [data[s][1] for s in range(n_streams)], k)
    be.fr.close()
    runs = {"stages": vo_native.run(ctx3, [d[0] for d in data], [d[1] for d in data], 5, 0.03, 0.03, engine="stages"),
            "resident_w1": vo_native.run(ctx3, [d[0] for d in data], [d[1] for d in data], 5, 0.03, 0.03, window=1),
            "resident_w5_2threads": vo_native.run(ctx3, [d[0] for d in data], [d[1] for d in data], 5, 0.03, 0.03, window=8, threads=2)}
    for name, (traj, stats, sec) in runs.items():
        _check_native(name, traj, stats, sec, V, data, n_streams, n_frames)
    # a window only changes how many frames are in flight, never a result: bit-identical trajectories
    assert np.array_equal(runs["resident_w1"][0], runs["resident_w5_2threads"][0])


def _check_native(name, traj, stats, sec, V, data, n_streams, n_frames):
    assert sec > 0, name
    for s in range(n_streams):
        st = V.streams[s]
        assert not st.lost and stats[s]["lost"] == 0, name
        assert stats[s]["keyframes"] == st.stats["keyframes"] and stats[s]["ba"] == st.stats["ba"], name
        assert stats[s]["keyframes"] >= 3 and stats[s]["ba"] >= 2, name
        for key in ("candidates", "projected", "inliers"):
            assert abs(stats[s][key] - st.stats[key]) <= 1e-3 * st.stats[key], (name, key)
        for k in range(n_frames):
            assert np.linalg.norm(se3.se3_log(se3.mul(traj[s, k], se3.inv(st.trajectory[k])))) < 1e-4, (name, s, k)
        # and both follow the exact ground truth of the sliding-crop stream
        assert np.linalg.norm(se3.se3_log(se3.mul(traj[s, -1], se3.inv(data[s][2][-1])))) < 3e-3
