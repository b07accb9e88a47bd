"""oracle/vo_cpu.cpp (the C++ tracking loop on the CPU oracle that bench.py times as the reference arm of BASELINE
config C5) against the Python loop on the same oracle (ygz_slam_b200/vo.py + oracle/vo_backend.py): same key-frames,
same local BAs, same trajectory -- so the timed CPU arm and the parity loop are the same computation."""
import numpy as np

from oracle import pyoracle
from oracle.vo_backend import OracleBackend
from ygz_slam_b200 import se3, synth, vo


def test_cpp_cpu_loop_matches_python_oracle_loop(oracle):
    n_streams, n_frames = 2, 13
    data = [synth.shift_stream(s, n_frames) for s in range(n_streams)]
    V = vo.VisualOdometry(OracleBackend(oracle), n_streams, kf_min_frames=5, kf_min_rot=0.03, kf_min_trans=0.03)
    for k in range(n_frames):
        V.add_frames([data[s][0][k] for s in range(n_streams)], [data[s][1] for s in range(n_streams)], k)
    traj, stats, sec, stage = pyoracle.vo_run(oracle, [d[0] for d in data], [d[1] for d in data], 5, 0.03, 0.03, warm=2, threads=2)
    assert sec > 0 and stage["sparse_align"] > 0 and stage["local_ba"] > 0
    for s in range(n_streams):
        st = V.streams[s]
        assert not st.lost and stats[s]["lost"] == 0
        assert stats[s]["keyframes"] == st.stats["keyframes"] >= 3 and stats[s]["ba"] == st.stats["ba"] >= 2
        for key in ("candidates", "projected", "inliers"):
            assert abs(stats[s][key] - st.stats[key]) <= 1e-3 * st.stats[key], key
        for k in range(n_frames):
            assert np.linalg.norm(se3.se3_log(se3.mul(traj[s, k], se3.inv(st.trajectory[k])))) < 1e-4, (s, k)   # host-side SE3 rounding can flip a borderline candidate / inlier (measured 2e-6)
        assert np.linalg.norm(se3.se3_log(se3.mul(traj[s, -1], se3.inv(data[s][2][-1])))) < 3e-3
