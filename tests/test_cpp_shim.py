"""The C++ shim (reference call surface over the C ABI) builds with g++ against libygz_b200.so (CPU suite) and
reproduces the oracle's numbers when driven like test/test_orb_match.cpp (GPU suite)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "tests" / "_build" / "cpp_shim_test"


def build_shim_test():
    from ygz_slam_b200 import capi
    capi.load_library()
    EXE.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", str(EXE), str(ROOT / "tests" / "cpp_shim_test.cpp"),
           f"-L{ROOT / 'ygz_slam_b200'}", "-lygz_b200", f"-Wl,-rpath,{ROOT / 'ygz_slam_b200'}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_shim_compiles_and_links_without_cuda_headers():
    """Host C++ only needs include/ygz_b200.h + the shim header: no CUDA toolkit on the include path."""
    build_shim_test()
    assert EXE.exists()
    r = subprocess.run([str(EXE)], capture_output=True)
    assert r.returncode == 2  # usage error, i.e. the binary starts and the shared library resolves


@pytest.mark.gpu
def test_shim_reproduces_oracle(oracle, tmp_path):
    from ygz_slam_b200 import se3, synth
    build_shim_test()
    g1, d1, T1 = synth.stream_frame(1)
    g2, _, T2 = synth.stream_frame(4)
    Trel = se3.mul(T2, se3.inv(T1))
    blob = tmp_path / "in.bin"
    with open(blob, "wb") as f:
        f.write(g1.tobytes()); f.write(g2.tobytes()); f.write(d1.astype(np.float32).tobytes()); f.write(Trel.astype(np.float64).tobytes())
    voc_file = tmp_path / "voc.bin"
    voc_data = synth.make_vocabulary(k=8, L=5, seed=21)
    voc_file.write_bytes(voc_data)
    r = subprocess.run([str(EXE), str(blob), str(voc_file)], capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    f1 = oracle.detect(oracle.build_pyramid(g1, 3))
    f2 = oracle.detect(oracle.build_pyramid(g2, 3))
    idx, dist = oracle.match_bf(f1["desc"], f2["desc"], True)
    want = f"features {f1['n']} {f2['n']} matches {(idx >= 0).sum()} dist_sum {dist[idx >= 0].sum()}"
    assert lines[0] == want
    ok, err = lines[1].split()[2], float(lines[1].split()[-1])
    assert ok == "1" and err < 2e-3
    assert int(lines[2].split()[-1]) > 0.8 * len(range(0, f1["n"], 4))
    # Tracker (KLT) against the oracle: same survivors (status && InFrame(pt, 20)) and mean disparity
    ref = np.stack([f1["px"], f1["py"]], 1).astype(np.float32)
    cur, st, _ = oracle.klt(g1, g2, ref, ref.copy())
    inside = (cur[:, 0] >= 20) & (cur[:, 0] < 620) & (cur[:, 1] >= 20) & (cur[:, 1] < 460)
    keep = st.astype(bool) & inside
    parts = lines[3].split()
    assert parts[2] == "1" and abs(int(parts[4]) - keep.sum()) <= 3
    want_disp = np.linalg.norm(ref[keep] - cur[keep], axis=1).mean()
    assert abs(float(parts[6]) - want_disp) < 0.02
    # both bundle-adjustment flavours (ba::LocalBAG2O, ba::LocalBA) on the test_local_ba.cpp fixture: noise-free
    # observations, so the reprojection error must collapse; key-frame 0 keeps its pose
    for line, name in ((lines[4], "g2o"), (lines[7], "ceres")):
        parts = line.split()
        assert parts[1] == name and float(parts[3]) > 1.0 and float(parts[5]) < 1e-3 and parts[7] == "0"
    # ba::OptimizeCurrent (pose + points, Huber 0.1) and ba::OptimizeCurrentPointOnly after perturbing the converged map
    oc, op = lines[5].split(), lines[6].split()
    assert oc[0] == "optimize_current" and float(oc[2]) > 0.5 and float(oc[4]) < 1e-3 and oc[6] == "0"
    assert op[0] == "optimize_point_only" and float(op[2]) > 0.5 and float(op[4]) < 1e-3
    # slot ownership (a live key-frame keeps its pyramid; exhaustion is reported, not silently recycled) and the 8b helpers
    sl = lines[8].split()
    assert sl[0] == "slots" and sl[4] == "1" and sl[6] == "1" and sl[8] == "1" and sl[10] == "1"
    assert abs(float(sl[12]) - (-(1.0 + 0.1 * 0.1 / 4.0))) < 1e-9      # JacobXYZ2Cam(0.1, -0.2, 2)(0, 4) = -(1 + x^2/z^2)
    # DBoW3 front (ORBVocabulary::loadFromBinaryFile, Frame::ComputeBoW, Matcher::SearchByBoW with the shim's th_low = 65)
    v = oracle.vocab_load(voc_data)
    _, n1, _, bw1, bv1 = oracle.bow_transform(v, f1["desc"], 4)
    _, n2, _, _, _ = oracle.bow_transform(v, f2["desc"], 4)
    om, ocnt = oracle.search_by_bow(f1["desc"], n1, f1["angle"], f2["desc"], n2, f2["angle"], th_low=65, knn_ratio=0.9, check_orientation=False)
    h = 0
    for i in np.flatnonzero(om >= 0):
        h = (h * 31 + int(i) * 7 + int(om[i])) % 1000003
    bl = lines[9].split()
    assert bl[0] == "bow" and int(bl[2]) == oracle.vocab_info(v)["words"] and int(bl[4]) == len(bw1)
    assert int(bl[6]) == len(set(n1[n1 >= 0])) and int(bl[8]) == int((n1 >= 0).sum())
    assert abs(float(bl[10]) - bv1.sum()) < 1e-10 and int(bl[12]) == ocnt and int(bl[14]) == int((om >= 0).sum()) and int(bl[16]) == h
    oracle.vocab_free(v)
    # Initializer::FindModels on the same matches: scores, inlier counts and F21(2,2) are the oracle's, bit for bit
    m1 = np.stack([f1["px"], f1["py"]], 1)[idx >= 0]
    m2 = np.stack([f2["px"], f2["py"]], 1)[idx[idx >= 0]]
    ro = oracle.initializer_ransac(m1, m2, oracle.initializer_sets(len(m1), 200))
    il = lines[10].split()
    assert il[0] == "initializer" and int(il[2]) == len(m1)
    assert il[6] == "%.3f" % ro["score_H"] and il[8] == "%.3f" % ro["score_F"]
    assert int(il[10]) == int(ro["inliers_H"].sum()) and int(il[12]) == int(ro["inliers_F"].sum())
    assert il[14] == "%.9e" % ro["F21"][2, 2]
    assert int(il[4]) == int(ro["score_H"] / (ro["score_H"] + ro["score_F"]) > 0.4)
    # Initializer::TryInitialize end to end (model choice + ReconstructF / ReconstructH)
    use_h = bool(ro["score_H"] / (ro["score_H"] + ro["score_F"]) > 0.4)
    rq = oracle.initializer_reconstruct(m1, m2, use_h, ro["H21"] if use_h else ro["F21"], ro["inliers_H"] if use_h else ro["inliers_F"])
    tl = lines[11].split()
    assert tl[0] == "try_initialize" and int(tl[2]) == int(rq["ok"])
    assert int(tl[4]) == (int(rq["triangulated"].sum()) if rq["ok"] else 0)
    if rq["ok"]:
        # quaternion round trip inside the shim's SE3: compare with a tolerance
        assert np.allclose([float(tl[6]), float(tl[7]), float(tl[8])], rq["t21"], atol=1e-9) and abs(float(tl[10]) - rq["R21"][0, 0]) < 1e-9
