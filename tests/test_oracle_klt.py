"""CPU suite: the KLT restatement against cv2.calcOpticalFlowPyrLK (the reference's own dependency)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _points(oracle, synth_frames, shift):
    g1 = synth_frames[0][0]
    f = oracle.detect(oracle.build_pyramid(g1, 3))
    ref = np.stack([f["px"], f["py"]], 1).astype(np.float32)
    # also a few points near / outside the border
    extra = np.array([[2.5, 3.0], [637.0, 10.0], [320.0, 478.5], [-30.0, 50.0], [700.0, 500.0]], np.float32)
    ref = np.concatenate([ref, extra])
    return ref, (ref + np.float32(shift)).astype(np.float32)


@pytest.mark.parametrize("shift", [0.0, 1.5, -4.0])
def test_klt_matches_cv2(oracle, synth_frames, shift):
    g1, g2 = synth_frames[0][0], synth_frames[2][0]
    ref, init = _points(oracle, synth_frames, shift)
    want, wst, werr = cv2.calcOpticalFlowPyrLK(g1, g2, ref.copy(), init.copy(), winSize=(21, 21), maxLevel=4,
                                               criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.001),
                                               flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    got, gst, gerr = oracle.klt(g1, g2, ref, init)
    wst = wst.reshape(-1).astype(bool)
    gst = gst.astype(bool)
    assert (wst != gst).mean() < 0.005
    both = wst & gst
    d = np.abs(got[both] - want.reshape(-1, 2)[both]).max(1)
    # OpenCV's SIMD build accumulates the f32 normal equations in a different order: tiny differences
    assert np.percentile(d, 99) < 0.01 and np.median(d) < 1e-3
    assert np.abs(gerr[both] - werr.reshape(-1)[both]).max() < 0.05
    assert both.sum() > 0.9 * len(ref)
