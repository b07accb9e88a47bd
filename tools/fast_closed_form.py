"""FAST corner score in closed form for any arc length (numpy, no cv2): shared by tools/make_fast_fixture.py and the tests."""
import numpy as np

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2),
        (-1, 3)]


def score_closed_form(img, xy, arc):
    """max over the 16 arcs of `arc` contiguous ring pixels of min(d) and of min(-d), minus 1 (d = ring pixel - centre)."""
    gi = img.astype(np.int32)
    x, y = xy[:, 0].astype(int), xy[:, 1].astype(int)
    d = np.stack([gi[y + dy, x + dx] - gi[y, x] for dx, dy in RING], 1)
    best = np.full(len(xy), -10 ** 9)
    for s in range(16):
        idx = [(s + k) % 16 for k in range(arc)]
        best = np.maximum(best, np.maximum(d[:, idx].min(1), (-d[:, idx]).min(1)))
    return (best - 1).astype(np.int32)
