"""One short stream through the device-resident tracking engine (`ygz_vo_run`: upload + pyramid, sparse alignment, local-map
projection + patch alignment, pose-only refinement, key-frame insertion, BA assembly, local BA) for
`compute-sanitizer --tool racecheck`:

    compute-sanitizer --tool racecheck --print-limit 20 python tools/racecheck_tracker.py

7 frames = the initial key-frame, 5 tracked frames, a second key-frame with its local BA, and one frame tracked against the
two-key-frame map.  The parity suite's tracking tests are too slow under racecheck.  No oracle here: the run must not lose
the stream and must insert the second key-frame.  RACECHECK_FRAMES / RACECHECK_WINDOW / RACECHECK_STREAMS / RACECHECK_THREADS widen it
(e.g. 2 streams x 16 frames on 2 host threads: three-key-frame BAs, two contexts and their upload streams side by side)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ygz_slam_b200 import Context, synth, vo_native  # noqa: E402


def main():
    n = int(os.environ.get("RACECHECK_FRAMES", "7"))
    window = int(os.environ.get("RACECHECK_WINDOW", "4"))
    streams = int(os.environ.get("RACECHECK_STREAMS", "1"))
    threads = int(os.environ.get("RACECHECK_THREADS", "1"))
    data = [synth.shift_stream(s, n)[:2] for s in range(streams)]
    ctx = Context(0)
    t0 = time.time()
    traj, stats, _ = vo_native.run(ctx, [d[0] for d in data], [d[1] for d in data], 5, 0.03, 0.03, window=window, threads=threads)[:3]
    for st in stats:
        print(f"tracker: {n} frames, window {window}, {streams} stream(s) on {threads} host thread(s): {st} t={time.time() - t0:.1f}s", flush=True)
        assert not st["lost"] and st["keyframes"] >= 2 and st["ba"] >= 1
    assert np.isfinite(traj).all()
    ctx.close()
    print("racecheck_tracker: done")


if __name__ == "__main__":
    main()
