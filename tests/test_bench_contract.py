"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys
the driver reads, on the same `config` as the GPU arm would use, and the rank > 0 processes of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--streams", "2", *args],
                       capture_output=True, text=True, env=env, cwd=str(ROOT), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    b = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in b, key
    assert b["impl"] == "reference" and b["metric"] == "tracked frames/sec" and b["unit"] == "frames/s" and b["higher_is_better"] is True
    assert b["value"] > 0 and b["e2e"]["value"] == b["value"] and b["e2e"]["h2d_bytes_per_step"] == 0 and b["e2e"]["d2h_bytes_per_step"] == 0
    cb = b["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == b["value"] and cb["streams_lost"] == 0
    assert cb["one_thread_frames_per_s"] > 0 and abs(sum(cb["stage_share"].values()) - 1.0) < 1e-9
    # the same config dict as the GPU arm builds for this command line (the driver compares the two)
    sys.path.insert(0, str(ROOT))
    import bench
    assert b["config"] == bench.vo_config(2, 10)
    assert "C5" in b["config"]["workload"]


def test_reference_arm_is_silent_on_other_ranks():
    lines = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert lines == []
