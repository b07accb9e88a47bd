"""CPU suite: the C-ABI library builds, loads and exports every symbol include/ygz_b200.h declares.
No compute calls (no GPU here); creating a context must fail loudly, not fall back to a CPU path."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ygz_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ygzb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ygz_slam_b200 import capi
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == syms


def test_no_cpu_fallback():
    import torch
    from ygz_slam_b200 import Context, YgzbError
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu suite")
    with pytest.raises(YgzbError):
        Context(0)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/ (checker, never shipped)."""
    for p in (ROOT / "ygz_slam_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp"}:
            txt = p.read_text(errors="ignore")
            assert "pyoracle" not in txt and "liboracle" not in txt and "oracle.h" not in txt, p


def test_header_is_plain_c_and_links(tmp_path):
    """include/ygz_b200.h must be usable from C (no C++ / CUDA / torch types): compile a pedantic C99 translation unit
    that takes the address of every declared entry point and link it against the library (no call is made: no GPU)."""
    import subprocess
    syms = declared_symbols()
    src = tmp_path / "abi.c"
    body = "\n".join(f"    p[{i}] = (fn)&{s_};" for i, s_ in enumerate(syms))
    src.write_text('#include "ygz_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n    fn p[%d];\n%s\n'
                   '    printf("%%d\\n", (int)(sizeof p / sizeof p[0]));\n    return p[0] == 0;\n}\n' % (len(syms), body))
    exe = tmp_path / "abi"
    libdir = ROOT / "ygz_slam_b200"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{libdir}",
                    "-lygz_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert int(out) == len(syms) and len(syms) >= 35


def _build_example(tmp_path):
    import subprocess
    exe = tmp_path / "extract_match"
    libdir = ROOT / "ygz_slam_b200"
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "extract_match.c"),
                    f"-L{libdir}", "-lygz_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True, capture_output=True, text=True)
    return exe


def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    """examples/extract_match.c (the test_orb_match shape in plain C) builds against the header + library; without an
    sm_100 device it must stop at ygzb_create with a message, not fall back to anything."""
    import subprocess
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked run of the example")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and "ygzb_create" in r.stderr and "sm_100" in r.stderr


@pytest.mark.gpu
def test_c_example_runs(tmp_path):
    import subprocess
    exe = _build_example(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, check=True)
    words = r.stdout.replace(",", "").split()
    n1, n2, matches, good = int(words[1]), int(words[3]), int(words[6]), int(words[9])
    assert n1 > 200 and n2 > 200 and matches > 100 and 0 < good <= matches


@pytest.mark.gpu
def test_blocking_synchronize():
    """ygzb_synchronize_blocking (sleeping wait on a blocking event) completes the same work as ygzb_synchronize."""
    import ctypes as C
    import numpy as np
    from ygz_slam_b200 import Context, synth
    ctx = Context(0)
    fr = ctx.frames(1)
    fr.upload(synth.stream_frame(1)[0][None])
    ctx.lib.ygzb_synchronize_blocking.argtypes = [C.c_void_p]
    for _ in range(3):
        assert ctx.lib.ygzb_synchronize_blocking(ctx.h) == 0
    feats = fr.detect([0])
    assert ctx.lib.ygzb_synchronize_blocking(ctx.h) == 0 and feats[0]["n"] > 1000
    assert ctx.lib.ygzb_synchronize_blocking(None) != 0
    fr.close()
    ctx.close()
