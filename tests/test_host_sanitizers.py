"""AddressSanitizer + UndefinedBehaviorSanitizer over the host code that sits between untrusted input and the render call:
the hierarchy builder (csrc/rtb200_bvh.hpp), the baseline JPEG decoder and the scene-JSON reader. No GPU, no CUDA: the
harness tests/host_sanitize.cpp compiles those sources directly with g++ -fsanitize=address,undefined."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "rust-raytracer_b200")


def _small_jpegs(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    import numpy as np
    rng = np.random.default_rng(5)
    base = (rng.random((37, 53, 3)) * 255).astype(np.uint8)          # odd sizes: partial MCUs on both edges
    base[8:24, 10:40] = (200, 40, 90)
    out = []
    cases = [("s444", dict(subsampling=0)), ("s422", dict(subsampling=1)), ("s420", dict(subsampling=2)),
             ("q30", dict(quality=30)), ("q98", dict(quality=98, subsampling=0)),
             ("opt", dict(optimize=True)), ("prog", dict(progressive=True))]
    for name, kw in cases:
        p = str(tmp_path / f"{name}.jpg")
        Image.fromarray(base).save(p, "JPEG", **kw)
        out.append(p)
    p = str(tmp_path / "grey.jpg")
    Image.fromarray(base[:, :, 0]).save(p, "JPEG")
    out.append(p)
    p = str(tmp_path / "tiny.jpg")
    Image.fromarray(base[:1, :1]).save(p, "JPEG")
    out.append(p)
    return out


def test_host_code_under_asan_ubsan(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "host_sanitize")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-ffp-contract=off", "-o", exe, os.path.join(REPO, "tests", "host_sanitize.cpp"),
           os.path.join(PKG, "host", "jpeg_decode.cpp"), os.path.join(PKG, "host", "scene_json.cpp")]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and ("asan" in b.stderr or "ubsan" in b.stderr or "sanitize" in b.stderr):
        pytest.skip("sanitizer runtimes not installed: " + b.stderr[-200:])
    assert b.returncode == 0, b.stderr[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([exe, REPO] + _small_jpegs(tmp_path), capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:] + "\n" + r.stderr[-6000:])
    assert "host_sanitize: ok" in r.stdout
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
