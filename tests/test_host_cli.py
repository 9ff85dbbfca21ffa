"""The drop-in CLI `raytracer <config_file> <output_file>` (reference main.rs:7-20) and the host-side scene staging."""
import json
import os
import subprocess

import numpy as np
import pytest

import rtb200 as R
from rtb200 import scenes

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "rust-raytracer_b200", "raytracer")


@pytest.fixture(scope="module", autouse=True)
def _cli_built():
    """The CLI is a build artefact (git-ignored); __graft_entry__.build() makes it, a bare checkout may not have it yet."""
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "rust-raytracer_b200"), "raytracer"])


def _run(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=300, **kw)


def test_usage_line_when_argc_is_not_3():
    for args in ([], ["a"], ["a", "b", "c"]):
        r = _run(args)
        assert r.returncode == 0 and r.stdout == f"Usage: {CLI} <config_file> <output_file>\n"     # main.rs:9-12


def test_unreadable_or_invalid_config_exits_like_a_panic(tmp_path):
    r = _run([str(tmp_path / "missing.json"), str(tmp_path / "o.png")])
    assert r.returncode == 101 and "Unable to read config file" in r.stderr                          # main.rs:14
    bad = tmp_path / "bad.json"; bad.write_text('{"width": 10, "height": ')
    r = _run([str(bad), str(tmp_path / "o.png")])
    assert r.returncode == 101 and "Unable to parse config json" in r.stderr                          # main.rs:15
    cfg = scenes._variant(scenes.cover_config(), 8, 6, 1, 2); cfg["objects"][1]["material"] = {"Plastic": {}}
    bad.write_text(json.dumps(cfg))
    r = _run([str(bad), str(tmp_path / "o.png")])
    assert r.returncode == 101 and "unknown variant `Plastic`" in r.stderr
    del cfg["objects"][1]["material"]; bad.write_text(json.dumps(cfg))
    r = _run([str(bad), str(tmp_path / "o.png")])
    assert r.returncode == 101 and "missing field `material`" in r.stderr


def test_jpeg_decoder_agrees_with_libjpeg_within_rounding():
    from PIL import Image
    for name in ("earth.jpg", "beach.jpg"):      # 4:4:4 with Adobe marker; 4:2:0 with restart intervals
        path = os.path.join(scenes.SCENES_DIR, "data", name)
        a = R._decode_jpeg(path).astype(int)
        b = np.asarray(Image.open(path).convert("RGB")).astype(int)
        assert a.shape == b.shape
        d = np.abs(a - b)
        assert d.max() <= 4 and d.mean() < 0.3
    assert R._decode_jpeg(os.path.join(scenes.SCENES_DIR, "data", "earth.jpg")).shape == (1024, 2048, 3)   # config.rs:135-146


def test_scene_json_schema_strings_of_the_reference_tests(tmp_path):
    """The literal JSON strings the reference's serialisation tests pin (config.rs:101,128; sphere.rs:101; camera.rs:134)
    parse into the same scene through the Python host mirror."""
    s1 = '{"width":100,"height":100,"samples_per_pixel":1,"max_depth":1,"sky":{"texture":""},"camera":{"look_from":{"x":0.0,"y":0.0,"z":0.0},"look_at":{"x":0.0,"y":0.0,"z":-1.0},"vup":{"x":0.0,"y":1.0,"z":0.0},"vfov":90.0,"aspect":1.0},"objects":[{"center":{"x":0.0,"y":0.0,"z":-1.0},"radius":0.5,"material":{"Lambertian":{"albedo":[0.8,0.3,0.3]}}}]}'
    sc = R.Scene.from_config(json.loads(s1))
    assert (sc.c.width, sc.c.height, sc.c.samples_per_pixel, sc.c.max_depth, sc.n_spheres, sc.c.sky.mode) == (100, 100, 1, 1, 1, R.RT_SKY_GRADIENT)
    assert list(sc._spheres[0].albedo) == [np.float32(0.8), np.float32(0.3), np.float32(0.3)]
    s2 = s1.replace('"sky":{"texture":""}', '"sky":null')
    assert R.Scene.from_config(json.loads(s2)).c.sky.mode == R.RT_SKY_NONE                     # config.rs:128 + raytracer.rs:138-140
    s3 = s1.replace('"sky":{"texture":""}', '"sky":{"texture":"data/earth.jpg"}')
    sk = R.Scene.from_config(json.loads(s3), scenes.SCENES_DIR).c.sky
    assert (sk.mode, sk.tex.width, sk.tex.height) == (R.RT_SKY_TEXTURE, 2048, 1024)         # config.rs:131-146


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_cli_without_gpu_fails_loudly(tmp_path):
    p = tmp_path / "s.json"; p.write_text(json.dumps(scenes._variant(scenes.cover_config(), 16, 12, 1, 4)))
    r = _run([str(p), str(tmp_path / "o.png")])
    assert r.returncode == 101 and "render failed" in r.stderr and not (tmp_path / "o.png").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["cover", "test_scene"])
def test_cli_renders_the_same_png_as_the_python_host(tmp_path, which):
    from PIL import Image
    cfg = scenes._variant(scenes.cover_config(), 96, 72, 4, 12) if which == "cover" else scenes._variant(scenes.test_scene_config(), 80, 60, 4, 6)
    p = tmp_path / "scene.json"; p.write_text(json.dumps(cfg))
    out = tmp_path / "out.png"
    r = _run([str(p), str(out)], cwd=scenes.SCENES_DIR)           # texture paths are relative to the CWD (materials.rs:214)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    assert lines[0] == "" and lines[1] == f"Rendering {out}" and lines[2].startswith("Frame time: ") and lines[2].endswith("ms")   # main.rs:18, raytracer.rs:263
    png = np.asarray(Image.open(out))
    ref, _ = R.render_rgb8(R.Scene.from_config(cfg, scenes.SCENES_DIR))
    assert png.shape == (cfg["height"], cfg["width"], 3) and np.array_equal(png, ref)


@pytest.mark.gpu
def test_cli_on_all_gpus_writes_the_same_png_as_on_one(tmp_path):
    """`RTB200_GPUS=0 raytracer scene.json out.png` (every GPU of the box through rtb200_render_rgb8_multi) must write the
    same image as the single-GPU run. Skips below 2 GPUs."""
    from PIL import Image
    if R.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cfg = scenes._variant(scenes.cover_config(), 160, 120, 8, 50)
    p = tmp_path / "scene.json"; p.write_text(json.dumps(cfg))
    one, many = tmp_path / "one.png", tmp_path / "all.png"
    r1 = _run([str(p), str(one)], cwd=scenes.SCENES_DIR)
    env = dict(os.environ, RTB200_GPUS="0", RTB200_STATS="1")
    r2 = subprocess.run([CLI, str(p), str(many)], capture_output=True, text=True, cwd=scenes.SCENES_DIR, env=env)
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    assert f"gpus={R.device_count()}" in r2.stderr
    assert np.array_equal(np.asarray(Image.open(one)), np.asarray(Image.open(many)))


def test_jpeg_decoder_survives_malformed_headers(tmp_path):
    """load_texture_image runs on scene-named files: malformed segments must produce an error (the reference's jpeg-decoder
    crate returns Err and the caller panics cleanly), never an out-of-bounds read. Mutates the header bytes of a small
    baseline JPEG (lengths, table ids, precisions, dimensions, truncations) and decodes every mutant."""
    import ctypes as C
    import io
    from PIL import Image
    rng = np.random.default_rng(5)
    img = Image.fromarray(rng.integers(0, 255, (24, 40, 3), dtype=np.uint8), "RGB")
    buf = io.BytesIO(); img.save(buf, format="JPEG", quality=80, subsampling=2); good = buf.getvalue()
    sos = good.index(b"\xff\xda")
    L = R.lib()

    def decode(data):
        p = tmp_path / "m.jpg"; p.write_bytes(data)
        out = C.c_void_p(); w = C.c_uint64(); h = C.c_uint64()
        rc = L.rtb200_decode_jpeg_file(str(p).encode(), C.byref(out), C.byref(w), C.byref(h))
        if rc == 0:
            assert 0 < w.value <= 65535 and 0 < h.value <= 65535
            L.rtb200_free(out)
        return rc

    assert decode(good) == 0
    n_err = 0
    for trial in range(300):
        d = bytearray(good)
        kind = trial % 4
        if kind == 0:                                   # one random header byte
            d[int(rng.integers(2, sos + 14))] = int(rng.integers(0, 256))
        elif kind == 1:                                 # a segment length set to something tiny / huge
            pos = [i for i in range(2, sos) if d[i] == 0xFF and d[i + 1] not in (0x00, 0xFF)]
            k = pos[int(rng.integers(len(pos)))]
            d[k + 2], d[k + 3] = [(0, 0), (0, 1), (0, 2), (0, 3), (255, 255), (0, int(rng.integers(4, 40)))][int(rng.integers(6))]
        elif kind == 2:                                 # truncation
            d = d[: int(rng.integers(4, len(d)))]
        else:                                           # several random header bytes
            for _ in range(4):
                d[int(rng.integers(2, sos + 14))] = int(rng.integers(0, 256))
        n_err += decode(bytes(d)) != 0
    assert n_err > 50
