"""Row n3 of SURVEY 8(f): the reference's command line (src/bin/craytracer.rs:321-374) and its EXR output over the B200 path."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

import craytracer_b200 as c
from craytracer_b200 import scenes
from craytracer_b200.__main__ import main as cli_main

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_exr(path):
    """Minimal reader for what cray_write_exr writes: single-part scan-line, FLOAT channels B, G, R, no compression."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", data, 0)
    assert magic == 20000630 and version == 2
    pos, attrs = 8, {}
    while data[pos] != 0:
        end = data.index(b"\0", pos)
        name = data[pos:end].decode()
        pos = end + 1
        end = data.index(b"\0", pos)
        typ = data[pos:end].decode()
        pos = end + 1
        size, = struct.unpack_from("<i", data, pos)
        pos += 4
        attrs[name] = (typ, data[pos:pos + size])
        pos += size
    pos += 1
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    chlist, channels, cp = attrs["channels"][1], [], 0
    while chlist[cp] != 0:
        end = chlist.index(b"\0", cp)
        pixel_type, = struct.unpack_from("<i", chlist, end + 1)
        channels.append((chlist[cp:end].decode(), pixel_type))
        cp = end + 1 + 16
    assert channels == [("B", 2), ("G", 2), ("R", 2)], channels  # 2 = FLOAT
    offsets = struct.unpack_from(f"<{h}Q", data, pos)
    film = np.empty((h, w, 3), dtype=np.float32)
    for row, off in enumerate(offsets):
        y, size = struct.unpack_from("<ii", data, off)
        assert y == row and size == w * 12
        line = np.frombuffer(data, dtype="<f4", count=3 * w, offset=off + 8).reshape(3, w)
        film[row, :, 2], film[row, :, 1], film[row, :, 0] = line[0], line[1], line[2]
    return film


def test_exr_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    film = rng.uniform(0, 40, size=(9, 13, 3)).astype(np.float32)
    film[0, 0] = [0.0, np.float32(1e-30), np.float32(3e38)]
    path = tmp_path / "t.exr"
    c.write_exr(path, film)
    assert np.array_equal(read_exr(path), film)
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    try:
        import cv2
    except Exception:
        return
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is not None:  # an independent OpenEXR implementation reads the same pixels
        assert np.array_equal(img[..., ::-1], film)


def test_exr_errors(tmp_path):
    with pytest.raises(c.CrayError):
        c.write_exr(tmp_path / "no" / "such" / "dir" / "x.exr", np.zeros((2, 2, 3), dtype=np.float32))


@pytest.mark.parametrize("name", ["simple", "materials", "test", "rounding-error", "cornell", "dragon", "staircase"])
def test_scene_files_match_the_generators(name):
    """scenes/*.cry are what tools/write_scenes.py writes, and they load through the file entry point."""
    make = {"test": scenes.test_scene, "rounding-error": scenes.rounding_error}.get(name) or scenes.CONFIGS[name]
    text = open(os.path.join(ROOT, "scenes", name + ".cry")).read()
    assert text.split("\n", 1)[1].strip() == make().strip()
    scenes.register_standins(dragon_triangles=3000, interior_triangles=3000)
    hs = c.load_scene(os.path.join(ROOT, "scenes", name + ".cry"), base_dir=scenes.ASSETS)
    assert hs.desc.n_primitives > 0 and hs.desc.n_lights > 0


def test_cli_help_runs_without_a_gpu():
    res = subprocess.run([sys.executable, "-m", "craytracer_b200", "--help"], cwd=ROOT, capture_output=True, text=True)
    assert res.returncode == 0 and "--scene" in res.stdout and "--output" in res.stdout and "--seed" in res.stdout


@pytest.mark.gpu
def test_cli_renders_a_scene_file_to_exr(tmp_path):
    text = scenes.simple(num_samples=3, width=64, height=40)
    scene_path = tmp_path / "simple_small.cry"
    scene_path.write_text(text)
    out = tmp_path / "out.exr"
    assert cli_main(["--scene", str(scene_path), "--output", str(out), "--seed", "5"]) == 0
    film = read_exr(out)
    gpu = c.Scene(c.parse_scene(text))
    ref, _ = gpu.render(seed=5, sample_begin=0, sample_end=3)
    assert film.shape == (40, 64, 3)
    assert np.abs(film - ref / np.float32(3)).max() <= 1e-5 * max(1.0, float(ref.max()))
    out2 = tmp_path / "out2.exr"
    assert cli_main(["-s", str(scene_path), "--output", str(out2), "--spp", "2", "--mode", "exact"]) == 0
    assert read_exr(out2).shape == (40, 64, 3)


# ---- the compiled command line (craytracer_b200/csrc/cli_main.cpp -> craytracer_b200/cray_b200), C ABI only ----

CLI = os.path.join(ROOT, "craytracer_b200", "cray_b200")


@pytest.fixture(autouse=True, scope="module")
def _compiled_cli_present():
    """Built by __graft_entry__.build() / make beside the library; (re)build it here if a snapshot arrived without it."""
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "craytracer_b200", "csrc"), "../cray_b200"], capture_output=True)


def test_compiled_cli_arguments_and_errors(tmp_path):
    """The reference's flags (struct Cli, craytracer.rs:321-334); a parse error is logged with its location and, like the
    reference's main (:346-355), is not a failing exit; without a device the render fails loudly (no CPU path)."""
    res = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert res.returncode == 0 and all(flag in res.stdout for flag in ("--scene", "--output", "--seed", "--preview"))
    res = subprocess.run([CLI], capture_output=True, text=True)
    assert res.returncode == 2 and "--scene" in res.stderr
    res = subprocess.run([CLI, "--scene", "x.cry", "--mode", "bogus"], capture_output=True, text=True)
    assert res.returncode == 2
    bad = tmp_path / "bad.cry"
    bad.write_text("{ camera: Perspective { origin: Point(0,0,-5) target: Point(0,0,0) } }")
    res = subprocess.run([CLI, "--scene", str(bad)], capture_output=True, text=True)
    assert res.returncode == 0 and "[ERROR]" in res.stderr and f"{bad}:1:" in res.stderr
    res = subprocess.run([CLI, "--scene", str(tmp_path / "missing.cry")], capture_output=True, text=True)
    assert res.returncode == 1 and "[ERROR]" in res.stderr
    good = tmp_path / "simple_small.cry"
    good.write_text(scenes.simple(num_samples=1, width=32, height=20))
    res = subprocess.run([CLI, "--scene", str(good), "--output", str(tmp_path / "o.exr")], capture_output=True, text=True,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 1 and "no CUDA device" in res.stderr and not (tmp_path / "o.exr").exists()


@pytest.mark.gpu
def test_compiled_cli_renders_the_same_film_as_the_library(tmp_path):
    text = scenes.simple(num_samples=3, width=64, height=40)
    scene_path = tmp_path / "simple_small.cry"
    scene_path.write_text(text)
    out = tmp_path / "out.exr"
    res = subprocess.run([CLI, "--scene", str(scene_path), "--output", str(out), "--seed", "5"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "Scene constructed in" in res.stderr and "Rendering finished in" in res.stderr and f"Output written to {out}" in res.stderr
    gpu = c.Scene(c.parse_scene(text))
    ref, _ = gpu.render(seed=5, sample_begin=0, sample_end=3)
    film = read_exr(out)
    assert film.shape == (40, 64, 3)
    assert np.array_equal(film, ref / np.float32(3))
    out2 = tmp_path / "out2.exr"
    res = subprocess.run([CLI, "-s", str(scene_path), "--output", str(out2), "--spp", "2", "--mode", "f32", "--preview"], capture_output=True, text=True)
    assert res.returncode == 0 and read_exr(out2).shape == (40, 64, 3)


@pytest.mark.gpu
def test_both_command_lines_split_long_sample_ranges(tmp_path):
    """One render call takes at most 2^32 - 1 samples; the command lines render a longer frame as several sample ranges and add
    the film sums (here forced by --samples-per-call: 7 samples as 3 + 3 + 1)."""
    text = scenes.simple(num_samples=7, width=64, height=40)
    scene_path = tmp_path / "simple_small.cry"
    scene_path.write_text(text)
    whole, split, py_split = tmp_path / "whole.exr", tmp_path / "split.exr", tmp_path / "py_split.exr"
    for out, extra in ((whole, []), (split, ["--samples-per-call", "3"])):
        res = subprocess.run([CLI, "--scene", str(scene_path), "--output", str(out), "--seed", "2"] + extra, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    assert cli_main(["--scene", str(scene_path), "--output", str(py_split), "--seed", "2", "--samples-per-call", "3"]) == 0
    a, b, p = read_exr(whole), read_exr(split), read_exr(py_split)
    # (each range's film is rounded to f32 before the ranges are added)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6) and np.allclose(a, p, rtol=1e-5, atol=1e-6)
    assert not np.array_equal(a, read_exr(split) * 0)   # a real image
