"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the CPU oracle, the one leg of
bench.py that may execute oracle/) prints exactly ONE JSON line on stdout carrying the keys the contract names, whatever the
libraries it loads write; our arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--spp", "2", "--steps", "1", "--warmup", "0", "--triangles", "2000", "--width", "48", "--height", "32"]


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + SMALL, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"] + SMALL, capture_output=True, text=True, timeout=120,
                         cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_has_no_cpu_path():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode != 0 and res.stdout.strip() == ""
    assert "no CUDA device" in res.stderr


def test_traffic_figures_are_reported_only_for_the_code_and_workload_they_were_captured_on(monkeypatch):
    """roofline.traffic comes from a committed ncu capture (profiles/kernel_traffic.json): it names the kernels' sources by sha256
    and the workload; anything else gets null rather than another build's figures."""
    import argparse

    sys.path.insert(0, ROOT)
    import bench

    with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
        data = json.load(f)
    w = data["workload"]
    args = argparse.Namespace(spp=w["spp"], width=w["width"], height=w["height"], triangles=w["triangles"])
    for key in ("extend", "shadow", "shade"):
        k = data["kernels"][key]
        assert k["dram_bytes_per_launch"] * k["launches_per_render"] == k["dram_bytes_per_render"]
        assert abs(k["dram_bytes_per_unit"] * k["units_per_render"] - k["dram_bytes_per_render"]) <= 1e-6 * k["dram_bytes_per_render"]
    got = bench.profiled_traffic(args, 1)
    if data["source_sha256"] == bench.source_sha256():
        assert got is not None and got["kernels"]["extend"]["dram_bytes_per_launch"] > 0
    else:
        assert got is None          # kernels edited since the capture: nothing is claimed until it is retaken
    args.spp += 1
    assert bench.profiled_traffic(args, 1) is None
    args.spp -= 1
    assert bench.profiled_traffic(args, 2) is None
    monkeypatch.setenv("CRAY_B200_LIB", "/some/variant.so")
    assert bench.profiled_traffic(args, 1) is None
