#!/usr/bin/env python3
"""DRAM traffic of the wavefront kernels -> profiles/kernel_traffic.json.

usage: kernel_traffic.py traffic.csv traffic.log lib_sha.txt "what was captured" [report.ncu-rep] [out.json]

`traffic.csv` is the ncu launch list of EVERY kernel launch of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-f32-leg`
(the default workload: 1024 spp) with gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum; `traffic.log` is
that run's stdout, whose JSON line holds the ray counts of one render.  The run renders the frame twice with the same seed (the
device-film leg and the host-film leg), so everything is halved to one render = one bench step.  With the path pool holding
the whole frame a render is one launch of each kernel per bounce, and the launches shrink bounce by bounce: the figures are
per-render sums and per-launch MEANS, the same way bench.py forms `roofline.achieved`.

`lib_sha.txt` is written ON THE GPU BOX by tools/final_capture.sh: the sha256 of the library's sources (bench.source_sha256;
first line) and of the binary (second line, informational: nvcc builds are not bit-reproducible).  bench.py reports these
figures as `roofline.traffic` only while the library it times is built from the same sources and the workload is the same.
The optional `--set full` report adds, per kernel, the figures of its longest captured launch (`full_set_launch`)."""
import csv
import io
import json
import re
import subprocess
import sys

# (ncu prints template arguments as `(bool)0, cray::ExtendSource` on the source page and as `0, ExtendSource` on the raw page)
CLASSES = [("extend", r"k_wide_persistent<(\(bool\))?0, (cray::)?ExtendSource, (\(bool\))?0>"), ("shadow", r"k_wide_persistent<(\(bool\))?1, (cray::)?ShadowSource, (\(bool\))?0>"),
           ("shade_classify", r"k_shade_classify"), ("shade_matte", r"k_shade_class<(\(unsigned int\))?0>"), ("shade_glass", r"k_shade_class<(\(unsigned int\))?1>"),
           ("shade_plastic", r"k_shade_class<(\(unsigned int\))?2>"), ("shade_metal", r"k_shade_class<(\(unsigned int\))?3>"), ("shade_miss", r"k_shade_miss"),
           ("generate", r"k_generate"), ("requeue", r"k_requeue"), ("begin_iteration", r"k_begin_iteration"), ("extend_f32", r"k_wide_persistent<(\(bool\))?0, (cray::)?ExtendSource, (\(bool\))?1>")]
SCALE_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
SCALE_MS = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
RENDERS = 2  # bench.py --steps 1 --warmup 0: the device-film leg and the host-film leg, same seed


def classify(name):
    for key, needle in CLASSES:
        if re.search(needle, name):
            return key
    return None


def launch_list(path):
    """ncu --csv --log-file: one row per (launch, metric) -> {id: {"kernel": .., metric: value}}"""
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    c = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    out = {}
    for r in rows[1:]:
        e = out.setdefault(int(r[c["ID"]]), {"kernel": r[c["Kernel Name"]]})
        v, unit, metric = float(r[c["Metric Value"]].replace(",", "")), r[c["Metric Unit"]], r[c["Metric Name"]]
        e[metric] = v * (SCALE_MS[unit] if metric == "gpu__time_duration.sum" else SCALE_B[unit])
    return out


def full_set(rep):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    best = {}
    for r in rows[2:]:
        key = classify(r[col["Kernel Name"]])
        if key is None:
            continue
        ms = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * SCALE_MS[units[col["gpu__time_duration.sum"]]]
        if key not in best or ms > best[key]["launch_ms_under_ncu"]:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * SCALE_B[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * SCALE_B[units[col["dram__bytes_write.sum"]]]
            best[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "launch_ms_under_ncu": ms}
    return best


def main():
    traffic_csv, traffic_log, sha_file, what = sys.argv[1:5]
    rest = sys.argv[5:]
    rep = next((a for a in rest if a.endswith(".ncu-rep")), None)
    out_path = next((a for a in rest if a.endswith(".json")), "profiles/kernel_traffic.json")
    line = next(json.loads(ln) for ln in open(traffic_log) if ln.startswith("{") and '"metric"' in ln)
    units = {"extend": (line["roofline"]["rays_in_kernel"], "closest-hit rays"), "shadow": (line["other_kernels"][1]["units"], "traced shadow rays"),
             "shade": (line["other_kernels"][0]["units"], "path vertices")}
    kernels = {}
    for e in launch_list(traffic_csv).values():
        key = classify(e["kernel"])
        if key is None:
            continue
        k = kernels.setdefault(key, {"launches_per_render": 0.0, "dram_read_per_render": 0.0, "dram_write_per_render": 0.0, "ms_under_ncu_per_render": 0.0})
        k["launches_per_render"] += 1.0 / RENDERS
        k["dram_read_per_render"] += e["dram__bytes_read.sum"] / RENDERS
        k["dram_write_per_render"] += e["dram__bytes_write.sum"] / RENDERS
        k["ms_under_ncu_per_render"] += e["gpu__time_duration.sum"] / RENDERS
    parts = [kernels[k] for k in kernels if k.startswith("shade_")]
    if parts:   # the shade stage = its six kernels, one launch of each per bounce
        kernels["shade"] = {f: sum(x[f] for x in parts) for f in parts[0]}
        kernels["shade"]["launches_per_render"] = kernels["shade_classify"]["launches_per_render"]
    for key, k in kernels.items():
        k["dram_bytes_per_render"] = k["dram_read_per_render"] + k["dram_write_per_render"]
        k["dram_bytes_per_launch"] = k["dram_bytes_per_render"] / k["launches_per_render"]
        if key in units:
            k["units_per_render"], k["unit_name"] = units[key]
            k["dram_bytes_per_unit"] = k["dram_bytes_per_render"] / k["units_per_render"]
    if rep:
        for key, v in full_set(rep).items():
            if key in kernels:
                kernels[key]["full_set_launch"] = v
    lines = [ln.split()[0] for ln in open(sha_file).read().splitlines() if ln.strip()]
    data = {"source_sha256": lines[0], "library_sha256": lines[1] if len(lines) > 1 else None, "source": what, "launch_list": traffic_csv, "report": rep,
            "workload": {"spp": line["config"]["spp"], "width": line["config"]["film"][0], "height": line["config"]["film"][1], "triangles": line["setup"]["triangles"], "renders_in_capture": RENDERS},
            "kernels": kernels}
    with open(out_path, "w") as f:
        json.dump(data, f, indent=1)
        f.write("\n")
    print(json.dumps({k: {f: v[f] for f in ("launches_per_render", "dram_bytes_per_launch", "dram_bytes_per_unit") if f in v} for k, v in kernels.items()}, indent=1))


if __name__ == "__main__":
    main()
