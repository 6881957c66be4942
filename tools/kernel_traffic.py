#!/usr/bin/env python3
"""DRAM traffic per launch of the wavefront kernels from one `ncu --set full` report -> profiles/kernel_traffic.json.

usage: kernel_traffic.py report.ncu-rep lib_sha.txt "what was captured" [out.json]

`lib_sha.txt` is written ON THE GPU BOX by tools/final_capture.sh beside the report: the sha256 of the library's sources
(bench.source_sha256; first line) and of the binary (second line, informational: nvcc builds are not bit-reproducible).  bench.py
reports these figures as `roofline.traffic` only while the library it times is built from the same sources.  Per kernel
class the launch with the longest duration is taken (the capture window sits where all path slots are live)."""
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
           ("generate", r"k_generate"), ("extend_f32", r"k_wide_persistent<(\(bool\))?0, (cray::)?ExtendSource, (\(bool\))?1>")]


def main():
    rep, sha_file, what = sys.argv[1], sys.argv[2], sys.argv[3]
    out_path = sys.argv[4] if len(sys.argv) > 4 else "profiles/kernel_traffic.json"
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}

    def to_bytes(v, unit):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return float(v.replace(",", "")) * scale

    def to_ms(v, unit):
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]
        return float(v.replace(",", "")) * scale

    kernels = {}
    for key, needle in CLASSES:
        best = None
        for r in rows[2:]:
            if not re.search(needle, r[col["Kernel Name"]]):
                continue
            ms = to_ms(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
            if best is None or ms > best[0]:
                rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
                wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                best = (ms, rd, wr)
        if best:
            kernels[key] = {"dram_bytes_per_launch": best[1] + best[2], "dram_read": best[1], "dram_write": best[2], "launch_ms_under_ncu": best[0]}
    # the shade stage = its six kernels (the longest launch of each)
    parts = [kernels[k] for k in kernels if k.startswith("shade_")]
    if parts:
        kernels["shade"] = {f: sum(x[f] for x in parts) for f in ("dram_bytes_per_launch", "dram_read", "dram_write", "launch_ms_under_ncu")}
    lines = [ln.split()[0] for ln in open(sha_file).read().splitlines() if ln.strip()]
    data = {"source_sha256": lines[0], "library_sha256": lines[1] if len(lines) > 1 else None, "source": what, "report": rep, "kernels": kernels}
    with open(out_path, "w") as f:
        json.dump(data, f, indent=1)
        f.write("\n")
    print(json.dumps(data, indent=1))


if __name__ == "__main__":
    main()
