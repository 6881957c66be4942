#!/usr/bin/env python3
"""Summarise an ncu report's source page per CUDA source line: share of stall samples, of executed instructions, and the
average number of active threads per instruction.  usage: ncu_lines.py report.ncu-rep [kernel-regex] [top]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
    pattern = re.compile(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] else None
    selected = True
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, {}
    seen_kernel = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            selected = pattern is None or bool(pattern.search(r[1]))
            continue
        if not selected:
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0]:  # a CUDA source line with aggregated metrics
            key = (cur_file, int(r[0]))
            d = lines.setdefault(key, {"src": r[1].strip(), "samples": 0.0, "inst": 0.0, "thr": 0.0})
            idx = {name: i for i, name in enumerate(hdr)}

            def f(name):
                try:
                    return float(r[idx[name]].replace(",", ""))
                except Exception:
                    return 0.0
            d["samples"] += f("# Samples")
            d["inst"] += f("Instructions Executed")
            d["thr"] += f("Thread Instructions Executed")
    ts = sum(d["samples"] for d in lines.values()) or 1
    ti = sum(d["inst"] for d in lines.values()) or 1
    tt = sum(d["thr"] for d in lines.values())
    print(f"total: samples={ts:.0f} warp-inst={ti:.0f} avg active threads/inst={tt / ti:.2f}")
    for (file, line), d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print(f"{file:16s}:{line:4d} samp={d['samples'] / ts * 100:5.1f}% inst={d['inst'] / ti * 100:5.1f}% thr/inst={d['thr'] / max(d['inst'], 1):5.1f}  {d['src'][:100]}")


if __name__ == "__main__":
    main()
