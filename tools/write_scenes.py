#!/usr/bin/env python3
"""Writes scenes/*.cry: the benchmark configurations of BASELINE.json (and the reference's two test scenes) in the reference's
.cry syntax, from the parametrised restatements in craytracer_b200/scenes.py at the reference files' own sizes and sample
counts.  Run from the repo root: python tools/write_scenes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from craytracer_b200 import scenes  # noqa: E402

FILES = {"simple": scenes.simple, "cornell": scenes.cornell, "materials": scenes.materials, "dragon": scenes.dragon, "staircase": scenes.staircase,
         "anthropic": scenes.anthropic, "test": scenes.test_scene, "rounding-error": scenes.rounding_error}

for name, make in FILES.items():
    path = os.path.join(ROOT, "scenes", name + ".cry")
    with open(path, "w") as f:
        f.write(f"// {name}.cry -- same scene as the reference's scenes/{name}.cry (written by tools/write_scenes.py)\n")
        f.write(make() + "\n")
    print(path)
