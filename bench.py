#!/usr/bin/env python3
"""Benchmark of the hot path on BASELINE.json's headline workload: dragon.cry, 600x400, 1024 spp.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

A step is one whole-frame render of the workload (every pixel, `spp` samples, up to max_depth bounces).  Prints ONE
JSON line on rank 0.  `value` is device-resident throughput (scene in HBM, film left on the device); `e2e` goes through
the host-buffer C-ABI call (film copied back to host memory every step).  See DESIGN.md "Measurement".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGORITHMIC_BYTES_PER_RAY = 904  # 32 ray + 16 hit + 8 levels x 80 B wide node + 3 x 72 B f64 triangles
ALGORITHMIC_BYTES_PER_RAY_F32 = 796  # the same with 3 x 36 B f32 triangles (the records as stored are 48 B)
WORKLOAD = "dragon.cry 600x400 (7 219 045-triangle procedural stand-in for xyzrgb_dragon.obj)"


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on init), so the
    process's fd 1 is pointed at stderr for the whole run and the result line goes to a private copy of the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


ALGORITHMIC_BYTES_PER_VERTEX = 336  # path state 64 + 64, hit 16, shading attributes 60, material 64, two rays 64, texel 4


KERNEL_SOURCES = ["wavefront.cu", "traverse.cuh", "shapes.cuh", "shading.cuh", "sampler.cuh", "cray_math.cuh", "device_types.cuh", "bvh_build.hpp", "Makefile"]


def source_sha256():
    """sha256 over the sources the wavefront kernels of libcray_b200.so are compiled from (kernels, their headers, the node and
    record layouts, the compiler flags).  (The binary itself cannot serve: two nvcc builds of the same sources differ in their
    internal symbol names.)"""
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES + ["../../include/cray_b200.h"]:
        path = os.path.normpath(os.path.join(ROOT, "craytracer_b200", "csrc", name))
        h.update(os.path.relpath(path, ROOT).encode() + b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def profiled_traffic(args=None, world=1):
    """DRAM bytes of each kernel from the committed ncu launch list of this workload (tools/final_capture.sh ->
    tools/kernel_traffic.py -> profiles/kernel_traffic.json): per render (= one bench step), per launch (the mean over the
    render's launches, one per bounce) and per ray.  The capture names the sources of the library it was taken on by sha256 and
    the workload (spp, film, triangles): if the library being timed now is built from other sources, or this run renders
    something else, the figures are not this run's and nothing is reported (null).  CRAY_B200_LIB (a tuning variant) never matches."""
    try:
        if os.environ.get("CRAY_B200_LIB"):
            return None
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            data = json.load(f)
        if data.get("source_sha256") != source_sha256():
            return None
        w = data.get("workload", {})
        if args is not None and (world != 1 or (w.get("spp"), w.get("width"), w.get("height"), w.get("triangles")) != (args.spp, args.width, args.height, args.triangles)):
            return None
        return data
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled every 250 ms DURING the timed region.  NVML is read in-process (pynvml): the
    `nvidia-smi -lms` loop of the profiling recipe costs this workload ~4 % (its query takes driver locks that delay the ~100
    kernel launches of a frame); nvidia-smi remains the fallback when pynvml is missing."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.sm, self.smax, self.reasons = [], [], set()
        self.source = None

    def _nvml_loop(self, nv, handle):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(handle))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.25)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            index = self.gpu
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                try:
                    index = int(visible.split(",")[self.gpu])
                except Exception:
                    pass
            handle = nv.nvmlDeviceGetHandleByIndex(index)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread:
            self.stop_flag.set()
            self.thread.join(timeout=2)
        elif self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            try:
                for line in open(self.path):
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        self.sm.append(float(parts[1]))
                        self.smax.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(self.NAMES, parts[5:9]):
                        if val.lower().startswith("active"):
                            self.reasons.add(name)
                os.unlink(self.path)
            except Exception:
                pass
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source"]}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": float(max(self.smax)) if self.smax else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def build_host_scene(args):
    import craytracer_b200 as c
    from craytracer_b200 import scenes
    scenes.register_standins(dragon_triangles=args.triangles)
    t0 = time.time()
    hs = c.parse_scene(scenes.dragon(num_samples=args.spp, width=args.width, height=args.height), base_dir=os.path.join(ROOT, "assets"))
    return hs, time.time() - t0


def cpu_baseline(hs, args, budget_s, threads=0):
    """The reference's algorithm (CPU oracle port) on the host cores over a bounded number of samples per pixel."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.OracleScene(hs)
    cores = threads or os.cpu_count() or 1
    t0 = time.time()
    _, counts = orc.render(args.width, args.height, seed=0, sample_begin=0, sample_end=1, threads=cores)
    dt1 = max(time.time() - t0, 1e-3)
    spp = int(max(1, min(args.spp, budget_s / dt1)))
    t0 = time.time()
    _, counts = orc.render(args.width, args.height, seed=0, sample_begin=0, sample_end=spp, threads=cores)
    dt = time.time() - t0
    rays = int(counts[0] + counts[1])
    return orc, {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                 "sample": f"{spp} of {args.spp} spp of the same frame ({rays} reference rays = Scene::intersect + Scene::intersects calls, in {dt:.1f} s); "
                           "rates are spp-independent",
                 "samples_per_s": args.width * args.height * spp / dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    hs, _ = build_host_scene(args)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.OracleScene(hs)
    cores = os.cpu_count() or 1
    # size each step to a few seconds of host work
    t0 = time.time()
    orc.render(args.width, args.height, seed=0, sample_begin=0, sample_end=1, threads=cores)
    dt1 = max(time.time() - t0, 1e-3)
    spp = int(max(1, min(args.spp, 6.0 / dt1)))
    for _ in range(args.warmup):
        orc.render(args.width, args.height, seed=0, sample_begin=0, sample_end=max(1, spp // 4), threads=cores)
    rays = 0
    t0 = time.time()
    for step in range(args.steps):
        _, counts = orc.render(args.width, args.height, seed=step, sample_begin=0, sample_end=spp, threads=cores)
        rays += int(counts[0] + counts[1])
    dt = time.time() - t0
    value = rays / dt / 1e6
    sample = f"each step = {spp} of {args.spp} spp of the frame on {cores} host threads; rates are spp-independent"
    line = {"impl": "reference", "metric": "Mrays/s on dragon.cry", "value": value, "unit": "Mrays/s",
            "value_counts": "reference rays = Scene::intersect + Scene::intersects calls (every one of them is traced by the CPU renderer)", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "spp": args.spp, "film": [args.width, args.height], "max_depth": 8, "sampler": "sobol"},
            "samples_per_s": args.width * args.height * spp * args.steps / dt,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import craytracer_b200 as c
    from craytracer_b200.distributed import shard_samples

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hs, parse_s = build_host_scene(args)
    t0 = time.time()
    also_f32 = world == 1 and args.mode == "fast" and not args.no_f32_leg
    scene = c.Scene(hs, device=local_rank, build=c.BUILD_EXACT | c.BUILD_FAST | (c.BUILD_F32 if args.mode == "f32" or also_f32 else 0))
    create_s = time.time() - t0
    mode = {"exact": c.TRAVERSE_EXACT, "fast": c.TRAVERSE_FAST, "f32": c.TRAVERSE_F32}[args.mode]
    lo, hi = shard_samples(args.spp, rank, world)
    n_film = args.width * args.height * 3
    film = torch.zeros(n_film, dtype=torch.float32, device="cuda")
    host_film = torch.zeros(n_film, dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream()

    host_np = host_film.numpy()
    job_bytes = 8 + 4 + 4  # the per-step host->device input of a render call: seed, sample_begin, sample_end (the scene is resident)

    def step(seed, e2e, mode=mode):
        if e2e and world == 1:
            # the reference-facing call with a HOST film buffer: cray_render (include/cray_b200.h, replaces render()
            # src/bin/craytracer.rs:224); device->host copy of the film inside the call
            st = scene.render_into(host_np, seed=seed, sample_begin=lo, sample_end=hi, mode=mode)
            np.divide(host_np, np.float32(args.spp), out=host_np)  # pixels /= num_samples (craytracer.rs:253-259)
            return st
        st = scene.render_device(film.data_ptr(), seed=seed, sample_begin=lo, sample_end=hi, mode=mode, stream=stream.cuda_stream)
        if world > 1:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            film.div_(float(args.spp))  # pixels /= num_samples (craytracer.rs:253-259)
            if e2e:
                host_film.copy_(film, non_blocking=True)  # the film the reference hands to on_render_finish
        return st

    def timed(e2e, mode=mode):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        totals = {"closest": 0, "shadow": 0, "shadow_traced": 0, "contact": 0, "launches": 0, "trace_ms": 0.0, "render_ms": 0.0, "iters": 0, "nan": 0, "shade_ms": 0.0, "shadow_ms": 0.0,
                  "generate_ms": 0.0}
        start.record(stream)
        for k in range(args.steps):
            t_host = time.perf_counter()
            st = step(k, e2e, mode)
            totals["host_ms"] = totals.get("host_ms", 0.0) + (time.perf_counter() - t_host) * 1e3
            totals["closest"] += st.closest_rays
            totals["shadow"] += st.shadow_rays
            totals["shadow_traced"] += st.shadow_rays_traced
            totals["contact"] += st.contact_rays
            totals["launches"] += st.kernel_launches + (1 if rank == 0 else 0)
            totals["trace_ms"] += st.trace_ms
            totals["shade_ms"] += st.shade_ms
            totals["shadow_ms"] += st.shadow_ms
            totals["generate_ms"] += st.generate_ms
            totals["render_ms"] += st.render_ms
            totals["iters"] += st.iterations
            totals["nan"] += st.nan_samples
        end.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device="cuda")
        counts = torch.tensor([totals["closest"], totals["shadow"], totals["launches"], totals["shadow_traced"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        return float(ms.item()), [float(x) for x in counts.tolist()], totals

    for k in range(args.warmup):
        step(1000 + k, False)  # both legs are warmed: the device-film leg (torch ops on the film load lazily) ...
        step(2000 + k, True)   # ... and the host-film leg through cray_render
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, counts_dev, totals = timed(False)
    ms_e2e, counts_e2e, _ = timed(True)
    clocks = sampler.stop() if sampler else None
    f32_leg = None
    if also_f32:  # the opt-in F32 traversal mode beside the headline (parity) mode: same frame, same timing rules
        step(3000, False, c.TRAVERSE_F32)
        ms32, counts32, totals32 = timed(False, c.TRAVERSE_F32)
        f32_leg = {"value": (counts32[0] + counts32[1]) / ms32 / 1e3, "unit": "Mrays/s", "ms_per_step": ms32 / args.steps,
                   "stage_ms_per_step": {"extend": totals32["trace_ms"] / args.steps, "shade": totals32["shade_ms"] / args.steps,
                                         "shadow": totals32["shadow_ms"] / args.steps, "generate": totals32["generate_ms"] / args.steps},
                   "extend_achieved_gbs": totals32["closest"] * ALGORITHMIC_BYTES_PER_RAY_F32 / (totals32["trace_ms"] * 1e-3) / 1e9,
                   "algorithmic_bytes_per_ray": ALGORITHMIC_BYTES_PER_RAY_F32,
                   "note": "CRAY_TRAVERSE_F32 (SURVEY 8f n4), opt-in: f32 watertight triangle tests on the same wide BVH, hits re-evaluated in f64; "
                           "not the headline because a ray grazing an edge may find the neighbouring triangle (tests/test_gpu_f32_mode.py)"}

    if rank == 0:
        # "reference rays": the Scene::intersect + Scene::intersects calls the reference makes for these samples (what the CPU arm
        # counts too, so the two arms' values are the same quantity).  "traced rays": what this implementation traced -- a path
        # vertex whose light sample cannot contribute (black contribution, all-specular material: the dragon is a conductor) makes
        # the reference's Scene::intersects call but needs no ray here.
        rays = counts_dev[0] + counts_dev[1]
        traced = counts_dev[0] + counts_dev[3]
        samples = args.width * args.height * args.spp * args.steps
        value = rays / ms_dev / 1e3
        peak, peak_kind = measured_peaks()
        # dominant kernel: k_wide_persistent<false> (closest-hit traversal); CUDA-event time of its launches on rank 0 inside the timed region
        extend_ms = totals["trace_ms"]
        bytes_per_ray = ALGORITHMIC_BYTES_PER_RAY_F32 if args.mode == "f32" else ALGORITHMIC_BYTES_PER_RAY
        achieved = totals["closest"] * bytes_per_ray / (extend_ms * 1e-3) / 1e9 if extend_ms > 0 else 0.0
        traffic = profiled_traffic(args, world) if args.mode == "fast" else None
        tk = (traffic or {}).get("kernels", {})

        def other(kernel, units, unit_name, per_unit, ms, key):
            got = units * per_unit / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            t = tk.get(key)
            return {"kernel": kernel, "bound": "hbm", "achieved": got, "peak": peak, "unit": "GB/s", "frac": got / peak, "units": units, "unit_name": unit_name,
                    "algorithmic_bytes_per_unit": per_unit, "kernel_ms": ms, "kernel_share_of_step": ms / max(totals["render_ms"], 1e-9),
                    "launches": totals["iters"], "algorithmic_bytes_per_launch": units * per_unit / max(totals["iters"], 1),
                    "traffic": t["dram_bytes_per_launch"] if t else None, "traffic_detail": t}

        line = {
            "metric": "Mrays/s on dragon.cry", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "spp": args.spp, "film": [args.width, args.height], "max_depth": 8, "sampler": "sobol", "traversal": args.mode, "parallelism": f"samples/{world}",
                       "l2": "per-step working set (triangle records 578 MB + 8-wide nodes 115 MB + shading records 924 MB) exceeds the 126 MB L2"},
            "value_counts": "reference rays = Scene::intersect + Scene::intersects calls of the reference for the same samples (the CPU arm counts the same)",
            "rays": {"reference_closest": counts_dev[0], "reference_shadow": counts_dev[1], "traced_shadow": counts_dev[3],
                     "reference_mrays_per_s": value, "traced_mrays_per_s": traced / ms_dev / 1e3,
                     "contact_rays_in_reference_order": totals["contact"]},
            "samples_per_s": samples / (ms_dev * 1e-3),
            "rays_per_sample": rays / samples,
            "e2e": {"value": (counts_e2e[0] + counts_e2e[1]) / ms_e2e / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": job_bytes, "d2h_bytes_per_step": n_film * 4,
                    "samples_per_s": samples / (ms_e2e * 1e-3), "traced_mrays_per_s": (counts_e2e[0] + counts_e2e[3]) / ms_e2e / 1e3,
                    "note": "cray_render through the C ABI with a host film buffer every step (N>1: device film + NCCL reduce + copy to pinned host memory); "
                            "the scene is resident like the reference's &Scene, so the per-step host->device input is the job description only; "
                            "scene upload is reported under setup.upload_ms"},
            "gpu_launches": int(counts_dev[2]),
            "roofline": {"bound": "hbm", "kernel": "k_wide_persistent<false, ExtendSource" + (", true" if args.mode == "f32" else "") + "> (closest-hit traversal, 8-wide BVH)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_kind, "algorithmic_bytes_per_ray": bytes_per_ray,
                         "rays_in_kernel": totals["closest"], "kernel_ms": extend_ms, "kernel_share_of_step": extend_ms / max(totals["render_ms"], 1e-9),
                         "launches": totals["iters"], "algorithmic_bytes_per_launch": totals["closest"] * bytes_per_ray / max(totals["iters"], 1),
                         "traffic": tk["extend"]["dram_bytes_per_launch"] if "extend" in tk else None, "traffic_detail": tk.get("extend"),
                         "traffic_capture": {k: v for k, v in (traffic or {}).items() if k != "kernels"} or None},
            "other_kernels": [
                other("shade stage: k_shade_classify + k_shade_class<matte|glass|plastic|metal> + k_shade_miss (one path vertex: emission, light sample, BSDF sample, Russian roulette)", totals["closest"], "path vertices", ALGORITHMIC_BYTES_PER_VERTEX,
                      totals["shade_ms"], "shade"),
                other("k_wide_persistent<true, ShadowSource> (any-hit traversal of the traced shadow rays)", totals["shadow_traced"], "traced shadow rays",
                      ALGORITHMIC_BYTES_PER_RAY, totals["shadow_ms"], "shadow")],
            "stage_ms_per_step": {"extend": totals["trace_ms"] / args.steps, "shade": totals["shade_ms"] / args.steps, "shadow": totals["shadow_ms"] / args.steps,
                                  "generate": totals["generate_ms"] / args.steps, "render_call": totals["render_ms"] / args.steps, "host_step": totals["host_ms"] / args.steps,
                                  "iterations": totals["iters"] / args.steps},
            "clocks": clocks,
            "setup": {"parse_and_standin_s": parse_s, "bvh_build_ms": scene.info.bvh_build_ms, "upload_ms": scene.info.upload_ms, "scene_create_s": create_s,
                      "wide_nodes": scene.info.wide_nodes, "wide_depth": scene.info.wide_depth, "triangles": hs.desc.n_triangles,
                      "contact_nodes": scene.info.contact_nodes, "contact_primitives": scene.info.contact_primitives},
            "dropped_samples": totals["nan"],
        }
        if f32_leg:
            line["f32_mode"] = f32_leg
        if world == 1 and not args.no_cpu_baseline:
            scene.close()
            _, base = cpu_baseline(hs, args, budget_s=args.cpu_budget)
            line["cpu_baseline"] = base
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_single_process(args):
    """All N GPUs from ONE process through the C ABI a Rust host would call: cray_scene_create_multi (one host-side BVH build,
    uploaded to every device) + cray_render_multi (a thread per device, sample ranges sharded, one ncclReduce of the films).
    `value` divides by the slowest GPU's CUDA-event time of each frame (the reduce excluded), `e2e` by the host's wall clock
    around the synchronous call (reduce and the copy of the film to host memory included)."""
    import torch

    import craytracer_b200 as c
    if not torch.cuda.is_available() or torch.cuda.device_count() < args.gpus:
        raise SystemExit(f"bench.py --single-process: {args.gpus} CUDA devices needed")
    hs, parse_s = build_host_scene(args)
    t0 = time.time()
    replicas = c.Scene.create_multi(hs, list(range(args.gpus)), build=c.BUILD_EXACT | c.BUILD_FAST | (c.BUILD_F32 if args.mode == "f32" else 0))
    create_s = time.time() - t0
    mode = {"exact": c.TRAVERSE_EXACT, "fast": c.TRAVERSE_FAST, "f32": c.TRAVERSE_F32}[args.mode]
    for k in range(args.warmup):
        c.render_multi(replicas, seed=1000 + k, sample_begin=0, sample_end=args.spp, mode=mode)
    sampler = ClockSampler(0)
    sampler.start()
    rays = traced = launches = 0
    device_ms = wall_ms = 0.0
    for k in range(args.steps):
        t_host = time.perf_counter()
        _, st = c.render_multi(replicas, seed=k, sample_begin=0, sample_end=args.spp, mode=mode)
        wall_ms += (time.perf_counter() - t_host) * 1e3
        device_ms += st.render_ms
        rays += st.closest_rays + st.shadow_rays
        traced += st.closest_rays + st.shadow_rays_traced
        launches += st.kernel_launches
    clocks = sampler.stop()
    samples = args.width * args.height * args.spp * args.steps
    n_film = args.width * args.height * 3
    info = replicas[0].info
    emit({"metric": "Mrays/s on dragon.cry", "value": rays / device_ms / 1e3, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "launch": "single process: cray_scene_create_multi + cray_render_multi (include/cray_b200.h)",
          "config": {"workload": WORKLOAD, "spp": args.spp, "film": [args.width, args.height], "max_depth": 8, "sampler": "sobol", "traversal": args.mode, "parallelism": f"samples/{args.gpus}"},
          "value_counts": "reference rays = Scene::intersect + Scene::intersects calls of the reference for the same samples",
          "rays": {"reference_mrays_per_s": rays / device_ms / 1e3, "traced_mrays_per_s": traced / device_ms / 1e3},
          "samples_per_s": samples / (device_ms * 1e-3),
          "e2e": {"value": rays / wall_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": n_film * 4, "samples_per_s": samples / (wall_ms * 1e-3),
                  "ms_per_step": wall_ms / args.steps, "note": "host wall clock around cray_render_multi: render on every GPU, ncclReduce, film to host memory"},
          "gpu_launches": int(launches), "clocks": clocks,
          "setup": {"parse_and_standin_s": parse_s, "bvh_build_ms": info.bvh_build_ms, "upload_ms": info.upload_ms, "scene_create_s": create_s,
                    "note": "ONE host-side BVH build for all devices (under torchrun every rank builds its own)"}})
    for r in replicas:
        r.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-f32-leg", action="store_true", help="skip the extra timed leg in the opt-in F32 traversal mode (N=1, --mode fast)")
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--width", type=int, default=600)
    ap.add_argument("--height", type=int, default=400)
    ap.add_argument("--triangles", type=int, default=7_219_045)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact", "f32"],
                    help="fast (default) and exact return the reference's hits bit for bit; f32 is the opt-in fast mode of SURVEY 8f n4")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of host work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all --gpus devices from this one process through cray_scene_create_multi / cray_render_multi instead of one rank per GPU")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.single_process:
        return run_single_process(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
