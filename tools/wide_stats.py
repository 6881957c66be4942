#!/usr/bin/env python3
"""Work counters of the persistent traversal kernels on the bench workload (GPU box, stats build of the library):
    make -C craytracer_b200/csrc VARIANT=stats EXTRA=-DCRAY_WIDE_STATS=1
    CRAY_B200_LIB=$PWD/craytracer_b200/libcray_b200_stats.so python tools/wide_stats.py [spp]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import craytracer_b200 as c  # noqa: E402
from craytracer_b200 import _abi, scenes  # noqa: E402


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    scenes.register_standins()
    hs = c.parse_scene(scenes.dragon(num_samples=spp), base_dir=os.path.join(ROOT, "assets"))
    scene = c.Scene(hs)
    out = np.zeros(24, dtype=np.uint64)
    _abi.lib().cray_debug_wide_stats(out.ctypes.data)  # clear
    _, st = scene.render(seed=0, sample_begin=0, sample_end=spp)
    rc = _abi.lib().cray_debug_wide_stats(out.ctypes.data)
    assert rc == 0, _abi.lib().cray_last_error()
    print(f"wide nodes {scene.info.wide_nodes}, depth {scene.info.wide_depth}; closest rays {st.closest_rays}, shadow rays traced+skipped {st.shadow_rays}, render {st.render_ms:.1f} ms")
    for name, v in (("closest", out[:12]), ("any", out[12:])):
        rays, iters, node_steps, node_phases, tests, rounds, refills, idle, empty, ihits, noprim, afterhit = [float(x) for x in v]
        if rays == 0:
            continue
        print(f"{name:8s} rays {rays:.0f}  node steps/ray {node_steps / rays:.2f}  tests/ray {tests / rays:.2f}  "
              f"warp iterations per 32 rays {iters / rays * 32:.1f}  lanes per node phase {node_steps / max(node_phases, 1):.1f}  "
              f"tests per round {tests / max(rounds, 1):.1f}  rounds per iteration {rounds / iters:.2f}  idle lanes/iteration {idle / iters:.1f}  "
              f"node steps that hit nothing {empty / node_steps:.3f}  that queued no primitive {noprim / node_steps:.3f}  taken after a first hit {afterhit / node_steps:.3f}  "
              f"interior children hit per step {ihits / node_steps:.2f}")


if __name__ == "__main__":
    main()
