"""The reference's SAH tree built on the GPU (csrc/bvh_build_gpu.cu) against the host builder -- which tests/test_host_scene.py
holds against the oracle node for node: same nodes, same boxes, same split axes, same leaf order (src/bvh.rs:234-336)."""
import time

import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import scenes

pytestmark = pytest.mark.gpu


def _compare(hs):
    t0 = time.time()
    host_nodes, host_order = o.product_bvh(hs)
    t1 = time.time()
    dev_nodes, dev_order = o.product_bvh(hs, device=0)
    t2 = time.time()
    assert len(host_nodes) == len(dev_nodes)
    assert np.array_equal(host_order, dev_order), "leaf order differs"
    for f in ("min", "max", "axis", "a", "b"):
        assert np.array_equal(host_nodes[f], dev_nodes[f]), f
    return len(host_nodes), t1 - t0, t2 - t1


@pytest.mark.parametrize("triangles", [70_001, 300_000, 2_000_003])
def test_device_tree_equals_host_tree_on_the_dragon_standin(triangles):
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, triangles, 0)
    hs = c.parse_scene(scenes.dragon(), base_dir="/nonexistent")
    n, t_host, t_dev = _compare(hs)
    print(f"{triangles} triangles: {n} nodes; host build {t_host:.2f} s, device build {t_dev:.2f} s (both incl. primitive bounds and the dump)")


def test_device_tree_equals_host_tree_on_the_interior_standin():
    scenes.register_standins(interior_triangles=400_000)
    hs = c.parse_scene(scenes.staircase(), base_dir=scenes.ASSETS)
    _compare(hs)


def test_scene_created_with_the_device_tree_renders_the_same_film(monkeypatch):
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 150_001, 0)
    hs = c.parse_scene(scenes.dragon(width=96, height=64), base_dir="/nonexistent")
    films = []
    for flag in ("1", "0"):
        monkeypatch.setenv("CRAY_GPU_BUILD", flag)
        gpu = c.Scene(hs)
        film, st = gpu.render(seed=3, sample_begin=0, sample_end=4)
        films.append((film, st.closest_rays, st.shadow_rays))
        gpu.close()
    # (the film sums are f64 atomics: their order, and with it the last bit, differs from run to run)
    assert np.allclose(films[0][0], films[1][0], rtol=1e-6, atol=1e-9) and films[0][1:] == films[1][1:]
