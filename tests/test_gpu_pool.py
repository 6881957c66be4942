"""The path pool at every size: a pool that holds the whole job renders it as ONE wave (a launch of each kernel per bounce, the
extend queue rebuilt from the previous one by k_requeue once every sample has started); a pool smaller than the job keeps
refilling the slots of finished paths (k_generate) until the samples run out.  Both schedules trace the same rays for the same
samples, so the ray counts are equal exactly and the films up to the order of the f64 film atomics -- and the small-pool film is
held against the oracle like any other (src/bin/craytracer.rs:148-206: a sample depends on (x, y, sample_index, seed) only)."""
import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import scenes

pytestmark = pytest.mark.gpu


def _scene(name):
    if name == "materials":      # every shade class, analytic shapes
        return c.parse_scene(scenes.materials(num_samples=8, width=64, height=48), base_dir=scenes.ASSETS)
    if name == "cornell":        # planar contacts: rays in the back part of the queue (reference-order traversal)
        return c.parse_scene(scenes.cornell(num_samples=8, width=48, height=48), base_dir=scenes.ASSETS)
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 40_001, 0)
    return c.parse_scene(scenes.dragon(num_samples=8, width=64, height=48), base_dir="/nonexistent")


@pytest.mark.parametrize("name", ["materials", "cornell", "dragon"])
@pytest.mark.parametrize("mode", [c.TRAVERSE_FAST, c.TRAVERSE_EXACT])
def test_small_pools_trace_the_same_rays_as_one_wave(name, mode, monkeypatch):
    hs = _scene(name)
    gpu = c.Scene(hs, build=c.BUILD_EXACT | c.BUILD_FAST)
    results = {}
    for log2 in ("28", "13", "10"):      # 24 576 samples (cornell 18 432): one wave, three or four pool-fulls, twenty-odd pool-fulls
        monkeypatch.setenv("CRAY_POOL_LOG2", log2)
        film, st = gpu.render(seed=5, sample_begin=0, sample_end=8, mode=mode)
        results[log2] = (film.astype(np.float64), (st.closest_rays, st.shadow_rays, st.shadow_rays_traced, st.contact_rays, st.nan_samples, st.samples), st.iterations)
    gpu.close()
    one = results["28"]
    assert one[2] <= hs.desc.max_depth + 1, "a pool that holds the job needs one iteration per bounce"
    for log2 in ("13", "10"):
        film, counts, iterations = results[log2]
        assert counts == one[1], (log2, counts, one[1])
        assert iterations > one[2], "the small pool should have needed more iterations"
        assert np.allclose(film, one[0], rtol=1e-6, atol=1e-9), log2
    if name == "cornell" and mode == c.TRAVERSE_FAST:
        assert one[1][3] > 0, "the Cornell box has coplanar contacts: some rays must have gone the reference-order way"


@pytest.mark.parametrize("pool", ["28", "12"])
def test_sample_order_does_not_change_the_samples(pool, monkeypatch):
    """Paths start pixel by pixel, all samples of a pixel together (the default), or `CRAY_SAMPLE_GROUP` samples of every pixel at
    a time (1 = the whole frame once per sample index; 3 does not divide the 8 samples: the last group is shorter)."""
    hs = _scene("materials")
    gpu = c.Scene(hs, build=c.BUILD_EXACT | c.BUILD_FAST)
    monkeypatch.setenv("CRAY_POOL_LOG2", pool)
    results = []
    for group in (None, "1", "3", "100"):
        if group is None:
            monkeypatch.delenv("CRAY_SAMPLE_GROUP", raising=False)
        else:
            monkeypatch.setenv("CRAY_SAMPLE_GROUP", group)
        film, st = gpu.render(seed=9, sample_begin=2, sample_end=10)
        results.append((film.astype(np.float64), (st.closest_rays, st.shadow_rays, st.shadow_rays_traced, st.nan_samples, st.samples)))
    gpu.close()
    for film, counts in results[1:]:
        assert counts == results[0][1]
        assert np.allclose(film, results[0][0], rtol=1e-6, atol=1e-9)


def test_small_pool_film_equals_the_oracle(monkeypatch):
    hs = c.parse_scene(scenes.materials(num_samples=4, width=40, height=26), base_dir=scenes.ASSETS)
    orc = o.OracleScene(hs)
    want, counts = orc.render(40, 26, seed=2, sample_begin=0, sample_end=4)
    monkeypatch.setenv("CRAY_POOL_LOG2", "10")
    gpu = c.Scene(hs, build=c.BUILD_EXACT | c.BUILD_FAST)
    film, st = gpu.render(seed=2, sample_begin=0, sample_end=4, mode=c.TRAVERSE_EXACT)
    gpu.close()
    orc.close()
    assert st.iterations > hs.desc.max_depth + 1
    assert (st.closest_rays, st.shadow_rays, st.nan_samples) == (int(counts[0]), int(counts[1]), int(counts[2]))
    rel_mse = float(np.mean((film - want) ** 2 / (want ** 2 + 1e-2)))
    assert rel_mse <= 1e-10, rel_mse            # the tolerance of tests/test_gpu_parity.py for equal-sample renders
