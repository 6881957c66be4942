"""Multi-GPU rendering: sample-index ranges per rank, scene replicated, one reduce of the film.

The reference's only parallelism is shared-memory work stealing over (tile, 8-sample batch) units
(src/bin/craytracer.rs:22-43, :271-291).  Units (pixel, sample_index) are independent, so rank r of N takes the
sample indices [r*spp/N, (r+1)*spp/N) of every pixel (perfect balance, the union is exactly the 1-GPU sample set) and
the per-rank SUM films are added with a single ``torch.distributed.reduce`` (NCCL over NVLink on GPUs, gloo in the CPU
tests).  No other exchange exists on this path.
"""
import torch
import torch.distributed as dist


def shard_samples(num_samples, rank, world_size, sample_begin=0):
    """Contiguous sample-index range of `rank`; ranges tile [sample_begin, sample_begin + num_samples) exactly."""
    lo = sample_begin + (num_samples * rank) // world_size
    hi = sample_begin + (num_samples * (rank + 1)) // world_size
    return lo, hi


def reduce_film(film_sum, num_samples, dst=0):
    """Sum the per-rank film tensors onto `dst` and divide by the total sample count there
    (pixels /= num_samples, src/bin/craytracer.rs:253-259).  Returns the mean film on dst, None elsewhere."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_sum, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return film_sum / float(num_samples)


def render_sharded(render_range, num_samples, film_shape, device, sample_begin=0):
    """render_range(lo, hi, out_tensor) must fill `out_tensor` (f32, film_shape) with the SUM film of samples [lo, hi).
    Returns (mean film on rank 0 | None, (lo, hi))."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_samples(num_samples, rank, world, sample_begin)
    film = torch.zeros(film_shape, dtype=torch.float32, device=device)
    if hi > lo:
        render_range(lo, hi, film)
    return reduce_film(film, num_samples), (lo, hi)
