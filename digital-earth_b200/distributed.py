"""Multi-GPU partition of the path: one process per GPU, torch.distributed for plumbing.

The path shards by independent units (every path sample is independent, renderer.py:305-330) with ONE
exchange step: a sum-reduce of the float accumulation buffers.  Work is split by SAMPLE SLICE -- rank r
renders sample indices [first, first+n) of every pixel -- which is perfectly balanced whatever the image
content, and whose union is exactly the single-GPU sample set because the RNG is keyed by (pixel, sample).
"""


def sample_slice(total_spp, rank, world):
    """(first_sample, n_samples) of `rank`; slices are contiguous, disjoint and cover [0, total_spp)."""
    base, rem = divmod(int(total_spp), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def frame_shard(n_frames, rank, world):
    """Flythrough batches (BASELINE configs[4]): frame f -> rank f mod world; no collective needed."""
    return list(range(rank, int(n_frames), int(world)))


def reduce_accumulation(buf, dst=0):
    """Sum the per-rank accumulation buffers onto `dst` (ncclReduce over NVLink on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return buf


def render_distributed(renderer, total_spp, rank, world, dst=0):
    """Render this rank's sample slice and reduce; the caller resolves on `dst` with spp=total_spp."""
    first, n = sample_slice(total_spp, rank, world)
    renderer.reset_framebuffer()
    if n:
        renderer.accumulate(n, first_sample=first)
    reduce_accumulation(renderer.color_buffer, dst)
    renderer.current_spp = total_spp
    return renderer.color_buffer
