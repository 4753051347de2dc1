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


def resolve_fused(renderer, total_spp, rank, world, dst=0):
    """Exchange step fused into the resolve: `dst` maps the other ranks' accumulation buffers (CUDA IPC over
    NVLink peer memory) and its resolve kernel sums them while tonemapping -- no reduce pass.  Returns the image on
    `dst`, None elsewhere.  All ranks must call it; ranks must be processes on one node."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return renderer.fetch_image(spp=total_spp)
    handles = [None] * world
    dist.all_gather_object(handles, renderer.export_accum_handle())
    torch.cuda.synchronize()
    dist.barrier()                       # every rank's samples are in its buffer
    img = None
    if rank == dst:
        peers = [renderer.open_peer(handles[k]) for k in range(world) if k != dst]
        img = renderer.fetch_image_peers(peers, total_spp).clone()
        torch.cuda.synchronize()
        renderer.close_peers()
    dist.barrier()                       # the peers' buffers may be reused from here on
    return img
