"""Multi-GPU partition of the path: one process per GPU, torch.distributed for plumbing.

The path shards by independent units (every path sample is independent, renderer.py:305-330) with ONE
exchange step: a sum-reduce of the float accumulation buffers.  Work is split by SAMPLE SLICE -- rank r
renders sample indices [first, first+n) of every pixel -- which is perfectly balanced whatever the image
content, and whose union is exactly the single-GPU sample set because the RNG is keyed by (pixel, sample);
by interleaved FILM TILE (16x8 pixels, the reference's launch block: tile t belongs to rank t mod N); or by
both (tile groups x sample slices, `partition`).
"""


def sample_slice(total_spp, rank, world):
    """(first_sample, n_samples) of `rank`; slices are contiguous, disjoint and cover [0, total_spp)."""
    base, rem = divmod(int(total_spp), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def tile_grid(width, height, tile_w=16, tile_h=8):
    """(tiles_x, tiles_y) of the film: the reference's 16x8 launch block (renderer.py:43-46,304-305) is the partition unit."""
    return (int(width) + tile_w - 1) // tile_w, (int(height) + tile_h - 1) // tile_h


def tile_slice(n_tiles, rank, world):
    """Film tiles of `rank` under the interleaved tile partition: t = ty * tiles_x + tx with t % world == rank.  Interleaving
    spreads sky, limb and cloud regions over all ranks (a contiguous split would hand one rank all of Apollo's empty space)."""
    return list(range(int(rank), int(n_tiles), int(world)))


def partition(total_spp, rank, world, tile_groups=1):
    """Tile (+ spp) partition of one frame over `world` ranks: `tile_groups` interleaved tile groups x world / tile_groups sample
    slices (BASELINE configs[3] uses e.g. 2 x 4 on 8 GPUs).  tile_groups = 1 is the pure spp slice, tile_groups = world
    the pure tile split.  Returns dict(tile_stride, tile_offset, first_sample, n_spp): rank r renders sample indices
    [first_sample, first_sample + n_spp) of the tiles t with t % tile_stride == tile_offset.  Every (pixel, sample) pair belongs
    to exactly one rank, so the sum of the ranks' buffers is the single-GPU frame."""
    world, g = int(world), int(tile_groups)
    if g < 1 or world % g:
        raise ValueError("tile_groups (%d) must divide the number of ranks (%d)" % (g, world))
    first, n = sample_slice(total_spp, rank // g, world // g)
    return {"tile_stride": g, "tile_offset": rank % g, "first_sample": first, "n_spp": n}


def frame_shard(n_frames, rank, world):
    """Flythrough batches (BASELINE configs[4]): frame f -> rank f mod world; no collective needed."""
    return list(range(rank, int(n_frames), int(world)))


def reduce_accumulation(buf, dst=0):
    """Sum the per-rank accumulation buffers onto `dst` (ncclReduce over NVLink on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return buf


def render_partition(renderer, part):
    """reset + render this rank's share `part` (from partition()) into its accumulation buffer; no exchange."""
    renderer.reset_framebuffer()
    if part["n_spp"]:
        tiles = (part["tile_stride"], part["tile_offset"]) if part["tile_stride"] > 1 else None
        renderer.accumulate(part["n_spp"], first_sample=part["first_sample"], tiles=tiles)


def render_distributed(renderer, total_spp, rank, world, dst=0, tile_groups=1):
    """Render this rank's share (spp slice, or tile group x spp slice) and sum-reduce; the caller resolves on `dst` with
    spp=total_spp.  Tiles of other groups stay zero in a rank's buffer, so one reduce(sum) serves every partition."""
    render_partition(renderer, partition(total_spp, rank, world, tile_groups))
    reduce_accumulation(renderer.color_buffer, dst)
    renderer.current_spp = total_spp
    return renderer.color_buffer


def resolve_fused(renderer, total_spp, rank, world, dst=0, tile_groups=1):
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
        others = [k for k in range(world) if k != dst]
        peers = [renderer.open_peer(handles[k]) for k in others]
        img = renderer.fetch_image_peers(peers, total_spp, tile_stride=tile_groups, own_offset=dst % tile_groups,
                                         peer_offsets=[k % tile_groups for k in others]).clone()
        torch.cuda.synchronize()
        renderer.close_peers()
    dist.barrier()                       # the peers' buffers may be reused from here on
    return img
