"""world_size-2 gloo test of the multi-GPU host logic on CPU: sample-slice partition + sum-reduce.
The CPU oracle stands in for the device renderer (checker only; the product never runs on CPU)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from digital_earth_b200 import distributed as dd  # noqa: E402


def test_sample_slices_cover_exactly():
    for spp in (1, 7, 64, 1024, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                f, n = dd.sample_slice(spp, r, world)
                seen += list(range(f, f + n))
            assert seen == list(range(spp))
    assert dd.frame_shard(10, 1, 4) == [1, 5, 9] and sum(len(dd.frame_shard(120, r, 8)) for r in range(8)) == 120


def _worker(rank, world, port, spp, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import digital_earth_b200 as de
    from oracle import oracle as orc
    tex = de.textures.synthetic(64, 32, seed=7)
    cfg = de.load_config(os.path.join(ROOT, "digital-earth_b200", "assets", "configs", "config - florida.txt"))
    s = orc.Scene(tex, 32, 16, cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"], aspect_scale=cfg["aspect_scale"],
                  sun_angle=cfg["sun_angle"], sun_path_rot=cfg["sun_path_rot"])
    first, n = dd.sample_slice(spp, rank, world)
    acc, _ = orc.render(s, n, first_sample=first, seed=3, nthreads=2)
    buf = torch.from_numpy(acc)
    dd.reduce_accumulation(buf, dst=0)
    if rank == 0:
        whole, _ = orc.render(s, spp, seed=3, nthreads=2)
        np.save(out_path, np.stack([buf.numpy(), whole]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slice_and_reduce_equals_single_rank(tmp_path):
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, port, 5, out), nprocs=2, join=True)
    got, want = np.load(out)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())  # identical (pixel, sample) keys: equal up to float summation order
    assert want.sum() > 0


def test_tile_and_combined_partitions_cover_every_pixel_sample_pair_once():
    for (w, h) in ((1920, 1080), (3840, 2160), (640, 360), (48, 24)):
        tx, ty = dd.tile_grid(w, h)
        assert (tx, ty) == ((w + 15) // 16, (h + 7) // 8)
        for world in (1, 2, 4, 8):
            tiles = sorted(t for r in range(world) for t in dd.tile_slice(tx * ty, r, world))
            assert tiles == list(range(tx * ty))
            sizes = [len(dd.tile_slice(tx * ty, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    for world, groups, spp in ((8, 2, 4096), (8, 1, 1024), (8, 8, 7), (4, 2, 5), (2, 2, 3), (1, 1, 9)):
        cover = {}
        for r in range(world):
            p = dd.partition(spp, r, world, groups)
            assert p["tile_stride"] == groups and 0 <= p["tile_offset"] < groups
            for sm in range(p["first_sample"], p["first_sample"] + p["n_spp"]):
                cover[(p["tile_offset"], sm)] = cover.get((p["tile_offset"], sm), 0) + 1
        assert cover == {(g, sm): 1 for g in range(groups) for sm in range(spp)}
    import pytest
    with pytest.raises(ValueError):
        dd.partition(16, 0, 8, 3)


def _tile_worker(rank, world, port, spp, groups, out_path):
    """tile group x spp slice on the oracle (window renders per film tile), gloo reduce, against the whole frame."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import digital_earth_b200 as de
    from oracle import oracle as orc
    W, H = 64, 32
    tex = de.textures.synthetic(64, 32, seed=7)
    cfg = de.load_config(os.path.join(ROOT, "digital-earth_b200", "assets", "configs", "config - florida.txt"))
    s = orc.Scene(tex, W, H, cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"], aspect_scale=cfg["aspect_scale"],
                  sun_angle=cfg["sun_angle"], sun_path_rot=cfg["sun_path_rot"])
    p = dd.partition(spp, rank, world, groups)
    tx, ty = dd.tile_grid(W, H)
    acc = np.zeros((H, W, 3), np.float32)
    for t in dd.tile_slice(tx * ty, p["tile_offset"], p["tile_stride"]):
        win = ((t % tx) * 16, (t // tx) * 8, 16, 8)
        part, _ = orc.render(s, p["n_spp"], first_sample=p["first_sample"], seed=3, window=win, nthreads=1)
        acc += part
    buf = torch.from_numpy(acc)
    dd.reduce_accumulation(buf, dst=0)
    if rank == 0:
        whole, _ = orc.render(s, spp, seed=3, nthreads=2)
        np.save(out_path, np.stack([buf.numpy(), whole]))
    dist.barrier()
    dist.destroy_process_group()


def test_tile_plus_spp_partition_and_reduce_equals_single_rank(tmp_path):
    for world, groups in ((2, 2), (4, 2)):
        with socket.socket() as so:
            so.bind(("127.0.0.1", 0))
            port = so.getsockname()[1]
        out = str(tmp_path / ("acc_%d_%d.npy" % (world, groups)))
        mp.spawn(_tile_worker, args=(world, port, 4, groups, out), nprocs=world, join=True)
        got, want = np.load(out)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
        assert want.sum() > 0
