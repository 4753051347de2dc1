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
