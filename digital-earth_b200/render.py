"""Headless renderer CLI: what `main.py` + the GGUI loop do, without a window.

    python -m digital_earth_b200.render --config "config - florida.txt" --res 1920x1080 --spp 1024 \\
        --textures synthetic:8192x4096 --out florida.png
    torchrun --nproc-per-node 8 -m digital_earth_b200.render --config ... --orbit 120 --out-dir frames/

One process per GPU.  A still frame is split by sample slice (or --tile-groups: film-tile groups x sample
slices) across ranks and sum-reduced with NCCL;
an orbit (camera and look-at rotated about +Y, BASELINE configs[4]) is sharded by frame, no collective.
"""
import argparse
import math
import os
import sys


def orbit_config(cfg, angle_rad):
    """Rotate camera position, look-at and up about the +Y (polar) axis."""
    c, s = math.cos(angle_rad), math.sin(angle_rad)

    def rot(v):
        return (c * v[0] + s * v[2], v[1], -s * v[0] + c * v[2])
    out = dict(cfg)
    out["cam_pos"], out["look_at"], out["up"] = rot(cfg["cam_pos"]), rot(cfg["look_at"]), rot(cfg["up"])
    return out


def parse_textures(spec, scene_hint=""):
    from . import textures
    if spec.startswith("synthetic"):
        w, h = (int(x) for x in (spec.split(":")[1] if ":" in spec else "2048x1024").split("x"))
        stormy = "hurricane" in scene_hint
        return textures.synthetic(w, h, cloud_cover=0.8 if stormy else 0.5, hurricane=stormy)
    return textures.load_directory(spec)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", required=True, help="10-line scene file (earth_viewer.py:100-126,213-236)")
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--batch", type=int, default=256, help="samples per launch")
    ap.add_argument("--textures", default="textures", help="directory with the NASA maps, or synthetic:WxH")
    ap.add_argument("--out", default=None, help="image file (default: screenshot/<main>-<timestamp>.jpg)")
    ap.add_argument("--orbit", type=int, default=0, help="render N frames of an orbit instead of one still")
    ap.add_argument("--degrees-per-frame", type=float, default=3.0)
    ap.add_argument("--out-dir", default="frames")
    ap.add_argument("--mode", default="wavefront", choices=["wavefront", "megakernel", "parity", "preview"])
    ap.add_argument("--tonemapper", default="opendrt", choices=["opendrt", "agx"])
    ap.add_argument("--save-accum", default=None, help=".npz checkpoint of the linear accumulation buffer + spp")
    ap.add_argument("--resume", default=None, help="continue a still from a --save-accum checkpoint: --spp is the new total")
    ap.add_argument("--tile-groups", type=int, default=1, help="multi-GPU still: interleaved 16x8 film-tile groups x sample slices (1 = sample slices only)")
    ap.add_argument("--no-frames", action="store_true", help="orbit: render and resolve every frame but do not write the PNGs (throughput runs)")
    a = ap.parse_args(argv)

    import numpy as np
    import torch
    import torch.distributed as dist
    from . import Renderer, load_config, save_screenshot
    from .distributed import frame_shard, partition, reduce_accumulation

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = (int(x) for x in a.res.split("x"))
    cfg = load_config(a.config)
    r = Renderer((W, H), cfg["up"], textures=parse_textures(a.textures, a.config), device=local, mode=a.mode)
    r.tonemapper = 1 if a.tonemapper == "agx" else 0
    r.copy_textures()

    def render_slice(first, n, keep=False, tiles=None):
        if not keep:
            r.reset_framebuffer()
        done = 0
        while done < n:
            k = min(a.batch, n - done)
            r.accumulate(k, first_sample=first + done, tiles=tiles)
            done += k

    if a.orbit:
        os.makedirs(a.out_dir, exist_ok=True)
        mine, dev_ms = frame_shard(a.orbit, rank, world), 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for f in mine:
            r.apply_config(orbit_config(cfg, math.radians(a.degrees_per_frame * f)))
            e0.record()
            render_slice(0, a.spp)
            img = r.fetch_image(spp=a.spp)
            e1.record()
            if a.no_frames:
                torch.cuda.synchronize()
            else:
                save_screenshot(img, os.path.join(a.out_dir, "frame_%04d.png" % f))
            dev_ms += e0.elapsed_time(e1)
        if mine:
            print("rank %d: %d frames of %dx%d x %d spp, %.1f ms/frame on the device (%.2f frames/s, %.1f M samples/s)"
                  % (rank, len(mine), W, H, a.spp, dev_ms / len(mine), 1e3 * len(mine) / dev_ms, W * H * a.spp * len(mine) / dev_ms / 1e3), flush=True)
        # whole-job figure: frames are independent, the job ends when the slowest rank has rendered its share (device time, max over ranks)
        tot = torch.tensor([dev_ms], device=r.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        if rank == 0:
            import json
            print(json.dumps({"flythrough": {"frames": a.orbit, "res": a.res, "spp": a.spp, "n_gpus": world, "device_ms_max_over_ranks": float(tot.item()),
                                             "frames_per_s": 1e3 * a.orbit / float(tot.item()), "path_samples_per_s": W * H * a.spp * a.orbit / (float(tot.item()) * 1e-3),
                                             "sharding": "frame f -> rank f mod N, no collective"}}), flush=True)
    else:
        r.apply_config(cfg)
        have = 0
        if a.resume:  # rank 0 carries the old sums; every rank renders its share of the NEW samples [have, spp)
            # every rank validates the checkpoint against ITS renderer (resolution, seed, scene, textures, integrator); rank 0 loads it
            have = r.load_accumulation(a.resume) if rank == 0 else r.check_checkpoint(a.resume)
            if have > a.spp:
                raise SystemExit("checkpoint already holds %d spp, --spp %d asks for fewer" % (have, a.spp))
        part = partition(a.spp - have, rank, world, a.tile_groups)
        render_slice(have + part["first_sample"], part["n_spp"], keep=bool(a.resume) and rank == 0,
                     tiles=(part["tile_stride"], part["tile_offset"]) if a.tile_groups > 1 else None)
        reduce_accumulation(r.color_buffer, dst=0)
        if rank == 0:
            r.current_spp = a.spp
            img = r.fetch_image(spp=a.spp)
            save_screenshot(img, a.out)
            if a.save_accum:
                r.save_accumulation(a.save_accum, extra=cfg)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
