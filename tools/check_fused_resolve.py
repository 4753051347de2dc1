"""2+ ranks (torchrun): the fused peer-memory resolve must give the image of ncclReduce + resolve.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_fused_resolve.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digital_earth_b200 as de  # noqa: E402
from digital_earth_b200.distributed import reduce_accumulation, resolve_fused, sample_slice  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, SPP = 256, 128, 16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = de.load_config(os.path.join(ROOT, "digital-earth_b200", "assets", "configs", "config - florida.txt"))
r = de.Renderer((W, H), cfg["up"], textures=de.textures.synthetic(512, 256), device=local)
r.apply_config(cfg)
first, n = sample_slice(SPP, rank, world)
r.reset_framebuffer(); r.accumulate(n, first_sample=first)
fused = resolve_fused(r, SPP, rank, world, dst=0)            # reads the peers' buffers in place
reduce_accumulation(r.color_buffer, dst=0)                   # now sum them with NCCL
if rank == 0:
    ref = r.fetch_image(spp=SPP)
    err = (fused - ref).abs().max().item()
    one = de.Renderer((W, H), cfg["up"], textures=de.textures.synthetic(512, 256), device=local)
    one.apply_config(cfg); one.reset_framebuffer(); one.accumulate(SPP)
    err1 = (one.fetch_image() - ref).abs().max().item()
    print("fused vs nccl max |diff| = %.3g ; %d-GPU vs 1-GPU image max |diff| = %.3g ; mean %.4f" % (err, world, err1, ref.mean().item()), flush=True)
    assert err <= 1e-5 and err1 <= 2e-3 and ref.mean().item() > 0.01
    one.close()
dist.barrier()
r.close()
dist.destroy_process_group()
