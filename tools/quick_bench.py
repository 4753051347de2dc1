"""Quick device timing of de_accumulate for the three shipped scenes (development aid)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digital_earth_b200 as de  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", default="640x360")
ap.add_argument("--tex", default="2048x1024")
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--modes", default="megakernel,wavefront")
ap.add_argument("--scenes", default="Apollo 11,florida,sunset hurricane")
ap.add_argument("--count", action="store_true")
a = ap.parse_args()
W, H = map(int, a.res.split("x"))
tw, th = map(int, a.tex.split("x"))
cfgdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "digital-earth_b200", "assets", "configs")
for scene in a.scenes.split(","):
    t0 = time.time()
    tex = de.textures.synthetic(tw, th, cloud_cover=0.8 if "sunset" in scene else 0.5, hurricane="sunset" in scene)
    r = de.Renderer((W, H), (0, 1, 0), textures=tex)
    r.apply_config(de.load_config(os.path.join(cfgdir, "config - %s.txt" % scene)))
    r.copy_textures()
    print("scene %-18s textures %dx%d in %.1fs" % (scene, tw, th, time.time() - t0), flush=True)
    ref = None
    for mode in a.modes.split(","):
        r.set_mode(mode)
        r.set_counting(False)
        r.reset_framebuffer(); r.accumulate(1); torch.cuda.synchronize()
        r.reset_framebuffer()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.accumulate(a.spp); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        acc = r.color_buffer.clone()
        mean = (acc / a.spp).mean().item()
        line = "  %-10s %8.2f ms  %8.1f Mpaths/s  mean %.5f" % (mode, ms, W * H * a.spp / ms / 1e3, mean)
        if ref is None:
            ref = acc
        else:
            rel = ((acc - ref).abs().mean() / ref.abs().mean()).item()
            line += "  mean|diff|/mean vs first %.4f" % rel
        if a.count:
            r.set_counting(True); r.reset_framebuffer(); r.accumulate(a.spp); c = r.counters(); prof = r.stage_profile() if mode == 'wavefront' else None; r.set_counting(False)
            p = max(c["paths"], 1)
            line += "  per path: seg %.2f rmo %.1f cloud %.1f sdf %.1f tex %.1f surf %.2f" % (
                c["segments"] / p, c["rmo_steps"] / p, c["cloud_steps"] / p, c["sdf_evals"] / p, c["tex_fetches"] / p, c["surface_hits"] / p)
        print(line, flush=True)
        if a.count and mode == 'wavefront' and prof:
            tot = sum(v[0] for v in prof.values())
            sw = prof.pop('SWITCHES', (0, 0, 0))[1]
            visits = sum(v[1] for v in prof.values())
            print('    stage share of warp cycles: ' + '  '.join('%s %.1f%% (%.1f slots/visit)' % (k, 100.0 * v[0] / tot, v[2] / max(v[1], 1)) for k, v in prof.items())
                  + '  | %.1f%% of %d visits change the SM\'s stage body' % (100.0 * sw / max(visits, 1), visits), flush=True)
    r.close()
