"""Launch-cost curve and in-kernel timeline of de_accumulate (development aid, round 2).

For each scene: device time of one accumulate(n) for n = 1 ... 256 spp (CUDA events, best of `--reps`), a linear fit
ms = a + b * spp (a = the per-launch constant: ramp + drain), and the kernel's own timeline (option "timeline": first / last
CTA to find the work counter exhausted, first / last CTA end, chunks claimed per CTA) for a few launch sizes.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digital_earth_b200 as de  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", default="1920x1080")
ap.add_argument("--tex", default="8192x4096")
ap.add_argument("--spps", default="1,2,4,8,16,32,64,128")
ap.add_argument("--timeline-spps", default="1,16,128")
ap.add_argument("--scenes", default="Apollo 11,florida,sunset hurricane")
ap.add_argument("--variants", default="space_tiles=0;space_tiles=1,space_async=0;space_tiles=1,space_async=1")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--flush", action="store_true", help="evict L2 before every timed launch")
ap.add_argument("--cta", type=int, default=0, help="print the drain diagnostics of the N slowest CTAs (and the fastest) per timeline launch")
a = ap.parse_args()
W, H = map(int, a.res.split("x"))
tw, th = map(int, a.tex.split("x"))
cfgdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "digital-earth_b200", "assets", "configs")
spps = [int(x) for x in a.spps.split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if a.flush else None
for scene in a.scenes.split(","):
    tex = de.textures.synthetic(tw, th, cloud_cover=0.8 if "sunset" in scene else 0.5, hurricane="sunset" in scene)
    r = de.Renderer((W, H), (0, 1, 0), textures=tex)
    r.apply_config(de.load_config(os.path.join(cfgdir, "config - %s.txt" % scene)))
    r.copy_textures()
    print("== %s  %dx%d  textures %dx%d" % (scene, W, H, tw, th), flush=True)
    for variant in a.variants.split(";"):
        opts = dict(kv.split("=") for kv in variant.split(",") if kv)
        for k, v in opts.items():
            r.set_option(k, int(v))
        r.set_option("timeline", 0)
        r.reset_framebuffer(); r.accumulate(2); torch.cuda.synchronize()
        ms = []
        for n in spps:
            best = 1e30
            for _ in range(a.reps):
                r.reset_framebuffer()
                if flush is not None:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r.accumulate(n); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ms.append(best)
        big = [i for i, n in enumerate(spps) if n >= 8]
        b, c = np.polyfit([spps[i] for i in big], [ms[i] for i in big], 1)
        print("  [%s]  " % variant + "  ".join("%dspp %.2fms" % (n, m) for n, m in zip(spps, ms)) + "   fit(>=8spp): %.3f ms/spp + %.2f ms  -> %.1f Mpaths/s asymptotic"
              % (b, c, W * H / b / 1e3), flush=True)
        r.set_option("timeline", 1)
        r.set_counting(True)        # the timeline is recorded by the instrumented build of the kernel
        for n in [int(x) for x in a.timeline_spps.split(",")]:
            r.reset_framebuffer()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r.accumulate(n); e1.record(); torch.cuda.synchronize()
            t = r.launch_timeline()
            print("    timeline %4d spp: event %.2f ms | counter exhausted %.2f..%.2f ms, CTAs end %.2f..%.2f ms, chunks/CTA %d..%d, tiles wf %d space %d"
                  % (n, e0.elapsed_time(e1), t["first_exhaust_ms"], t["last_exhaust_ms"], t["first_cta_end_ms"], t["last_cta_end_ms"], t["min_chunks_per_cta"],
                     t["max_chunks_per_cta"], t["wavefront_tiles"], t["space_tiles"]), flush=True)
            if a.cta:
                ct = sorted(r.cta_timeline(), key=lambda c: -c["end_ms"])
                for c in ct[:a.cta] + ct[-1:]:
                    print("      CTA end %.2f ms (exhaust %.2f, <64 alive at %.2f): visits after exhaustion " % (c["end_ms"], c["exhaust_ms"], c["few_ms"])
                          + " ".join("%s %d/%d" % (k, c["visits"][k], c["slots"][k]) for k in c["visits"]), flush=True)
        r.set_option("timeline", 0)
        r.set_counting(False)
    r.close()
