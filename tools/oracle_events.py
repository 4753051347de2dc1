"""Event counts per path of the REFERENCE algorithm (the oracle's counters) for the bench workloads -> profiles/oracle_events.json.
They are the N_* of SURVEY.md 8(d)'s work model; bench.py reads them for `roofline.achieved` (and refreshes them from its
cpu_baseline leg when that runs).  Test infrastructure: executes oracle/.
    python tools/oracle_events.py [scene:tex ...]      default: apollo:8192x4096 florida:2048x1024 sunset:8192x4096"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as orc  # noqa: E402


class A:
    pass


def main():
    jobs = sys.argv[1:] or ["apollo:8192x4096", "florida:2048x1024", "sunset:8192x4096"]
    path = os.path.join(ROOT, "profiles", "oracle_events.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for job in jobs:
        scene, tex = job.split(":")
        a = A(); a.scene, a.tex = scene, tex
        tw, th = map(int, tex.split("x"))
        cfg = bench.scene_cfg(a)
        s = orc.Scene(bench.textures_for(a, tw, th), 240, 136, cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"],
                      aspect_scale=cfg["aspect_scale"], sun_angle=cfg["sun_angle"], sun_path_rot=cfg["sun_path_rot"])
        _, cnt = orc.render(s, 4, seed=1)
        out["%s_%s" % (scene, tex)] = {k: int(v) for k, v in cnt.items()}
        print(job, {k: round(v / max(cnt["paths"], 1), 3) for k, v in cnt.items()}, "flop/path %.0f" % bench.flop_per_path(cnt), flush=True)
    out["note"] = "oracle (oracle/de_oracle.c) event counters, 240x136 x 4 spp of each view, seed 1; tools/oracle_events.py"
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
