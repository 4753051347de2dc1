"""Oracle frames of the 4096-spp image gate (tests/test_gpu_render.py::test_4096spp_rel_rmse_wavefront_vs_oracle), frozen.

    python tests/golden/gen_oracle_frames.py [view ...]        ->  tests/golden/oracle_frames_4096spp_v1.npz

For each shipped view: the CPU oracle's (oracle/de_oracle.c, pinned bit-exact to the reference's source) 128x64 x 4096-spp
accumulation buffer and per-pixel sums of squares on the synthetic 256x128 textures of the GPU render tests, seed 4242.  The render
is deterministic (Philox keyed by pixel and sample, one thread per pixel row), so the file is reproducible bit for bit:
tests/test_oracle_frames_cpu.py re-renders one view live and compares.  Why frozen: the three renders cost the GPU box's 16 host
cores ~4.5 minutes of every `pytest -m gpu` run; `DE_LIVE_ORACLE=1` makes the GPU test render them live again.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "oracle_frames_4096spp_v1.npz")
VIEWS = ("Apollo 11", "florida", "sunset hurricane")
W, H, TW, TH, SPP, SEED = 128, 64, 256, 128, 4096, 4242


def key(view):
    return view.replace(" ", "_")


def scene_for(view):
    import digital_earth_b200 as de
    from oracle import oracle as orc
    tex = de.textures.synthetic(TW, TH, cloud_cover=0.6, seed=3)
    cfg = de.load_config(os.path.join(ROOT, "digital-earth_b200", "assets", "configs", "config - %s.txt" % view))
    return orc, orc.Scene(tex, W, H, cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"], aspect_scale=cfg["aspect_scale"],
                          exposure=cfg["exposure"], selected_crf=cfg["selected_crf"], gamma=cfg["gamma"], sun_angle=cfg["sun_angle"],
                          sun_path_rot=cfg["sun_path_rot"])


def render(view):
    orc, s = scene_for(view)
    acc, acc2, _ = orc.render(s, SPP, seed=SEED, second_moment=True)
    return np.asarray(acc, np.float32), np.asarray(acc2, np.float32)


def load(view):
    """(acc, acc2) of `view` from the committed file, or None."""
    if not os.path.exists(OUT):
        return None
    z = np.load(OUT)
    k = key(view)
    if k + "_acc" not in z.files or tuple(z["params"]) != (W, H, TW, TH, SPP, SEED):
        return None
    return z[k + "_acc"], z[k + "_acc2"]


if __name__ == "__main__":
    have = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for view in sys.argv[1:] or VIEWS:
        acc, acc2 = render(view)
        have[key(view) + "_acc"], have[key(view) + "_acc2"] = acc, acc2
        print("%-18s mean %.8g  sum of squares %.8g" % (view, acc.mean() / SPP, acc2.sum()), flush=True)
    have["params"] = np.array([W, H, TW, TH, SPP, SEED], np.int64)
    np.savez(OUT, **have)
    print("wrote", OUT)
