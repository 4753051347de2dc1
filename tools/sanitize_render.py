"""One small wavefront render per shipped view, for compute-sanitizer (tools/r2_sanitize.sh): 64x32 pixels, 4 spp, 256x128 textures,
space tiles + second moments + counters on, then a resolve.  Prints a checksum so a silent no-op is visible."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digital_earth_b200 as de  # noqa: E402

cfgdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "digital-earth_b200", "assets", "configs")
tex = de.textures.synthetic(256, 128, cloud_cover=0.6, seed=3)
for scene in sys.argv[1:] or ["Apollo 11", "florida", "sunset hurricane"]:
    r = de.Renderer((64, 32), (0, 1, 0), textures=tex)
    r.apply_config(de.load_config(os.path.join(cfgdir, "config - %s.txt" % scene)))
    r.set_option("moments", 1)
    r.reset_framebuffer(); r.accumulate(4)
    img = r.fetch_image()
    print("%-18s accum sum %.6g  moment2 sum %.6g  image mean %.4f" % (scene, float(r.color_buffer.sum()), float(r.moment2.sum()), float(img.mean())), flush=True)
    r.close()
