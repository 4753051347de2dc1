"""Generate tests/golden/golden_preview_v1.npz by executing the reference's `ray_marcher`
(pathtracer.py:471-685, unreferenced upstream) on the Taichi stand-in -- same method, textures, image size
and configs as gen_golden.py, whose helpers are reused.  Build container only:
    python tests/golden/gen_golden_preview.py

RNG contract of the preview: the whole path draws from ONE stream (bounce key 1); the light-cone sample
(pathtracer.py:575) and the hemisphere sample (:620) each start on a multiple of 4.  Bounce key 0 carries
the wavelength and the pixel jitter exactly as for the path tracer (renderer.py:310-313).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (sets up the shim + reference imports)

ti, pt, colour, volume = gg.ti, gg.pt, gg.colour, gg.volume
f32 = np.float32


def main():
    t_all = time.time()
    rng = np.random.default_rng(20261018)
    G = {}
    luts = dict(np.load(os.path.join(gg.ROOT, "digital-earth_b200", "assets", "luts.npz")))
    tex = gg.synth.make_textures(gg.TEX_W, gg.TEX_H, cloud_cover=0.6, seed=7)   # the textures of golden_v1.npz
    T = {k: gg.shim_tex(tex[k], ti.Format.rgba8 if k in ("albedo", "stars") else ti.Format.r8) for k in tex}
    pt.TOPOGRAPHY_TEX_RES = (gg.TEX_W, gg.TEX_H)
    R = gg.make_renderer(gg.IMG_W, gg.IMG_H, luts)
    cfgs = {n: gg.load_config(os.path.join(gg.REF, "config - %s.txt" % n)) for n in ("Apollo 11", "florida", "sunset hurricane")}

    # ---- ray_marh_atmos / ray_march_transmittance on explicit rays (pathtracer.py:471-541)
    n = 48
    PR = 6371e3
    pos = gg.unit(rng, n) * (PR + rng.random((n, 1)) ** 2 * 150e3).astype(np.float32)
    pos = pos.astype(np.float32)
    dirs = gg.unit(rng, n)
    sun = gg.unit(rng, n)
    sun[: n // 2] = (pos[: n // 2] / np.linalg.norm(pos[: n // 2], axis=1, keepdims=True) + 0.6 * gg.unit(rng, n // 2)).astype(np.float32)
    sun = (sun / np.linalg.norm(sun, axis=1, keepdims=True)).astype(np.float32)
    t0 = (rng.random(n) * 2e4).astype(np.float32)
    t1 = (t0 + 1e3 + rng.random(n) ** 2 * 8e5).astype(np.float32)
    wl = np.array([f32(390.0) + f32(441.0) * f32((2 * int(k) + 1) / 512.0) for k in rng.integers(0, 256, n)], dtype=np.float32)
    o2, oT = [], []
    for i in range(n):
        ext = gg.vec3(volume.spectra_extinction_rayleigh(f32(wl[i])), volume.spectra_extinction_mie(f32(wl[i])),
                      volume.spectra_extinction_ozone(f32(wl[i]), R.O3_crossec_LUT_buff))
        scat = gg.vec2(ext.x * volume.rayleigh_albedo, ext.y * volume.aerosol_albedo)   # pathtracer.py:563
        ins, tr = pt.ray_marh_atmos(gg.V(pos[i]), gg.V(dirs[i]), f32(t0[i]), f32(t1[i]), gg.V(sun[i]), ext, scat, T["clouds"])
        o2.append([float(ins), float(tr)])
        oT.append(float(pt.ray_march_transmittance(gg.V(pos[i]), gg.V(sun[i]), ext)))
    G["rm_pos"], G["rm_dir"], G["rm_sun"], G["rm_t0"], G["rm_t1"], G["rm_wl"] = pos, dirs, sun, t0, t1, wl
    G["rm_atmos_out"], G["rm_T_out"] = np.array(o2, np.float32), np.array(oT, np.float32)

    # ---- whole preview samples: Renderer.render's prologue (renderer.py:305-314) + ray_marcher + :329-330
    cur = {}
    saved = gg.install_contract_hooks(cur)
    n_per = int(os.environ.get("DE_GOLDEN_PATHS", "64"))  # the committed fixture holds 64 per view
    for cname, cfg in cfgs.items():
        key = cname.split()[0].lower()
        gg.apply_config(R, cfg)
        sp = gg.scene_params_of(R)
        px = rng.integers(0, gg.IMG_W, n_per); py = rng.integers(0, gg.IMG_H, n_per); sm = rng.integers(0, 4, n_per)
        outs = []
        t0_ = time.time()
        for i in range(n_per):
            s_ = gg.Stream(5, int(py[i]) * gg.IMG_W + int(px[i]), int(sm[i])); cur["s"] = s_; ti._set_random_source(s_)
            wavelength, response, rcp = colour.spectrum_sample(R.CIE_LUT_tex, 441)
            pp = gg.PathParameters()
            pp.wavelength = wavelength
            pp.ray_dir = R.get_cast_dir(ti.I32(int(px[i])), ti.I32(int(py[i])))
            pp.ray_pos = R.camera_pos[None]
            s_.set_bounce(1)
            L = pt.ray_marcher(pp, sp, T["albedo"], T["topography"], T["ocean"], T["clouds"], T["bathymetry"], T["emissive"], T["stars"],
                               R.srgb_to_spectrum_buff, R.O3_crossec_LUT_buff)
            rgbc = colour.xyzToRGBMatrix_D65 @ (L * response * rcp)
            outs.append([*gg.arr(rgbc), wavelength, L])
        print("preview %-8s %d in %.1fs" % (key, n_per, time.time() - t0_), flush=True)
        G["prev_%s_px" % key], G["prev_%s_py" % key], G["prev_%s_sample" % key] = px.astype(np.int32), py.astype(np.int32), sm.astype(np.uint32)
        G["prev_%s_out" % key] = np.array(outs, np.float32)
    for name, fn in saved.items():
        setattr(pt, name, fn)
    G["prev_seed"] = np.uint32(5)
    outp = os.environ.get("DE_GOLDEN_PREVIEW_OUT") or os.path.join(HERE, "golden_preview_v1.npz")
    np.savez_compressed(outp, **G)
    print("wrote %s (%d arrays, %.1f kB) in %.1fs" % (outp, len(G), os.path.getsize(outp) / 1e3, time.time() - t_all))


if __name__ == "__main__":
    main()
