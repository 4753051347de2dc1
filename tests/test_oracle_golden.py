"""The CPU oracle (oracle/de_oracle.c) must reproduce, BIT FOR BIT, the golden vectors that
tests/golden/gen_golden.py produced by executing the reference's own source files
(/root/reference/{pathtracer,renderer}.py, lib/*.py) on the Taichi stand-in."""
import numpy as np
import pytest

from oracle import oracle as orc


def same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError("%d/%d mismatches; first at %s: oracle=%r golden=%r" % (bad.sum(), bad.size, tuple(i), a[tuple(i)], b[tuple(i)]))


def tiny_scene(g, key=None, W=None, H=None):
    tex = {k: g["tex_" + k] for k in orc.TEX_SLOTS}
    W = int(g["img_res"][0]) if W is None else W
    H = int(g["img_res"][1]) if H is None else H
    p = {}
    if key:
        sc = g["cfg_%s_scalars" % key]
        p = dict(cam_pos=g["cfg_%s_cam_pos" % key], look_at=g["cfg_%s_look_at" % key], up=g["cfg_%s_up" % key], fov=sc[0], aspect_scale=sc[1],
                 exposure=sc[2], selected_crf=int(sc[3]), gamma=sc[4], sun_angle=sc[5], sun_path_rot=sc[6])
    return orc.Scene(tex, W, H, **p)


def test_philox_known_answers(golden):
    # Random123 kat_vectors, Philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for i, (c, k, want) in enumerate(kat):
        assert tuple(orc.philox(c, k)) == want
        assert tuple(int(x) for x in golden["philox_kat"][i]) == want


def test_rsi(golden):
    same(orc.rsi(golden["rsi_pos"], golden["rsi_dir"], golden["rsi_r"]), golden["rsi_out"])


def test_density(golden):
    same(orc.density(golden["density_h"]), golden["density_out"])


def test_spectra(golden):
    same(orc.spectra(golden["spectra_wl"]), golden["spectra_out"])


def test_phase_eval(golden):
    same(orc.phase_eval(golden["phase_a"], golden["phase_b"], golden["phase_id"], golden["phase_reduce"]), golden["phase_eval_out"])


def test_phase_sample(golden):
    d, w = orc.phase_sample(golden["phase_a"], golden["phase_id"], golden["phase_reduce"], golden["phase_rand"])
    same(d, golden["phase_sample_dir"])
    same(w, golden["phase_sample_w"])


def test_direction_samplers(golden):
    same(orc.dir_sample(0, golden["dirs_n"], float(golden["dirs_cmax"]), golden["dirs_rand"]), golden["dirs_cone_out"])
    same(orc.dir_sample(1, golden["dirs_n"], 0.0, golden["dirs_rand"]), golden["dirs_hemi_out"])


def test_brdf(golden):
    g = golden
    same(orc.brdf(g["brdf_albedo"], g["brdf_ocean"], g["brdf_bathy"], g["brdf_v"], g["brdf_n"], g["brdf_l"]), g["brdf_out"])


def test_srgb_to_spectrum(golden):
    same(orc.srgb_to_spectrum(golden["s2s_rgb"], golden["s2s_wl"]), golden["s2s_out"])


def test_spectrum_sample(golden):
    same(orc.spectrum_sample(golden["specsample_rand"]), golden["specsample_out"])


def test_tex_fetch(golden):
    same(orc.tex_fetch(golden["tex_clouds"], golden["texfetch_pos"]), golden["texfetch_r8_out"])
    same(orc.tex_fetch(golden["tex_albedo"], golden["texfetch_pos"]), golden["texfetch_rgb8_out"])


@pytest.mark.parametrize("key", ["apollo", "florida", "sunset"])
def test_cast_dir(golden, key):
    s = tiny_scene(golden, key)
    same(orc.cast_dir(s, golden["cast_%s_u" % key], golden["cast_%s_v" % key], golden["cast_%s_rand" % key]), golden["cast_%s_out" % key])


def test_opendrt(golden):
    same(orc.opendrt(golden["tm_rgb"]), golden["opendrt_out"])


def test_agx(golden):
    same(orc.agx(golden["tm_rgb"]), golden["agx_out"])


@pytest.mark.parametrize("sel", [0, 5, 12])
def test_camera_response(golden, sel):
    s = tiny_scene(golden)
    s.s.selected_crf = sel
    same(orc.crf(s, golden["crf_rgb"]), golden["crf_out_%d" % sel])


def test_srgb_oetf(golden):
    same(orc.srgb_oetf(golden["oetf_in"]), golden["oetf_out"])


def test_resolve(golden):
    s = tiny_scene(golden, "apollo")
    same(orc.resolve(s, golden["resolve_accum"], int(golden["resolve_samples"])), golden["resolve_out"])


def test_intersect_land(golden):
    s = tiny_scene(golden)
    same(orc.intersect_land(s, golden["geo_pos"], golden["geo_dir"]), golden["intersect_land_out"])


def test_land_normal_and_material(golden):
    s = tiny_scene(golden)
    same(orc.land_normal(s, golden["surf_pos"]), golden["land_normal_out"])
    same(orc.land_material(s, golden["surf_pos"]), golden["land_material_out"])


def test_cloud_limits_and_density(golden):
    s = tiny_scene(golden)
    same(orc.cloud_limits(golden["cloud_pos"], golden["cloud_dir"], golden["cloud_land"]), golden["cloud_limits_out"])
    same(orc.clouds_density(s, golden["cloud_pos"]), golden["clouds_density_out"])


def test_fixed_ray_transmittance(golden):
    same(orc.raymarch_T(golden["rm_pos"], golden["rm_dir"], golden["rm_ext"]), golden["rm_out"])


def test_tracking(golden):
    g = golden
    s = tiny_scene(g)
    same(orc.tracking(0, s, g["trk_pos"], g["trk_dir"], g["trk_land"], g["trk_wl"], int(g["trk_seed"])), g["trk_interaction_out"])
    same(orc.tracking(1, s, g["trk_pos"], g["trk_dir"], g["trk_land"], g["trk_wl"], int(g["trk_seed"])), g["trk_transmittance_out"])


@pytest.mark.parametrize("key", ["apollo", "florida", "sunset"])
def test_full_paths(golden, key):
    g = golden
    s = tiny_scene(g, key)
    out = orc.trace_paths(s, g["path_%s_px" % key], g["path_%s_py" % key], g["path_%s_sample" % key], int(g["path_seed"]))
    same(out, g["path_%s_out" % key])


# ---- the deterministic preview integrator: the reference's ray_marcher, executed on the stand-in
def test_preview_marching_routines(golden, golden_preview):
    gp = golden_preview
    s = tiny_scene(golden, "florida")
    atmos, T = orc.ray_march(s, gp["rm_pos"], gp["rm_dir"], gp["rm_t0"], gp["rm_t1"], gp["rm_sun"], gp["rm_wl"])
    same(atmos, gp["rm_atmos_out"])   # pathtracer.py:501-541
    same(T, gp["rm_T_out"])           # pathtracer.py:471-499


@pytest.mark.parametrize("key", ["apollo", "florida", "sunset"])
def test_preview_full_samples(golden, golden_preview, key):
    gp = golden_preview
    out = orc.trace_paths(tiny_scene(golden, key), gp["prev_%s_px" % key], gp["prev_%s_py" % key], gp["prev_%s_sample" % key], int(gp["prev_seed"]),
                          integrator="ray_marcher")
    same(out, gp["prev_%s_out" % key])
    assert (gp["prev_%s_out" % key][:, 4] > 0).sum() >= 20   # the vectors are not all space background
