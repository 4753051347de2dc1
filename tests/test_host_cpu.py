"""CPU-side checks (no GPU): host logic of the package, C-ABI exports, data assets."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import digital_earth_b200 as de
from digital_earth_b200 import _lib, textures
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "digital-earth_b200", "assets", "configs")


def test_library_loads_and_exports_every_declared_symbol():
    __import__("importlib").import_module("digital_earth_b200.build").build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "de_api.h")).read()
    declared = set(re.findall(r"\b(de_[a-z0-9_A-Z]+)\s*\(", hdr)) - {"de_ctx"}
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    assert lib.de_abi_version() == 1


def test_render_kernel_fits_the_occupancy_it_is_designed_for():
    """The persistent kernel runs 1024 threads per SM next to a 226 KB pool: at most 64 registers per thread (65536 / 1024), a stack of a
    few words at most (spills are shared-memory-speed traffic the design has no room for), compiled for sm_100a."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    __import__("importlib").import_module("digital_earth_b200.build").build()
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    m = re.search(r"Function \S*k_render_wavefrontILb0\S*:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "k_render_wavefront<false> not found in libde.so"
    assert int(m.group(1)) <= 64 and int(m.group(2)) <= 64, m.groups()


def test_create_rejects_bad_resolution_without_gpu_work():
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    assert lib.de_create(ctypes.byref(ctx), 0, 100, 100) == -1  # W%16, H%8 (renderer.py:46)
    assert lib.de_create(None, 0, 64, 32) == -1


def test_renderer_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DeError):
        de.Renderer((64, 32), (0, 1, 0), textures=textures.synthetic(64, 32))


@pytest.mark.parametrize("name", ["Apollo 11", "florida", "sunset hurricane"])
def test_config_roundtrip(tmp_path, name):
    cfg = de.load_config(os.path.join(CFG, "config - %s.txt" % name))
    assert len(cfg["cam_pos"]) == 3 and isinstance(cfg["selected_crf"], int)
    p = tmp_path / "config.txt"
    de.save_config(str(p), cfg)
    assert de.load_config(str(p)) == cfg
    txt = p.read_text()
    assert len(txt.split("\n")) == 10 and not txt.endswith("\n")  # earth_viewer.py:213-222


def test_config_values_match_survey_appendix_b():
    a = de.load_config(os.path.join(CFG, "config - Apollo 11.txt"))
    assert a["selected_crf"] == 12 and abs(np.linalg.norm(a["cam_pos"]) - 57078.7e3) < 1e3
    f = de.load_config(os.path.join(CFG, "config - florida.txt"))
    assert abs(np.linalg.norm(f["cam_pos"]) - 7656.8e3) < 1e3 and abs(f["fov"] - 0.19467) < 1e-5


def test_screenshot_orientation(tmp_path):
    W, H = 32, 16
    img = np.zeros((W, H, 3), np.float32)
    img[0, 0] = (1, 0, 0)          # bottom-left pixel red
    img[W - 1, H - 1] = (0, 0, 2)  # top-right, over-range
    u8 = de.to_uint8_image(img)
    assert u8.shape == (H, W, 3)
    assert tuple(u8[H - 1, 0]) == (255, 0, 0) and tuple(u8[0, W - 1]) == (0, 0, 255)
    out = de.save_screenshot(img, str(tmp_path / "x.png"))
    assert os.path.exists(out)


def test_synthetic_textures_are_deterministic_and_shaped():
    a = textures.synthetic(128, 64, seed=3)
    b = textures.synthetic(128, 64, seed=3)
    for k in textures.SLOTS:
        assert a[k].dtype == np.uint8 and a[k].shape[:2] == (64, 128)
        assert np.array_equal(a[k], b[k])
    assert a["albedo"].shape == (64, 128, 3) and a["clouds"].ndim == 2
    cover = (textures.synthetic(256, 128, cloud_cover=0.8)["clouds"] > 0).mean()
    assert 0.7 < cover < 0.9


def test_image_array_ingest_flips_rows():
    img = np.zeros((4, 8, 3), np.uint8)
    img[0] = 9  # top row of the decoded file
    t = textures.from_image_array(img, rgb=True)
    assert t.shape == (4, 8, 3) and (t[3] == 9).all() and (t[0] == 0).all()
    assert textures.from_image_array(img, rgb=False).shape == (4, 8, 1)


def test_manifest_matches_reference_resolutions():
    assert textures.MANIFEST[2]["albedo"] == ("earth_color_21K.png", (21600, 10800))
    assert textures.MANIFEST[2]["ocean"][1] == (16200, 8100) and textures.MANIFEST[0]["stars"][1] == (8100, 4050)
    with pytest.raises(FileNotFoundError):
        textures.load_directory("/nonexistent")


# ---- known-answer checks on the data assets ----
def test_lut_known_answers(luts):
    cie = luts["cie"]
    assert cie.shape == (2, 441, 3)
    assert (np.diff(cie[0], axis=0) >= 0).all() and np.allclose(cie[0, -1], 1.0)
    assert np.allclose(cie[1].sum(0), 113.042, atol=2e-2)
    crf = luts["crf"]
    assert luts["crf_names"][0] == "Neutral.rf" and crf.shape == (16, 1024, 3)
    ramp = np.arange(1024, dtype=np.float64) / 1023.0
    assert np.abs(crf[0] - ramp[:, None]).max() < 1e-6  # Neutral.rf is the identity
    assert 5e-24 < luts["o3"].min() and luts["o3"].max() < 5.2e-21
    s = luts["srgb2spec"].astype(np.float32).sum(1)
    assert np.abs(s - 1.0).max() < 0.02  # white -> reflectance ~1


def test_oracle_analytic_invariants():
    # appendix A constants
    d0 = orc.density(np.array([0.0, 25000.0], np.float32))
    assert np.allclose(d0[0], [0.9963712, 1.06, 0.0832662], rtol=2e-6) and abs(d0[1, 2] - 1.0) < 1e-6
    sp = orc.spectra(np.array([400.0, 550.0, 830.0], np.float32))
    assert np.allclose(sp[1, :3], [1.167e-5, 2.049e-5, 8.391e-7], rtol=2e-3)
    assert np.allclose(sp[1, 3:], [2.80427e4, 1.87530e2], rtol=1e-4)
    # OpenDRT: 0.18 -> 0.11696, 64 -> 1.0 (OpenDRT.py:306-319)
    o = orc.opendrt(np.array([[0.18] * 3, [64.0] * 3], np.float32))
    assert np.allclose(o[0], 0.11696, atol=2e-6) and np.allclose(o[1], 1.0, atol=1e-6)
    # Klein-Nishina end points
    z = np.array([[0, 0, 1.0]], np.float32)
    assert abs(orc.phase_eval(z, z, [1], [0])[0] - 54.88302) < 1e-3
    assert abs(orc.phase_eval(z, -z, [1], [0])[0] - 9.145647e-3) < 1e-7
    # Neutral CRF is the identity to bilinear accuracy
    tex = textures.synthetic(64, 32)
    s = orc.Scene(tex, 32, 16)
    x = np.linspace(0.01, 0.99, 50, dtype=np.float32)
    assert np.abs(orc.crf(s, np.stack([x, x, x], 1)) - x[:, None]).max() < 2e-3


def test_phase_functions_integrate_to_one():
    # Rayleigh, Klein-Nishina, cloud (HG+Draine) over the sphere: 2*pi*int p(cos) dcos == 1
    mu = np.cos(np.linspace(0, np.pi, 200001))
    a = np.tile(np.array([[0, 0, 1.0]], np.float32), (len(mu), 1))
    b = np.stack([np.sqrt(1 - mu ** 2), np.zeros_like(mu), mu], 1).astype(np.float32)
    for pid, red, tol in ((0, 0, 1e-4), (1, 0, 2e-2), (3, 1, 2e-3), (4, 0, 1e-5)):
        p = orc.phase_eval(a, b, np.full(len(mu), pid), np.full(len(mu), red)).astype(np.float64)
        integral = -2 * np.pi * np.trapezoid(p, mu)
        assert abs(integral - 1.0) < tol, (pid, integral)


def test_samplers_match_their_pdfs():
    # importance samplers have weight 1 => the histogram of cos(theta) must match the phase pdf
    rng = np.random.default_rng(1)
    n = 200000
    view = np.tile(np.array([[0.3, 0.5, 0.8124038]], np.float32), (n, 1))
    view /= np.linalg.norm(view, axis=1, keepdims=True)
    for pid, red in ((3, 1), (0, 0)):
        rand = rng.integers(0, 2 ** 32, (n, 4), dtype=np.uint64).astype(np.uint32)
        d, w = orc.phase_sample(view, np.full(n, pid), np.full(n, red), rand)
        mu = (d * view).sum(1)
        edges = np.linspace(-1, 1, 21)
        hist = np.histogram(mu, edges, weights=w)[0] / n
        mid = np.linspace(-1, 1, 20001)
        b = np.stack([np.sqrt(1 - mid ** 2), np.zeros_like(mid), mid], 1).astype(np.float32)
        a = np.tile(np.array([[0, 0, 1.0]], np.float32), (len(mid), 1))
        p = orc.phase_eval(a, b, np.full(len(mid), pid), np.full(len(mid), red)).astype(np.float64) * 2 * np.pi
        cdf = np.concatenate([[0], np.cumsum((p[1:] + p[:-1]) * 0.5 * np.diff(mid))])
        want = np.diff(np.interp(edges, mid, cdf))
        assert np.abs(hist - want).max() < 6e-3, (pid, np.abs(hist - want).max())


def test_oracle_multithreaded_render_is_deterministic():
    tex = textures.synthetic(64, 32, seed=7)
    cfg = de.load_config(os.path.join(CFG, "config - florida.txt"))
    p = dict(cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"], aspect_scale=cfg["aspect_scale"],
             sun_angle=cfg["sun_angle"], sun_path_rot=cfg["sun_path_rot"])
    s = orc.Scene(tex, 32, 16, **p)
    a1, c1 = orc.render(s, 2, nthreads=1)
    a4, c4 = orc.render(s, 2, nthreads=4)
    assert np.array_equal(a1, a4) and c1 == c4 and c1["paths"] == 32 * 16 * 2
    b, _ = orc.render(s, 1, first_sample=0, nthreads=4)
    c, _ = orc.render(s, 1, first_sample=1, nthreads=4)
    assert np.allclose(b + c, a1, rtol=1e-6, atol=1e-9)  # sample slices add up (multi-GPU partition property)


def test_orbit_rotates_about_the_polar_axis():
    from digital_earth_b200.render import orbit_config
    cfg = de.load_config(os.path.join(CFG, "config - florida.txt"))
    o = orbit_config(cfg, np.radians(90.0))
    assert abs(np.linalg.norm(o["cam_pos"]) - np.linalg.norm(cfg["cam_pos"])) < 1e-3
    assert abs(o["cam_pos"][1] - cfg["cam_pos"][1]) < 1e-9            # latitude kept
    assert abs(np.dot(o["cam_pos"], cfg["cam_pos"]) - cfg["cam_pos"][1] ** 2) < 1e-3 * np.linalg.norm(cfg["cam_pos"]) ** 2  # 90 deg in the equatorial plane
    back = orbit_config(o, np.radians(-90.0))
    assert np.allclose(back["cam_pos"], cfg["cam_pos"]) and np.allclose(back["look_at"], cfg["look_at"])


def test_lut_provenance(luts):
    """The packed LUTs are byte-identical to what the reference's own generator scripts produce
    (verdict recorded by tests/golden/check_lut_provenance.py) and unchanged since (sha256)."""
    import hashlib
    import json
    prov = json.load(open(os.path.join(ROOT, "tests", "golden", "lut_provenance.json")))
    assert prov["ozone_regenerated_bit_exact"] and prov["srgb2spec_regenerated_bit_exact"]
    for k in ("cie", "srgb2spec", "o3", "crf"):
        assert hashlib.sha256(np.ascontiguousarray(luts[k]).tobytes()).hexdigest() == prov["sha256_" + k], k


def test_oracle_event_fixture_feeds_the_roofline_model():
    """bench.py's FLOP model uses the oracle's event counts of each bench view (profiles/oracle_events.json)."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    ev = json.load(open(os.path.join(ROOT, "profiles", "oracle_events.json")))
    for key, lo, hi in (("apollo_8192x4096", 5e3, 8e3), ("florida_2048x1024", 1.4e4, 2.2e4), ("sunset_8192x4096", 2.8e4, 4.2e4)):
        assert lo < bench.flop_per_path(ev[key]) < hi, key

    class A:
        scene, tex = "apollo", "8192x4096"
    assert bench.oracle_events(A)["paths"] > 0
    A.tex = "123x45"
    assert bench.oracle_events(A) is None


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours) on a tiny sample: one JSON line with the contract keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--tex", "256x128",
                          "--cpu-res", "96x64", "--cpu-spp", "1"], capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    # ranks other than 0 of a torchrun launch print nothing and exit 0
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=60,
                         env={**os.environ, "RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""
