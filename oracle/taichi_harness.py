"""Real-Taichi cross-check of the oracle (SURVEY.md 8c deliverable 3, BASELINE.md "Baseline B").  TEST INFRASTRUCTURE ONLY.

The oracle (oracle/de_oracle.c) is pinned to the reference's SOURCE executed on a Taichi stand-in; three assumptions about
the Taichi RUNTIME itself cannot be observed without Taichi:

  1. `for x in range(0, log2(res))` (lib/colour.py:26) runs 8 iterations (the f32 bound 8.78 is cast to i32);
  2. `Texture.sample_lod` (lib/math_utils.py:44, lib/colour.py:28,40,41) clamps to the edge at u = 1.0 and filters with
     fp32 weights between texel centres (i + 0.5) / N;
  3. `fast_math` moves exp / log / pow / sin / cos by no more than an ulp or two.

This module runs the UNMODIFIED reference (`/root/reference/renderer.py`, `pathtracer.py`, `lib/*.py`) through a real Taichi
install when one exists, headless, on the synthetic textures of the tests:

  * `available()`            -> (ok, reason); ok only if `taichi` imports AND a reference checkout is present
  * `probe_semantics()`      -> the three assumptions above measured with tiny Taichi kernels
  * `TaichiReference(...)`   -> the reference's `Renderer` on synthetic maps; `.render(cfg, spp)` returns the accumulation
                                 buffer in this repo's [H][W][3] layout and the wall time; `.fetch_image()` the tonemapped frame

Arch: `ti.cpu` first (the baseline BASELINE.json names).  Taichi's CPU backend has no texture support in the releases of the
reference's era; if constructing the reference's `ti.Texture`s fails there, the harness falls back to a GPU arch (`ti.cuda`,
then `ti.vulkan`) and says so in `.arch` -- a field-based re-implementation of `sample_lod` would no longer be the reference's
sampler, which is exactly what item 2 wants to observe.

Nothing here is imported by the product.  Neither Taichi nor the reference exist in the build container or on the GPU box of
this project, so tests/test_taichi_harness_cpu.py skips there; it is the hook for whoever has both.
"""
import importlib
import importlib.util
import os
import shutil
import sys
import tempfile
import time
import types

import numpy as np

REF = os.environ.get("DE_REFERENCE", "/root/reference")
_SLOT_FILES = {"albedo": "albedo.png", "topography": "topography.png", "ocean": "ocean.png", "clouds": "clouds.png",
               "bathymetry": "bathymetry.png", "emissive": "emissive.png", "stars": "stars.png"}


def available():
    if importlib.util.find_spec("taichi") is None:
        return False, "taichi is not installed"
    if not os.path.isfile(os.path.join(REF, "pathtracer.py")):
        return False, "no reference checkout at %s (set DE_REFERENCE)" % REF
    return True, "taichi %s, reference at %s" % (importlib.import_module("taichi").__version__, REF)


def _init(ti, arch_names, cores):
    last = None
    for name in arch_names:
        try:
            ti.init(arch=getattr(ti, name), cpu_max_num_threads=cores, default_fp=ti.f32, offline_cache=False, log_level=ti.WARN)
            return name
        except Exception as e:  # arch not built in / no device
            last = e
    raise RuntimeError("no usable Taichi arch among %s: %s" % (arch_names, last))


def probe_semantics(arch="cpu"):
    """Measure the three runtime assumptions (module docstring) with throw-away kernels.  Returns a dict; needs only Taichi."""
    import taichi as ti
    used = _init(ti, [arch], os.cpu_count() or 1)
    out = {"arch": used}

    count = ti.field(ti.i32, shape=())

    @ti.kernel
    def loop_bound(res: ti.f32):
        n = 0
        for x in range(0, ti.math.log2(res)):      # lib/colour.py:26 with res = 441
            n += 1
        count[None] = n
    loop_bound(441.0)
    out["bisection_iterations"] = int(count[None])  # the oracle assumes 8

    libm = ti.field(ti.f32, shape=5)

    @ti.kernel
    def transcendental(x: ti.f32):
        libm[0] = ti.exp(-x); libm[1] = ti.log(x); libm[2] = ti.pow(x, 1.5); libm[3] = ti.sin(x); libm[4] = ti.cos(x)
    transcendental(1.2345678)
    ref = np.array([np.exp(np.float32(-1.2345678)), np.log(np.float32(1.2345678)), np.float32(1.2345678) ** np.float32(1.5),
                    np.sin(np.float32(1.2345678)), np.cos(np.float32(1.2345678))], np.float32)
    out["libm_rel_err"] = [float(abs(a - b) / abs(b)) for a, b in zip(libm.to_numpy(), ref)]

    try:  # sample_lod: address mode at u = 1.0 and filter weights between centres
        tex = ti.Texture(ti.Format.r32f, (4, 2))
        src = ti.field(ti.f32, shape=(4, 2))
        src.from_numpy(np.array([[0.0, 10.0], [1.0, 11.0], [2.0, 12.0], [3.0, 13.0]], np.float32))
        res = ti.field(ti.f32, shape=4)

        @ti.kernel
        def fill(t: ti.types.rw_texture(num_dimensions=2, fmt=ti.Format.r32f, lod=0)):
            for i, j in src:
                t.store(ti.Vector([i, j]), ti.Vector([src[i, j], 0.0, 0.0, 0.0]))

        @ti.kernel
        def fetch(t: ti.types.texture(num_dimensions=2)):
            res[0] = t.sample_lod(ti.Vector([1.0, 0.25]), 0.0).x      # clamp: 3.0; repeat: 1.5
            res[1] = t.sample_lod(ti.Vector([0.5, 0.25]), 0.0).x      # halfway between centres 1 and 2: 1.5
            res[2] = t.sample_lod(ti.Vector([0.4375, 0.25]), 0.0).x   # x = 1.25 texels: 1.25 with fp32 weights
            res[3] = t.sample_lod(ti.Vector([0.0, 0.25]), 0.0).x      # clamp: 0.0; repeat: 1.5
        fill(tex); fetch(tex)
        v = res.to_numpy()
        out["sample_lod"] = {"u=1": float(v[0]), "u=0.5": float(v[1]), "u=0.4375": float(v[2]), "u=0": float(v[3]),
                             "clamp_to_edge": bool(abs(v[0] - 3.0) < 1e-6 and abs(v[3]) < 1e-6), "weight_error": float(abs(v[2] - 1.25))}
    except Exception as e:
        out["sample_lod"] = {"unsupported_on_this_arch": "%s: %s" % (type(e).__name__, e)}
    return out


class TaichiReference:
    """The reference's own `Renderer` (renderer.py:16) on synthetic maps, headless."""

    def __init__(self, textures, image_res, archs=("cpu", "cuda", "vulkan"), cores=None):
        ok, why = available()
        if not ok:
            raise RuntimeError(why)
        import taichi as ti
        from PIL import Image
        self.ti = ti
        self.cores = cores or os.cpu_count() or 1
        self.work = tempfile.mkdtemp(prefix="de_taichi_")
        # the reference opens textures/, LUT/ relative to the working directory (lib/textures.py:10-31, renderer.py:149)
        os.makedirs(os.path.join(self.work, "textures"))
        os.symlink(os.path.join(REF, "LUT"), os.path.join(self.work, "LUT"))
        consts = {}
        for slot, fname in _SLOT_FILES.items():
            a = np.ascontiguousarray(textures[slot])
            img = a[::-1] if a.ndim == 2 else a[::-1, :, :3]          # our rows run south -> north; image files top row first
            if a.ndim == 2:
                img = np.repeat(img[:, :, None], 3, axis=2)            # the reference reads channel 0 of an RGB file (renderer.py:68)
            Image.fromarray(img).save(os.path.join(self.work, "textures", fname))
            consts[slot] = ("textures/" + fname, (a.shape[1], a.shape[0]))
        # a stand-in for lib/textures.py holding OUR file names and resolutions; everything else is the reference's module namespace
        spec = importlib.util.spec_from_file_location("_ref_textures", os.path.join(REF, "lib", "textures.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for slot, prefix in (("albedo", "ALBEDO"), ("topography", "TOPOGRAPHY"), ("ocean", "OCEAN"), ("clouds", "CLOUDS"), ("bathymetry", "BATHYMETRY"),
                             ("emissive", "EMISSIVE"), ("stars", "STARS")):
            setattr(mod, prefix + "_TEX_FILE", consts[slot][0])
            setattr(mod, prefix + "_TEX_RES", consts[slot][1])
        self._cwd = os.getcwd()
        os.chdir(self.work)
        sys.path.insert(0, REF)
        err = None
        for arch in archs:
            try:
                self.arch = _init(ti, [arch], self.cores)
                for name in [m for m in sys.modules if m == "renderer" or m == "pathtracer" or m.startswith("lib.") or m == "lib"]:
                    del sys.modules[name]
                pkg = types.ModuleType("lib"); pkg.__path__ = [os.path.join(REF, "lib")]
                sys.modules["lib"] = pkg
                sys.modules["lib.textures"] = mod
                ref_renderer = importlib.import_module("renderer")
                self.renderer = ref_renderer.Renderer(image_res=tuple(image_res), up=(0.0, 1.0, 0.0))
                self.renderer.copy_textures()
                err = None
                break
            except Exception as e:     # e.g. ti.Texture unsupported on the CPU backend
                err = "%s on ti.%s: %s" % (type(e).__name__, arch, e)
                ti.reset()
        if err:
            self.close()
            raise RuntimeError("the reference Renderer could not be constructed: " + err)
        self.image_res = tuple(image_res)

    def apply_config(self, cfg):
        r = self.renderer
        r.set_camera_pos(*cfg["cam_pos"]); r.set_look_at(*cfg["look_at"]); r.set_up(*cfg["up"])
        r.set_fov(cfg["fov"]); r.set_aspect_scale(cfg["aspect_scale"]); r.set_exposure(cfg["exposure"])
        r.set_crf(cfg["selected_crf"]); r.set_gamma(cfg["gamma"]); r.set_sun_angle(cfg["sun_angle"]); r.set_sun_path_rot(cfg["sun_path_rot"])

    def render(self, cfg, spp):
        """reset + spp x accumulate() (the reference's only mode: one sample per launch); returns ([H][W][3] sums, seconds)."""
        self.apply_config(cfg)
        r = self.renderer
        r.reset_framebuffer()
        r.accumulate(); self.ti.sync()             # compile outside the timed region
        r.reset_framebuffer()
        t0 = time.perf_counter()
        for _ in range(spp):
            r.accumulate()
        self.ti.sync()
        dt = time.perf_counter() - t0
        return np.ascontiguousarray(r.color_buffer.to_numpy().transpose(1, 0, 2)), dt

    def fetch_image(self):
        return np.ascontiguousarray(self.renderer.fetch_image().to_numpy().transpose(1, 0, 2))

    def close(self):
        try:
            os.chdir(self._cwd)
        except Exception:
            pass
        shutil.rmtree(self.work, ignore_errors=True)
        if REF in sys.path:
            sys.path.remove(REF)
