"""ctypes front-end of the CPU oracle (oracle/de_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libde_oracle.so")
_ASSETS = os.path.join(os.path.dirname(_HERE), "digital-earth_b200", "assets")

TEX_SLOTS = ("albedo", "topography", "ocean", "clouds", "bathymetry", "emissive", "stars")


def build(force=False):
    src = os.path.join(_HERE, "de_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class OrcTex(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("c", C.c_int32)]


class OrcScene(C.Structure):
    _fields_ = [
        ("tex", OrcTex * 7),
        ("cie", C.c_void_p), ("srgb2spec", C.c_void_p), ("o3", C.c_void_p), ("crf", C.c_void_p),
        ("n_crf", C.c_int32),
        ("cam_pos", C.c_float * 3), ("look_at", C.c_float * 3), ("up", C.c_float * 3),
        ("fov", C.c_float), ("aspect_scale", C.c_float), ("sun_angle", C.c_float), ("sun_path_rot", C.c_float),
        ("land_height_scale", C.c_float), ("exposure", C.c_float), ("gamma", C.c_float),
        ("selected_crf", C.c_int32), ("crf_count", C.c_int32),
        ("vig_strength", C.c_float), ("vig_radius", C.c_float), ("vig_cx", C.c_float), ("vig_cy", C.c_float),
        ("tonemapper", C.c_int32), ("topo_tex_w", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
    ]


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("paths", "segments", "rmo_steps", "cloud_steps", "sdf_evals", "tex_fetches", "surface_hits", "rng_draws")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_abi_version.restype = C.c_int
    return _lib


def load_luts():
    z = np.load(os.path.join(_ASSETS, "luts.npz"))
    return {k: z[k] for k in ("cie", "srgb2spec", "o3", "crf")} | {"crf_names": [str(s) for s in z["crf_names"]]}


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """Owns the numpy buffers an orc_scene points to."""

    def __init__(self, textures, W, H, luts=None, **params):
        luts = luts or load_luts()
        self.keep = []
        s = OrcScene()
        for i, name in enumerate(TEX_SLOTS):
            t = np.ascontiguousarray(textures[name], dtype=np.uint8)
            if t.ndim == 2:
                t = t[:, :, None]
            self.keep.append(t)
            s.tex[i].data = t.ctypes.data
            s.tex[i].h, s.tex[i].w, s.tex[i].c = t.shape
        self.cie = _f32(luts["cie"])
        self.s2s = np.ascontiguousarray(luts["srgb2spec"], dtype=np.float16)
        self.o3 = _f32(luts["o3"])
        self.crf = _f32(luts["crf"])
        s.cie, s.srgb2spec, s.o3, s.crf = self.cie.ctypes.data, self.s2s.ctypes.data, self.o3.ctypes.data, self.crf.ctypes.data
        s.n_crf = self.crf.shape[0]
        s.crf_count = self.crf.shape[0]
        s.W, s.H = W, H
        s.topo_tex_w = self.keep[1].shape[1]
        # renderer.py:20-22,49-58 defaults
        d = dict(cam_pos=(-1.5e7, 0.0, 1.5e7), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=float(np.radians(27.0) * 0.5),
                 aspect_scale=1.0, sun_angle=float(np.radians(60.0)), sun_path_rot=float(np.radians(-45.0)),
                 land_height_scale=7800.0, exposure=2.5, gamma=1.0, selected_crf=0,
                 vig_strength=0.9, vig_radius=0.0, vig_cx=0.5, vig_cy=0.5, tonemapper=0)
        d.update(params)
        for k, v in d.items():
            if k in ("cam_pos", "look_at", "up"):
                getattr(s, k)[:] = [float(x) for x in v]
            else:
                setattr(s, k, v)
        self.s = s

    @property
    def ref(self):
        return C.byref(self.s)


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def _call(name, n, out_shape, *args):
    """Batch entry points are plain loops over n items; big batches are cut into row blocks run on host threads
    (ctypes releases the GIL).  Per-item results do not depend on the cut: no item reads another's state."""
    out = np.zeros(out_shape, dtype=np.float32)
    fn = getattr(lib(), name)
    per_item = name not in ("orc_tracking",)      # its Philox key is the item index
    nthr = min(os.cpu_count() or 1, 32)
    if per_item and n >= 50000 and nthr > 1:
        from concurrent.futures import ThreadPoolExecutor
        edges = np.linspace(0, n, nthr + 1).astype(int)

        def run(k):
            a, b = int(edges[k]), int(edges[k + 1])
            if b > a:
                sl = [_p(np.ascontiguousarray(x[a:b])) if isinstance(x, np.ndarray) and x.ndim >= 1 and x.shape[0] == n else (_p(x) if isinstance(x, np.ndarray) else x) for x in args]
                fn(C.c_int(b - a), *sl, _p(out[a:b]))
        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(run, range(nthr)))
        return out
    conv = [_p(a) if isinstance(a, np.ndarray) else a for a in args]
    fn(C.c_int(n), *conv, _p(out))
    return out


def rsi(pos, dirs, r):
    pos, dirs, r = _f32(pos), _f32(dirs), _f32(r)
    return _call("orc_rsi", len(r), (len(r), 2), pos, dirs, r)


def density(h):
    h = _f32(h)
    return _call("orc_density", len(h), (len(h), 3), h)


def spectra(wl, o3=None):
    wl = _f32(wl)
    o3 = _f32(load_luts()["o3"] if o3 is None else o3)
    return _call("orc_spectra", len(wl), (len(wl), 5), wl, o3)


def phase_eval(ray_dir, light_dir, ids, reduce):
    a, b = _f32(ray_dir), _f32(light_dir)
    i, r = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(reduce, np.int32)
    return _call("orc_phase_eval", len(i), (len(i),), a, b, i, r)


def phase_sample(ray_dir, ids, reduce, rand_u32):
    a = _f32(ray_dir)
    i, r = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(reduce, np.int32)
    u = np.ascontiguousarray(rand_u32, np.uint32)
    n = len(i)
    out_d = np.zeros((n, 3), np.float32)
    out_w = np.zeros(n, np.float32)
    lib().orc_phase_sample(C.c_int(n), _p(a), _p(i), _p(r), _p(u), _p(out_d), _p(out_w))
    return out_d, out_w


def dir_sample(kind, nrm, p, rand_u32):
    a = _f32(nrm)
    u = np.ascontiguousarray(rand_u32, np.uint32)
    n = len(a)
    out = np.zeros((n, 3), np.float32)
    lib().orc_dir_sample(C.c_int(n), C.c_int(kind), _p(a), C.c_float(p), _p(u), _p(out))
    return out


def brdf(albedo, ocean, bathy, v, nrm, l):
    args = [_f32(x) for x in (albedo, ocean, bathy, v, nrm, l)]
    n = len(args[0])
    return _call("orc_brdf", n, (n, 2), *args)


def srgb_to_spectrum(rgb, wl, lut=None):
    lut = np.ascontiguousarray(load_luts()["srgb2spec"] if lut is None else lut, np.float16)
    rgb, wl = _f32(rgb), _f32(wl)
    return _call("orc_srgb_to_spectrum", len(wl), (len(wl),), lut, rgb, wl)


def spectrum_sample(rand_u32, cie=None):
    cie = _f32(load_luts()["cie"] if cie is None else cie)
    u = np.ascontiguousarray(rand_u32, np.uint32)
    return _call("orc_spectrum_sample", len(u), (len(u), 5), cie, u)


def tex_fetch(tex_u8, pos):
    t = np.ascontiguousarray(tex_u8, np.uint8)
    if t.ndim == 2:
        t = t[:, :, None]
    ot = OrcTex(t.ctypes.data, t.shape[1], t.shape[0], t.shape[2])
    pos = _f32(pos)
    return _call("orc_tex_fetch", len(pos), (len(pos), 4), C.byref(ot), pos)


def cast_dir(scene, u, v, rand_u32):
    u, v = _f32(u), _f32(v)
    r = np.ascontiguousarray(rand_u32, np.uint32)
    return _call("orc_cast_dir", len(u), (len(u), 3), scene.ref, u, v, r)


def opendrt(rgb):
    rgb = _f32(rgb)
    return _call("orc_opendrt", len(rgb), rgb.shape, rgb)


def agx(rgb):
    rgb = _f32(rgb)
    return _call("orc_agx", len(rgb), rgb.shape, rgb)


def crf(scene, rgb):
    rgb = _f32(rgb)
    return _call("orc_crf", len(rgb), rgb.shape, scene.ref, rgb)


def srgb_oetf(x):
    x = _f32(x)
    return _call("orc_srgb_oetf", x.size, x.shape, x)


def resolve(scene, accum, samples):
    accum = _f32(accum)
    out = np.zeros_like(accum)
    lib().orc_resolve(scene.ref, _p(accum), C.c_int(samples), _p(out))
    return out


def intersect_land(scene, pos, dirs):
    pos, dirs = _f32(pos), _f32(dirs)
    return _call("orc_intersect_land", len(pos), (len(pos),), scene.ref, pos, dirs)


def intersect_land_iters(scene, pos, dirs):
    """(distance or -1, SDF evaluations) per ray; 250 evaluations and a hit = the reference's iteration-cap artefact."""
    pos, dirs = _f32(pos), _f32(dirs)
    return _call("orc_intersect_land_iters", len(pos), (len(pos), 2), scene.ref, pos, dirs)


def land_normal(scene, pos):
    pos = _f32(pos)
    return _call("orc_land_normal", len(pos), (len(pos), 3), scene.ref, pos)


def cloud_limits(pos, dirs, land):
    pos, dirs, land = _f32(pos), _f32(dirs), _f32(land)
    return _call("orc_cloud_limits", len(land), (len(land), 2), pos, dirs, land)


def clouds_density(scene, pos):
    pos = _f32(pos)
    return _call("orc_clouds_density", len(pos), (len(pos),), scene.ref, pos)


def land_material(scene, pos):
    pos = _f32(pos)
    return _call("orc_land_material", len(pos), (len(pos), 6), scene.ref, pos)


def raymarch_T(pos, dirs, ext):
    pos, dirs, ext = _f32(pos), _f32(dirs), _f32(ext)
    return _call("orc_raymarch_T", len(pos), (len(pos),), pos, dirs, ext)


def tracking(kind, scene, pos, dirs, land, wl, seed):
    pos, dirs, land, wl = _f32(pos), _f32(dirs), _f32(land), _f32(wl)
    n = len(land)
    out = np.zeros((n, 3), np.float32)
    lib().orc_tracking(C.c_int(n), C.c_int(kind), scene.ref, _p(pos), _p(dirs), _p(land), _p(wl), C.c_uint32(seed), _p(out))
    return out


INTEGRATORS = {"path_tracer": 0, "ray_marcher": 1}  # pathtracer.py:316 (what the renderer calls) / :543 (the preview)


def trace_paths(scene, px, py, sample, seed, counters=False, integrator="path_tracer"):
    px, py = np.ascontiguousarray(px, np.int32), np.ascontiguousarray(py, np.int32)
    sm = np.ascontiguousarray(sample, np.uint32)
    out = np.zeros((len(px), 5), np.float32)
    cnt = OrcCounters()
    lib().orc_trace_paths2(scene.ref, C.c_int(len(px)), _p(px), _p(py), _p(sm), C.c_uint32(seed), _p(out), C.byref(cnt), C.c_int(INTEGRATORS[integrator]))
    return (out, cnt.as_dict()) if counters else out


def ray_march(scene, pos, direction, t0, t1, sun, wavelength):
    """(in_scatter, transmittance) of ray_marh_atmos and ray_march_transmittance(pos, sun) per row (pathtracer.py:471-541)."""
    pos, direction, sun = (np.ascontiguousarray(a, np.float32) for a in (pos, direction, sun))
    t0, t1, wl = (np.ascontiguousarray(a, np.float32) for a in (t0, t1, wavelength))
    out2, outT = np.zeros((len(pos), 2), np.float32), np.zeros(len(pos), np.float32)
    lib().orc_ray_march(scene.ref, C.c_int(len(pos)), _p(pos), _p(direction), _p(t0), _p(t1), _p(sun), _p(wl), _p(out2), _p(outT))
    return out2, outT


def render(scene, spp, first_sample=0, seed=0, window=None, nthreads=None, second_moment=False, integrator="path_tracer"):
    """accum[H][W][3] (+ optional squared-sum buffer) and the event counters."""
    W, H = scene.s.W, scene.s.H
    x0, y0, w, h = window or (0, 0, W, H)
    accum = np.zeros((H, W, 3), np.float32)
    accum2 = np.zeros((H, W, 3), np.float32) if second_moment else None
    cnt = OrcCounters()
    nthreads = nthreads or os.cpu_count() or 1
    lib().orc_render2(scene.ref, C.c_int(x0), C.c_int(y0), C.c_int(w), C.c_int(h), C.c_int(spp), C.c_uint32(first_sample),
                      C.c_uint32(seed), _p(accum), _p(accum2) if second_moment else None, C.c_int(nthreads), C.byref(cnt),
                      C.c_int(INTEGRATORS[integrator]))
    return (accum, accum2, cnt.as_dict()) if second_moment else (accum, cnt.as_dict())
