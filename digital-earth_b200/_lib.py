"""ctypes binding of libde.so (include/de_api.h).  There is NO fallback: if the shared library
or a CUDA device is missing, loading / context creation raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DE_LIB_PATH") or os.path.join(_HERE, "libde.so")  # DE_LIB_PATH: tuning builds

DE_MODE_WAVEFRONT, DE_MODE_MEGAKERNEL, DE_MODE_PARITY = 0, 1, 2
DE_MODE_PREVIEW = 3
MODES = {"wavefront": DE_MODE_WAVEFRONT, "megakernel": DE_MODE_MEGAKERNEL, "parity": DE_MODE_PARITY, "preview": DE_MODE_PREVIEW}
TEX_SLOTS = ("albedo", "topography", "ocean", "clouds", "bathymetry", "emissive", "stars")


class DeParams(C.Structure):
    _fields_ = [
        ("cam_pos", C.c_float * 3), ("look_at", C.c_float * 3), ("up", C.c_float * 3),
        ("fov", C.c_float), ("aspect_scale", C.c_float), ("sun_angle", C.c_float), ("sun_path_rot", C.c_float),
        ("land_height_scale", C.c_float), ("exposure", C.c_float), ("gamma", C.c_float),
        ("selected_crf", C.c_int32), ("crf_count", C.c_int32),
        ("vignette_strength", C.c_float), ("vignette_radius", C.c_float), ("vignette_center", C.c_float * 2),
        ("tonemapper", C.c_int32), ("topo_tex_w", C.c_int32),
    ]


class DeCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "segments", "rmo_steps", "cloud_steps", "sdf_evals", "tex_fetches", "surface_hits", "rng_draws")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class DeError(RuntimeError):
    pass


_P, _I, _U, _F = C.c_void_p, C.c_int, C.c_uint32, C.c_float
_SIGS = {
    "de_abi_version": ([], _I),
    "de_create": ([C.POINTER(_P), _I, _I, _I], _I),
    "de_destroy": ([_P], None),
    "de_last_error": ([_P], C.c_char_p),
    "de_set_stream": ([_P, _P], _I),
    "de_set_mode": ([_P, _I], _I),
    "de_set_option": ([_P, C.c_char_p, _I], _I),
    "de_get_moment2": ([_P, C.POINTER(_P)], _I),
    "de_get_launch_timeline": ([_P, _P], _I),
    "de_get_cta_timeline": ([_P, _P, _I], _I),
    "de_bench_tex_gather": ([_P, _I, _I, C.POINTER(C.c_double)], _I),
    "de_set_params": ([_P, C.POINTER(DeParams)], _I),
    "de_upload_texture": ([_P, _I, _P, _I, _I, _I], _I),
    "de_upload_luts": ([_P, _P, _P, _P, _P, _I], _I),
    "de_reset": ([_P], _I),
    "de_accumulate": ([_P, _I, _U, _U, _I, _I, _I, _I], _I),
    "de_accumulate_tiles": ([_P, _I, _U, _U, _I, _I], _I),
    "de_resolve_peers_tiled": ([_P, C.POINTER(_P), C.POINTER(C.c_int), _I, _I, _I, _P, _I], _I),
    "de_get_accum": ([_P, C.POINTER(_P)], _I),
    "de_resolve": ([_P, _P, _P, _I], _I),
    "de_ipc_export_accum": ([_P, _P], _I),
    "de_ipc_open_peer": ([_P, _P, C.POINTER(_P)], _I),
    "de_ipc_close_peers": ([_P], _I),
    "de_resolve_peers": ([_P, C.POINTER(_P), _I, _P, _I], _I),
    "de_fetch_image_host": ([_P, _P, _I], _I),
    "de_sync": ([_P], _I),
    "de_get_counters": ([_P, C.POINTER(DeCounters)], _I),
    "de_set_counting": ([_P, _I], _I),
    "de_get_stage_profile": ([_P, _P], _I),
    "de_test_philox": ([_P, _P, _P, _I], _I),
    "de_test_rsi": ([_P, _P, _P, _P, _P, _I], _I),
    "de_test_density": ([_P, _P, _P, _I], _I),
    "de_test_spectra": ([_P, _P, _P, _I], _I),
    "de_test_phase_eval": ([_P, _P, _P, _P, _P, _P, _I], _I),
    "de_test_phase_sample": ([_P, _P, _P, _P, _P, _P, _P, _I], _I),
    "de_test_dir_sample": ([_P, _I, _P, _F, _P, _P, _I], _I),
    "de_test_brdf": ([_P, _P, _P, _P, _P, _P, _P, _P, _I], _I),
    "de_test_srgb2spec": ([_P, _P, _P, _P, _I], _I),
    "de_test_spectrum_sample": ([_P, _P, _P, _I], _I),
    "de_test_tex_fetch": ([_P, _I, _P, _P, _I], _I),
    "de_test_cast_dir": ([_P, _P, _P, _P, _P, _I], _I),
    "de_test_opendrt": ([_P, _P, _P, _I], _I),
    "de_test_agx": ([_P, _P, _P, _I], _I),
    "de_test_crf": ([_P, _P, _P, _I], _I),
    "de_test_srgb_oetf": ([_P, _P, _P, _I], _I),
    "de_test_intersect_land": ([_P, _P, _P, _P, _I], _I),
    "de_test_land_normal": ([_P, _P, _P, _I], _I),
    "de_test_land_material": ([_P, _P, _P, _I], _I),
    "de_test_cloud_limits": ([_P, _P, _P, _P, _P, _I], _I),
    "de_test_clouds_density": ([_P, _P, _P, _I], _I),
    "de_test_raymarch_T": ([_P, _P, _P, _P, _P, _I], _I),
    "de_test_tracking": ([_P, _I, _P, _P, _P, _P, _U, _P, _I], _I),
    "de_test_trace_paths": ([_P, _P, _P, _P, _U, _P, _I], _I),
    "de_test_trace_preview": ([_P, _P, _P, _P, _U, _P, _I], _I),
    "de_test_ray_march": ([_P, _P, _P, _P, _P, _P, _P, _P, _I], _I),
    "de_test_fast_cloud_bound": ([_P, _P, _P, _P, _P, _P, _I], _I),
    "de_test_fast_rmo_majorant": ([_P, _P, _P, _P, _P, _P, _P, _I], _I),
    "de_test_fast_land": ([_P, _P, _P, _P, _I], _I),
    "de_test_fast_rmo_bands": ([_P, _P, _P, _P, _P, _P, _P, _I, _P, _I], _I),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load():
    """dlopen libde.so; raises if it has not been built (python digital-earth_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DeError(f"{LIB_PATH} is missing -- build it with `python digital-earth_b200/build.py` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.argtypes, fn.restype = args, res
        if lib.de_abi_version() != 1:
            raise DeError("libde.so ABI version mismatch")
        _lib = lib
    return _lib


def check(ctx, rc):
    if rc != 0:
        msg = load().de_last_error(ctx)
        raise DeError("libde error %d: %s" % (rc, msg.decode() if msg else "?"))
