"""numpy front-end of the `de_test_*` entry points (include/de_api.h): the deterministic
sub-paths of the integrator evaluated on the GPU in parity arithmetic.  Used by tests/ and smoke()."""
import ctypes as C

import numpy as np

from . import _lib


class Hooks:
    def __init__(self, renderer):
        self.r = renderer
        self.torch = renderer._torch
        self.lib = renderer._lib
        if not renderer._textures_copied:
            renderer.copy_textures()
        renderer._bind_stream()
        renderer._push_params()

    def _dev(self, a, dtype):
        return self.torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device=self.r.device)

    def _call(self, name, n, out_shape, *args, out_dtype=np.float32):
        self.r._bind_stream()
        self.r._push_params()
        td = {np.float32: self.torch.float32, np.uint32: self.torch.int32}[out_dtype]
        out = self.torch.zeros(out_shape, dtype=td, device=self.r.device)
        conv, keep = [], []
        for a in args:
            if isinstance(a, self.torch.Tensor):
                keep.append(a)
                conv.append(C.c_void_p(a.data_ptr()))
            else:
                conv.append(a)
        _lib.check(self.r._ctx, getattr(self.lib, name)(self.r._ctx, *conv, C.c_void_p(out.data_ptr()), int(n)))
        self.torch.cuda.synchronize(self.r.device)
        res = out.cpu().numpy()
        return res.view(np.uint32) if out_dtype is np.uint32 else res

    def f(self, a): return self._dev(a, np.float32)
    def i(self, a): return self._dev(a, np.int32)
    def u(self, a): return self._dev(np.ascontiguousarray(a, np.uint32).view(np.int32), np.int32)

    def philox(self, ctr_key6):
        n = len(ctr_key6)
        return self._call("de_test_philox", n, (n, 4), self.u(ctr_key6), out_dtype=np.uint32)

    def rsi(self, pos, d, r): return self._call("de_test_rsi", len(r), (len(r), 2), self.f(pos), self.f(d), self.f(r))
    def density(self, h): return self._call("de_test_density", len(h), (len(h), 3), self.f(h))
    def spectra(self, wl): return self._call("de_test_spectra", len(wl), (len(wl), 5), self.f(wl))
    def phase_eval(self, a, b, ids, red): return self._call("de_test_phase_eval", len(ids), (len(ids),), self.f(a), self.f(b), self.i(ids), self.i(red))

    def phase_sample(self, a, ids, red, rand):
        n = len(ids)
        ow = self.torch.zeros(n, dtype=self.torch.float32, device=self.r.device)
        od = self.torch.zeros((n, 3), dtype=self.torch.float32, device=self.r.device)
        ta, ti, tr, tu = self.f(a), self.i(ids), self.i(red), self.u(rand)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        _lib.check(self.r._ctx, self.lib.de_test_phase_sample(self.r._ctx, p(ta), p(ti), p(tr), p(tu), p(od), p(ow), n))
        self.torch.cuda.synchronize(self.r.device)
        return od.cpu().numpy(), ow.cpu().numpy()

    def dir_sample(self, kind, nrm, cmax, rand): return self._call("de_test_dir_sample", len(nrm), (len(nrm), 3), int(kind), self.f(nrm), C.c_float(cmax), self.u(rand))
    def brdf(self, al, oc, ba, v, nr, l): return self._call("de_test_brdf", len(al), (len(al), 2), self.f(al), self.f(oc), self.f(ba), self.f(v), self.f(nr), self.f(l))
    def srgb2spec(self, rgb, wl): return self._call("de_test_srgb2spec", len(wl), (len(wl),), self.f(rgb), self.f(wl))
    def spectrum_sample(self, rand): return self._call("de_test_spectrum_sample", len(rand), (len(rand), 5), self.u(rand))
    def tex_fetch(self, slot, pos): return self._call("de_test_tex_fetch", len(pos), (len(pos), 4), int(slot), self.f(pos))
    def cast_dir(self, u, v, rand): return self._call("de_test_cast_dir", len(u), (len(u), 3), self.f(u), self.f(v), self.u(rand))
    def opendrt(self, rgb): return self._call("de_test_opendrt", len(rgb), (len(rgb), 3), self.f(rgb))
    def agx(self, rgb): return self._call("de_test_agx", len(rgb), (len(rgb), 3), self.f(rgb))
    def crf(self, rgb): return self._call("de_test_crf", len(rgb), (len(rgb), 3), self.f(rgb))
    def srgb_oetf(self, x): return self._call("de_test_srgb_oetf", len(x), (len(x),), self.f(x))
    def intersect_land(self, pos, d): return self._call("de_test_intersect_land", len(pos), (len(pos),), self.f(pos), self.f(d))
    def land_normal(self, pos): return self._call("de_test_land_normal", len(pos), (len(pos), 3), self.f(pos))
    def land_material(self, pos): return self._call("de_test_land_material", len(pos), (len(pos), 6), self.f(pos))
    def cloud_limits(self, pos, d, land): return self._call("de_test_cloud_limits", len(land), (len(land), 2), self.f(pos), self.f(d), self.f(land))
    def clouds_density(self, pos): return self._call("de_test_clouds_density", len(pos), (len(pos),), self.f(pos))
    def raymarch_T(self, pos, d, ext): return self._call("de_test_raymarch_T", len(pos), (len(pos),), self.f(pos), self.f(d), self.f(ext))
    def tracking(self, kind, pos, d, land, wl, seed): return self._call("de_test_tracking", len(land), (len(land), 3), int(kind), self.f(pos), self.f(d), self.f(land), self.f(wl), C.c_uint32(seed))
    def trace_preview(self, px, py, sample, seed): return self._call("de_test_trace_preview", len(px), (len(px), 5), self.i(px), self.i(py), self.u(sample), C.c_uint32(seed))
    def ray_march(self, pos, direction, t0, t1, sun, wl): return self._call("de_test_ray_march", len(pos), (len(pos), 2), self.f(pos), self.f(direction), self.f(t0), self.f(t1), self.f(sun), self.f(wl))
    def trace_paths(self, px, py, sample, seed): return self._call("de_test_trace_paths", len(px), (len(px), 5), self.i(px), self.i(py), self.u(sample), C.c_uint32(seed))

    # product-flavour work-removal bounds (the device functions the wavefront kernel calls)
    def fast_cloud_bound(self, pos, d, ts, tm): return self._call("de_test_fast_cloud_bound", len(ts), (len(ts), 4), self.f(pos), self.f(d), self.f(ts), self.f(tm))
    def fast_rmo_majorant(self, pos, d, ts, tm, ext): return self._call("de_test_fast_rmo_majorant", len(ts), (len(ts),), self.f(pos), self.f(d), self.f(ts), self.f(tm), self.f(ext))
    def fast_land(self, pos, d): return self._call("de_test_fast_land", len(pos), (len(pos), 3), self.f(pos), self.f(d))

    def fast_rmo_bands(self, pos, d, ts, tm, ext, tq):
        tq = np.ascontiguousarray(tq, np.float32)
        return self._call("de_test_fast_rmo_bands", len(ts), tq.shape, self.f(pos), self.f(d), self.f(ts), self.f(tm), self.f(ext), self.f(tq), int(tq.shape[1]))
