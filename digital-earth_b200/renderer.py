"""`Renderer` -- the reference's host driver class (renderer.py:16-401) on top of libde.so.

Same construction, attributes, setters, 0-d "fields" (`renderer.fov[None]` ...), `copy_textures`,
`reset_framebuffer`, `accumulate`, `fetch_image`.  Device memory is owned by torch tensors or by
the libde context and is only ever handed across the C-ABI as raw pointers.
"""
import ctypes as C

import numpy as np

from . import _lib, textures as _tex


class _Field0:
    """Taichi 0-d field look-alike: `f[None]` reads, `f[None] = v` writes (renderer.py:27-41)."""

    def __init__(self, owner, value, kind="f32"):
        self._owner, self._kind = owner, kind
        self._v = None
        self._set(value)

    def _set(self, v):
        if self._kind == "f32":
            self._v = float(np.float32(v))
        elif self._kind == "i32":
            self._v = int(v)
        else:
            self._v = tuple(float(np.float32(x)) for x in v)

    def __getitem__(self, key):
        assert key is None, "0-d field: index with [None]"
        return self._v

    def __setitem__(self, key, value):
        assert key is None, "0-d field: index with [None]"
        self._set(value)
        self._owner._dirty = True


class _CtxHandle:
    """Owns the libde context.  Shared by the Renderer and by every tensor that views context-owned device memory, so the
    context (and the memory) lives exactly as long as its last user: a `color_buffer` tensor kept after `Renderer.close()`
    stays valid, and the context is destroyed when that tensor goes away."""

    def __init__(self, lib, ctx):
        self.lib, self.ctx = lib, ctx

    def __del__(self):
        try:
            if self.ctx is not None and self.ctx.value:
                self.lib.de_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass


class _CudaView:
    """__cuda_array_interface__ wrapper; torch keeps this object (hence the context handle) alive with the storage."""

    def __init__(self, ptr, shape, handle):
        self._handle = handle
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": "<f4", "version": 2, "strides": None}


class Renderer:
    def __init__(self, image_res, up, textures=None, device=None, mode="wavefront", texture_quality=_tex.TEXTURE_QUALITY,
                 texture_dir="textures", assets_dir=None, seed=0):
        import torch
        if not torch.cuda.is_available():
            raise _lib.DeError("no CUDA device: the B200 renderer has no CPU fallback")
        self._torch = torch
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.image_res = (int(image_res[0]), int(image_res[1]))
        if self.image_res[0] % 16 or self.image_res[1] % 8:
            raise ValueError("image_res must be a multiple of (16, 8) (renderer.py:46)")
        self.aspect_ratio = image_res[0] / image_res[1]
        self.vignette_strength = 0.9   # renderer.py:20-22
        self.vignette_radius = 0.0
        self.vignette_center = [0.5, 0.5]
        self.current_spp = 0
        self.seed = int(seed)
        self.tonemapper = 0            # 0 OpenDRT (renderer.py:357), 1 AgX (renderer.py:356, commented out upstream)
        self._dirty = True
        self._ctx = C.c_void_p()
        rc = self._lib.de_create(C.byref(self._ctx), self.device.index, self.image_res[0], self.image_res[1])
        if rc != 0:
            raise _lib.DeError("de_create failed (%d)" % rc)
        self._handle = _CtxHandle(self._lib, self._ctx)
        self._moment2 = None
        self.set_mode(mode)

        # 0-d fields (renderer.py:27-41) with the defaults of renderer.py:49-56
        self.camera_pos = _Field0(self, (-15000000.0, 0.0, 15000000.0), "vec3")  # earth_viewer.py:27
        self.look_at = _Field0(self, (0.0, 0.0, 0.0), "vec3")
        self.up = _Field0(self, (0.0, 1.0, 0.0), "vec3")
        self.fov = _Field0(self, 0.0)
        self.aspect_scale = _Field0(self, 1.0)
        self.exposure = _Field0(self, 2.5)
        self.gamma = _Field0(self, 1.0)
        self.selected_crf = _Field0(self, 0, "i32")
        self.crf_count = _Field0(self, 0, "i32")
        self.sun_angle = _Field0(self, 0.0)
        self.sun_path_rot = _Field0(self, 0.0)
        self.set_up(*up)
        self.set_fov(np.radians(27.0) * 0.5)
        self.set_aspect_scale(1.0)
        self.set_exposure(2.5)
        self.set_gamma(1.0)
        self.set_crf(0)
        self.set_sun_angle(np.radians(60.0))
        self.set_sun_path_rot(np.radians(-45.0))
        self.land_height_scale = 7800.0  # renderer.py:58

        # textures (renderer.py:60-94): dict of arrays, a directory of the NASA maps, or None -> texture_dir
        if textures is None or isinstance(textures, str):
            textures = _tex.load_directory(textures or texture_dir, texture_quality)
        self._textures = {k: np.ascontiguousarray(textures[k], dtype=np.uint8) for k in _lib.TEX_SLOTS}
        self.topography_tex_res = self._textures["topography"].shape[1::-1]

        # LUTs + camera response functions (renderer.py:96-134,147-167)
        luts = _tex.load_luts(assets_dir)
        self._luts = luts
        self.crf_names = list(luts["crf_names"])
        self.crf_lut_res = (1024, len(self.crf_names))
        self.set_crf_count(self.crf_lut_res[1])
        self._textures_copied = False
        self._image = torch.empty((self.image_res[1], self.image_res[0], 3), dtype=torch.float32, device=self.device)
        ptr = C.c_void_p()
        self._check(self._lib.de_get_accum(self._ctx, C.byref(ptr)))
        self._accum = torch.as_tensor(_CudaView(ptr.value, (self.image_res[1], self.image_res[0], 3), self._handle), device=self.device)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        _lib.check(self._ctx, rc)

    def close(self):
        """Release the context.  Tensors handed out earlier (`color_buffer`, `moment2`) keep it alive until they die."""
        self._accum = self._moment2 = self._handle = None
        self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """Integrator options of include/de_api.h (`space_tiles`, `space_async`, `tile_order`, `moments`, `timeline`, `linear_textures`)."""
        self._bind_stream()
        self._check(self._lib.de_set_option(self._ctx, str(name).encode(), int(value)))
        if name == "moments":
            self._moment2 = None

    @property
    def moment2(self):
        """Per-pixel sums of squared sample contributions [H][W][3] (option `moments`): with `color_buffer` the sample
        variance of every pixel, for the image z-test of SURVEY 8d."""
        if self._moment2 is None:
            ptr = C.c_void_p()
            self._check(self._lib.de_get_moment2(self._ctx, C.byref(ptr)))
            self._moment2 = self._torch.as_tensor(_CudaView(ptr.value, (self.image_res[1], self.image_res[0], 3), self._handle), device=self.device)
        return self._moment2

    def launch_timeline(self):
        """Timeline of the last wavefront accumulate() with option `timeline` (ms relative to the first CTA's start)."""
        buf = np.zeros(8, np.uint64)
        self._check(self._lib.de_get_launch_timeline(self._ctx, buf.ctypes.data_as(C.c_void_p)))
        t0 = int(buf[0])
        ms = lambda k: (int(buf[k]) - t0) * 1e-6  # noqa: E731
        return {"first_exhaust_ms": ms(1), "last_exhaust_ms": ms(2), "first_cta_end_ms": ms(3), "last_cta_end_ms": ms(4),
                "min_chunks_per_cta": int(buf[5]), "max_chunks_per_cta": int(buf[6]),
                "wavefront_tiles": int(buf[7]) & 0xFFFFFFFF, "space_tiles": int(buf[7]) >> 32}

    def tex_gather_peak(self, slot=3, iters=4096):
        """Measured peak of the integrator's fetch instruction (tex2Dgather, L1-resident footprints) in requests/s."""
        if not self._textures_copied:
            self.copy_textures()
        self._bind_stream()
        v = C.c_double()
        self._check(self._lib.de_bench_tex_gather(self._ctx, int(slot), int(iters), C.byref(v)))
        return float(v.value)

    def cta_timeline(self):
        """Per-CTA drain diagnostics of the last wavefront accumulate() with option `timeline`: list of dicts, times in ms after the
        launch's first CTA start."""
        t0 = np.zeros(8, np.uint64)
        self._check(self._lib.de_get_launch_timeline(self._ctx, t0.ctypes.data_as(C.c_void_p)))
        buf = np.zeros((256, 24), np.uint64)
        n = self._lib.de_get_cta_timeline(self._ctx, buf.ctypes.data_as(C.c_void_p), 256)
        if n < 0:
            self._check(n)
        names = ("NEW", "SDF", "RMO", "CLOUD", "SDF_DONE", "RMO_DONE", "EVENT", "NEE_DONE", "SURFACE")
        ms = lambda v: (int(v) - int(t0[0])) * 1e-6 if int(v) else float("nan")  # noqa: E731
        return [{"exhaust_ms": ms(b[0]), "few_ms": ms(b[1]), "end_ms": ms(b[2]), "chunks": int(b[3]),
                 "visits": {k: int(b[4 + i]) for i, k in enumerate(names)}, "slots": {k: int(b[13 + i]) for i, k in enumerate(names)}} for b in buf[:n]]

    def set_mode(self, mode):
        """'wavefront' (product), 'megakernel' (1 thread/pixel baseline) or 'parity' (IEEE source-order)."""
        self.mode = mode
        self._check(self._lib.de_set_mode(self._ctx, _lib.MODES[mode]))

    def set_counting(self, enabled):
        self._check(self._lib.de_set_counting(self._ctx, int(bool(enabled))))

    def counters(self):
        c = _lib.DeCounters()
        self._check(self._lib.de_get_counters(self._ctx, C.byref(c)))
        return c.as_dict()

    def stage_profile(self):
        """Scheduler self-profile of the last counting accumulate(): {stage: (warp cycles, visits, slots)} + idle."""
        import numpy as np
        buf = np.zeros(32, np.uint64)
        self._check(self._lib.de_get_stage_profile(self._ctx, buf.ctypes.data_as(C.c_void_p)))
        names = ("NEW", "SDF", "RMO", "CLOUD", "SDF_DONE", "RMO_DONE", "EVENT", "NEE_DONE", "SURFACE")
        out = {n: tuple(int(x) for x in buf[3 * i:3 * i + 3]) for i, n in enumerate(names)}
        out["IDLE"] = (int(buf[3 * len(names)]), 0, 0)
        out["SWITCHES"] = (0, int(buf[3 * len(names) + 1]), 0)  # visits whose stage differs from the SM's previous visit
        return out

    def _bind_stream(self):
        self._check(self._lib.de_set_stream(self._ctx, C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)))

    def _params(self):
        p = _lib.DeParams()
        p.cam_pos[:] = self.camera_pos[None]
        p.look_at[:] = self.look_at[None]
        p.up[:] = self.up[None]
        p.fov, p.aspect_scale = self.fov[None], self.aspect_scale[None]
        p.sun_angle, p.sun_path_rot = self.sun_angle[None], self.sun_path_rot[None]
        p.land_height_scale = self.land_height_scale
        p.exposure, p.gamma = self.exposure[None], self.gamma[None]
        p.selected_crf, p.crf_count = self.selected_crf[None], self.crf_count[None]
        p.vignette_strength, p.vignette_radius = self.vignette_strength, self.vignette_radius
        p.vignette_center[:] = self.vignette_center
        p.tonemapper = int(self.tonemapper)
        p.topo_tex_w = int(self.topography_tex_res[0])
        return p

    def _push_params(self):
        p = self._params()
        key = bytes(p)
        if self._dirty or key != getattr(self, "_pushed", None):
            self._check(self._lib.de_set_params(self._ctx, C.byref(p)))
            self._pushed, self._dirty = key, False

    # ------------------------------------------------------------------ reference surface
    def copy_textures(self):
        """renderer.py:136-145: host arrays -> device textures + LUTs."""
        self._bind_stream()
        for i, name in enumerate(_lib.TEX_SLOTS):
            t = self._textures[name]
            ch = 1 if t.ndim == 2 else t.shape[2]
            self._check(self._lib.de_upload_texture(self._ctx, i, t.ctypes.data_as(C.c_void_p), t.shape[1], t.shape[0], ch))
        L = self._luts
        cie = np.ascontiguousarray(L["cie"], np.float32)
        s2s = np.ascontiguousarray(L["srgb2spec"], np.float16)
        o3 = np.ascontiguousarray(L["o3"], np.float32)
        crf = np.ascontiguousarray(L["crf"], np.float32)
        self._check(self._lib.de_upload_luts(self._ctx, cie.ctypes.data_as(C.c_void_p), s2s.ctypes.data_as(C.c_void_p),
                                             o3.ctypes.data_as(C.c_void_p), crf.ctypes.data_as(C.c_void_p), crf.shape[0]))
        self._textures_copied = True

    def load_crfs(self, directory=None):
        """renderer.py:147-167; returns (1024, n, 3) like the reference.  directory=None -> packaged curves."""
        if directory is not None:
            crf, names = _tex.load_crf_directory(directory)
            self._luts = dict(self._luts, crf=crf, crf_names=names)
            self.crf_names = list(names)
            self.crf_lut_res = (1024, len(names))
            self.set_crf_count(len(names))
            self._textures_copied = False
        return np.ascontiguousarray(self._luts["crf"].transpose(1, 0, 2))

    def set_camera_pos(self, x, y, z): self.camera_pos[None] = (x, y, z)          # renderer.py:224
    def set_up(self, x, y, z): self.up[None] = (x, y, z)                          # :228 (normalised on device)
    def set_look_at(self, x, y, z): self.look_at[None] = (x, y, z)                # :232
    def set_fov(self, fov): self.fov[None] = fov                                  # :236
    def set_aspect_scale(self, scale): self.aspect_scale[None] = scale            # :240
    def set_exposure(self, exposure): self.exposure[None] = exposure              # :244
    def set_gamma(self, gam): self.gamma[None] = gam                              # :248
    def set_crf(self, index): self.selected_crf[None] = index                     # :252
    def set_crf_count(self, num): self.crf_count[None] = num                      # :256
    def set_sun_angle(self, ang): self.sun_angle[None] = ang                      # :260
    def set_sun_path_rot(self, ang): self.sun_path_rot[None] = ang                # :264

    def apply_config(self, cfg):
        """Apply a parsed 10-line config (config.load_config)."""
        self.set_camera_pos(*cfg["cam_pos"]); self.set_look_at(*cfg["look_at"]); self.set_up(*cfg["up"])
        self.set_fov(cfg["fov"]); self.set_aspect_scale(cfg["aspect_scale"]); self.set_exposure(cfg["exposure"])
        self.set_crf(cfg["selected_crf"]); self.set_gamma(cfg["gamma"])
        self.set_sun_angle(cfg["sun_angle"]); self.set_sun_path_rot(cfg["sun_path_rot"])

    def reset_framebuffer(self):
        """renderer.py:367-369"""
        self._bind_stream()
        self.current_spp = 0
        self._check(self._lib.de_reset(self._ctx))

    def accumulate(self, n_spp=1, window=None, first_sample=None, tiles=None):
        """renderer.py:371-380 (+1 spp); n_spp > 1 renders several samples per pixel in one launch.
        first_sample: Philox sample index of the first sample (default: current_spp).
        tiles=(stride, offset): only the 16x8 film tiles t with t % stride == offset (multi-GPU tile partition)."""
        if not self._textures_copied:
            self.copy_textures()
        self._bind_stream()
        self._push_params()
        fs = self.current_spp if first_sample is None else int(first_sample)
        if tiles is not None:
            if window is not None:
                raise ValueError("a tile partition covers the whole frame: pass either window or tiles")
            self._check(self._lib.de_accumulate_tiles(self._ctx, int(n_spp), self.seed & 0xFFFFFFFF, fs & 0xFFFFFFFF, int(tiles[0]), int(tiles[1])))
        else:
            x0, y0, w, h = window or (0, 0, self.image_res[0], self.image_res[1])
            self._check(self._lib.de_accumulate(self._ctx, int(n_spp), self.seed & 0xFFFFFFFF, fs & 0xFFFFFFFF, x0, y0, w, h))
        self.current_spp += int(n_spp)

    def fetch_image(self, accum=None, spp=None):
        """renderer.py:382-384: resolve + tonemap; returns a (W, H, 3) float32 CUDA tensor in [0,1]
        indexed [x][y] with y up, like the reference's `_rendered_image` field."""
        if not self._textures_copied:
            self.copy_textures()
        self._bind_stream()
        self._push_params()
        spp = self.current_spp if spp is None else int(spp)
        src = C.c_void_p(accum.data_ptr()) if accum is not None else None
        self._check(self._lib.de_resolve(self._ctx, src, C.c_void_p(self._image.data_ptr()), max(spp, 1)))
        return self._image.permute(1, 0, 2)

    # ---- multi-GPU resolve fused with the accumulation exchange (include/de_api.h) ----
    def export_accum_handle(self):
        """64-byte CUDA IPC handle of this rank's accumulation buffer (bytes; send it to the resolving rank)."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.de_ipc_export_accum(self._ctx, buf))
        return buf.raw

    def open_peer(self, handle):
        """Map another rank's accumulation buffer (same node) into this process; returns its device address."""
        ptr = C.c_void_p()
        self._check(self._lib.de_ipc_open_peer(self._ctx, C.create_string_buffer(bytes(handle), 64), C.byref(ptr)))
        return ptr.value

    def close_peers(self):
        self._check(self._lib.de_ipc_close_peers(self._ctx))

    def fetch_image_peers(self, peers, spp, tile_stride=1, own_offset=0, peer_offsets=None):
        """fetch_image() of (own buffer + the peers' buffers): one kernel reads the other ranks' partial sums over
        NVLink peer memory while it resolves.  peers: device addresses (open_peer) or CUDA tensors [H][W][3] f32.
        With a tile partition (tile_stride > 1) every pixel reads only the buffers of the ranks that rendered its tile."""
        if not self._textures_copied:
            self.copy_textures()
        self._bind_stream()
        self._push_params()
        addrs = [int(p.data_ptr()) if hasattr(p, "data_ptr") else int(p) for p in peers]
        arr = (C.c_void_p * max(len(addrs), 1))(*addrs)
        offs = list(peer_offsets) if peer_offsets is not None else [0] * len(addrs)
        if len(offs) != len(addrs):
            raise ValueError("peer_offsets needs one entry per peer")
        oarr = (C.c_int * max(len(addrs), 1))(*offs)
        self._check(self._lib.de_resolve_peers_tiled(self._ctx, arr, oarr, len(addrs), int(tile_stride), int(own_offset),
                                                     C.c_void_p(self._image.data_ptr()), max(int(spp), 1)))
        return self._image.permute(1, 0, 2)

    @property
    def color_buffer(self):
        """The accumulation buffer (renderer.py:25) as a [H][W][3] float32 CUDA tensor (linear sRGB sums)."""
        return self._accum

    def sync(self):
        self._check(self._lib.de_sync(self._ctx))

    # progressive checkpoint: the linear accumulation buffer + the sample count is the
    # whole state of a progressive render, because a pixel's sample k is the same path in every launch
    def _accumulation_signature(self):
        """Everything the accumulated radiance depends on: camera, sun, terrain scale, integrator family and the texture set.
        (Exposure, gamma, response curve and tone mapper only act in fetch_image and may change between sessions.)"""
        import zlib
        p = self._params()
        scene = [float(x) for x in (*p.cam_pos, *p.look_at, *p.up, p.fov, p.aspect_scale, p.sun_angle, p.sun_path_rot, p.land_height_scale)] + [int(p.topo_tex_w)]
        tex = {}
        for name in _lib.TEX_SLOTS:
            t = self._textures[name]
            step = max(t.shape[0] // 64, 1), max(t.shape[1] // 128, 1)            # a 64 x 128 sub-grid identifies the map cheaply
            tex[name] = [list(t.shape), int(zlib.crc32(np.ascontiguousarray(t[::step[0], ::step[1]]).tobytes()))]
        return {"scene": scene, "integrator": "preview" if self.mode == "preview" else "path_tracer", "textures": tex}

    def save_accumulation(self, path, extra=None):
        """Write {accum [H][W][3] f32, spp, seed, image_res, signature of scene + textures + integrator} to an .npz; returns the path."""
        import json
        self.sync()
        meta = dict(extra or {})
        np.savez_compressed(path, accum=self._accum.cpu().numpy(), spp=np.int64(self.current_spp), seed=np.int64(self.seed),
                            image_res=np.asarray(self.image_res, np.int64), meta=np.array(sorted(map(str, meta.items()))),
                            signature=np.array(json.dumps(self._accumulation_signature(), sort_keys=True)))
        return path if str(path).endswith(".npz") else str(path) + ".npz"

    def check_checkpoint(self, path):
        """Validate a checkpoint against this renderer WITHOUT loading it (every rank of a distributed resume calls this): resolution,
        seed, and -- when the file carries one -- the scene / texture / integrator signature.  Returns its spp; raises ValueError."""
        import json
        with np.load(path) as z:
            spp, seed = int(z["spp"]), int(z["seed"]) if "seed" in z else self.seed
            res = tuple(int(x) for x in z["image_res"]) if "image_res" in z else tuple(z["accum"].shape[1::-1])
            sig = json.loads(str(z["signature"])) if "signature" in z else None
        if tuple(res) != tuple(self.image_res):
            raise ValueError("checkpoint is %dx%d, renderer is %dx%d" % (res[0], res[1], self.image_res[0], self.image_res[1]))
        if seed != self.seed:
            raise ValueError("checkpoint was rendered with seed %d, renderer has seed %d: resuming would repeat or skip sample streams" % (seed, self.seed))
        if sig is not None:
            mine = json.loads(json.dumps(self._accumulation_signature(), sort_keys=True))
            for key in ("scene", "integrator", "textures"):
                if sig.get(key) != mine[key]:
                    raise ValueError("checkpoint was rendered with a different %s: resuming would mix two images into one accumulation" % key)
        return spp

    def load_accumulation(self, path):
        """Resume from save_accumulation(): the next accumulate() continues at sample index `spp`."""
        import torch
        spp = self.check_checkpoint(path)
        with np.load(path) as z:
            acc = z["accum"]
        if acc.shape != tuple(self._accum.shape):
            raise ValueError("checkpoint buffer has shape %s, renderer expects %s" % (acc.shape, tuple(self._accum.shape)))
        self._bind_stream()
        self._accum.copy_(torch.from_numpy(np.ascontiguousarray(acc, dtype=np.float32)))
        self.current_spp = spp
        return spp

    def render(self, spp, batch=None):
        """Convenience: reset + accumulate `spp` samples (in batches) + fetch_image."""
        self.reset_framebuffer()
        batch = batch or spp
        done = 0
        while done < spp:
            n = min(batch, spp - done)
            self.accumulate(n)
            done += n
        return self.fetch_image()
