/* de_api.h -- C-ABI of libde.so: the B200-native drop-in for the Taichi kernels of
 * AntonioFerreras/Digital-Earth (renderer.py / pathtracer.py).
 *
 * The reference has no FFI: its seam is the Python class `Renderer` (renderer.py:16) whose
 * methods launch two Taichi kernels.  Every entry point below names the reference interface it
 * replaces (file:line relative to the reference root).  Plain pointers and sizes only; no torch
 * or CUDA types in any signature (streams and device pointers travel as void* / float*).
 * INTEGRATION.md shows the ctypes binding a maintainer adds to renderer.py.
 *
 * Conventions
 *   - every call returns 0 on success or a negative DE_ERR_* code; de_last_error() has the text;
 *     nothing throws across the ABI.
 *   - one ctx <-> one CUDA device <-> one stream.  Calls enqueue work asynchronously on the ctx
 *     stream (de_set_stream shares torch's current stream); de_sync() or a sync of that stream
 *     completes them.  A ctx is not thread-safe; different ctxs may be driven concurrently.
 *   - images / accumulation buffers are row-major [y][x][3] float32, y = 0 is the BOTTOM row
 *     (reference pixel (u,v)=(0,0) is bottom-left, renderer.py:269-279).  The reference's
 *     16x8 film tiling (renderer.py:43-46) is an internal scheduling detail here.
 *   - textures are uint8 row-major [y][x][c], y = 0 is v = 0 (south pole); i.e. the transpose of
 *     the reference's ti.tools.imread arrays ([x][y][c], y up; renderer.py:61-94).
 *   - RNG: per (pixel, sample, bounce) a stream of 32-bit slots; slot i = word i&3 of
 *     Philox4x32-10(key = (seed, y*W + x), counter = (sample_index, bounce, i>>2, 0)); bounce 0 =
 *     wavelength + pixel jitter, bounce k+1 = path segment k.  Each ti.random() of the reference
 *     takes the next slot; tracking passes, the light-direction sample, sample_phase and the
 *     hemisphere sample start on a multiple of 4, and a ratio-tracking trip owns two slots (see
 *     oracle/de_oracle.c header).  Every integrator flavour consumes the same stream, so a pixel's
 *     samples are the same paths in all of them.
 */
#ifndef DE_API_H
#define DE_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* libde.so is built with -fvisibility=hidden */
#endif

#define DE_ABI_VERSION 1

enum {
    DE_OK = 0,
    DE_ERR_INVALID = -1, /* bad argument (NULL, size, W%16 / H%8, slot ...) */
    DE_ERR_CUDA = -2,    /* a CUDA runtime call failed                      */
    DE_ERR_STATE = -3,   /* call order: textures / LUTs / params missing    */
    DE_ERR_NOMEM = -4
};

/* texture slots: renderer.py:61-94 */
enum { DE_TEX_ALBEDO = 0, DE_TEX_TOPOGRAPHY, DE_TEX_OCEAN, DE_TEX_CLOUDS, DE_TEX_BATHYMETRY, DE_TEX_EMISSIVE, DE_TEX_STARS, DE_TEX_COUNT };

/* integrator flavours */
enum {
    DE_MODE_WAVEFRONT = 0, /* product path: persistent-thread, stage-sorted, FMA + fast intrinsics */
    DE_MODE_MEGAKERNEL = 1,/* one thread per pixel, same fast arithmetic (baseline for profiles)   */
    DE_MODE_PARITY = 2,    /* one thread per pixel, IEEE source-order arithmetic (no FMA
                              contraction, accurate libm): comparable to the oracle per path       */
    DE_MODE_PREVIEW = 3    /* deterministic ray-marching integrator `ray_marcher` (pathtracer.py:543-685,
                              unreferenced upstream): 64-step view march x 16-step sun march, <= 3 surface
                              bounces, no clouds -- a noise-free atmosphere preview; fast arithmetic   */
};

typedef struct de_ctx de_ctx;

/* Scalar state of the reference Renderer: the 0-d Taichi fields written by the set_* kernels
 * (renderer.py:224-266), SceneParameters inputs (lib/parameters.py:10-15, renderer.py:293-302)
 * and the Python attributes read by _render_to_image (renderer.py:20-22,58). */
typedef struct DeParams {
    float cam_pos[3];        /* set_camera_pos                      renderer.py:224 */
    float look_at[3];        /* set_look_at                         renderer.py:232 */
    float up[3];             /* set_up (normalised on device)       renderer.py:228 */
    float fov;               /* tangent half-height                 renderer.py:236 */
    float aspect_scale;      /*                                     renderer.py:240 */
    float sun_angle;         /* radians                             renderer.py:260 */
    float sun_path_rot;      /* radians                             renderer.py:264 */
    float land_height_scale; /* 7800                                renderer.py:58  */
    float exposure;          /* stops                               renderer.py:244 */
    float gamma;             /*                                     renderer.py:248 */
    int32_t selected_crf;    /*                                     renderer.py:252 */
    int32_t crf_count;       /*                                     renderer.py:256 */
    float vignette_strength; /* 0.9                                 renderer.py:20  */
    float vignette_radius;   /* 0.0                                 renderer.py:21  */
    float vignette_center[2];/* (0.5, 0.5)                          renderer.py:22  */
    int32_t tonemapper;      /* 0 = OpenDRT (renderer.py:357), 1 = AgX (renderer.py:356) */
    int32_t topo_tex_w;      /* TOPOGRAPHY_TEX_RES[0] (lib/textures.py) used at pathtracer.py:20;
                                0 = width of the uploaded topography texture */
} DeParams;

/* event counters of the last de_accumulate (for the roofline's FLOP/path formula) */
typedef struct DeCounters {
    uint64_t paths, segments, rmo_steps, cloud_steps, sdf_evals, tex_fetches, surface_hits, rng_draws;
} DeCounters;

int de_abi_version(void);

/* Renderer.__init__(image_res, up)  renderer.py:17-58.  W%16==0 and H%8==0 as renderer.py:46. */
int de_create(de_ctx **out, int device, int width, int height);
void de_destroy(de_ctx *ctx);
const char *de_last_error(de_ctx *ctx);
int de_set_stream(de_ctx *ctx, void *cuda_stream);
int de_set_mode(de_ctx *ctx, int mode);

/* integrator options (name, value); unknown names fail with DE_ERR_INVALID:
 *   "space_tiles"     1 (default) film tiles that cannot see the atmosphere shell are rendered by a dedicated converged kernel
 *   "space_async"     1 (default) ... on a side stream, overlapping the persistent kernel's drain
 *   "tile_order"      1 (default) the persistent kernel starts with the film tiles that can produce long paths (cloud in sight, limb) and ends
 *                     with the clear ones, so the serial tail of a launch -- its longest path -- overlaps the rest of the work.  Changes the
 *                     order of the work only: every (pixel, sample) is the same path.
 *   "moments"         1 = keep per-pixel sums of squared sample contributions beside the accumulation buffer (image z-test,
 *                     SURVEY 8d); de_get_moment2 returns the buffer; de_reset clears it
 *   "timeline"        1 = record the wavefront kernel's launch timeline (de_get_launch_timeline); the instrumented build records
 *                     it, so de_set_counting(ctx, 1) must be on as well
 *   "linear_textures" 0 = release the row-major texture copies (read by the parity flavour and the hooks only) */
int de_set_option(de_ctx *ctx, const char *name, int value);

/* the eleven set_* kernels + direct field writes  renderer.py:224-266, earth_viewer.py:308-314 */
int de_set_params(de_ctx *ctx, const DeParams *p);

/* texture load + copy_*_texture kernels  renderer.py:61-94,136-143,170-210 (host pointer) */
int de_upload_texture(de_ctx *ctx, int slot, const uint8_t *host_texels, int w, int h, int channels);
/* LUT load + copy_CIE_LUT_texture / copy_CRF_LUT_texture  renderer.py:97-134,144-145,212-222
 *   cie [2][441][3] f32 (LUT/CIE.dat), srgb2spec [300][3] fp16 bits, o3 [441] f32, crf [n_crf][1024][3] f32 */
int de_upload_luts(de_ctx *ctx, const float *cie, const uint16_t *srgb2spec_f16, const float *o3, const float *crf, int n_crf);

/* Renderer.reset_framebuffer  renderer.py:367-369 */
int de_reset(de_ctx *ctx);
/* Renderer.accumulate == Renderer.render kernel  renderer.py:283-330,371-380, generalised:
 * n_spp samples per pixel with sample indices [first_sample, first_sample+n_spp) over the pixel
 * window [x0,x0+w) x [y0,y0+h) (whole frame: 0,0,W,H).  Adds into the ctx accumulation buffer. */
int de_accumulate(de_ctx *ctx, int n_spp, uint32_t seed, uint32_t first_sample, int x0, int y0, int w, int h);
/* Multi-GPU tile partition (the film tile is the reference's 16x8 block, renderer.py:43-46,304-305): the whole frame,
 * but only the film tiles t = ty * ceil(W/16) + tx with t % tile_stride == tile_offset.  Ranks with different offsets write
 * disjoint pixels; combine with an spp slice through first_sample / n_spp (e.g. 8 GPUs = 2 tile groups x 4 sample slices). */
int de_accumulate_tiles(de_ctx *ctx, int n_spp, uint32_t seed, uint32_t first_sample, int tile_stride, int tile_offset);
/* device pointer of the accumulation buffer ([H][W][3] f32 linear sRGB sums; color_buffer,
 * renderer.py:25,330) so the caller can view it as a tensor / hand it to NCCL */
int de_get_accum(de_ctx *ctx, float **dev_ptr);
/* second-moment buffer ([H][W][3] f32 sums of squared per-sample RGB contributions), option "moments" */
int de_get_moment2(de_ctx *ctx, float **dev_ptr);
/* Renderer.fetch_image == _render_to_image kernel  renderer.py:346-365,382-384.
 * dev_out: device [H][W][3] f32 in [0,1].  accum_override (device, may be NULL) resolves another
 * buffer of the same shape, e.g. an NCCL-reduced one. */
int de_resolve(de_ctx *ctx, const float *accum_override, float *dev_out, int spp_total);
/* ---- multi-GPU: resolve fused with the accumulation exchange (one process per GPU) -------------------
 * The reference is single-device; its film buffer is linear in the samples (renderer.py:329-330), so
 * ranks that rendered disjoint sample slices only need their buffers SUMMED before _render_to_image
 * (renderer.py:346-365).  Either reduce them with NCCL and call de_resolve, or let the resolving rank read the
 * other ranks' buffers in place over NVLink peer memory:
 *   every rank:      de_ipc_export_accum(ctx, handle)            64-byte CUDA IPC handle of its buffer;
 *                    exchange the handles (torch.distributed.all_gather_object / any channel);
 *   resolving rank:  de_ipc_open_peer(ctx, handle_k, &ptr_k)     for every OTHER rank k (same node);
 *                    de_resolve_peers(ctx, ptrs, n, out, spp)    out = tonemap((own + sum_k ptr_k) / spp);
 *                    de_ipc_close_peers(ctx).
 * The other ranks must have finished (stream-synchronised + a barrier) before de_resolve_peers runs and must
 * keep their buffers untouched until it has completed.  peer pointers may be any device-accessible [H][W][3]
 * float buffers (IPC-opened, peer-enabled, or on the same device); at most 15. */
int de_ipc_export_accum(de_ctx *ctx, void *handle64);
int de_ipc_open_peer(de_ctx *ctx, const void *handle64, float **dev_ptr);
int de_ipc_close_peers(de_ctx *ctx);
int de_resolve_peers(de_ctx *ctx, const float *const *peer_accums, int n_peers, float *dev_out, int spp_total);
/* de_resolve_peers for a tile (+ spp) partition: peer k rendered the tiles with t % tile_stride == peer_tile_offsets[k], this rank
 * those with own_tile_offset; a pixel only reads the buffers of the ranks that rendered its tile (1/stride of the peer traffic). */
int de_resolve_peers_tiled(de_ctx *ctx, const float *const *peer_accums, const int *peer_tile_offsets, int n_peers, int tile_stride, int own_tile_offset,
                           float *dev_out, int spp_total);
/* convenience for non-torch callers: resolve + copy to host memory, synchronous */
int de_fetch_image_host(de_ctx *ctx, float *host_out, int spp_total);
int de_sync(de_ctx *ctx);
int de_get_counters(de_ctx *ctx, DeCounters *out); /* synchronises */
int de_set_counting(de_ctx *ctx, int enabled);     /* counters cost atomics: off by default */
/* scheduler self-profile of the last counting de_accumulate (wavefront mode): out32[3*s+{0,1,2}] = warp cycles,
 * visits and slots handled by stage s (NEW, SDF, RMO, CLOUD, SDF_DONE, RMO_DONE, EVENT, NEE_DONE, SURFACE), out32[27] = idle cycles */
int de_get_stage_profile(de_ctx *ctx, uint64_t *out32);

/* launch timeline of the last de_accumulate in wavefront mode with option "timeline" (synchronises): globaltimer ns
 * out8 = {first CTA start, first CTA to find the work counter exhausted, last CTA to, first CTA end, last CTA end,
 *         min chunks claimed by a CTA, max chunks, (space tiles << 32) | tiles rendered by the persistent kernel} */
int de_get_launch_timeline(de_ctx *ctx, uint64_t *out8);

/* measurement aid for the bench line's texel-rate roofline: sustained rate of the integrator's own fetch instruction (tex2Dgather on the
 * block-linear copy of texture `slot`, footprints L1-resident, all SMs saturated), in gather requests per second (4 texels each) */
int de_bench_tex_gather(de_ctx *ctx, int slot, int iters, double *gathers_per_second);
/* per-CTA drain diagnostics of the same launch: out[24 * k + ...] = {ns when CTA k found the counter exhausted, ns when fewer than 64
 * of its paths were alive, ns at its end, chunks claimed, stage visits[9], slots handled[9] after exhaustion}; returns the number
 * of CTAs written (>= 0) or a DE_ERR_* code */
int de_get_cta_timeline(de_ctx *ctx, uint64_t *out, int max_ctas);

/* ---- test hooks: the deterministic sub-paths of SURVEY 8(a), DEVICE pointers, n items --------
 * Each evaluates the IEEE source-order (parity) flavour of one reference function. */
int de_test_philox(de_ctx *, const uint32_t *ctr4_key2, uint32_t *out4, int n);
int de_test_rsi(de_ctx *, const float *pos3, const float *dir3, const float *r, float *out2, int n);        /* math_utils.py:17 */
int de_test_density(de_ctx *, const float *h, float *out3, int n);                                           /* volume_rendering_models.py:270 */
int de_test_spectra(de_ctx *, const float *wavelength, float *out5, int n);                                  /* :194-224, colour.py:51 */
int de_test_phase_eval(de_ctx *, const float *ray3, const float *light3, const int32_t *id, const int32_t *reduce, float *out, int n); /* pathtracer.py:235 */
int de_test_phase_sample(de_ctx *, const float *ray3, const int32_t *id, const int32_t *reduce, const uint32_t *rand4, float *out_dir3, float *out_w, int n); /* pathtracer.py:249 */
int de_test_dir_sample(de_ctx *, int kind, const float *n3, float cos_max, const uint32_t *rand2, float *out3, int n); /* sampling.py:25,30 */
int de_test_brdf(de_ctx *, const float *albedo, const float *ocean, const float *bathy, const float *v3, const float *n3, const float *l3, float *out2, int n); /* surface_rendering_models.py:9 */
int de_test_srgb2spec(de_ctx *, const float *rgb3, const float *wavelength, float *out, int n);              /* colour.py:62 */
int de_test_spectrum_sample(de_ctx *, const uint32_t *rand, float *out5, int n);                             /* colour.py:12 */
int de_test_tex_fetch(de_ctx *, int slot, const float *pos3, float *out4, int n);                            /* math_utils.py:38 */
int de_test_cast_dir(de_ctx *, const float *u, const float *v, const uint32_t *rand2, float *out3, int n);   /* renderer.py:269 */
int de_test_opendrt(de_ctx *, const float *rgb3, float *out3, int n);                                        /* OpenDRT.py:221 */
int de_test_agx(de_ctx *, const float *rgb3, float *out3, int n);                                            /* AgX.py:131 */
int de_test_crf(de_ctx *, const float *rgb3, float *out3, int n);                                            /* renderer.py:333 */
int de_test_srgb_oetf(de_ctx *, const float *x, float *out, int n);                                          /* colour.py:74 */
int de_test_intersect_land(de_ctx *, const float *pos3, const float *dir3, float *out, int n);               /* pathtracer.py:27 */
int de_test_land_normal(de_ctx *, const float *pos3, float *out3, int n);                                    /* pathtracer.py:16 */
int de_test_land_material(de_ctx *, const float *pos3, float *out6, int n);                                  /* pathtracer.py:284 */
int de_test_cloud_limits(de_ctx *, const float *pos3, const float *dir3, const float *land, float *out2, int n); /* pathtracer.py:145 */
int de_test_clouds_density(de_ctx *, const float *pos3, float *out, int n);                                  /* pathtracer.py:48 */
int de_test_raymarch_T(de_ctx *, const float *pos3, const float *dir3, const float *ext3, float *out, int n);/* pathtracer.py:471 */
/* kind 0: sample_interaction -> (event, t, id); kind 1: sample_transmittance -> (T, 0, 0); Philox key (seed, i), bounce 1 */
int de_test_tracking(de_ctx *, int kind, const float *pos3, const float *dir3, const float *land, const float *wavelength, uint32_t seed, float *out3, int n); /* pathtracer.py:172,211 */
/* individual path samples in PARITY arithmetic: out5 = rgb contribution, wavelength, radiance   renderer.py:305-330 */
int de_test_ray_march(de_ctx *, const float *pos3, const float *dir3, const float *t0, const float *t1, const float *sun3, const float *wavelength, float *out2, int n); /* pathtracer.py:501-541 */
int de_test_trace_preview(de_ctx *, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out5, int n); /* renderer.py:305-330 with pathtracer.py:543-685 */
/* ---- hooks on the PRODUCT flavour's work-removal bounds: the very device functions the wavefront kernel calls
 * (csrc/de_device.cuh), so their validity can be checked ray by ray against dense samples of the oracle -------------- */
/* cloud_pass_setup over [t_start, t_max]: out4 = (texture bound c_max, density bound (0 = skip the pass), t_start', t_max') */
int de_test_fast_cloud_bound(de_ctx *, const float *pos3, const float *dir3, const float *t_start, const float *t_max, float *out4, int n);
/* rmo_segment_majorant: bound of sigma.rho (extinctions ext3) over [t_start, t_max] */
int de_test_fast_rmo_majorant(de_ctx *, const float *pos3, const float *dir3, const float *t_start, const float *t_max, const float *ext3, float *out, int n);
/* the altitude-band walk of an rmo pass (rmo_band_walk, as the wavefront loop runs it): t_query[n][n_query] ascending ray parameters in
 * [t_start, t_max]; out[n][n_query] = the majorant in force when the walk is at that parameter */
int de_test_fast_rmo_bands(de_ctx *, const float *pos3, const float *dir3, const float *t_start, const float *t_max, const float *ext3, const float *t_query,
                           int n_query, float *out, int n);
/* product-flavour intersect_land: out3 = (1 if land_surely_missed fired, distance or -1, SDF evaluations) */
int de_test_fast_land(de_ctx *, const float *pos3, const float *dir3, float *out3, int n);
int de_test_trace_paths(de_ctx *, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out5, int n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DE_API_H */
