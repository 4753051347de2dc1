// de_api.cu -- the C-ABI of libde.so (include/de_api.h): context, uploads, dispatch.
// Host code only; every piece of arithmetic on the render path runs in the device TUs.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/de_api.h"
#include "de_launch.h"
#include "de_wavefront.h"

struct de_ctx {
    int device = 0, W = 0, H = 0, mode = DE_MODE_WAVEFRONT;
    cudaStream_t stream = nullptr;
    bool have_params = false, have_luts = false, counting = false, derived_dirty = true;
    bool have_tex[DE_TEX_COUNT] = {};
    DeParams params{};
    DevScene scene{};
    uint8_t *d_tex[DE_TEX_COUNT] = {};
    cudaArray_t arr[DE_TEX_COUNT] = {};
    float *d_cie = nullptr, *d_s2s = nullptr, *d_o3 = nullptr, *d_crf = nullptr, *d_cdf = nullptr;
    LambdaRow *d_lam = nullptr;
    DevDerived *d_derived = nullptr;
    float *d_accum = nullptr, *d_image = nullptr;
    float *d_accum2 = nullptr;                     // optional per-pixel second moments (de_set_option "moments")
    bool opt_space_tiles = true, opt_space_async = true, opt_timeline = false, opt_tile_order = true;
    unsigned long long param_version = 1;          // bumps with every de_set_params (tile classification cache key)
    uint8_t *d_cloud_max = nullptr;
    unsigned long long *d_counters = nullptr;
    DeWavefrontState *wf = nullptr;
    std::vector<void *> ipc_open;  // peer buffers opened with de_ipc_open_peer
    std::string err;
};

namespace {
// Density bounds per altitude band for the product flavour's rmo majorant (de_device.cuh, rmo_band_*): the reference's fits
// (volume_rendering_models.py:229-273) in double precision at the band's bottom (Rayleigh, aerosol: decreasing with altitude) and the
// maximum of the ozone fit over the band (it peaks at 25 km), each taken 100 m beyond the band (f32 positions of a far camera are
// +- 15 m off the ideal ray) and 0.2 % up (MUFU exponentials, the 1.3e-5 step of the aerosol fit at 11.5 km).
double fit_rayl(double h) { return 3.68082 * std::exp(-(h + 24239.99) * (h + 24239.99) / 532307548.4168) / 1.225; }
double fit_mie(double h) {
    double d;
    if (h > 11500.0) d = 0.0918 * std::exp(-1.0e-6 * (h - 11500.0) * (h - 11500.0));
    else if (h > 2400.0) d = 0.3000 * std::exp(-2.5e-9 * (h + 2500.0) * (h + 2500.0)) - 0.092;
    else if (h > 1300.0) d = 0.6500 * std::exp(-5.0e-6 * (h - 1300.0) * (h - 1300.0)) + 0.18899;
    else d = 1.0 - h / 8136.646;
    return d * 1.06;
}
double fit_ozone(double h) {
    const double k = h * 0.001, d2 = (k - 25.0) * (k - 25.0);
    const double c = -0.000015 * (k - 15.0) * (k - 15.0) * (k - 15.0);
    return 0.625 * std::exp(-d2 / 49.0) + 0.375 * std::exp(-d2 / 256.0) + (c > 0.0 ? c : 0.0);
}
void build_rmo_bands(DevScene &s) {
    const double edge[kDeRmoBands + 1] = {0.0, 4000.0, 12000.0, 30000.0, 110000.0};
    const double margin = 100.0, up = 1.002;
    for (int k = 0; k < kDeRmoBands; ++k) {
        const double lo = edge[k] > margin ? edge[k] - margin : 0.0, hi = edge[k + 1] + margin;
        s.band_r[k] = (float)(6371000.0 + edge[k]);
        double m_mie = 0.0, m_oz = 0.0, m_ray = 0.0;
        for (int j = 0; j <= 4000; ++j) {            // dense scan: no monotonicity assumed
            const double h = lo + (hi - lo) * j / 4000.0;
            m_ray = std::fmax(m_ray, fit_rayl(h)); m_mie = std::fmax(m_mie, fit_mie(h)); m_oz = std::fmax(m_oz, fit_ozone(h));
        }
        s.band_dr[k] = (float)(m_ray * up); s.band_dm[k] = (float)(m_mie * up); s.band_do[k] = (float)(m_oz * up);
    }
}
int fail(de_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(ctx, DE_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(e_)); \
    } while (0)
#define NEED(cond, msg) do { if (!(cond)) return fail(ctx, DE_ERR_INVALID, msg); } while (0)
#define ENTER()                                                  \
    if (!ctx) return DE_ERR_INVALID;                             \
    CU(cudaSetDevice(ctx->device))

int check_launch(de_ctx *ctx, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return DE_OK;
}

// keep the by-value kernel argument in sync with params + uploads
int refresh_scene(de_ctx *ctx) {
    DevScene &s = ctx->scene;
    const DeParams &p = ctx->params;
    s.cam_pos = make_float3(p.cam_pos[0], p.cam_pos[1], p.cam_pos[2]);
    s.look_at = make_float3(p.look_at[0], p.look_at[1], p.look_at[2]);
    s.up = make_float3(p.up[0], p.up[1], p.up[2]);
    s.fov = p.fov; s.aspect_scale = p.aspect_scale; s.sun_angle = p.sun_angle; s.sun_path_rot = p.sun_path_rot;
    s.aspect_ratio = (float)((double)ctx->W / (double)ctx->H);  // renderer.py:19 (Python float -> f32 at use)
    s.land_height_scale = p.land_height_scale; s.exposure = p.exposure; s.gamma = p.gamma;
    s.selected_crf = p.selected_crf; s.crf_count = p.crf_count > 0 ? p.crf_count : s.n_crf;
    s.vig_strength = p.vignette_strength; s.vig_radius = p.vignette_radius; s.vig_cx = p.vignette_center[0]; s.vig_cy = p.vignette_center[1];
    s.tonemapper = p.tonemapper;
    s.topo_tex_w = p.topo_tex_w > 0 ? p.topo_tex_w : s.tex[DE_TEX_TOPOGRAPHY].w;
    s.W = ctx->W; s.H = ctx->H;
    s.derived = ctx->d_derived; s.lam = ctx->d_lam; s.cdf = ctx->d_cdf;
    s.cie = ctx->d_cie; s.s2s = ctx->d_s2s; s.o3 = ctx->d_o3; s.crf = ctx->d_crf;
    s.counters = ctx->counting ? ctx->d_counters : nullptr;
    if (ctx->derived_dirty && ctx->have_params && s.topo_tex_w > 0) {
        de_exact::launch_prepare(s, ctx->d_derived, ctx->stream);
        int rc = check_launch(ctx, "prepare");
        if (rc) return rc;
        ctx->derived_dirty = false;
    }
    return DE_OK;
}
int ready_to_render(de_ctx *ctx) {
    if (!ctx->have_params) return fail(ctx, DE_ERR_STATE, "de_set_params has not been called");
    if (!ctx->have_luts) return fail(ctx, DE_ERR_STATE, "de_upload_luts has not been called");
    for (int i = 0; i < DE_TEX_COUNT; ++i) {
        if (!ctx->have_tex[i]) return fail(ctx, DE_ERR_STATE, "texture slot " + std::to_string(i) + " has not been uploaded");
        if (ctx->mode == DE_MODE_PARITY && !ctx->scene.tex[i].data)
            return fail(ctx, DE_ERR_STATE, "the parity flavour reads the row-major texture copies, which were released (option linear_textures)");
    }
    return refresh_scene(ctx);
}
}  // namespace

extern "C" {

int de_abi_version(void) { return DE_ABI_VERSION; }

int de_create(de_ctx **out, int device, int width, int height) {
    if (!out) return DE_ERR_INVALID;
    *out = nullptr;
    if (width <= 0 || height <= 0 || width % 16 != 0 || height % 8 != 0) return DE_ERR_INVALID;  // renderer.py:46
    de_ctx *ctx = new (std::nothrow) de_ctx();
    if (!ctx) return DE_ERR_NOMEM;
    ctx->device = device; ctx->W = width; ctx->H = height;
    build_rmo_bands(ctx->scene);
    auto bail = [&](cudaError_t e) { (void)e; de_destroy(ctx); return DE_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e);
    size_t npx = (size_t)width * height;
    if ((e = cudaMalloc(&ctx->d_accum, npx * 3 * sizeof(float))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_image, npx * 3 * sizeof(float))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_derived, sizeof(DevDerived))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_lam, sizeof(LambdaRow) * 512)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_cdf, sizeof(float) * 512)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->d_counters, sizeof(DeCounters))) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(ctx->d_accum, 0, npx * 3 * sizeof(float))) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(ctx->d_counters, 0, sizeof(DeCounters))) != cudaSuccess) return bail(e);
    *out = ctx;
    return DE_OK;
}

void de_destroy(de_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < DE_TEX_COUNT; ++i) {
        if (ctx->scene.tex[i].obj) cudaDestroyTextureObject(ctx->scene.tex[i].obj);
        if (ctx->arr[i]) cudaFreeArray(ctx->arr[i]);
        cudaFree(ctx->d_tex[i]);
    }
    de_wavefront_free(ctx->wf);
    for (void *p : ctx->ipc_open) cudaIpcCloseMemHandle(p);
    cudaFree(ctx->d_cie); cudaFree(ctx->d_s2s); cudaFree(ctx->d_o3); cudaFree(ctx->d_crf); cudaFree(ctx->d_cdf);
    cudaFree(ctx->d_cloud_max);
    cudaFree(ctx->d_lam); cudaFree(ctx->d_derived); cudaFree(ctx->d_accum); cudaFree(ctx->d_accum2); cudaFree(ctx->d_image); cudaFree(ctx->d_counters);
    delete ctx;
}

const char *de_last_error(de_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int de_set_stream(de_ctx *ctx, void *cuda_stream) {
    ENTER();
    cudaStream_t ns = (cudaStream_t)cuda_stream;
    if (ns != ctx->stream) {  // work already queued on the old stream (render, k_prepare) is ordered before anything on the new one
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e1 = cudaEventRecord(ev, ctx->stream), e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(ns, ev, 0) : e1;
        cudaEventDestroy(ev);
        if (e2 != cudaSuccess) return fail(ctx, DE_ERR_CUDA, std::string("de_set_stream: ") + cudaGetErrorString(e2));
        ctx->stream = ns;
    }
    return DE_OK;
}
int de_set_option(de_ctx *ctx, const char *name, int value) {
    ENTER();
    NEED(name, "name is NULL");
    const std::string n(name);
    if (n == "space_tiles") ctx->opt_space_tiles = value != 0;
    else if (n == "space_async") ctx->opt_space_async = value != 0;
    else if (n == "tile_order") ctx->opt_tile_order = value != 0;
    else if (n == "timeline") ctx->opt_timeline = value != 0;
    else if (n == "linear_textures") {
        // The row-major copies of the maps are read by the parity flavour and the test hooks only; the product integrators
        // sample the block-linear arrays.  0 frees them (2.3 GB at the NASA resolution); they come back with the next upload.
        NEED(value == 0, "linear_textures can only be released (0); upload the textures again to restore them");
        CU(cudaDeviceSynchronize());
        for (int i = 0; i < DE_TEX_COUNT; ++i) { cudaFree(ctx->d_tex[i]); ctx->d_tex[i] = nullptr; ctx->scene.tex[i].data = nullptr; }
    }
    else if (n == "moments") {
        if (value && !ctx->d_accum2) {
            size_t bytes = (size_t)ctx->W * ctx->H * 3 * sizeof(float);
            CU(cudaMalloc(&ctx->d_accum2, bytes));
            CU(cudaMemsetAsync(ctx->d_accum2, 0, bytes, ctx->stream));
        } else if (!value && ctx->d_accum2) {
            CU(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_accum2); ctx->d_accum2 = nullptr;
        }
    } else return fail(ctx, DE_ERR_INVALID, "unknown option '" + n + "'");
    return DE_OK;
}
int de_get_moment2(de_ctx *ctx, float **dev_ptr) {
    ENTER();
    NEED(dev_ptr, "dev_ptr is NULL");
    if (!ctx->d_accum2) return fail(ctx, DE_ERR_STATE, "second moments are off: de_set_option(ctx, \"moments\", 1) first");
    *dev_ptr = ctx->d_accum2;
    return DE_OK;
}
int de_get_launch_timeline(de_ctx *ctx, uint64_t *out8) {
    ENTER();
    NEED(out8, "out is NULL");
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ctx->wf) return fail(ctx, DE_ERR_STATE, "the wavefront integrator has not run yet");
    unsigned long long buf[40];
    if (de_wavefront_profile(ctx->wf, buf) != 0) return fail(ctx, DE_ERR_CUDA, "timeline copy failed");
    for (int k = 0; k < 7; ++k) out8[k] = buf[32 + k];
    unsigned int tc[2] = {0u, 0u};
    de_wavefront_tile_counts(ctx->wf, tc);
    out8[7] = ((uint64_t)tc[1] << 32) | tc[0];
    return DE_OK;
}
int de_set_mode(de_ctx *ctx, int mode) {
    ENTER();
    NEED(mode >= DE_MODE_WAVEFRONT && mode <= DE_MODE_PREVIEW, "unknown mode");
    ctx->mode = mode;
    return DE_OK;
}
int de_get_stage_profile(de_ctx *ctx, uint64_t *out32) {
    ENTER();
    NEED(out32, "out is NULL");
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ctx->wf) return fail(ctx, DE_ERR_STATE, "the wavefront integrator has not run yet");
    unsigned long long buf[40];
    if (de_wavefront_profile(ctx->wf, buf) != 0) return fail(ctx, DE_ERR_CUDA, "profile copy failed");
    for (int k = 0; k < 32; ++k) out32[k] = buf[k];
    return DE_OK;
}
int de_get_cta_timeline(de_ctx *ctx, uint64_t *out, int max_ctas) {
    if (!ctx) return DE_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return DE_ERR_CUDA;
    if (!out || max_ctas <= 0) return fail(ctx, DE_ERR_INVALID, "out is NULL / max_ctas <= 0");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(ctx, DE_ERR_CUDA, "sync failed");
    if (!ctx->wf) return fail(ctx, DE_ERR_STATE, "the wavefront integrator has not run yet");
    const int n = de_wavefront_cta_stats(ctx->wf, (unsigned long long *)out, max_ctas);
    return n >= 0 ? n : fail(ctx, DE_ERR_CUDA, "cta timeline copy failed");
}
int de_bench_tex_gather(de_ctx *ctx, int slot, int iters, double *gathers_per_second) {
    ENTER();
    NEED(slot >= 0 && slot < DE_TEX_COUNT && ctx->have_tex[slot] && ctx->scene.tex[slot].obj, "texture slot not uploaded");
    NEED(iters > 0 && gathers_per_second, "bad arguments");
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const int ctas = sms * 8;                                       // 2048 threads per SM: every TEX unit saturated
    float *scratch = nullptr;
    CU(cudaMalloc(&scratch, (size_t)ctas * 256 * sizeof(float)));
    const float ms = de_fast::bench_tex_gather(ctx->scene.tex[slot].obj, ctx->scene.tex[slot].w, ctx->scene.tex[slot].h, ctas, iters, scratch, ctx->stream);
    cudaFree(scratch);
    if (!(ms > 0.0f)) return fail(ctx, DE_ERR_CUDA, "tex gather microbenchmark failed");
    *gathers_per_second = (double)ctas * 256.0 * (double)iters / ((double)ms * 1e-3);
    return check_launch(ctx, "tex_gather_peak");
}
int de_set_counting(de_ctx *ctx, int enabled) {
    ENTER();
    ctx->counting = enabled != 0;
    return DE_OK;
}

int de_set_params(de_ctx *ctx, const DeParams *p) {
    ENTER();
    NEED(p, "params is NULL");
    NEED(p->selected_crf >= 0, "selected_crf < 0");
    ctx->params = *p;
    ctx->have_params = true;
    ctx->derived_dirty = true;
    ++ctx->param_version;
    return DE_OK;
}

int de_upload_texture(de_ctx *ctx, int slot, const uint8_t *host, int w, int h, int channels) {
    ENTER();
    NEED(slot >= 0 && slot < DE_TEX_COUNT, "bad texture slot");
    NEED(host && w > 0 && h > 0, "bad texture");
    bool rgb = slot == DE_TEX_ALBEDO || slot == DE_TEX_STARS;
    NEED(channels == (rgb ? 3 : 1), "albedo/stars take 3 channels, the other maps 1");
    size_t bytes = (size_t)w * h * channels;
    DevTex &t = ctx->scene.tex[slot];
    CU(cudaStreamSynchronize(ctx->stream));  // a render in flight may still sample the old texture (destroying an object does not synchronise)
    if (ctx->wf) CU(cudaDeviceSynchronize());
    ctx->have_tex[slot] = false;             // the slot is empty until the new map is completely in place
    if (t.obj) { cudaDestroyTextureObject(t.obj); t.obj = 0; }
    if (ctx->arr[slot]) { cudaFreeArray(ctx->arr[slot]); ctx->arr[slot] = nullptr; }
    cudaFree(ctx->d_tex[slot]); ctx->d_tex[slot] = nullptr;
    t.data = nullptr; t.w = t.h = t.c = 0;
    if (slot == DE_TEX_CLOUDS) { ctx->scene.cloud_max = nullptr; }
    CU(cudaMalloc(&ctx->d_tex[slot], bytes));
    CU(cudaMemcpyAsync(ctx->d_tex[slot], host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    // block-linear copy behind a point-sampled, clamped, unnormalised-coordinate texture object:
    // the TEX path of the wavefront integrator gathers the 2x2 footprint with one instruction.
    cudaChannelFormatDesc fd = rgb ? cudaCreateChannelDesc<uchar4>() : cudaCreateChannelDesc<unsigned char>();
    CU(cudaMallocArray(&ctx->arr[slot], &fd, (size_t)w, (size_t)h, cudaArrayTextureGather));
    if (rgb) {
        std::vector<uint8_t> rgba((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; ++i) { rgba[4 * i] = host[3 * i]; rgba[4 * i + 1] = host[3 * i + 1]; rgba[4 * i + 2] = host[3 * i + 2]; rgba[4 * i + 3] = 0; }
        CU(cudaMemcpy2DToArray(ctx->arr[slot], 0, 0, rgba.data(), (size_t)w * 4, (size_t)w * 4, (size_t)h, cudaMemcpyHostToDevice));
    } else {
        CU(cudaMemcpy2DToArray(ctx->arr[slot], 0, 0, host, (size_t)w, (size_t)w, (size_t)h, cudaMemcpyHostToDevice));
    }
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = ctx->arr[slot];
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;  // unorm8 -> [0,1] in the TEX unit
    td.normalizedCoords = 0;
    CU(cudaCreateTextureObject(&t.obj, &rd, &td, nullptr));
    t.data = ctx->d_tex[slot]; t.w = w; t.h = h; t.c = channels;
    ctx->have_tex[slot] = true;
    if (slot == DE_TEX_CLOUDS) {  // coarse max-map for the local tracking majorant (product flavour)
        int b = w / 256 > 8 ? w / 256 : 8, cw = (w + b - 1) / b, ch = (h + b - 1) / b;
        cudaFree(ctx->d_cloud_max); ctx->d_cloud_max = nullptr;
        CU(cudaMalloc(&ctx->d_cloud_max, (size_t)cw * ch));
        de_fast::launch_build_cloud_max(ctx->d_tex[slot], w, h, b, ctx->d_cloud_max, cw, ch, ctx->stream);
        int rc_ = check_launch(ctx, "build_cloud_max");
        if (rc_) return rc_;
        ctx->scene.cloud_max = ctx->d_cloud_max; ctx->scene.cm_w = cw; ctx->scene.cm_h = ch; ctx->scene.cm_b = b;
        ++ctx->param_version;  // the work order of the wavefront kernel looks at the cloud map (tile classification cache key)
    }
    if (slot == DE_TEX_TOPOGRAPHY) ctx->derived_dirty = true;
    CU(cudaStreamSynchronize(ctx->stream));  // host buffer may be released by the caller
    return DE_OK;
}

int de_upload_luts(de_ctx *ctx, const float *cie, const uint16_t *s2s, const float *o3, const float *crf, int n_crf) {
    ENTER();
    NEED(cie && s2s && o3 && crf && n_crf > 0, "bad LUT arguments");
    // rgba16f CIE texture (renderer.py:97,212-216): quantise to fp16 once, keep as f32
    std::vector<float> cie_q(2 * 441 * 3), s2s_f(300 * 3);
    for (size_t i = 0; i < cie_q.size(); ++i) cie_q[i] = __half2float(__float2half_rn(cie[i]));
    for (size_t i = 0; i < s2s_f.size(); ++i) { __half_raw r; r.x = s2s[i]; s2s_f[i] = __half2float(__half(r)); }
    cudaFree(ctx->d_cie); cudaFree(ctx->d_s2s); cudaFree(ctx->d_o3); cudaFree(ctx->d_crf);
    ctx->d_cie = ctx->d_s2s = ctx->d_o3 = ctx->d_crf = nullptr;
    CU(cudaMalloc(&ctx->d_cie, cie_q.size() * 4)); CU(cudaMalloc(&ctx->d_s2s, s2s_f.size() * 4));
    CU(cudaMalloc(&ctx->d_o3, 441 * 4)); CU(cudaMalloc(&ctx->d_crf, (size_t)n_crf * 1024 * 3 * 4));
    CU(cudaMemcpy(ctx->d_cie, cie_q.data(), cie_q.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_s2s, s2s_f.data(), s2s_f.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_o3, o3, 441 * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_crf, crf, (size_t)n_crf * 1024 * 3 * 4, cudaMemcpyHostToDevice));
    ctx->scene.n_crf = n_crf;
    ctx->scene.cie = ctx->d_cie; ctx->scene.s2s = ctx->d_s2s; ctx->scene.o3 = ctx->d_o3; ctx->scene.crf = ctx->d_crf;
    de_exact::launch_build_lambda(ctx->scene, ctx->d_lam, ctx->d_cdf, ctx->stream);
    int rc = check_launch(ctx, "build_lambda");
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_luts = true;
    return DE_OK;
}

int de_reset(de_ctx *ctx) {
    ENTER();
    CU(cudaMemsetAsync(ctx->d_accum, 0, (size_t)ctx->W * ctx->H * 3 * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(DeCounters), ctx->stream));
    if (ctx->d_accum2) CU(cudaMemsetAsync(ctx->d_accum2, 0, (size_t)ctx->W * ctx->H * 3 * sizeof(float), ctx->stream));
    return DE_OK;
}

static int accumulate_impl(de_ctx *ctx, int n_spp, uint32_t seed, uint32_t first_sample, int x0, int y0, int w, int h, int stride, int offset) {
    NEED(n_spp > 0, "n_spp <= 0");
    NEED(x0 >= 0 && y0 >= 0 && w > 0 && h > 0 && x0 + w <= ctx->W && y0 + h <= ctx->H, "window outside the frame");
    NEED(stride >= 1 && offset >= 0 && offset < stride, "tile partition: need 0 <= offset < stride");
    int rc = ready_to_render(ctx);
    if (rc) return rc;
    float *a2 = ctx->d_accum2;
    if (ctx->mode == DE_MODE_PARITY) de_exact::launch_render_mega(ctx->scene, ctx->d_accum, a2, n_spp, seed, first_sample, x0, y0, w, h, stride, offset, ctx->counting, ctx->stream);
    else if (ctx->mode == DE_MODE_PREVIEW) de_fast::launch_render_preview(ctx->scene, ctx->d_accum, a2, n_spp, seed, first_sample, x0, y0, w, h, stride, offset, ctx->counting, ctx->stream);
    else if (ctx->mode == DE_MODE_MEGAKERNEL) de_fast::launch_render_mega(ctx->scene, ctx->d_accum, a2, n_spp, seed, first_sample, x0, y0, w, h, stride, offset, ctx->counting, ctx->stream);
    else {
        if (!ctx->wf) {
            ctx->wf = de_wavefront_alloc(ctx->device);
            if (!ctx->wf) return fail(ctx, DE_ERR_NOMEM, "wavefront state allocation failed");
        }
        DeWavefrontJob job;
        job.accum = ctx->d_accum; job.accum2 = a2; job.n_spp = n_spp; job.seed = seed; job.first_sample = first_sample;
        job.x0 = x0; job.y0 = y0; job.w = w; job.h = h;
        job.count = ctx->counting; job.timeline = ctx->opt_timeline;
        job.space_tiles = ctx->opt_space_tiles; job.space_async = ctx->opt_space_async; job.tile_order = ctx->opt_tile_order;
        job.param_version = ctx->param_version;
        job.tile_stride = stride; job.tile_offset = offset;
        if (de_wavefront_render(ctx->wf, ctx->scene, job, ctx->stream) != 0) return fail(ctx, DE_ERR_NOMEM, "wavefront tile buffers: allocation failed");
    }
    return check_launch(ctx, "render");
}
int de_accumulate(de_ctx *ctx, int n_spp, uint32_t seed, uint32_t first_sample, int x0, int y0, int w, int h) {
    ENTER();
    return accumulate_impl(ctx, n_spp, seed, first_sample, x0, y0, w, h, 1, 0);
}
int de_accumulate_tiles(de_ctx *ctx, int n_spp, uint32_t seed, uint32_t first_sample, int tile_stride, int tile_offset) {
    ENTER();
    return accumulate_impl(ctx, n_spp, seed, first_sample, 0, 0, ctx->W, ctx->H, tile_stride, tile_offset);
}

int de_get_accum(de_ctx *ctx, float **dev_ptr) {
    ENTER();
    NEED(dev_ptr, "dev_ptr is NULL");
    *dev_ptr = ctx->d_accum;
    return DE_OK;
}

int de_resolve(de_ctx *ctx, const float *accum_override, float *dev_out, int spp_total) {
    ENTER();
    NEED(dev_out, "dev_out is NULL");
    NEED(spp_total > 0, "spp_total <= 0");
    if (!ctx->have_params || !ctx->have_luts) return fail(ctx, DE_ERR_STATE, "params / LUTs missing");
    int rc = refresh_scene(ctx);
    if (rc) return rc;
    const float *src = accum_override ? accum_override : ctx->d_accum;
    // _render_to_image always runs in IEEE source-order arithmetic (the 1e-5 tonemap gate of the north star holds for every
    // mode): the kernel is 25 MB in / 25 MB out, fast intrinsics would buy nothing
    de_exact::launch_resolve(ctx->scene, src, dev_out, spp_total, ctx->stream);
    return check_launch(ctx, "resolve");
}

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI documents a 64-byte handle");
int de_ipc_export_accum(de_ctx *ctx, void *handle64) {
    ENTER();
    NEED(handle64, "handle64 is NULL");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_accum));
    std::memcpy(handle64, &h, sizeof(h));
    return DE_OK;
}

int de_ipc_open_peer(de_ctx *ctx, const void *handle64, float **dev_ptr) {
    ENTER();
    NEED(handle64 && dev_ptr, "handle64 / dev_ptr is NULL");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc_open.push_back(p);
    *dev_ptr = static_cast<float *>(p);
    return DE_OK;
}

int de_ipc_close_peers(de_ctx *ctx) {
    ENTER();
    CU(cudaStreamSynchronize(ctx->stream));  // a resolve may still be reading them
    int rc = DE_OK;
    for (void *p : ctx->ipc_open) {
        cudaError_t e = cudaIpcCloseMemHandle(p);
        if (e != cudaSuccess) rc = fail(ctx, DE_ERR_CUDA, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e));
    }
    ctx->ipc_open.clear();
    return rc;
}

int de_resolve_peers_tiled(de_ctx *ctx, const float *const *peer_accums, const int *peer_tile_offsets, int n_peers, int tile_stride, int own_tile_offset,
                           float *dev_out, int spp_total) {
    ENTER();
    NEED(dev_out, "dev_out is NULL");
    NEED(spp_total > 0, "spp_total <= 0");
    NEED(n_peers >= 0 && n_peers <= kDeMaxPeers, "n_peers out of range (0..15)");
    NEED(n_peers == 0 || peer_accums, "peer_accums is NULL");
    NEED(tile_stride >= 1 && own_tile_offset >= 0 && own_tile_offset < tile_stride, "tile partition: need 0 <= offset < stride");
    NEED(tile_stride == 1 || n_peers == 0 || peer_tile_offsets, "peer_tile_offsets is NULL");
    for (int k = 0; k < n_peers; ++k) {
        NEED(peer_accums[k], "a peer pointer is NULL");
        NEED(tile_stride == 1 || (peer_tile_offsets[k] >= 0 && peer_tile_offsets[k] < tile_stride), "a peer tile offset is out of range");
    }
    if (!ctx->have_params || !ctx->have_luts) return fail(ctx, DE_ERR_STATE, "params / LUTs missing");
    int rc = refresh_scene(ctx);
    if (rc) return rc;
    de_exact::launch_resolve_peers(ctx->scene, ctx->d_accum, peer_accums, peer_tile_offsets, n_peers, tile_stride, own_tile_offset, dev_out, spp_total, ctx->stream);
    return check_launch(ctx, "resolve_peers");
}
int de_resolve_peers(de_ctx *ctx, const float *const *peer_accums, int n_peers, float *dev_out, int spp_total) {
    return de_resolve_peers_tiled(ctx, peer_accums, nullptr, n_peers, 1, 0, dev_out, spp_total);
}

int de_fetch_image_host(de_ctx *ctx, float *host_out, int spp_total) {
    ENTER();
    NEED(host_out, "host_out is NULL");
    int rc = de_resolve(ctx, nullptr, ctx->d_image, spp_total);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_out, ctx->d_image, (size_t)ctx->W * ctx->H * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

int de_sync(de_ctx *ctx) {
    ENTER();
    CU(cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

int de_get_counters(de_ctx *ctx, DeCounters *out) {
    ENTER();
    NEED(out, "out is NULL");
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, ctx->d_counters, sizeof(DeCounters), cudaMemcpyDeviceToHost));
    return DE_OK;
}

// ---------------------------------------------------------------- test hooks
#define HOOK_PRE(needs_scene)                                                    \
    ENTER();                                                                     \
    NEED(n >= 0, "n < 0");                                                       \
    if (n == 0) return DE_OK;                                                    \
    if (needs_scene) {                                                           \
        int rc_ = refresh_scene(ctx); if (rc_) return rc_;                       \
        for (int i_ = 0; i_ < DE_TEX_COUNT; ++i_) NEED(!ctx->have_tex[i_] || ctx->scene.tex[i_].data, "row-major texture copies were released (option linear_textures)"); \
    }
#define HOOK_POST(name) return check_launch(ctx, name)

int de_test_philox(de_ctx *ctx, const uint32_t *in6, uint32_t *out4, int n) { HOOK_PRE(false); de_exact::t_philox(in6, out4, n, ctx->stream); HOOK_POST("philox"); }
int de_test_rsi(de_ctx *ctx, const float *pos, const float *dir, const float *r, float *out, int n) { HOOK_PRE(false); de_exact::t_rsi(pos, dir, r, out, n, ctx->stream); HOOK_POST("rsi"); }
int de_test_density(de_ctx *ctx, const float *h, float *out, int n) { HOOK_PRE(false); de_exact::t_density(h, out, n, ctx->stream); HOOK_POST("density"); }
int de_test_spectra(de_ctx *ctx, const float *wl, float *out, int n) { HOOK_PRE(true); de_exact::t_spectra(ctx->scene, wl, out, n, ctx->stream); HOOK_POST("spectra"); }
int de_test_phase_eval(de_ctx *ctx, const float *a, const float *b, const int32_t *id, const int32_t *red, float *out, int n) { HOOK_PRE(false); de_exact::t_phase_eval(a, b, id, red, out, n, ctx->stream); HOOK_POST("phase_eval"); }
int de_test_phase_sample(de_ctx *ctx, const float *a, const int32_t *id, const int32_t *red, const uint32_t *rand, float *od, float *ow, int n) { HOOK_PRE(false); de_exact::t_phase_sample(a, id, red, rand, od, ow, n, ctx->stream); HOOK_POST("phase_sample"); }
int de_test_dir_sample(de_ctx *ctx, int kind, const float *nrm, float cmax, const uint32_t *rand, float *out, int n) { HOOK_PRE(false); de_exact::t_dir_sample(kind, nrm, cmax, rand, out, n, ctx->stream); HOOK_POST("dir_sample"); }
int de_test_brdf(de_ctx *ctx, const float *al, const float *oc, const float *ba, const float *v, const float *nr, const float *l, float *out, int n) { HOOK_PRE(false); de_exact::t_brdf(al, oc, ba, v, nr, l, out, n, ctx->stream); HOOK_POST("brdf"); }
int de_test_srgb2spec(de_ctx *ctx, const float *rgb, const float *wl, float *out, int n) { HOOK_PRE(true); de_exact::t_srgb2spec(ctx->scene, rgb, wl, out, n, ctx->stream); HOOK_POST("srgb2spec"); }
int de_test_spectrum_sample(de_ctx *ctx, const uint32_t *rand, float *out, int n) { HOOK_PRE(true); de_exact::t_spectrum_sample(ctx->scene, rand, out, n, ctx->stream); HOOK_POST("spectrum_sample"); }
int de_test_tex_fetch(de_ctx *ctx, int slot, const float *pos, float *out, int n) {
    HOOK_PRE(true);
    NEED(slot >= 0 && slot < DE_TEX_COUNT && ctx->have_tex[slot], "texture slot not uploaded");
    de_exact::t_tex_fetch(ctx->scene, slot, pos, out, n, ctx->stream); HOOK_POST("tex_fetch");
}
int de_test_cast_dir(de_ctx *ctx, const float *u, const float *v, const uint32_t *rand, float *out, int n) { HOOK_PRE(true); de_exact::t_cast_dir(ctx->scene, u, v, rand, out, n, ctx->stream); HOOK_POST("cast_dir"); }
int de_test_opendrt(de_ctx *ctx, const float *rgb, float *out, int n) { HOOK_PRE(false); de_exact::t_opendrt(rgb, out, n, ctx->stream); HOOK_POST("opendrt"); }
int de_test_agx(de_ctx *ctx, const float *rgb, float *out, int n) { HOOK_PRE(false); de_exact::t_agx(rgb, out, n, ctx->stream); HOOK_POST("agx"); }
int de_test_crf(de_ctx *ctx, const float *rgb, float *out, int n) { HOOK_PRE(true); de_exact::t_crf(ctx->scene, rgb, out, n, ctx->stream); HOOK_POST("crf"); }
int de_test_srgb_oetf(de_ctx *ctx, const float *x, float *out, int n) { HOOK_PRE(false); de_exact::t_srgb_oetf(x, out, n, ctx->stream); HOOK_POST("srgb_oetf"); }
int de_test_intersect_land(de_ctx *ctx, const float *pos, const float *dir, float *out, int n) { HOOK_PRE(true); de_exact::t_intersect_land(ctx->scene, pos, dir, out, n, ctx->stream); HOOK_POST("intersect_land"); }
int de_test_land_normal(de_ctx *ctx, const float *pos, float *out, int n) { HOOK_PRE(true); de_exact::t_land_normal(ctx->scene, pos, out, n, ctx->stream); HOOK_POST("land_normal"); }
int de_test_land_material(de_ctx *ctx, const float *pos, float *out, int n) { HOOK_PRE(true); de_exact::t_land_material(ctx->scene, pos, out, n, ctx->stream); HOOK_POST("land_material"); }
int de_test_cloud_limits(de_ctx *ctx, const float *pos, const float *dir, const float *land, float *out, int n) { HOOK_PRE(false); de_exact::t_cloud_limits(pos, dir, land, out, n, ctx->stream); HOOK_POST("cloud_limits"); }
int de_test_clouds_density(de_ctx *ctx, const float *pos, float *out, int n) { HOOK_PRE(true); de_exact::t_clouds_density(ctx->scene, pos, out, n, ctx->stream); HOOK_POST("clouds_density"); }
int de_test_raymarch_T(de_ctx *ctx, const float *pos, const float *dir, const float *ext, float *out, int n) { HOOK_PRE(false); de_exact::t_raymarch_T(pos, dir, ext, out, n, ctx->stream); HOOK_POST("raymarch_T"); }
int de_test_tracking(de_ctx *ctx, int kind, const float *pos, const float *dir, const float *land, const float *wl, uint32_t seed, float *out, int n) { HOOK_PRE(true); de_exact::t_tracking(ctx->scene, kind, pos, dir, land, wl, seed, out, n, ctx->stream); HOOK_POST("tracking"); }
int de_test_ray_march(de_ctx *ctx, const float *pos, const float *dir, const float *t0, const float *t1, const float *sun, const float *wl, float *out2, int n) {
    HOOK_PRE(true); de_exact::t_ray_march(ctx->scene, pos, dir, t0, t1, sun, wl, out2, n, ctx->stream); HOOK_POST("ray_march");
}
int de_test_trace_preview(de_ctx *ctx, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out, int n) {
    ENTER();
    if (n <= 0) return DE_OK;
    int rc = ready_to_render(ctx);
    if (rc) return rc;
    de_exact::t_trace_preview(ctx->scene, px, py, sample, seed, out, n, ctx->stream);
    HOOK_POST("trace_preview");
}
int de_test_fast_cloud_bound(de_ctx *ctx, const float *pos, const float *dir, const float *ts, const float *tm, float *out4, int n) {
    HOOK_PRE(true);
    NEED(ctx->have_tex[DE_TEX_CLOUDS], "cloud texture not uploaded");
    de_fast::t_fast_cloud_bound(ctx->scene, pos, dir, ts, tm, out4, n, ctx->stream); HOOK_POST("fast_cloud_bound");
}
int de_test_fast_rmo_majorant(de_ctx *ctx, const float *pos, const float *dir, const float *ts, const float *tm, const float *ext3, float *out, int n) {
    HOOK_PRE(false); de_fast::t_fast_rmo_majorant(pos, dir, ts, tm, ext3, out, n, ctx->stream); HOOK_POST("fast_rmo_majorant");
}
int de_test_fast_rmo_bands(de_ctx *ctx, const float *pos, const float *dir, const float *ts, const float *tm, const float *ext3, const float *t_query, int n_query,
                           float *out, int n) {
    HOOK_PRE(false);
    NEED(n_query > 0 && t_query, "bad query arguments");
    int rc = refresh_scene(ctx);
    if (rc) return rc;
    de_fast::t_fast_rmo_bands(ctx->scene, pos, dir, ts, tm, ext3, t_query, n_query, out, n, ctx->stream); HOOK_POST("fast_rmo_bands");
}
int de_test_fast_land(de_ctx *ctx, const float *pos, const float *dir, float *out3, int n) {
    HOOK_PRE(true);
    NEED(ctx->have_tex[DE_TEX_TOPOGRAPHY], "topography texture not uploaded");
    de_fast::t_fast_land(ctx->scene, pos, dir, out3, n, ctx->stream); HOOK_POST("fast_land");
}
int de_test_trace_paths(de_ctx *ctx, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out, int n) {
    ENTER();
    if (n <= 0) return DE_OK;
    int rc = ready_to_render(ctx);
    if (rc) return rc;
    de_exact::t_trace_paths(ctx->scene, px, py, sample, seed, out, n, ctx->stream);
    HOOK_POST("trace_paths");
}

}  // extern "C"
