// de_launch.h -- host-callable launchers exported by the two device translation units.
#pragma once
#include "de_scene.h"

struct DeWavefrontState;  // de_wavefront.cuh
constexpr int kDeMaxPeers = 15;  // other ranks whose accumulation buffers one resolve can sum (16-GPU node)

#define DE_DECLARE_COMMON                                                                                                                \
    void launch_render_mega(const DevScene &s, float *accum, float *accum2, int n_spp, uint32_t seed, uint32_t first_sample, int x0, int y0, \
                            int w, int h, int tile_stride, int tile_offset, bool count, cudaStream_t st);                                \
    void launch_render_preview(const DevScene &s, float *accum, float *accum2, int n_spp, uint32_t seed, uint32_t first_sample, int x0,     \
                               int y0, int w, int h, int tile_stride, int tile_offset, bool count, cudaStream_t st);

namespace de_fast {
DE_DECLARE_COMMON
void launch_build_cloud_max(const uint8_t *tex, int w, int h, int b, uint8_t *out, int cw, int ch, cudaStream_t st);
// hooks on the product flavour's work-removal bounds (device pointers)
void t_fast_cloud_bound(const DevScene &s, const float *pos, const float *dir, const float *ts, const float *tm, float *out4, int n, cudaStream_t st);
void t_fast_rmo_majorant(const float *pos, const float *dir, const float *ts, const float *tm, const float *ext, float *out, int n, cudaStream_t st);
void t_fast_land(const DevScene &s, const float *pos, const float *dir, float *out3, int n, cudaStream_t st);
void t_fast_rmo_bands(const DevScene &s, const float *pos, const float *dir, const float *ts, const float *tm, const float *ext, const float *tq, int nq, float *out, int n,
                      cudaStream_t st);
// ms of `ctas` x 256 threads x `iters` independent L1-resident tex2Dgather requests (-1 on error); scratch: ctas * 256 floats
float bench_tex_gather(cudaTextureObject_t obj, int w, int h, int ctas, int iters, float *scratch, cudaStream_t st);
}
namespace de_exact {
DE_DECLARE_COMMON
void launch_resolve(const DevScene &s, const float *accum, float *out, int spp, cudaStream_t st);
void launch_resolve_peers(const DevScene &s, const float *accum, const float *const *peers, const int *peer_offsets, int n_peers, int tile_stride, int own_offset,
                          float *out, int spp, cudaStream_t st);
void launch_prepare(const DevScene &s, DevDerived *out, cudaStream_t st);
void launch_build_lambda(const DevScene &s, LambdaRow *lam, float *cdf, cudaStream_t st);
// test hooks (parity arithmetic); all pointers are device pointers
void t_philox(const uint32_t *in6, uint32_t *out4, int n, cudaStream_t st);
void t_rsi(const float *pos, const float *dir, const float *r, float *out, int n, cudaStream_t st);
void t_density(const float *h, float *out, int n, cudaStream_t st);
void t_spectra(const DevScene &s, const float *wl, float *out, int n, cudaStream_t st);
void t_phase_eval(const float *a, const float *b, const int32_t *id, const int32_t *red, float *out, int n, cudaStream_t st);
void t_phase_sample(const float *a, const int32_t *id, const int32_t *red, const uint32_t *rand, float *od, float *ow, int n, cudaStream_t st);
void t_dir_sample(int kind, const float *nrm, float cmax, const uint32_t *rand, float *out, int n, cudaStream_t st);
void t_brdf(const float *al, const float *oc, const float *ba, const float *v, const float *nr, const float *l, float *out, int n, cudaStream_t st);
void t_srgb2spec(const DevScene &s, const float *rgb, const float *wl, float *out, int n, cudaStream_t st);
void t_spectrum_sample(const DevScene &s, const uint32_t *rand, float *out, int n, cudaStream_t st);
void t_tex_fetch(const DevScene &s, int slot, const float *pos, float *out, int n, cudaStream_t st);
void t_cast_dir(const DevScene &s, const float *u, const float *v, const uint32_t *rand, float *out, int n, cudaStream_t st);
void t_opendrt(const float *rgb, float *out, int n, cudaStream_t st);
void t_agx(const float *rgb, float *out, int n, cudaStream_t st);
void t_crf(const DevScene &s, const float *rgb, float *out, int n, cudaStream_t st);
void t_srgb_oetf(const float *x, float *out, int n, cudaStream_t st);
void t_intersect_land(const DevScene &s, const float *pos, const float *dir, float *out, int n, cudaStream_t st);
void t_land_normal(const DevScene &s, const float *pos, float *out, int n, cudaStream_t st);
void t_land_material(const DevScene &s, const float *pos, float *out, int n, cudaStream_t st);
void t_cloud_limits(const float *pos, const float *dir, const float *land, float *out, int n, cudaStream_t st);
void t_clouds_density(const DevScene &s, const float *pos, float *out, int n, cudaStream_t st);
void t_raymarch_T(const float *pos, const float *dir, const float *ext, float *out, int n, cudaStream_t st);
void t_tracking(const DevScene &s, int kind, const float *pos, const float *dir, const float *land, const float *wl, uint32_t seed, float *out, int n, cudaStream_t st);
void t_ray_march(const DevScene &s, const float *pos, const float *dir, const float *t0, const float *t1, const float *sun, const float *wl, float *out2, int n, cudaStream_t st);
void t_trace_preview(const DevScene &s, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out, int n, cudaStream_t st);
void t_trace_paths(const DevScene &s, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out, int n, cudaStream_t st);
}  // namespace de_exact
