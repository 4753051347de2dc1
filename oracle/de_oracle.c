/* de_oracle.c -- CPU ORACLE for the Digital-Earth spectral path tracer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (digital-earth_b200/) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, as the checker / baseline.
 *
 * What it is: a plain-C, IEEE-binary32, source-order restatement of the
 * reference's Taichi device code (file:line cited at every function, relative
 * to /root/reference).  Build with -ffp-contract=off -fno-fast-math
 * (oracle/Makefile) so no FMA contraction or reassociation happens.
 *
 * Pinning: Taichi itself is absent from this environment, so the reference
 * cannot be executed through its real compiler ("parity unpinned" at the
 * Taichi-runtime boundary).  What IS pinned: tests/golden/ (.npz files) hold outputs
 * of the reference's own unmodified source files executed through the Taichi
 * stand-in in oracle/ti_shim (scalar f32, glibc libm); tests/test_oracle_golden.py
 * requires this file to reproduce them bit for bit, plus the known-answer
 * checks of SURVEY.md section 4.
 *
 * Semantics fixed where Taichi's are not visible in the reference (see
 * DESIGN.md): literals fold in f64 only when every operand is a Python-scope
 * constant; max/min = fmaxf/fminf; float->int casts truncate; bilinear fetch =
 * manual FP32 lerp a+f*(b-a), x then y, texel centres (i+.5)/N, clamp-to-edge;
 * the CIE LUT is rounded to fp16 before filtering (rgba16f texture,
 * renderer.py:97); xi = (u32>>8)*2^-24.
 *
 * RNG contract (shared by oracle, golden generator, and every CUDA integrator flavour):
 * per (pixel, sample, bounce) a stream of 32-bit SLOTS; slot i is word i&3 of
 * Philox4x32-10(key=(seed,pixel), counter=(sample,bounce,i>>2,0)); bounce 0 = wavelength +
 * pixel jitter, bounce k+1 = path segment k.  Each ti.random() of the reference takes the next
 * slot, with two rules that make the stream SIMT-friendly (one Philox block per two tracking
 * steps, no per-lane phase):
 *   (1) the slot index is rounded up to a multiple of 4 on entry to
 *       sample_interaction_delta_tracking, transmittance_ratio_tracking, the light-direction
 *       sample_cone_oriented, sample_phase and sample_hemisphere_cosine_weighted;
 *   (2) every loop trip of transmittance_ratio_tracking owns two slots (the second is unused),
 *       exactly like a delta-tracking trip (free flight + acceptance test).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- types -- */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { float x, y, z, w; } v4;

typedef struct {
    const uint8_t *data; /* row-major [y][x][c], y=0 is v=0 (south pole row) */
    int32_t w, h, c;
} orc_tex;

enum { T_ALBEDO = 0, T_TOPO, T_OCEAN, T_CLOUDS, T_BATHY, T_EMISSIVE, T_STARS, T_COUNT };

typedef struct {
    orc_tex tex[T_COUNT];
    const float *cie;          /* LUT/CIE.dat as stored: [row 2][x 441][3] f32 */
    const uint16_t *srgb2spec; /* LUT/srgb2spec.dat: [300][3] fp16 bits        */
    const float *o3;           /* LUT/ozone_cross_section.dat: [441] f32       */
    const float *crf;          /* [n_crf][1024][3] f32                         */
    int32_t n_crf;
    float cam_pos[3], look_at[3], up[3];
    float fov, aspect_scale, sun_angle, sun_path_rot, land_height_scale;
    float exposure, gamma;
    int32_t selected_crf, crf_count;
    float vig_strength, vig_radius, vig_cx, vig_cy;
    int32_t tonemapper; /* 0 OpenDRT, 1 AgX */
    int32_t topo_tex_w; /* TOPOGRAPHY_TEX_RES[0] (lib/textures.py) -> normal epsilon */
    int32_t W, H;
} orc_scene;

typedef struct {
    uint64_t paths, segments, rmo_steps, cloud_steps, sdf_evals, tex_fetches, surface_hits, rng_draws;
} orc_counters;

/* ------------------------------------------------------------------ RNG -- */
typedef struct {
    uint32_t key0, key1, sample, bounce, draw;
    uint32_t buf[4];
    const uint32_t *list; /* explicit draw list (unit tests) when non-NULL */
    uint32_t list_pos;
    int buf_valid;
    orc_counters *cnt;
} orc_rng;

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t rng_u32(orc_rng *r) {
    if (r->cnt) r->cnt->rng_draws++;
    if (r->list) return r->list[r->list_pos++];
    if ((r->draw & 3u) == 0 || !r->buf_valid) {
        uint32_t ctr[4] = { r->sample, r->bounce, r->draw >> 2, 0u }, key[2] = { r->key0, r->key1 };
        philox4x32_10(ctr, key, r->buf);
        r->buf_valid = 1;
    }
    return r->buf[(r->draw++) & 3u];
}
/* contract rule (1) / (2); no-ops for the explicit draw lists of the unit-test entry points */
static void rng_align(orc_rng *r) { if (!r->list) { r->draw = (r->draw + 3u) & ~3u; r->buf_valid = 0; } }
static void rng_skip(orc_rng *r) { if (!r->list) { r->draw += 1u; r->buf_valid = 0; } }
/* ti.random(f32) */
static float rnd(orc_rng *r) { return (float)(rng_u32(r) >> 8) * (1.0f / 16777216.0f); }
static void rng_bounce(orc_rng *r, uint32_t b) { r->bounce = b; r->draw = 0; r->buf_valid = 0; }

/* ------------------------------------------------------------ vector ops -- */
static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scl3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 neg3(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } /* (self*other).sum() */
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 norm3(v3 a) { float inv = 1.0f / len3(a); return scl3(a, inv); }    /* invlen * self */
static inline v3 cross3(v3 a, v3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float sqr(float x) { return x * x; }                        /* math_utils.py:9 */
static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float saturate(float x) { return clampf(x, 0.0f, 1.0f); }  /* math_utils.py:46 */
static inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline v3 mix3(v3 x, v3 y, float a) { return add3(scl3(x, 1.0f - a), scl3(y, a)); }
static inline float smoothstep(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
/* ops.pow: exponent exactly 2 is a multiply (LLVM InstCombine / NVVM / gcc all fold pow(x,2.0) -> x*x
 * without fast-math, so this is what Taichi's backends execute); everything else is libm powf. */
static inline float pow_ti(float x, float y) { return y == 2.0f ? x * x : powf(x, y); }
static inline float log2_ti(float x) { return logf(x) / 0.6931471805599453f; } /* taichi.math.log2 */

static const float PI_F = (float)3.141592653589793;
static const float TWO_PI_F = (float)(2.0 * 3.141592653589793);

/* lib/volume_rendering_models.py:8-44 */
#define PLANET_R 6371000.0f
#define ATMOS_UPPER 6481000.0f
#define CLOUDS_LOWER 6375000.0f
#define CLOUDS_UPPER 6381000.0f
#define CLOUDS_THICKNESS 6000.0f
#define CLOUDS_EXTINCT 0.1f
#define CLOUDS_DENSITY 0.029f
#define MIE_ASYMMETRY 3000.0f
#define RAYLEIGH_ALBEDO 1.0f /* volume_rendering_models.py:27-28 */
#define AEROSOL_ALBEDO 0.95f
enum { RAYLEIGH_ID = 0, MIE_ID = 1, OZONE_ID = 2, CLOUD_ID = 3, ISOTROPIC_CLOUD_ID = 4 };
enum { NULL_EVENT = 0, ABSORB_EVENT = 1, SCATTER_EVENT = 2 };

/* ------------------------------------------------------------- textures -- */
static float half_to_float(uint16_t h) {
    uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u, bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else { int sh = 0; while (!(m & 1024u)) { m <<= 1; ++sh; } m &= 1023u; bits = s | ((uint32_t)(113 - sh) << 23) | (m << 13); }
    } else if (e == 31) bits = s | 0x7F800000u | (m << 13);
    else bits = s | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4); return f;
}
static uint16_t float_to_half_rn(float f) { /* round-to-nearest-even, as numpy astype(float16) */
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000u; int32_t e = (int32_t)((x >> 23) & 255u) - 127 + 15; uint32_t m = x & 0x7FFFFFu;
    if (((x >> 23) & 255u) == 255u) return (uint16_t)(s | 0x7C00u | (m ? 0x200u : 0));
    if (e >= 31) return (uint16_t)(s | 0x7C00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m |= 0x800000u; int sh = 14 - e; uint32_t r = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        if (rem > half || (rem == half && (r & 1u))) ++r;
        return (uint16_t)(s | r);
    }
    uint32_t r = ((uint32_t)e << 10) | (m >> 13), rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    return (uint16_t)(s | r);
}

typedef struct { float c[4]; } texel4;
typedef float (*texel_fn)(const void *ctx, int x, int y, int c);

/* ti.Texture.sample_lod(uv, 0) as fixed by this oracle: see header. */
static texel4 bilinear(texel_fn fetch, const void *ctx, int w, int h, int nc, float u, float v) {
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 > w - 1 ? w - 1 : x0); x1 = x1 < 0 ? 0 : (x1 > w - 1 ? w - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
    texel4 out = { { 0, 0, 0, 0 } };
    for (int c = 0; c < nc; ++c) {
        float t00 = fetch(ctx, x0, y0, c), t10 = fetch(ctx, x1, y0, c);
        float t01 = fetch(ctx, x0, y1, c), t11 = fetch(ctx, x1, y1, c);
        float a = t00 + fx * (t10 - t00), b = t01 + fx * (t11 - t01);
        out.c[c] = a + fy * (b - a);
    }
    return out;
}
/* renderer.py:170-210: u8 -> /255.0 -> unorm8 texture */
static float fetch_u8(const void *ctx, int x, int y, int c) {
    const orc_tex *t = (const orc_tex *)ctx;
    return (float)t->data[((size_t)y * t->w + x) * t->c + c] / 255.0f;
}
/* renderer.py:97,212-216: f32 field -> rgba16f texture (fp16 quantisation) */
static float fetch_cie(const void *ctx, int x, int y, int c) {
    const float *cie = (const float *)ctx;
    return half_to_float(float_to_half_rn(cie[((size_t)y * 441 + x) * 3 + c]));
}
typedef struct { const float *crf; int n; } crf_ctx;
static float fetch_crf(const void *ctx, int x, int y, int c) {
    const crf_ctx *k = (const crf_ctx *)ctx;
    return k->crf[((size_t)y * 1024 + x) * 3 + c];
}

/* math_utils.py:17-23.  NOTE the select tests the sqrt, so a miss yields a NaN pair. */
static v2 rsi(v3 pos, v3 dir, float r) {
    float b = dot3(pos, dir);
    float discr = b * b - dot3(pos, pos) + r * r;
    discr = sqrtf(discr);
    v2 o;
    if (discr < 0.0f) { o.x = -1.0f; o.y = -1.0f; }
    else { o.x = -b + -discr; o.y = -b + discr; }
    return o;
}
/* math_utils.py:25-28 */
static v2 sphere_UV_map(v3 n) {
    v2 uv;
    uv.x = (atan2f(n.z, -n.x) / PI_F + 1.0f) / 2.0f;
    uv.y = asinf(n.y) / PI_F + 0.5f;
    return uv;
}
/* math_utils.py:38-44 */
static texel4 sample_sphere_texture(const orc_tex *t, v3 pos, orc_counters *cnt) {
    v2 uv = sphere_UV_map(norm3(pos));
    float u = uv.x * 1.0f, v = uv.y * 1.0f;
    u = u - floorf(u); v = v - floorf(v); /* fract */
    if (cnt) cnt->tex_fetches++;
    return bilinear(fetch_u8, t, t->w, t->h, t->c, u, v);
}

/* ------------------------------------------------- volume model (a10,a11) -- */
/* volume_rendering_models.py:229-246 */
static float get_ozone_density(float h) {
    float h_km = h * 0.001f;
    float d2 = h_km - (float)(25000.0 * 0.001);
    d2 = d2 * d2;
    float peak_density = 1.0f;
    float d = (peak_density - 0.375f) * expf(-d2 / 49.0f);
    d += 0.375f * expf(-d2 / 256.0f);
    d += fmaxf(0.0f, -0.000015f * pow_ti(h_km - 15.0f, 3.0f));
    return d;
}
/* volume_rendering_models.py:248-252 */
static float get_rayl_density(float h) {
    float density_sea_level = 1.225f;
    return 3.68082f * expf(-pow_ti(h + 24239.99f, 2.0f) / 532307548.4168f) / density_sea_level;
}
/* volume_rendering_models.py:254-267 */
static float get_mie_density(float h) {
    float dens = 0.0f;
    if (h > 11500.0f) dens = 0.0918f * expf(-1.0e-6f * pow_ti(h - 11500.0f, 2.0f));
    else if (h > 2400.0f) dens = 0.3000f * expf(-2.5e-9f * pow_ti(h + 2500.00f, 2.0f)) - 0.092f;
    else if (h > 1300.0f) dens = 0.6500f * expf(-5.0e-6f * pow_ti(h - 1300.00f, 2.0f)) + 0.18899f;
    else dens = 1.0f - h / 8136.646f;
    return dens * 1.06f; /* turbidity */
}
/* volume_rendering_models.py:270-273 */
static v3 get_density(float h) {
    h = fmaxf(h, 0.0f);
    return V3(get_rayl_density(h), get_mie_density(h), get_ozone_density(h));
}
/* volume_rendering_models.py:275-277 */
static float get_elevation(v3 p) { return sqrtf(p.x * p.x + p.y * p.y + p.z * p.z) - PLANET_R; }

/* volume_rendering_models.py:48-51 */
static float air(float wl) {
    float rcp = 1.0f / (wl * wl);
    return (float)(1.0 + 8.06051e-5) + 2.480990e-2f / (132.274f - rcp) + 1.74557e-4f / (39.32957f - rcp);
}
/* volume_rendering_models.py:194-200 */
static float spectra_extinction_mie(float wl) {
    float junge = 4.0f;
    float c = (float)((0.6544 * 1.06 - 0.6510) * 4e-18);
    float K = (0.773335f - 0.00386891f * wl) / (1.0f - 0.00546759f * wl);
    return 0.434f * c * PI_F * pow_ti(TWO_PI_F / (wl * 1e-9f), junge - 2.0f) * K;
}
/* volume_rendering_models.py:203-217 */
static float spectra_extinction_rayleigh(float wl) {
    float wn = wl * 1e-9f;
    float F_N2 = 1.034f + 3.17e-4f * (1.0f / pow_ti(wl, 2.0f));
    float F_O2 = 1.096f + 1.385e-3f * (1.0f / pow_ti(wl, 2.0f)) + 1.448e-4f * (1.0f / pow_ti(wl, 4.0f));
    float CCO2 = 0.0421f;
    float king = (78.084f * F_N2 + 20.946f * F_O2 + 0.934f + CCO2 * 1.15f) / ((float)(78.084 + 20.946 + 0.934) + CCO2);
    float n = sqr(air(wl * 1e-3f)) - 1.0f;
    return (((float)(8.0 * 31.006276680299816) * pow_ti(n, 2.0f)) / ((float)(3.0 * 2.5035422e25) * pow_ti(wn, 4.0f))) * king;
}
/* volume_rendering_models.py:219-224 */
static float spectra_extinction_ozone(float wl, const float *o3) {
    float ext = 0.0f;
    if (wl >= 390.0f && wl < 831.0f)
        ext = (float)(0.0001 * (2.5035422e25 * 0.012588 * 8e-6)) * o3[(int)(wl - 390.0f)];
    return ext;
}
/* colour.py:51-60 */
static float plancks(float T, float wl) {
    float h = 6.62607015e-16f, c = 2.9e17f, k = 1.38e-5f;
    float p1 = 2.0f * h * pow_ti(c, 2.0f) / pow_ti(wl, 5.0f);
    float p2 = expf((h * c) / (wl * k * T)) - 1.0f;
    return p1 / p2;
}
/* math_utils.py:13-15 */
static float cone_angle_to_solid_angle(float x) { return TWO_PI_F * (1.0f - cosf(x)); }

/* -------------------------------------------------------- phases (a12) -- */
/* volume_rendering_models.py:61-63 */
static float rayleigh_phase(float c) { return (float)(3.0 / (16.0 * 3.141592653589793)) * (1.0f + c * c); }
/* :87-89 */
static float klein_nishina_phase(float c, float e) {
    return e / (TWO_PI_F * (e * (1.0f - c) + 1.0f) * logf(2.0f * e + 1.0f));
}
/* :65-67 */
static float mie_phase(float c) { return klein_nishina_phase(c, MIE_ASYMMETRY); }
/* :73-75 */
static float hg_phase(float c, float g) {
    return (1 - g * g) / ((float)(4.0 * 3.141592653589793) * pow_ti(1.0f + g * g - 2 * g * c, 1.5f));
}
/* :121-122 */
/* draine_phase, cloud_params, sample_draine: "An Approximate Mie Scattering Function for Fog and Cloud Rendering" as included in
 * volume_rendering_models.py:99-183.  SPDX-FileCopyrightText: Copyright (c) <2023> NVIDIA CORPORATION & AFFILIATES. All rights
 * reserved.  SPDX-License-Identifier: MIT -- full notice in NOTICE.md. */
static float draine_phase(float c, float g, float a) {
    return ((1 - g * g) * (1 + a * c * c)) / (4.f * (1 + (a * (1 + 2 * g * g)) / 3.f) * PI_F * pow_ti(1 + g * g - 2 * g * c, 1.5f));
}
typedef struct { float g_hg, g_draine, alpha_draine, w_draine; } cloud_par;
/* :155-160 / :166-171 */
static cloud_par cloud_params(int reduce_peak) {
    float d = 8.0f;
    cloud_par p;
    p.g_hg = reduce_peak ? 0.91f : expf(-0.0990567f / (d - 1.67154f));
    p.g_draine = expf(-2.20679f / (d + 3.91029f) - 0.428934f);
    p.alpha_draine = expf(3.62489f - 8.29288f / (d + 5.52825f));
    p.w_draine = expf(-0.599085f / (d - 0.641583f) - 0.665888f);
    return p;
}
/* :154-162 */
static float cloud_phase(float c, int reduce_peak) {
    cloud_par p = cloud_params(reduce_peak);
    return mixf(hg_phase(c, p.g_hg), draine_phase(c, p.g_draine, p.alpha_draine), p.w_draine);
}
/* math_utils.py:55-60 */
static void make_orthonormal_basis(v3 n, v3 *x, v3 *y) {
    v3 h = fabsf(n.y) > 0.9f ? V3(1.0f, 0.0f, 0.0f) : V3(0.0f, 1.0f, 0.0f);
    *y = norm3(cross3(n, h));
    *x = cross3(n, *y);
}
/* math_utils.py:67-69 */
static v3 spherical_direction(float st, float ct, float phi, v3 x, v3 y, v3 z) {
    return add3(add3(scl3(x, st * cosf(phi)), scl3(y, st * sinf(phi))), scl3(z, ct));
}
/* volume_rendering_models.py:78-85 */
static v3 sample_hg_phase(v3 view, float g, orc_rng *r) {
    float sqr_term = (1 - g * g) / (1 - g + 2 * g * rnd(r));
    float ct = (1 + g * g - sqr_term * sqr_term) / (2 * g);
    float st = sqrtf(fmaxf(0.0f, 1 - ct * ct));
    float phi = TWO_PI_F * rnd(r);
    v3 t, b; make_orthonormal_basis(view, &t, &b);
    return spherical_direction(st, ct, phi, t, b, view);
}
/* :91-98 */
static v3 sample_klein_nishina_phase(v3 view, float e, orc_rng *r) {
    float ct = (-pow_ti(2.0f * e + 1.0f, 1.0f - rnd(r)) + e + 1.0f) / e;
    float st = sqrtf(fmaxf(0.0f, 1 - ct * ct));
    float phi = TWO_PI_F * rnd(r);
    v3 t, b; make_orthonormal_basis(view, &t, &b);
    return spherical_direction(st, ct, phi, t, b, view);
}
/* :125-150 (NVIDIA "approximate Mie" Draine sampler, MIT) -- operation order kept */
static v3 sample_draine(v3 view, float g, float a, orc_rng *r) {
    float xi = rnd(r);
    float g2 = g * g, g3 = g * g2, g4 = g2 * g2, g6 = g2 * g4;
    float pgp1_2 = (1 + g2) * (1 + g2);
    float T1 = (-1 + g2) * (4 * g2 + a * pgp1_2); (void)T1;
    float T1a = -a + a * g4;
    float T1a3 = T1a * T1a * T1a;
    float T2 = -1296 * (-1 + g2) * (a - a * g2) * (T1a) * (4 * g2 + a * pgp1_2);
    float T3 = 3 * g2 * (1 + g * (-1 + 2 * xi)) + a * (2 + g2 + g3 * (1 + 2 * g2) * (-1 + 2 * xi));
    float T4a = 432 * T1a3 + T2 + 432 * (a - a * g2) * T3 * T3;
    float T4b = -144 * a * g2 + 288 * a * g4 - 144 * a * g6;
    float T4b3 = T4b * T4b * T4b;
    float T4 = T4a + sqrtf(-4 * T4b3 + T4a * T4a);
    float T4p3 = pow_ti(T4, (float)(1.0 / 3.0));
    const float c48 = (float)(48 * 1.2599210498948732), c3 = (float)(3. * 1.2599210498948732);
    float T6 = (2 * T1a + (c48 * (-(a * g2) + 2 * a * g4 - a * g6)) / T4p3 + T4p3 / c3) / (a - a * g2);
    float T5 = 6 * (1 + g2) + T6;
    float ct = (1 + g2 - pow_ti(-0.5f * sqrtf(T5) + sqrtf(6 * (1 + g2) - (8 * T3) / (a * (-1 + g2) * sqrtf(T5)) - T6) / 2.f, 2.0f)) / (2.f * g);
    float st = sqrtf(fmaxf(0.0f, 1 - ct * ct));
    float phi = TWO_PI_F * rnd(r);
    v3 t, b; make_orthonormal_basis(view, &t, &b);
    return spherical_direction(st, ct, phi, t, b, view);
}
/* :164-183 */
static v3 sample_cloud_phase(v3 view, int reduce_peak, orc_rng *r) {
    cloud_par p = cloud_params(reduce_peak);
    if (rnd(r) < p.w_draine) return sample_draine(view, p.g_draine, p.alpha_draine, r);
    return sample_hg_phase(view, p.g_hg, r);
}
/* sampling.py:41-44 */
static v3 sample_sphere(float r0, float r1) {
    r0 *= TWO_PI_F; r1 = r1 * 2.0f - 1.0f;
    float s = sqrtf(1.0f - r1 * r1);
    return norm3(V3(sinf(r0) * s, cosf(r0) * s, r1));
}
/* sampling.py:13-23 */
static v3 sample_cone(float cmax, orc_rng *r) {
    float u0 = rnd(r), u1 = rnd(r);
    float ct = (1.0f - u0) + u0 * cmax;
    float st = sqrtf(1.0f - ct * ct);
    float phi = TWO_PI_F * u1;
    return V3(st * cosf(phi), st * sinf(phi), ct);
}
/* sampling.py:25-28 + math_utils.py:62-65 */
static v3 sample_cone_oriented(float cmax, v3 n, orc_rng *r) {
    v3 x, y; make_orthonormal_basis(n, &x, &y);
    v3 s = sample_cone(cmax, r);
    return V3(x.x * s.x + y.x * s.y + n.x * s.z, x.y * s.x + y.y * s.y + n.y * s.z, x.z * s.x + y.z * s.y + n.z * s.z);
}
/* sampling.py:30-39 */
static v3 sample_hemisphere_cosine_weighted(v3 n, orc_rng *r) {
    float u0 = rnd(r), u1 = rnd(r);
    float a = 1.0f - 2.0f * u0;
    float b = sqrtf(1.0f - a * a);
    a *= (float)(1.0 - 1e-5);
    b *= (float)(1.0 - 1e-5);
    float phi = TWO_PI_F * u1;
    return norm3(V3(n.x + b * cosf(phi), n.y + b * sinf(phi), n.z + a));
}
/* pathtracer.py:235-247 */
static float evaluate_phase(v3 ray_dir, v3 light_dir, int id, int reduce_peak) {
    float phase = 0.0f, c = dot3(ray_dir, light_dir);
    if (id == RAYLEIGH_ID) phase += rayleigh_phase(c);
    else if (id == MIE_ID) phase += klein_nishina_phase(c, MIE_ASYMMETRY);
    else if (id == CLOUD_ID) phase += cloud_phase(c, reduce_peak);
    else if (id == ISOTROPIC_CLOUD_ID) phase += (float)(1.0 / (4.0 * 3.141592653589793));
    return phase;
}
/* pathtracer.py:249-261 */
static v3 sample_phase(v3 ray_dir, int id, int reduce_peak, orc_rng *r, float *phase_div_pdf) {
    v3 d;
    *phase_div_pdf = 1.0f;
    if (id == RAYLEIGH_ID || id == ISOTROPIC_CLOUD_ID) {
        float r0 = rnd(r), r1 = rnd(r);
        d = sample_sphere(r0, r1);
        *phase_div_pdf = evaluate_phase(ray_dir, d, id, reduce_peak) * (float)(4.0 * 3.141592653589793);
    } else if (id == MIE_ID) d = sample_klein_nishina_phase(ray_dir, MIE_ASYMMETRY, r);
    else d = sample_cloud_phase(ray_dir, reduce_peak, r);
    return d;
}
/* pathtracer.py:263-270 */
static int sample_scatter_event(int id, orc_rng *r) {
    static const float albedos[4] = { 1.0f, 0.95f, 0.0f, 0.99f };
    if (id == ISOTROPIC_CLOUD_ID) id = CLOUD_ID;
    return rnd(r) < albedos[id];
}

/* ------------------------------------------------ surface model (a13,a14) -- */
/* surface_rendering_models.py:39-51 */
static float disney_diffuse(float rough, float nl, float nv, float lh) {
    float R_R = 2.0f * rough * sqr(lh);
    float F_L = pow_ti(1.0f - nl, 5.0f), F_V = pow_ti(1.0f - nv, 5.0f);
    float f_lambert = (float)(1.0 / 3.141592653589793);
    float f_retro = f_lambert * R_R * (F_L + F_V + F_L * F_V * (R_R - 1.0f));
    return f_lambert * (1.0f - 0.5f * F_L) * (1.0f - 0.5f * F_V) + f_retro;
}
/* :110-122 */
static float fresnel_dielectric(float vh, float F0) {
    F0 = sqrtf(F0);
    F0 = (1.0f + F0) / (1.0f - F0);
    float sI = sqrtf(saturate(1.0f - sqr(vh)));
    float sT = sI / fmaxf(F0, 1e-8f);
    float cT = sqrtf(1.0f - sqr(sT));
    float Rs = sqr((vh - (F0 * cT)) / fmaxf(vh + (F0 * cT), 1e-8f));
    float Rp = sqr((cT - (F0 * vh)) / fmaxf(cT + (F0 * vh), 1e-8f));
    return saturate((Rs + Rp) * 0.5f);
}
/* :82-85 */
static float GGX_D(float nh, float a2) {
    float den = (a2 - 1.0f) * nh * nh + 1.0f;
    return a2 / (PI_F * den * den);
}
/* :88-91 */
static float lambda_smith(float nx, float a2) {
    float x2 = nx * nx;
    return (-1.0f + sqrtf(a2 * (1.0f - x2) / x2 + 1.0f)) * 0.5f;
}
/* :100-104 */
static float G2_smith(float nl, float nv, float a2) {
    float lv = lambda_smith(nv, a2), ll = lambda_smith(nl, a2);
    return 1.0f / (1.0f + lv + ll);
}
/* :69-80 */
static float GGX_smith_specular(float rough, float F0, float nl, float nv, float lh, float nh) {
    float a2 = rough * rough;
    float D = GGX_D(nh, a2), G = G2_smith(nl, nv, a2), F = fresnel_dielectric(lh, F0);
    return D * G * F / fmaxf(4.0f * nl * nv, 1e-5f);
}
/* :146-152 */
static float beckmann_isotropic_ndf(float nh, float alpha) {
    float c2 = nh * nh, a2 = alpha * alpha;
    float exponent = (1.0f - c2) / (a2 * c2);
    float denom = PI_F * a2 * c2 * c2;
    return expf(-exponent) / fmaxf(denom, 1e-5f);
}
/* :169-171 */
static float G2_VCavity(float nl, float nv, float nh, float vh) {
    return fminf(1.0f, fminf(2.0f * nv * nh / vh, 2.0f * nl * nh / vh));
}
/* :53-67 */
static float beckmann_specular(float rough, float F0, float nl, float nv, float lh, float nh) {
    float alpha = rough;
    alpha *= alpha * 2.0f;
    float D = beckmann_isotropic_ndf(nh, alpha), V = G2_VCavity(nl, nv, nh, lh), F = fresnel_dielectric(lh, F0);
    return D * V * F;
}
/* :9-37 */
static float earth_brdf(float albedo, float oceanness, float bathymetry, v3 v, v3 n, v3 l, float *n_dot_l_out) {
    v3 h = norm3(add3(v, l));
    float nl = saturate(dot3(n, l)), nv = saturate(dot3(n, v));
    float lh = saturate(dot3(l, h)), nh = saturate(dot3(n, h));
    float land_roughness = 0.73f;
    float ocean_roughness = mixf((float)(0.23 + 0.02), (float)(0.23 - 0.04), smoothstep(0.3f, 0.7f, bathymetry));
    float land_F0 = 0.04f, ocean_F0 = 0.02f;
    float diffuse = disney_diffuse(land_roughness, nl, nv, lh);
    float land_spec = GGX_smith_specular(land_roughness, land_F0, nl, nv, lh, nh);
    float ocean_ggx = GGX_smith_specular(ocean_roughness, ocean_F0, nl, nv, lh, nh);
    float ocean_beck = 0.65f * beckmann_specular(ocean_roughness, ocean_F0, nl, nv, lh, nh);
    float ocean_spec = mixf(ocean_beck, ocean_ggx, clampf(smoothstep(0.2f, 0.95f, nv), 0.05f, 0.94f));
    float blender = smoothstep(0.6f, 1.0f, oceanness);
    float brdf = albedo * diffuse * 0.28f + mixf(land_spec, ocean_spec, blender) * 0.5f;
    *n_dot_l_out = nl;
    return brdf;
}
/* colour.py:88-95 */
static float lum(v3 x) { return dot3(x, V3(0.2126729f, 0.7151522f, 0.0721750f)); }
static v3 lum3(v3 x) { float y = lum(x); return V3(y, y, y); }
/* colour.py:62-71 (sign of f reproduced: f = w - (lambda-400) <= 0) */
static float srgb_to_spectrum(const uint16_t *lut, v3 rgb, float wl) {
    int w = (int)(wl - 400.0f);
    float f = (float)w - (wl - 400.0f);
    float power = 0.0f;
    if (w > 0 && w < 299) {
        v3 a = V3(half_to_float(lut[w * 3]), half_to_float(lut[w * 3 + 1]), half_to_float(lut[w * 3 + 2]));
        v3 b = V3(half_to_float(lut[w * 3 + 3]), half_to_float(lut[w * 3 + 4]), half_to_float(lut[w * 3 + 5]));
        power = dot3(rgb, mix3(a, b, f));
    }
    return power;
}
typedef struct { v3 albedo_srgb; float ocean, bathymetry, emissive; } land_material;
/* pathtracer.py:284-313 */
static land_material get_land_material(const orc_scene *s, v3 pos, orc_counters *cnt) {
    land_material m;
    m.ocean = sample_sphere_texture(&s->tex[T_OCEAN], pos, cnt).c[0];
    texel4 at = sample_sphere_texture(&s->tex[T_ALBEDO], pos, cnt);
    v3 tex = V3(at.c[0], at.c[1], at.c[2]);
    v3 land = mix3(lum3(tex), tex, 6.5f);
    float greenery = pow_ti(land.y / lum(land), 2.0f);
    greenery = smoothstep(1.5f, 1.9f, greenery);
    float den = greenery * 0.7f + 1.0f;
    land = V3(1.0f * tex.x / den, 1.0f * tex.y / den, 1.0f * tex.z / den);
    land = mix3(lum3(land), land, 1.4f - greenery * 0.45f);
    v3 tinted = mul3(land, V3(255.0f, 128.0f, 64.0f));
    tinted = V3(tinted.x / 255.0f, tinted.y / 255.0f, tinted.z / 255.0f);
    land = mix3(land, tinted, 0.2f * (1.0f - greenery));
    v3 ocean_albedo = scl3(mix3(lum3(tex), tex, 0.75f), 0.9f);
    m.albedo_srgb = mix3(land, ocean_albedo, m.ocean);
    m.bathymetry = sample_sphere_texture(&s->tex[T_BATHY], pos, cnt).c[0];
    m.emissive = sample_sphere_texture(&s->tex[T_EMISSIVE], pos, cnt).c[0];
    return m;
}

/* ------------------------------------------------------ geometry (a4,a7) -- */
/* pathtracer.py:11-14 */
static float land_sdf(const orc_scene *s, v3 pos, float scale, orc_counters *cnt) {
    if (cnt) cnt->sdf_evals++;
    return len3(pos) - PLANET_R - scale * sample_sphere_texture(&s->tex[T_TOPO], pos, cnt).c[0];
}
/* pathtracer.py:16-25 */
static v3 land_normal(const orc_scene *s, v3 pos, float scale, orc_counters *cnt) {
    float d = land_sdf(s, pos, scale, cnt);
    float e = (float)(3.141592653589793 * 6371e3 / (double)s->topo_tex_w);
    float z = 0.0f;
    v3 n = V3(d - land_sdf(s, V3(pos.x - e, pos.y - z, pos.z - z), scale, cnt),
              d - land_sdf(s, V3(pos.x - z, pos.y - e, pos.z - z), scale, cnt),
              d - land_sdf(s, V3(pos.x - z, pos.y - z, pos.z - e), scale, cnt));
    return norm3(n);
}
/* pathtracer.py:27-46 */
static float intersect_land(const orc_scene *s, v3 pos, v3 dir, float height_scale, orc_counters *cnt) {
    float ray_dist = 0.0f;
    float max_ray_dist = (float)(6371e3 * 10.0);
    v2 rd = rsi(pos, dir, ATMOS_UPPER);
    if (rd.x > 0.0f) ray_dist = rd.x;
    for (int i = 0; i < 250; ++i) {
        v3 ro = add3(pos, scl3(dir, ray_dist));
        float dist = land_sdf(s, ro, height_scale, cnt);
        ray_dist += dist;
        if (ray_dist > max_ray_dist || fabsf(dist) < ray_dist * 0.0001f) break;
    }
    return ray_dist < max_ray_dist ? ray_dist : -1.0f;
}
/* pathtracer.py:48-65 */
static float get_clouds_density(const orc_scene *s, v3 pos, orc_counters *cnt) {
    float r = len3(pos), density = 0.0f;
    if (r > CLOUDS_LOWER && r < CLOUDS_UPPER) {
        float h = (r - CLOUDS_LOWER) / CLOUDS_THICKNESS;
        float cloud_texture = sample_sphere_texture(&s->tex[T_CLOUDS], pos, cnt).c[0];
        float column_height = cloud_texture;
        float split = 0.2f;
        density = (h - split < column_height * (1.0f - split) && split - h < column_height * split) ? fmaxf(cloud_texture, 0.4f) : 0.0f;
    }
    return density * CLOUDS_DENSITY;
}
/* pathtracer.py:67-71 */
static v4 get_atmos_density(const orc_scene *s, v3 pos, orc_counters *cnt) {
    v3 rmo = get_density(get_elevation(pos));
    v4 d = { rmo.x, rmo.y, rmo.z, get_clouds_density(s, pos, cnt) };
    return d;
}
/* pathtracer.py:145-169 */
static void intersect_cloud_limits(v3 pos, v3 dir, float land_isection, float *t_start, float *t_max) {
    float ts = 0.0f, tm = 0.0f, elevation = len3(pos);
    v2 lo = rsi(pos, dir, CLOUDS_LOWER), up = rsi(pos, dir, CLOUDS_UPPER);
    if (elevation >= CLOUDS_UPPER) {
        ts = fmaxf(0.0f, up.x);
        tm = lo.y >= 0.0f ? lo.x : up.y;
        if (up.y < 0.0f) tm = -1.0f;
    } else if (elevation >= CLOUDS_LOWER) {
        ts = 0.0f;
        tm = lo.y >= 0.0f ? lo.x : up.y;
    } else {
        ts = lo.y;
        tm = up.y;
        if (land_isection > 0.0f) tm = -1.0f;
    }
    *t_start = ts; *t_max = tm;
}

/* ------------------------------------------------- tracking (a5,a6,a8) -- */
/* pathtracer.py:77-115 */
static int delta_tracking(const orc_scene *s, v3 pos, v3 dir, float t_start, float t_max, v4 ext, float max_ext,
                          orc_rng *r, orc_counters *cnt, int is_cloud, float *t_out, int *id_out) {
    float t = t_start;
    pos = add3(pos, scl3(dir, t));
    int id = 0, event = NULL_EVENT;
    rng_align(r);
    while (t < t_max) {
        float t_step = -logf(rnd(r)) / max_ext;
        pos = add3(pos, scl3(dir, t_step));
        t += t_step;
        if (t >= t_max) break;
        if (cnt) { if (is_cloud) cnt->cloud_steps++; else cnt->rmo_steps++; }
        v4 d = get_atmos_density(s, pos, cnt);
        float es[4] = { ext.x * d.x, ext.y * d.y, ext.z * d.z, ext.w * d.w };
        float sum = ((es[0] + es[1]) + es[2]) + es[3];
        float rand = rnd(r);
        if (rand < sum / max_ext) {
            float cmf = 0.0f;
            while (id < 3) {
                cmf += es[id];
                if (rand < cmf / max_ext) break;
                id += 1;
            }
            event = sample_scatter_event(id, r) ? SCATTER_EVENT : ABSORB_EVENT;
            break;
        }
    }
    *t_out = t; *id_out = id;
    return event;
}
/* pathtracer.py:117-143 */
static float ratio_tracking(const orc_scene *s, v3 pos, v3 dir, float t_start, float t_max, v4 ext, float max_ext,
                            orc_rng *r, orc_counters *cnt, int is_cloud) {
    float t = t_start;
    pos = add3(pos, scl3(dir, t));
    float T = 1.0f;
    rng_align(r);
    while (t < t_max) {
        float t_step = -logf(rnd(r)) / max_ext;
        rng_skip(r);
        pos = add3(pos, scl3(dir, t_step));
        t += t_step;
        if (t >= t_max) break;
        if (cnt) { if (is_cloud) cnt->cloud_steps++; else cnt->rmo_steps++; }
        v4 d = get_atmos_density(s, pos, cnt);
        float sum = ((ext.x * d.x + ext.y * d.y) + ext.z * d.z) + ext.w * d.w;
        T *= 1.0f - sum / max_ext;
        if (T < 1e-5f) break;
    }
    return T;
}
/* pathtracer.py:172-207 */
static int sample_interaction(const orc_scene *s, v3 pos, v3 dir, float land_isection, v4 ext, float max_rmo, float max_cloud,
                              orc_rng *r, orc_counters *cnt, float *t_out, int *id_out) {
    v2 atm = rsi(pos, dir, ATMOS_UPPER);
    float t_start = fmaxf(0.0f, atm.x);
    float t_max = land_isection >= 0.0f ? land_isection : atm.y;
    if (atm.y < 0.0f) t_max = -1.0f;
    v4 rmo_ext = { ext.x, ext.y, ext.z, 0.0f };
    float rmo_t; int rmo_id;
    int rmo_event = delta_tracking(s, pos, dir, t_start, t_max, rmo_ext, max_rmo, r, cnt, 0, &rmo_t, &rmo_id);
    intersect_cloud_limits(pos, dir, land_isection, &t_start, &t_max);
    int event = rmo_event, id = rmo_id;
    float t = rmo_t;
    if (rmo_event == NULL_EVENT || rmo_t > t_start) {
        v4 cl_ext = { 0.0f, 0.0f, 0.0f, ext.w };
        float cloud_t; int cloud_id;
        int cloud_event = delta_tracking(s, pos, dir, t_start, t_max, cl_ext, max_cloud, r, cnt, 1, &cloud_t, &cloud_id);
        if (cloud_event > 0 && (cloud_t < rmo_t || rmo_event == NULL_EVENT)) {
            t = cloud_t; id = CLOUD_ID; event = cloud_event;
        }
    }
    *t_out = t; *id_out = id;
    return event;
}
/* pathtracer.py:211-232 */
static float sample_transmittance(const orc_scene *s, v3 pos, v3 dir, float land_isection, v4 ext, float max_rmo, float max_cloud,
                                  orc_rng *r, orc_counters *cnt) {
    v2 atm = rsi(pos, dir, ATMOS_UPPER);
    float t_start = fmaxf(0.0f, atm.x);
    float t_max = land_isection >= 0.0f ? land_isection : atm.y;
    if (atm.y < 0.0f) t_max = -1.0f;
    v4 rmo_ext = { ext.x, ext.y, ext.z, 0.0f };
    float T = ratio_tracking(s, pos, dir, t_start, t_max, rmo_ext, max_rmo, r, cnt, 0);
    intersect_cloud_limits(pos, dir, land_isection, &t_start, &t_max);
    v4 cl_ext = { 0.0f, 0.0f, 0.0f, ext.w };
    T *= ratio_tracking(s, pos, dir, t_start, t_max, cl_ext, max_cloud, r, cnt, 1);
    return T;
}

/* --------------------------------------------------- scene params (a18) -- */
typedef struct { v3 light_direction; float sun_cos_angle, sun_angular_radius, land_height_scale; } scene_params;
/* renderer.py:293-302 */
static scene_params make_scene_params(const orc_scene *s) {
    scene_params p;
    p.land_height_scale = s->land_height_scale;
    float sun_radius = 6.95e8f, sun_distance = 1.4959e11f;
    p.sun_angular_radius = sun_radius / sun_distance;
    p.sun_cos_angle = cosf(p.sun_angular_radius);
    float rx = -sinf(s->sun_path_rot), ry = cosf(s->sun_path_rot);
    p.light_direction = V3(-sinf(s->sun_angle), cosf(s->sun_angle) * rx, cosf(s->sun_angle) * ry);
    return p;
}

/* --------------------------------------------------- the integrator (a1) -- */
/* pathtracer.py:316-469 */
static float path_tracer(const orc_scene *s, const scene_params *sc, float wavelength, v3 ray_pos, v3 ray_dir,
                         orc_rng *r, orc_counters *cnt) {
    const v3 path_ray_dir = ray_dir;
    float sun_power = plancks(5778.0f, wavelength);
    float nightlights_power = plancks(2700.0f, wavelength) * 0.0001f;
    float sun_irradiance = sun_power * cone_angle_to_solid_angle(sc->sun_angular_radius);
    v3 d0 = get_density(0.0f);
    v3 max_rmo_d = V3(d0.x, d0.y, get_ozone_density(25000.0f));
    float max_density_cloud = CLOUDS_DENSITY;
    v4 ext;
    ext.x = spectra_extinction_rayleigh(wavelength);
    ext.y = spectra_extinction_mie(wavelength);
    ext.z = spectra_extinction_ozone(wavelength, s->o3);
    ext.w = CLOUDS_EXTINCT;
    int primary_miss = 0;
    float in_scattering = 0.0f, throughput = 1.0f;
    for (int scatter_count = 0; scatter_count < 25; ++scatter_count) {
        rng_bounce(r, (uint32_t)scatter_count + 1u);
        if (cnt) cnt->segments++;
        if (scatter_count > 9) ext.w = 0.02f;
        float max_ext_rmo = (ext.x * max_rmo_d.x + ext.y * max_rmo_d.y) + ext.z * max_rmo_d.z;
        float max_ext_cloud = ext.w * max_density_cloud;
        float earth_isect = intersect_land(s, ray_pos, ray_dir, sc->land_height_scale, cnt);
        float interaction_dist; int id;
        int event = sample_interaction(s, ray_pos, ray_dir, earth_isect, ext, max_ext_rmo, max_ext_cloud, r, cnt, &interaction_dist, &id);
        if (scatter_count > 9 && id == CLOUD_ID) id = ISOTROPIC_CLOUD_ID;
        rng_align(r);
        v3 light_dir = sample_cone_oriented(sc->sun_cos_angle, sc->light_direction, r);
        if (event == ABSORB_EVENT) break;
        else if (event == SCATTER_EVENT) {
            v3 ipos = add3(ray_pos, scl3(ray_dir, interaction_dist));
            int direct_visibility = rsi(ipos, light_dir, PLANET_R).y > 0.0f;
            float direct_T = 0.0f;
            if (!direct_visibility)
                direct_T = sample_transmittance(s, ipos, light_dir, -1.0f, ext, max_ext_rmo, max_ext_cloud, r, cnt);
            float direct_phase = evaluate_phase(ray_dir, light_dir, id, scatter_count > 0);
            in_scattering += throughput * direct_T * sun_irradiance * direct_phase;
            float pdp;
            rng_align(r);
            v3 sd = sample_phase(ray_dir, id, scatter_count > 0, r, &pdp);
            ray_dir = sd; ray_pos = ipos; throughput *= pdp;
        } else if (earth_isect > 0.0f) {
            if (cnt) cnt->surface_hits++;
            v3 land_pos = add3(ray_pos, scl3(ray_dir, earth_isect));
            v3 nrm = land_normal(s, land_pos, sc->land_height_scale, cnt);
            land_material m = get_land_material(s, land_pos, cnt);
            float albedo = srgb_to_spectrum(s->srgb2spec, m.albedo_srgb, wavelength);
            in_scattering += throughput * m.emissive * nightlights_power;
            v3 offset_pos = scl3(land_pos, 1.0f + 0.0001f * sc->land_height_scale / 12000.0f);
            int vis = intersect_land(s, offset_pos, light_dir, sc->land_height_scale, cnt) < 0.0f;
            float direct_T = sample_transmittance(s, offset_pos, light_dir, vis ? -1.0f : 0.0f, ext, max_ext_rmo, max_ext_cloud, r, cnt);
            float ndl;
            float dbrdf = earth_brdf(albedo, m.ocean, m.bathymetry, neg3(ray_dir), nrm, light_dir, &ndl);
            in_scattering += throughput * direct_T * (float)vis * sun_irradiance * dbrdf * ndl;
            v3 view_dir = neg3(ray_dir);
            rng_align(r);
            ray_dir = sample_hemisphere_cosine_weighted(nrm, r);
            ray_pos = offset_pos;
            float unused;
            float brdf = earth_brdf(albedo, m.ocean, m.bathymetry, view_dir, nrm, ray_dir, &unused);
            throughput *= brdf * PI_F;
        } else {
            if (scatter_count == 0) primary_miss = 1;
            break;
        }
        if (scatter_count > 3) {
            float p = fmaxf(0.05f, 1.0f - throughput);
            if (rnd(r) < p) break;
            throughput /= 1.0f - p;
        }
    }
    if (primary_miss) {
        if (dot3(sc->light_direction, path_ray_dir) > sc->sun_cos_angle) in_scattering += sun_power;
        texel4 st = sample_sphere_texture(&s->tex[T_STARS], path_ray_dir, cnt);
        float stars_power = srgb_to_spectrum(s->srgb2spec, V3(st.c[0], st.c[1], st.c[2]), wavelength);
        in_scattering += stars_power * sun_power * 0.0000001f;
    }
    if (isinf(in_scattering) || isnan(in_scattering) || in_scattering < 0.0f) in_scattering = 0.0f;
    return in_scattering;
}

/* ------------------------------------------ deterministic preview integrator -- */
/* pathtracer.py:471-499: 16-step optical depth towards the light; 0 when the planet is in the way
 * (also a deterministic test vehicle on its own: orc_raymarch_T) */
static float ray_march_transmittance(v3 ray_pos, v3 ray_dir, v3 rmo_ext) {
    const int steps = 16;
    float r_steps = 1.0f / (float)steps;
    float transmittance = 0.0f;
    int visibility = rsi(ray_pos, ray_dir, PLANET_R).y > 0.0f;
    if (!visibility) {
        v2 atm = rsi(ray_pos, ray_dir, ATMOS_UPPER);
        float t_max = atm.y;
        if (atm.y < 0.0f) t_max = -1.0f;
        float dd = t_max * r_steps;
        v3 ray_step = scl3(ray_dir, dd);
        v3 od = V3(0.0f, 0.0f, 0.0f);
        for (int i = 0; i < steps; ++i) {
            v3 density = get_density(get_elevation(ray_pos));
            od = add3(od, scl3(density, dd));
            ray_pos = add3(ray_pos, ray_step);
        }
        transmittance = expf(-dot3(rmo_ext, od));
    }
    return transmittance;
}
/* pathtracer.py:501-541: 64-step single scattering of Rayleigh + Mie along the view segment */
static void ray_march_atmos(v3 ray_pos, v3 ray_dir, float t_start, float t_max, v3 sun_dir, v3 rmo_ext, v2 rm_scat,
                            float *in_scatter_out, float *transmittance_out) {
    const int steps = 64;
    float r_steps = 1.0f / (float)steps;
    float dd = (t_max - t_start) * r_steps;
    v3 ray_step = scl3(ray_dir, dd);
    ray_pos = add3(ray_pos, scl3(ray_dir, t_start));
    float cos_theta = dot3(ray_dir, sun_dir);
    float ph_r = rayleigh_phase(cos_theta), ph_m = mie_phase(cos_theta);
    float transmittance = 1.0f, in_scatter = 0.0f;
    for (int i = 0; i < steps; ++i) {
        float h = get_elevation(ray_pos);
        v3 density = get_density(h);
        float step_od = dot3(rmo_ext, scl3(density, dd));
        float step_T = saturate(expf(-step_od));
        float step_integral = saturate((1.0f - step_T) / step_od);
        float visible = transmittance * step_integral;
        float sun_T = ray_march_transmittance(ray_pos, sun_dir, rmo_ext);
        float step_scat = rm_scat.x * (density.x * ph_r) + rm_scat.y * (density.y * ph_m); /* vec2.dot */
        in_scatter += step_scat * sun_T * visible * dd;
        transmittance *= step_T;
        ray_pos = add3(ray_pos, ray_step);
    }
    *in_scatter_out = in_scatter; *transmittance_out = transmittance;
}
/* pathtracer.py:543-685 (dead code upstream; kept as a noise-free preview).  RNG: the light-cone sample and the
 * hemisphere sample each start on a multiple of 4 of ONE stream (bounce key 1) for the whole path. */
static float ray_marcher(const orc_scene *s, const scene_params *sc, float wavelength, v3 ray_pos, v3 ray_dir, orc_rng *r, orc_counters *cnt) {
    const v3 path_ray_dir = ray_dir;
    float sun_power = plancks(5778.0f, wavelength);
    float nightlights_power = plancks(2700.0f, wavelength) * 0.0001f;
    float sun_irradiance = sun_power * cone_angle_to_solid_angle(sc->sun_angular_radius);
    v3 ext = V3(spectra_extinction_rayleigh(wavelength), spectra_extinction_mie(wavelength), spectra_extinction_ozone(wavelength, s->o3));
    v2 scat; scat.x = ext.x * RAYLEIGH_ALBEDO; scat.y = ext.y * AEROSOL_ALBEDO;
    int primary_miss = 0;
    float accum = 0.0f, throughput = 1.0f;
    rng_bounce(r, 1u);
    for (int scatter_count = 0; scatter_count < 3; ++scatter_count) {
        float earth_isect = intersect_land(s, ray_pos, ray_dir, sc->land_height_scale, cnt);
        v2 atm = rsi(ray_pos, ray_dir, ATMOS_UPPER);
        float t_start = fmaxf(0.0f, atm.x);
        float t_max = earth_isect > 0.0f ? earth_isect : atm.y;
        if (atm.y < 0.0f) { primary_miss = scatter_count == 0; break; }
        rng_align(r);
        v3 light_dir = sample_cone_oriented(sc->sun_cos_angle, sc->light_direction, r);
        float in_scatter, transmittance;
        ray_march_atmos(ray_pos, ray_dir, t_start, t_max, light_dir, ext, scat, &in_scatter, &transmittance);
        accum += throughput * in_scatter;
        throughput *= transmittance;
        if (earth_isect > 0.0f) {
            v3 land_pos = add3(ray_pos, scl3(ray_dir, earth_isect));
            v3 nrm = land_normal(s, land_pos, sc->land_height_scale, cnt);
            land_material m = get_land_material(s, land_pos, cnt);
            float albedo = srgb_to_spectrum(s->srgb2spec, m.albedo_srgb, wavelength);
            accum += throughput * m.emissive * nightlights_power;
            v3 offset_pos = scl3(land_pos, 1.0f + 0.0001f * sc->land_height_scale / 12000.0f);
            int vis = intersect_land(s, offset_pos, light_dir, sc->land_height_scale, cnt) < 0.0f;
            float ndl;
            float dbrdf = earth_brdf(albedo, m.ocean, m.bathymetry, neg3(ray_dir), nrm, light_dir, &ndl);
            accum += throughput * 1.0f * (float)vis * sun_irradiance * dbrdf * ndl;
            v3 view_dir = neg3(ray_dir);
            rng_align(r);
            ray_dir = sample_hemisphere_cosine_weighted(nrm, r);
            ray_pos = offset_pos;
            float unused;
            float brdf = earth_brdf(albedo, m.ocean, m.bathymetry, view_dir, nrm, ray_dir, &unused);
            throughput *= brdf * PI_F;
        }
    }
    if (primary_miss) {
        if (dot3(sc->light_direction, path_ray_dir) > sc->sun_cos_angle) accum += sun_power;
        texel4 st = sample_sphere_texture(&s->tex[T_STARS], path_ray_dir, cnt);
        float stars_power = srgb_to_spectrum(s->srgb2spec, V3(st.c[0], st.c[1], st.c[2]), wavelength);
        accum += stars_power * sun_power * 0.0000001f;
    }
    if (isinf(accum) || isnan(accum) || accum < 0.0f) accum = 0.0f;
    return accum;
}

/* colour.py:12-48 */
static void spectrum_sample(const float *cie, float sample, float *wavelength, v3 *response, float *rcp_pdf) {
    float lo = 0.0f, hi = 1.0f, mid = (lo + hi) / 2.0f;
    const float third = (float)(1.0 / 3.0);
    for (int x = 0; x < 8; ++x) { /* range(0, log2(441)) -> int(8.78) */
        texel4 t = bilinear(fetch_cie, cie, 441, 2, 3, mid, 0.25f);
        float val = saturate((third * t.c[0] + third * t.c[1]) + third * t.c[2]);
        if (val < sample) lo = mid;
        else if (val > sample) hi = mid;
        else break;
        mid = (lo + hi) / 2.0f;
    }
    *wavelength = 390.0f + 441.0f * mid;
    texel4 rs = bilinear(fetch_cie, cie, 441, 2, 3, mid, 0.75f);
    texel4 mx = bilinear(fetch_cie, cie, 441, 2, 3, 1.0f, 0.25f);
    *response = V3(rs.c[0], rs.c[1], rs.c[2]);
    float pdf = dot3(*response, V3(mx.c[0], mx.c[1], mx.c[2]));
    *rcp_pdf = 0.0f;
    if (pdf > 1e-3f && !(isinf(pdf) || isnan(pdf))) *rcp_pdf = 1.0f / pdf;
}
/* renderer.py:269-279 (renderer.py:230: up is normalised when set) */
static v3 get_cast_dir(const orc_scene *s, float u, float v, float xi_u, float xi_v) {
    float fov = s->fov;
    v3 cam = V3(s->cam_pos[0], s->cam_pos[1], s->cam_pos[2]);
    v3 look = V3(s->look_at[0], s->look_at[1], s->look_at[2]);
    v3 up = norm3(V3(s->up[0], s->up[1], s->up[2]));
    v3 d = norm3(sub3(look, cam));
    float aspect_ratio = (float)((double)s->W / (double)s->H);
    float fu = (2 * fov * (u + xi_u) / (float)s->H - fov * aspect_ratio - 1e-5f) * s->aspect_scale;
    float fv = 2 * fov * (v + xi_v) / (float)s->H - fov - 1e-5f;
    v3 du = norm3(cross3(d, up));
    v3 dv = norm3(cross3(du, d));
    return norm3(add3(add3(d, scl3(du, fu)), scl3(dv, fv)));
}
/* colour.py:6-10 */
static const float XYZ2RGB[9] = { (float)3.2409699419, (float)-1.5373831776, (float)-0.4986107603,
                                  (float)-0.9692436363, (float)1.8759675015, (float)0.0415550574,
                                  (float)0.0556300797, (float)-0.2039769589, (float)1.0569715142 };
/* renderer.py:305-330: one path sample for pixel (u,v) -> linear sRGB contribution */
static v3 render_sample2(const orc_scene *s, const scene_params *sc, int u, int v, uint32_t sample_index, uint32_t seed,
                         orc_counters *cnt, float *wl_out, float *L_out, int integrator);
static v3 render_sample(const orc_scene *s, const scene_params *sc, int u, int v, uint32_t sample_index, uint32_t seed,
                        orc_counters *cnt, float *wl_out, float *L_out) {
    return render_sample2(s, sc, u, v, sample_index, seed, cnt, wl_out, L_out, 0);
}
/* integrator: 0 = path_tracer (what renderer.py:317 calls), 1 = ray_marcher (the preview) */
static v3 render_sample2(const orc_scene *s, const scene_params *sc, int u, int v, uint32_t sample_index, uint32_t seed,
                         orc_counters *cnt, float *wl_out, float *L_out, int integrator) {
    orc_rng r; memset(&r, 0, sizeof r);
    r.key0 = seed; r.key1 = (uint32_t)(v * s->W + u); r.sample = sample_index; r.cnt = cnt;
    rng_bounce(&r, 0);
    float wl, rcp; v3 resp;
    spectrum_sample(s->cie, rnd(&r), &wl, &resp, &rcp);
    float xu = rnd(&r), xv = rnd(&r);
    v3 dir = get_cast_dir(s, (float)u, (float)v, xu, xv);
    v3 pos = V3(s->cam_pos[0], s->cam_pos[1], s->cam_pos[2]);
    float L = integrator == 1 ? ray_marcher(s, sc, wl, pos, dir, &r, cnt) : path_tracer(s, sc, wl, pos, dir, &r, cnt);
    if (cnt) cnt->paths++;
    v3 xyz = scl3(scl3(resp, L), rcp);
    if (wl_out) *wl_out = wl;
    if (L_out) *L_out = L;
    return V3((XYZ2RGB[0] * xyz.x + XYZ2RGB[1] * xyz.y) + XYZ2RGB[2] * xyz.z,
              (XYZ2RGB[3] * xyz.x + XYZ2RGB[4] * xyz.y) + XYZ2RGB[5] * xyz.z,
              (XYZ2RGB[6] * xyz.x + XYZ2RGB[7] * xyz.y) + XYZ2RGB[8] * xyz.z);
}

/* ------------------------------------------------------- tonemap (a17) -- */
/* The OpenDRT functions below follow OpenDRT v0.2.2 by Jed Smith (https://github.com/jedypod/open-display-transform) as ported in
 * lib/OpenDRT.py -- License: GPL v3 (NOTICE.md). */
/* OpenDRT.py:92-97 */
static float sdivf(float a, float b) { return fabsf(b) < 1e-4f ? 0.0f : a / b; }
/* OpenDRT.py:111-116 */
static float spow_ti(float a, float b) { return a <= 0.0f ? a : pow_ti(a, b); }
/* OpenDRT.py:78-83 */
static float _logf10(float x) { return log2_ti(x) / log2_ti(10.0f); }
/* OpenDRT.py:200-208 (forward only) */
static float tonescale_fwd(float x, float m, float s, float c) { return spow_ti(m * x / (x + s), c); }
/* OpenDRT.py:211-218 */
static float flare_fwd(float x, float fl) { return spow_ti(x, 2.0f) / (x + fl); }
static float flare_inv(float x, float fl) { return (x + sqrtf(x * (4.0f * fl + x))) / 2.0f; }
static v3 vdot(const float m[9], v3 v) { /* OpenDRT.py:86-88: v @ m (row vector) */
    return V3((v.x * m[0] + v.y * m[3]) + v.z * m[6], (v.x * m[1] + v.y * m[4]) + v.z * m[7], (v.x * m[2] + v.y * m[5]) + v.z * m[8]);
}
static v3 narrow_hue_angles(v3 v) { /* OpenDRT.py:191-197 */
    return V3(fminf(2.0f, fmaxf(0.0f, v.x - (v.y + v.z))), fminf(2.0f, fmaxf(0.0f, v.y - (v.x + v.z))), fminf(2.0f, fmaxf(0.0f, v.z - (v.x + v.y))));
}
/* OpenDRT.py:221-484 with in_gamut=rec709, display=Rec709, EOTF=lin (OpenDRT.py:39-55) */
static v3 openDR_transform(float pR, float pG, float pB) {
    static const float rec709_to_xyz[9] = { (float)0.412390917540, (float)0.357584357262, (float)0.180480793118,
                                            (float)0.212639078498, (float)0.715168714523, (float)0.072192311287,
                                            (float)0.019330825657, (float)0.119194783270, (float)0.950532138348 };
    static const float xyz_to_rec709[9] = { (float)3.2409699419, (float)-1.53738317757, (float)-0.498610760293,
                                            (float)-0.969243636281, (float)1.87596750151, (float)0.041555057407,
                                            (float)0.055630079697, (float)-0.203976958889, (float)1.05697151424 };
    const float Lp = 100.0f, gb = 0.12f, c = 1.0f, fl = 0.005f, rw = 0.25f, bw = 0.35f, dch = 0.35f, dch_toe = 0.0f;
    const float hs_r = 0.3f, hs_g = -0.1f, hs_b = -0.2f, v_p = 0.5f;
    float ds = (float)(100.0 / 100.0);
    float clamp_max = ds * Lp / 100.0f;
    float px = 128.0f * _logf10(Lp) / _logf10(100.0f) - 64.0f;
    float py = (float)(100.0 / 100.0);
    float gx = 0.18f;
    float gy = (float)(11.696 / 100.0) * (1.0f + gb * _logf10(py) / _logf10(2.0f));
    float s0 = flare_inv(gy, fl), m0 = flare_inv(py, fl);
    float ip = (float)(1.0 / 1.0);
    float s = (px * gx * (pow_ti(m0, ip) - pow_ti(s0, ip))) / (px * pow_ti(s0, ip) - gx * pow_ti(m0, ip));
    float m = pow_ti(m0, ip) * (s + px) / px;

    v3 rgb = V3(pR, pG, pB);
    rgb = vdot(rec709_to_xyz, rgb);
    rgb = vdot(xyz_to_rec709, rgb);
    float mx = fmaxf(rgb.x, fmaxf(rgb.y, rgb.z)), mn = fminf(rgb.x, fminf(rgb.y, rgb.z));
    v3 h_rgb = V3(sdivf(rgb.x - mn, mx), sdivf(rgb.y - mn, mx), sdivf(rgb.z - mn, mx));
    h_rgb = narrow_hue_angles(h_rgb);
    v3 w = V3(rw, 1.0f, bw);
    float wl = len3(w);
    w = V3(w.x / wl, w.y / wl, w.z / wl);
    w = mul3(w, V3(fmaxf(rgb.x, 1e-5f), fmaxf(rgb.y, 1e-5f), fmaxf(rgb.z, 1e-5f)));
    float lumv = len3(w);
    v3 rats = V3(sdivf(rgb.x, lumv), sdivf(rgb.y, lumv), sdivf(rgb.z, lumv));
    float ts = tonescale_fwd(lumv, m, s, c);
    ts = flare_fwd(ts, fl);
    ts *= ds;
    float dch_s = dch / s;
    float ccf = sdivf(1.0f, lumv * dch_s + 1.0f);
    float toe_ccf = (float)(0.0 + 1.0) * sdivf(lumv, lumv + dch_toe) * ccf;
    v3 hs_w = scl3(h_rgb, 1.0f - ccf);
    rats = V3(rats.x + hs_w.z * hs_b - hs_w.y * hs_g, rats.y + hs_w.x * hs_r - hs_w.z * hs_b, rats.z + hs_w.y * hs_g - hs_w.x * hs_r);
    float omt = 1.0f - toe_ccf;
    rats = V3(omt + rats.x * toe_ccf, omt + rats.y * toe_ccf, omt + rats.z * toe_ccf);
    rats = V3(fmaxf(rats.x, 0.0f), fmaxf(rats.y, 0.0f), fmaxf(rats.z, 0.0f));
    float rmx = fmaxf(rats.x, fmaxf(rats.y, rats.z)), rmn = fminf(rats.x, fminf(rats.y, rats.z));
    float rats_ch = sdivf(rmx - rmn, rmx);
    float chf = spow_ti(rats_ch * ts, v_p);
    v3 rn = V3(sdivf(rats.x, rmx), sdivf(rats.y, rmx), sdivf(rats.z, rmx));
    rats = add3(scl3(rn, chf), scl3(rats, 1.0f - chf));
    rgb = scl3(rats, ts);
    rgb = V3(fminf(rgb.x, clamp_max), fminf(rgb.y, clamp_max), fminf(rgb.z, clamp_max));
    return rgb;
}

/* ---- AgX (lib/AgX.py) ---- */
typedef struct { float m[3][3]; } m33;
static v3 m33_mul(const m33 *a, v3 v) {
    return V3((a->m[0][0] * v.x + a->m[0][1] * v.y) + a->m[0][2] * v.z, (a->m[1][0] * v.x + a->m[1][1] * v.y) + a->m[1][2] * v.z,
              (a->m[2][0] * v.x + a->m[2][1] * v.y) + a->m[2][2] * v.z);
}
/* AgX.py:24-43 */
static m33 InverseMat(const m33 *mm) {
    const float (*m)[3] = mm->m;
    float d = m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    float id = 1.0f / d;
    m33 c;
    c.m[0][0] = id * (m[1][1] * m[2][2] - m[2][1] * m[1][2]);
    c.m[0][1] = id * (m[0][2] * m[2][1] - m[0][1] * m[2][2]);
    c.m[0][2] = id * (m[0][1] * m[1][2] - m[0][2] * m[1][1]);
    c.m[1][0] = id * (m[1][2] * m[2][0] - m[1][0] * m[2][2]);
    c.m[1][1] = id * (m[0][0] * m[2][2] - m[0][2] * m[2][0]);
    c.m[1][2] = id * (m[1][0] * m[0][2] - m[0][0] * m[1][2]);
    c.m[2][0] = id * (m[1][0] * m[2][1] - m[2][0] * m[1][1]);
    c.m[2][1] = id * (m[2][0] * m[0][1] - m[0][0] * m[2][1]);
    c.m[2][2] = id * (m[0][0] * m[1][1] - m[1][0] * m[0][1]);
    return c;
}
/* AgX.py:45-58 */
static v3 Unproject(v2 xy) {
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    if (xy.y != 0.0f) { Y = 1.0f; X = (xy.x * Y) / xy.y; Z = ((1.0f - xy.x - xy.y) * Y) / xy.y; }
    return V3(X, Y, Z);
}
/* AgX.py:60-75 */
static m33 PrimariesToMatrix(v2 r, v2 g, v2 b, v2 w) {
    v3 R = Unproject(r), G = Unproject(g), B = Unproject(b), W = Unproject(w);
    m33 t = { { { R.x, G.x, B.x }, { 1.0f, 1.0f, 1.0f }, { R.z, G.z, B.z } } };
    m33 inv = InverseMat(&t);
    v3 sc = m33_mul(&inv, W);
    m33 o = { { { sc.x * R.x, sc.y * G.x, sc.z * B.x }, { sc.x * R.y, sc.y * G.y, sc.z * B.y }, { sc.x * R.z, sc.y * G.z, sc.z * B.z } } };
    return o;
}
static v2 V2(float x, float y) { v2 r = { x, y }; return r; }
/* AgX.py:77-85 */
static m33 ComputeCompressionMatrix(v2 r, v2 g, v2 b, v2 w, float compression) {
    float sf = 1.0f / (1.0f - compression);
    v2 R = V2((r.x - w.x) * sf + w.x, (r.y - w.y) * sf + w.y);
    v2 G = V2((g.x - w.x) * sf + w.x, (g.y - w.y) * sf + w.y);
    v2 B = V2((b.x - w.x) * sf + w.x, (b.y - w.y) * sf + w.y);
    return PrimariesToMatrix(R, G, B, w);
}
/* AgX.py:98-129 */
static float AgXScale(float xp, float yp, float sp, float power) {
    return pow_ti(pow_ti((sp * xp), -power) * (pow_ti((sp * (xp / yp)), power) - 1.0f), -1.0f / power);
}
static float AgXHyperbolic(float x, float power) { return x / pow_ti(1.0f + pow_ti(x, power), 1.0f / power); }
static float AgXTerm(float x, float xp, float sp, float scale) { return (sp * (x - xp)) / scale; }
static float AgXFullCurve(float x, float xp, float yp, float sp, float toe, float shoulder) {
    float sxp = x >= xp ? 1.0f - xp : xp, syp = x >= xp ? 1.0f - yp : yp;
    float toe_scale = AgXScale(sxp, syp, sp, toe), shoulder_scale = AgXScale(sxp, syp, sp, shoulder);
    float scale = x >= xp ? shoulder_scale : -toe_scale;
    float curve;
    if (scale < 0.0f) curve = scale * AgXHyperbolic(AgXTerm(x, xp, sp, scale), toe) + yp;
    else curve = scale * AgXHyperbolic(AgXTerm(x, xp, sp, scale), shoulder) + yp;
    return curve;
}
/* AgX.py:131-160 */
static v3 agx_display_transform(v3 col) {
    const float MIDDLE_GREY = 0.18f, SLOPE = 2.3f, TOE = 1.9f, SHOULDER = 3.1f, COMPRESSION = 0.15f, MIN_EV = -10.0f, MAX_EV = 6.5f, SATURATION = 1.4f;
    v2 pr = V2(0.64f, 0.33f), pg = V2(0.3f, 0.6f), pb = V2(0.15f, 0.06f), pw = V2(0.3127f, 0.3290f);
    m33 sRGB_to_XYZ = PrimariesToMatrix(pr, pg, pb, pw);
    m33 adjusted_to_XYZ = ComputeCompressionMatrix(pr, pg, pb, pw, COMPRESSION);
    m33 XYZ_to_adjusted = InverseMat(&adjusted_to_XYZ);
    v3 xyz = m33_mul(&sRGB_to_XYZ, col);
    v3 adj = m33_mul(&XYZ_to_adjusted, xyz);
    float x_pivot = (float)(10.0 / (6.5 - -10.0)), y_pivot = 0.5f;
    float total = MAX_EV - MIN_EV;
    v3 lg = V3(clampf(log2_ti(adj.x / MIDDLE_GREY), MIN_EV, MAX_EV), clampf(log2_ti(adj.y / MIDDLE_GREY), MIN_EV, MAX_EV), clampf(log2_ti(adj.z / MIDDLE_GREY), MIN_EV, MAX_EV));
    lg = V3((lg.x - MIN_EV) / total, (lg.y - MIN_EV) / total, (lg.z - MIN_EV) / total);
    v3 o = V3(AgXFullCurve(lg.x, x_pivot, y_pivot, SLOPE, TOE, SHOULDER), AgXFullCurve(lg.y, x_pivot, y_pivot, SLOPE, TOE, SHOULDER), AgXFullCurve(lg.z, x_pivot, y_pivot, SLOPE, TOE, SHOULDER));
    o = V3(clampf(o.x, 0.0f, 1.0f), clampf(o.y, 0.0f, 1.0f), clampf(o.z, 0.0f, 1.0f));
    o = mix3(lum3(o), o, SATURATION);
    return V3(clampf(o.x, 0.0f, 1.0f), clampf(o.y, 0.0f, 1.0f), clampf(o.z, 0.0f, 1.0f));
}
/* renderer.py:333-344 */
static v3 camera_response(const float *crf, int n_crf, int selected, int crf_count, v3 t) {
    t = V3(clampf(t.x, 0.0f, 1.0f), clampf(t.y, 0.0f, 1.0f), clampf(t.z, 0.0f, 1.0f));
    float slice_v = ((float)selected + 0.5f) / (float)crf_count;
    float u_off = (float)(0.5 / 1024.0);
    v3 ul = V3(fminf(t.x + u_off, 1.0f - u_off), fminf(t.y + u_off, 1.0f - u_off), fminf(t.z + u_off, 1.0f - u_off));
    crf_ctx k = { crf, n_crf };
    float r = bilinear(fetch_crf, &k, 1024, n_crf, 3, ul.x, slice_v).c[0];
    float g = bilinear(fetch_crf, &k, 1024, n_crf, 3, ul.y, slice_v).c[1];
    float b = bilinear(fetch_crf, &k, 1024, n_crf, 3, ul.z, slice_v).c[2];
    return V3(clampf(r, 0.0f, 1.0f), clampf(g, 0.0f, 1.0f), clampf(b, 0.0f, 1.0f));
}
/* colour.py:74-79 */
static float srgb_transfer1(float lin) {
    float lo = lin * 12.92f;
    float hi = (pow_ti(fabsf(lin), (float)(1.0 / 2.4)) * 1.055f) - 0.055f;
    float st = 0.0031308f >= lin ? 1.0f : 0.0f; /* step(linear, 0.0031308) */
    return hi * (1.0f - st) + lo * st;
}
/* renderer.py:346-365: one pixel of _render_to_image */
static v3 resolve_pixel(const orc_scene *s, int i, int j, v3 color, int samples) {
    float u = 1.0f * (float)i / (float)s->W, v = 1.0f * (float)j / (float)s->H;
    float du = u - s->vig_cx, dv = v - s->vig_cy;
    float darken = 1.0f - s->vig_strength * fmaxf(sqrtf(du * du + dv * dv) - s->vig_radius, 0.0f);
    float ex = pow_ti(2.0f, s->exposure), ns = (float)samples;
    v3 lin = V3(color.x / ns * darken * ex, color.y / ns * darken * ex, color.z / ns * darken * ex);
    v3 tm = s->tonemapper == 1 ? agx_display_transform(lin) : openDR_transform(lin.x, lin.y, lin.z);
    v3 cam = camera_response(s->crf, s->n_crf, s->selected_crf, s->crf_count, tm);
    v3 g = V3(pow_ti(cam.x, s->gamma), pow_ti(cam.y, s->gamma), pow_ti(cam.z, s->gamma));
    return V3(srgb_transfer1(g.x), srgb_transfer1(g.y), srgb_transfer1(g.z));
}

/* ================================================================= API ==
 * Batch entry points (ctypes).  All arrays are caller-owned host memory. */
#define LD3(p, i) V3((p)[3 * (i)], (p)[3 * (i) + 1], (p)[3 * (i) + 2])
#define ST3(p, i, v) do { (p)[3 * (i)] = (v).x; (p)[3 * (i) + 1] = (v).y; (p)[3 * (i) + 2] = (v).z; } while (0)

ORC_API void orc_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }
ORC_API void orc_rsi(int n, const float *pos, const float *dir, const float *r, float *out) {
    for (int i = 0; i < n; ++i) { v2 o = rsi(LD3(pos, i), LD3(dir, i), r[i]); out[2 * i] = o.x; out[2 * i + 1] = o.y; }
}
ORC_API void orc_density(int n, const float *h, float *out) {
    for (int i = 0; i < n; ++i) { v3 d = get_density(h[i]); ST3(out, i, d); }
}
/* out[5n]: sigma_rayleigh, sigma_mie, sigma_ozone, planck(5778), planck(2700) */
ORC_API void orc_spectra(int n, const float *wl, const float *o3, float *out) {
    for (int i = 0; i < n; ++i) {
        out[5 * i] = spectra_extinction_rayleigh(wl[i]); out[5 * i + 1] = spectra_extinction_mie(wl[i]);
        out[5 * i + 2] = spectra_extinction_ozone(wl[i], o3); out[5 * i + 3] = plancks(5778.0f, wl[i]); out[5 * i + 4] = plancks(2700.0f, wl[i]);
    }
}
ORC_API void orc_phase_eval(int n, const float *ray_dir, const float *light_dir, const int32_t *id, const int32_t *reduce, float *out) {
    for (int i = 0; i < n; ++i) out[i] = evaluate_phase(LD3(ray_dir, i), LD3(light_dir, i), id[i], reduce[i]);
}
/* rand: 4 u32 per item (consumed in order); out_dir[3n], out_w[n] */
ORC_API void orc_phase_sample(int n, const float *ray_dir, const int32_t *id, const int32_t *reduce, const uint32_t *rand, float *out_dir, float *out_w) {
    for (int i = 0; i < n; ++i) {
        orc_rng r; memset(&r, 0, sizeof r); r.list = rand + 4 * i;
        v3 d = sample_phase(LD3(ray_dir, i), id[i], reduce[i], &r, &out_w[i]); ST3(out_dir, i, d);
    }
}
/* kind 0: sample_cone_oriented(cos_max=p, n), 1: sample_hemisphere_cosine_weighted(n); rand: 2 u32 per item */
ORC_API void orc_dir_sample(int n, int kind, const float *nrm, float p, const uint32_t *rand, float *out) {
    for (int i = 0; i < n; ++i) {
        orc_rng r; memset(&r, 0, sizeof r); r.list = rand + 2 * i;
        v3 d = kind == 0 ? sample_cone_oriented(p, LD3(nrm, i), &r) : sample_hemisphere_cosine_weighted(LD3(nrm, i), &r);
        ST3(out, i, d);
    }
}
/* in: albedo, ocean, bathy [n], v,nrm,l [3n]; out[2n] = brdf, n_dot_l */
ORC_API void orc_brdf(int n, const float *albedo, const float *ocean, const float *bathy, const float *v, const float *nrm, const float *l, float *out) {
    for (int i = 0; i < n; ++i) out[2 * i] = earth_brdf(albedo[i], ocean[i], bathy[i], LD3(v, i), LD3(nrm, i), LD3(l, i), &out[2 * i + 1]);
}
ORC_API void orc_srgb_to_spectrum(int n, const uint16_t *lut, const float *rgb, const float *wl, float *out) {
    for (int i = 0; i < n; ++i) out[i] = srgb_to_spectrum(lut, LD3(rgb, i), wl[i]);
}
/* out[5n]: wavelength, response xyz, rcp_pdf */
ORC_API void orc_spectrum_sample(int n, const float *cie, const uint32_t *rand, float *out) {
    for (int i = 0; i < n; ++i) {
        float xi = (float)(rand[i] >> 8) * (1.0f / 16777216.0f); v3 resp;
        spectrum_sample(cie, xi, &out[5 * i], &resp, &out[5 * i + 4]); out[5 * i + 1] = resp.x; out[5 * i + 2] = resp.y; out[5 * i + 3] = resp.z;
    }
}
ORC_API void orc_tex_fetch(int n, const orc_tex *t, const float *pos, float *out) {
    for (int i = 0; i < n; ++i) { texel4 q = sample_sphere_texture(t, LD3(pos, i), NULL); memcpy(out + 4 * i, q.c, 16); }
}
ORC_API void orc_cast_dir(int n, const orc_scene *s, const float *u, const float *v, const uint32_t *rand, float *out) {
    for (int i = 0; i < n; ++i) {
        float a = (float)(rand[2 * i] >> 8) * (1.0f / 16777216.0f), b = (float)(rand[2 * i + 1] >> 8) * (1.0f / 16777216.0f);
        v3 d = get_cast_dir(s, u[i], v[i], a, b); ST3(out, i, d);
    }
}
ORC_API void orc_opendrt(int n, const float *rgb, float *out) { for (int i = 0; i < n; ++i) { v3 o = openDR_transform(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]); ST3(out, i, o); } }
ORC_API void orc_agx(int n, const float *rgb, float *out) { for (int i = 0; i < n; ++i) { v3 o = agx_display_transform(LD3(rgb, i)); ST3(out, i, o); } }
ORC_API void orc_crf(int n, const orc_scene *s, const float *rgb, float *out) {
    for (int i = 0; i < n; ++i) { v3 o = camera_response(s->crf, s->n_crf, s->selected_crf, s->crf_count, LD3(rgb, i)); ST3(out, i, o); }
}
ORC_API void orc_srgb_oetf(int n, const float *x, float *out) { for (int i = 0; i < n; ++i) out[i] = srgb_transfer1(x[i]); }
/* accum, out: row-major [y][x][3] */
ORC_API void orc_resolve(const orc_scene *s, const float *accum, int samples, float *out) {
    for (int j = 0; j < s->H; ++j) for (int i = 0; i < s->W; ++i) {
        size_t k = (size_t)j * s->W + i; v3 o = resolve_pixel(s, i, j, LD3(accum, k), samples); ST3(out, k, o);
    }
}
ORC_API void orc_intersect_land(int n, const orc_scene *s, const float *pos, const float *dir, float *out) {
    for (int i = 0; i < n; ++i) out[i] = intersect_land(s, LD3(pos, i), LD3(dir, i), s->land_height_scale, NULL);
}
/* out[2n] = (distance or -1, SDF evaluations): 250 evaluations with a hit = the march ran into its iteration cap */
ORC_API void orc_intersect_land_iters(int n, const orc_scene *s, const float *pos, const float *dir, float *out) {
    for (int i = 0; i < n; ++i) {
        orc_counters c; memset(&c, 0, sizeof c);
        out[2 * i] = intersect_land(s, LD3(pos, i), LD3(dir, i), s->land_height_scale, &c);
        out[2 * i + 1] = (float)c.sdf_evals;
    }
}
ORC_API void orc_land_normal(int n, const orc_scene *s, const float *pos, float *out) {
    for (int i = 0; i < n; ++i) { v3 o = land_normal(s, LD3(pos, i), s->land_height_scale, NULL); ST3(out, i, o); }
}
ORC_API void orc_cloud_limits(int n, const float *pos, const float *dir, const float *land, float *out) {
    for (int i = 0; i < n; ++i) intersect_cloud_limits(LD3(pos, i), LD3(dir, i), land[i], &out[2 * i], &out[2 * i + 1]);
}
ORC_API void orc_clouds_density(int n, const orc_scene *s, const float *pos, float *out) {
    for (int i = 0; i < n; ++i) out[i] = get_clouds_density(s, LD3(pos, i), NULL);
}
/* out[6n]: albedo_srgb rgb, ocean, bathymetry, emissive */
ORC_API void orc_land_material(int n, const orc_scene *s, const float *pos, float *out) {
    for (int i = 0; i < n; ++i) { land_material m = get_land_material(s, LD3(pos, i), NULL); ST3(out, 2 * i, m.albedo_srgb); out[6 * i + 3] = m.ocean; out[6 * i + 4] = m.bathymetry; out[6 * i + 5] = m.emissive; }
}
ORC_API void orc_raymarch_T(int n, const float *pos, const float *dir, const float *ext, float *out) {
    for (int i = 0; i < n; ++i) out[i] = ray_march_transmittance(LD3(pos, i), LD3(dir, i), LD3(ext, i));
}
/* Stochastic sub-paths driven by the Philox stream (key=(seed,i), sample 0, bounce 1).
 * kind 0: sample_interaction -> out[3n] = event, t, id ; kind 1: sample_transmittance -> out[3n] = T,0,0 */
ORC_API void orc_tracking(int n, int kind, const orc_scene *s, const float *pos, const float *dir, const float *land, const float *wl, uint32_t seed, float *out) {
    v3 d0 = get_density(0.0f); float o3max = get_ozone_density(25000.0f);
    for (int i = 0; i < n; ++i) {
        orc_rng r; memset(&r, 0, sizeof r); r.key0 = seed; r.key1 = (uint32_t)i; rng_bounce(&r, 1);
        v4 ext = { spectra_extinction_rayleigh(wl[i]), spectra_extinction_mie(wl[i]), spectra_extinction_ozone(wl[i], s->o3), CLOUDS_EXTINCT };
        float mr = (ext.x * d0.x + ext.y * d0.y) + ext.z * o3max, mc = ext.w * CLOUDS_DENSITY;
        if (kind == 0) { float t; int id; int ev = sample_interaction(s, LD3(pos, i), LD3(dir, i), land[i], ext, mr, mc, &r, NULL, &t, &id); out[3 * i] = (float)ev; out[3 * i + 1] = t; out[3 * i + 2] = (float)id; }
        else { out[3 * i] = sample_transmittance(s, LD3(pos, i), LD3(dir, i), land[i], ext, mr, mc, &r, NULL); out[3 * i + 1] = 0; out[3 * i + 2] = 0; }
    }
}
/* Individual path samples: out[5n] = rgb contribution, wavelength, radiance */
ORC_API void orc_trace_paths(const orc_scene *s, int n, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out, orc_counters *cnt) {
    scene_params sc = make_scene_params(s);
    for (int i = 0; i < n; ++i) { v3 c = render_sample(s, &sc, px[i], py[i], sample[i], seed, cnt, &out[5 * i + 3], &out[5 * i + 4]); out[5 * i] = c.x; out[5 * i + 1] = c.y; out[5 * i + 2] = c.z; }
}

ORC_API void orc_trace_paths2(const orc_scene *s, int n, const int32_t *px, const int32_t *py, const uint32_t *sample, uint32_t seed, float *out, orc_counters *cnt, int integrator) {
    scene_params sc = make_scene_params(s);
    for (int i = 0; i < n; ++i) { v3 c = render_sample2(s, &sc, px[i], py[i], sample[i], seed, cnt, &out[5 * i + 3], &out[5 * i + 4], integrator); out[5 * i] = c.x; out[5 * i + 1] = c.y; out[5 * i + 2] = c.z; }
}
/* the preview's two marching routines on explicit inputs (pathtracer.py:471-541): out[n][2] = (in_scatter, transmittance) of
 * ray_marh_atmos over [t0[i], t1[i]] with light direction sun[i]; outT[n] = ray_march_transmittance(pos, sun) */
ORC_API void orc_ray_march(const orc_scene *s, int n, const float *pos, const float *dir, const float *t0, const float *t1, const float *sun,
                           const float *wl, float *out2, float *outT) {
    for (int i = 0; i < n; ++i) {
        v3 ext = V3(spectra_extinction_rayleigh(wl[i]), spectra_extinction_mie(wl[i]), spectra_extinction_ozone(wl[i], s->o3));
        v2 scat; scat.x = ext.x * RAYLEIGH_ALBEDO; scat.y = ext.y * AEROSOL_ALBEDO;
        v3 p = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), d = V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), l = V3(sun[3 * i], sun[3 * i + 1], sun[3 * i + 2]);
        ray_march_atmos(p, d, t0[i], t1[i], l, ext, scat, &out2[2 * i], &out2[2 * i + 1]);
        outT[i] = ray_march_transmittance(p, l, ext);
    }
}

/* Multi-threaded render (the CPU baseline): accum[y][x][3] += spp samples per pixel in the
 * window [x0,x0+w) x [y0,y0+h); optional accum2 accumulates squared luminance-like moments per channel. */
typedef struct {
    const orc_scene *s; scene_params sc; int x0, y0, w, h, spp; uint32_t first_sample, seed;
    float *accum, *accum2; volatile int32_t *next_row; orc_counters cnt; int integrator;
} render_job;
static void *render_worker(void *arg) {
    render_job *j = (render_job *)arg;
    for (;;) {
        int row = __atomic_fetch_add(j->next_row, 1, __ATOMIC_RELAXED);
        if (row >= j->h) break;
        int y = j->y0 + row;
        for (int x = j->x0; x < j->x0 + j->w; ++x) {
            size_t k = ((size_t)y * j->s->W + x) * 3;
            for (int sp = 0; sp < j->spp; ++sp) {
                v3 c = render_sample2(j->s, &j->sc, x, y, j->first_sample + (uint32_t)sp, j->seed, &j->cnt, NULL, NULL, j->integrator);
                j->accum[k] += c.x; j->accum[k + 1] += c.y; j->accum[k + 2] += c.z;
                if (j->accum2) { j->accum2[k] += c.x * c.x; j->accum2[k + 1] += c.y * c.y; j->accum2[k + 2] += c.z * c.z; }
            }
        }
    }
    return NULL;
}
ORC_API void orc_render2(const orc_scene *s, int x0, int y0, int w, int h, int spp, uint32_t first_sample, uint32_t seed,
                         float *accum, float *accum2, int nthreads, orc_counters *cnt_out, int integrator);
ORC_API void orc_render(const orc_scene *s, int x0, int y0, int w, int h, int spp, uint32_t first_sample, uint32_t seed,
                        float *accum, float *accum2, int nthreads, orc_counters *cnt_out) {
    orc_render2(s, x0, y0, w, h, spp, first_sample, seed, accum, accum2, nthreads, cnt_out, 0);
}
ORC_API void orc_render2(const orc_scene *s, int x0, int y0, int w, int h, int spp, uint32_t first_sample, uint32_t seed,
                         float *accum, float *accum2, int nthreads, orc_counters *cnt_out, int integrator) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    volatile int32_t next_row = 0;
    render_job *jobs = (render_job *)calloc((size_t)nthreads, sizeof(render_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    scene_params sc = make_scene_params(s);
    for (int t = 0; t < nthreads; ++t) {
        render_job *j = &jobs[t];
        j->s = s; j->sc = sc; j->x0 = x0; j->y0 = y0; j->w = w; j->h = h; j->spp = spp; j->first_sample = first_sample; j->seed = seed;
        j->accum = accum; j->accum2 = accum2; j->next_row = &next_row; j->integrator = integrator;
        pthread_create(&th[t], NULL, render_worker, j);
    }
    orc_counters tot; memset(&tot, 0, sizeof tot);
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        const uint64_t *a = (const uint64_t *)&jobs[t].cnt; uint64_t *b = (uint64_t *)&tot;
        for (size_t q = 0; q < sizeof(orc_counters) / 8; ++q) b[q] += a[q];
    }
    if (cnt_out) *cnt_out = tot;
    free(jobs); free(th);
}
ORC_API int orc_abi_version(void) { return 1; }
