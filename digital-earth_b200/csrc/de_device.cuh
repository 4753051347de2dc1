// de_device.cuh -- device-side math library of the B200 spectral path tracer.
//
// Compiled twice (see build.py):
//   * DE_EXACT=1, -fmad=false, IEEE div/sqrt, accurate libdevice functions  -> namespace de_exact
//     ("parity" flavour: evaluates the reference's expressions in source order so results are
//      comparable with the CPU oracle to ~1e-6 relative);
//   * DE_EXACT=0, -use_fast_math (FMA contraction, MUFU intrinsics)         -> namespace de_fast
//     (product flavour used by the wavefront / megakernel integrators).
// Reference locations are cited as file:line relative to AntonioFerreras/Digital-Earth.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "de_scene.h"

#ifndef DE_EXACT
#define DE_EXACT 0
#endif
#if DE_EXACT
#define DE_NS de_exact
#else
#define DE_NS de_fast
#endif
#define DE_DEV __device__ __forceinline__

namespace DE_NS {

// ---------------------------------------------------------------- constants (volume_rendering_models.py:8-44)
constexpr float kPlanetR = 6371000.0f;
constexpr float kAtmosUpper = 6481000.0f;
constexpr float kCloudsLower = 6375000.0f;
constexpr float kCloudsUpper = 6381000.0f;
constexpr float kCloudsThickness = 6000.0f;
constexpr float kCloudsExtinct = 0.1f;
constexpr float kCloudsDensity = 0.029f;
constexpr float kMieAsymmetry = 3000.0f;
constexpr float kRayleighAlbedo = 1.0f, kAerosolAlbedo = 0.95f;  // volume_rendering_models.py:27-28
constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kFourPi = 12.56637061435917295385f;
constexpr int kRayleigh = 0, kMie = 1, kOzone = 2, kCloud = 3, kIsoCloud = 4;
constexpr int kNullEvent = 0, kAbsorbEvent = 1, kScatterEvent = 2;
constexpr int kLambdaBins = 512;  // mid = j/512, j in 1..511 (colour.py:26: 8 bisection steps)

// ---------------------------------------------------------------- float3 helpers (source-order arithmetic)
DE_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
DE_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
DE_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
DE_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
DE_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
DE_DEV float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
DE_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
DE_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
DE_DEV float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
DE_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
DE_DEV float3 normalize(float3 a) {
#if DE_EXACT
    float inv = 1.0f / length(a);  // taichi Vector.normalized(): invlen * self
#else
    float inv = rsqrtf(dot(a, a));
#endif
    return a * inv;
}
DE_DEV float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DE_DEV float sqr(float x) { return x * x; }
DE_DEV float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
DE_DEV float saturate(float x) { return clampf(x, 0.0f, 1.0f); }
DE_DEV float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
DE_DEV float3 mix3(float3 x, float3 y, float a) { return x * (1.0f - a) + y * a; }
DE_DEV float smoothstep(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// ops.pow: exponent 2 is a multiply in every backend (LLVM/NVVM fold); otherwise powf
DE_DEV float pow_ti(float x, float y) { return y == 2.0f ? x * x : powf(x, y); }
DE_DEV float log2_ti(float x) { return logf(x) / 0.6931471805599453f; }  // taichi.math.log2

// ---------------------------------------------------------------- Philox4x32-10 stream
// RNG contract (include/de_api.h): slot i of (pixel, sample, bounce) = word i&3 of
// Philox4x32-10(key=(seed,pixel), ctr=(sample,bounce,i>>2,0)); align() / skip() implement the two
// consumption rules (consumers start on a block boundary; a ratio-tracking trip owns two slots).
struct Rng {
    uint32_t key0, key1, sample, bounce, draw;
    uint32_t b0, b1, b2, b3;
    bool valid;
    DE_DEV void init(uint32_t seed, uint32_t pixel, uint32_t sample_index) {
        key0 = seed; key1 = pixel; sample = sample_index; bounce = 0; draw = 0; valid = false;
    }
    DE_DEV void set_bounce(uint32_t b) { bounce = b; draw = 0; valid = false; }
    DE_DEV void align() { draw = (draw + 3u) & ~3u; valid = false; }
    DE_DEV void skip() { draw += 1u; valid = false; }
    DE_DEV void refill() {
        uint32_t c0 = sample, c1 = bounce, c2 = draw >> 2, c3 = 0u, k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        b0 = c0; b1 = c1; b2 = c2; b3 = c3;
        valid = true;
    }
    DE_DEV uint32_t next_u32() {
        uint32_t lane = draw & 3u;
        if (lane == 0u || !valid) refill();
        ++draw;
        return lane == 0u ? b0 : (lane == 1u ? b1 : (lane == 2u ? b2 : b3));
    }
    // ti.random(f32) = (u32 >> 8) * 2^-24
    DE_DEV float next() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};
struct ListRng {  // explicit draws for the unit-test hooks
    const uint32_t *p;
    DE_DEV float next() { return (float)((*p++) >> 8) * (1.0f / 16777216.0f); }
    DE_DEV void align() {}
    DE_DEV void skip() {}
};
DE_DEV float u32_to_unit(uint32_t v) { return (float)(v >> 8) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------- counters
enum { C_PATHS = 0, C_SEGMENTS, C_RMO, C_CLOUD, C_SDF, C_TEX, C_SURF, C_DRAWS, C_COUNT };
struct Counters {
    unsigned int v[C_COUNT];
    DE_DEV void clear() {
#pragma unroll
        for (int i = 0; i < C_COUNT; ++i) v[i] = 0;
    }
    DE_DEV void flush(unsigned long long *g) {
        if (!g) return;
#pragma unroll
        for (int i = 0; i < C_COUNT; ++i)
            if (v[i]) atomicAdd(&g[i], (unsigned long long)v[i]);
    }
};
#define DE_COUNT(c, k) do { if (COUNT) (c).v[k]++; } while (0)

// ---------------------------------------------------------------- textures
// Manual FP32 bilinear: texel centres (i+.5)/N, clamp-to-edge, lerp a+f*(b-a), x then y.
struct Bilin { int x0, x1, y0, y1; float fx, fy; };
DE_DEV Bilin bilin_setup(int w, int h, float u, float v) {
    Bilin b;
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    b.fx = x - x0f; b.fy = y - y0f;
    int x0 = (int)x0f, y0 = (int)y0f;
    b.x0 = min(max(x0, 0), w - 1); b.x1 = min(max(x0 + 1, 0), w - 1);
    b.y0 = min(max(y0, 0), h - 1); b.y1 = min(max(y0 + 1, 0), h - 1);
    return b;
}
DE_DEV float lerp2(float t00, float t10, float t01, float t11, float fx, float fy) {
    float a = t00 + fx * (t10 - t00), b = t01 + fx * (t11 - t01);
    return a + fy * (b - a);
}
DE_DEV float unorm8(uint8_t t) {
#if DE_EXACT
    return (float)t / 255.0f;  // renderer.py:173: cast(u8,f32)/255.0
#else
    return (float)t * (1.0f / 255.0f);
#endif
}
#if !DE_EXACT
// Product flavour: the 2x2 footprint comes from ONE texture instruction (tex2Dgather on the
// block-linear copy, unorm8 -> float in the TEX unit); the weights stay ours (manual FP32 lerp,
// as the oracle defines the filter).  The gather is addressed at the footprint's centre corner
// (i0+1, j0+1), so texel selection never depends on the unit's fixed-point rounding, and the
// clamp address mode reproduces the clamped indices at the borders.
DE_DEV float tex_r8_gather(const DevTex &t, float u, float v) {
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    float4 g = tex2Dgather<float4>(t.obj, x0f + 1.0f, y0f + 1.0f, 0);  // (x,y,z,w) = t01, t11, t10, t00
    return lerp2(g.w, g.z, g.x, g.y, fx, fy);
}
DE_DEV float3 tex_rgb8_gather(const DevTex &t, float u, float v) {
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    float4 r = tex2Dgather<float4>(t.obj, x0f + 1.0f, y0f + 1.0f, 0);
    float4 g = tex2Dgather<float4>(t.obj, x0f + 1.0f, y0f + 1.0f, 1);
    float4 b = tex2Dgather<float4>(t.obj, x0f + 1.0f, y0f + 1.0f, 2);
    return f3(lerp2(r.w, r.z, r.x, r.y, fx, fy), lerp2(g.w, g.z, g.x, g.y, fx, fy), lerp2(b.w, b.z, b.x, b.y, fx, fy));
}
#endif
DE_DEV float tex_r8(const DevTex &t, float u, float v) {
#if !DE_EXACT
#if DE_TEX_OBJ_ONLY  // the wavefront kernel: every uploaded map has its texture object (de_upload_texture fails otherwise)
    return tex_r8_gather(t, u, v);
#else
    if (t.obj) return tex_r8_gather(t, u, v);
#endif
#endif
    Bilin b = bilin_setup(t.w, t.h, u, v);
    const uint8_t *r0 = t.data + (size_t)b.y0 * t.w, *r1 = t.data + (size_t)b.y1 * t.w;
    return lerp2(unorm8(__ldg(r0 + b.x0)), unorm8(__ldg(r0 + b.x1)), unorm8(__ldg(r1 + b.x0)), unorm8(__ldg(r1 + b.x1)), b.fx, b.fy);
}
DE_DEV float3 tex_rgb8(const DevTex &t, float u, float v) {
#if !DE_EXACT
#if DE_TEX_OBJ_ONLY
    return tex_rgb8_gather(t, u, v);
#else
    if (t.obj) return tex_rgb8_gather(t, u, v);
#endif
#endif
    Bilin b = bilin_setup(t.w, t.h, u, v);
    const uint8_t *p00 = t.data + ((size_t)b.y0 * t.w + b.x0) * 3, *p10 = t.data + ((size_t)b.y0 * t.w + b.x1) * 3;
    const uint8_t *p01 = t.data + ((size_t)b.y1 * t.w + b.x0) * 3, *p11 = t.data + ((size_t)b.y1 * t.w + b.x1) * 3;
    float3 o;
    o.x = lerp2(unorm8(__ldg(p00)), unorm8(__ldg(p10)), unorm8(__ldg(p01)), unorm8(__ldg(p11)), b.fx, b.fy);
    o.y = lerp2(unorm8(__ldg(p00 + 1)), unorm8(__ldg(p10 + 1)), unorm8(__ldg(p01 + 1)), unorm8(__ldg(p11 + 1)), b.fx, b.fy);
    o.z = lerp2(unorm8(__ldg(p00 + 2)), unorm8(__ldg(p10 + 2)), unorm8(__ldg(p01 + 2)), unorm8(__ldg(p11 + 2)), b.fx, b.fy);
    return o;
}
#if !DE_EXACT
// Product flavour of atan2 / asin for the equirect mapping: minimax atan(a) = a*P(a^2) on [0,1]
// (max error 3.2e-7 rad) and Abramowitz-Stegun 4.4.46 for asin (2.2e-8 rad); the libdevice
// versions were 18.6 % of all issued instructions (profiles/r1_wavefront.md).  3e-7 rad is
// 0.001 texel on a 21600-wide map.
DE_DEV float fast_atan2(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = __fdividef(mn, fmaxf(mx, 1e-30f));
    float s = a * a;
    float p = 0.006811773870140314f;
    p = fmaf(p, s, -0.03360416740179062f);
    p = fmaf(p, s, 0.07962362468242645f);
    p = fmaf(p, s, -0.1323333978652954f);
    p = fmaf(p, s, 0.19807815551757812f);
    p = fmaf(p, s, -0.3331736922264099f);
    p = fmaf(p, s, 0.9999961256980896f);
    float r = p * a;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = x < 0.0f ? 3.14159265358979324f - r : r;
    return copysignf(r, y);
}
DE_DEV float fast_asin(float x) {
    float ax = fminf(fabsf(x), 1.0f);
    float p = -0.0012624911f;
    p = fmaf(p, ax, 0.0066700901f);
    p = fmaf(p, ax, -0.0170881256f);
    p = fmaf(p, ax, 0.0308918810f);
    p = fmaf(p, ax, -0.0501743046f);
    p = fmaf(p, ax, 0.0889789874f);
    p = fmaf(p, ax, -0.2145988016f);
    p = fmaf(p, ax, 1.5707963050f);
    float r = 1.57079632679489662f - sqrtf(1.0f - ax) * p;
    return copysignf(r, x);
}
#endif
// math_utils.py:25-28
DE_DEV float2 sphere_UV_map(float3 n) {
#if DE_EXACT
    return make_float2((atan2f(n.z, -n.x) / kPi + 1.0f) / 2.0f, asinf(n.y) / kPi + 0.5f);
#else
    return make_float2(fmaf(fast_atan2(n.z, -n.x), 0.5f / kPi, 0.5f), fmaf(fast_asin(n.y), 1.0f / kPi, 0.5f));
#endif
}
// math_utils.py:38-44 (uv -> fract(uv))
DE_DEV float2 sphere_uv(float3 pos) {
    float2 uv = sphere_UV_map(normalize(pos));
    uv.x = uv.x - floorf(uv.x);
    uv.y = uv.y - floorf(uv.y);
    return uv;
}
#if !DE_EXACT
// product flavour: position + its 1/|pos| (the caller already needs r), no second normalisation.
// u, v come out in [0,1]; the fract() of math_utils.py:44 only matters at the exact pole / date line
// (measure zero) and the clamp address mode keeps the fetch in range there.
DE_DEV float sample_sphere_r8_inv(const DevTex &t, float3 pos, float inv_r) {
    float u = fmaf(fast_atan2(pos.z, -pos.x), 0.5f / kPi, 0.5f);
    float v = fmaf(fast_asin(pos.y * inv_r), 1.0f / kPi, 0.5f);
    return tex_r8(t, u, v);
}
#endif
DE_DEV float sample_sphere_r8(const DevTex &t, float3 pos) { float2 uv = sphere_uv(pos); return tex_r8(t, uv.x, uv.y); }
DE_DEV float3 sample_sphere_rgb8(const DevTex &t, float3 pos) { float2 uv = sphere_uv(pos); return tex_rgb8(t, uv.x, uv.y); }
#if !DE_EXACT
// Upper bound of the cloud texture along the part [ts, tm] of a ray (product flavour).
// A straight ray projects onto a great-circle arc: longitude is monotone along it (arcs shorter than
// pi that stay away from the poles), latitude leaves the endpoint range by at most ~theta^2/8 * tan(lat).  The
// bound is the maximum of the dilated coarse map over that lat-long box; 1.0 (no information)
// whenever the box is unsafe (date-line crossing, polar caps, long arcs, too many cells).
#if defined(DE_WF_SHRINK) && DE_WF_SHRINK
__device__ __noinline__ float2 sphere_uv_ool(float px, float py, float pz);  // de_wavefront.cu: one shared copy of the equirect mapping (code size)
#define DE_CMAX_UV(p) sphere_uv_ool((p).x, (p).y, (p).z)
#else
#define DE_CMAX_UV(p) sphere_uv(p)
#endif
DE_DEV float cloud_segment_cmax(const DevScene &s, float3 o, float3 d, float ts, float tm) {
    if (!s.cloud_max) return 1.0f;
    float theta = (tm - ts) * (1.0f / 6375000.0f);
    if (!(theta < 0.25f)) return 1.0f;
    const float3 pa = o + d * ts, pb = o + d * tm;
    float2 a = DE_CMAX_UV(pa), b = DE_CMAX_UV(pb);
    if (fabsf(a.x - b.x) > 0.4f) return 1.0f;
    // latitude: sin(lat) is a sinusoid of amplitude <= 1 along the arc, so it leaves the end points' range by at most
    // 1 - cos(theta/2) <= theta^2/8; inside |lat| <= 81 deg (which the coarse test with the trivial bound theta/2 guarantees)
    // d(v)/d(sin lat) <= 1/(pi cos 81deg), hence 0.2544 theta^2 in v.  (The trivial bound alone made the box 10-60x too tall.)
    const float pad_coarse = theta * (0.5f / kPi) + 1e-4f;
    if (fminf(a.y, b.y) - pad_coarse < 0.05f || fmaxf(a.y, b.y) + pad_coarse > 0.95f) return 1.0f;
    const float pad = 0.2544f * theta * theta + 1e-4f;
    float vlo = fminf(a.y, b.y) - pad, vhi = fmaxf(a.y, b.y) + pad;
    float sx = (float)s.tex[3].w / (float)s.cm_b, sy = (float)s.tex[3].h / (float)s.cm_b;
    int cu0 = max((int)((fminf(a.x, b.x) - 1e-4f) * sx), 0), cu1 = min((int)((fmaxf(a.x, b.x) + 1e-4f) * sx), s.cm_w - 1);
    int cv0 = max((int)(vlo * sy), 0), cv1 = min((int)(vhi * sy), s.cm_h - 1);
    if ((cu1 - cu0 + 1) * (cv1 - cv0 + 1) > 48) return 1.0f;
    unsigned m = 0u;
    for (int cv = cv0; cv <= cv1; ++cv)
        for (int cu = cu0; cu <= cu1; ++cu) m = max(m, (unsigned)__ldg(s.cloud_max + cv * s.cm_w + cu));
    return (float)m * (1.0f / 255.0f);
}
// density majorant of get_clouds_density given a bound on the texture value (pathtracer.py:63-65)
DE_DEV float cloud_density_bound(float cmax) { return cmax > 0.0f ? fmaxf(cmax, 0.4f) * kCloudsDensity : 0.0f; }
#endif
// generic float LUT texture [h][w][nc] (CIE 441x2x3, CRF 1024xNx3)
DE_DEV float tex_f32(const float *d, int w, int h, int nc, int c, float u, float v) {
    Bilin b = bilin_setup(w, h, u, v);
    const float *r0 = d + (size_t)b.y0 * w * nc + c, *r1 = d + (size_t)b.y1 * w * nc + c;
    return lerp2(__ldg(r0 + b.x0 * nc), __ldg(r0 + b.x1 * nc), __ldg(r1 + b.x0 * nc), __ldg(r1 + b.x1 * nc), b.fx, b.fy);
}

// ---------------------------------------------------------------- geometry
// math_utils.py:17-23.  A miss yields a NaN pair (the select tests the sqrt); every caller only
// uses >0 / >=0 / <0 tests, which NaN fails like (-1,-1) would.
DE_DEV float2 rsi(float3 pos, float3 dir, float r) {
    float b = dot(pos, dir);
    float discr = b * b - dot(pos, pos) + r * r;
    discr = sqrtf(discr);
    if (discr < 0.0f) return make_float2(-1.0f, -1.0f);
    return make_float2(-b + -discr, -b + discr);
}

#if !DE_EXACT
// Product flavour: density bound AND interval of one cloud tracking pass over [ts, tm] (0 = nothing to track).
// Where the texture is <= cmax the cloud layer ends at height 0.2 + 0.8 cmax of the shell (pathtracer.py:63: the density is
// zero unless hgt - 0.2 < 0.8 c), so the part of the pass above that sphere could only produce null collisions: cutting it
// leaves the distribution of real collisions unchanged and removes more than half of the cloud steps
// (profiles/r1_bench.md).  The 1e-3 (6 m) margin covers rounding of hgt; a ray that misses the sphere yields rsi's NaN pair.
// pos_noise: bound on |fl(o + d t) - (o + d t)| for t <= tm (two f32 roundings per component of magnitude <= |o| + t).
DE_DEV float pos_noise(float3 o, float tm) { return 3e-7f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + tm); }
DE_DEV float cloud_pass_setup(const DevScene &s, float3 o, float3 d, float &ts, float &tm) {
    const float cmax = cloud_segment_cmax(s, o, d, ts, tm);
    float bound = cloud_density_bound(cmax);
    if (bound > 0.0f && cmax < 0.99f) {
        // The top sphere is intersected from the pass's entry point (|p| ~ 6.4e6 m): from a far camera (Apollo: |o| = 5.7e7 m,
        // o.o ~ 3e15 with an ulp of 2.7e8) the discriminant of rsi(o, ...) is off by tens of metres of radius.  Tracked positions
        // are o + d t in f32, i.e. up to pos_noise() away from the ideal ray, so the margin grows with the camera distance.
        const float margin = 1e-3f + pos_noise(o, tm) * (1.0f / kCloudsThickness);
        const float2 top = rsi(o + d * ts, d, kCloudsLower + kCloudsThickness * (0.2f + 0.8f * cmax + margin));
        const float t0 = ts;
        ts = fmaxf(ts, t0 + top.x); tm = fminf(tm, t0 + top.y);
        if (!(top.y >= 0.0f && ts < tm)) bound = 0.0f;
    }
    return bound;
}
// Exact miss test for intersect_land (pathtracer.py:27-46), product flavour.  From the marching start
// point p (distance s0 already travelled) the terrain SDF is >= alt - scale (heightmap <= 1).  The loop
// only stops early when |dist| < 1e-4 * ray_dist.  If the lowest altitude the ray can still reach
// (perigee if it is approaching, the start point if it is receding) clears the tallest terrain by more
// than 1e-4 * (distance travelled until then) -- afterwards altitude grows like x^2/2r, which beats
// 1e-4 x by construction -- every iterate keeps dist > threshold, so the reference's stopping test can
// never fire: its loop runs ray_dist past 10 R and returns -1.  100 m of slack covers f32 cancellation
// for cameras at 5.7e7 m.  One caveat, a deviation on purpose: a ray that skims the
// tallest terrain so closely that the reference's march is still crawling after its 250 iterations is
// returned by the reference as a "hit" at wherever it got to (pathtracer.py:37,46); here it is a miss.
// Measured frequency: 2.5e-6 of isotropic rays started 0.1-12 km above the ground.
DE_DEV bool land_surely_missed(float3 p, float3 dir, float s0, float scale) {
    float b = dot(p, dir), r2 = dot(p, p);
    float rmin2 = b >= 0.0f ? r2 : r2 - b * b;
    float need = kPlanetR + scale + 1e-4f * (s0 + fmaxf(-b, 0.0f)) + 100.0f;
    return rmin2 > need * need;
}
// Same argument at an iterate of the march (s = distance travelled so far): above every terrain by more
// than the stopping tolerance can still reach, and receding => the reference's stopping test cannot fire
// any more (same iteration-cap caveat as above).
DE_DEV bool march_surely_missed(float3 ro, float3 dir, float r2, float s, float scale) {
    float need = kPlanetR + scale + 1e-4f * s + 100.0f;
    return r2 > need * need && dot(ro, dir) > 0.0f;
}
// Nothing can be hit above the terrain-top sphere R+scale: distance from p (s0 already travelled) to that
// sphere if p is outside and the ray enters it, else 0.  The reference's iterates from the atmosphere top
// stop at the first one within 1e-4*t of the surface, so WHERE inside that band the march stops depends on
// the iterate sequence; skipping the approach changes it.  The skip is therefore only taken when the band
// is at most 400 m (t <= 4e6 m: low orbits, secondary rays); far cameras (Apollo: band 5.7 km) keep the
// reference's iterates exactly.
DE_DEV float skip_to_terrain_top(float3 p, float3 dir, float s0, float scale) {
    const float Rg = kPlanetR + scale + 16.0f;
    float b = dot(p, dir), r = sqrtf(dot(p, p));
    if (r <= Rg || b >= 0.0f) return 0.0f;
    float disc = b * b - (r - Rg) * (r + Rg);
    if (!(disc > 0.0f)) return 0.0f;
    float skip = fmaxf(-b - sqrtf(disc), 0.0f);
    return s0 + skip <= 4.0e6f ? skip : 0.0f;
}
#endif
// ---------------------------------------------------------------- medium densities (volume_rendering_models.py:229-277)
DE_DEV float get_ozone_density(float h) {
    float h_km = h * 0.001f;
    float d2 = h_km - 25.0f;
    d2 = d2 * d2;
    float d = (1.0f - 0.375f) * expf(-d2 / 49.0f);
    d += 0.375f * expf(-d2 / 256.0f);
    d += fmaxf(0.0f, -0.000015f * pow_ti(h_km - 15.0f, 3.0f));
    return d;
}
DE_DEV float get_rayl_density(float h) {
    return 3.68082f * expf(-pow_ti(h + 24239.99f, 2.0f) / 532307548.4168f) / 1.225f;
}
DE_DEV float get_mie_density(float h) {
    float dens;
    if (h > 11500.0f) dens = 0.0918f * expf(-1.0e-6f * pow_ti(h - 11500.0f, 2.0f));
    else if (h > 2400.0f) dens = 0.3000f * expf(-2.5e-9f * pow_ti(h + 2500.00f, 2.0f)) - 0.092f;
    else if (h > 1300.0f) dens = 0.6500f * expf(-5.0e-6f * pow_ti(h - 1300.00f, 2.0f)) + 0.18899f;
    else dens = 1.0f - h / 8136.646f;
    return dens * 1.06f;
}
DE_DEV float3 get_density(float h) {
    h = fmaxf(h, 0.0f);
    return f3(get_rayl_density(h), get_mie_density(h), get_ozone_density(h));
}
DE_DEV float get_elevation(float3 p) { return sqrtf(p.x * p.x + p.y * p.y + p.z * p.z) - kPlanetR; }

#if !DE_EXACT
// Majorant of sigma.rho over the part [ts, tm] of a ray (product flavour).  The Rayleigh and aerosol
// fits decrease with altitude (the aerosol fit steps up by 1.3e-5 at 11.5 km: 1.001 covers it) and the
// ozone fit is bounded by its 25 km peak value 1 below the peak and decreases above it, so the densities
// at the LOWEST point of the segment bound the whole segment.  The reference uses the sea-level values
// everywhere (pathtracer.py:336,355); any valid majorant leaves delta / ratio tracking unbiased.
DE_DEV float rmo_segment_majorant(float3 ext, float3 o, float3 d, float ts, float tm) {
    // perigee from the segment's entry point p (|p| <= 6.5e6 m, r^2 resolves 0.3 m): evaluated from a far camera
    // (|o| = 5.7e7 m) the cancellation in o.o + t (2 o.d + t) left hmin up to ~70 m too high
    const float3 p = o + d * ts;
    const float b = dot(p, d), r2 = dot(p, p);
    const float tp = fminf(fmaxf(-b, 0.0f), tm - ts);       // perigee clamped to the segment
    const float slack = 2.0f + pos_noise(o, tm);            // tracked positions are fl(o + d t): up to pos_noise() below the ideal ray
    float hmin = fmaxf(sqrtf(fmaxf(r2 + tp * (2.0f * b + tp), 0.0f)) - kPlanetR - slack, 0.0f);
    float oz = hmin < 25000.0f ? 1.0f : get_ozone_density(hmin);
    return 1.001f * (ext.x * get_rayl_density(hmin) + ext.y * get_mie_density(hmin)) + ext.z * oz;
}
#endif
#if !DE_EXACT
// Altitude bands for the rmo majorant (product flavour).  Along a straight ray the altitude is convex in t, and every density fit
// falls with altitude (ozone: rises to its 25 km peak, then falls), so inside the band [H_k, H_k+1) the densities at the band's
// bottom (ozone: its maximum over the band) bound sigma.rho.  A tracking pass walks the bands: the free flight is sampled against
// the current band's majorant; a flight that would leave the band is cut at the band's exit sphere and restarted there against the
// next band's majorant (the exponential is memoryless, so the collision law is unchanged).  From space to the ground the reference's
// sea-level majorant spends ~110 km-equivalents of candidates, the bands {0, 4, 12, 30 km} about 15.
DE_DEV float rmo_band_majorant(const DevScene &s, float3 ext, int k) { return ext.x * s.band_dr[k] + ext.y * s.band_dm[k] + ext.z * s.band_do[k]; }
DE_DEV int rmo_band_of(const DevScene &s, float r) {
    int k = 0;
#pragma unroll
    for (int j = 1; j < kDeRmoBands; ++j) k += r >= s.band_r[j] ? 1 : 0;
    return k;
}
// Distance from q (a point of the ray inside band k, |q| ~ 6.4e6 m so the quadratic resolves < 1 m) along d to where the ray leaves the
// band.  Descending rays leave through the bottom if they reach it, everything else through the top; the top band has no exit (the pass
// ends at t_max first).
DE_DEV float rmo_band_exit(const DevScene &s, float3 q, float3 d, int k) {
    const float b = dot(q, d), r2 = dot(q, q);
    if (b < 0.0f && k > 0) {
        const float rl = s.band_r[k];
        const float disc = b * b - (r2 - rl * rl);
        if (disc > 0.0f) return fmaxf(-b - sqrtf(disc), 0.0f);
    }
    if (k == kDeRmoBands - 1) return 3.0e38f;
    const float rh = s.band_r[k + 1];
    return fmaxf(-b + sqrtf(fmaxf(b * b - (r2 - rh * rh), 0.0f)), 0.0f);
}
// At the exit point q of band k: the band the ray enters (a sphere about the centre is left inwards only while descending, outwards only
// while ascending) and the distance to THAT band's exit.
DE_DEV float rmo_band_cross(const DevScene &s, float3 q, float3 d, int k, int &k_new) {
    k_new = min(max(k + (dot(q, d) < 0.0f ? -1 : 1), 0), kDeRmoBands - 1);
    return rmo_band_exit(s, q, d, k_new);
}
// State of a pass walking the bands: the ray parameter t, the band it is in, where it leaves it, and the majorant in force
// (never above the whole segment's majorant m_seg).
struct RmoWalk { float t, tlim, max_ext; int band; };
// Advance by the optical depth tau (sampled against the majorants in force): returns the ray parameter reached, cutting the flight at
// every band exit before t_max and continuing with what is left of tau.  `advance(t, band)` -> (t of the NEW band's exit, new band), called
// at the exit point t of `band`, is a functor so the kernel can keep that rarely taken code out of line.
template <class Adv> DE_DEV float rmo_band_walk(const DevScene &s, float3 ext, float m_seg, float t_max, float tau, RmoWalk &w, Adv advance) {
    float tn = w.t + tau / w.max_ext;
#pragma unroll 1
    for (int it = 0; it < 2 * kDeRmoBands + 2 && tn >= w.tlim && w.tlim < t_max; ++it) {
        tau -= (w.tlim - w.t) * w.max_ext;
        w.t = w.tlim;
        const float2 adv = advance(w.t, w.band);
        w.band = __float_as_int(adv.y); w.tlim = adv.x;
        w.max_ext = fminf(m_seg, rmo_band_majorant(s, ext, w.band));
        tn = w.t + fmaxf(tau, 0.0f) / w.max_ext;
    }
    if (tn >= w.tlim && w.tlim < t_max) {  // (never in exact arithmetic: more crossings than bands) the segment's majorant is always valid
        w.tlim = 3.0e38f; w.max_ext = m_seg;
        tn = w.t + fmaxf(tau, 0.0f) / w.max_ext;
    }
    w.t = tn;
    return tn;
}
#endif
// ---------------------------------------------------------------- spectra (volume_rendering_models.py:48-51,194-224; colour.py:51-60)
DE_DEV float air(float wl) {
    float rcp = 1.0f / (wl * wl);
    return (float)(1.0 + 8.06051e-5) + 2.480990e-2f / (132.274f - rcp) + 1.74557e-4f / (39.32957f - rcp);
}
DE_DEV float spectra_extinction_mie(float wl) {
    const float c = (float)((0.6544 * 1.06 - 0.6510) * 4e-18);
    float K = (0.773335f - 0.00386891f * wl) / (1.0f - 0.00546759f * wl);
    return 0.434f * c * kPi * pow_ti(kTwoPi / (wl * 1e-9f), 4.0f - 2.0f) * K;
}
DE_DEV float spectra_extinction_rayleigh(float wl) {
    float wn = wl * 1e-9f;
    float F_N2 = 1.034f + 3.17e-4f * (1.0f / pow_ti(wl, 2.0f));
    float F_O2 = 1.096f + 1.385e-3f * (1.0f / pow_ti(wl, 2.0f)) + 1.448e-4f * (1.0f / pow_ti(wl, 4.0f));
    float CCO2 = 0.0421f;
    float king = (78.084f * F_N2 + 20.946f * F_O2 + 0.934f + CCO2 * 1.15f) / ((float)(78.084 + 20.946 + 0.934) + CCO2);
    float n = sqr(air(wl * 1e-3f)) - 1.0f;
    return (((float)(8.0 * 31.006276680299816) * pow_ti(n, 2.0f)) / ((float)(3.0 * 2.5035422e25) * pow_ti(wn, 4.0f))) * king;
}
DE_DEV float spectra_extinction_ozone(float wl, const float *o3) {
    float ext = 0.0f;
    if (wl >= 390.0f && wl < 831.0f) ext = (float)(0.0001 * (2.5035422e25 * 0.012588 * 8e-6)) * __ldg(o3 + (int)(wl - 390.0f));
    return ext;
}
DE_DEV float plancks(float T, float wl) {
    float h = 6.62607015e-16f, c = 2.9e17f, k = 1.38e-5f;
    float p1 = 2.0f * h * pow_ti(c, 2.0f) / pow_ti(wl, 5.0f);
    float p2 = expf((h * c) / (wl * k * T)) - 1.0f;
    return p1 / p2;
}
DE_DEV float cone_angle_to_solid_angle(float x) { return kTwoPi * (1.0f - cosf(x)); }  // math_utils.py:13

// ---------------------------------------------------------------- phase functions (volume_rendering_models.py:61-183)
DE_DEV float rayleigh_phase(float c) { return (float)(3.0 / (16.0 * 3.141592653589793)) * (1.0f + c * c); }
DE_DEV float klein_nishina_phase(float c, float e) { return e / (kTwoPi * (e * (1.0f - c) + 1.0f) * logf(2.0f * e + 1.0f)); }
DE_DEV float hg_phase(float c, float g) { return (1 - g * g) / (kFourPi * pow_ti(1.0f + g * g - 2 * g * c, 1.5f)); }
DE_DEV float draine_phase(float c, float g, float a) {
    return ((1 - g * g) * (1 + a * c * c)) / (4.f * (1 + (a * (1 + 2 * g * g)) / 3.f) * kPi * pow_ti(1 + g * g - 2 * g * c, 1.5f));
}
struct CloudPar { float g_hg, g_draine, alpha_draine, w_draine; };
DE_DEV CloudPar cloud_params(bool reduce_peak) {
#if !DE_EXACT
    {   // droplet size is the constant 8 (volume_rendering_models.py:155): fold the four exps (product flavour)
        CloudPar q;
        q.g_hg = reduce_peak ? 0.91f : 0.98446935f;
        q.g_draine = 0.54106367f; q.alpha_draine = 20.325685f; q.w_draine = 0.47364232f;
        return q;
    }
#endif
    const float d = 8.0f;
    CloudPar p;
    p.g_hg = reduce_peak ? 0.91f : expf(-0.0990567f / (d - 1.67154f));
    p.g_draine = expf(-2.20679f / (d + 3.91029f) - 0.428934f);
    p.alpha_draine = expf(3.62489f - 8.29288f / (d + 5.52825f));
    p.w_draine = expf(-0.599085f / (d - 0.641583f) - 0.665888f);
    return p;
}
DE_DEV float cloud_phase(float c, bool reduce_peak) {
    CloudPar p = cloud_params(reduce_peak);
    return mixf(hg_phase(c, p.g_hg), draine_phase(c, p.g_draine, p.alpha_draine), p.w_draine);
}
// math_utils.py:55-69
DE_DEV void make_orthonormal_basis(float3 n, float3 &x, float3 &y) {
    float3 h = fabsf(n.y) > 0.9f ? f3(1.0f, 0.0f, 0.0f) : f3(0.0f, 1.0f, 0.0f);
    y = normalize(cross(n, h));
    x = cross(n, y);
}
DE_DEV float3 spherical_direction(float st, float ct, float phi, float3 x, float3 y, float3 z) {
    float s, c;
#if DE_EXACT
    s = sinf(phi); c = cosf(phi);
#else
    __sincosf(phi, &s, &c);
#endif
    return (st * c) * x + (st * s) * y + ct * z;
}
template <class R> DE_DEV float3 sample_hg_phase(float3 view, float g, R &rng) {
    float sqr_term = (1 - g * g) / (1 - g + 2 * g * rng.next());
    float ct = (1 + g * g - sqr_term * sqr_term) / (2 * g);
    float st = sqrtf(fmaxf(0.0f, 1 - ct * ct));
    float phi = kTwoPi * rng.next();
    float3 t, b;
    make_orthonormal_basis(view, t, b);
    return spherical_direction(st, ct, phi, t, b, view);
}
template <class R> DE_DEV float3 sample_klein_nishina_phase(float3 view, float e, R &rng) {
    float ct = (-pow_ti(2.0f * e + 1.0f, 1.0f - rng.next()) + e + 1.0f) / e;
    float st = sqrtf(fmaxf(0.0f, 1 - ct * ct));
    float phi = kTwoPi * rng.next();
    float3 t, b;
    make_orthonormal_basis(view, t, b);
    return spherical_direction(st, ct, phi, t, b, view);
}
// volume_rendering_models.py:125-150 -- Draine inverse CDF of "An Approximate Mie Scattering Function for Fog and Cloud Rendering";
// term order kept.  SPDX-FileCopyrightText: Copyright (c) <2023> NVIDIA CORPORATION & AFFILIATES. All rights reserved.
// SPDX-License-Identifier: MIT -- full notice in NOTICE.md (also covers draine_phase and cloud_params above).
template <class R> DE_DEV float3 sample_draine(float3 view, float g, float a, R &rng) {
    float xi = rng.next();
    float g2 = g * g, g3 = g * g2, g4 = g2 * g2, g6 = g2 * g4;
    float pgp1_2 = (1 + g2) * (1 + g2);
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
    float phi = kTwoPi * rng.next();
    float3 t, b;
    make_orthonormal_basis(view, t, b);
    return spherical_direction(st, ct, phi, t, b, view);
}
template <class R> DE_DEV float3 sample_cloud_phase(float3 view, bool reduce_peak, R &rng) {
    CloudPar p = cloud_params(reduce_peak);
    if (rng.next() < p.w_draine) return sample_draine(view, p.g_draine, p.alpha_draine, rng);
    return sample_hg_phase(view, p.g_hg, rng);
}
// sampling.py:41-44
DE_DEV float3 sample_sphere(float r0, float r1) {
    r0 *= kTwoPi; r1 = r1 * 2.0f - 1.0f;
    float s = sqrtf(1.0f - r1 * r1);
    return normalize(f3(sinf(r0) * s, cosf(r0) * s, r1));
}
// sampling.py:13-28
template <class R> DE_DEV float3 sample_cone_oriented(float cmax, float3 n, R &rng) {
    float3 x, y;
    make_orthonormal_basis(n, x, y);
    float u0 = rng.next(), u1 = rng.next();
    float ct = (1.0f - u0) + u0 * cmax;
    float st = sqrtf(1.0f - ct * ct);
    float phi = kTwoPi * u1;
    float3 s = f3(st * cosf(phi), st * sinf(phi), ct);
    return f3(x.x * s.x + y.x * s.y + n.x * s.z, x.y * s.x + y.y * s.y + n.y * s.z, x.z * s.x + y.z * s.y + n.z * s.z);
}
// sampling.py:30-39
template <class R> DE_DEV float3 sample_hemisphere_cosine_weighted(float3 n, R &rng) {
    float u0 = rng.next(), u1 = rng.next();
    float a = 1.0f - 2.0f * u0;
    float b = sqrtf(1.0f - a * a);
    a *= (float)(1.0 - 1e-5);
    b *= (float)(1.0 - 1e-5);
    float phi = kTwoPi * u1;
    return normalize(f3(n.x + b * cosf(phi), n.y + b * sinf(phi), n.z + a));
}
// pathtracer.py:235-247
DE_DEV float evaluate_phase(float3 ray_dir, float3 light_dir, int id, bool reduce_peak) {
    float phase = 0.0f, c = dot(ray_dir, light_dir);
    if (id == kRayleigh) phase = rayleigh_phase(c);
    else if (id == kMie) phase = klein_nishina_phase(c, kMieAsymmetry);
    else if (id == kCloud) phase = cloud_phase(c, reduce_peak);
    else if (id == kIsoCloud) phase = (float)(1.0 / (4.0 * 3.141592653589793));
    return phase;
}
// pathtracer.py:249-261
template <class R> DE_DEV float3 sample_phase(float3 ray_dir, int id, bool reduce_peak, R &rng, float &phase_div_pdf) {
    phase_div_pdf = 1.0f;
    if (id == kRayleigh || id == kIsoCloud) {
        float r0 = rng.next(), r1 = rng.next();
        float3 d = sample_sphere(r0, r1);
        phase_div_pdf = evaluate_phase(ray_dir, d, id, reduce_peak) * kFourPi;
        return d;
    }
    if (id == kMie) return sample_klein_nishina_phase(ray_dir, kMieAsymmetry, rng);
    return sample_cloud_phase(ray_dir, reduce_peak, rng);
}
// pathtracer.py:263-270
template <class R> DE_DEV bool sample_scatter_event(int id, R &rng) {
    if (id == kIsoCloud) id = kCloud;
    float albedo = id == kRayleigh ? 1.0f : (id == kMie ? 0.95f : (id == kOzone ? 0.0f : 0.99f));
    return rng.next() < albedo;
}

// ---------------------------------------------------------------- ground BRDF (surface_rendering_models.py)
DE_DEV float disney_diffuse(float rough, float nl, float nv, float lh) {
    float R_R = 2.0f * rough * sqr(lh);
    float F_L = pow_ti(1.0f - nl, 5.0f), F_V = pow_ti(1.0f - nv, 5.0f);
    const float f_lambert = (float)(1.0 / 3.141592653589793);
    float f_retro = f_lambert * R_R * (F_L + F_V + F_L * F_V * (R_R - 1.0f));
    return f_lambert * (1.0f - 0.5f * F_L) * (1.0f - 0.5f * F_V) + f_retro;
}
DE_DEV float fresnel_dielectric(float vh, float F0) {
    F0 = sqrtf(F0);
    F0 = (1.0f + F0) / (1.0f - F0);
    float sI = sqrtf(saturate(1.0f - sqr(vh)));
    float sT = sI / fmaxf(F0, 1e-8f);
    float cT = sqrtf(1.0f - sqr(sT));
    float Rs = sqr((vh - (F0 * cT)) / fmaxf(vh + (F0 * cT), 1e-8f));
    float Rp = sqr((cT - (F0 * vh)) / fmaxf(cT + (F0 * vh), 1e-8f));
    return saturate((Rs + Rp) * 0.5f);
}
DE_DEV float GGX_D(float nh, float a2) {
    float den = (a2 - 1.0f) * nh * nh + 1.0f;
    return a2 / (kPi * den * den);
}
DE_DEV float lambda_smith(float nx, float a2) {
    float x2 = nx * nx;
    return (-1.0f + sqrtf(a2 * (1.0f - x2) / x2 + 1.0f)) * 0.5f;
}
DE_DEV float G2_smith(float nl, float nv, float a2) {
    float lv = lambda_smith(nv, a2), ll = lambda_smith(nl, a2);
    return 1.0f / (1.0f + lv + ll);
}
DE_DEV float GGX_smith_specular(float rough, float F0, float nl, float nv, float lh, float nh) {
    float a2 = rough * rough;
    float D = GGX_D(nh, a2), G = G2_smith(nl, nv, a2), F = fresnel_dielectric(lh, F0);
    return D * G * F / fmaxf(4.0f * nl * nv, 1e-5f);
}
DE_DEV float beckmann_isotropic_ndf(float nh, float alpha) {
    float c2 = nh * nh, a2 = alpha * alpha;
    float exponent = (1.0f - c2) / (a2 * c2);
    float denom = kPi * a2 * c2 * c2;
    return expf(-exponent) / fmaxf(denom, 1e-5f);
}
DE_DEV float G2_VCavity(float nl, float nv, float nh, float vh) {
    return fminf(1.0f, fminf(2.0f * nv * nh / vh, 2.0f * nl * nh / vh));
}
DE_DEV float beckmann_specular(float rough, float F0, float nl, float nv, float lh, float nh) {
    float alpha = rough;
    alpha *= alpha * 2.0f;
    return beckmann_isotropic_ndf(nh, alpha) * G2_VCavity(nl, nv, nh, lh) * fresnel_dielectric(lh, F0);
}
// surface_rendering_models.py:9-37
DE_DEV float earth_brdf(float albedo, float oceanness, float bathymetry, float3 v, float3 n, float3 l, float &n_dot_l) {
    float3 h = normalize(v + l);
    float nl = saturate(dot(n, l)), nv = saturate(dot(n, v));
    float lh = saturate(dot(l, h)), nh = saturate(dot(n, h));
    const float land_roughness = 0.73f;
    float ocean_roughness = mixf((float)(0.23 + 0.02), (float)(0.23 - 0.04), smoothstep(0.3f, 0.7f, bathymetry));
    float diffuse = disney_diffuse(land_roughness, nl, nv, lh);
    float land_spec = GGX_smith_specular(land_roughness, 0.04f, nl, nv, lh, nh);
    float ocean_ggx = GGX_smith_specular(ocean_roughness, 0.02f, nl, nv, lh, nh);
    float ocean_beck = 0.65f * beckmann_specular(ocean_roughness, 0.02f, nl, nv, lh, nh);
    float ocean_spec = mixf(ocean_beck, ocean_ggx, clampf(smoothstep(0.2f, 0.95f, nv), 0.05f, 0.94f));
    float blender = smoothstep(0.6f, 1.0f, oceanness);
    n_dot_l = nl;
    return albedo * diffuse * 0.28f + mixf(land_spec, ocean_spec, blender) * 0.5f;
}
// colour.py:88-95
DE_DEV float lum(float3 x) { return dot(x, f3(0.2126729f, 0.7151522f, 0.0721750f)); }
DE_DEV float3 lum3(float3 x) { float y = lum(x); return f3(y, y, y); }
// colour.py:62-71; the sign of f (<= 0, extrapolating) is the reference's
DE_DEV void srgb_to_spectrum_coeff(const float *lut, float wl, float3 &coeff, bool &valid) {
    int w = (int)(wl - 400.0f);
    float f = (float)w - (wl - 400.0f);
    valid = w > 0 && w < 299;
    coeff = f3(0, 0, 0);
    if (valid) {
        float3 a = f3(__ldg(lut + w * 3), __ldg(lut + w * 3 + 1), __ldg(lut + w * 3 + 2));
        float3 b = f3(__ldg(lut + w * 3 + 3), __ldg(lut + w * 3 + 4), __ldg(lut + w * 3 + 5));
        coeff = mix3(a, b, f);
    }
}
DE_DEV float srgb_to_spectrum(const float *lut, float3 rgb, float wl) {
    float3 c; bool valid;
    srgb_to_spectrum_coeff(lut, wl, c, valid);
    return valid ? dot(rgb, c) : 0.0f;
}
struct LandMaterial { float3 albedo_srgb; float ocean, bathymetry, emissive; };
// pathtracer.py:284-313: colour grading of the albedo texel
DE_DEV float3 grade_albedo(float3 tex, float ocean) {
    float3 land = mix3(lum3(tex), tex, 6.5f);
    float greenery = pow_ti(land.y / lum(land), 2.0f);
    greenery = smoothstep(1.5f, 1.9f, greenery);
    land = (1.0f * tex) / (greenery * 0.7f + 1.0f);
    land = mix3(lum3(land), land, 1.4f - greenery * 0.45f);
    land = mix3(land, (land * f3(255.0f, 128.0f, 64.0f)) / 255.0f, 0.2f * (1.0f - greenery));
    float3 ocean_albedo = mix3(lum3(tex), tex, 0.75f) * 0.9f;
    return mix3(land, ocean_albedo, ocean);
}

// ---------------------------------------------------------------- colour / camera
// colour.py:12-48 evaluated literally (8 LUT-bisection steps)
DE_DEV void spectrum_sample(const float *cie, float sample, float &wavelength, float3 &response, float &rcp_pdf, float &mid_out) {
    float lo = 0.0f, hi = 1.0f, mid = (lo + hi) / 2.0f;
    const float third = (float)(1.0 / 3.0);
    for (int x = 0; x < 8; ++x) {
        float r = tex_f32(cie, 441, 2, 3, 0, mid, 0.25f), g = tex_f32(cie, 441, 2, 3, 1, mid, 0.25f), b = tex_f32(cie, 441, 2, 3, 2, mid, 0.25f);
        float val = saturate((third * r + third * g) + third * b);
        if (val < sample) lo = mid;
        else if (val > sample) hi = mid;
        else break;
        mid = (lo + hi) / 2.0f;
    }
    wavelength = 390.0f + 441.0f * mid;
    response = f3(tex_f32(cie, 441, 2, 3, 0, mid, 0.75f), tex_f32(cie, 441, 2, 3, 1, mid, 0.75f), tex_f32(cie, 441, 2, 3, 2, mid, 0.75f));
    float3 mx = f3(tex_f32(cie, 441, 2, 3, 0, 1.0f, 0.25f), tex_f32(cie, 441, 2, 3, 1, 1.0f, 0.25f), tex_f32(cie, 441, 2, 3, 2, 1.0f, 0.25f));
    float pdf = dot(response, mx);
    rcp_pdf = (pdf > 1e-3f && !(isinf(pdf) || isnan(pdf))) ? 1.0f / pdf : 0.0f;
    mid_out = mid;
}
// Same bisection on the precomputed thresholds cdf[j] = saturate(mean CIE CDF(j/512)); returns j (mid = j/512).
DE_DEV int spectrum_bin(const float *cdf, float sample) {
    int lo = 0, hi = 512, mid = 256;
#pragma unroll 1
    for (int x = 0; x < 8; ++x) {
        float val = cdf[mid];
        if (val < sample) lo = mid;
        else if (val > sample) hi = mid;
        else break;
        mid = (lo + hi) >> 1;
    }
    return mid;
}
// renderer.py:269-279 given the precomputed basis
DE_DEV float3 get_cast_dir(const DevScene &s, const DevDerived &dv, float u, float v, float xi_u, float xi_v) {
    float fov = s.fov;
    float fu = (2 * fov * (u + xi_u) / (float)s.H - fov * s.aspect_ratio - 1e-5f) * s.aspect_scale;
    float fv = 2 * fov * (v + xi_v) / (float)s.H - fov - 1e-5f;
    return normalize((dv.cam_d + fu * dv.cam_du) + fv * dv.cam_dv);
}
DE_DEV float3 xyz_to_rgb(float3 xyz) {  // colour.py:6-10
    const float kXYZ2RGB[9] = {(float)3.2409699419, (float)-1.5373831776, (float)-0.4986107603, (float)-0.9692436363, (float)1.8759675015,
                               (float)0.0415550574, (float)0.0556300797, (float)-0.2039769589, (float)1.0569715142};
    return f3((kXYZ2RGB[0] * xyz.x + kXYZ2RGB[1] * xyz.y) + kXYZ2RGB[2] * xyz.z, (kXYZ2RGB[3] * xyz.x + kXYZ2RGB[4] * xyz.y) + kXYZ2RGB[5] * xyz.z,
              (kXYZ2RGB[6] * xyz.x + kXYZ2RGB[7] * xyz.y) + kXYZ2RGB[8] * xyz.z);
}

// ---------------------------------------------------------------- tone mapping (OpenDRT.py, AgX.py, renderer.py:333-365)
// openDR_transform and its helpers follow OpenDRT v0.2.2 by Jed Smith (https://github.com/jedypod/open-display-transform) as ported
// in the reference's lib/OpenDRT.py -- License: GPL v3; the AgX functions follow Troy Sobotka's AgX (shader translation by Olivier
// Groulx).  See NOTICE.md.
DE_DEV float sdivf(float a, float b) { return fabsf(b) < 1e-4f ? 0.0f : a / b; }
DE_DEV float spowf(float a, float b) { return a <= 0.0f ? a : pow_ti(a, b); }
DE_DEV float logf10_ti(float x) { return log2_ti(x) / log2_ti(10.0f); }
DE_DEV float flare_inv(float x, float fl) { return (x + sqrtf(x * (4.0f * fl + x))) / 2.0f; }
DE_DEV float3 vdot(const float *m, float3 v) {  // v @ m (OpenDRT.py:86-88)
    return f3((v.x * m[0] + v.y * m[3]) + v.z * m[6], (v.x * m[1] + v.y * m[4]) + v.z * m[7], (v.x * m[2] + v.y * m[5]) + v.z * m[8]);
}
DE_DEV float3 narrow_hue_angles(float3 v) {
    return f3(fminf(2.0f, fmaxf(0.0f, v.x - (v.y + v.z))), fminf(2.0f, fmaxf(0.0f, v.y - (v.x + v.z))), fminf(2.0f, fmaxf(0.0f, v.z - (v.x + v.y))));
}
struct OpenDrtPar { float s, m, ds, clamp_max; };
// OpenDRT.py:270-319 with Lp=100, gb=.12, c=1, fl=.005 (OpenDRT.py:44-48)
DE_DEV OpenDrtPar opendrt_params() {
    const float Lp = 100.0f, gb = 0.12f, fl = 0.005f;
    OpenDrtPar p;
    p.ds = 1.0f;
    p.clamp_max = p.ds * Lp / 100.0f;
    float px = 128.0f * logf10_ti(Lp) / logf10_ti(100.0f) - 64.0f;
    float py = 1.0f;
    float gx = 0.18f;
    float gy = (float)(11.696 / 100.0) * (1.0f + gb * logf10_ti(py) / logf10_ti(2.0f));
    float s0 = flare_inv(gy, fl), m0 = flare_inv(py, fl);
    float ip = 1.0f;
    p.s = (px * gx * (pow_ti(m0, ip) - pow_ti(s0, ip))) / (px * pow_ti(s0, ip) - gx * pow_ti(m0, ip));
    p.m = pow_ti(m0, ip) * (p.s + px) / px;
    return p;
}
DE_DEV float3 openDR_transform(float3 in, const OpenDrtPar &P) {
    const float rec709_to_xyz[9] = {(float)0.412390917540, (float)0.357584357262, (float)0.180480793118, (float)0.212639078498, (float)0.715168714523,
                                    (float)0.072192311287, (float)0.019330825657, (float)0.119194783270, (float)0.950532138348};
    const float xyz_to_rec709[9] = {(float)3.2409699419, (float)-1.53738317757, (float)-0.498610760293, (float)-0.969243636281, (float)1.87596750151,
                                    (float)0.041555057407, (float)0.055630079697, (float)-0.203976958889, (float)1.05697151424};
    const float c = 1.0f, fl = 0.005f, rw = 0.25f, bw = 0.35f, dch = 0.35f, dch_toe = 0.0f, hs_r = 0.3f, hs_g = -0.1f, hs_b = -0.2f, v_p = 0.5f;
    float3 rgb = vdot(xyz_to_rec709, vdot(rec709_to_xyz, in));
    float mx = fmaxf(rgb.x, fmaxf(rgb.y, rgb.z)), mn = fminf(rgb.x, fminf(rgb.y, rgb.z));
    float3 h_rgb = narrow_hue_angles(f3(sdivf(rgb.x - mn, mx), sdivf(rgb.y - mn, mx), sdivf(rgb.z - mn, mx)));
    float3 w = f3(rw, 1.0f, bw);
    w = w / length(w);
    w = w * f3(fmaxf(rgb.x, 1e-5f), fmaxf(rgb.y, 1e-5f), fmaxf(rgb.z, 1e-5f));
    float lumv = length(w);
    float3 rats = f3(sdivf(rgb.x, lumv), sdivf(rgb.y, lumv), sdivf(rgb.z, lumv));
    float ts = spowf(P.m * lumv / (lumv + P.s), c);
    ts = spowf(ts, 2.0f) / (ts + fl);
    ts *= P.ds;
    float dch_s = dch / P.s;
    float ccf = sdivf(1.0f, lumv * dch_s + 1.0f);
    float toe_ccf = 1.0f * sdivf(lumv, lumv + dch_toe) * ccf;
    float3 hs_w = (1.0f - ccf) * h_rgb;
    rats = f3(rats.x + hs_w.z * hs_b - hs_w.y * hs_g, rats.y + hs_w.x * hs_r - hs_w.z * hs_b, rats.z + hs_w.y * hs_g - hs_w.x * hs_r);
    float omt = 1.0f - toe_ccf;
    rats = f3(omt + rats.x * toe_ccf, omt + rats.y * toe_ccf, omt + rats.z * toe_ccf);
    rats = f3(fmaxf(rats.x, 0.0f), fmaxf(rats.y, 0.0f), fmaxf(rats.z, 0.0f));
    float rmx = fmaxf(rats.x, fmaxf(rats.y, rats.z)), rmn = fminf(rats.x, fminf(rats.y, rats.z));
    float rats_ch = sdivf(rmx - rmn, rmx);
    float chf = spowf(rats_ch * ts, v_p);
    float3 rn = f3(sdivf(rats.x, rmx), sdivf(rats.y, rmx), sdivf(rats.z, rmx));
    rats = rn * chf + rats * (1.0f - chf);
    rgb = rats * ts;
    return f3(fminf(rgb.x, P.clamp_max), fminf(rgb.y, P.clamp_max), fminf(rgb.z, P.clamp_max));
}
// ---- AgX (AgX.py) ----
struct M33 { float m[3][3]; };
DE_DEV float3 mul(const M33 &a, float3 v) {
    return f3((a.m[0][0] * v.x + a.m[0][1] * v.y) + a.m[0][2] * v.z, (a.m[1][0] * v.x + a.m[1][1] * v.y) + a.m[1][2] * v.z,
              (a.m[2][0] * v.x + a.m[2][1] * v.y) + a.m[2][2] * v.z);
}
DE_DEV M33 InverseMat(const M33 &q) {
    const float(*m)[3] = q.m;
    float d = m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    float id = 1.0f / d;
    M33 c;
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
DE_DEV float3 Unproject(float2 xy) {
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    if (xy.y != 0.0f) { Y = 1.0f; X = (xy.x * Y) / xy.y; Z = ((1.0f - xy.x - xy.y) * Y) / xy.y; }
    return f3(X, Y, Z);
}
DE_DEV M33 PrimariesToMatrix(float2 r, float2 g, float2 b, float2 w) {
    float3 R = Unproject(r), G = Unproject(g), B = Unproject(b), W = Unproject(w);
    M33 t = {{{R.x, G.x, B.x}, {1.0f, 1.0f, 1.0f}, {R.z, G.z, B.z}}};
    float3 sc = mul(InverseMat(t), W);
    M33 o = {{{sc.x * R.x, sc.y * G.x, sc.z * B.x}, {sc.x * R.y, sc.y * G.y, sc.z * B.y}, {sc.x * R.z, sc.y * G.z, sc.z * B.z}}};
    return o;
}
DE_DEV float AgXScale(float xp, float yp, float sp, float power) {
    return pow_ti(pow_ti(sp * xp, -power) * (pow_ti(sp * (xp / yp), power) - 1.0f), -1.0f / power);
}
DE_DEV float AgXHyperbolic(float x, float power) { return x / pow_ti(1.0f + pow_ti(x, power), 1.0f / power); }
DE_DEV float AgXFullCurve(float x, float xp, float yp, float sp, float toe, float shoulder) {
    float sxp = x >= xp ? 1.0f - xp : xp, syp = x >= xp ? 1.0f - yp : yp;
    float toe_scale = AgXScale(sxp, syp, sp, toe), shoulder_scale = AgXScale(sxp, syp, sp, shoulder);
    float scale = x >= xp ? shoulder_scale : -toe_scale;
    float term = (sp * (x - xp)) / scale;
    return scale * AgXHyperbolic(term, scale < 0.0f ? toe : shoulder) + yp;
}
struct AgxPar { M33 srgb_to_xyz, xyz_to_adjusted; };
DE_DEV AgxPar agx_params() {
    float2 pr = make_float2(0.64f, 0.33f), pg = make_float2(0.3f, 0.6f), pb = make_float2(0.15f, 0.06f), pw = make_float2(0.3127f, 0.3290f);
    AgxPar p;
    p.srgb_to_xyz = PrimariesToMatrix(pr, pg, pb, pw);
    float sf = 1.0f / (1.0f - 0.15f);
    float2 R = make_float2((pr.x - pw.x) * sf + pw.x, (pr.y - pw.y) * sf + pw.y);
    float2 G = make_float2((pg.x - pw.x) * sf + pw.x, (pg.y - pw.y) * sf + pw.y);
    float2 B = make_float2((pb.x - pw.x) * sf + pw.x, (pb.y - pw.y) * sf + pw.y);
    p.xyz_to_adjusted = InverseMat(PrimariesToMatrix(R, G, B, pw));
    return p;
}
DE_DEV float3 agx_display_transform(float3 col, const AgxPar &P) {
    const float MIDDLE_GREY = 0.18f, SLOPE = 2.3f, TOE = 1.9f, SHOULDER = 3.1f, MIN_EV = -10.0f, MAX_EV = 6.5f, SATURATION = 1.4f;
    float3 adj = mul(P.xyz_to_adjusted, mul(P.srgb_to_xyz, col));
    const float x_pivot = (float)(10.0 / (6.5 - -10.0)), y_pivot = 0.5f;
    float total = MAX_EV - MIN_EV;
    float3 lg = f3(clampf(log2_ti(adj.x / MIDDLE_GREY), MIN_EV, MAX_EV), clampf(log2_ti(adj.y / MIDDLE_GREY), MIN_EV, MAX_EV), clampf(log2_ti(adj.z / MIDDLE_GREY), MIN_EV, MAX_EV));
    lg = f3((lg.x - MIN_EV) / total, (lg.y - MIN_EV) / total, (lg.z - MIN_EV) / total);
    float3 o = f3(AgXFullCurve(lg.x, x_pivot, y_pivot, SLOPE, TOE, SHOULDER), AgXFullCurve(lg.y, x_pivot, y_pivot, SLOPE, TOE, SHOULDER), AgXFullCurve(lg.z, x_pivot, y_pivot, SLOPE, TOE, SHOULDER));
    o = f3(saturate(o.x), saturate(o.y), saturate(o.z));
    o = mix3(lum3(o), o, SATURATION);
    return f3(saturate(o.x), saturate(o.y), saturate(o.z));
}
// renderer.py:333-344
DE_DEV float3 camera_response(const DevScene &s, float3 t) {
    t = f3(saturate(t.x), saturate(t.y), saturate(t.z));
    float slice_v = ((float)s.selected_crf + 0.5f) / (float)s.crf_count;
    const float u_off = (float)(0.5 / 1024.0);
    float3 ul = f3(fminf(t.x + u_off, 1.0f - u_off), fminf(t.y + u_off, 1.0f - u_off), fminf(t.z + u_off, 1.0f - u_off));
    float r = tex_f32(s.crf, 1024, s.n_crf, 3, 0, ul.x, slice_v);
    float g = tex_f32(s.crf, 1024, s.n_crf, 3, 1, ul.y, slice_v);
    float b = tex_f32(s.crf, 1024, s.n_crf, 3, 2, ul.z, slice_v);
    return f3(saturate(r), saturate(g), saturate(b));
}
// colour.py:74-79
DE_DEV float srgb_transfer1(float lin) {
    float lo = lin * 12.92f;
    float hi = (pow_ti(fabsf(lin), (float)(1.0 / 2.4)) * 1.055f) - 0.055f;
    float st = 0.0031308f >= lin ? 1.0f : 0.0f;
    return hi * (1.0f - st) + lo * st;
}
// renderer.py:346-365
DE_DEV float3 resolve_pixel(const DevScene &s, const OpenDrtPar &OP, const AgxPar &AP, int i, int j, float3 color, int samples) {
    float u = 1.0f * (float)i / (float)s.W, v = 1.0f * (float)j / (float)s.H;
    float du = u - s.vig_cx, dv = v - s.vig_cy;
    float darken = 1.0f - s.vig_strength * fmaxf(sqrtf(du * du + dv * dv) - s.vig_radius, 0.0f);
    float ex = pow_ti(2.0f, s.exposure), ns = (float)samples;
    float3 lin = f3(color.x / ns * darken * ex, color.y / ns * darken * ex, color.z / ns * darken * ex);
    float3 tm = s.tonemapper == 1 ? agx_display_transform(lin, AP) : openDR_transform(lin, OP);
    float3 cam = camera_response(s, tm);
    float3 g = f3(pow_ti(cam.x, s.gamma), pow_ti(cam.y, s.gamma), pow_ti(cam.z, s.gamma));
    return f3(srgb_transfer1(g.x), srgb_transfer1(g.y), srgb_transfer1(g.z));
}

}  // namespace DE_NS
