// de_wavefront.cu -- the product integrator: persistent-thread, stage-sorted wavefront.
//
// Why: ncu on the one-thread-per-pixel kernel (profiles/r1_megakernel.md) shows 4.9 of 32 lanes
// active per issued instruction -- path length and stage mix diverge, memory does not matter
// (L1 97 %, L2 98 % hits, DRAM idle).  This kernel keeps every lane of a warp inside the SAME
// inner loop:
//   * each warp owns a pool of WF_N path states in shared memory (SoA, ~112 B per path), so
//     state never touches HBM;
//   * a path is a small state machine  SDF -> RMO -> CLOUD -> EVENT -> (SDF) -> RMO -> CLOUD ->
//     NEE_DONE -> SDF ...  (pathtracer.py:349-453 cut at its loop boundaries);
//   * the warp repeatedly ballots the pool, picks the most populated stage and runs a burst of
//     that stage's loop body with all lanes converged; a lane whose path leaves the stage takes
//     over a spare pool member of the same stage (warp-ballot compaction);
//   * terminated paths are replaced at once from a global atomic work counter (path
//     regeneration), 32 consecutive pixels of one 16x8 film tile and one sample index at a time.
// The random stream of a path is the one the parity kernel uses (Philox key (seed,pixel), counter
// (sample,bounce,draw)), so a pixel's samples are the same paths in every integrator flavour.
#include "de_integrator.cuh"
#include "de_launch.h"
#include "de_wavefront.h"

namespace de_fast {

#ifndef WF_N
#define WF_N 96        // pool slots per warp
#endif
#ifndef WF_WARPS
#define WF_WARPS 16    // warps per CTA, one persistent CTA per SM
#endif
#ifndef WF_BURST
#define WF_BURST 64    // max loop iterations per burst
#endif
#ifndef WF_MIN_ACTIVE
#define WF_MIN_ACTIVE 16
#endif

// Stages.  Loop stages (SDF, RMO, CLOUD) run bursts of a small loop body; the others are one-shot
// bodies executed converged over up to 32 members.  A loop body never runs transition code: a
// finished lane records its result and flips the stage, the transition happens later for a whole
// group at once.
enum : uint32_t { ST_DEAD = 0, ST_NEW, ST_SDF, ST_RMO, ST_CLOUD, ST_SDF_DONE, ST_RMO_DONE, ST_EVENT, ST_NEE_DONE, ST_COUNT };

// pk word: stage[0:4) ratio[4] shadow[5] surface[6] vis[7] sc[8:13) lam[13:22) ev[22:24) rmo_ev[24:26) rmo_id[26:28) id[28:31)
#define PK_STAGE(p) ((p)&15u)
#define PK_RATIO 16u
#define PK_SHADOW 32u
#define PK_SURFACE 64u
#define PK_VIS 128u
#define PK_SC(p) (((p) >> 8) & 31u)
#define PK_LAM(p) (((p) >> 13) & 511u)
#define PK_EV(p) (((p) >> 22) & 3u)
#define PK_RMO_EV(p) (((p) >> 24) & 3u)
#define PK_RMO_ID(p) (((p) >> 26) & 3u)
#define PK_ID(p) (((p) >> 28) & 7u)
DE_DEV uint32_t pk_set(uint32_t p, int shift, uint32_t mask, uint32_t v) { return (p & ~(mask << shift)) | ((v & mask) << shift); }
#define PK_SET_STAGE(p, v) pk_set(p, 0, 15u, v)
#define PK_SET_SC(p, v) pk_set(p, 8, 31u, v)
#define PK_SET_LAM(p, v) pk_set(p, 13, 511u, v)
#define PK_SET_EV(p, v) pk_set(p, 22, 3u, v)
#define PK_SET_RMO_EV(p, v) pk_set(p, 24, 3u, v)
#define PK_SET_RMO_ID(p, v) pk_set(p, 26, 3u, v)
#define PK_SET_ID(p, v) pk_set(p, 28, 7u, v)

struct WarpPool {  // SoA: lane l touching slot s hits bank s%32
    float ox[WF_N], oy[WF_N], oz[WF_N], dx[WF_N], dy[WF_N], dz[WF_N];
    float thr[WF_N], L[WF_N];
    uint32_t pix[WF_N], sample[WF_N], pk[WF_N], draw[WF_N];  // draw: rng draw index [0:24) | sdf iteration [24:32)
    float t[WF_N], tmax[WF_N], aux[WF_N], isect[WF_N];       // aux: rmo_t (delta) or transmittance (ratio)
    float mdx[WF_N], mdy[WF_N], mdz[WF_N];                    // main ray direction while the NEE ray is tracked
    float nx[WF_N], ny[WF_N], nz[WF_N], m0[WF_N], m1[WF_N], m2[WF_N];  // surface normal, albedo, ocean, bathymetry
    float na[WF_N], nb[WF_N];                                 // NEE factors: phase | brdf, n.l
    uint8_t members[WF_N];
};

struct WfParams {
    float *accum;
    unsigned int *next;  // global work counter (units of 32 paths)
    unsigned int n_chunks;
    int n_spp, x0, y0, w, h, tiles_x;
    uint32_t seed, first_sample;
};

// One Philox4x32-10 block; deliberately NOT inlined: ~70 instructions that would otherwise be
// replicated at every draw site and blow the instruction cache (profiles/r1_wavefront.md).
__device__ __noinline__ uint4 philox_block(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2) {
    uint32_t c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// Philox stream that can be resumed from (bounce, draw) kept in the pool
struct RngW {
    uint32_t key0, key1, sample, bounce, draw;
    uint32_t b0, b1, b2, b3;
    bool valid;
    DE_DEV void refill() {
        uint4 b = philox_block(key0, key1, sample, bounce, draw >> 2);
        b0 = b.x; b1 = b.y; b2 = b.z; b3 = b.w;
        valid = true;
    }
    DE_DEV void align() { draw = (draw + 3u) & ~3u; valid = false; }
    DE_DEV void skip() { draw += 1u; valid = false; }
    DE_DEV float next() {
        uint32_t lane = draw & 3u;
        if (lane == 0u || !valid) refill();
        ++draw;
        uint32_t v = lane == 0u ? b0 : (lane == 1u ? b1 : (lane == 2u ? b2 : b3));
        return (float)(v >> 8) * (1.0f / 16777216.0f);
    }
};

// Out-of-line equirect fetches for the one-shot stages (normals, materials, stars); the loop
// stages keep their single fetch inline.
__device__ __noinline__ float fetch_r8_ool(const uint8_t *data, int w, int h, float px, float py, float pz) {
    DevTex t; t.data = data; t.w = w; t.h = h; t.c = 1; t.obj = 0;
    return sample_sphere_r8(t, f3(px, py, pz));
}
__device__ __noinline__ float3 fetch_rgb8_ool(const uint8_t *data, int w, int h, float px, float py, float pz) {
    DevTex t; t.data = data; t.w = w; t.h = h; t.c = 3; t.obj = 0;
    return sample_sphere_rgb8(t, f3(px, py, pz));
}
DE_DEV float r8_ool(const DevTex &t, float3 p) { return fetch_r8_ool(t.data, t.w, t.h, p.x, p.y, p.z); }
DE_DEV float3 rgb8_ool(const DevTex &t, float3 p) { return fetch_rgb8_ool(t.data, t.w, t.h, p.x, p.y, p.z); }

struct Ctx {  // per-warp context
    const DevScene &s;
    const DevDerived &dv;
    const WfParams &P;
    WarpPool &pool;
    Counters &cn;
    int lane;
};

DE_DEV RngW load_rng(const Ctx &c, int slot, uint32_t pk) {
    RngW r;
    r.key0 = c.P.seed; r.key1 = c.pool.pix[slot]; r.sample = c.pool.sample[slot];
    r.bounce = PK_SC(pk) + 1u; r.draw = c.pool.draw[slot] & 0xFFFFFFu; r.valid = false;
    return r;
}
DE_DEV void store_draw(const Ctx &c, int slot, uint32_t draw, uint32_t iter) { c.pool.draw[slot] = (draw & 0xFFFFFFu) | (iter << 24); }
DE_DEV float3 ld_o(const Ctx &c, int s) { return f3(c.pool.ox[s], c.pool.oy[s], c.pool.oz[s]); }
DE_DEV float3 ld_d(const Ctx &c, int s) { return f3(c.pool.dx[s], c.pool.dy[s], c.pool.dz[s]); }
DE_DEV void st_o(const Ctx &c, int s, float3 v) { c.pool.ox[s] = v.x; c.pool.oy[s] = v.y; c.pool.oz[s] = v.z; }
DE_DEV void st_d(const Ctx &c, int s, float3 v) { c.pool.dx[s] = v.x; c.pool.dy[s] = v.y; c.pool.dz[s] = v.z; }
DE_DEV float cloud_ext_of(uint32_t sc) { return sc > 9u ? 0.02f : kCloudsExtinct; }  // pathtracer.py:351-352

// ------------------------------------------------------------------ transitions (run converged inside one-shot stages)
// ratio tracking through the cloud shell (second half of sample_transmittance, pathtracer.py:229-231)
DE_DEV uint32_t setup_cloud_ratio(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d) {
    float ts, tm;
    intersect_cloud_limits(o, d, c.pool.isect[slot], ts, tm);
    if (ts < tm) {
        c.pool.t[slot] = ts; c.pool.tmax[slot] = tm;
        return PK_SET_STAGE(pk, ST_CLOUD) | PK_RATIO;
    }
    return PK_SET_STAGE(pk, ST_NEE_DONE);
}
// outcome of sample_interaction (pathtracer.py:200-207) -> EVENT stage
DE_DEV uint32_t finish_interaction(const Ctx &c, int slot, uint32_t pk, uint32_t ev, float t, uint32_t id) {
    c.pool.t[slot] = t;
    pk = PK_SET_EV(pk, ev);
    pk = PK_SET_ID(pk, id);
    return PK_SET_STAGE(pk, ST_EVENT) & ~PK_RATIO;
}
// cloud half of sample_interaction (pathtracer.py:189-198); rmo result is in pk / aux
DE_DEV uint32_t setup_cloud_delta(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d) {
    float ts, tm;
    intersect_cloud_limits(o, d, c.pool.isect[slot], ts, tm);
    uint32_t rmo_ev = PK_RMO_EV(pk);
    float rmo_t = c.pool.aux[slot];
    if ((rmo_ev == kNullEvent || rmo_t > ts) && ts < tm) {
        c.pool.t[slot] = ts; c.pool.tmax[slot] = tm;
        return PK_SET_STAGE(pk, ST_CLOUD) & ~PK_RATIO;
    }
    return finish_interaction(c, slot, pk, rmo_ev, rmo_t, PK_RMO_ID(pk));
}
// rmo half of sample_interaction / sample_transmittance (pathtracer.py:180-186, 219-227);
// isect[slot] holds the land intersection that bounds the ray
DE_DEV uint32_t setup_rmo(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d, bool ratio) {
    float land = c.pool.isect[slot];
    float2 atm = rsi(o, d, kAtmosUpper);
    float t_start = fmaxf(0.0f, atm.x);
    float t_max = land >= 0.0f ? land : atm.y;
    if (atm.y < 0.0f) t_max = -1.0f;
    pk = ratio ? (pk | PK_RATIO) : (pk & ~PK_RATIO);
    if (ratio) c.pool.aux[slot] = 1.0f;
    if (t_start < t_max) {
        c.pool.t[slot] = t_start; c.pool.tmax[slot] = t_max;
        return PK_SET_STAGE(pk, ST_RMO);
    }
    if (!ratio) {  // no atmosphere on the way: NULL event at t_start
        pk = PK_SET_RMO_EV(pk, kNullEvent);
        pk = PK_SET_RMO_ID(pk, 0u);
        c.pool.aux[slot] = t_start;
    }
    return PK_SET_STAGE(pk, ST_RMO_DONE);
}
// intersect_land prologue (pathtracer.py:29-35)
DE_DEV uint32_t setup_sdf(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d, uint32_t draw) {
    float ray_dist = 0.0f;
    float2 rd = rsi(o, d, kAtmosUpper);
    if (rd.x > 0.0f) ray_dist = rd.x;
    c.pool.t[slot] = ray_dist;
    store_draw(c, slot, draw, 0u);
    return PK_SET_STAGE(pk, ST_SDF);
}
// top of the scatter loop (pathtracer.py:349-359)
DE_DEV uint32_t begin_segment(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d) {
    pk &= ~(PK_SHADOW | PK_SURFACE | PK_RATIO | PK_VIS);
    return setup_sdf(c, slot, pk, o, d, 0u);
}

// ST_SDF_DONE: what follows intersect_land -- main ray: sample_interaction; shadow ray: visibility +
// sample_transmittance (pathtracer.py:422-430).  t[slot] holds the intersection distance.
DE_DEV void stage_sdf_done(Ctx &c, int n_members) {
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    if (slot < 0) return;
    uint32_t pk = c.pool.pk[slot];
    float3 o = ld_o(c, slot), d = ld_d(c, slot);
    float isect = c.pool.t[slot];
    bool shadow = (pk & PK_SHADOW) != 0u;
    if (shadow) {
        bool vis = isect < 0.0f;
        pk = vis ? (pk | PK_VIS) : (pk & ~PK_VIS);
        isect = vis ? -1.0f : 0.0f;
        pk &= ~PK_SHADOW;
    }
    c.pool.isect[slot] = isect;
    c.pool.pk[slot] = setup_rmo(c, slot, pk, o, d, shadow);
}
// ST_RMO_DONE: between the rmo pass and the cloud pass of either tracker
DE_DEV void stage_rmo_done(Ctx &c, int n_members) {
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    if (slot < 0) return;
    uint32_t pk = c.pool.pk[slot];
    float3 o = ld_o(c, slot), d = ld_d(c, slot);
    c.pool.pk[slot] = (pk & PK_RATIO) ? setup_cloud_ratio(c, slot, pk, o, d) : setup_cloud_delta(c, slot, pk, o, d);
}

// ------------------------------------------------------------------ path start / end
// ST_NEW: Renderer.render prologue for one sample (renderer.py:305-314); warp-collective work claim
template <bool COUNT> DE_DEV bool stage_new(Ctx &c, int n_members) {
    const unsigned full = 0xFFFFFFFFu;
    const WfParams &P = c.P;
    unsigned chunk = 0u;
    if (c.lane == 0) chunk = atomicAdd(P.next, 1u);
    chunk = __shfl_sync(full, chunk, 0);
    if (chunk >= P.n_chunks) return false;  // no work left: caller retires the ST_NEW slots
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    if (slot < 0) return true;  // unreachable: the caller passes exactly 32 members
    unsigned index = chunk * 32u + (unsigned)c.lane;
    unsigned in_tile = index & 127u, ts = index >> 7;
    unsigned sp = ts % (unsigned)P.n_spp, tile = ts / (unsigned)P.n_spp;
    int px = P.x0 + (int)(tile % (unsigned)P.tiles_x) * kDeTileW + (int)(in_tile & 15u);
    int py = P.y0 + (int)(tile / (unsigned)P.tiles_x) * kDeTileH + (int)(in_tile >> 4);
    if (px >= P.x0 + P.w || py >= P.y0 + P.h) return true;  // outside the window: slot stays ST_NEW
    RngW rng;
    rng.key0 = P.seed; rng.key1 = (uint32_t)(py * c.s.W + px); rng.sample = P.first_sample + sp; rng.bounce = 0u; rng.draw = 0u; rng.valid = false;
    int bin = spectrum_bin(c.s.cdf, rng.next());
    float xu = rng.next(), xv = rng.next();
    float3 dir = get_cast_dir(c.s, c.dv, (float)px, (float)py, xu, xv);
    c.pool.pix[slot] = rng.key1; c.pool.sample[slot] = rng.sample;
    c.pool.thr[slot] = 1.0f; c.pool.L[slot] = 0.0f;
    st_o(c, slot, c.s.cam_pos); st_d(c, slot, dir);
    uint32_t pk = PK_SET_LAM(0u, (uint32_t)bin);
    c.pool.pk[slot] = begin_segment(c, slot, pk, c.s.cam_pos, dir);
    DE_COUNT(c.cn, C_SEGMENTS);
    return true;
}
// pathtracer.py:455-469 + renderer.py:329-330
template <bool COUNT> DE_DEV void end_path(Ctx &c, int slot, uint32_t pk, bool primary_miss, float3 dir) {
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    float Lr = c.pool.L[slot];
    if (primary_miss) {
        if (dot(c.dv.light_dir, dir) > c.dv.sun_cos_angle) Lr += lr.sun_power;
        DE_COUNT(c.cn, C_TEX);
        float3 st = rgb8_ool(c.s.tex[6], dir);
        float stars_power = lr.s2s_valid != 0.0f ? dot(st, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
        Lr += stars_power * lr.sun_power * 0.0000001f;
    }
    if (isinf(Lr) || isnan(Lr) || Lr < 0.0f) Lr = 0.0f;
    DE_COUNT(c.cn, C_PATHS);
    if (Lr != 0.0f) {
        float3 rgb = xyz_to_rgb((Lr * f3(lr.resp_x, lr.resp_y, lr.resp_z)) * lr.rcp_pdf);
        float *a = c.P.accum + (size_t)c.pool.pix[slot] * 3;
        atomicAdd(a, rgb.x); atomicAdd(a + 1, rgb.y); atomicAdd(a + 2, rgb.z);
    }
    c.pool.pk[slot] = ST_NEW;
}

// ------------------------------------------------------------------ loop stages
// SDF sphere tracing, one iteration per loop trip (pathtracer.py:37-44)
template <bool COUNT> DE_DEV void burst_sdf(Ctx &c, int n_members) {
    const unsigned full = 0xFFFFFFFFu;
    int next = min(n_members, 32);
    const int min_active = min(WF_MIN_ACTIVE, (next + 1) / 2);
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    bool active = slot >= 0;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    float t = 0.0f;
    uint32_t iter = 0u;
    if (active) { o = ld_o(c, slot); d = ld_d(c, slot); t = c.pool.t[slot]; iter = c.pool.draw[slot] >> 24; }
    const float scale = c.s.land_height_scale;
    for (int it = 0; it < WF_BURST; ++it) {
        if (active) {
            float3 ro = o + d * t;
            DE_COUNT(c.cn, C_SDF); DE_COUNT(c.cn, C_TEX);
            float dist = length(ro) - kPlanetR - scale * sample_sphere_r8(c.s.tex[1], ro);
            t += dist;
            ++iter;
            if (t > 63710000.0f || fabsf(dist) < t * 0.0001f || iter >= 250u) {
                c.pool.t[slot] = t < 63710000.0f ? t : -1.0f;
                c.pool.pk[slot] = PK_SET_STAGE(c.pool.pk[slot], ST_SDF_DONE);
                active = false;
            }
        }
        unsigned am = __ballot_sync(full, active);
        if (next < n_members) {
            unsigned need = ~am;
            int rank = __popc(need & ((1u << c.lane) - 1u));
            if (!active && next + rank < n_members) {
                slot = c.pool.members[next + rank];
                o = ld_o(c, slot); d = ld_d(c, slot); t = c.pool.t[slot]; iter = c.pool.draw[slot] >> 24;
                active = true;
            }
            next = min(n_members, next + __popc(need));
            am = __ballot_sync(full, active);
        }
        if (__popc(am) < min_active) break;
    }
    if (active) { c.pool.t[slot] = t; c.pool.draw[slot] = (c.pool.draw[slot] & 0xFFFFFFu) | (iter << 24); }
}

// delta / ratio tracking through Rayleigh+Mie+ozone (IS_CLOUD=false) or the cloud shell (true)
// (pathtracer.py:91-112,130-141).  One loop trip = ONE Philox block = TWO collision candidates
// (words 0,1 and 2,3: free flight + acceptance test; a ratio step leaves its second word unused),
// so the RNG is issued converged with static word selection.  A real collision only records the
// slot of its scatter/absorb draw (pathtracer.py:270); ST_EVENT evaluates it for the winner.
template <bool COUNT, bool IS_CLOUD> DE_DEV void burst_track(Ctx &c, int n_members) {
    const unsigned full = 0xFFFFFFFFu;
    int next = min(n_members, 32);
    const int min_active = min(WF_MIN_ACTIVE, (next + 1) / 2);
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    bool active = slot >= 0;
    float3 d = f3(0, 0, 1), pos = f3(0, 0, 0), ext = f3(0, 0, 0);
    float t = 0.0f, tmax = 0.0f, T = 1.0f, max_ext = 1.0f, ext_cloud = 0.0f;
    uint32_t pk = 0u, blk = 0u, key1 = 0u, smp = 0u;
    auto load = [&]() {
        d = ld_d(c, slot);
        t = c.pool.t[slot]; tmax = c.pool.tmax[slot]; T = c.pool.aux[slot];
        pk = c.pool.pk[slot];
        key1 = c.pool.pix[slot]; smp = c.pool.sample[slot];
        blk = ((c.pool.draw[slot] & 0xFFFFFFu) + 3u) >> 2;  // passes start on a block boundary; trips end on one
        pos = ld_o(c, slot) + d * t;
        if (IS_CLOUD) { ext_cloud = cloud_ext_of(PK_SC(pk)); max_ext = ext_cloud * kCloudsDensity; }
        else { const LambdaRow &lr = c.s.lam[PK_LAM(pk)]; ext = f3(lr.ext_r, lr.ext_m, lr.ext_o); max_ext = lr.max_ext_rmo; }
    };
    if (active) load();
    for (int it = 0; it < WF_BURST / 2; ++it) {
        if (active) {
            const bool ratio = (pk & PK_RATIO) != 0u;
            const uint4 rb = philox_block(c.P.seed, key1, smp, PK_SC(pk) + 1u, blk);
            bool done = false;
            uint32_t ev = 0u, id = IS_CLOUD ? 3u : 0u, draw_after = 0u;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                float t_step = -logf(u32_to_unit(h ? rb.z : rb.x)) / max_ext;
                pos = pos + d * t_step;
                t += t_step;
                if (t >= tmax) { done = true; draw_after = 4u * blk + 2u * h + 1u; break; }
                float es0 = 0.0f, es1 = 0.0f, es2 = 0.0f, sum;
                if (IS_CLOUD) {
                    DE_COUNT(c.cn, C_CLOUD);
                    sum = ext_cloud * get_clouds_density<COUNT>(c.s, pos, c.cn);
                } else {
                    DE_COUNT(c.cn, C_RMO);
                    float3 dens = get_density(get_elevation(pos));
                    es0 = ext.x * dens.x; es1 = ext.y * dens.y; es2 = ext.z * dens.z;
                    sum = (es0 + es1) + es2;
                }
                if (ratio) {
                    T *= 1.0f - sum / max_ext;
                    if (T < 1e-5f) { done = true; draw_after = 4u * blk + 2u * h + 2u; break; }
                } else {
                    float rand = u32_to_unit(h ? rb.w : rb.y);
                    if (rand < sum / max_ext) {
                        if (!IS_CLOUD) {
                            float cmf = es0;
                            if (!(rand < cmf / max_ext)) {
                                id = 1u; cmf += es1;
                                if (!(rand < cmf / max_ext)) { id = 2u; cmf += es2; if (!(rand < cmf / max_ext)) id = 3u; }
                            }
                        }
                        ev = 1u; done = true; draw_after = 4u * blk + 2u * h + 3u;  // the slot before draw_after decides scatter vs absorb
                        break;
                    }
                }
            }
            ++blk;
            if (done) {
                store_draw(c, slot, draw_after, 0u);
                uint32_t npk;
                if (ratio) {
                    c.pool.aux[slot] = T;
                    npk = PK_SET_STAGE(pk, IS_CLOUD ? ST_NEE_DONE : ST_RMO_DONE);
                } else if (IS_CLOUD) {
                    uint32_t rmo_ev = PK_RMO_EV(pk);
                    float rmo_t = c.pool.aux[slot];
                    if (ev > 0u && (t < rmo_t || rmo_ev == 0u)) {
                        c.pool.nb[slot] = __uint_as_float(draw_after - 1u);
                        npk = finish_interaction(c, slot, pk, 1u, t, kCloud);
                    } else npk = finish_interaction(c, slot, pk, rmo_ev, rmo_t, PK_RMO_ID(pk));
                } else {
                    npk = PK_SET_RMO_EV(pk, ev);
                    npk = PK_SET_RMO_ID(npk, id);
                    c.pool.aux[slot] = t;
                    if (ev) c.pool.nb[slot] = __uint_as_float(draw_after - 1u);
                    npk = PK_SET_STAGE(npk, ST_RMO_DONE);
                }
                c.pool.pk[slot] = npk;
                active = false;
            }
        }
        unsigned am = __ballot_sync(full, active);
        if (next < n_members) {
            unsigned need = ~am;
            int rank = __popc(need & ((1u << c.lane) - 1u));
            if (!active && next + rank < n_members) { slot = c.pool.members[next + rank]; load(); active = true; }
            next = min(n_members, next + __popc(need));
            am = __ballot_sync(full, active);
        }
        if (__popc(am) < min_active) break;
    }
    if (active) { c.pool.t[slot] = t; if (pk & PK_RATIO) c.pool.aux[slot] = T; store_draw(c, slot, 4u * blk, 0u); }
}

// ST_EVENT: after sample_interaction, pathtracer.py:369-444 up to the point where the NEE ray is traced
template <bool COUNT> DE_DEV void stage_event(Ctx &c, int n_members) {
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    if (slot < 0) return;
    uint32_t pk = c.pool.pk[slot];
    const uint32_t sc = PK_SC(pk);
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    float3 o = ld_o(c, slot), d = ld_d(c, slot);
    RngW rng = load_rng(c, slot, pk);
    uint32_t ev = PK_EV(pk), id = PK_ID(pk);
    if (ev) {  // a real collision: scatter or absorb (pathtracer.py:108-111,263-270) from the recorded slot
        uint32_t ds = __float_as_uint(c.pool.nb[slot]);
        uint4 b = philox_block(rng.key0, rng.key1, rng.sample, rng.bounce, ds >> 2);
        uint32_t w = (ds & 3u) == 0u ? b.x : ((ds & 3u) == 1u ? b.y : ((ds & 3u) == 2u ? b.z : b.w));
        float albedo = id == 0u ? 1.0f : (id == 1u ? 0.95f : (id == 2u ? 0.0f : 0.99f));
        ev = u32_to_unit(w) < albedo ? (uint32_t)kScatterEvent : (uint32_t)kAbsorbEvent;
    }
    if (sc > 9u && id == (uint32_t)kCloud) id = kIsoCloud;
    rng.align();
    float3 light_dir = sample_cone_oriented(c.dv.sun_cos_angle, c.dv.light_dir, rng);
    if (ev == (uint32_t)kAbsorbEvent) { end_path<COUNT>(c, slot, pk, false, d); return; }
    if (ev == (uint32_t)kScatterEvent) {
        float3 ipos = o + d * c.pool.t[slot];
        bool blocked = rsi(ipos, light_dir, kPlanetR).y > 0.0f;
        c.pool.na[slot] = evaluate_phase(d, light_dir, (int)id, sc > 0u);
        c.pool.mdx[slot] = d.x; c.pool.mdy[slot] = d.y; c.pool.mdz[slot] = d.z;
        st_o(c, slot, ipos); st_d(c, slot, light_dir);
        pk = PK_SET_ID(pk, id) & ~PK_SURFACE;
        store_draw(c, slot, rng.draw, 0u);
        if (blocked) { c.pool.aux[slot] = 0.0f; c.pool.pk[slot] = PK_SET_STAGE(pk, ST_NEE_DONE); }
        else { c.pool.isect[slot] = -1.0f; c.pool.pk[slot] = setup_rmo(c, slot, pk, ipos, light_dir, true); }
        return;
    }
    float earth_isect = c.pool.isect[slot];
    if (earth_isect > 0.0f) {
        DE_COUNT(c.cn, C_SURF);
        float3 land_pos = o + d * earth_isect;
        // land_normal (pathtracer.py:16-25) and get_land_material (:284-313) on the shared fetch routine
        const float hs = c.s.land_height_scale, eps = c.dv.normal_eps;
        if (COUNT) { c.cn.v[C_SDF] += 4; c.cn.v[C_TEX] += 8; }
        float sd0 = length(land_pos) - kPlanetR - hs * r8_ool(c.s.tex[1], land_pos);
        float3 px_ = f3(land_pos.x - eps, land_pos.y, land_pos.z), py_ = f3(land_pos.x, land_pos.y - eps, land_pos.z), pz_ = f3(land_pos.x, land_pos.y, land_pos.z - eps);
        float3 nrm = normalize(f3(sd0 - (length(px_) - kPlanetR - hs * r8_ool(c.s.tex[1], px_)), sd0 - (length(py_) - kPlanetR - hs * r8_ool(c.s.tex[1], py_)),
                                  sd0 - (length(pz_) - kPlanetR - hs * r8_ool(c.s.tex[1], pz_))));
        LandMaterial m;
        m.ocean = r8_ool(c.s.tex[2], land_pos);
        m.albedo_srgb = grade_albedo(rgb8_ool(c.s.tex[0], land_pos), m.ocean);
        m.bathymetry = r8_ool(c.s.tex[4], land_pos);
        m.emissive = r8_ool(c.s.tex[5], land_pos);
        float albedo = lr.s2s_valid != 0.0f ? dot(m.albedo_srgb, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
        c.pool.L[slot] += c.pool.thr[slot] * m.emissive * lr.nightlights_power;
        float3 offset_pos = land_pos * (1.0f + 0.0001f * c.s.land_height_scale / 12000.0f);
        float ndl;
        float dbrdf = earth_brdf(albedo, m.ocean, m.bathymetry, -d, nrm, light_dir, ndl);
        c.pool.na[slot] = dbrdf; c.pool.nb[slot] = ndl;
        c.pool.nx[slot] = nrm.x; c.pool.ny[slot] = nrm.y; c.pool.nz[slot] = nrm.z;
        c.pool.m0[slot] = albedo; c.pool.m1[slot] = m.ocean; c.pool.m2[slot] = m.bathymetry;
        c.pool.mdx[slot] = d.x; c.pool.mdy[slot] = d.y; c.pool.mdz[slot] = d.z;
        st_o(c, slot, offset_pos); st_d(c, slot, light_dir);
        pk |= PK_SURFACE | PK_SHADOW;
        c.pool.pk[slot] = setup_sdf(c, slot, pk, offset_pos, light_dir, rng.draw);
        return;
    }
    end_path<COUNT>(c, slot, pk, sc == 0u, d);  // escaped (pathtracer.py:441-444)
}

// ST_NEE_DONE: after the NEE transmittance, pathtracer.py:394-401 / 431-439, Russian roulette :447-453, next segment
template <bool COUNT> DE_DEV void stage_nee_done(Ctx &c, int n_members) {
    int slot = c.lane < n_members ? c.pool.members[c.lane] : -1;
    if (slot < 0) return;
    uint32_t pk = c.pool.pk[slot];
    uint32_t sc = PK_SC(pk);
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    RngW rng = load_rng(c, slot, pk);
    float3 o = ld_o(c, slot);  // interaction position / offset position
    float3 main_d = f3(c.pool.mdx[slot], c.pool.mdy[slot], c.pool.mdz[slot]);
    float T = c.pool.aux[slot], thr = c.pool.thr[slot], Lacc = c.pool.L[slot];
    float3 nd;
    if (pk & PK_SURFACE) {
        float vis = (pk & PK_VIS) ? 1.0f : 0.0f;
        Lacc += thr * T * vis * lr.sun_irradiance * c.pool.na[slot] * c.pool.nb[slot];
        float3 nrm = f3(c.pool.nx[slot], c.pool.ny[slot], c.pool.nz[slot]);
        rng.align();
        nd = sample_hemisphere_cosine_weighted(nrm, rng);
        float unused;
        float brdf = earth_brdf(c.pool.m0[slot], c.pool.m1[slot], c.pool.m2[slot], -main_d, nrm, nd, unused);
        thr *= brdf * kPi;
    } else {
        Lacc += thr * T * lr.sun_irradiance * c.pool.na[slot];
        float pdp;
        rng.align();
        nd = sample_phase(main_d, (int)PK_ID(pk), sc > 0u, rng, pdp);
        thr *= pdp;
    }
    c.pool.L[slot] = Lacc;
    bool terminate = false;
    if (sc > 3u) {
        float p = fmaxf(0.05f, 1.0f - thr);
        if (rng.next() < p) terminate = true;
        else thr /= 1.0f - p;
    }
    ++sc;
    if (terminate || sc >= 25u) { end_path<COUNT>(c, slot, pk, false, nd); return; }
    c.pool.thr[slot] = thr;
    st_d(c, slot, nd);
    pk = PK_SET_SC(pk, sc);
    DE_COUNT(c.cn, C_SEGMENTS);
    c.pool.pk[slot] = begin_segment(c, slot, pk, o, nd);
}

// ------------------------------------------------------------------ the persistent kernel
template <bool COUNT> __global__ void __launch_bounds__(WF_WARPS * 32, 1) k_render_wavefront(const __grid_constant__ DevScene s, const __grid_constant__ WfParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpPool *pools = reinterpret_cast<WarpPool *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    const DevDerived dv = *s.derived;
    Counters cn;
    cn.clear();
    Ctx c{s, dv, P, pools[warp], cn, lane};
    WarpPool &pool = c.pool;
    for (int k = lane; k < WF_N; k += 32) pool.pk[k] = ST_NEW;
    __syncwarp();
    bool work_left = true;
    for (;;) {
        // 1. census of stages (warp ballots over the pool's stage words)
        int cnt[ST_COUNT];
#pragma unroll
        for (int q = 0; q < ST_COUNT; ++q) cnt[q] = 0;
        for (int k = 0; k < WF_N; k += 32) {
            uint32_t st = PK_STAGE(pool.pk[k + lane]);
#pragma unroll
            for (int q = 1; q < ST_COUNT; ++q) cnt[q] += __popc(__ballot_sync(full, st == (uint32_t)q));
        }
        if (!work_left) cnt[ST_NEW] = 0;
        // regeneration needs a whole chunk of 32 free slots while work is plentiful; the other
        // stages compete on population
        int best = 0, best_n = 0;
#pragma unroll
        for (int q = 1; q < ST_COUNT; ++q) {
            int n = cnt[q];
            if (q == ST_NEW && n < 32) n = 0;
            if (n > best_n) { best_n = n; best = q; }
        }
        if (best_n == 0) break;  // pool drained (free slots always come in groups >= 32 until then: WF_N >= 64)
        // 2. compact the members of that stage (warp-ballot compaction)
        int base = 0;
        for (int k = 0; k < WF_N; k += 32) {
            bool is = PK_STAGE(pool.pk[k + lane]) == (uint32_t)best;
            unsigned m = __ballot_sync(full, is);
            if (is) pool.members[base + __popc(m & ((1u << lane) - 1u))] = (uint8_t)(k + lane);
            base += __popc(m);
        }
        __syncwarp();
        // 3. run it
        if (best == ST_SDF) burst_sdf<COUNT>(c, best_n);
        else if (best == ST_RMO) burst_track<COUNT, false>(c, best_n);
        else if (best == ST_CLOUD) burst_track<COUNT, true>(c, best_n);
        else {
            // a work chunk is exactly 32 paths: regeneration only runs on whole groups of 32 free slots
            const int lim = best == ST_NEW ? (best_n & ~31) : best_n;
            for (int b = 0; b < lim; b += 32) {
                if (b) { uint8_t mv = lane + b < best_n ? pool.members[lane + b] : 0; __syncwarp(); pool.members[lane] = mv; __syncwarp(); }
                int n = min(32, best_n - b);
                if (best == ST_SDF_DONE) stage_sdf_done(c, n);
                else if (best == ST_RMO_DONE) stage_rmo_done(c, n);
                else if (best == ST_EVENT) stage_event<COUNT>(c, n);
                else if (best == ST_NEE_DONE) stage_nee_done<COUNT>(c, n);
                else { if (!stage_new<COUNT>(c, n)) { work_left = false; break; } }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    if (COUNT) cn.flush(s.counters);
}

}  // namespace de_fast

struct DeWavefrontState {
    int device = 0, sm_count = 0;
    unsigned int *d_next = nullptr;
    bool attr_set = false;
};

DeWavefrontState *de_wavefront_alloc(int device) {
    DeWavefrontState *st = new DeWavefrontState();
    st->device = device;
    cudaDeviceGetAttribute(&st->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaMalloc(&st->d_next, sizeof(unsigned int)) != cudaSuccess) { delete st; return nullptr; }
    return st;
}
void de_wavefront_free(DeWavefrontState *st) {
    if (!st) return;
    cudaFree(st->d_next);
    delete st;
}
void de_wavefront_render(DeWavefrontState *st, const DevScene &s, float *accum, int n_spp, uint32_t seed, uint32_t first_sample, int x0, int y0,
                         int w, int h, bool count, cudaStream_t stream) {
    using namespace de_fast;
    size_t smem = sizeof(WarpPool) * WF_WARPS;
    if (!st->attr_set) {
        cudaFuncSetAttribute(k_render_wavefront<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_render_wavefront<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        st->attr_set = true;
    }
    WfParams P;
    P.accum = accum; P.next = st->d_next;
    P.tiles_x = (w + kDeTileW - 1) / kDeTileW;
    int tiles = P.tiles_x * ((h + kDeTileH - 1) / kDeTileH);
    P.n_chunks = (unsigned)tiles * (unsigned)n_spp * 4u;
    P.n_spp = n_spp; P.x0 = x0; P.y0 = y0; P.w = w; P.h = h; P.seed = seed; P.first_sample = first_sample;
    cudaMemsetAsync(st->d_next, 0, sizeof(unsigned int), stream);
    int grid = st->sm_count;
    if (count) k_render_wavefront<true><<<grid, WF_WARPS * 32, smem, stream>>>(s, P);
    else k_render_wavefront<false><<<grid, WF_WARPS * 32, smem, stream>>>(s, P);
}
