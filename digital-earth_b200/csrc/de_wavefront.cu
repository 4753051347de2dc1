// de_wavefront.cu -- the product integrator: persistent-thread, stage-sorted wavefront.
//
// Why: ncu on the one-thread-per-pixel kernel (profiles/r1_wavefront.md) shows 4.9 of 32 lanes
// active per issued instruction -- path length and stage mix diverge, memory does not matter
// (L1 97 %, L2 98 % hits, DRAM idle).  This kernel keeps every lane of a warp inside the SAME
// inner loop:
//   * one persistent CTA per SM owns a pool of WF_SLOTS (2464) path states: the 16 words every stage needs in shared
//     memory (SoA, 64 B per path), the 13 words only the shading stages touch as one 64-B record per path in global
//     memory that stays in L2 (WF_COLD) -- path state never touches HBM, and the pool is large enough for the stage
//     queues to hold several groups, which is what keeps a stage body in the 32 KB instruction cache between visits
//     (the kernel is bound by instruction fetch: profiles/r2_bench.md);
//   * a path is a small state machine  SDF -> SDF_DONE -> RMO -> RMO_DONE -> CLOUD -> EVENT ->
//     (SURFACE -> SDF ->) RMO -> RMO_DONE -> CLOUD -> NEE_DONE -> SDF ...  (pathtracer.py:349-453 cut
//     at its loop boundaries); every stage has a ring queue of ready slots in shared memory;
//   * a warp pops up to 32 slots of one queue -- the stage the whole SM is working on while that has a
//     full group (the instruction caches then hold one body), else the fullest -- and runs the stage
//     with all lanes converged: loop stages (SDF / RMO / CLOUD) in bursts, refilling idle lanes from
//     the same queue, the transition and shading stages as one-shot bodies; a finished lane only
//     records its result and pushes the slot to the next stage's queue (one reservation per group);
//   * code size is throughput here (profiles/r1_bench.md): one copy of every large helper per stage,
//     cold math out of line, no dead fallback paths;
//   * terminated paths free their slot; free slots are refilled 32 at a time from a global atomic
//     work counter (path regeneration): 32 neighbouring pixels of one 16x8 film tile, one sample.
// The random stream of a path is the contract of include/de_api.h (Philox key (seed,pixel), counter
// (sample,bounce,slot>>2)), so a pixel's samples are the same paths in every integrator flavour.
#define DE_TEX_OBJ_ONLY 1
#ifndef WF_SHRINK
#define WF_SHRINK 0  // 1: one shared out-of-line copy of the equirect mapping inside the cloud bound, of the rmo segment majorant (2 inlined copies) and
                     // of the terrain-march prologue (3 inlined copies): the kernel is bound by instruction fetch (profiles/r2_bench.md)
#endif
#define DE_WF_SHRINK WF_SHRINK
#include "de_integrator.cuh"
#include "de_launch.h"
#include "de_wavefront.h"

namespace de_fast {

#ifndef WF_RMO_BANDS
#define WF_RMO_BANDS 0  // 1: altitude-band majorants for the rmo passes (de_device.cuh: rmo_band_*; one more state word per path).  Valid (the
                        // device functions are checked ray by ray in tests/test_gpu_bounds.py) and it cuts the rmo candidates per path 3-7x, yet the
                        // frame is not faster (profiles/r2_bench.md): an rmo pass is 1-3 candidates either way, what it costs is the stage visit
                        // (pop, state load, one Philox block, flush, push), not the candidates.  Off; kept for the record.
#endif
#ifndef WF_COLD
#define WF_COLD 1      // 1: the 13 words of a path's state that only the shading stages touch (throughput, radiance, NEE factors, main direction,
                       // normal, material) live in global memory (64 B per slot, L2-resident: 148 x 2464 x 64 B = 23 MB) instead of shared
                       // memory; the pool then holds 2464 paths of 16 hot words instead of 1728 of 28.  Why: the kernel is bound by
                       // instruction fetch (profiles/r2_bench.md) and how often a stage body is re-fetched falls with the number of paths
                       // WAITING in the stage queues (pool size minus the 32 x 32 in flight): 750 -> 1480.
#endif
#ifndef WF_SLOTS
#if WF_COLD
#define WF_SLOTS 2464
#elif WF_RMO_BANDS
#define WF_SLOTS 1664  // path states per CTA (one CTA per SM): 188 KB of state (29 words each) + 36 KB of queues
#else
#define WF_SLOTS 1728  // 28 words each
#endif
#endif
#ifndef WF_WARPS
#define WF_WARPS 32
#endif
#ifndef WF_BURST
#define WF_BURST 64    // max loop iterations per burst
#endif
#ifndef WF_MIN_ACTIVE
#define WF_MIN_ACTIVE 24  // a burst ends when fewer lanes than this are busy and the queue is dry
#endif
#ifndef WF_PHASE
#define WF_PHASE 32  // SM-wide phase: all warps prefer one stage while it has a full group (I-cache: +11-23 %, profiles/r1_bench.md)
#endif
#ifndef WF_STICKY
#define WF_STICKY 64
#endif
#ifndef WF_CYCLIC
#define WF_CYCLIC 0  // 1: when the phase stage runs out of full groups the SM moves on to the NEXT stage in dataflow order that has one (NEW, SDF,
                     // SDF_DONE, RMO, RMO_DONE, CLOUD, EVENT, SURFACE, NEE_DONE, ...) instead of the fullest queue: the pool's paths then travel
                     // as a wave and the warps of the SM sit on a few adjacent stage bodies (instruction cache, profiles/r2_bench.md)
#endif
#ifndef WF_MIN_FRAC8
#define WF_MIN_FRAC8 6   // ... or this many eighths of the lanes the burst started with
#endif
#ifndef WF_REFILL_MIN
#define WF_REFILL_MIN 10 // idle lanes that trigger a mid-burst refill
#endif
// ring capacity per stage queue (power of two > WF_SLOTS); an entry is slot | lap tag << WF_SLOT_BITS in 16 bits
constexpr int WF_SLOT_BITS = WF_SLOTS < 2048 ? 11 : 12;
constexpr int WF_RING = 1 << WF_SLOT_BITS;
constexpr unsigned WF_TAG_MASK = (1u << (16 - WF_SLOT_BITS)) - 1u;
static_assert(WF_SLOTS < WF_RING && WF_SLOTS % 32 == 0, "ring / slot-id encoding");

// Stages.  Loop stages (SDF, RMO, CLOUD) run bursts of a small loop body; the others are one-shot
// bodies executed converged over up to 32 slots.  A loop body never runs transition code.
enum : uint32_t { ST_NEW = 0, ST_SDF, ST_RMO, ST_CLOUD, ST_SDF_DONE, ST_RMO_DONE, ST_EVENT, ST_NEE_DONE, ST_SURFACE, ST_COUNT };

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

struct WarpPool {  // CTA-wide pool, SoA: lane l touching slot s hits bank s%32
    float ox[WF_SLOTS], oy[WF_SLOTS], oz[WF_SLOTS], dx[WF_SLOTS], dy[WF_SLOTS], dz[WF_SLOTS];
    uint32_t pix[WF_SLOTS], sample[WF_SLOTS], pk[WF_SLOTS], draw[WF_SLOTS];  // draw: rng slot index [0:24) | sdf iteration [24:32)
    float t[WF_SLOTS], tmax[WF_SLOTS], aux[WF_SLOTS], isect[WF_SLOTS];       // aux: rmo_t (delta) or transmittance (ratio)
#if WF_COLD
    uint32_t dsl[WF_SLOTS];                                                   // rng slot of the scatter / absorb decision of the pending collision
#else
    float thr[WF_SLOTS], L[WF_SLOTS];
    float mdx[WF_SLOTS], mdy[WF_SLOTS], mdz[WF_SLOTS];                        // main ray direction while the NEE ray is tracked
    float nx[WF_SLOTS], ny[WF_SLOTS], nz[WF_SLOTS], m0[WF_SLOTS], m1[WF_SLOTS], m2[WF_SLOTS];  // surface normal, albedo, ocean, bathymetry
    float na[WF_SLOTS], nb[WF_SLOTS];                                         // NEE factors: phase | brdf, n.l (nb doubles as the decision-slot word)
#endif
    float cmj[WF_SLOTS];                                                      // tracking pass: local majorant (cloud: density bound; rmo: sigma.rho bound of the whole segment)
#if WF_RMO_BANDS
    float tlim[WF_SLOTS];                                                     // rmo pass: where the ray leaves its current altitude band (draw[24:32) holds the band)
#endif
    // per-stage MPMC ring queues of ready slots: entry = slot | (lap & 31) << 11
    uint16_t ring[ST_COUNT][WF_RING];
    unsigned int q_tail[ST_COUNT], q_head[ST_COUNT];
    int q_avail[ST_COUNT];
    int retired;     // slots that found no more work
    int phase;       // SM-wide preferred stage (WF_PHASE)
    int last_visit;  // counting build: stage of the SM's latest visit
    int work_left;
    unsigned int n_chunks;  // work units of this launch: tiles x samples x 4 quarter-tiles
    unsigned int claimed;   // chunks this CTA claimed (timeline only)
    // drain diagnostics (timeline only): stage visits / slots handled after this CTA found the work counter exhausted
    unsigned int dr_visits[ST_COUNT], dr_slots[ST_COUNT];
    unsigned long long t_exhaust, t_few;  // ... when it did, and when fewer than 64 of its paths were still alive
};

static_assert(sizeof(WarpPool) <= 227 * 1024, "the pool must fit the 227 KB of shared memory a CTA can opt into on sm_100");

struct WfParams {
    float *accum;
    float4 *cold;        // WF_COLD: [gridDim.x][WF_SLOTS][4] shading state of the paths in flight
    float *accum2;       // optional per-pixel second moments (sum of squared contributions), or nullptr
    unsigned int *next;  // global work counter (units of 32 paths)
    const unsigned int *tile_list;  // film tiles this kernel renders (k_classify_tiles: everything that is not pure space)
    const unsigned int *n_tiles;    // their number (device: the classification never syncs with the host)
    int n_spp, x0, y0, w, h, tiles_x;
    uint32_t seed, first_sample;
    unsigned long long *prof;  // counting build: per stage {cycles, visits, lanes}, + idle cycles at index ST_COUNT
    unsigned long long *cta_stats;  // optional with timeline: [gridDim.x][24] per-CTA drain diagnostics
    unsigned long long *timeline;  // optional, counting build only: globaltimer ns {first CTA start, first / last CTA to see the work counter exhausted,
                                   // first / last CTA end, min / max chunks claimed by a CTA}
};

#ifndef WF_SPACE_SHORTCUT
#define WF_SPACE_SHORTCUT 1
#endif
#ifndef WF_CHAIN
#define WF_CHAIN 0   // >0: one-shot stages hand their largest group of successors (>= this many lanes) straight to the next stage, no queue
                     // round trip.  Measured 8-10 % SLOWER at 8/16/24 (profiles/r1_bench.md): it breaks the SM-wide phase, and the
                     // instruction cache matters more than the queue traffic.  Kept for the record, off.
#endif
#ifndef WF_DRAIN_FAST
#define WF_DRAIN_FAST 0  // 1: once the work counter is exhausted bursts run until every lane is done and one-shot stages hand their
                         // successors straight to the next stage.  Measured (profiles/r2_tail.md): the drain is NOT shorter (it is the serial
                         // latency of the longest path, not queue hops) and the extra code costs 8 % in steady state (no_instruction stalls
                         // 1.6 -> 2.8 per issue).  Kept for the record, off.
#endif
#ifndef WF_CHAIN_TOPUP
#define WF_CHAIN_TOPUP 4  // idle lanes of a chained group that trigger a top-up from the next stage's queue
#endif
#ifndef WF_OOL_MASK
#define WF_OOL_MASK 0  // bit 0: end_path out of line, bit 1: setup_sdf out of line
#endif
#ifndef WF_TRACK_PHILOX_INLINE
#define WF_TRACK_PHILOX_INLINE 0
#endif
#ifndef WF_FUSE_RMO
#define WF_FUSE_RMO 0    // ST_SDF_DONE also runs the FIRST trip (one Philox block, two candidates) of the rmo pass it sets up and, when that ends the
                         // pass (it usually does: with the segment majorant an rmo pass averages ~2 candidates), the ST_RMO_DONE transition as well:
                         // SDF_DONE -> RMO -> RMO_DONE collapses into one stage visit for most rays (two pops, pushes and state round trips less);
                         // the NEE ray of a scatter event enters through the same stage.  The random stream is untouched (a pass starts on a block
                         // boundary, a trip is one block), so the paths are the same paths.  profiles/r2_bench.md
#endif
#ifndef WF_FUSE_CLOUD
#define WF_FUSE_CLOUD 0  // the ST_RMO_DONE transition also runs the first trip of the cloud pass it sets up (same idea as WF_FUSE_RMO)
#endif
static_assert(!(WF_FUSE_RMO && WF_RMO_BANDS), "the fused first trip does not walk altitude bands");
#ifndef WF_IDLE_EXP
#define WF_IDLE_EXP 5    // a warp that finds every queue empty sleeps 64 ns << min(consecutive empty rounds, 5) (64 ns ... 2 us) instead of a flat 64 ns:
                         // half of all scheduling rounds are such polls, and they compete with working warps for issue slots (+0.5 ... 1.4 %)
#endif
#ifndef WF_BACKOFF_NS
#define WF_BACKOFF_NS 0  // sleep after a pop that lost the race for the last group of a queue (idle warps otherwise spin through the scheduler)
#endif
#ifndef WF_PHILOX_UNROLL
#define WF_PHILOX_UNROLL 5  // partly rolled: 10 unrolled rounds are 2.8 KB of the hottest shared code (I-cache); 5 measured best
#endif
constexpr int kPhiloxUnroll = WF_PHILOX_UNROLL;
// One Philox4x32-10 block; deliberately NOT inlined: ~70 instructions that would otherwise be
// replicated at every draw site and blow the instruction cache (profiles/r1_wavefront.md).
DE_DEV uint4 philox_block_inl(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2) {
    uint32_t c3 = 0u;
#pragma unroll kPhiloxUnroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __noinline__ uint4 philox_block(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2) {
    uint32_t c3 = 0u;
#pragma unroll kPhiloxUnroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
DE_DEV unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
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
__device__ __noinline__ float fetch_r8_ool(cudaTextureObject_t obj, int w, int h, float px, float py, float pz) {
    DevTex t; t.data = nullptr; t.w = w; t.h = h; t.c = 1; t.obj = obj;
    return sample_sphere_r8(t, f3(px, py, pz));
}
__device__ __noinline__ float3 fetch_rgb8_ool(cudaTextureObject_t obj, int w, int h, float px, float py, float pz) {
    DevTex t; t.data = nullptr; t.w = w; t.h = h; t.c = 3; t.obj = obj;
    return sample_sphere_rgb8(t, f3(px, py, pz));
}
// cold, large bodies shared by the one-shot stages (kept out of line for the instruction cache)
__device__ __noinline__ float2 brdf_ool(float albedo, float ocean, float bathy, float vx, float vy, float vz, float nx, float ny, float nz, float lx, float ly, float lz) {
    float ndl;
    float b = earth_brdf(albedo, ocean, bathy, f3(vx, vy, vz), f3(nx, ny, nz), f3(lx, ly, lz), ndl);
    return make_float2(b, ndl);
}
__device__ __noinline__ float phase_eval_ool(float ax, float ay, float az, float bx, float by, float bz, int id, bool reduce) {
    return evaluate_phase(f3(ax, ay, az), f3(bx, by, bz), id, reduce);
}
struct BlockRng {  // up to four draws from one Philox block, static order
    uint4 b; int i;
    DE_DEV float next() { uint32_t v = i == 0 ? b.x : (i == 1 ? b.y : (i == 2 ? b.z : b.w)); ++i; return u32_to_unit(v); }
    DE_DEV void align() {}
};
// sample_phase (pathtracer.py:249-261) on one Philox block; returns (dir, phase/pdf) and the number of words used in .w of the int
__device__ __noinline__ float4 phase_sample_ool(float dx, float dy, float dz, int id, bool reduce, uint4 blk, int *used) {
    BlockRng r{blk, 0};
    float pdp;
    float3 d = sample_phase(f3(dx, dy, dz), id, reduce, r, pdp);
    *used = r.i;
    return make_float4(d.x, d.y, d.z, pdp);
}
__device__ __noinline__ float2 sphere_uv_ool(float px, float py, float pz) { return sphere_uv(f3(px, py, pz)); }
#if WF_SHRINK
__device__ __noinline__ float rmo_majorant_ool(float ex, float em, float eo, float ox, float oy, float oz, float dx, float dy, float dz, float ts, float tm) {
    return rmo_segment_majorant(f3(ex, em, eo), f3(ox, oy, oz), f3(dx, dy, dz), ts, tm);
}
// intersect_land prologue (pathtracer.py:29-35): where the march starts, or -1 when the ray surely misses the terrain
__device__ __noinline__ float sdf_start_ool(float ox, float oy, float oz, float dx, float dy, float dz, float scale) {
    const float3 o = f3(ox, oy, oz), d = f3(dx, dy, dz);
    float ray_dist = 0.0f;
    const float2 rd = rsi(o, d, kAtmosUpper);
    if (rd.x > 0.0f) ray_dist = rd.x;
    const float3 p = o + d * ray_dist;
    if (land_surely_missed(p, d, ray_dist, scale)) return -1.0f;
    return ray_dist + skip_to_terrain_top(p, d, ray_dist, scale);
}
#endif
DE_DEV float r8_ool(const DevTex &t, float3 p) { return fetch_r8_ool(t.obj, t.w, t.h, p.x, p.y, p.z); }
DE_DEV float3 rgb8_ool(const DevTex &t, float3 p) { return fetch_rgb8_ool(t.obj, t.w, t.h, p.x, p.y, p.z); }

// The ray leaves altitude band k at parameter t: returns (t where it leaves the band it enters, that band as float bits).  Out of line: called
// rarely and under divergence.  k < 0: t is the start of the pass, the band is looked up from the altitude there.
__device__ __noinline__ float2 rmo_band_advance(const DevScene &s, float ox, float oy, float oz, float dx, float dy, float dz, float t, int k) {
    const float3 d = f3(dx, dy, dz), q = f3(ox, oy, oz) + d * t;
    int kn;
    float ds;
    if (k < 0) { kn = rmo_band_of(s, sqrtf(dot(q, q))); ds = rmo_band_exit(s, q, d, kn); }
    else ds = rmo_band_cross(s, q, d, k, kn);
    return make_float2(t + ds, __int_as_float(kn));
}
struct Ctx {  // per-warp context
    const DevScene &s;
    const DevDerived &dv;
    const WfParams &P;
    WarpPool &pool;
    Counters &cn;
    int lane;
    float4 *cold;  // this CTA's part of WfParams::cold
};

// ---- stage queues -------------------------------------------------------------------------
// push: reserve a ring position, write the lap-tagged entry, publish it.  pop: take up to `want`
// published entries (semaphore style), then read the reserved positions, spinning on the lap tag
// for the rare entry whose producer reserved earlier but has not written yet.
DE_DEV void q_push(WarpPool &p, uint32_t st, int slot) {
    unsigned int pos = atomicAdd(&p.q_tail[st], 1u);
    volatile uint16_t *r = p.ring[st];
    r[pos & (WF_RING - 1)] = (uint16_t)((unsigned)slot | (((pos / WF_RING) & WF_TAG_MASK) << WF_SLOT_BITS));
    __threadfence_block();
    atomicAdd(&p.q_avail[st], 1);
}
// warp-aggregated push of the lanes in `mask` (all to stage st): one reservation for the group
#ifndef WF_PUSHGROUP_INLINE
#define WF_PUSHGROUP_INLINE 1  // measured: out of line costs 4 % (it is called under divergence)
#endif
#if WF_PUSHGROUP_INLINE
DE_DEV
#else
__device__ __noinline__
#endif
void q_push_group(WarpPool &p, uint32_t st, int slot, unsigned mask, int lane) {
    int n = __popc(mask), leader = __ffs(mask) - 1;
    unsigned int base = 0u;
    if (lane == leader) base = atomicAdd(&p.q_tail[st], (unsigned)n);
    base = __shfl_sync(mask, base, leader);
    unsigned int pos = base + (unsigned)__popc(mask & ((1u << lane) - 1u));
    volatile uint16_t *r = p.ring[st];
    r[pos & (WF_RING - 1)] = (uint16_t)((unsigned)slot | (((pos / WF_RING) & WF_TAG_MASK) << WF_SLOT_BITS));
    __threadfence_block();
    __syncwarp(mask);
    if (lane == leader) atomicAdd(&p.q_avail[st], n);
}
// one-shot stages: every lane with a slot pushes it to stage PK_STAGE(npk); grouped per target stage
// (out of line on purpose: warp-collective, called converged from many sites -- inlined copies of the
// queue code were 27 KB of the kernel and the instruction cache is the scarce resource here)
__device__ __noinline__ void q_push_sorted(WarpPool &p, bool has, uint32_t st, int slot, int lane) {
    unsigned todo = __ballot_sync(0xFFFFFFFFu, has);
    while (todo) {
        int leader = __ffs(todo) - 1;
        uint32_t lst = __shfl_sync(0xFFFFFFFFu, st, leader);
        unsigned grp = __ballot_sync(0xFFFFFFFFu, has && st == lst);
        if (has && st == lst) q_push_group(p, lst, slot, grp, lane);
        todo &= ~grp;
    }
}
// warp-collective: returns the number of slots obtained (<= want); lane i < n receives its slot
__device__ __noinline__ int2 q_pop2(WarpPool &p, uint32_t st, int want, int lane) {
    unsigned int base = 0u;
    int n = 0, slot;
    if (lane == 0) {
        int a = atomicSub(&p.q_avail[st], want);
        n = a >= want ? want : (a > 0 ? a : 0);
        if (n < want) atomicAdd(&p.q_avail[st], want - n);
        if (n) base = atomicAdd(&p.q_head[st], (unsigned)n);
    }
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    slot = -1;
    if (lane < n) {
        unsigned int pos = base + (unsigned)lane;
        const unsigned tag = (pos / WF_RING) & WF_TAG_MASK;
        volatile uint16_t *r = p.ring[st];
        unsigned e;
        do { e = r[pos & (WF_RING - 1)]; } while ((e >> WF_SLOT_BITS) != tag);
        slot = (int)(e & (unsigned)(WF_RING - 1));
    }
    return make_int2(n, slot);
}
DE_DEV int q_pop(WarpPool &p, uint32_t st, int want, int lane, int &slot) {
    int2 r = q_pop2(p, st, want, lane);
    slot = r.y;
    return r.x;
}

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
// Shading state of a slot: q0 = (throughput, radiance, NEE factor a, NEE factor b), q1 = (main direction, albedo), q2 = (normal | light
// direction, oceanness), q3 = (bathymetry).  WF_COLD: 64 B per slot in global memory, L2 only (.cg: written by one warp, read by another of the
// same CTA after a queue hand-over, which carries a __threadfence_block); otherwise the pool's arrays.
#if WF_COLD
DE_DEV float4 cold_ld(const Ctx &c, int slot, int q) { return __ldcg(c.cold + slot * 4 + q); }
DE_DEV void cold_st(const Ctx &c, int slot, int q, float4 v) { __stcg(c.cold + slot * 4 + q, v); }
DE_DEV void cold_st_thr_L(const Ctx &c, int slot, float thr, float L) { __stcg(reinterpret_cast<float2 *>(c.cold + slot * 4), make_float2(thr, L)); }
DE_DEV void cold_st_na(const Ctx &c, int slot, float na) { __stcg(reinterpret_cast<float *>(c.cold + slot * 4) + 2, na); }
DE_DEV float cold_ld_L(const Ctx &c, int slot) { return __ldcg(reinterpret_cast<const float *>(c.cold + slot * 4) + 1); }
#define WF_DSL(c, slot) (c).pool.dsl[slot]
#else
DE_DEV float4 cold_ld(const Ctx &c, int s, int q) {
    const WarpPool &p = c.pool;
    return q == 0 ? make_float4(p.thr[s], p.L[s], p.na[s], p.nb[s]) : q == 1 ? make_float4(p.mdx[s], p.mdy[s], p.mdz[s], p.m0[s])
         : q == 2 ? make_float4(p.nx[s], p.ny[s], p.nz[s], p.m1[s]) : make_float4(p.m2[s], 0.0f, 0.0f, 0.0f);
}
DE_DEV void cold_st(const Ctx &c, int s, int q, float4 v) {
    WarpPool &p = c.pool;
    if (q == 0) { p.thr[s] = v.x; p.L[s] = v.y; p.na[s] = v.z; p.nb[s] = v.w; }
    else if (q == 1) { p.mdx[s] = v.x; p.mdy[s] = v.y; p.mdz[s] = v.z; p.m0[s] = v.w; }
    else if (q == 2) { p.nx[s] = v.x; p.ny[s] = v.y; p.nz[s] = v.z; p.m1[s] = v.w; }
    else p.m2[s] = v.x;
}
DE_DEV void cold_st_thr_L(const Ctx &c, int s, float thr, float L) { c.pool.thr[s] = thr; c.pool.L[s] = L; }
DE_DEV void cold_st_na(const Ctx &c, int s, float na) { c.pool.na[s] = na; }
DE_DEV float cold_ld_L(const Ctx &c, int s) { return c.pool.L[s]; }
#define WF_DSL(c, slot) (*reinterpret_cast<uint32_t *>(&(c).pool.nb[slot]))
#endif
DE_DEV float cloud_ext_of(uint32_t sc) { return sc > 9u ? 0.02f : kCloudsExtinct; }  // pathtracer.py:351-352

// ------------------------------------------------------------------ transitions (run converged inside one-shot stages)
// outcome of sample_interaction (pathtracer.py:200-207) -> EVENT stage
DE_DEV uint32_t finish_interaction(const Ctx &c, int slot, uint32_t pk, uint32_t ev, float t, uint32_t id) {
    c.pool.t[slot] = t;
    pk = PK_SET_EV(pk, ev);
    pk = PK_SET_ID(pk, id);
    return PK_SET_STAGE(pk, ST_EVENT) & ~PK_RATIO;
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
        const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
        c.pool.t[slot] = t_start; c.pool.tmax[slot] = t_max;
#if WF_SHRINK
        c.pool.cmj[slot] = fminf(lr.max_ext_rmo, rmo_majorant_ool(lr.ext_r, lr.ext_m, lr.ext_o, o.x, o.y, o.z, d.x, d.y, d.z, t_start, t_max));
#else
        c.pool.cmj[slot] = fminf(lr.max_ext_rmo, rmo_segment_majorant(f3(lr.ext_r, lr.ext_m, lr.ext_o), o, d, t_start, t_max));  // local majorant
#endif
#if WF_RMO_BANDS
        {   // altitude band of the entry point and where the ray leaves it
            const float2 adv = rmo_band_advance(c.s, o.x, o.y, o.z, d.x, d.y, d.z, t_start, -1);
            c.pool.tlim[slot] = adv.x;
            c.pool.draw[slot] = (c.pool.draw[slot] & 0xFFFFFFu) | ((uint32_t)__float_as_int(adv.y) << 24);
        }
#endif
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
#if WF_OOL_MASK & 2
__device__ __noinline__
#else
DE_DEV
#endif
uint32_t setup_sdf(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d, uint32_t draw) {
#if WF_SHRINK
    store_draw(c, slot, draw, 0u);
    const float t0 = sdf_start_ool(o.x, o.y, o.z, d.x, d.y, d.z, c.s.land_height_scale);
    c.pool.t[slot] = t0;
    return PK_SET_STAGE(pk, t0 < 0.0f ? ST_SDF_DONE : ST_SDF);
#endif
    float ray_dist = 0.0f;
    float2 rd = rsi(o, d, kAtmosUpper);
    if (rd.x > 0.0f) ray_dist = rd.x;
    store_draw(c, slot, draw, 0u);
    if (land_surely_missed(o + d * ray_dist, d, ray_dist, c.s.land_height_scale)) {  // exact: the march would return -1
        c.pool.t[slot] = -1.0f;
        return PK_SET_STAGE(pk, ST_SDF_DONE);
    }
    c.pool.t[slot] = ray_dist + skip_to_terrain_top(o + d * ray_dist, d, ray_dist, c.s.land_height_scale);
    return PK_SET_STAGE(pk, ST_SDF);
}
// top of the scatter loop (pathtracer.py:349-359)
DE_DEV uint32_t begin_segment(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d) {
    pk &= ~(PK_SHADOW | PK_SURFACE | PK_RATIO | PK_VIS);
    return setup_sdf(c, slot, pk, o, d, 0u);
}

#if WF_FUSE_RMO
// First trip of an rmo pass (burst_track's loop body, IS_CLOUD = false, for a pass set up a moment ago): one Philox block, two collision
// candidates.  Returns the pk of ST_RMO_DONE when the pass ended, of ST_RMO (state stored for the loop stage) when it needs more trips.
template <bool COUNT> DE_DEV uint32_t rmo_first_trip(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d) {
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    const float3 ext = f3(lr.ext_r, lr.ext_m, lr.ext_o);
    const float inv_max = 1.0f / c.pool.cmj[slot], tmax = c.pool.tmax[slot];
    float t = c.pool.t[slot], T = 1.0f;
    const bool ratio = (pk & PK_RATIO) != 0u;
    const uint32_t blk = ((c.pool.draw[slot] & 0xFFFFFFu) + 3u) >> 2;
    const uint4 rb = philox_block(c.P.seed, c.pool.pix[slot], c.pool.sample[slot], PK_SC(pk) + 1u, blk);
    bool done = false;
    uint32_t ev = 0u, id = 0u, draw_after = 0u;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        t -= __logf(u32_to_unit(h ? rb.z : rb.x)) * inv_max;
        const float3 pos = o + d * t;
        if (t >= tmax) { done = true; draw_after = 4u * blk + 2u * h + 1u; break; }
        DE_COUNT(c.cn, C_RMO);
        const float3 dens = get_density(get_elevation(pos));
        const float es0 = ext.x * dens.x, es1 = ext.y * dens.y, es2 = ext.z * dens.z, sum = (es0 + es1) + es2;
        if (ratio) {
            T *= 1.0f - sum * inv_max;
            if (T < 1e-5f) { done = true; draw_after = 4u * blk + 2u * h + 2u; break; }
        } else {
            const float rand = u32_to_unit(h ? rb.w : rb.y);
            if (rand < sum * inv_max) {
                float cmf = es0;
                if (!(rand < cmf * inv_max)) {
                    id = 1u; cmf += es1;
                    if (!(rand < cmf * inv_max)) { id = 2u; cmf += es2; if (!(rand < cmf * inv_max)) id = 3u; }
                }
                ev = 1u; done = true; draw_after = 4u * blk + 2u * h + 3u;
                break;
            }
        }
    }
    if (!done) {  // the loop stage carries on with the next block
        c.pool.t[slot] = t;
        if (ratio) c.pool.aux[slot] = T;
        store_draw(c, slot, 4u * (blk + 1u), 0u);
        return pk;
    }
    store_draw(c, slot, draw_after, 0u);
    if (ratio) { c.pool.aux[slot] = T; return PK_SET_STAGE(pk, ST_RMO_DONE); }
    pk = PK_SET_RMO_EV(pk, ev);
    pk = PK_SET_RMO_ID(pk, id);
    c.pool.aux[slot] = t;
    if (ev) WF_DSL(c, slot) = draw_after - 1u;
    return PK_SET_STAGE(pk, ST_RMO_DONE);
}
#endif
template <bool COUNT> DE_DEV uint32_t rmo_done_body(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d);
// ST_SDF_DONE: what follows intersect_land -- main ray: sample_interaction; shadow ray: visibility +
// sample_transmittance (pathtracer.py:422-430).  t[slot] holds the intersection distance.  (WF_FUSE_RMO: the NEE ray of a scatter
// event arrives here as a shadow ray that found nothing, t = -1.)
template <bool COUNT> DE_DEV uint32_t stage_sdf_done(Ctx &c, int slot) {
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
#if WF_FUSE_RMO
    pk = setup_rmo(c, slot, pk, o, d, shadow);
    if (PK_STAGE(pk) == ST_RMO) pk = rmo_first_trip<COUNT>(c, slot, pk, o, d);
    if (PK_STAGE(pk) == ST_RMO_DONE) pk = rmo_done_body<COUNT>(c, slot, pk, o, d);
    return pk;
#else
    return setup_rmo(c, slot, pk, o, d, shadow);
#endif
}
// ST_RMO_DONE: between the rmo pass and the cloud pass of either tracker.  Delta tracking
// (sample_interaction, pathtracer.py:189-207) enters the shell only if the rmo pass found no collision
// before it; ratio tracking (sample_transmittance, :229-231) always does.  One copy of the cloud-bound
// code serves both (instruction cache).
template <bool COUNT> DE_DEV uint32_t rmo_done_body(const Ctx &c, int slot, uint32_t pk, float3 o, float3 d) {
    const bool ratio = (pk & PK_RATIO) != 0u;
    float ts, tm;
    intersect_cloud_limits(o, d, c.pool.isect[slot], ts, tm);
    const uint32_t rmo_ev = PK_RMO_EV(pk);
    const float rmo_t = c.pool.aux[slot];  // delta: distance of the rmo collision; ratio: transmittance so far
    bool enter = ts < tm;
    if (!ratio) enter = enter && (rmo_ev == kNullEvent || rmo_t > ts);
    if (enter) {
        const float bound = cloud_pass_setup(c.s, o, d, ts, tm);  // local majorant; the pass ends at the top of the local cloud layer
        if (bound > 0.0f) {
            c.pool.tmax[slot] = tm; c.pool.cmj[slot] = bound;
#if WF_FUSE_CLOUD
            {   // first trip of the pass (burst_track's loop body, IS_CLOUD = true)
                const float ext_cloud = cloud_ext_of(PK_SC(pk)), inv_max = 1.0f / (ext_cloud * bound);
                float t = ts, T = rmo_t;
                const uint32_t blk = ((c.pool.draw[slot] & 0xFFFFFFu) + 3u) >> 2;
                const uint4 rb = philox_block(c.P.seed, c.pool.pix[slot], c.pool.sample[slot], PK_SC(pk) + 1u, blk);
                bool done = false;
                uint32_t ev = 0u, draw_after = 0u;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    t -= __logf(u32_to_unit(h ? rb.z : rb.x)) * inv_max;
                    const float3 pos = o + d * t;
                    if (t >= tm) { done = true; draw_after = 4u * blk + 2u * h + 1u; break; }
                    DE_COUNT(c.cn, C_CLOUD);
                    float r2 = dot(pos, pos), inv_r = rsqrtf(r2), r = r2 * inv_r, dens = 0.0f;
                    if (r > kCloudsLower && r < kCloudsUpper) {
                        DE_COUNT(c.cn, C_TEX);
                        float hgt = (r - kCloudsLower) * (1.0f / kCloudsThickness);
                        float cl = sample_sphere_r8_inv(c.s.tex[3], pos, inv_r);
                        dens = (hgt - 0.2f < cl * 0.8f && 0.2f - hgt < cl * 0.2f) ? fmaxf(cl, 0.4f) : 0.0f;
                    }
                    const float sum = ext_cloud * (dens * kCloudsDensity);
                    if (ratio) {
                        T *= 1.0f - sum * inv_max;
                        if (T < 1e-5f) { done = true; draw_after = 4u * blk + 2u * h + 2u; break; }
                    } else if (u32_to_unit(h ? rb.w : rb.y) < sum * inv_max) { ev = 1u; done = true; draw_after = 4u * blk + 2u * h + 3u; break; }
                }
                if (done) {
                    store_draw(c, slot, draw_after, 0u);
                    if (ratio) { c.pool.aux[slot] = T; return PK_SET_STAGE(pk, ST_NEE_DONE); }
                    if (ev > 0u && (t < rmo_t || rmo_ev == 0u)) {
                        WF_DSL(c, slot) = draw_after - 1u;
                        return finish_interaction(c, slot, pk, 1u, t, kCloud);
                    }
                    return finish_interaction(c, slot, pk, rmo_ev, rmo_t, PK_RMO_ID(pk));
                }
                c.pool.t[slot] = t;
                if (ratio) c.pool.aux[slot] = T;
                store_draw(c, slot, 4u * (blk + 1u), 0u);
                return PK_SET_STAGE(pk, ST_CLOUD);
            }
#else
            c.pool.t[slot] = ts;
            return PK_SET_STAGE(pk, ST_CLOUD);  // PK_RATIO stays as it is
#endif
        }
    }
    if (ratio) return PK_SET_STAGE(pk, ST_NEE_DONE);
    return finish_interaction(c, slot, pk, rmo_ev, rmo_t, PK_RMO_ID(pk));
}
template <bool COUNT> DE_DEV uint32_t stage_rmo_done(Ctx &c, int slot) { return rmo_done_body<COUNT>(c, slot, c.pool.pk[slot], ld_o(c, slot), ld_d(c, slot)); }

// ------------------------------------------------------------------ path start / end
// ST_NEW (free slots): Renderer.render prologue for one sample (renderer.py:305-314).  Called with
// exactly 32 free slots, one per lane; claims one work chunk = 32 neighbouring pixels, one sample.
// Returns the new pk (ST_SDF), ST_NEW to hand the slot back (pixel outside the window), or ~0u
// when no work is left.
#if WF_OOL_MASK & 1
#define WF_ENDPATH_ATTR __device__ __noinline__
#else
#define WF_ENDPATH_ATTR DE_DEV
#endif
template <bool COUNT> WF_ENDPATH_ATTR uint32_t end_path(Ctx &c, int slot, uint32_t pk, bool primary_miss, float3 dir, float Lr);
template <bool COUNT> DE_DEV uint32_t stage_new(Ctx &c, int slot) {
    const unsigned full = 0xFFFFFFFFu;
    const WfParams &P = c.P;
    unsigned chunk = 0u;
    if (c.lane == 0) chunk = atomicAdd(P.next, 1u);
    chunk = __shfl_sync(full, chunk, 0);
    if (chunk >= c.pool.n_chunks) return ~0u;
    if (COUNT && P.timeline && c.lane == 0) atomicAdd(&c.pool.claimed, 1u);
    // chunk = (tile slot * n_spp + sample) * 4 + quarter; kept in chunk units: the path index itself exceeds 32 bits for
    // 4K x 4096 spp (3.4e10 paths).  The tile comes from the list of tiles that can see the planet (k_classify_tiles).
    unsigned in_tile = (chunk & 3u) * 32u + (unsigned)c.lane, ts = chunk >> 2;
    unsigned sp = ts % (unsigned)P.n_spp, tile = __ldg(P.tile_list + ts / (unsigned)P.n_spp);
    int px = P.x0 + (int)(tile % (unsigned)P.tiles_x) * kDeTileW + (int)(in_tile & 15u);
    int py = P.y0 + (int)(tile / (unsigned)P.tiles_x) * kDeTileH + (int)(in_tile >> 4);
    if (px >= P.x0 + P.w || py >= P.y0 + P.h) return ST_NEW;
    RngW rng;
    rng.key0 = P.seed; rng.key1 = (uint32_t)(py * c.s.W + px); rng.sample = P.first_sample + sp; rng.bounce = 0u; rng.draw = 0u; rng.valid = false;
    int bin = spectrum_bin(c.s.cdf, rng.next());
    float xu = rng.next(), xv = rng.next();
    float3 dir = get_cast_dir(c.s, c.dv, (float)px, (float)py, xu, xv);
    c.pool.pix[slot] = rng.key1; c.pool.sample[slot] = rng.sample;
    DE_COUNT(c.cn, C_SEGMENTS);
#if WF_SPACE_SHORTCUT
    // A primary ray that misses the atmosphere shell interacts with nothing (every later test of the
    // segment uses a smaller sphere): it is a primary miss right here (pathtracer.py:441-444,455-466),
    // instead of five queue hops.  rsi keeps the reference's NaN-on-miss behaviour, hence the negation.
    if (!(rsi(c.s.cam_pos, dir, kAtmosUpper).y >= 0.0f)) return end_path<COUNT>(c, slot, PK_SET_LAM(0u, (uint32_t)bin), true, dir, 0.0f);
#endif
    cold_st_thr_L(c, slot, 1.0f, 0.0f);
    st_o(c, slot, c.s.cam_pos); st_d(c, slot, dir);
    return begin_segment(c, slot, PK_SET_LAM(0u, (uint32_t)bin), c.s.cam_pos, dir);
}
// pathtracer.py:455-469 + renderer.py:329-330; frees the slot
template <bool COUNT> WF_ENDPATH_ATTR uint32_t end_path(Ctx &c, int slot, uint32_t pk, bool primary_miss, float3 dir, float Lr) {
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
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
        if (c.P.accum2) {  // per-pixel second moments for the image z-test; off in production
            float *a2 = c.P.accum2 + (size_t)c.pool.pix[slot] * 3;
            atomicAdd(a2, rgb.x * rgb.x); atomicAdd(a2 + 1, rgb.y * rgb.y); atomicAdd(a2 + 2, rgb.z * rgb.z);
        }
    }
    return ST_NEW;
}

// ------------------------------------------------------------------ loop stages
// Burst bookkeeping at the converged point of a loop trip.  Finished lanes keep (slot, next stage)
// in registers; only when enough lanes idle (or none is busy) are they pushed -- grouped per target
// queue, one reservation per group -- and the idle lanes refilled from this stage's own queue.
// Returns the active mask after the refill; lanes that received a slot have take_new = true.
#ifndef WF_FLUSH_OOL
#define WF_FLUSH_OOL 1  // the flush / refill path of a burst lives out of line, so the steady-state trip is compact code
#endif
// slow path of burst_sync: push the finished lanes, pop replacements.  Returns the slot this lane received, or -1.
#if WF_FLUSH_OOL
__device__ __noinline__
#else
DE_DEV
#endif
int burst_flush(WarpPool &pool, uint32_t st, unsigned am, bool pending, uint32_t pend_st, int pend_slot, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    q_push_sorted(pool, pending, pend_st, pend_slot, lane);
    int av = 0;
    if (lane == 0) av = *(volatile int *)&pool.q_avail[st];
    if (__shfl_sync(full, av, 0) <= 0) return -1;  // warp-uniform decision
    int got_slot;
    const int n = q_pop(pool, st, 32 - __popc(am), lane, got_slot);
    if (n == 0) return -1;
    const int rank = __popc(~am & ((1u << lane) - 1u));
    const int mine = __shfl_sync(full, got_slot, rank & 31);
    return (!((am >> lane) & 1u) && rank < n) ? mine : -1;
}
DE_DEV unsigned burst_sync(Ctx &c, uint32_t st, bool active, bool &pending, uint32_t pend_st, int pend_slot, int &slot, bool &take_new) {
    const unsigned full = 0xFFFFFFFFu;
    take_new = false;
    unsigned am = __ballot_sync(full, active);
    int idle = 32 - __popc(am);
    if (idle < WF_REFILL_MIN && am != 0u) return am;
    const int got = burst_flush(c.pool, st, am, pending, pend_st, pend_slot, c.lane);
    pending = false;
    if (got >= 0) { slot = got; take_new = true; }
    return am | __ballot_sync(full, take_new);
}

// SDF sphere tracing, one iteration per loop trip (pathtracer.py:37-44)
template <bool COUNT> DE_DEV void burst_sdf(Ctx &c, int slot) {
    const unsigned full = 0xFFFFFFFFu;
    bool active = slot >= 0;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    float t = 0.0f;
    uint32_t iter = 0u;
    auto load = [&]() { o = ld_o(c, slot); d = ld_d(c, slot); t = c.pool.t[slot]; iter = c.pool.draw[slot] >> 24; };
    if (active) load();
    int min_active = min(WF_MIN_ACTIVE, (__popc(__ballot_sync(full, active)) * WF_MIN_FRAC8 + 7) / 8);  // leave the burst when this few lanes are busy
#if WF_DRAIN_FAST
    if (!*(volatile int *)&c.pool.work_left) min_active = 1;  // draining: nothing to re-pack with
#endif
    const float scale = c.s.land_height_scale;
    bool pending = false;
    int pend_slot = -1;
    for (int it = 0; it < WF_BURST; ++it) {
        if (active) {
            float3 ro = o + d * t;
            DE_COUNT(c.cn, C_SDF); DE_COUNT(c.cn, C_TEX);
            float r2 = dot(ro, ro), inv_r = rsqrtf(r2);
            const bool gone = march_surely_missed(ro, d, r2, t, scale);  // exact: the march would run off to 10 R
            float dist = r2 * inv_r - kPlanetR - scale * sample_sphere_r8_inv(c.s.tex[1], ro, inv_r);
            t += dist;
            ++iter;
            if (gone || t > 63710000.0f || fabsf(dist) < t * 0.0001f || iter >= 250u) {
                c.pool.t[slot] = (t < 63710000.0f && !gone) ? t : -1.0f;
                c.pool.pk[slot] = PK_SET_STAGE(c.pool.pk[slot], ST_SDF_DONE);
                active = false; pending = true; pend_slot = slot;
            }
        }
        bool take;
        unsigned am = burst_sync(c, ST_SDF, active, pending, ST_SDF_DONE, pend_slot, slot, take);
        if (take) { load(); active = true; }
        if (__popc(am) < min_active) break;
    }
    q_push_sorted(c.pool, pending, ST_SDF_DONE, pend_slot, c.lane);
    unsigned am = __ballot_sync(full, active);
    if (active) {
        c.pool.t[slot] = t; c.pool.draw[slot] = (c.pool.draw[slot] & 0xFFFFFFu) | (iter << 24);
        q_push_group(c.pool, ST_SDF, slot, am, c.lane);
    }
}

// delta / ratio tracking through Rayleigh+Mie+ozone (IS_CLOUD=false) or the cloud shell (true)
// (pathtracer.py:91-112,130-141).  One loop trip = ONE Philox block = TWO collision candidates
// (words 0,1 and 2,3: free flight + acceptance test; a ratio step leaves its second word unused),
// so the RNG is issued converged with static word selection.  A real collision only records the
// slot of its scatter/absorb draw (pathtracer.py:270); ST_EVENT evaluates it for the winner.
#ifndef WF_TRACK_TEMPLATE
#define WF_TRACK_TEMPLATE 0
#endif
#if WF_TRACK_TEMPLATE
template <bool COUNT, bool IS_CLOUD> DE_DEV void burst_track(Ctx &c, int slot) {
#else
template <bool COUNT> DE_DEV void burst_track(Ctx &c, int slot, const bool IS_CLOUD) {  // IS_CLOUD is warp-uniform: one copy of the loop
#endif
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t ST_SELF = IS_CLOUD ? ST_CLOUD : ST_RMO;
    bool active = slot >= 0;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1), ext = f3(0, 0, 0);
    float t = 0.0f, tmax = 0.0f, T = 1.0f, max_ext = 1.0f, inv_max = 1.0f, ext_cloud = 0.0f;
    float tlim = 3.0e38f;   // rmo: exit of the current altitude band (cloud passes have none)
    int band = 0;
    uint32_t pk = 0u, blk = 0u, key1 = 0u, smp = 0u;
    auto load = [&]() {
        d = ld_d(c, slot);
        t = c.pool.t[slot]; tmax = c.pool.tmax[slot]; T = c.pool.aux[slot];
        pk = c.pool.pk[slot];
        key1 = c.pool.pix[slot]; smp = c.pool.sample[slot];
        const uint32_t dw = c.pool.draw[slot];
        blk = ((dw & 0xFFFFFFu) + 3u) >> 2;  // passes start on a block boundary; trips end on one
        o = ld_o(c, slot);
        if (IS_CLOUD) { ext_cloud = cloud_ext_of(PK_SC(pk)); max_ext = ext_cloud * c.pool.cmj[slot]; tlim = 3.0e38f; }
        else {
            const LambdaRow &lr = c.s.lam[PK_LAM(pk)]; ext = f3(lr.ext_r, lr.ext_m, lr.ext_o); max_ext = c.pool.cmj[slot];
#if WF_RMO_BANDS
            band = (int)(dw >> 24); tlim = c.pool.tlim[slot];
            max_ext = fminf(max_ext, rmo_band_majorant(c.s, ext, band));
#endif
        }
        inv_max = 1.0f / max_ext;
    };
    if (active) load();
    int min_active = min(WF_MIN_ACTIVE, (__popc(__ballot_sync(full, active)) * WF_MIN_FRAC8 + 7) / 8);  // leave the burst when this few lanes are busy
#if WF_DRAIN_FAST
    if (!*(volatile int *)&c.pool.work_left) min_active = 1;
#endif
    bool pending = false;
    int pend_slot = -1;
    uint32_t pend_st = ST_RMO_DONE;
    for (int it = 0; it < WF_BURST / 2; ++it) {
        if (active) {
            const bool ratio = (pk & PK_RATIO) != 0u;
#if WF_TRACK_PHILOX_INLINE
            const uint4 rb = philox_block_inl(c.P.seed, key1, smp, PK_SC(pk) + 1u, blk);
#else
            const uint4 rb = philox_block(c.P.seed, key1, smp, PK_SC(pk) + 1u, blk);
#endif
            bool done = false;
            uint32_t ev = 0u, id = IS_CLOUD ? 3u : 0u, draw_after = 0u;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
#if WF_RMO_BANDS
                if (!IS_CLOUD) {  // free flight through the altitude bands (de_device.cuh: rmo_band_walk)
                    RmoWalk w{t, tlim, max_ext, band};
                    t = rmo_band_walk(c.s, ext, c.pool.cmj[slot], tmax, -__logf(u32_to_unit(h ? rb.z : rb.x)), w,
                                      [&](float tt, int kk) { return rmo_band_advance(c.s, o.x, o.y, o.z, d.x, d.y, d.z, tt, kk); });
                    tlim = w.tlim; band = w.band;
                    if (w.max_ext != max_ext) { max_ext = w.max_ext; inv_max = 1.0f / max_ext; }
                } else t -= __logf(u32_to_unit(h ? rb.z : rb.x)) * inv_max;
#else
                t -= __logf(u32_to_unit(h ? rb.z : rb.x)) * inv_max;
#endif
                const float3 pos = o + d * t;  // from the origin every time: independent of where bursts were cut
                if (t >= tmax) { done = true; draw_after = 4u * blk + 2u * h + 1u; break; }
                float es0 = 0.0f, es1 = 0.0f, es2 = 0.0f, sum;
                if (IS_CLOUD) {  // get_clouds_density (pathtracer.py:48-65) with one shared rsqrt
                    DE_COUNT(c.cn, C_CLOUD);
                    float r2 = dot(pos, pos), inv_r = rsqrtf(r2), r = r2 * inv_r, dens = 0.0f;
                    if (r > kCloudsLower && r < kCloudsUpper) {
                        DE_COUNT(c.cn, C_TEX);
                        float hgt = (r - kCloudsLower) * (1.0f / kCloudsThickness);
                        float cl = sample_sphere_r8_inv(c.s.tex[3], pos, inv_r);
                        dens = (hgt - 0.2f < cl * 0.8f && 0.2f - hgt < cl * 0.2f) ? fmaxf(cl, 0.4f) : 0.0f;
                    }
                    sum = ext_cloud * (dens * kCloudsDensity);
                } else {
                    DE_COUNT(c.cn, C_RMO);
                    float3 dens = get_density(get_elevation(pos));
                    es0 = ext.x * dens.x; es1 = ext.y * dens.y; es2 = ext.z * dens.z;
                    sum = (es0 + es1) + es2;
                }
                if (ratio) {
                    T *= 1.0f - sum * inv_max;
                    if (T < 1e-5f) { done = true; draw_after = 4u * blk + 2u * h + 2u; break; }
                } else {
                    float rand = u32_to_unit(h ? rb.w : rb.y);
                    if (rand < sum * inv_max) {
                        if (!IS_CLOUD) {
                            float cmf = es0;
                            if (!(rand < cmf * inv_max)) {
                                id = 1u; cmf += es1;
                                if (!(rand < cmf * inv_max)) { id = 2u; cmf += es2; if (!(rand < cmf * inv_max)) id = 3u; }
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
                        WF_DSL(c, slot) = draw_after - 1u;
                        npk = finish_interaction(c, slot, pk, 1u, t, kCloud);
                    } else npk = finish_interaction(c, slot, pk, rmo_ev, rmo_t, PK_RMO_ID(pk));
                } else {
                    npk = PK_SET_RMO_EV(pk, ev);
                    npk = PK_SET_RMO_ID(npk, id);
                    c.pool.aux[slot] = t;
                    if (ev) WF_DSL(c, slot) = draw_after - 1u;
                    npk = PK_SET_STAGE(npk, ST_RMO_DONE);
                }
                c.pool.pk[slot] = npk;
                pend_st = PK_STAGE(npk);  // RMO_DONE, or EVENT (delta) / NEE_DONE (ratio) after the cloud pass
                active = false; pending = true; pend_slot = slot;
            }
        }
        bool take;
        unsigned am = burst_sync(c, ST_SELF, active, pending, pend_st, pend_slot, slot, take);
        if (take) { load(); active = true; }
        if (__popc(am) < min_active) break;
    }
    q_push_sorted(c.pool, pending, pend_st, pend_slot, c.lane);
    unsigned am = __ballot_sync(full, active);
    if (active) {
        c.pool.t[slot] = t; if (pk & PK_RATIO) c.pool.aux[slot] = T; store_draw(c, slot, 4u * blk, (uint32_t)band);
#if WF_RMO_BANDS
        if (!IS_CLOUD) c.pool.tlim[slot] = tlim;
#endif
        q_push_group(c.pool, ST_SELF, slot, am, c.lane);
    }
}

// ST_EVENT: after sample_interaction, pathtracer.py:369-444 up to the point where the NEE ray is traced
template <bool COUNT> DE_DEV uint32_t stage_event(Ctx &c, int slot) {
    uint32_t pk = c.pool.pk[slot];
    const uint32_t sc = PK_SC(pk);
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    float3 o = ld_o(c, slot), d = ld_d(c, slot);
    RngW rng = load_rng(c, slot, pk);
    uint32_t ev = PK_EV(pk), id = PK_ID(pk);
    if (ev) {  // a real collision: scatter or absorb (pathtracer.py:108-111,263-270) from the recorded slot
        uint32_t ds = WF_DSL(c, slot);
        uint4 b = philox_block(rng.key0, rng.key1, rng.sample, rng.bounce, ds >> 2);
        uint32_t w = (ds & 3u) == 0u ? b.x : ((ds & 3u) == 1u ? b.y : ((ds & 3u) == 2u ? b.z : b.w));
        float albedo = id == 0u ? 1.0f : (id == 1u ? 0.95f : (id == 2u ? 0.0f : 0.99f));
        ev = u32_to_unit(w) < albedo ? (uint32_t)kScatterEvent : (uint32_t)kAbsorbEvent;
    }
    if (sc > 9u && id == (uint32_t)kCloud) id = kIsoCloud;
    rng.align();
    float3 light_dir = sample_cone_oriented(c.dv.sun_cos_angle, c.dv.light_dir, rng);
    const float earth_isect = c.pool.isect[slot];
    const bool absorbed = ev == (uint32_t)kAbsorbEvent;
    if (absorbed || (ev != (uint32_t)kScatterEvent && !(earth_isect > 0.0f)))  // absorbed, or escaped (pathtracer.py:441-444)
        return end_path<COUNT>(c, slot, pk, !absorbed && sc == 0u, d, cold_ld_L(c, slot));
    if (ev == (uint32_t)kScatterEvent) {
        float3 ipos = o + d * c.pool.t[slot];
        bool blocked = rsi(ipos, light_dir, kPlanetR).y > 0.0f;
        cold_st_na(c, slot, phase_eval_ool(d.x, d.y, d.z, light_dir.x, light_dir.y, light_dir.z, (int)id, sc > 0u));
        cold_st(c, slot, 1, make_float4(d.x, d.y, d.z, 0.0f));
        st_o(c, slot, ipos); st_d(c, slot, light_dir);
        pk = PK_SET_ID(pk, id) & ~PK_SURFACE;
        store_draw(c, slot, rng.draw, 0u);
        if (blocked) { c.pool.aux[slot] = 0.0f; return PK_SET_STAGE(pk, ST_NEE_DONE); }
#if WF_FUSE_RMO
        c.pool.t[slot] = -1.0f;  // "a shadow ray that found nothing": ST_SDF_DONE sets isect = -1, ratio tracking, and runs the pass's first trip
        return PK_SET_STAGE(pk | PK_SHADOW, ST_SDF_DONE);
#else
        c.pool.isect[slot] = -1.0f;
        return setup_rmo(c, slot, pk, ipos, light_dir, true);
#endif
    }
    // surface hit: shading is a long body of its own (four height fetches, four material maps, the BRDF) that a
    // fifth of the events need -- it runs as stage ST_SURFACE over a full group instead of a few lanes of this one
    cold_st(c, slot, 2, make_float4(light_dir.x, light_dir.y, light_dir.z, 0.0f));
    store_draw(c, slot, rng.draw, 0u);
    return PK_SET_STAGE(pk, ST_SURFACE);
}

// ST_SURFACE: pathtracer.py:405-439 up to the shadow ray -- normal, material, emission, BRDF towards the light
template <bool COUNT> DE_DEV uint32_t stage_surface(Ctx &c, int slot) {
    uint32_t pk = c.pool.pk[slot];
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    const float3 o = ld_o(c, slot), d = ld_d(c, slot);
    const float4 q2 = cold_ld(c, slot, 2), q0 = cold_ld(c, slot, 0);
    const float3 light_dir = f3(q2.x, q2.y, q2.z);
    const float earth_isect = c.pool.isect[slot];
    const uint32_t draw = c.pool.draw[slot] & 0xFFFFFFu;
    {
        DE_COUNT(c.cn, C_SURF);
        float3 land_pos = o + d * earth_isect;
        // land_normal (pathtracer.py:16-25) and get_land_material (:284-313) on the shared fetch routine
        const float hs = c.s.land_height_scale, eps = c.dv.normal_eps;
        if (COUNT) { c.cn.v[C_SDF] += 4; c.cn.v[C_TEX] += 8; }
        float sd0 = length(land_pos) - kPlanetR - hs * r8_ool(c.s.tex[1], land_pos);
        float3 px_ = f3(land_pos.x - eps, land_pos.y, land_pos.z), py_ = f3(land_pos.x, land_pos.y - eps, land_pos.z), pz_ = f3(land_pos.x, land_pos.y, land_pos.z - eps);
        float3 nrm = normalize(f3(sd0 - (length(px_) - kPlanetR - hs * r8_ool(c.s.tex[1], px_)), sd0 - (length(py_) - kPlanetR - hs * r8_ool(c.s.tex[1], py_)),
                                  sd0 - (length(pz_) - kPlanetR - hs * r8_ool(c.s.tex[1], pz_))));
        LandMaterial m;  // one equirect mapping for the four material maps
        const float2 muv = sphere_uv_ool(land_pos.x, land_pos.y, land_pos.z);
        m.ocean = tex_r8(c.s.tex[2], muv.x, muv.y);
        m.albedo_srgb = grade_albedo(tex_rgb8(c.s.tex[0], muv.x, muv.y), m.ocean);
        m.bathymetry = tex_r8(c.s.tex[4], muv.x, muv.y);
        m.emissive = tex_r8(c.s.tex[5], muv.x, muv.y);
        float albedo = lr.s2s_valid != 0.0f ? dot(m.albedo_srgb, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
        const float L_new = q0.y + q0.x * m.emissive * lr.nightlights_power;
        float3 offset_pos = land_pos * (1.0f + 0.0001f * c.s.land_height_scale / 12000.0f);
        float2 bn = brdf_ool(albedo, m.ocean, m.bathymetry, -d.x, -d.y, -d.z, nrm.x, nrm.y, nrm.z, light_dir.x, light_dir.y, light_dir.z);
        cold_st(c, slot, 0, make_float4(q0.x, L_new, bn.x, bn.y));
        cold_st(c, slot, 1, make_float4(d.x, d.y, d.z, albedo));
        cold_st(c, slot, 2, make_float4(nrm.x, nrm.y, nrm.z, m.ocean));
        cold_st(c, slot, 3, make_float4(m.bathymetry, 0.0f, 0.0f, 0.0f));
        st_o(c, slot, offset_pos); st_d(c, slot, light_dir);
        pk |= PK_SURFACE | PK_SHADOW;
        return setup_sdf(c, slot, pk, offset_pos, light_dir, draw);
    }
}

// ST_NEE_DONE: after the NEE transmittance, pathtracer.py:394-401 / 431-439, Russian roulette :447-453, next segment
template <bool COUNT> DE_DEV uint32_t stage_nee_done(Ctx &c, int slot) {
    uint32_t pk = c.pool.pk[slot];
    uint32_t sc = PK_SC(pk);
    const LambdaRow &lr = c.s.lam[PK_LAM(pk)];
    RngW rng = load_rng(c, slot, pk);
    float3 o = ld_o(c, slot);  // interaction position / offset position
    const float4 q0 = cold_ld(c, slot, 0), q1 = cold_ld(c, slot, 1);
    float3 main_d = f3(q1.x, q1.y, q1.z);
    float T = c.pool.aux[slot], thr = q0.x, Lacc = q0.y;
    float3 nd;
    if (pk & PK_SURFACE) {
        float vis = (pk & PK_VIS) ? 1.0f : 0.0f;
        const float4 q2 = cold_ld(c, slot, 2), q3 = cold_ld(c, slot, 3);
        Lacc += thr * T * vis * lr.sun_irradiance * q0.z * q0.w;
        float3 nrm = f3(q2.x, q2.y, q2.z);
        rng.align();
        nd = sample_hemisphere_cosine_weighted(nrm, rng);
        float brdf = brdf_ool(q1.w, q2.w, q3.x, -main_d.x, -main_d.y, -main_d.z, nrm.x, nrm.y, nrm.z, nd.x, nd.y, nd.z).x;
        thr *= brdf * kPi;
    } else {
        Lacc += thr * T * lr.sun_irradiance * q0.z;
        rng.align();
        uint4 blk = philox_block(rng.key0, rng.key1, rng.sample, rng.bounce, rng.draw >> 2);
        int used = 0;
        float4 sp = phase_sample_ool(main_d.x, main_d.y, main_d.z, (int)PK_ID(pk), sc > 0u, blk, &used);
        nd = f3(sp.x, sp.y, sp.z);
        thr *= sp.w;
        rng.b0 = blk.x; rng.b1 = blk.y; rng.b2 = blk.z; rng.b3 = blk.w; rng.valid = true; rng.draw += (uint32_t)used;  // the roulette draw follows in the same block
    }
    bool terminate = false;
    if (sc > 3u) {
        float p = fmaxf(0.05f, 1.0f - thr);
        if (rng.next() < p) terminate = true;
        else thr /= 1.0f - p;
    }
    ++sc;
    if (terminate || sc >= 25u) return end_path<COUNT>(c, slot, pk, false, nd, Lacc);
    cold_st_thr_L(c, slot, thr, Lacc);
    st_d(c, slot, nd);
    pk = PK_SET_SC(pk, sc);
    DE_COUNT(c.cn, C_SEGMENTS);
    return begin_segment(c, slot, pk, o, nd);
}

// ------------------------------------------------------------------ the persistent kernel
template <bool COUNT> __global__ void __launch_bounds__(WF_WARPS * 32, 1) k_render_wavefront(const __grid_constant__ DevScene s, const __grid_constant__ WfParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpPool &pool = *reinterpret_cast<WarpPool *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xFFFFFFFFu;
    const DevDerived dv = *s.derived;
    Counters cn;
    cn.clear();
    Ctx c{s, dv, P, pool, cn, lane, P.cold ? P.cold + (size_t)blockIdx.x * WF_SLOTS * 4 : nullptr};
    // all slots start free (queue ST_NEW); ring entries carry lap tag 31 until first written
    for (int k = threadIdx.x; k < ST_COUNT * WF_RING; k += blockDim.x) (&pool.ring[0][0])[k] = 0xFFFFu;
    __syncthreads();
    for (int k = threadIdx.x; k < WF_SLOTS; k += blockDim.x) pool.ring[ST_NEW][k] = (uint16_t)k;
    if (threadIdx.x < ST_COUNT) {
        pool.q_tail[threadIdx.x] = threadIdx.x == ST_NEW ? WF_SLOTS : 0u;
        pool.q_head[threadIdx.x] = 0u;
        pool.q_avail[threadIdx.x] = threadIdx.x == ST_NEW ? WF_SLOTS : 0;
    }
    if (threadIdx.x == 0) {
        pool.retired = 0; pool.work_left = 1; pool.phase = 0; pool.last_visit = -1; pool.claimed = 0u; pool.t_exhaust = 0ull; pool.t_few = 0ull;
        for (int k = 0; k < (int)ST_COUNT; ++k) { pool.dr_visits[k] = 0u; pool.dr_slots[k] = 0u; }
        pool.n_chunks = __ldg(P.n_tiles) * (unsigned)P.n_spp * 4u;
        if (COUNT && P.timeline) atomicMin(&P.timeline[0], globaltimer_ns());
    }
    __syncthreads();
    int last_st = -1;
#if WF_IDLE_EXP
    int idle_streak = 0;
#endif
    bool chained = false;  // the warp already holds slots of stage `st` (handed over by the previous one-shot stage)
    uint32_t st = 0u;
    int slot = -1, n = 0;
    bool work_left = true;
    for (;;) {
        if (!chained) {
            // 1. pick the fullest stage queue (free slots only count in whole chunks of 32 while work remains)
            int wl = 0;
            if (lane == 0) wl = *(volatile int *)&pool.work_left | (*(volatile int *)&pool.retired >= WF_SLOTS ? 2 : 0) | (*(volatile int *)&pool.phase << 8);
            wl = __shfl_sync(full, wl, 0);  // warp-uniform snapshot
            const int phase = wl >> 8;
            work_left = (wl & 1) != 0;
            int av = lane < (int)ST_COUNT ? *(volatile int *)&pool.q_avail[lane] : 0;
            if (lane == (int)ST_NEW && work_left && av < 32) av = 0;
            // plain fullest-queue policy: stage affinity per sub-partition and role-specialised warps were
            // both measured slower (profiles/r1_wavefront.md, "scheduling experiments")
            int key = av > 0 ? (av << 4) | lane : 0;
#if WF_CYCLIC
            {   // stage -> position in the dataflow cycle (nibble table), distance from the phase stage; full groups rank by that distance
                const unsigned long long kOrd = 0x786425310ULL;
                int dist = (int)((kOrd >> (4 * (lane < (int)ST_COUNT ? lane : 0))) & 15ull) - (int)((kOrd >> (4 * phase)) & 15ull);
                if (dist < 0) dist += (int)ST_COUNT;
                if (av >= 32) key = (1 << 20) | (((int)ST_COUNT - dist) << 16) | lane;
            }
#else
#if WF_STICKY
            if (lane == last_st && av >= WF_STICKY) key += 1 << 20;  // stay on the stage whose code is warm while it has a full group
#endif
#if WF_PHASE
            if (lane == phase && av >= WF_PHASE) key += 1 << 21;     // SM-wide phase: everybody on the same body while it lasts
#endif
#endif
            key = __reduce_max_sync(full, key);
            if (key == 0) {
                if (wl & 2) break;  // every slot found the work counter exhausted
                long long t0i = COUNT ? clock64() : 0;
#if WF_IDLE_EXP
                __nanosleep(64u << min(idle_streak, WF_IDLE_EXP));
                ++idle_streak;
#else
                __nanosleep(WF_BACKOFF_NS > 64 ? WF_BACKOFF_NS : 64);
#endif
                if (COUNT && lane == 0 && P.prof) atomicAdd(&P.prof[3 * ST_COUNT], (unsigned long long)(clock64() - t0i));
                continue;
            }
            st = (uint32_t)(key & 15);
            last_st = (int)st;
#if WF_IDLE_EXP
            idle_streak = 0;
#endif
#if WF_PHASE
            if ((int)st != phase && lane == 0) pool.phase = (int)st;  // the phase stage ran low: whoever notices moves the SM on
#endif
            // 2. take up to 32 ready slots
            n = q_pop(pool, st, 32, lane, slot);
            if (n == 0) {
#if WF_BACKOFF_NS
                __nanosleep(WF_BACKOFF_NS);
#endif
                continue;
            }
        }
        chained = false;
        const long long t0s = COUNT ? clock64() : 0;
        const uint32_t st_run = st;
        const int n_run = n;
        // 3. run the stage
        if (st == ST_SDF) burst_sdf<COUNT>(c, slot);
#if WF_TRACK_TEMPLATE
        else if (st == ST_RMO) burst_track<COUNT, false>(c, slot);
        else if (st == ST_CLOUD) burst_track<COUNT, true>(c, slot);
#else
        else if (st == ST_RMO || st == ST_CLOUD) burst_track<COUNT>(c, slot, st == ST_CLOUD);
#endif
        else {
            uint32_t npk = 0u;
            bool has = slot >= 0;
            if (st == ST_NEW) {
                if (!work_left) {
                    if (lane == 0) {
                        const int before = atomicAdd(&pool.retired, n);
                        if (COUNT && P.timeline && before < WF_SLOTS - 64 && before + n >= WF_SLOTS - 64) pool.t_few = globaltimer_ns();
                    }
                    continue;
                }
                if (n < 32) { q_push_sorted(pool, has, ST_NEW, slot, lane); continue; }  // lost a race for a whole chunk
                npk = stage_new<COUNT>(c, slot);
                if (npk == ~0u) {  // work counter exhausted
                    if (lane == 0) {
                        if (COUNT && P.timeline && atomicExch(&pool.work_left, 0) != 0) {
                            const unsigned long long tn = globaltimer_ns();
                            atomicMin(&P.timeline[1], tn); atomicMax(&P.timeline[2], tn);
                            pool.t_exhaust = tn;
                        }
                        pool.work_left = 0; atomicAdd(&pool.retired, 32);
                    }
                    continue;
                }
            } else if (has) {
                if (st == ST_SDF_DONE) npk = stage_sdf_done<COUNT>(c, slot);
                else if (st == ST_RMO_DONE) npk = stage_rmo_done<COUNT>(c, slot);
                else if (st == ST_EVENT) npk = stage_event<COUNT>(c, slot);
                else if (st == ST_SURFACE) npk = stage_surface<COUNT>(c, slot);
                else npk = stage_nee_done<COUNT>(c, slot);
            }
            if (has) c.pool.pk[slot] = npk;
            const uint32_t tgt = PK_STAGE(npk);
#if WF_CHAIN || WF_DRAIN_FAST
            // hand the largest group of successors straight to its stage: no queue round trip for them
            const unsigned grp = __match_any_sync(full, has ? tgt : 15u);
            const unsigned best = __reduce_max_sync(full, has ? ((unsigned)__popc(grp) << 4) | tgt : 0u);
            const uint32_t cst = best & 15u;
            const int cn_ = (int)(best >> 4);
#if WF_CHAIN
            const int chain_min = WF_CHAIN;
#else
            const int chain_min = work_left ? 64 : 1;  // only while draining
#endif
            if (cn_ >= chain_min && cst != ST_NEW) {
                const bool keep = has && tgt == cst;
                q_push_sorted(pool, has && !keep, tgt, slot, lane);
                if (!keep) slot = -1;
                n = cn_;
                if (WF_CHAIN && 32 - cn_ >= WF_CHAIN_TOPUP && cst != ST_NEW) {  // fill the idle lanes from the successor's queue
                    int av2 = 0;
                    if (lane == 0) av2 = *(volatile int *)&pool.q_avail[cst];
                    if (__shfl_sync(full, av2, 0) > 0) {
                        int got;
                        const int m = q_pop(pool, cst, 32 - cn_, lane, got);
                        const unsigned km = __ballot_sync(full, keep);
                        const int rank = __popc(~km & ((1u << lane) - 1u));
                        const int mine = __shfl_sync(full, got, rank & 31);
                        if (!keep && rank < m) slot = mine;
                        n += m;
                    }
                }
                st = cst;
                chained = true;
            } else
#endif
            q_push_sorted(pool, has, tgt, slot, lane);
        }
        if (COUNT && P.timeline && lane == 0 && !work_left) { atomicAdd(&pool.dr_visits[st_run], 1u); atomicAdd(&pool.dr_slots[st_run], (unsigned)n_run); }
        if (COUNT && lane == 0 && P.prof) {
            if (atomicExch(&pool.last_visit, (int)st_run) != (int)st_run) atomicAdd(&P.prof[3 * ST_COUNT + 1], 1ull);  // the SM changed stage body
            atomicAdd(&P.prof[3 * st_run], (unsigned long long)(clock64() - t0s));
            atomicAdd(&P.prof[3 * st_run + 1], 1ull);
            atomicAdd(&P.prof[3 * st_run + 2], (unsigned long long)n_run);
        }
    }
    if (COUNT) cn.flush(s.counters);
    if (COUNT && P.timeline) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long tn = globaltimer_ns();
            atomicMin(&P.timeline[3], tn); atomicMax(&P.timeline[4], tn);
            atomicMin(&P.timeline[5], (unsigned long long)pool.claimed); atomicMax(&P.timeline[6], (unsigned long long)pool.claimed);
            if (P.cta_stats) {
                unsigned long long *o = P.cta_stats + (size_t)blockIdx.x * 24;
                o[0] = pool.t_exhaust; o[1] = pool.t_few; o[2] = tn; o[3] = pool.claimed;
                for (int k = 0; k < (int)ST_COUNT; ++k) { o[4 + k] = pool.dr_visits[k]; o[4 + ST_COUNT + k] = pool.dr_slots[k]; }
            }
        }
    }
}



// ------------------------------------------------------------------ film tiles that cannot see the planet
// Apollo 11 (BASELINE configs[1]): 63 % of the samples are primary rays that miss the atmosphere shell.  They interact with
// nothing (pathtracer.py:441-444,455-466: sun disc + star map), so a tile ALL of whose jittered primary rays miss the
// 6 481 km shell needs no path state, no queues and no scheduler: k_space_tiles renders its samples one thread per pixel,
// fully converged, and the persistent kernel only receives the list of the remaining tiles.
//
// Classification (f64, one thread per tile, ordered compaction so the work order is deterministic): the rays of a tile are
// dir(px, py) over the rectangle [x, x+16] x [y, y+8] of continuous film coordinates (pixel + jitter in [0,1), renderer.py:
// 269-279).  Seen from a camera outside the shell the hit directions form the cap of half-angle asin(R_atm / |cam|) around
// -cam.  All rays of the tile lie within alpha = max corner angle of the tile-centre direction (the sub-level sets of the
// angular distance are convex in the film plane, so its maximum over a rectangle sits on a corner); the tile is space iff
// angle(centre, -cam) > cap + alpha + 2e-5 rad.  The margin is ~30x the f32 rounding of get_cast_dir / rsi at |cam| = 5.7e7 m
// (where 2e-5 rad = 1.1 km): every sample of a space tile is also a miss of the reference's own f32 test, so the samples
// are the same paths with the same values as the generic route (tests/test_gpu_render.py compares the two bit patterns).
struct D3 { double x, y, z; };
DE_DEV D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
DE_DEV double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DE_DEV D3 dcross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DE_DEV double dangle(D3 a, D3 b) { D3 c = dcross(a, b); return atan2(sqrt(ddot(c, c)), ddot(a, b)); }
DE_DEV D3 film_dir(const DevScene &s, const DevDerived &dv, double px, double py) {  // get_cast_dir before normalisation
    const double fov = s.fov;
    const double fu = (2.0 * fov * px / (double)s.H - fov * (double)s.aspect_ratio - 1e-5) * (double)s.aspect_scale;
    const double fv = 2.0 * fov * py / (double)s.H - fov - 1e-5;
    return d3(dv.cam_d.x + fu * dv.cam_du.x + fv * dv.cam_dv.x, dv.cam_d.y + fu * dv.cam_du.y + fv * dv.cam_dv.y,
              dv.cam_d.z + fu * dv.cam_du.z + fv * dv.cam_dv.z);
}
DE_DEV bool tile_is_space(const DevScene &s, const DevDerived &dv, int x0, int y0, int x1, int y1) {
    const D3 cam = d3(s.cam_pos.x, s.cam_pos.y, s.cam_pos.z);
    const double dist = sqrt(ddot(cam, cam));
    if (!(dist > (double)kAtmosUpper * 1.0001)) return false;  // camera inside (or grazing) the shell: every ray is in the medium
    const double cap = asin((double)kAtmosUpper / dist);
    const D3 axis = d3(-cam.x, -cam.y, -cam.z);
    const D3 c = film_dir(s, dv, 0.5 * (x0 + x1), 0.5 * (y0 + y1));
    double alpha = dangle(c, film_dir(s, dv, x0, y0));
    alpha = fmax(alpha, dangle(c, film_dir(s, dv, x1, y0)));
    alpha = fmax(alpha, dangle(c, film_dir(s, dv, x0, y1)));
    alpha = fmax(alpha, dangle(c, film_dir(s, dv, x1, y1)));
    const double phi = dangle(c, axis);
    return phi > cap + alpha + 2e-5;  // NaN (degenerate camera) compares false: generic route
}
// Can the paths of this tile get long?  The tail of a launch is its longest path (profiles/r2_tail.md): 15-25 segments of multiple scattering in
// cloud, a serial chain of ~500 stage visits.  A tile all of whose corner rays meet the cloud-top sphere at more than ~20 degrees above the horizon and
// whose lat-long footprint (padded by one cell) is empty in the dilated cloud max-map starts only clear-air / ground paths (a handful of segments).
// f32 and approximate on purpose: this only ORDERS the work (clear tiles last), no sample depends on it.
DE_DEV bool tile_is_clear(const DevScene &s, const DevDerived &dv, int x0, int y0, int x1, int y1) {
    if (!s.cloud_max) return false;
    float ulo = 2.0f, uhi = -1.0f, vlo = 2.0f, vhi = -1.0f;
    for (int k = 0; k < 5; ++k) {
        const float fx = k == 4 ? 0.5f * (float)(x0 + x1) : (float)((k & 1) ? x1 : x0), fy = k == 4 ? 0.5f * (float)(y0 + y1) : (float)((k & 2) ? y1 : y0);
        const float3 d = get_cast_dir(s, dv, fx, fy, 0.0f, 0.0f);
        const float2 hit = rsi(s.cam_pos, d, kCloudsUpper);
        if (!(hit.x > 0.0f)) return false;                       // misses the shell, or the camera is inside it
        const float3 p = s.cam_pos + d * hit.x;
        if (!(-dot(p, d) > 0.35f * kCloudsUpper)) return false;  // grazing: long slant paths, the footprint test means little
        const float2 uv = sphere_uv(p);
        ulo = fminf(ulo, uv.x); uhi = fmaxf(uhi, uv.x); vlo = fminf(vlo, uv.y); vhi = fmaxf(vhi, uv.y);
    }
    if (uhi - ulo > 0.25f || vlo < 0.03f || vhi > 0.97f) return false;  // date line / polar cap
    const float sx = (float)s.tex[3].w / (float)s.cm_b, sy = (float)s.tex[3].h / (float)s.cm_b;
    const int cu0 = max((int)(ulo * sx) - 1, 0), cu1 = min((int)(uhi * sx) + 1, s.cm_w - 1), cv0 = max((int)(vlo * sy) - 1, 0), cv1 = min((int)(vhi * sy) + 1, s.cm_h - 1);
    if ((cu1 - cu0 + 1) * (cv1 - cv0 + 1) > 256) return false;
    for (int cv = cv0; cv <= cv1; ++cv)
        for (int cu = cu0; cu <= cu1; ++cu)
            if (__ldg(s.cloud_max + cv * s.cm_w + cu)) return false;
    return true;
}
// one CTA; cls[t] = 1 for space tiles, 2 for tiles of another rank (tile partition: t % stride != offset), 0 / 3 for the persistent kernel's tiles
// (3 = clear, see above); wf_list = the 0-tiles in tile order, then the 3-tiles in tile order; counts = {n_wf, n_space}
__global__ void __launch_bounds__(1024) k_classify_tiles(const __grid_constant__ DevScene s, int x0, int y0, int w, int h, int tiles_x, int n_tiles, int enable,
                                                        int order, int stride, int offset, unsigned char *cls, unsigned int *wf_list, unsigned int *counts) {
    __shared__ unsigned int warp_sum[32];
    __shared__ unsigned int base_wf, n_space;
    const DevDerived dv = *s.derived;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { base_wf = 0u; n_space = 0u; }
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 classifies and lists the tiles that may hold long paths, pass 1 appends the clear ones
        for (int b = 0; b < n_tiles; b += 1024) {
            const int t = b + (int)threadIdx.x;
            bool keep = false;
            if (t < n_tiles) {
                unsigned char c;
                if (pass == 0) {
                    const int tx = t % tiles_x, ty = t / tiles_x;
                    const int px0 = x0 + tx * kDeTileW, py0 = y0 + ty * kDeTileH, px1 = min(px0 + kDeTileW, x0 + w), py1 = min(py0 + kDeTileH, y0 + h);
                    const bool mine = t % stride == offset;
                    const bool space = mine && enable && tile_is_space(s, dv, px0, py0, px1, py1);
                    const bool clear = mine && !space && order && tile_is_clear(s, dv, px0, py0, px1, py1);
                    c = !mine ? 2 : (space ? 1 : (clear ? 3 : 0));
                    cls[t] = c;
                    if (space) atomicAdd(&n_space, 1u);
                } else c = cls[t];
                keep = c == (pass == 0 ? 0 : 3);
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (lane == 0) warp_sum[wid] = __popc(bal);
            __syncthreads();
            unsigned int off = base_wf;
            for (int k = 0; k < wid; ++k) off += warp_sum[k];
            if (keep) wf_list[off + __popc(bal & ((1u << lane) - 1u))] = (unsigned)t;
            __syncthreads();
            if (threadIdx.x == 0) { unsigned int tot = 0u; for (int k = 0; k < 32; ++k) tot += warp_sum[k]; base_wf += tot; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) { counts[0] = base_wf; counts[1] = n_space; }
}

// Renderer.render for tiles that cannot see the planet: one thread per pixel, n samples in registers, one add per pixel.
// Same random stream, same expressions as stage_new + end_path(primary_miss), minus the (certainly failing) shell test.
#ifndef WF_SPACE_SPLIT
#define WF_SPACE_SPLIT 64  // samples per CTA: enough CTAs at low spp, 3 atomics per pixel and 64 samples at high spp
#endif
template <bool COUNT> __global__ void __launch_bounds__(kDeTileW * kDeTileH) k_space_tiles(const __grid_constant__ DevScene s, const __grid_constant__ WfParams P,
                                                                                            const unsigned char *__restrict__ cls) {
    const unsigned tile = blockIdx.x;
    if (cls[tile] != 1) return;
    __shared__ float cdf_s[kLambdaBins];
    for (int k = threadIdx.x; k < kLambdaBins; k += blockDim.x) cdf_s[k] = s.cdf[k];
    __syncthreads();
    const int px = P.x0 + (int)(tile % (unsigned)P.tiles_x) * kDeTileW + (int)(threadIdx.x & 15u);
    const int py = P.y0 + (int)(tile / (unsigned)P.tiles_x) * kDeTileH + (int)(threadIdx.x >> 4);
    if (px >= P.x0 + P.w || py >= P.y0 + P.h) return;
    const DevDerived dv = *s.derived;
    const uint32_t pix = (uint32_t)(py * s.W + px);
    const int s0 = (int)blockIdx.y * WF_SPACE_SPLIT, s1 = min(s0 + WF_SPACE_SPLIT, P.n_spp);
    float3 acc = f3(0.0f, 0.0f, 0.0f), acc2 = f3(0.0f, 0.0f, 0.0f);
    unsigned ntex = 0u;
#pragma unroll 1
    for (int sp = s0; sp < s1; ++sp) {
        const uint4 b = philox_block(P.seed, pix, P.first_sample + (uint32_t)sp, 0u, 0u);
        const int bin = spectrum_bin(cdf_s, u32_to_unit(b.x));
        const float3 dir = get_cast_dir(s, dv, (float)px, (float)py, u32_to_unit(b.y), u32_to_unit(b.z));
        const LambdaRow &lr = s.lam[bin];
        float Lr = 0.0f;
        if (dot(dv.light_dir, dir) > dv.sun_cos_angle) Lr += lr.sun_power;
        ++ntex;
        const float3 st = sample_sphere_rgb8(s.tex[6], dir);
        const float stars_power = lr.s2s_valid != 0.0f ? dot(st, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
        Lr += stars_power * lr.sun_power * 0.0000001f;
        if (isinf(Lr) || isnan(Lr) || Lr < 0.0f) Lr = 0.0f;
        if (Lr != 0.0f) {
            const float3 rgb = xyz_to_rgb((Lr * f3(lr.resp_x, lr.resp_y, lr.resp_z)) * lr.rcp_pdf);
            acc = acc + rgb;
            acc2 = acc2 + rgb * rgb;
        }
    }
    if (acc.x != 0.0f || acc.y != 0.0f || acc.z != 0.0f) {
        float *a = P.accum + (size_t)pix * 3;
        atomicAdd(a, acc.x); atomicAdd(a + 1, acc.y); atomicAdd(a + 2, acc.z);
        if (P.accum2) {
            float *a2 = P.accum2 + (size_t)pix * 3;
            atomicAdd(a2, acc2.x); atomicAdd(a2 + 1, acc2.y); atomicAdd(a2 + 2, acc2.z);
        }
    }
    if (COUNT && s.counters) {  // same bookkeeping as the generic route: one segment and one star fetch per path
        const unsigned n = (unsigned)(s1 - s0);
        const unsigned tot = __reduce_add_sync(__activemask(), n), tt = __reduce_add_sync(__activemask(), ntex);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(__activemask()) - 1)) {
            atomicAdd(&s.counters[C_PATHS], (unsigned long long)tot);
            atomicAdd(&s.counters[C_SEGMENTS], (unsigned long long)tot);
            atomicAdd(&s.counters[C_TEX], (unsigned long long)tt);
        }
    }
}

}  // namespace de_fast

struct DeWavefrontState {
    int device = 0, sm_count = 0;
    unsigned int *d_next = nullptr;
    float4 *d_cold = nullptr;                  // WF_COLD: shading state of the paths in flight, [sm_count][WF_SLOTS][4]
    unsigned long long *d_prof = nullptr;      // [0,32): stage profile, [32,40): launch timeline
    unsigned long long *d_cta = nullptr;       // [sm_count][24] per-CTA drain diagnostics (timeline mode)
    bool attr_set = false;
    // tile classification, cached per (parameter version, window)
    unsigned char *d_cls = nullptr;
    unsigned int *d_wf_list = nullptr, *d_counts = nullptr;
    int tiles_cap = 0;
    unsigned long long cls_version = ~0ull;
    int cls_win[4] = {-1, -1, -1, -1}, cls_enable = -1, cls_order = -1, cls_part[2] = {-1, -1};
    // k_space_tiles runs on a side stream so its CTAs fill the SMs the persistent kernel's drain leaves idle
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

DeWavefrontState *de_wavefront_alloc(int device) {
    DeWavefrontState *st = new DeWavefrontState();
    st->device = device;
    cudaDeviceGetAttribute(&st->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaMalloc(&st->d_next, sizeof(unsigned int)) != cudaSuccess) { delete st; return nullptr; }
    if (cudaMalloc(&st->d_prof, sizeof(unsigned long long) * 40) != cudaSuccess) { cudaFree(st->d_next); delete st; return nullptr; }
    if (cudaMalloc(&st->d_counts, sizeof(unsigned int) * 2) != cudaSuccess) { cudaFree(st->d_next); cudaFree(st->d_prof); delete st; return nullptr; }
    cudaMemset(st->d_prof, 0, sizeof(unsigned long long) * 40);
#if WF_COLD
    if (cudaMalloc(&st->d_cold, sizeof(float4) * 4 * (size_t)WF_SLOTS * (size_t)st->sm_count) != cudaSuccess) {
        cudaFree(st->d_next); cudaFree(st->d_prof); cudaFree(st->d_counts); delete st; return nullptr;
    }
#endif
    if (cudaMalloc(&st->d_cta, sizeof(unsigned long long) * 24 * (size_t)st->sm_count) != cudaSuccess) st->d_cta = nullptr;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = numerically greatest = lowest priority
    if (cudaStreamCreateWithPriority(&st->side, cudaStreamNonBlocking, lo) != cudaSuccess) st->side = nullptr;
    if (cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming) != cudaSuccess) st->ev_fork = nullptr;
    if (cudaEventCreateWithFlags(&st->ev_join, cudaEventDisableTiming) != cudaSuccess) st->ev_join = nullptr;
    return st;
}
void de_wavefront_free(DeWavefrontState *st) {
    if (!st) return;
    cudaFree(st->d_next);
    cudaFree(st->d_prof);
    cudaFree(st->d_cold);
    cudaFree(st->d_cls); cudaFree(st->d_wf_list); cudaFree(st->d_counts); cudaFree(st->d_cta);
    if (st->side) cudaStreamDestroy(st->side);
    if (st->ev_fork) cudaEventDestroy(st->ev_fork);
    if (st->ev_join) cudaEventDestroy(st->ev_join);
    delete st;
}
int de_wavefront_render(DeWavefrontState *st, const DevScene &s, const DeWavefrontJob &job, cudaStream_t stream) {
    using namespace de_fast;
    size_t smem = sizeof(WarpPool);
    if (!st->attr_set) {
        cudaFuncSetAttribute(k_render_wavefront<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_render_wavefront<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        st->attr_set = true;
    }
    WfParams P;
    P.accum = job.accum; P.accum2 = job.accum2; P.next = st->d_next; P.cold = st->d_cold;
    P.tiles_x = (job.w + kDeTileW - 1) / kDeTileW;
    const long long tiles = (long long)P.tiles_x * ((job.h + kDeTileH - 1) / kDeTileH);
    P.x0 = job.x0; P.y0 = job.y0; P.w = job.w; P.h = job.h; P.seed = job.seed;
    const bool count = job.count;
    P.prof = count ? st->d_prof : nullptr;
    P.timeline = job.timeline ? st->d_prof + 32 : nullptr;
    P.cta_stats = job.timeline ? st->d_cta : nullptr;
    if (count) cudaMemsetAsync(st->d_prof, 0, sizeof(unsigned long long) * 32, stream);
    if (job.timeline) {
        const unsigned long long init[8] = {~0ull, ~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull, 0ull};
        cudaMemcpyAsync(st->d_prof + 32, init, sizeof(init), cudaMemcpyHostToDevice, stream);  // pageable source: staged before the call returns
    }
    // tiles that cannot see the planet go to k_space_tiles; the classification only changes with the camera or the window
    if ((int)tiles > st->tiles_cap) {
        cudaFree(st->d_cls); cudaFree(st->d_wf_list);
        st->d_cls = nullptr; st->d_wf_list = nullptr; st->tiles_cap = 0;
        if (cudaMalloc(&st->d_cls, (size_t)tiles) != cudaSuccess || cudaMalloc(&st->d_wf_list, (size_t)tiles * sizeof(unsigned int)) != cudaSuccess) return -1;
        st->tiles_cap = (int)tiles;
        st->cls_version = ~0ull;
    }
    const int enable = job.space_tiles ? 1 : 0, order = job.tile_order ? 1 : 0;
    if (st->cls_version != job.param_version || st->cls_win[0] != job.x0 || st->cls_win[1] != job.y0 || st->cls_win[2] != job.w || st->cls_win[3] != job.h ||
        st->cls_enable != enable || st->cls_order != order || st->cls_part[0] != job.tile_stride || st->cls_part[1] != job.tile_offset) {
        k_classify_tiles<<<1, 1024, 0, stream>>>(s, job.x0, job.y0, job.w, job.h, P.tiles_x, (int)tiles, enable, order, job.tile_stride, job.tile_offset, st->d_cls,
                                                st->d_wf_list, st->d_counts);
        st->cls_version = job.param_version; st->cls_enable = enable; st->cls_order = order; st->cls_part[0] = job.tile_stride; st->cls_part[1] = job.tile_offset;
        st->cls_win[0] = job.x0; st->cls_win[1] = job.y0; st->cls_win[2] = job.w; st->cls_win[3] = job.h;
    }
    P.tile_list = st->d_wf_list; P.n_tiles = st->d_counts;
    // the work counter is 32 bits wide: at most 2^31 chunks (2^36 paths) per launch, more samples go in several launches
    const long long max_spp = ((1LL << 31) / (tiles * 4) > 1) ? (1LL << 31) / (tiles * 4) : 1;
    for (long long done = 0; done < job.n_spp; done += max_spp) {
        const int batch = (int)((job.n_spp - done < max_spp) ? job.n_spp - done : max_spp);
        P.n_spp = batch; P.first_sample = job.first_sample + (uint32_t)done;
        cudaMemsetAsync(st->d_next, 0, sizeof(unsigned int), stream);
        const bool async = enable && job.space_async && st->side && st->ev_fork && st->ev_join;
        cudaStream_t sstream = async ? st->side : stream;
        if (async) { cudaEventRecord(st->ev_fork, stream); cudaStreamWaitEvent(st->side, st->ev_fork, 0); }
        const int grid = st->sm_count;
        if (count) k_render_wavefront<true><<<grid, WF_WARPS * 32, smem, stream>>>(s, P);
        else k_render_wavefront<false><<<grid, WF_WARPS * 32, smem, stream>>>(s, P);
        if (enable) {  // (disjoint pixels: the two kernels never touch the same accumulator)
            const dim3 sg((unsigned)tiles, (unsigned)((batch + WF_SPACE_SPLIT - 1) / WF_SPACE_SPLIT));
            if (count) k_space_tiles<true><<<sg, kDeTileW * kDeTileH, 0, sstream>>>(s, P, st->d_cls);
            else k_space_tiles<false><<<sg, kDeTileW * kDeTileH, 0, sstream>>>(s, P, st->d_cls);
        }
        if (async) { cudaEventRecord(st->ev_join, st->side); cudaStreamWaitEvent(stream, st->ev_join, 0); }
    }
    return 0;
}

int de_wavefront_profile(DeWavefrontState *st, unsigned long long *out40) {
    if (!st) return -1;
    return cudaMemcpy(out40, st->d_prof, sizeof(unsigned long long) * 40, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
int de_wavefront_tile_counts(DeWavefrontState *st, unsigned int *out2) {
    if (!st) return -1;
    return cudaMemcpy(out2, st->d_counts, sizeof(unsigned int) * 2, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
int de_wavefront_cta_stats(DeWavefrontState *st, unsigned long long *out, int max_ctas) {
    if (!st || !st->d_cta) return -1;
    const int n = st->sm_count < max_ctas ? st->sm_count : max_ctas;
    return cudaMemcpy(out, st->d_cta, sizeof(unsigned long long) * 24 * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess ? n : -2;
}
