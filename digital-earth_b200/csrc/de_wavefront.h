// de_wavefront.h -- host interface of the persistent-thread wavefront integrator (de_wavefront.cu).
#pragma once
#include "de_scene.h"

struct DeWavefrontState;
struct DeWavefrontJob {
    float *accum = nullptr, *accum2 = nullptr;  // [H][W][3] sums and (optional) sums of squares
    int n_spp = 1;
    uint32_t seed = 0, first_sample = 0;
    int x0 = 0, y0 = 0, w = 0, h = 0;
    bool count = false;        // counting build: event counters + per-stage cycle profile
    bool timeline = false;     // with count: record the launch timeline (ramp / drain), see de_get_launch_timeline
    bool space_tiles = true;   // render tiles that cannot see the planet in k_space_tiles
    bool space_async = true;   // ... on a low-priority side stream, overlapping the persistent kernel's drain
    bool tile_order = true;    // work order of the persistent kernel: tiles that can produce long paths (cloud in sight, limb) first, clear tiles last
    unsigned long long param_version = 0;  // bumps whenever the camera changes (tile classification cache)
    int tile_stride = 1, tile_offset = 0;  // multi-GPU tile partition: render the film tiles t of the window with t % stride == offset
};
DeWavefrontState *de_wavefront_alloc(int device);
void de_wavefront_free(DeWavefrontState *st);
int de_wavefront_render(DeWavefrontState *st, const DevScene &s, const DeWavefrontJob &job, cudaStream_t stream);
// out40[3*stage + {0,1,2}] = {cycles, visits, slots} per stage (counting build), out40[3*ST_COUNT] = idle cycles (warp-level sums);
// out40[32..38] = launch timeline in globaltimer ns / chunks (DeWavefrontJob::timeline)
int de_wavefront_profile(DeWavefrontState *st, unsigned long long *out40);
int de_wavefront_tile_counts(DeWavefrontState *st, unsigned int *out2);  // {tiles in the persistent kernel, space tiles}
// timeline mode: per CTA 24 words {t_exhaust, t_few (< 64 paths alive), t_end, chunks claimed, visits[9], slots[9] after exhaustion}; returns CTAs
int de_wavefront_cta_stats(DeWavefrontState *st, unsigned long long *out, int max_ctas);
