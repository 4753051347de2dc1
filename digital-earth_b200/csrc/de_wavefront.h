// de_wavefront.h -- host interface of the persistent-thread wavefront integrator (de_wavefront.cu).
#pragma once
#include "de_scene.h"

struct DeWavefrontState;
DeWavefrontState *de_wavefront_alloc(int device);
void de_wavefront_free(DeWavefrontState *st);
void de_wavefront_render(DeWavefrontState *st, const DevScene &s, float *accum, int n_spp, uint32_t seed, uint32_t first_sample, int x0, int y0,
                         int w, int h, bool count, cudaStream_t stream);
// counting build only: out32[3*stage + {0,1,2}] = {cycles, visits, slots} per stage, out32[3*ST_COUNT] = idle cycles (warp-level sums)
int de_wavefront_profile(DeWavefrontState *st, unsigned long long *out32);
