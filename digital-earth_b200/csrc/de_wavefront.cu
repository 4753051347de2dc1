// de_wavefront.cu -- persistent-thread wavefront integrator (fast arithmetic flavour only).
#include "de_integrator.cuh"
#include "de_launch.h"
#include "de_wavefront.h"

struct DeWavefrontState {
    int device = 0, sm_count = 0;
    unsigned int *d_next = nullptr;  // work counter
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
    // TEMPORARY (milestone 1): route to the fast megakernel until the stage machine lands.
    de_fast::launch_render_mega(s, accum, n_spp, seed, first_sample, x0, y0, w, h, count, stream);
}
