// de_scene.h -- plain-data description of the scene as it lives in HBM; shared by the host-side
// C-ABI (de_api.cu) and both device flavours (de_exact / de_fast).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kDeRmoBands = 4;

struct DevTex {
    const uint8_t *data;  // row-major [y][x][c], y=0 south
    int w, h, c;
    cudaTextureObject_t obj;  // block-linear copy for the TEX-gather path (0 = absent)
};

// One row per possible wavelength sample; everything that depends on lambda only
// (pathtracer.py:332-343, colour.py:12-71).  Built on device by build_lambda_table().
struct LambdaRow {
    float wavelength;
    float ext_r, ext_m, ext_o;   // sigma Rayleigh / Mie / ozone [1/m]
    float max_ext_rmo;           // majorant (pathtracer.py:355)
    float sun_power, sun_irradiance, nightlights_power;
    float resp_x, resp_y, resp_z, rcp_pdf;  // CIE response and 1/pdf (colour.py:39-46)
    float s2s_r, s2s_g, s2s_b, s2s_valid;   // srgb2spec coefficients at lambda (colour.py:62-71)
};

struct DevDerived {  // SceneParameters + camera basis, computed once per set_params on device
    float3 light_dir;
    float sun_cos_angle, sun_angular_radius;
    float3 cam_d, cam_du, cam_dv;  // renderer.py:272-277
    float3 up_n;
    float normal_eps;              // pi*planet_r/TOPOGRAPHY_TEX_RES[0] (pathtracer.py:20)
};

struct DevScene {
    DevTex tex[7];
    const float *cie;      // [2][441][3], already rounded through fp16 (rgba16f texture, renderer.py:97)
    const float *s2s;      // [300][3] f32 (fp16 values expanded)
    const float *o3;       // [441]
    const float *crf;      // [n_crf][1024][3]
    const LambdaRow *lam;  // [512]
    const float *cdf;      // [512] mean CIE CDF at mid=j/512 (bisection thresholds)
    const DevDerived *derived;
    int n_crf;
    float3 cam_pos, look_at, up;
    float fov, aspect_scale, aspect_ratio, sun_angle, sun_path_rot, land_height_scale;
    float exposure, gamma;
    int selected_crf, crf_count;
    float vig_strength, vig_radius, vig_cx, vig_cy;
    int tonemapper, topo_tex_w;
    int W, H;
    unsigned long long *counters;  // DeCounters layout, or nullptr
    // coarse max-map of the cloud texture (product flavour: local tracking majorant); cell = cm_b x cm_b
    // texels, dilated by one texel for the bilinear footprint; nullptr disables it
    const uint8_t *cloud_max;
    int cm_w, cm_h, cm_b;
    // altitude bands of the rmo tracking majorant (product flavour): band k = altitudes [band_r[k], band_r[k+1]) (radii; band 0 reaches down
    // to the centre, the last band up to the atmosphere top) with density bounds (Rayleigh, aerosol, ozone) valid over the whole band
    float band_r[kDeRmoBands];
    float band_dr[kDeRmoBands], band_dm[kDeRmoBands], band_do[kDeRmoBands];
};


// launch geometry shared by host and device
constexpr int kDeTileW = 16, kDeTileH = 8;  // film tile = one 128-thread CTA (renderer.py:43-46)
