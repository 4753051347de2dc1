// de_integrator.cuh -- the reference integrator (pathtracer.py:11-469) one thread per path.
// Used by the PARITY flavour (DE_EXACT=1) and by the fast megakernel baseline; the wavefront
// kernel (de_wavefront.cuh) runs the same mathematics as a stage machine.
#pragma once
#include "de_device.cuh"

namespace DE_NS {

// pathtracer.py:11-14
template <bool COUNT> DE_DEV float land_sdf(const DevScene &s, float3 pos, float scale, Counters &cn) {
    DE_COUNT(cn, C_SDF); DE_COUNT(cn, C_TEX);
    return length(pos) - kPlanetR - scale * sample_sphere_r8(s.tex[1], pos);
}
// pathtracer.py:16-25
template <bool COUNT> DE_DEV float3 land_normal(const DevScene &s, float eps, float3 pos, float scale, Counters &cn) {
    float d = land_sdf<COUNT>(s, pos, scale, cn);
    float3 n = f3(d - land_sdf<COUNT>(s, f3(pos.x - eps, pos.y - 0.0f, pos.z - 0.0f), scale, cn),
                  d - land_sdf<COUNT>(s, f3(pos.x - 0.0f, pos.y - eps, pos.z - 0.0f), scale, cn),
                  d - land_sdf<COUNT>(s, f3(pos.x - 0.0f, pos.y - 0.0f, pos.z - eps), scale, cn));
    return normalize(n);
}
// pathtracer.py:27-46
template <bool COUNT> DE_DEV float intersect_land(const DevScene &s, float3 pos, float3 dir, float height_scale, Counters &cn) {
    float ray_dist = 0.0f;
    const float max_ray_dist = 63710000.0f;
    float2 rd = rsi(pos, dir, kAtmosUpper);
    if (rd.x > 0.0f) ray_dist = rd.x;
#if !DE_EXACT
    if (land_surely_missed(pos + dir * ray_dist, dir, ray_dist, height_scale)) return -1.0f;
    ray_dist += skip_to_terrain_top(pos + dir * ray_dist, dir, ray_dist, height_scale);
#endif
    for (int i = 0; i < 250; ++i) {
        float3 ro = pos + dir * ray_dist;
#if !DE_EXACT
        if (march_surely_missed(ro, dir, dot(ro, ro), ray_dist, height_scale)) return -1.0f;
#endif
        float dist = land_sdf<COUNT>(s, ro, height_scale, cn);
        ray_dist += dist;
        if (ray_dist > max_ray_dist || fabsf(dist) < ray_dist * 0.0001f) break;
    }
    return ray_dist < max_ray_dist ? ray_dist : -1.0f;
}
// pathtracer.py:48-65
template <bool COUNT> DE_DEV float get_clouds_density(const DevScene &s, float3 pos, Counters &cn) {
    float r = length(pos), density = 0.0f;
    if (r > kCloudsLower && r < kCloudsUpper) {
        float h = (r - kCloudsLower) / kCloudsThickness;
        DE_COUNT(cn, C_TEX);
        float c = sample_sphere_r8(s.tex[3], pos);
        const float split = 0.2f;
        density = (h - split < c * (1.0f - split) && split - h < c * split) ? fmaxf(c, 0.4f) : 0.0f;
    }
    return density * kCloudsDensity;
}
// pathtracer.py:145-169
DE_DEV void intersect_cloud_limits(float3 pos, float3 dir, float land_isection, float &t_start, float &t_max) {
    float elevation = length(pos);
    float2 lo = rsi(pos, dir, kCloudsLower), up = rsi(pos, dir, kCloudsUpper);
    if (elevation >= kCloudsUpper) {
        t_start = fmaxf(0.0f, up.x);
        t_max = lo.y >= 0.0f ? lo.x : up.y;
        if (up.y < 0.0f) t_max = -1.0f;
    } else if (elevation >= kCloudsLower) {
        t_start = 0.0f;
        t_max = lo.y >= 0.0f ? lo.x : up.y;
    } else {
        t_start = lo.y;
        t_max = up.y;
        if (land_isection > 0.0f) t_max = -1.0f;
    }
}
// pathtracer.py:77-115.  IS_CLOUD selects which half of get_atmos_density is live: the other
// half is multiplied by a zero extinction in the reference (pathtracer.py:98,185,197), so
// skipping it is exact (0*finite == 0 and x+0 == x).
template <bool COUNT, bool IS_CLOUD, class R>
DE_DEV int delta_tracking(const DevScene &s, float3 pos, float3 dir, float t_start, float t_max, float3 ext_rmo, float ext_cloud, float max_ext,
                          R &rng, Counters &cn, float &t_out, int &id_out) {
    float t = t_start;
    pos = pos + dir * t;
    int id = 0, event = kNullEvent;
    rng.align();
    while (t < t_max) {
        float t_step = -logf(rng.next()) / max_ext;
        pos = pos + dir * t_step;
        t += t_step;
        if (t >= t_max) break;
        float es0 = 0.0f, es1 = 0.0f, es2 = 0.0f, sum;
        if (IS_CLOUD) {
            DE_COUNT(cn, C_CLOUD);
            sum = ext_cloud * get_clouds_density<COUNT>(s, pos, cn);
        } else {
            DE_COUNT(cn, C_RMO);
            float3 d = get_density(get_elevation(pos));
            es0 = ext_rmo.x * d.x; es1 = ext_rmo.y * d.y; es2 = ext_rmo.z * d.z;
            sum = (es0 + es1) + es2;
        }
        float rand = rng.next();
        if (rand < sum / max_ext) {
            if (IS_CLOUD) id = 3;
            else {
                float cmf = es0;
                if (!(rand < cmf / max_ext)) {
                    id = 1; cmf += es1;
                    if (!(rand < cmf / max_ext)) {
                        id = 2; cmf += es2;
                        if (!(rand < cmf / max_ext)) id = 3;
                    }
                }
            }
            event = sample_scatter_event(id, rng) ? kScatterEvent : kAbsorbEvent;
            break;
        }
    }
    t_out = t; id_out = id;
    return event;
}
// pathtracer.py:117-143
template <bool COUNT, bool IS_CLOUD, class R>
DE_DEV float ratio_tracking(const DevScene &s, float3 pos, float3 dir, float t_start, float t_max, float3 ext_rmo, float ext_cloud, float max_ext, R &rng, Counters &cn) {
    float t = t_start;
    pos = pos + dir * t;
    float T = 1.0f;
    rng.align();
    while (t < t_max) {
        float t_step = -logf(rng.next()) / max_ext;
        rng.skip();
        pos = pos + dir * t_step;
        t += t_step;
        if (t >= t_max) break;
        float sum;
        if (IS_CLOUD) {
            DE_COUNT(cn, C_CLOUD);
            sum = ext_cloud * get_clouds_density<COUNT>(s, pos, cn);
        } else {
            DE_COUNT(cn, C_RMO);
            float3 d = get_density(get_elevation(pos));
            sum = (ext_rmo.x * d.x + ext_rmo.y * d.y) + ext_rmo.z * d.z;
        }
        T *= 1.0f - sum / max_ext;
        if (T < 1e-5f) break;
    }
    return T;
}
// pathtracer.py:172-207
template <bool COUNT, class R>
DE_DEV int sample_interaction(const DevScene &s, float3 pos, float3 dir, float land_isection, float3 ext_rmo, float ext_cloud, float max_rmo, float max_cloud,
                              R &rng, Counters &cn, float &t_out, int &id_out) {
    float2 atm = rsi(pos, dir, kAtmosUpper);
    float t_start = fmaxf(0.0f, atm.x);
    float t_max = land_isection >= 0.0f ? land_isection : atm.y;
    if (atm.y < 0.0f) t_max = -1.0f;
    float rmo_t; int rmo_id;
#if !DE_EXACT
    if (t_start < t_max) max_rmo = fminf(max_rmo, rmo_segment_majorant(ext_rmo, pos, dir, t_start, t_max));
#endif
    int rmo_event = delta_tracking<COUNT, false>(s, pos, dir, t_start, t_max, ext_rmo, ext_cloud, max_rmo, rng, cn, rmo_t, rmo_id);
    intersect_cloud_limits(pos, dir, land_isection, t_start, t_max);
    int event = rmo_event, id = rmo_id;
    float t = rmo_t;
    if (rmo_event == kNullEvent || rmo_t > t_start) {
        float cloud_t; int cloud_id;
#if !DE_EXACT
        if (t_start < t_max) {  // product flavour: local majorant + layer top from the coarse cloud max-map (unbiased: any bound works)
            float bound = cloud_pass_setup(s, pos, dir, t_start, t_max);
            if (bound == 0.0f) t_max = t_start; else max_cloud = ext_cloud * bound;
        }
#endif
        int cloud_event = delta_tracking<COUNT, true>(s, pos, dir, t_start, t_max, ext_rmo, ext_cloud, max_cloud, rng, cn, cloud_t, cloud_id);
        if (cloud_event > 0 && (cloud_t < rmo_t || rmo_event == kNullEvent)) { t = cloud_t; id = kCloud; event = cloud_event; }
    }
    t_out = t; id_out = id;
    return event;
}
// pathtracer.py:211-232
template <bool COUNT, class R>
DE_DEV float sample_transmittance(const DevScene &s, float3 pos, float3 dir, float land_isection, float3 ext_rmo, float ext_cloud, float max_rmo, float max_cloud, R &rng, Counters &cn) {
    float2 atm = rsi(pos, dir, kAtmosUpper);
    float t_start = fmaxf(0.0f, atm.x);
    float t_max = land_isection >= 0.0f ? land_isection : atm.y;
    if (atm.y < 0.0f) t_max = -1.0f;
#if !DE_EXACT
    if (t_start < t_max) max_rmo = fminf(max_rmo, rmo_segment_majorant(ext_rmo, pos, dir, t_start, t_max));
#endif
    float T = ratio_tracking<COUNT, false>(s, pos, dir, t_start, t_max, ext_rmo, ext_cloud, max_rmo, rng, cn);
    intersect_cloud_limits(pos, dir, land_isection, t_start, t_max);
#if !DE_EXACT
    if (t_start < t_max) {
        float bound = cloud_pass_setup(s, pos, dir, t_start, t_max);
        if (bound == 0.0f) t_max = t_start; else max_cloud = ext_cloud * bound;
    }
#endif
    T *= ratio_tracking<COUNT, true>(s, pos, dir, t_start, t_max, ext_rmo, ext_cloud, max_cloud, rng, cn);
    return T;
}
// pathtracer.py:284-313
template <bool COUNT> DE_DEV LandMaterial get_land_material(const DevScene &s, float3 pos, Counters &cn) {
    LandMaterial m;
    float2 uv = sphere_uv(pos);
    if (COUNT) cn.v[C_TEX] += 4;
    m.ocean = tex_r8(s.tex[2], uv.x, uv.y);
    m.albedo_srgb = grade_albedo(tex_rgb8(s.tex[0], uv.x, uv.y), m.ocean);
    m.bathymetry = tex_r8(s.tex[4], uv.x, uv.y);
    m.emissive = tex_r8(s.tex[5], uv.x, uv.y);
    return m;
}

// pathtracer.py:316-469
template <bool COUNT, class R>
DE_DEV float path_tracer(const DevScene &s, const DevDerived &dv, const LambdaRow &lr, float3 ray_pos, float3 ray_dir, R &rng, Counters &cn) {
    const float3 path_ray_dir = ray_dir;
    const float3 ext_rmo = f3(lr.ext_r, lr.ext_m, lr.ext_o);
    float ext_cloud = kCloudsExtinct;
    bool primary_miss = false;
    float in_scattering = 0.0f, throughput = 1.0f;
    for (int scatter_count = 0; scatter_count < 25; ++scatter_count) {
        rng.set_bounce((uint32_t)scatter_count + 1u);
        DE_COUNT(cn, C_SEGMENTS);
        if (scatter_count > 9) ext_cloud = 0.02f;
        float max_ext_rmo = lr.max_ext_rmo;
        float max_ext_cloud = ext_cloud * kCloudsDensity;
        float earth_isect = intersect_land<COUNT>(s, ray_pos, ray_dir, s.land_height_scale, cn);
        float interaction_dist; int id;
        int event = sample_interaction<COUNT>(s, ray_pos, ray_dir, earth_isect, ext_rmo, ext_cloud, max_ext_rmo, max_ext_cloud, rng, cn, interaction_dist, id);
        if (scatter_count > 9 && id == kCloud) id = kIsoCloud;
        rng.align();
        float3 light_dir = sample_cone_oriented(dv.sun_cos_angle, dv.light_dir, rng);
        if (event == kAbsorbEvent) break;
        else if (event == kScatterEvent) {
            float3 ipos = ray_pos + ray_dir * interaction_dist;
            bool direct_visibility = rsi(ipos, light_dir, kPlanetR).y > 0.0f;
            float direct_T = 0.0f;
            if (!direct_visibility) direct_T = sample_transmittance<COUNT>(s, ipos, light_dir, -1.0f, ext_rmo, ext_cloud, max_ext_rmo, max_ext_cloud, rng, cn);
            float direct_phase = evaluate_phase(ray_dir, light_dir, id, scatter_count > 0);
            in_scattering += throughput * direct_T * lr.sun_irradiance * direct_phase;
            float pdp;
            rng.align();
            float3 sd = sample_phase(ray_dir, id, scatter_count > 0, rng, pdp);
            ray_dir = sd; ray_pos = ipos; throughput *= pdp;
        } else if (earth_isect > 0.0f) {
            DE_COUNT(cn, C_SURF);
            float3 land_pos = ray_pos + ray_dir * earth_isect;
            float3 nrm = land_normal<COUNT>(s, dv.normal_eps, land_pos, s.land_height_scale, cn);
            LandMaterial m = get_land_material<COUNT>(s, land_pos, cn);
            float albedo = lr.s2s_valid != 0.0f ? dot(m.albedo_srgb, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
            in_scattering += throughput * m.emissive * lr.nightlights_power;
            float3 offset_pos = land_pos * (1.0f + 0.0001f * s.land_height_scale / 12000.0f);
            bool vis = intersect_land<COUNT>(s, offset_pos, light_dir, s.land_height_scale, cn) < 0.0f;
            float direct_T = sample_transmittance<COUNT>(s, offset_pos, light_dir, vis ? -1.0f : 0.0f, ext_rmo, ext_cloud, max_ext_rmo, max_ext_cloud, rng, cn);
            float ndl;
            float dbrdf = earth_brdf(albedo, m.ocean, m.bathymetry, -ray_dir, nrm, light_dir, ndl);
            in_scattering += throughput * direct_T * (vis ? 1.0f : 0.0f) * lr.sun_irradiance * dbrdf * ndl;
            float3 view_dir = -ray_dir;
            rng.align();
            ray_dir = sample_hemisphere_cosine_weighted(nrm, rng);
            ray_pos = offset_pos;
            float unused;
            float brdf = earth_brdf(albedo, m.ocean, m.bathymetry, view_dir, nrm, ray_dir, unused);
            throughput *= brdf * kPi;
        } else {
            if (scatter_count == 0) primary_miss = true;
            break;
        }
        if (scatter_count > 3) {
            float p = fmaxf(0.05f, 1.0f - throughput);
            if (rng.next() < p) break;
            throughput /= 1.0f - p;
        }
    }
    if (primary_miss) {
        if (dot(dv.light_dir, path_ray_dir) > dv.sun_cos_angle) in_scattering += lr.sun_power;
        DE_COUNT(cn, C_TEX);
        float3 st = sample_sphere_rgb8(s.tex[6], path_ray_dir);
        float stars_power = lr.s2s_valid != 0.0f ? dot(st, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
        in_scattering += stars_power * lr.sun_power * 0.0000001f;
    }
    if (isinf(in_scattering) || isnan(in_scattering) || in_scattering < 0.0f) in_scattering = 0.0f;
    return in_scattering;
}

// ------------------------------------------------------------------ deterministic preview
// pathtracer.py:471-499: 16-step optical depth towards the light; 0 when the planet is in the way
DE_DEV float ray_march_transmittance(float3 pos, float3 dir, float3 ext) {
    const int steps = 16;
    const float r_steps = 1.0f / (float)steps;
    float T = 0.0f;
    const bool visibility = rsi(pos, dir, kPlanetR).y > 0.0f;
    if (!visibility) {
        float2 atm = rsi(pos, dir, kAtmosUpper);
        float t_max = atm.y;
        if (atm.y < 0.0f) t_max = -1.0f;
        const float dd = t_max * r_steps;
        const float3 step = dir * dd;
        float3 od = f3(0.0f, 0.0f, 0.0f);
        for (int k = 0; k < steps; ++k) {
            float3 d = get_density(get_elevation(pos));
            od = od + d * dd;
            pos = pos + step;
        }
        T = expf(-dot(ext, od));
    }
    return T;
}
// pathtracer.py:501-541: 64-step single scattering (Rayleigh + Mie) over [t_start, t_max]
DE_DEV void ray_march_atmos(float3 pos, float3 dir, float t_start, float t_max, float3 sun_dir, float3 ext, float2 scat, float &in_scatter, float &transmittance) {
    const int steps = 64;
    const float r_steps = 1.0f / (float)steps;
    const float dd = (t_max - t_start) * r_steps;
    const float3 step = dir * dd;
    pos = pos + dir * t_start;
    const float cos_theta = dot(dir, sun_dir);
    const float ph_r = rayleigh_phase(cos_theta), ph_m = klein_nishina_phase(cos_theta, kMieAsymmetry);
    transmittance = 1.0f;
    in_scatter = 0.0f;
#pragma unroll 1
    for (int i = 0; i < steps; ++i) {
        float3 density = get_density(get_elevation(pos));
        float step_od = dot(ext, density * dd);
        float step_T = saturate(expf(-step_od));
        float step_integral = saturate((1.0f - step_T) / step_od);
        float visible = transmittance * step_integral;
        float sun_T = ray_march_transmittance(pos, sun_dir, ext);
        float step_scat = scat.x * (density.x * ph_r) + scat.y * (density.y * ph_m);
        in_scatter += step_scat * sun_T * visible * dd;
        transmittance *= step_T;
        pos = pos + step;
    }
}
// pathtracer.py:543-685 (unreferenced upstream): <= 3 surface bounces, marched atmosphere, no clouds.  One random stream
// for the whole path (bounce key 1); the light-cone and hemisphere samples start on multiples of 4.
template <bool COUNT, class R>
DE_DEV float ray_marcher(const DevScene &s, const DevDerived &dv, const LambdaRow &lr, float3 ray_pos, float3 ray_dir, R &rng, Counters &cn) {
    const float3 path_ray_dir = ray_dir;
    const float3 ext = f3(lr.ext_r, lr.ext_m, lr.ext_o);
    const float2 scat = make_float2(lr.ext_r * kRayleighAlbedo, lr.ext_m * kAerosolAlbedo);
    bool primary_miss = false;
    float accum = 0.0f, throughput = 1.0f;
    rng.set_bounce(1u);
    for (int scatter_count = 0; scatter_count < 3; ++scatter_count) {
        DE_COUNT(cn, C_SEGMENTS);
        float earth_isect = intersect_land<COUNT>(s, ray_pos, ray_dir, s.land_height_scale, cn);
        float2 atm = rsi(ray_pos, ray_dir, kAtmosUpper);
        float t_start = fmaxf(0.0f, atm.x);
        float t_max = earth_isect > 0.0f ? earth_isect : atm.y;
        if (atm.y < 0.0f) { primary_miss = scatter_count == 0; break; }
        rng.align();
        float3 light_dir = sample_cone_oriented(dv.sun_cos_angle, dv.light_dir, rng);
        float in_scatter, transmittance;
        ray_march_atmos(ray_pos, ray_dir, t_start, t_max, light_dir, ext, scat, in_scatter, transmittance);
        accum += throughput * in_scatter;
        throughput *= transmittance;
        if (earth_isect > 0.0f) {
            DE_COUNT(cn, C_SURF);
            float3 land_pos = ray_pos + ray_dir * earth_isect;
            float3 nrm = land_normal<COUNT>(s, dv.normal_eps, land_pos, s.land_height_scale, cn);
            LandMaterial m = get_land_material<COUNT>(s, land_pos, cn);
            float albedo = lr.s2s_valid != 0.0f ? dot(m.albedo_srgb, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
            accum += throughput * m.emissive * lr.nightlights_power;
            float3 offset_pos = land_pos * (1.0f + 0.0001f * s.land_height_scale / 12000.0f);
            bool vis = intersect_land<COUNT>(s, offset_pos, light_dir, s.land_height_scale, cn) < 0.0f;
            float ndl;
            float dbrdf = earth_brdf(albedo, m.ocean, m.bathymetry, -ray_dir, nrm, light_dir, ndl);
            accum += throughput * 1.0f * (vis ? 1.0f : 0.0f) * lr.sun_irradiance * dbrdf * ndl;
            float3 view_dir = -ray_dir;
            rng.align();
            ray_dir = sample_hemisphere_cosine_weighted(nrm, rng);
            ray_pos = offset_pos;
            float unused;
            float brdf = earth_brdf(albedo, m.ocean, m.bathymetry, view_dir, nrm, ray_dir, unused);
            throughput *= brdf * kPi;
        }
    }
    if (primary_miss) {
        if (dot(dv.light_dir, path_ray_dir) > dv.sun_cos_angle) accum += lr.sun_power;
        DE_COUNT(cn, C_TEX);
        float3 st = sample_sphere_rgb8(s.tex[6], path_ray_dir);
        float stars_power = lr.s2s_valid != 0.0f ? dot(st, f3(lr.s2s_r, lr.s2s_g, lr.s2s_b)) : 0.0f;
        accum += stars_power * lr.sun_power * 0.0000001f;
    }
    if (isinf(accum) || isnan(accum) || accum < 0.0f) accum = 0.0f;
    return accum;
}

// renderer.py:305-330: one path sample for pixel (u,v) -> linear sRGB contribution
template <bool COUNT, bool PREVIEW = false>
DE_DEV float3 render_sample(const DevScene &s, const DevDerived &dv, int u, int v, uint32_t sample_index, uint32_t seed, Counters &cn, float *wl_out, float *L_out) {
    Rng rng;
    rng.init(seed, (uint32_t)(v * s.W + u), sample_index);
    int bin = spectrum_bin(s.cdf, rng.next());
    LambdaRow lr = s.lam[bin];
    float xu = rng.next(), xv = rng.next();
    float3 dir = get_cast_dir(s, dv, (float)u, (float)v, xu, xv);
    float L = PREVIEW ? ray_marcher<COUNT>(s, dv, lr, s.cam_pos, dir, rng, cn) : path_tracer<COUNT>(s, dv, lr, s.cam_pos, dir, rng, cn);
    if (COUNT) { cn.v[C_PATHS]++; }
    if (wl_out) *wl_out = lr.wavelength;
    if (L_out) *L_out = L;
    float3 xyz = (L * f3(lr.resp_x, lr.resp_y, lr.resp_z)) * lr.rcp_pdf;
    return xyz_to_rgb(xyz);
}

}  // namespace DE_NS
