// crt_render.h records -> the shading structs (host only).
#pragma once
#include <cstddef>
#include <cstring>

#include "crt_render.h"
#include "scene_desc.h"
#include "volume.h"

namespace shade {

static_assert(offsetof(Material, pmf) == sizeof(crt_material), "crt_material must be the authored prefix of shade::Material");
static_assert(sizeof(Light) == sizeof(crt_light), "crt_light layout");
static_assert(sizeof(crt_material) == 236 && sizeof(crt_light) == 116, "documented sizes");
static_assert(sizeof(crt_volume) == 168 && offsetof(crt_volume, grid_data) == 160, "crt_volume layout (render_types.VOLUME_DTYPE)");

inline Material material_from_abi(const crt_material& m) {
    Material o{};
    std::memcpy(&o, &m, sizeof m);
    prepare_material(o);
    return o;
}
inline void material_to_abi(const Material& m, crt_material* out) { std::memcpy(out, &m, sizeof *out); }
inline Light light_from_abi(const crt_light& l) {
    Light o;
    std::memcpy(&o, &l, sizeof l);
    o.env = -1;  // maps are attached through crt_world_set_environment only
    return o;
}
// 3x3 inverse by cofactors on column vectors (what glam's Mat3A::inverse computes)
inline void mat3_inverse(V3 c0, V3 c1, V3 c2, V3& r0, V3& r1, V3& r2) {
    V3 t0 = cross(c1, c2), t1 = cross(c2, c0), t2 = cross(c0, c1);
    float inv_det = 1.0f / dot(c2, t2);
    // inverse = transpose([t0 t1 t2]) * inv_det; columns of the inverse:
    r0 = v3(t0.x, t1.x, t2.x) * inv_det;
    r1 = v3(t0.y, t1.y, t2.y) * inv_det;
    r2 = v3(t0.z, t1.z, t2.z) * inv_det;
}
inline void light_set_environment(Light& l, int32_t env, const float m[9]) {
    l.a = v3(m[0], m[1], m[2]);
    l.b = v3(m[3], m[4], m[5]);
    l.c = v3(m[6], m[7], m[8]);
    mat3_inverse(l.a, l.b, l.c, l.wa, l.wb, l.wc);
    l.env = env;
}
// `data` overrides crt_volume::grid_data (the caller's copy of the voxels: host vector or device block)
inline VolumeRegion volume_from_abi(const crt_volume& v, const float* host_data, const float* bind_data) {
    VolumeFieldDesc f;
    f.field = v.field;
    f.scale = v.noise_scale;
    f.octaves = v.noise_octaves > 1u ? v.noise_octaves : 1u;
    f.gain = v.noise_gain;
    f.lacunarity = v.noise_lacunarity;
    f.threshold = v.noise_threshold;
    f.seed = v.noise_seed;
    f.nx = v.grid_dims[0];
    f.ny = v.grid_dims[1];
    f.nz = v.grid_dims[2];
    f.data = host_data;  // max_value is folded into the majorant on the host
    VolumeRegion r = volume_region_new(v.local_to_world, v3(v.half_extent[0], v.half_extent[1], v.half_extent[2]), v3(v.sigma_s[0], v.sigma_s[1], v.sigma_s[2]),
                                       v3(v.sigma_a[0], v.sigma_a[1], v.sigma_a[2]), v.g, v3(v.emission[0], v.emission[1], v.emission[2]), v.density_scale, f);
    r.data = bind_data;
    return r;
}
inline bool volume_abi_valid(const crt_volume& v) {
    if (v.field > FIELD_GRID) return false;
    if (v.field == FIELD_GRID && (!v.grid_data || v.grid_dims[0] == 0 || v.grid_dims[1] == 0 || v.grid_dims[2] == 0)) return false;
    return true;
}
inline Camera camera_from_abi(const crt_camera& c) {
    return camera_new(v3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]), v3(c.lookat[0], c.lookat[1], c.lookat[2]), v3(c.vup[0], c.vup[1], c.vup[2]), c.vfov_deg,
                      c.aspect, c.aperture, c.focus_dist);
}
inline RenderSettings settings_from_abi(const crt_render_settings& s) {
    RenderSettings o;
    o.width = s.width;
    o.height = s.height;
    o.samples_per_pixel = s.samples_per_pixel;
    o.max_depth = s.max_depth;
    o.min_samples_per_pixel = s.min_samples_per_pixel;
    o.variance_threshold = s.variance_threshold;
    o.frame = s.frame;
    o.sampling_strategy = s.sampling_strategy;
    o.adaptive = s.adaptive;
    return o;
}

}  // namespace shade
