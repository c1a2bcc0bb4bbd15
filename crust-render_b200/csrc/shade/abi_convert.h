// crt_render.h records -> the shading structs (host only).
#pragma once
#include <cstddef>
#include <cstring>

#include "crt_render.h"
#include "scene_desc.h"

namespace shade {

static_assert(offsetof(Material, pmf) == sizeof(crt_material), "crt_material must be the authored prefix of shade::Material");
static_assert(sizeof(Light) == sizeof(crt_light), "crt_light layout");
static_assert(sizeof(crt_material) == 236 && sizeof(crt_light) == 76, "documented sizes");

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
    return o;
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
