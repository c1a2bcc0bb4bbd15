// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of crust-core's integrator (crates/crust-core/src/tracer.rs:515-636 render_pixel,
// :1086-1558 trace_path with its forward walk + backward gather, :930-1035 emission ownership /
// escaped_emission / shadow_transmittance) over the oracle's crust-rt restatement, one path at a
// time exactly like the reference (one ray per Scene::intersect call, VertexRec list, gather from
// the deepest vertex). The wavefront CUDA integrator (csrc/render.cu) is structured differently
// (queues, forward throughput accumulation) and is checked against this file.
//
// What is shared with the product: the leaf shading math in crust-render_b200/csrc/shade/*.h
// (OpenPBR lobes, lights, camera, filter, sampler) is one host+device restatement; compiled here
// for the host it is pinned by the reference's own unit-test properties (tests/test_shade_kats.py).
// PARITY UNPINNED for sample values: the reference's sampler is the third-party crate
// `openqmc-rs = "0.1"` (crates/crust-core/Cargo.toml:17), not vendored; see shade/base.h.
// Path guiding is out of scope: `guiding == None` everywhere below. Volume regions follow
// tracer.rs:1016-1077 (shadow_transmittance, volume_nee) and :1166-1249 (the region scatter vertex) over
// shade/volume.h's restatement of volume.rs.
#pragma once
#include <thread>

#include "../crust-render_b200/csrc/shade/scene_desc.h"
#include "../crust-render_b200/csrc/shade/volume.h"
#include "crt_oracle.hpp"

namespace orcr {
using namespace shade;
// names that also exist in namespace orc (the crust-rt restatement has its own vector type)
using shade::V3;
using shade::v3;
using shade::splat;
using shade::dot;
using shade::cross;
using shade::normalize;
using shade::length;
using shade::length_squared;
using shade::lerp;
using shade::vmin;
using shade::vmax;
using shade::max_element;
using shade::is_zero;
using orc::Ctx;
using orc::Scene;
using orc::MASK_ALL;
using orc::MASK_CAMERA;
using orc::MASK_INDIRECT;
using orc::MASK_SHADOW;

inline V3 S(orc::V3 v) { return V3{v.x, v.y, v.z}; }
inline orc::V3 O(V3 v) { return orc::V3{v.x, v.y, v.z}; }

struct RayStats {  // stats.rs:128-146
    uint64_t camera_rays = 0, closest_hit = 0, shadow_rays = 0, vertices = 0, rr_tested = 0, rr_killed = 0, ended_escaped = 0,
             ended_depth = 0;
    void merge(const RayStats& o) {
        camera_rays += o.camera_rays;
        closest_hit += o.closest_hit;
        shadow_rays += o.shadow_rays;
        vertices += o.vertices;
        rr_tested += o.rr_tested;
        rr_killed += o.rr_killed;
        ended_escaped += o.ended_escaped;
        ended_depth += o.ended_depth;
    }
};

// rt_world.rs:142-293 — the kernel scene plus the geom_id-indexed material table.
struct World {
    std::shared_ptr<Scene> scene;
    std::vector<shade::Material> materials;
    std::vector<uint32_t> geom_material;  // geom_id -> index into materials
    std::vector<shade::Light> lights;     // LightList, light.rs:404-446
    std::vector<shade::EnvMapTables> env_tables;  // environment maps of dome lights (environment.rs)
    std::vector<shade::EnvMap> envs;              // views into env_tables, indexed by Light::env
    std::vector<shade::VolumeRegion> volumes;     // Volumes (volume.rs:364-384); grid voxels owned by `grids`
    std::vector<std::vector<float>> grids;
    const shade::Material& material(uint32_t geom_id) const { return materials[geom_material[geom_id]]; }
};

// ray.rs:13-17 (crust-core): the kernel ray plus the carried interior medium (here: the index of the
// material whose `medium` it is, -1 = vacuum).
struct PathRay {
    V3 origin, dir;
    float time = 0.0f;
    uint32_t mask = MASK_ALL;
    int medium = -1;
};
struct WorldHit {  // rt_world.rs:45-58
    shade::SurfaceHit rec;
    const shade::Material* mat;
    uint32_t geom_id, prim_id;
};

inline bool world_intersect(const World& w, const PathRay& r, float t_min, float t_max, Ctx& cx, WorldHit& out) {  // rt_world.rs:207-232
    orc::Ray kr{O(r.origin), O(r.dir), r.time, r.mask};
    auto h = orc::scene_intersect(*w.scene, kr, t_min, t_max, cx);
    if (!h) return false;
    out.rec.p = S(kr.at(h->t));
    out.rec.normal = S(h->normal);
    out.rec.t = h->t;
    out.rec.front_face = h->front_face;
    out.mat = &w.material(h->geom_id);
    out.geom_id = h->geom_id;
    out.prim_id = h->prim_id;
    return true;
}
inline bool world_occluded(const World& w, const PathRay& r, float t_min, float t_max, Ctx& cx) {  // rt_world.rs:235-237
    orc::Ray kr{O(r.origin), O(r.dir), r.time, r.mask};
    return w.scene->bvh.hit_any(kr, t_min, t_max, cx);
}

struct VertexRec {  // tracer.rs:832-863
    V3 atten, segment_emit, emit_here, nee, factor, next_emit;
    float next_emit_weight;
};
struct PrevBounce {  // PrevVertex, tracer.rs:898-920: Surface(PrevBounce) or Phase{pos, pdf}
    bool phase = false;
    V3 phase_pos{};
    PathRay ray;
    shade::SurfaceHit rec;
    const shade::Material* mat;
    V3 dir;
    float pdf;
    bool delta;
};

// tracer.rs:930-953
inline float bounce_emission_weight(const PrevBounce& p, const World& w, const WorldHit& hit, uint32_t strategy) {
    V3 value;
    float pdf;
    V3 from = p.phase_pos;
    if (!p.phase) {
        if (p.delta || !shade::material_eval(*p.mat, p.ray.dir, p.rec, p.dir, value, pdf)) return 1.0f;
        from = p.rec.p;
    }
    int li = shade::light_find_by_geom(w.lights.data(), (int)w.lights.size(), hit.geom_id);
    if (li < 0) return 1.0f;
    float light_pdf = std::max(shade::light_pdf_at_point(w.lights[li], from, hit.rec.p) / (float)w.lights.size(), 1e-6f);
    return shade::strategy_bounce_weight(strategy, p.pdf, light_pdf);
}
// tracer.rs:966-1009
inline V3 escaped_emission(const PrevBounce* prev, const World& w, V3 direction, uint32_t strategy, bool& covered) {
    covered = false;
    V3 radiance = shade::splat(0.0f);
    if (w.lights.empty()) return radiance;
    bool competing = false;
    float bounce_pdf = 0.0f;
    if (prev) {
        V3 value;
        float pdf;
        competing = prev->phase || (!prev->delta && shade::material_eval(*prev->mat, prev->ray.dir, prev->rec, prev->dir, value, pdf));
        bounce_pdf = prev->pdf;
    }
    float n_lights = (float)w.lights.size();
    for (const shade::Light& l : w.lights) {
        V3 emitted;
        float pdf;
        if (!shade::light_escaped(l, direction, emitted, pdf, w.envs.data())) continue;
        covered = true;
        float weight = 1.0f;
        if (competing && shade::strategy_samples_lights(strategy)) weight = shade::strategy_bounce_weight(strategy, bounce_pdf, std::max(pdf / n_lights, 1e-6f));
        radiance += emitted * weight;
    }
    return radiance;
}

// tracer.rs:1016-1035: ZERO when a surface occludes the segment, else the transmittance through every region crossed
inline V3 shadow_transmittance(const World& w, const PathRay& shadow_ray, float distance, PathSampler vertex, RayStats& stats, Ctx& cx) {
    stats.shadow_rays++;
    if (world_occluded(w, shadow_ray, 0.001f, distance - 0.001f, cx)) return splat(0.0f);
    if (w.volumes.empty()) return splat(1.0f);
    Rng rng = sampler_rng(new_domain(vertex, K_NEE_SHADOW));
    return volumes_transmittance(w.volumes.data(), (int)w.volumes.size(), shadow_ray.origin, shadow_ray.dir, 0.001f, distance - 0.001f, rng);
}
// tracer.rs:1040-1077
inline V3 volume_nee(V3 p, V3 wi, const PhaseMix& phase, const World& w, uint32_t strategy, PathSampler vertex, float time, RayStats& stats, Ctx& cx) {
    if (!strategy_samples_lights(strategy)) return splat(0.0f);
    float nee[4];
    draw_sample4(new_domain(vertex, K_NEE), nee);
    int li = light_pick((int)w.lights.size(), nee[0]);
    if (li < 0) return splat(0.0f);
    float n_lights = (float)w.lights.size();
    LightSample s;
    if (!light_sample_li(w.lights[li], p, nee[1], nee[2], s, w.envs.data())) return splat(0.0f);
    PathRay shadow_ray{p, s.direction, time, MASK_SHADOW, -1};
    V3 tr = shadow_transmittance(w, shadow_ray, s.distance, vertex, stats, cx);
    if (is_zero(tr)) return splat(0.0f);
    float light_pdf = fmaxf(s.pdf / n_lights, 1e-6f);
    float phase_val = phase_pdf(phase, dot(wi, s.direction));
    float weight = strategy_light_weight(strategy, light_pdf, phase_val);
    return s.radiance * phase_val * tr * weight / light_pdf;
}

// tracer.rs:1086-1558
inline V3 trace_path(const PathRay& r, const World& w, int depth, uint32_t strategy, shade::PathSampler sampler, std::vector<VertexRec>& records,
                            RayStats& stats, Ctx& cx) {
    PathSampler path = new_domain(sampler, K_PATH);
    records.clear();
    PathRay ray = r;
    int remaining = depth;
    bool has_prev = false;
    PrevBounce prev{};
    V3 beta = splat(1.0f);
    V3 terminal = splat(0.0f);

    for (;;) {
        PathSampler v = new_domain(path, (int32_t)records.size());
        if (remaining <= 0) {  // :1123-1148
            stats.ended_depth++;
            if (has_prev) {
                stats.closest_hit++;
                WorldHit hit;
                if (world_intersect(w, ray, 0.001f, INFINITY, cx, hit)) {
                    float cos_o = fabsf(dot(normalize(ray.dir), hit.rec.normal));
                    V3 emitted = material_emitted_directional(*hit.mat, cos_o);
                    if (length_squared(emitted) > 0.0f) {
                        if (ray.medium >= 0) emitted *= medium_transmittance(w.materials[ray.medium].medium, hit.rec.t);
                        if (!w.volumes.empty()) {
                            Rng rng = sampler_rng(new_domain(v, K_VOLUME));
                            emitted *= volumes_transmittance(w.volumes.data(), (int)w.volumes.size(), ray.origin, ray.dir, 0.001f, hit.rec.t, rng);
                        }
                        records.back().next_emit = emitted;
                        records.back().next_emit_weight = bounce_emission_weight(prev, w, hit, strategy);
                    }
                }
            }
            break;
        }
        stats.closest_hit++;
        WorldHit hit;
        bool have_hit = world_intersect(w, ray, 0.001f, INFINITY, cx, hit);
        float t_surf = have_hit ? hit.rec.t : INFINITY;
        // free flight in a scattering carried medium (:1158-1164)
        float t_med = INFINITY;
        const Medium* med = ray.medium >= 0 ? &w.materials[ray.medium].medium : nullptr;
        if (med && medium_is_scattering(*med)) {
            float sigma_t_max = fmaxf(medium_sigma_t_max(*med), 1e-4f);
            t_med = -logf(draw_rnd1(new_domain(v, K_MEDIUM))) / sigma_t_max;
        }
        // volume-region interaction clipped to the event that would otherwise end the segment (:1166-1182)
        V3 vol_tr = splat(1.0f), vol_emit = splat(0.0f);
        if (!w.volumes.empty()) {
            Rng rng = sampler_rng(new_domain(v, K_VOLUME));
            VolumeEvent ev;
            volumes_sample_interaction(w.volumes.data(), (int)w.volumes.size(), ray.origin, ray.dir, 0.001f, fminf(t_surf, t_med), rng, ev);
            if (ev.scatter) {  // region scatter vertex (:1184-1244)
                V3 wi = normalize(ray.dir);
                float ps[4];
                draw_sample4(new_domain(v, K_PHASE), ps);
                V3 dir = phase_sample(ev.phase, wi, ps[0], ps[1], ps[2]);
                float phase_pdf_dir = fmaxf(phase_pdf(ev.phase, dot(wi, dir)), 1e-6f);
                V3 nee = volume_nee(ev.p, wi, ev.phase, w, strategy, v, ray.time, stats, cx);
                VertexRec vrec{ev.weight, ev.emitted, splat(0.0f), nee, splat(1.0f), splat(0.0f), 1.0f};
                beta *= ev.weight;
                bool survived = true;
                if ((int)records.size() >= RR_START_BOUNCE) {
                    stats.rr_tested++;
                    float p_survive = clampf(max_element(beta), RR_MIN_PROB, 1.0f);
                    if (p_survive < 1.0f) {
                        if (draw_rnd1(new_domain(v, K_RR)) >= p_survive) {
                            survived = false;
                            stats.rr_killed++;
                            vrec.factor = splat(0.0f);
                        } else {
                            vrec.factor = vrec.factor / p_survive;
                            beta = beta / p_survive;
                        }
                    }
                }
                stats.vertices++;
                records.push_back(vrec);
                if (!survived) break;
                prev = PrevBounce{};
                prev.phase = true;
                prev.phase_pos = ev.p;
                prev.pdf = phase_pdf_dir;
                has_prev = true;
                int m = ray.medium;
                float time = ray.time;
                ray = PathRay{ev.p, dir, time, MASK_INDIRECT, m};
                remaining -= 1;
                continue;
            }
            vol_tr = ev.weight;
            vol_emit = ev.emitted;
        }

        if (t_med < t_surf) {  // carried-medium scatter vertex (:1251-1318)
            float sigma_bar = fmaxf(medium_sigma_t_max(*med), 1e-4f);
            V3 pos = ray.origin + t_med * ray.dir;
            float ph[4];
            draw_sample4(new_domain(v, K_PHASE), ph);
            V3 dir = sample_henyey_greenstein(normalize(ray.dir), med->g, ph[0], ph[1]);
            V3 e = (splat(sigma_bar) - (med->sigma_a + med->sigma_s)) * t_med;
            V3 factor = med->sigma_s / sigma_bar * v3(expf(e.x), expf(e.y), expf(e.z));
            VertexRec vrec{vol_tr, vol_emit, splat(0), splat(0), factor, splat(0), 1.0f};
            beta *= vol_tr * factor;
            bool survived = true;
            if ((int)records.size() >= RR_START_BOUNCE) {
                stats.rr_tested++;
                float p_survive = clampf(max_element(beta), RR_MIN_PROB, 1.0f);
                if (p_survive < 1.0f) {
                    if (draw_rnd1(new_domain(v, K_RR)) >= p_survive) {
                        survived = false;
                        stats.rr_killed++;
                        vrec.factor = splat(0.0f);
                    } else {
                        vrec.factor = vrec.factor / p_survive;
                        beta = beta / p_survive;
                    }
                }
            }
            stats.vertices++;
            records.push_back(vrec);
            if (!survived) break;
            int m = ray.medium;
            float time = ray.time;
            ray = PathRay{pos, dir, time, MASK_INDIRECT, m};
            remaining -= 1;
            has_prev = false;
            continue;
        }

        if (!have_hit) {  // background (:1321-1342)
            stats.ended_escaped++;
            V3 unit_direction = normalize(ray.dir);
            bool covered;
            V3 background = escaped_emission(has_prev ? &prev : nullptr, w, unit_direction, strategy, covered);
            if (!covered) {
                float t = 0.5f * (unit_direction.y + 1.0f);
                background += (1.0f - t) * v3(1.0f, 1.0f, 1.0f) + t * v3(0.5f, 0.7f, 1.0f);
            }
            terminal = vol_emit + vol_tr * background;
            break;
        }
        const SurfaceHit rec = hit.rec;
        const Material* mat = hit.mat;

        // arriving-segment attenuation (:1352-1361)
        V3 med_arrival = splat(1.0f);
        if (med) {
            if (medium_is_scattering(*med)) {
                float sigma_bar = fmaxf(medium_sigma_t_max(*med), 1e-4f);
                V3 e = (splat(sigma_bar) - (med->sigma_a + med->sigma_s)) * rec.t;
                med_arrival = v3(expf(e.x), expf(e.y), expf(e.z));
            } else {
                med_arrival = medium_transmittance(*med, rec.t);
            }
        }
        V3 atten = vol_tr * med_arrival;

        // emission ownership (:1369-1381)
        float cos_o = fabsf(dot(normalize(ray.dir), rec.normal));
        V3 emitted = material_emitted_directional(*mat, cos_o);
        V3 emit_here = splat(0.0f);
        if (has_prev) {
            if (length_squared(emitted) > 0.0f) {
                records.back().next_emit = atten * emitted;
                records.back().next_emit_weight = bounce_emission_weight(prev, w, hit, strategy);
            }
        } else {
            emit_here = emitted;
        }

        // 1. light sampling (:1394-1445)
        V3 nee = splat(0.0f);
        float nee_s[4];
        draw_sample4(new_domain(v, K_NEE), nee_s);
        int li = strategy_samples_lights(strategy) ? light_pick((int)w.lights.size(), nee_s[0]) : -1;
        LightSample ls;
        if (li >= 0 && light_sample_li(w.lights[li], rec.p, nee_s[1], nee_s[2], ls, w.envs.data())) {
            float n_lights = (float)w.lights.size();
            PathRay shadow_ray{rec.p, ls.direction, ray.time, MASK_SHADOW, -1};
            V3 shadow_tr = shadow_transmittance(w, shadow_ray, ls.distance, v, stats, cx);
            if (!is_zero(shadow_tr)) {
                float cosine = fabsf(dot(rec.normal, ls.direction));
                float light_pdf = fmaxf(ls.pdf / n_lights, 1e-6f);
                V3 brdf_value;
                float brdf_pdf;
                if (material_eval(*mat, ray.dir, rec, ls.direction, brdf_value, brdf_pdf)) {
                    float weight = strategy_light_weight(strategy, light_pdf, brdf_pdf);
                    nee += ls.radiance * brdf_value * cosine * shadow_tr * weight / light_pdf;
                }
            }
        }

        VertexRec vrec{atten, vol_emit, emit_here, nee, splat(0.0f), splat(0.0f), 1.0f};

        // 2. BSDF sampling (:1459-1530)
        float bs[4];
        draw_sample4(new_domain(v, K_BSDF), bs);
        ScatterSample sample;
        if (material_scatter(*mat, ray.dir, rec, bs, sample)) {
            V3 dir = normalize(sample.dir);
            float cosine = sample.delta ? 1.0f : fabsf(dot(rec.normal, dir));
            V3 factor = sample.value * cosine / sample.pdf;
            beta *= atten * factor;
            bool survived = true;
            if ((int)records.size() >= RR_START_BOUNCE) {
                stats.rr_tested++;
                float p_survive = clampf(max_element(beta), RR_MIN_PROB, 1.0f);
                if (p_survive < 1.0f) {
                    if (draw_rnd1(new_domain(v, K_RR)) >= p_survive) {
                        survived = false;
                        stats.rr_killed++;
                    } else {
                        factor = factor / p_survive;
                        beta = beta / p_survive;
                    }
                }
            }
            if (survived) {
                vrec.factor = factor;
                prev = PrevBounce{false, splat(0.0f), ray, rec, mat, dir, sample.pdf, sample.delta};
                has_prev = true;
                stats.vertices++;
                records.push_back(vrec);
                int m = sample.enters_medium ? (int)(mat - w.materials.data()) : -1;
                ray = PathRay{sample.origin, sample.dir, ray.time, MASK_INDIRECT, m};
                remaining -= 1;
                continue;
            }
        }
        stats.vertices++;
        records.push_back(vrec);
        break;
    }

    // backward gather (:1537-1556)
    V3 radiance = terminal;
    for (size_t k = records.size(); k-- > 0;) {
        const VertexRec& vr = records[k];
        radiance = vr.segment_emit + vr.atten * (vr.emit_here + vr.nee + vr.factor * (vr.next_emit * vr.next_emit_weight + radiance));
    }
    return radiance;
}

struct PixelOut {
    V3 mean;
    double variance;
    uint32_t taken;
};

// tracer.rs:515-636
inline PixelOut render_pixel(const World& w, const shade::Camera& cam, const shade::Filter& filter, const shade::RenderSettings& st, uint32_t i, uint32_t j,
                             std::vector<VertexRec>& scratch, RayStats& stats, Ctx& cx) {
    V3 sum = splat(0.0f);
    float weight_sum = 0.0f;
    double lum_sum = 0.0, lum_sq = 0.0;
    double threshold = (double)st.variance_threshold;
    uint32_t min_spp = std::max(st.min_samples_per_pixel, 2u);
    uint32_t taken = 0;
    int32_t tile = (int32_t)(i >> 8) + (int32_t)(j >> 8) * 4096;
    bool motion = w.scene->has_motion;
    for (uint32_t sample = 0; sample < st.samples_per_pixel; ++sample) {
        PathSampler root = new_domain(sampler_new((int32_t)i, (int32_t)j, st.frame, (int32_t)sample), tile);
        float cs[4];
        draw_sample4(new_domain(root, K_CAMERA), cs);
        float fx, wx, fy, wy;
        filter_sample(filter, cs[0], fx, wx);
        filter_sample(filter, cs[1], fy, wy);
        float u = ((float)i + fx) / (float)st.width;
        float v = ((float)j + fy) / (float)st.height;
        float time = 0.0f;
        if (motion) {
            float ts[4];
            draw_sample4(new_domain(root, K_TIME), ts);
            time = ts[0];
        }
        PathRay r;
        camera_get_ray(cam, u, v, cs[2], cs[3], r.origin, r.dir);
        r.time = time;
        r.mask = MASK_CAMERA;
        stats.camera_rays++;
        V3 color = trace_path(r, w, (int)st.max_depth, st.sampling_strategy, root, scratch, stats, cx) * (wx * wy);
        sum += color;
        weight_sum += wx * wy;
        double lum = (double)luminance(color);
        lum_sum += lum;
        lum_sq += lum * lum;
        taken = sample + 1;
        if (st.adaptive && threshold > 0.0 && taken >= min_spp && taken % 4 == 0) {
            double n = (double)taken;
            double var_of_mean = std::max((lum_sq - lum_sum * lum_sum / n) / (n - 1.0) / n, 0.0);
            double mean = std::max(lum_sum / n, 1e-4);
            if (std::sqrt(var_of_mean) / mean < threshold) break;
        }
    }
    double n = (double)taken;
    double variance = taken >= 2 ? std::max((lum_sq - lum_sum * lum_sum / n) / (n - 1.0) / n, 0.0) : INFINITY;
    V3 mean = weight_sum > 0.0f ? sum / weight_sum : sum / (float)taken;
    return PixelOut{mean, variance, taken};
}

// tracer.rs:408-513 render_pass (scanline flavour; rows are distributed over `threads`). `fb` is
// width*height*4 floats, row j = image row j counted from the BOTTOM (buffer.rs:40-44; the flip is
// applied on read, :74-77), alpha = samples taken. Only pixels with (tile id % tile_world) ==
// tile_rank are rendered (tile_world = 1: all) — the multi-GPU tile split of csrc/render.cu.
inline RayStats render(const World& w, const shade::Camera& cam, const shade::Filter& filter, const shade::RenderSettings& st, float* fb, double* var_map,
                       int threads, uint32_t tile_size, uint32_t tile_rank, uint32_t tile_world) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<RayStats> per(threads);
    std::vector<std::thread> pool;
    std::atomic<uint32_t> next_row{0};
    uint32_t tiles_x = tile_size ? (st.width + tile_size - 1) / tile_size : 1;
    auto work = [&](int tix) {
        Ctx cx;
        std::vector<VertexRec> scratch;
        scratch.reserve(st.max_depth + 1);
        for (;;) {
            uint32_t j = next_row.fetch_add(1);
            if (j >= st.height) break;
            for (uint32_t i = 0; i < st.width; ++i) {
                if (tile_world > 1) {
                    uint32_t tid = (j / tile_size) * tiles_x + (i / tile_size);
                    if (tid % tile_world != tile_rank) continue;
                }
                PixelOut p = render_pixel(w, cam, filter, st, i, j, scratch, per[tix], cx);
                float* o = fb + 4 * ((size_t)j * st.width + i);
                o[0] = p.mean.x;
                o[1] = p.mean.y;
                o[2] = p.mean.z;
                o[3] = (float)p.taken;
                if (var_map) var_map[(size_t)j * st.width + i] = p.variance;
            }
        }
    };
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    RayStats total;
    for (auto& s : per) total.merge(s);
    return total;
}

}  // namespace orcr
