/* crt_render.h — C ABI of the wavefront path tracer that sits on top of crt.h.
 *
 * BASELINE.json's north_star moves crust-core's integrator loop onto the device
 * (a one-ray-per-call CPU seam cannot feed a GPU), so besides the crust-rt seam
 * (crt.h) the library exports the engine-side records the integrator reads and
 * one render entry. Every `file:line` is relative to
 * /root/reference/crates/crust-core/src.
 *
 *   crt_world      rt_world.rs:142-293   World: committed scene + geom_id-indexed material table
 *                  light.rs:404-446      LightList
 *   crt_material   material/openpbr.rs:66-173 OpenPBR parameters (spec defaults), material/emissive.rs
 *   crt_light      light.rs:21-304       AreaLight{Sphere,Rect}, DistantLight, uniform DomeLight
 *   crt_camera     camera.rs:27-59       Camera::new arguments
 *   crt_render_settings tracer.rs:639-661 RenderSettings + filter.rs:28-60 PixelFilter
 *   crt_ray_stats  stats.rs:128-146      RayStats
 *   crt_volume     volume.rs:168-231     VolumeRegion::new arguments (free-standing participating media)
 *   crt_render     tracer.rs:194-212     Renderer::render_with_stats (non-guided pass)
 *
 * No CPU fallback: crt_render fails with CRT_ERR_NO_DEVICE without a CUDA device.
 */
#ifndef CRT_RENDER_H
#define CRT_RENDER_H
#include "crt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CRT_MAT_OPENPBR 0u
#define CRT_MAT_EMISSIVE 1u

/* openpbr.rs:66-128 (field order), defaults :130-173 via crt_material_default(). 236 bytes. */
typedef struct crt_material {
    uint32_t kind;
    float base_weight;
    float base_color[3];
    float base_diffuse_roughness, base_metalness;
    float specular_weight;
    float specular_color[3];
    float specular_roughness, specular_ior, specular_roughness_anisotropy;
    float transmission_weight;
    float transmission_color[3];
    float transmission_depth;
    float transmission_scatter[3];
    float transmission_scatter_anisotropy, transmission_dispersion_scale, transmission_dispersion_abbe_number;
    float subsurface_weight;
    float subsurface_color[3];
    float subsurface_radius;
    float subsurface_radius_scale[3];
    float subsurface_scatter_anisotropy;
    float fuzz_weight;
    float fuzz_color[3];
    float fuzz_roughness;
    float coat_weight;
    float coat_color[3];
    float coat_roughness, coat_roughness_anisotropy, coat_ior, coat_darkening;
    float thin_film_weight, thin_film_thickness, thin_film_ior;
    float emission_luminance;
    float emission_color[3];
    float geometry_opacity;
    uint32_t geometry_thin_walled;
    float emissive_color[3]; /* CRT_MAT_EMISSIVE: Emissive::color (emissive.rs:12-14) */
} crt_material;

#define CRT_LIGHT_SPHERE 0u  /* light.rs:21-44  SphereShape: a = center, r = radius */
#define CRT_LIGHT_RECT 1u    /* light.rs:55-81  RectShape: a = origin, b = edge_u, c = edge_v, n = unit normal */
#define CRT_LIGHT_DISTANT 2u /* light.rs:234-304: n = unit travel direction, r = cos(half angle), s = cone solid angle */
#define CRT_LIGHT_DOME 3u    /* light.rs:320-389, uniform (no environment map): radiance = tint */
#define CRT_NO_GEOM 0xFFFFFFFFu

/* 116 bytes */
typedef struct crt_light {
    uint32_t kind;
    uint32_t geom_id; /* world geom_id of the emissive geometry an area light shares (light.rs:150-160), else CRT_NO_GEOM */
    float radiance[3];
    float a[3], b[3], c[3], n[3];
    float r, s;
    int32_t env;      /* set to -1; crt_world_set_environment fills it and the two matrices below */
    float w2l[9];     /* dome with a map: world_to_light columns (light_to_world columns travel in a, b, c) */
} crt_light;

/* camera.rs:27-35 */
typedef struct crt_camera {
    float lookfrom[3], lookat[3], vup[3];
    float vfov_deg, aspect, aperture, focus_dist;
} crt_camera;

#define CRT_FILTER_BOX 0u
#define CRT_FILTER_TRIANGLE 1u
#define CRT_FILTER_GAUSSIAN 2u
#define CRT_FILTER_BLACKMAN 3u
#define CRT_FILTER_MITCHELL 4u
#define CRT_STRATEGY_POWER_MIS 0u
#define CRT_STRATEGY_BALANCE_MIS 1u
#define CRT_STRATEGY_LIGHT_ONLY 2u
#define CRT_STRATEGY_BSDF_ONLY 3u

typedef struct crt_render_settings {
    /* tracer.rs:639-661 */
    uint32_t width, height;
    uint32_t samples_per_pixel, max_depth, min_samples_per_pixel;
    float variance_threshold; /* 0 disables adaptive early stopping */
    int32_t frame;            /* sampler seed */
    uint32_t sampling_strategy;
    uint32_t adaptive; /* 1 = final-pass behaviour (tracer.rs:216-223) */
    uint32_t pixel_filter;
    float filter_radius;
    /* multi-GPU image split: square tiles of `tile_size` pixels, tile t rendered by the
     * process with tile_rank == t % tile_world. tile_world <= 1: the whole image. Pixels of other
     * ranks are left untouched in the framebuffer. */
    uint32_t tile_size, tile_rank, tile_world;
    /* wavefront pool: at most this many paths in flight (0 = library default). */
    uint32_t max_paths;
    /* crt_render_steal only: tiles per stolen chunk (0 = auto: about 4 chunks per process of tile_world). Chunk c = tiles
     * {c + k * n_chunks}: spread over the image. */
    uint32_t steal_chunk_tiles;
} crt_render_settings;

/* stats.rs:128-146 */
typedef struct crt_ray_stats {
    uint64_t camera_rays, closest_hit, shadow_rays, vertices, rr_tested, rr_killed, ended_escaped, ended_depth;
} crt_ray_stats;

typedef struct crt_world crt_world;

/* openpbr.rs:130-173 Default for OpenPBR */
void crt_material_default(crt_material* out);
/* World (rt_world.rs:142-242 WorldBuilder::commit) + LightList. Retains `scene`; inputs are copied.
 * geom_material[g] = index into `materials` of the material bound to geom_id g (n_geoms must equal
 * the scene's geometry count). */
int crt_world_new(crt_scene* scene, const crt_material* materials, size_t n_materials, const uint32_t* geom_material,
                  size_t n_geoms, const crt_light* lights, size_t n_lights, crt_world** out);
void crt_world_release(crt_world* w);
/* DomeLight::new(tint, Some(map), light_to_world) (light.rs:327-335, environment.rs:117-141): gives dome light
 * `light_index` of the world a lat-long environment map (`rgb`: width * height * 3 floats, row 0 = the +Y pole,
 * u = 0.5 looks down -Z) with luminance importance sampling. `light_to_world`: 9 floats, glam Mat3A column order. */
int crt_world_set_environment(crt_world* w, uint32_t light_index, uint32_t width, uint32_t height, const float* rgb,
                              const float light_to_world[9]);

/* volume.rs:22-48 DensityField */
#define CRT_FIELD_HOMOGENEOUS 0u /* density 1 inside the box */
#define CRT_FIELD_NOISE 1u       /* hash-based value-noise fBm ("smoke") */
#define CRT_FIELD_GRID 2u        /* voxel grid, x-fastest, trilinear, edge clamp */
#define CRT_MAX_VOLUMES 16u

/* VolumeRegion::new (volume.rs:194-231): an oriented box [-half, +half]^3 under `local_to_world` filled with a density
 * field. sigma_s / sigma_a are per-channel coefficients at density 1 BEFORE density_scale is folded in. */
typedef struct crt_volume {
    float local_to_world[16]; /* glam Mat4, column-major; must be affine and invertible */
    float half_extent[3];
    float sigma_s[3], sigma_a[3];
    float g; /* Henyey-Greenstein anisotropy, clamped to +-0.99 */
    float emission[3];
    float density_scale;
    uint32_t field; /* CRT_FIELD_* */
    float noise_scale;
    uint32_t noise_octaves;
    float noise_gain, noise_lacunarity, noise_threshold;
    uint32_t noise_seed;
    uint32_t grid_dims[3];
    const float* grid_data; /* HOST pointer, grid_dims[0]*[1]*[2] floats (copied) */
} crt_volume;
/* Scene::with_volumes (tracer.rs:167-170): replaces the world's volume regions (n = 0 clears them). Regions live
 * outside the BVH; the integrator delta-tracks through them on every segment and ratio-tracks shadow rays.
 * Any number of regions, like the reference's Vec (volume.rs:356-366). The one limit left is per ray: a single segment is
 * walked through at most CRT_MAX_VOLUMES regions that it crosses at once; regions beyond that are ignored for that
 * segment and bit 1 of crt_device_status() is set. */
int crt_world_set_volumes(crt_world* w, const crt_volume* volumes, size_t n);

/* Renderer::render_with_stats (tracer.rs:194-212): one non-guided pass. `framebuffer` is a DEVICE
 * pointer to width*height float4 {r, g, b, samples taken}; row j is image row j counted from the
 * bottom (buffer.rs:40-44; the reference flips on read, :74-77). Work is enqueued on `stream`;
 * the call synchronises that stream before returning because adaptive rounds are decided from
 * device-side counts. `stats` may be NULL. */
int crt_render(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings, float* framebuffer,
               crt_ray_stats* stats, void* stream);
/* Same with a HOST framebuffer (D2H inside). */
int crt_render_host(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings,
                    float* framebuffer, crt_ray_stats* stats);

/* Work stealing across the per-GPU processes (the reference balances tiles over rayon workers,
 * tracer.rs:423-468): one process creates a cursor in its device memory and hands the 64-byte CUDA IPC handle to
 * the others (e.g. a torch.distributed broadcast); all then call crt_render_steal on the same frame and pull
 * chunks of tiles off the cursor with system-scope atomics until it is exhausted. Whole tiles only, so per-pixel
 * sampler streams — and therefore the image — do not depend on who rendered what. The owner resets the cursor
 * between frames (with a barrier on either side). Pixels this process did not render are left untouched. */
typedef struct crt_steal_cursor crt_steal_cursor;
int crt_steal_cursor_create(crt_steal_cursor** out, unsigned char ipc_handle[64]);
int crt_steal_cursor_open(const unsigned char ipc_handle[64], crt_steal_cursor** out);
int crt_steal_cursor_reset(crt_steal_cursor* c);
void crt_steal_cursor_close(crt_steal_cursor* c);
int crt_render_steal(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings, float* framebuffer,
                     crt_ray_stats* stats, void* stream, crt_steal_cursor* cursor);

/* Peer-memory frame gather. The process that owns the final frame creates it here and hands the 64-byte CUDA
 * IPC handle to the other per-GPU processes; they map it and pass the mapped pointer to crt_render / crt_render_steal as
 * `framebuffer`, so the film kernel stores their tiles straight into the owner's HBM over NVLink / NVSwitch peer access.
 * Pixel ownership is disjoint, so there is nothing to reduce and no per-process full frame to zero: the frame is complete
 * on the owner once every process has returned from its render call (which synchronises its stream) and one barrier —
 * e.g. the small RayStats reduce — has passed. */
int crt_shared_frame_create(size_t bytes, void** device_ptr, unsigned char ipc_handle[64]);
int crt_shared_frame_open(const unsigned char ipc_handle[64], void** device_ptr);
int crt_shared_frame_close(void* device_ptr, int owner);

/* Kernel launches issued by the last crt_render* call of this thread, and the device time it took
 * (CUDA events on the render stream), for bench.py. */
int crt_render_last_info(uint64_t* launches, double* device_ms, uint64_t* rounds);

/* Test aid: the Material trait surface of ONE material evaluated on the device for n independent queries (host buffers,
 * copies inside), so the shading math can be checked against a witness that shares no source with it
 * (oracle/openpbr_ref.py, tests/test_gpu_shade_independent.py).
 *   crt_debug_material_eval    Material::eval (openpbr.rs:1138-1158): value_pdf[4i..] = brdf*|cos| (rgb), pdf; some[i] = Some?
 *   crt_debug_material_scatter Material::scatter_importance given the 4-D BSDF block s (openpbr.rs:1026-1136):
 *                              out[8i..] = direction (xyz), value (rgb), pdf, delta flag (0 / 1); some[i] = Some?
 * ray_dir, normal, wi: 3 floats per query; front_face: 1 byte per query. */
int crt_debug_material_eval(const crt_material* material, size_t n, const float* ray_dir, const float* normal, const uint8_t* front_face,
                            const float* wi, float* value_pdf, uint8_t* some);
int crt_debug_material_scatter(const crt_material* material, size_t n, const float* ray_dir, const float* normal, const uint8_t* front_face,
                               const float* s4, float* out8, uint8_t* some);

#ifdef __cplusplus
}
#endif
#endif /* CRT_RENDER_H */
