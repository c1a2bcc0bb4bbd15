// Wavefront path tracer for sm_100a (include/crt_render.h).
//
// Restates crust-core's integrator (crates/crust-core/src/tracer.rs:515-636 render_pixel,
// :1086-1558 trace_path) as a queue-driven wavefront instead of one recursive walk per pixel:
//
//   raygen  -> extend queue 0                      camera.rs:71-84, filter.rs, tracer.rs:559-584
//   loop k = 0 .. max_depth:
//     trace   closest-hit over extend queue k      traverse.cu (persistent BVH4 kernel), count read on device
//     shade   per hit: emission + MIS ownership, NEE sample -> shadow queue k, BSDF sample + Russian
//             roulette -> extend queue k+1         tracer.rs:1321-1530 (queues fed through global atomics)
//     trace   any-hit over shadow queue k
//     resolve unoccluded NEE contributions          tracer.rs:1416-1443
//   accumulate the finished samples per pixel IN SAMPLE ORDER, then the adaptive test every 4th
//   sample (tracer.rs:602-617) decides which pixels enter the next round.
//
// The reference records a VertexRec per vertex and gathers backwards (tracer.rs:1537-1556); here the
// same sum is accumulated forwards with the running throughput `beta` (identical terms, different
// float association). Path state lives in HBM as structure-of-arrays indexed by path slot; queues
// hold (ray, path slot) so the traversal kernels read contiguous 32-byte ray records. Nothing in a
// round synchronises with the host; the host reads one counter per adaptive round.
//
// Volume regions (volume.rs; tracer.rs:1166-1249 scatter vertex, :1040-1077 volume NEE, :1016-1035 shadow
// transmittance) ride the same queues: the delta-tracking walk runs inside shade, a region scatter vertex feeds the
// shadow and extend queues like a surface vertex, and nee_resolve ratio-tracks the unoccluded shadow rays. Those code
// paths are compiled into separate kernel instances (<VOL = true>) launched only for worlds that have regions.
// Out of scope here: path guiding, Ptex.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "crt_render.h"
#include "internal.h"
#include "shade/abi_convert.h"

using namespace shade;

namespace {

#define CKR(call)                                                  \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return crt::set_cuda_error(_e, #call); \
    } while (0)

constexpr int kBlock = 128;
constexpr int kShadeBins = 64;  // class 0 = miss, then one class per lobe signature of the world's materials (crt_world::mat_class)
#ifndef CRT_SHADE_BLOCK
#define CRT_SHADE_BLOCK 128  // threads per block of the shading kernels
#endif
#ifndef CRT_SHADE_SYNC
#define CRT_SHADE_SYNC 0  // 1: the warps of a block start every queue entry together (block-wide barrier per grid-stride step), so warps that share an SM sub-partition fetch the same instruction lines at about the same time
#endif
#ifndef CRT_SHADE_MINBLOCKS
#define CRT_SHADE_MINBLOCKS 8  // resident blocks per SM shade_kernel is compiled for: 64 registers with spills beat 128 registers at half the occupancy (measured 4/5/6/8 on B200)
#endif
constexpr uint32_t FLAG_HAS_PREV = 1u, FLAG_PREV_DELTA = 2u;  // PrevBounce, tracer.rs:898-905
constexpr float kTMin = 0.001f;                               // tracer.rs:1152 / 1029
// per-sample RayStats (stats.rs:128-146), packed into the path slot: x = closest_hit (12 bits) | shadow_rays (12) |
// rr_killed (1) | ended_escaped (1) | ended_depth (1); y = vertices (12) | rr_tested (12). max_depth <= 4095.
constexpr uint32_t CNT_SHADOW_SHIFT = 12, CNT_RR_KILLED = 1u << 24, CNT_ESCAPED = 1u << 25, CNT_DEPTH = 1u << 26, CNT_RR_TESTED_SHIFT = 12;

struct PathState {       // structure of arrays, one slot per path in flight
    float4* beta_pdf;    // throughput (tracer.rs:1113) + pdf of the bounce that produced the current ray
    float4* L_w;         // radiance accumulated so far + the sample's FIS weight wx*wy (tracer.rs:565-566)
    float4* prevp;       // previous vertex position (MIS `from`, tracer.rs:945) + flag word
    uint2* sampler;      // K_PATH domain pattern, sample index (tracer.rs:1101)
    float* time;         // shutter time of the path
    uint2* counts;       // RayStats of this sample, packed (see CNT_*): only accepted samples are counted
};
struct Queues {
    float4* rays[2];     // extend queue, double-buffered: 2 x float4 per ray (crt_ray)
    uint32_t* path[2];   // path slot of each queued ray
    float* time[2];      // shutter time per queued ray (null when nothing moves)
    float4* hits;        // closest-hit records of the current extend queue
    uint8_t* ff;         // front_face of the current extend queue
    float4* srays;       // shadow queue
    float* stime;
    uint32_t* spath;
    float4* scontrib;    // NEE contribution that lands if the shadow ray is unoccluded
    uint8_t* occ;
    // shading order of the current extend queue: its entries grouped by class (miss, then one class per material), so a
    // warp of shade threads runs one material's code; null = queue order
    uint32_t* order;
    uint32_t* bins;      // [class][sort block] counts, scanned in place into offsets
    // per wavefront, all zeroed by one memset: ext_count[D+2] | sh_count[D+2] | ext_fetch[D+2] | sh_fetch[D+2]
    uint32_t* counts;
};
struct PixelState {      // one entry per pixel this process renders (position in `pixels`)
    const uint32_t* pixels;  // pixel id = j * width + i
    float4* sum_w;           // Σ w·L, Σ w (tracer.rs:526-531)
    double2* lum;            // Σ lum, Σ lum² in f64 (tracer.rs:533-534)
    uint32_t* taken;         // samples accumulated so far; bit 31 = the adaptive test has stopped this pixel
};
constexpr uint32_t PX_DONE = 0x80000000u;
struct RenderParams {
    Camera cam;
    Filter filter;
    RenderSettings st;
    const Material* materials;
    const uint32_t* geom_material;
    const uint8_t* mat_class;   // shading class of each material (1 .. kShadeBins - 1), see material_signature
    const Light* lights;
    const int32_t* geom_light;  // LightList::find_by_geom (light.rs:436-440) as a table
    const EnvMap* envs;         // environment maps of dome lights, indexed by Light::env
    const VolumeRegion* volumes;  // Volumes (volume.rs:364-384), outside the BVH
    int n_volumes;
    int n_lights;
    int motion;
    PathState ps;
    Queues q;
    PixelState px;
    unsigned long long* stats;  // crt_ray_stats
    unsigned int* status;       // sticky device status word (bit 1: a ray segment crossed more volume regions than it can walk)
};

__device__ __forceinline__ V3 xyz(float4 a) { return v3(a.x, a.y, a.z); }
__device__ __forceinline__ float4 f4(V3 a, float w) { return make_float4(a.x, a.y, a.z, w); }

// Warp-aggregated queue slot allocation (one global atomic per warp per push site).
__device__ __forceinline__ uint32_t queue_push(uint32_t* counter) {
    unsigned m = __activemask();
    unsigned lane = threadIdx.x & 31u;
    int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
__device__ __forceinline__ void stat_add(unsigned long long* dst, uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

// ---- raygen (tracer.rs:543-584) --------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) raygen_kernel(const RenderParams P, const uint32_t* __restrict__ active, uint32_t nb, uint32_t s_base,
                                                        uint32_t s_count) {
    const uint32_t total = nb * s_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.q.counts[0] = total;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        uint32_t a = p % nb, sl = p / nb;  // sample-major: a warp covers neighbouring pixels of one sample index
        uint32_t pixel = P.px.pixels[active[a]];
        uint32_t i = pixel % P.st.width, j = pixel / P.st.width;
        uint32_t sample = s_base + sl;
        int32_t tile = (int32_t)(i >> 8) + (int32_t)(j >> 8) * 4096;
        PathSampler root = new_domain(sampler_new((int32_t)i, (int32_t)j, P.st.frame, (int32_t)sample), tile);
        float cs[4];
        draw_sample4(new_domain(root, K_CAMERA), cs);
        float fx, wx, fy, wy;
        filter_sample(P.filter, cs[0], fx, wx);
        filter_sample(P.filter, cs[1], fy, wy);
        float u = ((float)i + fx) / (float)P.st.width;
        float v = ((float)j + fy) / (float)P.st.height;
        float time = 0.0f;
        if (P.motion) {
            float ts[4];
            draw_sample4(new_domain(root, K_TIME), ts);
            time = ts[0];
        }
        V3 o, d;
        camera_get_ray(P.cam, u, v, cs[2], cs[3], o, d);
        P.q.rays[0][2 * (size_t)p] = f4(o, kTMin);
        P.q.rays[0][2 * (size_t)p + 1] = f4(d, INFINITY);
        P.q.path[0][p] = p;
        if (P.q.time[0]) P.q.time[0][p] = time;
        PathSampler path = new_domain(root, K_PATH);
        P.ps.beta_pdf[p] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        P.ps.L_w[p] = make_float4(0.0f, 0.0f, 0.0f, wx * wy);
        P.ps.prevp[p] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
        P.ps.sampler[p] = make_uint2(path.pattern, path.index);
        P.ps.time[p] = time;
        P.ps.counts[p] = make_uint2(P.st.max_depth == 0 ? CNT_DEPTH : 0u, 0u);
    }
}

// bounce_emission_weight, tracer.rs:930-953. A surface bounce that produced a sample always has an
// evaluable material (scatter and eval share the `v_local.z <= 0` gate, openpbr.rs:1033-1037,
// 1148-1151), so "eval is None" reduces to the delta flag.
__device__ __forceinline__ float bounce_emission_weight(const RenderParams& P, uint32_t flags, float prev_pdf, V3 prev_p, uint32_t geom_id, V3 hit_p) {
    if (flags & FLAG_PREV_DELTA) return 1.0f;
    int li = P.geom_light[geom_id];
    if (li < 0) return 1.0f;
    float light_pdf = fmaxf(light_pdf_at_point(P.lights[li], prev_p, hit_p) / (float)P.n_lights, 1e-6f);
    return strategy_bounce_weight(P.st.sampling_strategy, prev_pdf, light_pdf);
}

// ---- shade: one path vertex per queue entry (tracer.rs:1118-1535) --------------------------------
// ---- shading order: a counting sort of the extend queue by class ------------------------------------------------
// Threads of a warp that shade different materials (or a miss next to a hit) execute each other's code paths one after the
// other; OpenPBR's lobes are thousands of instructions each. Grouping the queue by class costs two light passes over the
// hit records and makes most warps uniform. Only the ORDER in which entries are shaded changes — every entry is keyed by
// its path slot, so images and RayStats are unaffected.
__device__ __forceinline__ uint32_t shade_class(const RenderParams& P, uint32_t i) {
    const uint32_t geom_id = __float_as_uint(P.q.hits[2 * (size_t)i].w);
    if (geom_id == CRT_INVALID_ID) return 0u;
    return P.mat_class[P.geom_material[geom_id]];
}
// A STABLE counting sort: entries of one class keep their queue order, which is also the order in which their continuation
// rays enter the next queue — neighbouring paths stay neighbours there, and the traversal kernel depends on that coherence
// (an unstable scatter by atomics cost the stress scene 8 %). Each block owns one contiguous chunk of the queue.
constexpr int kSortBlocks = 1024;
__device__ __forceinline__ void sort_chunk(uint32_t n, uint32_t& begin, uint32_t& end) {
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + kBlock - 1) / kBlock * kBlock;
    begin = min(n, blockIdx.x * per);
    end = min(n, begin + per);
}
// pass 1: per-block class counts -> bins[class * gridDim + block]
__global__ void __launch_bounds__(kBlock) classify_count_kernel(const RenderParams P, int k, int nc) {  // nc = classes in use
    __shared__ uint32_t cnt[kShadeBins];
    for (int c = threadIdx.x; c < nc; c += kBlock) cnt[c] = 0;
    __syncthreads();
    uint32_t begin, end;
    sort_chunk(P.q.counts[k], begin, end);
    for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock) {
        const uint32_t c = shade_class(P, i);
        const unsigned peers = __match_any_sync(__activemask(), c);
        if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&cnt[c], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nc; c += kBlock) P.q.bins[(size_t)c * gridDim.x + blockIdx.x] = cnt[c];
}
// pass 2: exclusive scan of the (class-major, block-minor) counts, in place; one block
__global__ void __launch_bounds__(1024) classify_scan_kernel(uint32_t* bins, uint32_t total) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (total + 1023u) / 1024u;
    const uint32_t lo = min(total, threadIdx.x * per), hi = min(total, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += bins[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        uint32_t v = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t v = bins[i];
        bins[i] = run;
        run += v;
    }
}
// pass 3: every entry goes to (its block's offset for its class) + (its rank among the block's earlier entries of the class)
__global__ void __launch_bounds__(kBlock) classify_scatter_kernel(const RenderParams P, int k, int nc) {
    __shared__ uint32_t running[kShadeBins];            // entries of each class this block has placed so far
    __shared__ uint32_t warp_cnt[kBlock / 32][kShadeBins];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < nc; c += kBlock) running[c] = P.q.bins[(size_t)c * gridDim.x + blockIdx.x];
    uint32_t begin, end;
    sort_chunk(P.q.counts[k], begin, end);
    for (uint32_t base_i = begin; base_i < end; base_i += kBlock) {  // uniform trip count: the barriers below are safe
        for (int c = threadIdx.x; c < (kBlock / 32) * nc; c += kBlock) warp_cnt[c / nc][c % nc] = 0;
        __syncthreads();
        const uint32_t i = base_i + threadIdx.x;
        const bool live = i < end;
        uint32_t c = 0, rank = 0;
        if (live) {
            c = shade_class(P, i);
            const unsigned peers = __match_any_sync(__activemask(), c);
            rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if ((int)lane == __ffs(peers) - 1) warp_cnt[warp][c] = (uint32_t)__popc(peers);
        }
        __syncthreads();
        if (live) {
            uint32_t before = running[c];
            for (unsigned w = 0; w < warp; ++w) before += warp_cnt[w][c];
            P.q.order[before + rank] = i;
        }
        __syncthreads();
        for (int cc = threadIdx.x; cc < nc; cc += kBlock) {
            uint32_t add = 0;
            for (int w = 0; w < kBlock / 32; ++w) add += warp_cnt[w][cc];
            running[cc] += add;
        }
        __syncthreads();
    }
}

// PASS selects what one launch does with a vertex: SHADE_FUSED everything (worlds with volume regions); SHADE_VERTEX
// everything except the BSDF sampling of a surface vertex; SHADE_SCATTER only that — BSDF sample, Russian roulette,
// continuation ray (tracer.rs:1459-1530). The fused kernel is 230 KB of machine code and ncu attributes 44 % of its stalls to
// instruction fetch (`stall_no_inst`): light sampling + material_eval and material_scatter are each a few thousand
// instructions, and a warp walking both overruns the instruction caches. Two launches over the same queue, each with
// half the code, cost a second read of the hit and path records and no arithmetic twice except the segment attenuation.
enum : int { SHADE_FUSED = 0, SHADE_VERTEX = 1, SHADE_SCATTER = 2 };

template <bool VOL, int PASS>
__global__ void __launch_bounds__(CRT_SHADE_BLOCK, (VOL ? 4 : CRT_SHADE_MINBLOCKS) * 128 / CRT_SHADE_BLOCK) shade_kernel(const RenderParams P, int k, int cur, const uint32_t* __restrict__ order) {
    static_assert(!VOL || PASS == SHADE_FUSED, "worlds with volume regions use the fused kernel");
    const int D = (int)P.st.max_depth;
    const uint32_t n = P.q.counts[k];
    uint32_t* next_count = P.q.counts + k + 1;
    uint32_t* sh_count = P.q.counts + (D + 2) + k;
    const int nxt = cur ^ 1;

#if CRT_SHADE_SYNC
    for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += gridDim.x * blockDim.x) {  // uniform trip count: the barrier is safe
        __syncthreads();
        const uint32_t j = j0 + threadIdx.x;
        if (j >= n) continue;
#else
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
#endif
        const uint32_t i = order ? order[j] : j;
        const uint32_t path = P.q.path[cur][i];
        const float4 ra = P.q.rays[cur][2 * (size_t)i], rb = P.q.rays[cur][2 * (size_t)i + 1];
        const V3 ro = xyz(ra), rd = xyz(rb);
        const float4 h0 = P.q.hits[2 * (size_t)i], h1 = P.q.hits[2 * (size_t)i + 1];
        const uint32_t geom_id = __float_as_uint(h0.w);
        const bool have_hit = geom_id != CRT_INVALID_ID;
        const float t_surf = h0.x;
        const float4 bp = P.ps.beta_pdf[path];
        V3 beta = xyz(bp);
        const float prev_pdf = bp.w;
        const float4 pp = P.ps.prevp[path];
        const uint32_t flags = __float_as_uint(pp.w);
        const bool has_prev = (flags & FLAG_HAS_PREV) != 0;
        const int medium = (int)((flags >> 8) & 0xFFFFu) - 1;
        const uint2 smp = P.ps.sampler[path];
        const PathSampler v = new_domain(PathSampler{smp.x, smp.y}, k);
        const float time = P.ps.time[path];
        V3 add = splat(0.0f);
        [[maybe_unused]] bool vol_overflow = false;
        uint2 cnt = P.ps.counts[path];
        if (PASS != SHADE_SCATTER) cnt.x += 1u;  // this entry is one traced closest-hit query (stats.rs:133)
        const Medium* med = medium >= 0 ? &P.materials[medium].medium : nullptr;

        if (k >= D) {
            if (PASS == SHADE_SCATTER) continue;
            // Depth exhausted (tracer.rs:1123-1148): the ray was traced for bounce-hit emission only.
            cnt.x |= CNT_DEPTH;
            if (have_hit) {
                const Material& mat = P.materials[P.geom_material[geom_id]];
                V3 normal = v3(h1.y, h1.z, h1.w);
                float cos_o = fabsf(dot(normalize(rd), normal));
                V3 emitted = material_emitted_directional(mat, cos_o);
                if (length_squared(emitted) > 0.0f) {
                    if (med) emitted *= medium_transmittance(*med, t_surf);
                    if (VOL) {  // tracer.rs:1138-1141
                        Rng rng = sampler_rng(new_domain(v, K_VOLUME));
                        emitted *= volumes_transmittance(P.volumes, P.n_volumes, ro, rd, kTMin, t_surf, rng, &vol_overflow);
                    }
                    float w = bounce_emission_weight(P, flags, prev_pdf, xyz(pp), geom_id, ro + t_surf * rd);
                    add = beta * (emitted * w);
                }
            }
        } else {
            // Free flight in a scattering carried medium (tracer.rs:1158-1164).
            float t_med = INFINITY;
            if (med && medium_is_scattering(*med)) {
                float sigma_t_max = fmaxf(medium_sigma_t_max(*med), 1e-4f);
                t_med = -logf(draw_rnd1(new_domain(v, K_MEDIUM))) / sigma_t_max;
            }
            bool cont = false;      // does the path continue with a new extend ray?
            V3 no = splat(0.0f), nd = splat(0.0f);
            float new_pdf = 0.0f;
            V3 new_prevp = splat(0.0f);
            uint32_t new_flags = 0;
            // Volume-region interaction, clipped to whatever event would otherwise end the segment (tracer.rs:1166-1182).
            V3 vol_tr = splat(1.0f);
            bool vol_scatter = false;
            if (VOL) {
                Rng rng = sampler_rng(new_domain(v, K_VOLUME));
                VolumeEvent ev;
                volumes_sample_interaction(P.volumes, P.n_volumes, ro, rd, kTMin, fminf(t_surf, t_med), rng, ev, &vol_overflow);
                add = beta * ev.emitted;  // segment_emit: already weighted by the walk, not attenuated again (:1552)
                if (ev.scatter) {
                    // Volume-region scatter vertex (tracer.rs:1184-1244): NEE with the phase function in place of brdf*cos
                    // (volume_nee, :1040-1077), continuation factor ONE (the HG value and pdf cancel), prev = Phase.
                    vol_scatter = true;
                    const V3 wi = normalize(rd);
                    float ps[4];
                    draw_sample4(new_domain(v, K_PHASE), ps);
                    V3 dir = phase_sample(ev.phase, wi, ps[0], ps[1], ps[2]);
                    float phase_pdf_dir = fmaxf(phase_pdf(ev.phase, dot(wi, dir)), 1e-6f);
                    beta *= ev.weight;
                    if (strategy_samples_lights(P.st.sampling_strategy)) {
                        float nee_s[4];
                        draw_sample4(new_domain(v, K_NEE), nee_s);
                        int li = light_pick(P.n_lights, nee_s[0]);
                        LightSample ls;
                        if (li >= 0 && light_sample_li(P.lights[li], ev.p, nee_s[1], nee_s[2], ls, P.envs)) {
                            float light_pdf = fmaxf(ls.pdf / (float)P.n_lights, 1e-6f);
                            float phase_val = phase_pdf(ev.phase, dot(wi, ls.direction));
                            float weight = strategy_light_weight(P.st.sampling_strategy, light_pdf, phase_val);
                            V3 contrib = beta * (ls.radiance * phase_val * weight / light_pdf);  // x shadow transmittance in nee_resolve
                            cnt.x += 1u << CNT_SHADOW_SHIFT;
                            uint32_t s = queue_push(sh_count);
                            P.q.srays[2 * (size_t)s] = f4(ev.p, kTMin);
                            P.q.srays[2 * (size_t)s + 1] = f4(ls.direction, ls.distance - 0.001f);
                            if (P.q.stime) P.q.stime[s] = time;
                            P.q.spath[s] = path;
                            P.q.scontrib[s] = f4(contrib, 0.0f);
                        }
                    }
                    bool survived = true;
                    if (k >= RR_START_BOUNCE) {
                        cnt.y += 1u << CNT_RR_TESTED_SHIFT;
                        float p_survive = clampf(max_element(beta), RR_MIN_PROB, 1.0f);
                        if (p_survive < 1.0f) {
                            if (draw_rnd1(new_domain(v, K_RR)) >= p_survive) {
                                survived = false;
                                cnt.x |= CNT_RR_KILLED;
                            } else {
                                beta = beta / p_survive;
                            }
                        }
                    }
                    cnt.y += 1u;
                    if (survived) {
                        cont = true;
                        no = ev.p;
                        nd = dir;
                        new_pdf = phase_pdf_dir;
                        new_prevp = ev.p;
                        new_flags = FLAG_HAS_PREV | ((uint32_t)(medium + 1) << 8);  // the carried medium is preserved (:1236-1240)
                    }
                } else {
                    vol_tr = ev.weight;
                }
            }

            if (vol_scatter) {
                // the region scatter vertex above was this segment's event: it ends before the surface and the carried medium's
                // free flight (the walk was clipped to both, tracer.rs:1172-1173)
            } else if (t_med < t_surf) {
                if (PASS == SHADE_SCATTER) continue;
                // Carried-medium scatter vertex (tracer.rs:1251-1318): no NEE, prev = None afterwards.
                float sigma_bar = fmaxf(medium_sigma_t_max(*med), 1e-4f);
                V3 pos = ro + t_med * rd;
                float ph[4];
                draw_sample4(new_domain(v, K_PHASE), ph);
                V3 dir = sample_henyey_greenstein(normalize(rd), med->g, ph[0], ph[1]);
                V3 e = (splat(sigma_bar) - (med->sigma_a + med->sigma_s)) * t_med;
                V3 factor = med->sigma_s / sigma_bar * v3(expf(e.x), expf(e.y), expf(e.z));
                beta *= vol_tr * factor;
                bool survived = true;
                if (k >= RR_START_BOUNCE) {
                    cnt.y += 1u << CNT_RR_TESTED_SHIFT;
                    float p_survive = clampf(max_element(beta), RR_MIN_PROB, 1.0f);
                    if (p_survive < 1.0f) {
                        if (draw_rnd1(new_domain(v, K_RR)) >= p_survive) {
                            survived = false;
                            cnt.x |= CNT_RR_KILLED;
                        } else {
                            beta = beta / p_survive;
                        }
                    }
                }
                cnt.y += 1u;
                if (survived) {
                    if (k + 1 >= D) {
                        cnt.x |= CNT_DEPTH;  // remaining <= 0 with prev = None: the reference ends without tracing (:1123-1128)
                    } else {
                        cont = true;
                        no = pos;
                        nd = dir;
                        new_flags = (uint32_t)(medium + 1) << 8;
                    }
                }
            } else if (!have_hit) {
                if (PASS == SHADE_SCATTER) continue;
                // Background (tracer.rs:1321-1342) with escaped_emission (:966-1009).
                cnt.x |= CNT_ESCAPED;
                V3 unit = normalize(rd);
                V3 background = splat(0.0f);
                bool covered = false;
                if (P.n_lights > 0) {
                    bool competing = has_prev && !(flags & FLAG_PREV_DELTA);
                    for (int l = 0; l < P.n_lights; ++l) {
                        V3 emitted;
                        float pdf;
                        if (!light_escaped(P.lights[l], unit, emitted, pdf, P.envs)) continue;
                        covered = true;
                        float w = 1.0f;
                        if (competing && strategy_samples_lights(P.st.sampling_strategy))
                            w = strategy_bounce_weight(P.st.sampling_strategy, prev_pdf, fmaxf(pdf / (float)P.n_lights, 1e-6f));
                        background += emitted * w;
                    }
                }
                if (!covered) {
                    float t = 0.5f * (unit.y + 1.0f);
                    background += (1.0f - t) * v3(1.0f, 1.0f, 1.0f) + t * v3(0.5f, 0.7f, 1.0f);
                }
                add += beta * (vol_tr * background);
            } else {
                // Surface vertex.
                const uint32_t mat_index = P.geom_material[geom_id];
                const Material& mat = P.materials[mat_index];
                SurfaceHit rec;
                rec.p = ro + t_surf * rd;  // rt_world.rs:221
                rec.normal = v3(h1.y, h1.z, h1.w);
                rec.t = t_surf;
                rec.front_face = P.q.ff[i] != 0;
                // attenuation over the arriving segment (tracer.rs:1352-1361)
                V3 atten = splat(1.0f);
                if (med) {
                    if (medium_is_scattering(*med)) {
                        float sigma_bar = fmaxf(medium_sigma_t_max(*med), 1e-4f);
                        V3 e = (splat(sigma_bar) - (med->sigma_a + med->sigma_s)) * rec.t;
                        atten = v3(expf(e.x), expf(e.y), expf(e.z));
                    } else {
                        atten = medium_transmittance(*med, rec.t);
                    }
                }
                atten = vol_tr * atten;  // tracer.rs:1362
                if (PASS != SHADE_SCATTER) {
                // emission ownership (tracer.rs:1369-1381)
                float cos_o = fabsf(dot(normalize(rd), rec.normal));
                V3 emitted = material_emitted_directional(mat, cos_o);
                if (has_prev) {
                    if (length_squared(emitted) > 0.0f) {
                        float w = bounce_emission_weight(P, flags, prev_pdf, xyz(pp), geom_id, rec.p);
                        add += beta * ((atten * emitted) * w);
                    }
                } else {
                    add += beta * (atten * emitted);
                }
                // 1. light sampling (tracer.rs:1394-1445): the shadow ray is queued with the contribution
                //    it would deliver; every sampled light costs one occlusion query like the reference.
                float nee_s[4];
                draw_sample4(new_domain(v, K_NEE), nee_s);
                int li = strategy_samples_lights(P.st.sampling_strategy) ? light_pick(P.n_lights, nee_s[0]) : -1;
                LightSample ls;
                if (li >= 0 && light_sample_li(P.lights[li], rec.p, nee_s[1], nee_s[2], ls, P.envs)) {
                    V3 contrib = splat(0.0f);
                    float cosine = fabsf(dot(rec.normal, ls.direction));
                    float light_pdf = fmaxf(ls.pdf / (float)P.n_lights, 1e-6f);
                    V3 brdf_value;
                    float brdf_pdf;
                    if (material_eval(mat, rd, rec, ls.direction, brdf_value, brdf_pdf)) {
                        float weight = strategy_light_weight(P.st.sampling_strategy, light_pdf, brdf_pdf);
                        V3 nee = ls.radiance * brdf_value * cosine * weight / light_pdf;
                        contrib = beta * (atten * nee);
                    }
                    cnt.x += 1u << CNT_SHADOW_SHIFT;
                    uint32_t s = queue_push(sh_count);
                    P.q.srays[2 * (size_t)s] = f4(rec.p, kTMin);
                    P.q.srays[2 * (size_t)s + 1] = f4(ls.direction, ls.distance - 0.001f);
                    if (P.q.stime) P.q.stime[s] = time;
                    P.q.spath[s] = path;
                    P.q.scontrib[s] = f4(contrib, 0.0f);
                }
                cnt.y += 1u;  // stats.vertices: counted whether or not the path goes on (:1531-1533)
                }
                if (PASS != SHADE_VERTEX) {
                // 2. BSDF sampling + Russian roulette (tracer.rs:1459-1530)
                float bs[4];
                draw_sample4(new_domain(v, K_BSDF), bs);
                ScatterSample sample;
                if (material_scatter(mat, rd, rec, bs, sample)) {
                    V3 dir = normalize(sample.dir);
                    float cosine = sample.delta ? 1.0f : fabsf(dot(rec.normal, dir));
                    V3 factor = sample.value * cosine / sample.pdf;
                    beta *= atten * factor;
                    bool survived = true;
                    if (k >= RR_START_BOUNCE) {
                        cnt.y += 1u << CNT_RR_TESTED_SHIFT;
                        float p_survive = clampf(max_element(beta), RR_MIN_PROB, 1.0f);
                        if (p_survive < 1.0f) {
                            if (draw_rnd1(new_domain(v, K_RR)) >= p_survive) {
                                survived = false;
                                cnt.x |= CNT_RR_KILLED;
                            } else {
                                beta = beta / p_survive;
                            }
                        }
                    }
                    if (survived) {
                        cont = true;
                        no = sample.origin;
                        nd = sample.dir;
                        new_pdf = sample.pdf;
                        new_prevp = rec.p;
                        new_flags = FLAG_HAS_PREV | (sample.delta ? FLAG_PREV_DELTA : 0u) | (sample.enters_medium ? (mat_index + 1u) << 8 : 0u);
                    }
                }
                }
            }
            if (cont) {
                P.ps.beta_pdf[path] = f4(beta, new_pdf);
                P.ps.prevp[path] = f4(new_prevp, __uint_as_float(new_flags));
                uint32_t s = queue_push(next_count);
                P.q.rays[nxt][2 * (size_t)s] = f4(no, kTMin);
                P.q.rays[nxt][2 * (size_t)s + 1] = f4(nd, INFINITY);
                P.q.path[nxt][s] = path;
                if (P.q.time[nxt]) P.q.time[nxt][s] = time;
            }
        }
        P.ps.counts[path] = cnt;
        if (VOL && vol_overflow) atomicOr(P.status, 2u);
        if (add.x != 0.0f || add.y != 0.0f || add.z != 0.0f) {
            float4 lw = P.ps.L_w[path];
            P.ps.L_w[path] = make_float4(lw.x + add.x, lw.y + add.y, lw.z + add.z, lw.w);
        }
    }
}

// ---- NEE resolve: unoccluded contributions land on their path (tracer.rs:1416-1443) ---------------
template <bool VOL>
__global__ void __launch_bounds__(kBlock) nee_resolve_kernel(const RenderParams P, int k) {
    const int D = (int)P.st.max_depth;
    const uint32_t n = P.q.counts[(D + 2) + k];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (P.q.occ[i]) continue;
        float4 c = P.q.scontrib[i];
        if (c.x == 0.0f && c.y == 0.0f && c.z == 0.0f) continue;
        uint32_t path = P.q.spath[i];
        if (VOL) {
            // shadow_transmittance (tracer.rs:1030-1034): the unoccluded segment pays the regions it crosses, walked with
            // the vertex's K_NEE_SHADOW stream
            const uint2 smp = P.ps.sampler[path];
            Rng rng = sampler_rng(new_domain(new_domain(PathSampler{smp.x, smp.y}, k), K_NEE_SHADOW));
            const float4 ra = P.q.srays[2 * (size_t)i], rb = P.q.srays[2 * (size_t)i + 1];
            bool vol_overflow = false;
            V3 tr = volumes_transmittance(P.volumes, P.n_volumes, xyz(ra), xyz(rb), kTMin, rb.w, rng, &vol_overflow);
            if (vol_overflow) atomicOr(P.status, 2u);
            c.x *= tr.x;
            c.y *= tr.y;
            c.z *= tr.z;
        }
        float4 lw = P.ps.L_w[path];
        P.ps.L_w[path] = make_float4(lw.x + c.x, lw.y + c.y, lw.z + c.z, lw.w);
    }
}

// ---- film: add the finished samples of each pixel in sample order and apply the adaptive stop
// (tracer.rs:599-617). A round may hold more samples than the reference would have taken for a pixel: the
// test runs after every 4th sample exactly where the reference runs it, and samples past the stopping point
// are discarded — neither their radiance nor their RayStats count — so results do not depend on the round size.
__global__ void __launch_bounds__(kBlock) accumulate_kernel(const RenderParams P, const uint32_t* __restrict__ active, uint32_t nb, uint32_t s_count) {
    const double threshold = (double)P.st.variance_threshold;
    const bool adaptive = P.st.adaptive && threshold > 0.0;
    const uint32_t min_spp = max(P.st.min_samples_per_pixel, 2u);
    uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < nb; a += gridDim.x * blockDim.x) {
        uint32_t pos = active[a];
        float4 sw = P.px.sum_w[pos];
        double2 lum = P.px.lum[pos];
        uint32_t taken = P.px.taken[pos];
        bool done = (taken & PX_DONE) != 0;
        taken &= ~PX_DONE;
        for (uint32_t sl = 0; sl < s_count && !done; ++sl) {
            const size_t p = (size_t)sl * nb + a;
            float4 lw = P.ps.L_w[p];
            V3 color = xyz(lw) * lw.w;
            sw.x += color.x;
            sw.y += color.y;
            sw.z += color.z;
            sw.w += lw.w;
            double l = (double)luminance(color);
            lum.x += l;
            lum.y += l * l;
            ++taken;
            uint2 cnt = P.ps.counts[p];
            c[0] += 1u;
            c[1] += cnt.x & 0xFFFu;
            c[2] += (cnt.x >> CNT_SHADOW_SHIFT) & 0xFFFu;
            c[3] += cnt.y & 0xFFFu;
            c[4] += (cnt.y >> CNT_RR_TESTED_SHIFT) & 0xFFFu;
            c[5] += (cnt.x & CNT_RR_KILLED) ? 1u : 0u;
            c[6] += (cnt.x & CNT_ESCAPED) ? 1u : 0u;
            c[7] += (cnt.x & CNT_DEPTH) ? 1u : 0u;
            if (taken >= P.st.samples_per_pixel) {
                done = true;
            } else if (adaptive && taken >= min_spp && taken % 4u == 0u) {
                double nn = (double)taken;
                double var_of_mean = fmax((lum.y - lum.x * lum.x / nn) / (nn - 1.0) / nn, 0.0);
                double mean = fmax(lum.x / nn, 1e-4);
                done = sqrt(var_of_mean) / mean < threshold;
            }
        }
        P.px.sum_w[pos] = sw;
        P.px.lum[pos] = lum;
        P.px.taken[pos] = taken | (done ? PX_DONE : 0u);
    }
    for (int k = 0; k < 8; ++k) stat_add(P.stats + k, c[k]);
}

// ---- which pixels enter the next round ----------------------------------------------------------
__global__ void __launch_bounds__(kBlock) adaptive_kernel(const RenderParams P, const uint32_t* __restrict__ active_in, uint32_t n_in, uint32_t* __restrict__ active_out,
                                                          uint32_t* __restrict__ count_out) {
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_in; a += gridDim.x * blockDim.x) {
        uint32_t pos = active_in[a];
        if (!(P.px.taken[pos] & PX_DONE)) active_out[queue_push(count_out)] = pos;
    }
}

__global__ void iota_kernel(uint32_t* __restrict__ out, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

// ---- resolve the film (tracer.rs:627-634) ------------------------------------------------------------
// `packed`: pixel `pos` of this process' pixel list goes to fb[pos] (the host path downloads only the pixels it owns and
// scatters them on the host); otherwise to its place in the full frame.
__global__ void __launch_bounds__(kBlock) film_kernel(const RenderParams P, uint32_t n, float4* __restrict__ fb, int packed) {
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x) {
        float4 sw = P.px.sum_w[pos];
        uint32_t taken = P.px.taken[pos] & ~PX_DONE;
        V3 mean = sw.w > 0.0f ? xyz(sw) / sw.w : xyz(sw) / (float)taken;
        fb[packed ? pos : P.px.pixels[pos]] = f4(mean, (float)taken);
    }
}

// crt_steal_cursor: one u32 in device memory of the creating process; peers map it through CUDA IPC and bump
// it with system-scope atomics (over NVLink / NVSwitch peer access).
__global__ void fetch_chunk_kernel(unsigned int* cursor, uint32_t* out) { *out = atomicAdd_system(cursor, 1u); }

// ---- test aid: one material's eval / scatter for n queries (crt_debug_material_*) ----------------------------------------
__global__ void debug_material_kernel(Material m, uint32_t n, const float* __restrict__ ray_dir, const float* __restrict__ normal,
                                      const uint8_t* __restrict__ front_face, const float* __restrict__ arg, float* __restrict__ out,
                                      uint8_t* __restrict__ some, int scatter) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SurfaceHit rec;
    rec.p = v3(0.0f, 0.0f, 0.0f);
    rec.normal = v3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
    rec.t = 1.0f;
    rec.front_face = front_face[i] != 0;
    const V3 rd = v3(ray_dir[3 * i], ray_dir[3 * i + 1], ray_dir[3 * i + 2]);
    if (!scatter) {
        V3 value = v3(0.0f, 0.0f, 0.0f);
        float pdf = 0.0f;
        bool ok = material_eval(m, rd, rec, v3(arg[3 * i], arg[3 * i + 1], arg[3 * i + 2]), value, pdf);
        some[i] = ok ? 1 : 0;
        out[4 * i] = value.x, out[4 * i + 1] = value.y, out[4 * i + 2] = value.z, out[4 * i + 3] = pdf;
    } else {
        const float s4[4] = {arg[4 * i], arg[4 * i + 1], arg[4 * i + 2], arg[4 * i + 3]};
        ScatterSample ss{};
        bool ok = material_scatter(m, rd, rec, s4, ss);
        some[i] = ok ? 1 : 0;
        float* o = out + 8 * (size_t)i;
        o[0] = ss.dir.x, o[1] = ss.dir.y, o[2] = ss.dir.z, o[3] = ss.value.x, o[4] = ss.value.y, o[5] = ss.value.z, o[6] = ss.pdf, o[7] = ss.delta ? 1.0f : 0.0f;
    }
}

struct Scratch {  // grow-only device buffers reused across crt_render calls on one world
    size_t cap_paths = 0, cap_pixels = 0, cap_counts = 0;
    void* paths = nullptr;   // everything sized by the path pool
    void* pixels = nullptr;  // everything sized by the pixel count
    uint32_t* counts = nullptr;
    unsigned long long* stats = nullptr;
    uint32_t* active_count = nullptr;  // 2 words
    uint32_t* h_pinned = nullptr;      // pinned host mirror for small read-backs
    void* host_fb = nullptr;           // crt_render_host: the device-side frame (whole frame, or this process' pixels packed)
    size_t cap_host_fb = 0;
    char* h_stage = nullptr;           // crt_render_host: pinned staging between the device frame and the caller's memory
    size_t cap_stage = 0;
    std::vector<uint32_t> h_pixels;    // crt_render_host with a tile split: host copy of the pixel list (scatter order)
    cudaStream_t host_stream = nullptr;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // the static pixel list of the last frame stays on the device: same image size and tile deal => no host work.
    // The deal is remembered as the real tuple (no hash: distinct deals must never alias).
    struct PixelDeal {
        uint32_t w = 0, h = 0, tile = 0, rank = 0, world = 0;
        bool operator==(const PixelDeal& o) const { return w == o.w && h == o.h && tile == o.tile && rank == o.rank && world == o.world; }
    };
    PixelDeal pixel_deal, h_deal;  // of the device pixel list / of h_pixels
    bool pixel_deal_valid = false, h_deal_valid = false;
    uint32_t pixel_count = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~Scratch() {
        cudaFree(paths);
        cudaFree(pixels);
        cudaFree(counts);
        cudaFree(stats);
        cudaFree(active_count);
        cudaFree(host_fb);
        if (h_stage) cudaFreeHost(h_stage);
        if (host_stream) cudaStreamDestroy(host_stream);
        for (cudaEvent_t e : stage_ev)
            if (e) cudaEventDestroy(e);
        if (h_pinned) cudaFreeHost(h_pinned);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

thread_local uint64_t t_last_launches = 0, t_last_rounds = 0, t_last_chunks = 0;
thread_local double t_last_ms = 0.0;

}  // namespace

struct crt_steal_cursor {
    unsigned int* dev = nullptr;
    bool owner = false;
};

struct crt_world {
    crt_scene* scene = nullptr;
    std::vector<Material> materials;
    std::vector<uint32_t> geom_material;
    std::vector<Light> lights;
    Material* d_materials = nullptr;
    uint32_t* d_geom_material = nullptr;
    Light* d_lights = nullptr;
    int32_t* d_geom_light = nullptr;
    uint8_t* d_mat_class = nullptr;
    int n_signatures = 0;              // distinct lobe signatures among the materials
    std::vector<float*> d_env_arrays;  // one block of device floats per environment map
    std::vector<EnvMap> envs;          // device views
    EnvMap* d_envs = nullptr;
    std::vector<VolumeRegion> volumes;  // device views (grid data pointers are device addresses)
    std::vector<float*> d_grids;
    VolumeRegion* d_volumes = nullptr;
    std::mutex mu;
    Scratch scratch;
    ~crt_world() {
        cudaFree(d_materials);
        cudaFree(d_geom_material);
        cudaFree(d_lights);
        cudaFree(d_geom_light);
        cudaFree(d_mat_class);
        cudaFree(d_envs);
        for (float* p : d_env_arrays) cudaFree(p);
        for (float* p : d_grids) cudaFree(p);
        cudaFree(d_volumes);
        if (scene) crt_scene_release(scene);
    }
};

extern "C" {

void crt_material_default(crt_material* out) {
    if (!out) return;
    Material m = material_default();
    material_to_abi(m, out);
}

int crt_world_new(crt_scene* scene, const crt_material* materials, size_t n_materials, const uint32_t* geom_material, size_t n_geoms,
                  const crt_light* lights, size_t n_lights, crt_world** out) {
    if (!scene || !out || (n_materials && !materials) || (n_geoms && !geom_material) || (n_lights && !lights))
        return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_new: null argument");
    if (n_geoms != scene->n_geoms) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_new: geom_material must have one entry per geometry of the scene");
    for (size_t g = 0; g < n_geoms; ++g)
        if (geom_material[g] >= n_materials) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_new: material index out of range");
    if (n_materials > 0xFFFEu) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_new: at most 65534 distinct materials");
    for (size_t l = 0; l < n_lights; ++l)
        if (lights[l].kind > CRT_LIGHT_DOME) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_new: unknown light kind");
    int rc = crt::require_device();
    if (rc) return rc;
    auto* w = new crt_world();
    crt_scene_retain(scene);
    w->scene = scene;
    for (size_t i = 0; i < n_materials; ++i) w->materials.push_back(material_from_abi(materials[i]));
    w->geom_material.assign(geom_material, geom_material + n_geoms);
    for (size_t i = 0; i < n_lights; ++i) w->lights.push_back(light_from_abi(lights[i]));
    // Shading classes: materials that take the same branches through the OpenPBR code (same set of active lobes) share one.
    std::vector<uint8_t> mat_class(std::max<size_t>(n_materials, 1), 1);
    {
        std::vector<uint32_t> seen;
        for (size_t i = 0; i < n_materials; ++i) {
            const Material& m = w->materials[i];
            uint32_t sig = m.kind == MAT_EMISSIVE ? 0x80000000u
                                                  : (m.transmission_weight > 0.0f ? 1u : 0u) | (m.coat_weight > 0.0f ? 2u : 0u) | (m.fuzz_weight > 0.0f ? 4u : 0u) |
                                                        (m.geometry_thin_walled ? 8u : 0u) | (m.base_metalness > 0.0f ? 16u : 0u) | (m.subsurface_weight > 0.0f ? 32u : 0u) |
                                                        (m.thin_film_weight > 0.0f ? 64u : 0u) | (m.specular_weight > 0.0f ? 128u : 0u);
            size_t c = std::find(seen.begin(), seen.end(), sig) - seen.begin();
            if (c == seen.size()) seen.push_back(sig);
            mat_class[i] = (uint8_t)std::min<size_t>(1 + c, kShadeBins - 1);
        }
        w->n_signatures = (int)seen.size();
    }
    std::vector<int32_t> geom_light(n_geoms, -1);
    for (size_t g = 0; g < n_geoms; ++g) geom_light[g] = light_find_by_geom(w->lights.data(), (int)w->lights.size(), (uint32_t)g);
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes ? bytes : 16);
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t e;
    if ((e = up((void**)&w->d_materials, w->materials.data(), w->materials.size() * sizeof(Material))) != cudaSuccess ||
        (e = up((void**)&w->d_geom_material, w->geom_material.data(), n_geoms * sizeof(uint32_t))) != cudaSuccess ||
        (e = up((void**)&w->d_lights, w->lights.data(), w->lights.size() * sizeof(Light))) != cudaSuccess ||
        (e = up((void**)&w->d_geom_light, geom_light.data(), n_geoms * sizeof(int32_t))) != cudaSuccess ||
        (e = up((void**)&w->d_mat_class, mat_class.data(), mat_class.size())) != cudaSuccess) {
        delete w;
        return crt::set_cuda_error(e, "crt_world_new upload");
    }
    *out = w;
    return CRT_OK;
}

void crt_world_release(crt_world* w) { delete w; }

int crt_world_set_environment(crt_world* w, uint32_t light_index, uint32_t width, uint32_t height, const float* rgb, const float light_to_world[9]) {
    if (!w || !rgb || !light_to_world) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_set_environment: null argument");
    if (light_index >= w->lights.size() || w->lights[light_index].kind != LIGHT_DOME)
        return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_set_environment: light_index does not name a dome light");
    EnvMapTables t;
    if (width == 0 || height == 0 || width > 65536u || height > 65536u || !t.build((int32_t)width, (int32_t)height, rgb))
        return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_set_environment: malformed map (environment.rs:118-120)");
    std::lock_guard<std::mutex> lk(w->mu);
    // one device block: pixels | marg_func | marg_cdf | cond_func | cond_cdf | cond_integral
    const size_t np = t.pixels.size(), nmf = t.marg_func.size(), nmc = t.marg_cdf.size(), ncf = t.cond_func.size(), ncc = t.cond_cdf.size(), nci = t.cond_integral.size();
    float* d = nullptr;
    CKR(cudaMalloc((void**)&d, (np + nmf + nmc + ncf + ncc + nci) * sizeof(float)));
    w->d_env_arrays.push_back(d);
    size_t o = 0;
    auto put = [&](const std::vector<float>& v) -> const float* {
        cudaMemcpy(d + o, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
        const float* at = d + o;
        o += v.size();
        return at;
    };
    EnvMap m;
    m.width = t.width, m.height = t.height;
    m.pixels = put(t.pixels);
    m.marg_func = put(t.marg_func);
    m.marg_cdf = put(t.marg_cdf);
    m.cond_func = put(t.cond_func);
    m.cond_cdf = put(t.cond_cdf);
    m.cond_integral = put(t.cond_integral);
    m.marg_integral = t.marg_integral;
    m.pad = 0;
    w->envs.push_back(m);
    cudaFree(w->d_envs);
    w->d_envs = nullptr;
    CKR(cudaMalloc((void**)&w->d_envs, w->envs.size() * sizeof(EnvMap)));
    CKR(cudaMemcpy(w->d_envs, w->envs.data(), w->envs.size() * sizeof(EnvMap), cudaMemcpyHostToDevice));
    light_set_environment(w->lights[light_index], (int32_t)w->envs.size() - 1, light_to_world);
    CKR(cudaMemcpy(w->d_lights, w->lights.data(), w->lights.size() * sizeof(Light), cudaMemcpyHostToDevice));
    return CRT_OK;
}

int crt_world_set_volumes(crt_world* w, const crt_volume* volumes, size_t n) {
    if (!w || (n && !volumes)) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_set_volumes: null argument");
    for (size_t i = 0; i < n; ++i)
        if (!volume_abi_valid(volumes[i])) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_world_set_volumes: unknown field kind or malformed grid");
    int rc = crt::require_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(w->mu);
    for (float* p : w->d_grids) cudaFree(p);
    w->d_grids.clear();
    w->volumes.clear();
    cudaFree(w->d_volumes);
    w->d_volumes = nullptr;
    for (size_t i = 0; i < n; ++i) {
        const crt_volume& v = volumes[i];
        float* d = nullptr;
        if (v.field == CRT_FIELD_GRID) {
            size_t cells = (size_t)v.grid_dims[0] * v.grid_dims[1] * v.grid_dims[2];
            CKR(cudaMalloc((void**)&d, cells * sizeof(float)));
            w->d_grids.push_back(d);
            CKR(cudaMemcpy(d, v.grid_data, cells * sizeof(float), cudaMemcpyHostToDevice));
        }
        w->volumes.push_back(volume_from_abi(v, v.grid_data, d));
    }
    if (n) {
        CKR(cudaMalloc((void**)&w->d_volumes, n * sizeof(VolumeRegion)));
        CKR(cudaMemcpy(w->d_volumes, w->volumes.data(), n * sizeof(VolumeRegion), cudaMemcpyHostToDevice));
    }
    return CRT_OK;
}

int crt_steal_cursor_create(crt_steal_cursor** out, unsigned char ipc_handle[64]) {
    if (!out) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_steal_cursor_create: null argument");
    int rc = crt::require_device();
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    auto* c = new crt_steal_cursor();
    c->owner = true;
    cudaError_t e = cudaMalloc(&c->dev, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->dev, 0, sizeof(unsigned int));
    if (e == cudaSuccess && ipc_handle) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, c->dev);
        if (e == cudaSuccess) memcpy(ipc_handle, &h, 64);
    }
    if (e != cudaSuccess) {
        cudaFree(c->dev);
        delete c;
        return crt::set_cuda_error(e, "crt_steal_cursor_create");
    }
    *out = c;
    return CRT_OK;
}

int crt_steal_cursor_open(const unsigned char ipc_handle[64], crt_steal_cursor** out) {
    if (!out || !ipc_handle) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_steal_cursor_open: null argument");
    int rc = crt::require_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    auto* c = new crt_steal_cursor();
    cudaError_t e = cudaIpcOpenMemHandle((void**)&c->dev, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        delete c;
        return crt::set_cuda_error(e, "cudaIpcOpenMemHandle (peer access between the two GPUs is required)");
    }
    *out = c;
    return CRT_OK;
}

int crt_steal_cursor_reset(crt_steal_cursor* c) {
    if (!c || !c->owner) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_steal_cursor_reset: only the creating process resets the cursor");
    CKR(cudaMemset(c->dev, 0, sizeof(unsigned int)));
    return CRT_OK;
}

void crt_steal_cursor_close(crt_steal_cursor* c) {
    if (!c) return;
    if (c->owner)
        cudaFree(c->dev);
    else
        cudaIpcCloseMemHandle(c->dev);
    delete c;
}

int crt_render_steal(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings, float* framebuffer, crt_ray_stats* stats, void* stream,
                     crt_steal_cursor* cursor);

int crt_shared_frame_create(size_t bytes, void** device_ptr, unsigned char ipc_handle[64]) {
    if (!device_ptr || !ipc_handle || bytes == 0) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_shared_frame_create: null argument");
    int rc = crt::require_device();
    if (rc) return rc;
    void* p = nullptr;
    CKR(cudaMalloc(&p, bytes));  // a plain cudaMalloc block: the IPC handle then names exactly this frame (offset 0)
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return crt::set_cuda_error(e, "crt_shared_frame_create");
    }
    memcpy(ipc_handle, &h, 64);
    *device_ptr = p;
    return CRT_OK;
}

int crt_shared_frame_open(const unsigned char ipc_handle[64], void** device_ptr) {
    if (!device_ptr || !ipc_handle) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_shared_frame_open: null argument");
    int rc = crt::require_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return crt::set_cuda_error(e, "cudaIpcOpenMemHandle (peer access between the two GPUs is required)");
    *device_ptr = p;
    return CRT_OK;
}

int crt_shared_frame_close(void* device_ptr, int owner) {
    if (!device_ptr) return CRT_OK;
    cudaError_t e = owner ? cudaFree(device_ptr) : cudaIpcCloseMemHandle(device_ptr);
    if (e != cudaSuccess) return crt::set_cuda_error(e, "crt_shared_frame_close");
    return CRT_OK;
}

int crt_render_last_info(uint64_t* launches, double* device_ms, uint64_t* rounds) {
    if (launches) *launches = t_last_launches;
    if (device_ms) *device_ms = t_last_ms;
    if (rounds) *rounds = t_last_rounds;
    return CRT_OK;
}

}  // extern "C"

namespace {

template <typename T>
T* carve(char*& p, size_t count) {
    T* r = (T*)p;
    p += (count * sizeof(T) + 255) & ~(size_t)255;
    return r;
}
size_t carve_bytes(size_t count, size_t elem) { return (count * elem + 255) & ~(size_t)255; }

// Runs `fn(first, last)` over [0, n) on up to `threads` host threads (the calling thread takes the first slice).
template <class F>
void host_parallel_for(size_t n, unsigned threads, F&& fn) {
    threads = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n));
    std::vector<std::thread> pool;
    const size_t per = (n + threads - 1) / threads;
    for (unsigned t = 1; t < threads; ++t) {
        size_t a = std::min(n, t * per), b = std::min(n, a + per);
        if (a < b) pool.emplace_back([=, &fn] { fn(a, b); });
    }
    fn(0, std::min(n, per));
    for (std::thread& t : pool) t.join();
}

bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// `host_fb` != nullptr (crt_render_host): the frame ends in host memory. The device frame is the world's own; with a tile
// split only the pixels this process owns are resolved (packed, in pixel-list order), copied to pinned staging with one
// asynchronous copy and scattered into the caller's frame on the host — the other ranks' pixels are neither uploaded nor
// downloaded. A whole-frame render is copied straight into pinned caller memory, or through pinned staging and a
// multi-threaded memcpy into pageable memory.
int render_impl(crt_world* w, const crt_camera* camera, const crt_render_settings* s, float* d_fb, crt_ray_stats* stats_out, cudaStream_t stream,
                crt_steal_cursor* cursor, float* host_fb = nullptr) {
    if (!w || !camera || !s || (!d_fb && !host_fb)) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render: null argument");
    if (s->width == 0 || s->height == 0 || (uint64_t)s->width * s->height > 0x7FFFFFFFu) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render: bad image size");
    if (s->samples_per_pixel == 0 || s->samples_per_pixel > 65536u) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render: samples_per_pixel must be in 1..65536 (sampler index range)");
    if (s->max_depth > 4000u) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render: max_depth must be <= 4000 (12-bit per-sample ray counters)");
    if (s->pixel_filter > CRT_FILTER_MITCHELL || s->sampling_strategy > CRT_STRATEGY_BSDF_ONLY) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render: unknown filter or strategy");
    if (s->tile_world > 1 && (s->tile_size == 0 || s->tile_rank >= s->tile_world)) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render: bad tile split");
    int rc = crt::require_device();
    if (rc) return rc;
    if (w->scene->host_only) return crt::set_error(CRT_ERR_NO_DEVICE, "scene was committed host-only: nothing to render with");
    std::lock_guard<std::mutex> lk(w->mu);
    Scratch& sc = w->scratch;
    // L2 window over the tree's top levels for this call's traversal launches only (cleared on every exit path)
    struct Hint {
        bool on;
        cudaStream_t st;
        ~Hint() {
            if (on) crt::clear_scene_hint(st);
        }
    } hint{crt::hint_scene_top_levels(w->scene, stream), stream};
    const int sms = crt::device_sm_count();
    const int D = (int)s->max_depth;
    uint64_t launches = 0, trace_launches = 0, rounds = 0, chunks_done = 0;

    // ---- which pixels this process renders ----
    // Tiles are walked in 8x4 pixel blocks so a warp of camera rays covers a compact footprint.
    const uint32_t W = s->width, H = s->height;
    const uint32_t T = s->tile_size ? s->tile_size : 32u, tiles_x = (W + T - 1) / T, tiles_y = (H + T - 1) / T, n_tiles = tiles_x * tiles_y;
    auto append_tile = [&](uint32_t t, std::vector<uint32_t>& out) {
        uint32_t x0 = (t % tiles_x) * T, y0 = (t / tiles_x) * T;
        for (uint32_t by = y0; by < std::min(y0 + T, H); by += 4)
            for (uint32_t bx = x0; bx < std::min(x0 + T, W); bx += 8)
                for (uint32_t y = by; y < std::min({by + 4, y0 + T, H}); ++y)
                    for (uint32_t x = bx; x < std::min({bx + 8, x0 + T, W}); ++x) out.push_back(y * W + x);
    };
    std::vector<uint32_t> pixels;
    Scratch::PixelDeal deal;
    deal.w = W, deal.h = H;
    if (s->tile_world > 1) deal.tile = T, deal.rank = s->tile_rank, deal.world = s->tile_world;  // a whole-frame render ignores the tile fields
    const bool pixels_cached = !cursor && sc.pixel_deal_valid && sc.pixel_deal == deal && sc.pixels != nullptr;
    // 0 = auto: about four chunks per participating process — coarse enough that each chunk's adaptive rounds still
    // fill the GPU (small chunks are launch-latency-bound), fine enough to absorb a slow or busy GPU
    const uint32_t auto_chunk = std::max(1u, n_tiles / (4u * std::max(1u, s->tile_world)));
    const uint32_t steal_chunk = cursor ? (s->steal_chunk_tiles ? s->steal_chunk_tiles : auto_chunk) : 0u;
    const uint32_t n_chunks = steal_chunk ? (n_tiles + steal_chunk - 1) / steal_chunk : 0u;
    size_t max_pixels;  // largest pixel set one render_pixel_set call will see (buffer sizing)
    if (cursor) {
        max_pixels = std::min<size_t>((size_t)steal_chunk * T * T, (size_t)W * H);
    } else if (pixels_cached) {
        max_pixels = sc.pixel_count;
    } else if (s->tile_world <= 1) {
        pixels.reserve((size_t)W * H);
        for (uint32_t by = 0; by < H; by += 4)
            for (uint32_t bx = 0; bx < W; bx += 8)
                for (uint32_t y = by; y < std::min(by + 4, H); ++y)
                    for (uint32_t x = bx; x < std::min(bx + 8, W); ++x) pixels.push_back(y * W + x);
        max_pixels = pixels.size();
    } else {
        // static deal: tile t belongs to rank t % world (interleaved, so adaptive cost is decorrelated from rank)
        for (uint32_t t = s->tile_rank; t < n_tiles; t += s->tile_world) append_tile(t, pixels);
        max_pixels = pixels.size();
    }
    const bool packed = host_fb != nullptr && s->tile_world > 1;
    if (host_fb) {
        const size_t need_fb = (packed ? max_pixels : (size_t)W * H) * 16;
        if (need_fb > sc.cap_host_fb) {
            cudaFree(sc.host_fb);
            sc.host_fb = nullptr;
            sc.cap_host_fb = 0;
            CKR(cudaMalloc(&sc.host_fb, need_fb ? need_fb : 16));
            sc.cap_host_fb = need_fb;
        }
        d_fb = (float*)sc.host_fb;
        if (packed && !(sc.h_deal_valid && sc.h_deal == deal)) {
            // host copy of the pixel list = the scatter order; kept per deal (a device-frame render of another deal in
            // between leaves it alone, so it carries its own key)
            if (!pixels_cached) {
                sc.h_pixels = pixels;
            } else {
                sc.h_pixels.clear();
                for (uint32_t t = s->tile_rank; t < n_tiles; t += s->tile_world) append_tile(t, sc.h_pixels);
            }
            sc.h_deal = deal;
            sc.h_deal_valid = true;
        }
    }
    if (!sc.ev0) {
        CKR(cudaEventCreate(&sc.ev0));
        CKR(cudaEventCreate(&sc.ev1));
        CKR(cudaMalloc(&sc.stats, 8 * sizeof(unsigned long long)));
        CKR(cudaMalloc(&sc.active_count, 2 * sizeof(uint32_t)));
        CKR(cudaMallocHost(&sc.h_pinned, 64 * sizeof(uint32_t)));
    }
    CKR(cudaEventRecord(sc.ev0, stream));
    CKR(cudaMemsetAsync(sc.stats, 0, 8 * sizeof(unsigned long long), stream));
    if (max_pixels == 0) {
        CKR(cudaEventRecord(sc.ev1, stream));
        CKR(cudaStreamSynchronize(stream));
        if (stats_out) *stats_out = crt_ray_stats{};
        t_last_launches = 0;
        t_last_rounds = 0;
        t_last_ms = 0.0;
        return CRT_OK;
    }

    // ---- buffers ----
    size_t cap = s->max_paths ? s->max_paths : ((size_t)1 << 24);  // 16 Mi paths (~5.8 GB of queues + state): fewer, longer launches; measured 4..64 Mi on B200
    cap = std::max<size_t>(cap, 1024);
    cap = std::min<size_t>(cap, max_pixels * s->samples_per_pixel);
    cap = std::max<size_t>(cap, 1);
    const bool motion = w->scene->has_motion;
    {
        size_t need = 3 * carve_bytes(cap, 16) + 2 * carve_bytes(cap, 8) + carve_bytes(cap, 4)                       // path state
                      + 2 * (carve_bytes(2 * cap, 16) + 2 * carve_bytes(cap, 4)) + carve_bytes(2 * cap, 16) + carve_bytes(cap, 1)  // extend
                      + carve_bytes(2 * cap, 16) + 2 * carve_bytes(cap, 4) + carve_bytes(cap, 16) + carve_bytes(cap, 1)           // shadow
                      + carve_bytes(cap, 4) + carve_bytes((size_t)kShadeBins * kSortBlocks, 4);                                     // shading order
        if (need > sc.cap_paths) {
            cudaFree(sc.paths);
            sc.paths = nullptr;
            sc.cap_paths = 0;
            CKR(cudaMalloc(&sc.paths, need));
            sc.cap_paths = need;
        }
        size_t need_px = carve_bytes(max_pixels, 4) * 4 + carve_bytes(max_pixels, 16) * 2;
        if (need_px > sc.cap_pixels) {
            cudaFree(sc.pixels);
            sc.pixels = nullptr;
            sc.cap_pixels = 0;
            sc.pixel_deal_valid = false;
            CKR(cudaMalloc(&sc.pixels, need_px));
            sc.cap_pixels = need_px;
        }
        size_t need_counts = 4 * (size_t)(D + 2);
        if (need_counts > sc.cap_counts) {
            cudaFree(sc.counts);
            sc.counts = nullptr;
            sc.cap_counts = 0;
            CKR(cudaMalloc(&sc.counts, need_counts * sizeof(uint32_t)));
            sc.cap_counts = need_counts;
        }
    }
    RenderParams P;
    P.cam = camera_from_abi(*camera);
    P.filter = filter_new(s->pixel_filter, s->filter_radius);
    P.st = settings_from_abi(*s);
    P.materials = w->d_materials;
    P.geom_material = w->d_geom_material;
    P.mat_class = w->d_mat_class;
    P.lights = w->d_lights;
    P.geom_light = w->d_geom_light;
    P.envs = w->d_envs;
    P.volumes = w->d_volumes;
    P.n_volumes = (int)w->volumes.size();
    P.n_lights = (int)w->lights.size();
    P.motion = motion ? 1 : 0;
    P.stats = sc.stats;
    P.status = crt::device_status_word();
    {
        char* p = (char*)sc.paths;
        P.ps.beta_pdf = carve<float4>(p, cap);
        P.ps.L_w = carve<float4>(p, cap);
        P.ps.prevp = carve<float4>(p, cap);
        P.ps.sampler = carve<uint2>(p, cap);
        P.ps.time = carve<float>(p, cap);
        P.ps.counts = carve<uint2>(p, cap);
        for (int b = 0; b < 2; ++b) {
            P.q.rays[b] = carve<float4>(p, 2 * cap);
            P.q.path[b] = carve<uint32_t>(p, cap);
            float* t = carve<float>(p, cap);
            P.q.time[b] = motion ? t : nullptr;
        }
        P.q.hits = carve<float4>(p, 2 * cap);
        P.q.ff = carve<uint8_t>(p, cap);
        P.q.srays = carve<float4>(p, 2 * cap);
        float* st = carve<float>(p, cap);
        P.q.stime = motion ? st : nullptr;
        P.q.spath = carve<uint32_t>(p, cap);
        P.q.scontrib = carve<float4>(p, cap);
        P.q.occ = carve<uint8_t>(p, cap);
        P.q.order = carve<uint32_t>(p, cap);
        P.q.bins = carve<uint32_t>(p, (size_t)kShadeBins * kSortBlocks);
        P.q.counts = sc.counts;
    }
    auto shade_grid = [&](size_t n) {
        size_t blocks = (n + CRT_SHADE_BLOCK - 1) / CRT_SHADE_BLOCK;
        return (int)std::max<size_t>(1, std::min<size_t>(blocks, (size_t)sms * (8 * 128 / CRT_SHADE_BLOCK)));
    };
    auto grid_for = [&](size_t n) {
        size_t blocks = (n + kBlock - 1) / kBlock;
        return (int)std::max<size_t>(1, std::min<size_t>(blocks, (size_t)sms * 8));
    };

    const bool split_shade = !getenv("CRT_SHADE_FUSED");  // development switch: the fused kernel for A/B timing
    // Worth it only where warps would otherwise mix several code paths (measured: openpbr_showcase, 7 signatures, -30 %;
    // veach_mis, 3 signatures, +1.5 %; the two-signature stress scene +8 %: the sorted order costs the next queue's locality)
    const bool sort_shade = !getenv("CRT_SHADE_UNSORTED") && (w->n_signatures >= 3 || getenv("CRT_SHADE_SORTED"));
    // without light sampling the vertex pass is nearly empty and the second launch costs more than it saves (measured on the
    // sky-only cornellbox / MedCity scenes)
    const bool PL_has_nee = P.n_lights > 0 && strategy_samples_lights(P.st.sampling_strategy);
    // ---- one wavefront: `nb` pixels (active[off .. off+nb)) x samples [s_base, s_base + s_count) ----
    auto wavefront = [&](const uint32_t* active, uint32_t nb, uint32_t s_base, uint32_t s_count) -> int {
        const uint32_t total = nb * s_count;
        CKR(cudaMemsetAsync(sc.counts, 0, 4 * (size_t)(D + 2) * sizeof(uint32_t), stream));
        raygen_kernel<<<grid_for(total), kBlock, 0, stream>>>(P, active, nb, s_base, s_count);
        ++launches;
        if (D == 0) {
            // max_depth 0: every path ends at `remaining <= 0` before tracing anything (tracer.rs:1123-1148)
            return CRT_OK;
        }
        // Upper bound on the length of extend queue k (and of shadow queue k): a path yields at most one continuation and
        // one shadow ray per vertex, so queue lengths only shrink with depth; refreshed whenever the host looks at a count.
        // Sizes the grids of a thinning wavefront's launches (a full persistent grid costs ~10 us to start and drain).
        uint32_t bound = total;
        for (int k = 0; k <= D; ++k) {
            const int cur = k & 1;
            crt::QueueTrace qt;
            qt.rays = P.q.rays[cur];
            qt.time = P.q.time[cur];
            qt.mask_uniform = k == 0 ? CRT_MASK_CAMERA : CRT_MASK_INDIRECT;
            qt.hits = P.q.hits;
            qt.flag = P.q.ff;
            qt.n_dev = sc.counts + k;
            qt.capacity = bound;
            qt.counter = sc.counts + 2 * (D + 2) + k;
            int r = crt::enqueue_queue_trace(w->scene, false, qt, stream);
            if (r) return r;
            // group the queue by class when it is long enough to pay for the three extra launches
            const uint32_t* order = nullptr;
            if (sort_shade && bound >= (1u << 17) && k < D) {
                const int sort_blocks = (int)std::min<size_t>(kSortBlocks, ((size_t)bound + 16 * kBlock - 1) / (16 * kBlock));
                const int nc = std::min(kShadeBins, w->n_signatures + 1);  // miss + one class per signature
                classify_count_kernel<<<sort_blocks, kBlock, 0, stream>>>(P, k, nc);
                classify_scan_kernel<<<1, 1024, 0, stream>>>(P.q.bins, (uint32_t)(nc * sort_blocks));
                classify_scatter_kernel<<<sort_blocks, kBlock, 0, stream>>>(P, k, nc);
                launches += 3;
                order = P.q.order;
            }
            if (P.n_volumes > 0) {
                shade_kernel<true, SHADE_FUSED><<<shade_grid(bound), CRT_SHADE_BLOCK, 0, stream>>>(P, k, cur, order);
            } else if (split_shade && PL_has_nee) {
                shade_kernel<false, SHADE_VERTEX><<<shade_grid(bound), CRT_SHADE_BLOCK, 0, stream>>>(P, k, cur, order);
                if (k < D) {  // a depth-exhausted vertex has no continuation to sample
                    shade_kernel<false, SHADE_SCATTER><<<shade_grid(bound), CRT_SHADE_BLOCK, 0, stream>>>(P, k, cur, order);
                    ++launches;
                }
            } else {
                shade_kernel<false, SHADE_FUSED><<<shade_grid(bound), CRT_SHADE_BLOCK, 0, stream>>>(P, k, cur, order);
            }
            launches += 2;
            ++trace_launches;
            if (k < D && P.n_lights > 0 && strategy_samples_lights(P.st.sampling_strategy)) {
                crt::QueueTrace qs;
                qs.rays = P.q.srays;
                qs.time = P.q.stime;
                qs.mask_uniform = CRT_MASK_SHADOW;
                qs.hits = nullptr;
                qs.flag = P.q.occ;
                qs.n_dev = sc.counts + (D + 2) + k;
                qs.capacity = bound;
                qs.counter = sc.counts + 3 * (D + 2) + k;
                r = crt::enqueue_queue_trace(w->scene, true, qs, stream);
                if (r) return r;
                if (P.n_volumes > 0)
                    nee_resolve_kernel<true><<<grid_for(bound), kBlock, 0, stream>>>(P, k);
                else
                    nee_resolve_kernel<false><<<grid_for(bound), kBlock, 0, stream>>>(P, k);
                launches += 2;
                ++trace_launches;
            }
            // Every few bounces look at the next queue's length so a dead wavefront stops launching.
            if (k >= 5 && k < D && ((k - 5) % 4) == 0) {
                CKR(cudaMemcpyAsync(sc.h_pinned, sc.counts + k + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                CKR(cudaStreamSynchronize(stream));
                if (sc.h_pinned[0] == 0) break;
                bound = std::min(bound, sc.h_pinned[0]);
            }
        }
        return CRT_OK;
    };

    // ---- one pixel set: all rounds of all its pixels, then its film pixels ----
    uint32_t* d_active[2] = {nullptr, nullptr};
    auto render_pixel_set = [&](const std::vector<uint32_t>& pixels, uint32_t n_pix, bool upload) -> int {
        if (n_pix == 0) return CRT_OK;
        {
            char* p = (char*)sc.pixels;
            uint32_t* d_pixels = carve<uint32_t>(p, n_pix);
            d_active[0] = carve<uint32_t>(p, n_pix);
            d_active[1] = carve<uint32_t>(p, n_pix);
            P.px.taken = carve<uint32_t>(p, n_pix);
            P.px.sum_w = carve<float4>(p, n_pix);
            P.px.lum = (double2*)carve<float4>(p, n_pix);
            P.px.pixels = d_pixels;
            if (upload) {
                CKR(cudaMemcpyAsync(d_pixels, pixels.data(), (size_t)n_pix * 4, cudaMemcpyHostToDevice, stream));
                CKR(cudaStreamSynchronize(stream));  // `pixels` may be reused by the caller
            }
            iota_kernel<<<(int)std::max<size_t>(1, std::min<size_t>(((size_t)n_pix + kBlock - 1) / kBlock, (size_t)sms * 8)), kBlock, 0, stream>>>(d_active[0], n_pix);
            ++launches;
            CKR(cudaMemsetAsync(P.px.taken, 0, (size_t)n_pix * 4, stream));
            CKR(cudaMemsetAsync(P.px.sum_w, 0, (size_t)n_pix * 16, stream));
            CKR(cudaMemsetAsync(P.px.lum, 0, (size_t)n_pix * 16, stream));
        }
        // ---- rounds (tracer.rs:558-618): samples are taken in order; the adaptive test runs after every
        //      4th sample once min_spp is reached, so rounds end on those boundaries ----
        const uint32_t spp = s->samples_per_pixel;
        const bool adaptive = s->adaptive && s->variance_threshold > 0.0f;
        const uint32_t min_spp = std::max(s->min_samples_per_pixel, 2u);
        uint32_t n_active = n_pix, s_base = 0;
        int cur_list = 0;
        double stop_rate = 1.0;  // fraction of the active pixels that stopped per 4-sample group in the last round
        while (n_active > 0 && s_base < spp) {
            uint32_t round_samples;
            if (!adaptive)
                round_samples = spp - s_base;
            else if (s_base == 0)
                round_samples = std::min(spp, (min_spp + 3u) / 4u * 4u);
            else {
                // later rounds: a multiple of 4 large enough to keep ~1 Mi paths in flight (samples a pixel turns out
                // not to need are discarded by accumulate_kernel, so this only trades a little work for fewer rounds)
                uint32_t groups = (uint32_t)std::min<size_t>(8, std::max<size_t>(1, ((size_t)1 << 20) / ((size_t)n_active * 4)));
                // ... and longer rounds while few pixels are stopping: a round costs its launches whatever its length, a
                // pixel that stops inside a G-group round wastes at most G - 1 groups of samples
                const uint32_t by_rate = stop_rate > 0.10 ? 1u : (stop_rate > 0.03 ? 2u : (stop_rate > 0.01 ? 4u : 8u));
                groups = std::max(groups, by_rate);
                round_samples = std::min(spp - s_base, 4u * groups);
            }
            // split the round into wavefronts that fit the pool
            uint32_t sub = (uint32_t)std::max<size_t>(1, std::min<size_t>(round_samples, cap / n_active));
            uint32_t block_px = (uint32_t)std::min<size_t>(n_active, cap / sub);
            for (uint32_t s0 = 0; s0 < round_samples; s0 += sub) {
                uint32_t sc_count = std::min(sub, round_samples - s0);
                for (uint32_t off = 0; off < n_active; off += block_px) {
                    uint32_t nb = std::min(block_px, n_active - off);
                    int r = wavefront(d_active[cur_list] + off, nb, s_base + s0, sc_count);
                    if (r) return r;
                    accumulate_kernel<<<grid_for(nb), kBlock, 0, stream>>>(P, d_active[cur_list] + off, nb, sc_count);
                    ++launches;
                }
            }
            s_base += round_samples;
            ++rounds;
            if (s_base >= spp) break;
            CKR(cudaMemsetAsync(sc.active_count, 0, sizeof(uint32_t), stream));
            adaptive_kernel<<<grid_for(n_active), kBlock, 0, stream>>>(P, d_active[cur_list], n_active, d_active[cur_list ^ 1], sc.active_count);
            ++launches;
            CKR(cudaMemcpyAsync(sc.h_pinned, sc.active_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CKR(cudaStreamSynchronize(stream));
            if (s_base > round_samples)  // not after the first round: its stop count covers min_spp samples, not one group
                stop_rate = (double)(n_active - sc.h_pinned[0]) / (double)n_active / (double)(round_samples / 4u);
            else
                stop_rate = (double)(n_active - sc.h_pinned[0]) / (double)n_active;
            n_active = sc.h_pinned[0];
            cur_list ^= 1;
        }
        film_kernel<<<grid_for(n_pix), kBlock, 0, stream>>>(P, n_pix, (float4*)d_fb, packed ? 1 : 0);
        ++launches;
        return CRT_OK;
    };
    if (cursor) {
        // work stealing: every process pulls chunks of `steal_chunk` tiles off one system-scope cursor until the
        // frame is exhausted; chunk c = tiles {c + k * n_chunks}, i.e. spread over the whole image
        for (;;) {
            fetch_chunk_kernel<<<1, 1, 0, stream>>>(cursor->dev, sc.active_count + 1);
            ++launches;
            CKR(cudaMemcpyAsync(sc.h_pinned + 1, sc.active_count + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CKR(cudaStreamSynchronize(stream));
            const uint32_t c = sc.h_pinned[1];
            if (c >= n_chunks) break;
            pixels.clear();
            for (uint32_t t = c; t < n_tiles; t += n_chunks) append_tile(t, pixels);
            sc.pixel_deal_valid = false;  // the pixel scratch now holds a chunk's list
            int r = render_pixel_set(pixels, (uint32_t)pixels.size(), true);
            if (r) return r;
            ++chunks_done;
        }
    } else {
        if (!pixels_cached) sc.pixel_deal_valid = false;  // the list is being replaced: not valid again until the set has rendered
        int r = render_pixel_set(pixels, pixels_cached ? sc.pixel_count : (uint32_t)pixels.size(), !pixels_cached);
        if (r) return r;
        sc.pixel_deal = deal;
        sc.pixel_deal_valid = true;
        sc.pixel_count = pixels_cached ? sc.pixel_count : (uint32_t)pixels.size();
    }
    CKR(cudaGetLastError());
    CKR(cudaMemcpyAsync(sc.h_pinned + 16, sc.stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CKR(cudaEventRecord(sc.ev1, stream));
    CKR(cudaStreamSynchronize(stream));
    float ms = 0.0f;
    CKR(cudaEventElapsedTime(&ms, sc.ev0, sc.ev1));
    crt::count_launches(launches - trace_launches);  // traversal launches were already counted by enqueue_queue_trace
    if (stats_out) {
        const unsigned long long* st = (const unsigned long long*)(sc.h_pinned + 16);
        stats_out->camera_rays = st[0];
        stats_out->closest_hit = st[1];
        stats_out->shadow_rays = st[2];
        stats_out->vertices = st[3];
        stats_out->rr_tested = st[4];
        stats_out->rr_killed = st[5];
        stats_out->ended_escaped = st[6];
        stats_out->ended_depth = st[7];
    }
    t_last_launches = launches;
    t_last_rounds = rounds;
    t_last_ms = (double)ms;
    t_last_chunks = chunks_done;
    if (host_fb) {
        const size_t n_out = packed ? (size_t)(pixels_cached ? sc.pixel_count : sc.h_pixels.size()) : (size_t)W * H;
        const size_t bytes = n_out * 16;
        if (!packed && host_pointer_is_pinned(host_fb)) {
            CKR(cudaMemcpyAsync(host_fb, d_fb, bytes, cudaMemcpyDeviceToHost, stream));
            CKR(cudaStreamSynchronize(stream));
            return CRT_OK;
        }
        if (bytes > sc.cap_stage) {
            if (sc.h_stage) cudaFreeHost(sc.h_stage);
            sc.h_stage = nullptr;
            sc.cap_stage = 0;
            CKR(cudaMallocHost((void**)&sc.h_stage, bytes));
            sc.cap_stage = bytes;
        }
        CKR(cudaMemcpyAsync(sc.h_stage, d_fb, bytes, cudaMemcpyDeviceToHost, stream));
        CKR(cudaStreamSynchronize(stream));
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned threads = std::min(8u, std::max(1u, hw / std::max(1u, s->tile_world)));
        if (!packed) {
            char* dst = (char*)host_fb;
            const char* src = sc.h_stage;
            host_parallel_for((bytes + ((size_t)1 << 20) - 1) >> 20, threads, [=](size_t a, size_t b) {
                const size_t o0 = a << 20, o1 = std::min(bytes, b << 20);
                memcpy(dst + o0, src + o0, o1 - o0);
            });
        } else {
            // runs of consecutive pixel indices (8 pixels = 128 bytes in the 8x4-block order) are copied whole
            const uint32_t* px = sc.h_pixels.data();
            const float4* src = (const float4*)sc.h_stage;
            float4* dst = (float4*)host_fb;
            host_parallel_for((n_out + 4095) / 4096, threads, [=](size_t a, size_t b) {
                const size_t i1 = std::min(n_out, b * 4096);
                for (size_t i = a * 4096; i < i1;) {
                    size_t j = i + 1;
                    while (j < i1 && px[j] == px[j - 1] + 1) ++j;
                    memcpy(dst + px[i], src + i, (j - i) * sizeof(float4));
                    i = j;
                }
            });
        }
    }
    return CRT_OK;
}

}  // namespace

extern "C" {

int crt_render(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings, float* framebuffer, crt_ray_stats* stats, void* stream) {
    return render_impl(const_cast<crt_world*>(w), camera, settings, framebuffer, stats, (cudaStream_t)stream, nullptr);
}

int crt_render_steal(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings, float* framebuffer, crt_ray_stats* stats, void* stream,
                     crt_steal_cursor* cursor) {
    if (!cursor) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render_steal: null cursor");
    return render_impl(const_cast<crt_world*>(w), camera, settings, framebuffer, stats, (cudaStream_t)stream, cursor);
}

int crt_render_host(const crt_world* w, const crt_camera* camera, const crt_render_settings* settings, float* framebuffer, crt_ray_stats* stats) {
    if (!settings || !framebuffer || !w) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_render_host: null argument");
    int rc = crt::require_device();
    if (rc) return rc;
    // Pixels of other ranks stay as given: with a tile split only this process' pixels are written.
    return render_impl(const_cast<crt_world*>(w), camera, settings, nullptr, stats, nullptr, nullptr, framebuffer);
}

}  // extern "C"


namespace {
int debug_material(const crt_material* material, size_t n, const float* ray_dir, const float* normal, const uint8_t* front_face, const float* arg, size_t arg_w,
                   float* out, size_t out_w, uint8_t* some, int scatter) {
    if (!material || (n && (!ray_dir || !normal || !front_face || !arg || !out || !some))) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_debug_material: null argument");
    if (n > 0x7FFFFFFFu) return crt::set_error(CRT_ERR_INVALID_ARG, "crt_debug_material: too many queries");
    int rc = crt::require_device();
    if (rc) return rc;
    if (n == 0) return CRT_OK;
    const Material m = material_from_abi(*material);
    float *d_rd = nullptr, *d_n = nullptr, *d_arg = nullptr, *d_out = nullptr;
    uint8_t *d_ff = nullptr, *d_some = nullptr;
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dst, const void* src, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(dst, bytes);
        if (e == cudaSuccess && src) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    up((void**)&d_rd, ray_dir, 12 * n);
    up((void**)&d_n, normal, 12 * n);
    up((void**)&d_ff, front_face, n);
    up((void**)&d_arg, arg, 4 * arg_w * n);
    up((void**)&d_out, nullptr, 4 * out_w * n);
    up((void**)&d_some, nullptr, n);
    if (e == cudaSuccess) {
        debug_material_kernel<<<(unsigned)((n + 127) / 128), 128>>>(m, (uint32_t)n, d_rd, d_n, d_ff, d_arg, d_out, d_some, scatter);
        crt::count_launches(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, 4 * out_w * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(some, d_some, n, cudaMemcpyDeviceToHost);
    cudaFree(d_rd), cudaFree(d_n), cudaFree(d_ff), cudaFree(d_arg), cudaFree(d_out), cudaFree(d_some);
    if (e != cudaSuccess) return crt::set_cuda_error(e, "crt_debug_material");
    return CRT_OK;
}
}  // namespace

extern "C" {
int crt_debug_material_eval(const crt_material* material, size_t n, const float* ray_dir, const float* normal, const uint8_t* front_face, const float* wi,
                            float* value_pdf, uint8_t* some) {
    return debug_material(material, n, ray_dir, normal, front_face, wi, 3, value_pdf, 4, some, 0);
}
int crt_debug_material_scatter(const crt_material* material, size_t n, const float* ray_dir, const float* normal, const uint8_t* front_face, const float* s4,
                               float* out8, uint8_t* some) {
    return debug_material(material, n, ray_dir, normal, front_face, s4, 4, out8, 8, some, 1);
}
}  // extern "C"
