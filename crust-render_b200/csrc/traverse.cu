// Persistent-warp BVH4 traversal kernels (closest-hit / any-hit) for sm_100a.
// See traverse.cuh for the semantics contract and reference citations.
#include "traverse.cuh"

namespace crt {

namespace {

// Stack entry: .x = child word (bit 31 = leaf record offset, else node index,
// 0xFFFFFFFF = "return to the instancing level"), .y = entry distance bits.
#define CRT_RETURN_MARK 0xFFFFFFFFu
#define CRT_NO_REC 0xFFFFFFFFu

template <bool ANY_HIT>
struct StackT {
    typedef uint2 type;
};
template <>
struct StackT<true> {
    typedef uint32_t type;
};

// What an instance descent must restore (prim.rs:345-365 returns into the caller's leaf loop), generic nesting depth:
// kept in local memory. Scenes whose instances nest one level deep (every bundled sample but nested_instancing) use the
// MAXD == 1 variant instead, which parks the same words in shared memory and keeps the hit chain in one register.
struct LevelSave {
    LevelRay r;  // the caller's ray with its slab reciprocals and shear: restored, not recomputed (2 divisions + 4 reciprocals)
    const DevNode* nodes;
    const float4* leafs;
    const DevScene* scene;  // descriptor of the caller's level (nullptr = top level)
    uint32_t resume;        // next record of the instancing leaf, or CRT_NO_REC
    uint32_t inst;          // the instance record (offset in the caller's leafs), for the hit chain
};

// Capacity of the traversal stack per ray (entries). A stack never needs more than 3 entries per BVH4 level plus one, and
// the builders bound the depth (bvh.rs:152 MAX_DEPTH = 60 binary levels on the host; Morton-code length on the device),
// so these are upper bounds, not typical sizes (measured depth at 10 M triangles: 23): the first CRT_SSTACK entries
// live in shared memory, the rest in local memory that only deep trees ever touch. The reference spills to the heap
// (bvh.rs:711-755) and is unbounded; here an overflow beyond the bound sets a sticky status bit.
template <int MAXD>
struct StackCap {
#ifdef CRT_STACK_CAP_ALL  // test builds only: proves that tests/test_gpu_parity.py's deep-stack scene needs more than the old 128
    static constexpr int value = CRT_STACK_CAP_ALL;
#else
    static constexpr int value = MAXD == 0 ? 512 : (MAXD == 1 ? 1024 : 2048);
#endif
};

template <bool ANY_HIT, uint32_t FEAT, int MAXD>
struct MinBlocks {
    static constexpr int value = MAXD > 1 ? CRT_MINBLOCKS_INST : (MAXD == 1 ? (ANY_HIT ? CRT_MINBLOCKS_INST1_ANY : CRT_MINBLOCKS_INST1) : (ANY_HIT ? CRT_MINBLOCKS_ANY : CRT_MINBLOCKS_CLOSEST));
};

// COUNT: the counting build of the same kernel (crt_trace_counters): what the traversal actually touches, summed over the
// launch — nodes fetched, popped entries culled by their entry distance, leaves entered, triangle records tested, other
// primitive records tested, instance descents — for the bytes-per-ray figure of bench.py. Never used by a timed launch.
template <bool ANY_HIT, uint32_t FEAT, int MAXD, bool COUNT>
__global__ void __launch_bounds__(CRT_BLOCK, MinBlocks<ANY_HIT, FEAT, MAXD>::value) trace_kernel(const TraceParams p) {
    typedef typename StackT<ANY_HIT>::type Entry;
    constexpr bool INST = MAXD > 0;
    static_assert(INST == ((FEAT & FEAT_INSTANCE) != 0), "MAXD > 0 exactly for the variants with instance descent");
    constexpr int kLStack = StackCap<MAXD>::value - CRT_SSTACK;
    __shared__ Entry sstack[CRT_SSTACK][CRT_BLOCK];
    // MAXD == 1: the top-level ray parked during an instance descent (o, d, inv, shear, packed axes, resume record)
    __shared__ uint32_t ssave[MAXD == 1 ? 14 : 1][MAXD == 1 ? CRT_BLOCK : 1];
    Entry lstack[kLStack];  // spill (local memory, touched only by deep trees)

    const unsigned lane = threadIdx.x & 31u;
    const unsigned tid = threadIdx.x;
    // wavefront queues pass their length by device pointer (written by the producing kernel)
    const uint32_t n_rays = p.n_dev ? min(__ldg(p.n_dev), p.n) : p.n;

    // ---- per-ray state -------------------------------------------------------
    enum : int { IDLE = 0, RUN = 1, DONE = 2 };  // DONE = traversal finished, result not yet written
    int state = IDLE;
    uint32_t ray_idx = 0;
    LevelRay r;  // at depth 0 this is the ray as given (the world ray): finalize() relies on it
    float t_min = 0.0f, closest = 0.0f, time = 0.0f;
    [[maybe_unused]] bool tmin_nonneg = true;  // t_min >= +0: entry distances order by their bit patterns (node_visit)
    uint32_t ray_mask = 0;
    const DevNode* nodes = nullptr;
    const float4* leafs = nullptr;
    int sp = 0;
    uint32_t rec = CRT_NO_REC;  // leaf mode: next record to test (16-byte units from `leafs`)
    // packet state (bvh.rs:530-562)
    float pkt_tmax = 0.0f;
    uint32_t pkt_deferred = 0;
    // best hit: record offset in the leafs of the level it was found at
    uint32_t best_rec = CRT_NO_REC;
    F3 best_n = f3(0, 0, 0);
    bool any_found = false;
    bool overflow = false;  // a stack entry was dropped (reported once, at kernel exit)
    [[maybe_unused]] uint32_t c_nodes = 0, c_culled = 0, c_leaves = 0, c_tris = 0, c_prims = 0, c_inst = 0, c_rays = 0;  // COUNT only
    // instancing
    int depth = 0;
    int best_depth = 0;
    [[maybe_unused]] uint32_t cur_inst = CRT_NO_REC, best_inst = CRT_NO_REC;  // MAXD == 1: the instance record being traversed / of the best hit
    [[maybe_unused]] const DevScene* cur_scene = nullptr;                    // MAXD > 1: descriptor of the current level (nullptr = top)
    LevelSave saved[MAXD > 1 ? MAXD : 1];
    uint32_t best_chain[MAXD > 1 ? MAXD : 1];

#if CRT_SADDR
    // The thread's column of the shared stack as a 32-bit shared-window address held in a register: without it the
    // compiler rebuilds the address (CTA window base + tid * 8) from special registers at every push and pop.
    uint32_t sbase;
    {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(&sstack[0][tid]);
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(a));
    }
    constexpr uint32_t kSlotBytes = CRT_BLOCK * sizeof(Entry);
    auto sst = [&](int slot, Entry e) {
        if constexpr (ANY_HIT)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + (uint32_t)slot * kSlotBytes), "r"(e) : "memory");
        else
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sbase + (uint32_t)slot * kSlotBytes), "r"(e.x), "r"(e.y) : "memory");
    };
    auto sld = [&](int slot) -> Entry {
        Entry e;
        if constexpr (ANY_HIT)
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(sbase + (uint32_t)slot * kSlotBytes) : "memory");
        else
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(sbase + (uint32_t)slot * kSlotBytes) : "memory");
        return e;
    };
#else
    auto sst = [&](int slot, Entry e) { sstack[slot][tid] = e; };
    auto sld = [&](int slot) -> Entry { return sstack[slot][tid]; };
#endif
    // Beyond the depth bound of any tree the builders emit. A dropped entry could be an instance-return marker, after which
    // the walk is meaningless, so the ray is abandoned (it reports a miss) and the sticky bit of crt_device_status() is set.
    auto abandon_ray = [&]() {
        overflow = true;
        sp = 0;
        depth = 0;
        rec = CRT_NO_REC;
        best_rec = CRT_NO_REC;
        best_inst = CRT_NO_REC;
        best_depth = 0;
        any_found = false;
    };
    auto push = [&](Entry e) {
        if (sp < CRT_SSTACK)
            sst(sp, e);
        else if (sp < CRT_SSTACK + kLStack)
            lstack[sp - CRT_SSTACK] = e;
        else
            return abandon_ray();
        ++sp;
    };
    auto pop = [&]() -> Entry {
        --sp;
        if (sp < CRT_SSTACK) return sld(sp);
        if (sp < CRT_SSTACK + kLStack) return lstack[sp - CRT_SSTACK];
        Entry e;
        if constexpr (ANY_HIT) e = 0u; else e = make_uint2(0u, 0x7f800000u);
        return e;
    };
    auto make_entry = [&](uint32_t child, float tn) -> Entry {
        if constexpr (ANY_HIT) return child; else return make_uint2(child, __float_as_uint(tn));
    };
    auto accept = [&](uint32_t at, float t) {
        closest = t;
        best_rec = at;
        if constexpr (MAXD == 1) best_inst = cur_inst;
        if constexpr (MAXD > 1) {
            best_depth = depth;
            for (int k = 0; k < depth; ++k) best_chain[k] = saved[k].inst;
        }
    };

    // ---- result of a finished ray (runs for all DONE lanes of a warp together) ----
    // A closest-hit ray is DONE only with an empty stack, i.e. back at depth 0: `r` holds the world ray again.
    auto finalize = [&]() {
        if constexpr (ANY_HIT) {
            p.flag[ray_idx] = any_found ? 1 : 0;
        } else {
            float4 h0, h1;
            uint8_t ff = 0;
            if (best_rec == CRT_NO_REC) {
                h0 = make_float4(__int_as_float(0x7f800000), 0.0f, 0.0f, __uint_as_float(CRT_INVALID_ID));
                h1 = make_float4(__uint_as_float(CRT_INVALID_ID), 0.0f, 0.0f, 0.0f);
            } else {
                // Re-derive the local ray at the hit level (same arithmetic as the descent).
                LevelRay lr;
                const F3 world_d = r.d;
                F3 o = r.o, d = r.d;
                const DevScene* sc = nullptr;  // descriptor of the hit level (nullptr = top)
                const float4* hit_leafs = (const float4*)p.scene.leafs;
                Xf w2l[INST ? MAXD : 1];
                if constexpr (MAXD == 1) {
                    if (best_inst != CRT_NO_REC) {
                        const float4* ir = (const float4*)p.scene.leafs + best_inst;
                        w2l[0] = instance_w2l(p.scene, ir, time);
                        o = xf_point(w2l[0], o);
                        d = xf_vec(w2l[0], d);
                        sc = instance_child(ir);
                        hit_leafs = instance_child_leafs(ir);
                        best_depth = 1;
                    }
                }
                if constexpr (MAXD > 1) {
                    for (int k = 0; k < best_depth; ++k) {
                        const DevScene& owner = sc ? *sc : p.scene;
                        const float4* ir = (const float4*)owner.leafs + best_chain[k];
                        w2l[k] = instance_w2l(owner, ir, time);
                        o = xf_point(w2l[k], o);
                        d = xf_vec(w2l[k], d);
                        sc = instance_child(ir);
                        hit_leafs = instance_child_leafs(ir);
                    }
                }
                const DevScene& hs = sc ? *sc : p.scene;
                const float4* hit_rec = hit_leafs + best_rec;
                float4 q0 = __ldg(hit_rec);
                uint32_t flags = __float_as_uint(q0.w);
                uint32_t kind = flags & REC_KIND_MASK;
                float u = 0.0f, v = 0.0f, t = closest;
                F3 n = best_n;
                uint32_t prim;
                if (kind == REC_TRI) {
                    float4 q1 = __ldg(hit_rec + 1), q2 = __ldg(hit_rec + 2);
                    prim = __float_as_uint(q1.w);
                    if constexpr (MAXD == 0)
                        lr = r;  // no instancing: the loop's ray is the hit level's ray, shear constants included
                    else
                        lr.set(o, d, true);
                    F3 v0 = xyz(q0), v1 = xyz(q1), v2 = xyz(q2);
                    TriEval e = tri_eval_exact(lr.o, lr.d, lr.sx, lr.sy, lr.sz, lr.kz, v0, v1, v2);
                    float inv_det = 1.0f / e.det;
                    t = e.t_scaled * inv_det;
                    u = e.e1 * inv_det;
                    v = e.e2 * inv_det;
                    if ((FEAT & FEAT_NORMALS) && (flags & REC_HAS_NORMALS)) {  // prim.rs:77-78
                        const float* nn = hs.normals + 9 * (size_t)prim;
                        n = normalize((ld3(nn) * (1.0f - u - v) + ld3(nn + 3) * u) + ld3(nn + 6) * v);
                    } else {
                        n = normalize(cross(v1 - v0, v2 - v0));  // prim.rs:80-84
                    }
                } else if (kind == REC_SPHERE) {
                    prim = __float_as_uint(__ldg(hit_rec + 1).y);
                } else {
                    prim = __float_as_uint(__ldg(hit_rec + 1).w);
                }
                uint2 ids = __ldg((const uint2*)hs.prim_ids + prim);
                uint32_t geom_id = ids.x;
                if constexpr (INST) {
                    const DevScene* oc = nullptr;
                    uint32_t chain_geom[INST ? MAXD : 1];
                    for (int k = 0; k < best_depth; ++k) {  // geom_id of every instance on the chain (outermost first)
                        const DevScene& owner = oc ? *oc : p.scene;
                        uint32_t at;
                        if constexpr (MAXD == 1) at = best_inst; else at = best_chain[k];
                        const float4* ir = (const float4*)owner.leafs + at;
                        chain_geom[k] = __float_as_uint(__ldg(ir + 4).z);
                        if (k + 1 < best_depth) oc = instance_child(ir);
                    }
                    for (int k = best_depth - 1; k >= 0; --k) {  // prim.rs:359-363, innermost first
                        const Xf& m = w2l[k];                    // normal_mat = w2l.matrix3^T
                        F3 nx = f3(m.x.x, m.y.x, m.z.x), ny = f3(m.x.y, m.y.y, m.z.y), nz = f3(m.x.z, m.y.z, m.z.z);
                        n = normalize((nx * n.x + ny * n.y) + nz * n.z);
                        geom_id = chain_geom[k];
                    }
                }
                bool front = dot(world_d, n) < 0.0f;  // scene.rs:356-359
                if (!front) n = f3(-n.x, -n.y, -n.z);
                ff = front ? 1 : 0;
                h0 = make_float4(t, u, v, __uint_as_float(geom_id));
                h1 = make_float4(__uint_as_float(ids.y), n.x, n.y, n.z);
            }
            p.hits[2 * (size_t)ray_idx] = h0;
            p.hits[2 * (size_t)ray_idx + 1] = h1;
            if (p.flag) p.flag[ray_idx] = ff;
        }
    };

    // ---- node visit: 4-wide slab test (bvh.rs:790-808) + ordered push ------------
    auto node_visit = [&](uint32_t child) {
        if constexpr (COUNT) ++c_nodes;
        float tn[4];
        [[maybe_unused]] float tf[4];
        bool hit[4];
        uint32_t cw[4];
#if CRT_SIGNSEL
        // The near plane of an axis is the box's low face when the ray travels in +axis and the high face otherwise, so the
        // planes are picked by address (three byte offsets per ray) instead of by 24 min / max. Bit-identical to
        // min(t0, t1) / max(t0, t1): lo <= hi and rounded subtraction and multiplication are monotone, safe_inv keeps
        // infinities and NaN out of `inv`, and the sign of a zero plays no part in the comparisons that follow
        // (tests/test_slab_signselect.py). Empty lanes of the device copy carry lo = +inf, hi = -inf (scene.cu
        // mark_empty_lanes), which fails `near <= far` for every ray, so the child word need not be looked at here.
        const char* nb = (const char*)(nodes + child);
        const float4 nx4 = __ldg((const float4*)(nb + r.near_x)), ny4 = __ldg((const float4*)(nb + r.near_y)),
                     nz4 = __ldg((const float4*)(nb + r.near_z));
        const float4 fx4 = __ldg((const float4*)(nb + (48u - r.near_x))), fy4 = __ldg((const float4*)(nb + (80u - r.near_y))),
                     fz4 = __ldg((const float4*)(nb + (112u - r.near_z)));
        const uint4 ch = __ldg((const uint4*)(nb + 96));
        cw[0] = ch.x, cw[1] = ch.y, cw[2] = ch.z, cw[3] = ch.w;
        const float nx[4] = {nx4.x, nx4.y, nx4.z, nx4.w}, ny[4] = {ny4.x, ny4.y, ny4.z, ny4.w}, nz[4] = {nz4.x, nz4.y, nz4.z, nz4.w};
        const float fx[4] = {fx4.x, fx4.y, fx4.z, fx4.w}, fy[4] = {fy4.x, fy4.y, fy4.z, fy4.w}, fz[4] = {fz4.x, fz4.y, fz4.z, fz4.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            float tnx = (nx[l] - r.o.x) * r.inv.x, tfx = (fx[l] - r.o.x) * r.inv.x;
            float tny = (ny[l] - r.o.y) * r.inv.y, tfy = (fy[l] - r.o.y) * r.inv.y;
            float tnz = (nz[l] - r.o.z) * r.inv.z, tfz = (fz[l] - r.o.z) * r.inv.z;
            float near_ = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), t_min);
            float far_ = fminf(fminf(fminf(tfx, tfy), tfz), closest);
            tn[l] = near_;
            tf[l] = far_;
            hit[l] = near_ <= far_;
        }
#else
        const float4* np = (const float4*)(nodes + child);
        float4 lox = __ldg(np), loy = __ldg(np + 1), loz = __ldg(np + 2);
        float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
        uint4 ch = __ldg((const uint4*)(np + 6));
        cw[0] = ch.x, cw[1] = ch.y, cw[2] = ch.z, cw[3] = ch.w;
        const float lx[4] = {lox.x, lox.y, lox.z, lox.w}, ly[4] = {loy.x, loy.y, loy.z, loy.w},
                    lz[4] = {loz.x, loz.y, loz.z, loz.w};
        const float hx[4] = {hix.x, hix.y, hix.z, hix.w}, hy[4] = {hiy.x, hiy.y, hiy.z, hiy.w},
                    hz[4] = {hiz.x, hiz.y, hiz.z, hiz.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            float t0x = (lx[l] - r.o.x) * r.inv.x, t1x = (hx[l] - r.o.x) * r.inv.x;
            float t0y = (ly[l] - r.o.y) * r.inv.y, t1y = (hy[l] - r.o.y) * r.inv.y;
            float t0z = (lz[l] - r.o.z) * r.inv.z, t1z = (hz[l] - r.o.z) * r.inv.z;
            float near_ = fmaxf(fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z)), t_min);
            float far_ = fminf(fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z)), closest);
            tn[l] = near_;
            tf[l] = far_;
            hit[l] = (near_ <= far_) & (cw[l] != CRT_EMPTY_LANE);
        }
#endif
        if constexpr (ANY_HIT) {
            // bvh.rs:598-608 — no ordering needed for a boolean answer; queue internal
            // lanes first so leaf lanes (the cheap early exits) pop first.
#if CRT_SADDR && CRT_KEY_ADDR && CRT_ANY_ORDER
            // A boolean answer does not depend on the visiting order, its cost does: the lanes are queued by a per-lane key
            // (leaf lanes still pop first) — 1: nearest entry first, 2: longest overlap with the ray's range first, 3: farthest first.
            if (sp + 4 <= CRT_SSTACK && (CRT_ANY_ORDER == 2 || tmin_nonneg)) {
                uint32_t u[4], hb[4];
                // 4: a query with a finite far end is a visibility test towards a point (a light): what blocks it for many
                // origins sits near that point, so its lanes are visited farthest first; an open-ended one nearest first
                [[maybe_unused]] const uint32_t far_first = closest < __int_as_float(0x7f800000) ? 0x7FFFFFFFu : 0u;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    uint32_t k;
                    if (CRT_ANY_ORDER == 4) k = __float_as_uint(tn[l]) ^ far_first;
                    else if (CRT_ANY_ORDER == 1) k = __float_as_uint(tn[l]);
                    else if (CRT_ANY_ORDER == 2) k = 0x7FFFFFFFu - (__float_as_uint(tf[l] - tn[l]) & 0x7FFFFFFFu);
                    else k = 0x7FFFFFFFu - (__float_as_uint(tn[l]) & 0x7FFFFFFFu);
                    u[l] = k | (~cw[l] & CRT_LEAF_BIT);
                    hb[l] = hit[l] ? kSlotBytes : 0u;
                }
                const uint32_t nbytes = (hb[0] + hb[1]) + (hb[2] + hb[3]);
                const uint32_t top = sbase + (uint32_t)sp * kSlotBytes + nbytes - kSlotBytes;
                uint32_t a0 = top, a1 = top, a2 = top, a3 = top;
                if (u[0] > u[1]) a0 -= hb[1]; else a1 -= hb[0];
                if (u[0] > u[2]) a0 -= hb[2]; else a2 -= hb[0];
                if (u[0] > u[3]) a0 -= hb[3]; else a3 -= hb[0];
                if (u[1] > u[2]) a1 -= hb[2]; else a2 -= hb[1];
                if (u[1] > u[3]) a1 -= hb[3]; else a3 -= hb[1];
                if (u[2] > u[3]) a2 -= hb[3]; else a3 -= hb[2];
                if (hit[0]) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a0), "r"(cw[0]) : "memory");
                if (hit[1]) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a1), "r"(cw[1]) : "memory");
                if (hit[2]) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a2), "r"(cw[2]) : "memory");
                if (hit[3]) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a3), "r"(cw[3]) : "memory");
                sp += (int)(nbytes / kSlotBytes);
                return;
            }
#endif
#if CRT_SADDR && CRT_KEY_ADDR
            if (sp + 4 <= CRT_SSTACK) {
                // Hot path, every slot in shared memory: two running shared-window addresses (internal lanes from the group's
                // base, leaf lanes above all of them), each advanced by a slot for a hit lane of its class — selects and adds,
                // no per-lane branch, no bounds test. Same slots as the eight conditional pushes below.
                uint32_t ib[4], lb[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const bool leaf = (cw[l] & CRT_LEAF_BIT) != 0u;
                    ib[l] = (hit[l] && !leaf) ? kSlotBytes : 0u;
                    lb[l] = (hit[l] && leaf) ? kSlotBytes : 0u;
                }
                const uint32_t b0 = sbase + (uint32_t)sp * kSlotBytes;
                uint32_t ai = b0, al = b0 + ((ib[0] + ib[1]) + (ib[2] + ib[3]));
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t at = (cw[l] & CRT_LEAF_BIT) ? al : ai;
                    if (hit[l]) asm volatile("st.shared.u32 [%0], %1;" ::"r"(at), "r"(cw[l]) : "memory");
                    ai += ib[l];
                    al += lb[l];
                }
                sp += (int)((al - b0) / kSlotBytes);
                return;
            }
#endif
#pragma unroll
            for (int l = 0; l < 4; ++l)
                if (hit[l] && !(cw[l] & CRT_LEAF_BIT)) push(make_entry(cw[l], 0.0f));
#pragma unroll
            for (int l = 0; l < 4; ++l)
                if (hit[l] && (cw[l] & CRT_LEAF_BIT)) push(make_entry(cw[l], 0.0f));
        } else {
            // Near-to-far order with the reference's stable insertion sort
            // (bvh.rs:472-486): lane m precedes lane l iff tn[m] < tn[l], or equal and
            // m < l. Internal lanes go on the stack far-to-near, then the leaf lanes
            // far-to-near on top of them (bvh.rs:491-505). Branch-free: with g(m,l) = [tn[m] > tn[l]] "before"
            // masks are assembled from the six pair bits, and each lane's slot is its rank
            // among the hit lanes of its own class.
            auto hit_mask = [&]() {  // built only on the paths that rank with masks: the hot path below never needs it
                uint32_t m = 0;
#pragma unroll
                for (int l = 0; l < 4; ++l) m |= hit[l] ? (1u << l) : 0u;
                return m;
            };
            const int base = sp;
#if CRT_KEY_ORDER
            if (tmin_nonneg) {
                // Every entry distance is >= t_min >= +0, so its bit pattern orders like the number and the sign bit is free:
                // with the sign bit set on INTERNAL lanes the unsigned keys order the whole pop sequence at once — leaf lanes
                // first, each class near to far — and one rank among all hit lanes replaces the two per-class ranks (no leaf mask,
                // no class counts, no per-lane mask select: -20 instructions per visit). Same order as below by construction:
                // within a class the keys compare like tn, ties fall to the lane index through the >= / > asymmetry of the pair bits.
                uint32_t u[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) u[l] = __float_as_uint(tn[l]) | (~cw[l] & CRT_LEAF_BIT);
#if CRT_SADDR && CRT_KEY_ADDR
                if (base + 4 <= CRT_SSTACK) {
                    // Hot path, every slot in shared memory: the slots are computed as shared-window byte addresses. All hit
                    // lanes start at the top slot of the group; for every pair the lane that is popped later moves one slot
                    // down if the other lane is a hit — six compares, twelve predicated subtractions, no masks, no popcounts.
                    uint32_t hb[4];
#pragma unroll
                    for (int l = 0; l < 4; ++l) hb[l] = hit[l] ? kSlotBytes : 0u;
                    const uint32_t nbytes = (hb[0] + hb[1]) + (hb[2] + hb[3]);
                    const uint32_t top = sbase + (uint32_t)base * kSlotBytes + nbytes - kSlotBytes;
                    uint32_t a0 = top, a1 = top, a2 = top, a3 = top;
                    if (u[0] > u[1]) a0 -= hb[1]; else a1 -= hb[0];  // u[a] > u[b] (a < b): b is popped before a; else a before b
                    if (u[0] > u[2]) a0 -= hb[2]; else a2 -= hb[0];
                    if (u[0] > u[3]) a0 -= hb[3]; else a3 -= hb[0];
                    if (u[1] > u[2]) a1 -= hb[2]; else a2 -= hb[1];
                    if (u[1] > u[3]) a1 -= hb[3]; else a3 -= hb[1];
                    if (u[2] > u[3]) a2 -= hb[3]; else a3 -= hb[2];
                    if (hit[0]) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a0), "r"(cw[0]), "r"(__float_as_uint(tn[0])) : "memory");
                    if (hit[1]) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a1), "r"(cw[1]), "r"(__float_as_uint(tn[1])) : "memory");
                    if (hit[2]) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a2), "r"(cw[2]), "r"(__float_as_uint(tn[2])) : "memory");
                    if (hit[3]) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a3), "r"(cw[3]), "r"(__float_as_uint(tn[3])) : "memory");
                    sp = base + (int)(nbytes / kSlotBytes);
                    return;
                }
#endif
                const uint32_t g01 = u[0] > u[1] ? 1u : 0u, g02 = u[0] > u[2] ? 1u : 0u, g03 = u[0] > u[3] ? 1u : 0u,
                               g12 = u[1] > u[2] ? 1u : 0u, g13 = u[1] > u[3] ? 1u : 0u, g23 = u[2] > u[3] ? 1u : 0u;
                uint32_t B[4];  // B[l]: the lanes popped before lane l
                B[0] = (g01 << 1) | (g02 << 2) | (g03 << 3);
                B[1] = (g01 ^ 1u) | (g12 << 2) | (g13 << 3);
                B[2] = (g02 ^ 1u) | ((g12 ^ 1u) << 1) | (g23 << 3);
                B[3] = (g03 ^ 1u) | ((g13 ^ 1u) << 1) | ((g23 ^ 1u) << 2);
                const uint32_t H = hit_mask();
                const int n_hit = __popc(H);
                if (base + 4 <= CRT_SSTACK) {  // hot path: every slot is in shared memory
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const int slot = base + n_hit - 1 - __popc(B[l] & H);
                        if (hit[l]) sst(slot, make_entry(cw[l], tn[l]));
                    }
                } else {
                    bool fits = true;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        if (!hit[l]) continue;
                        const int slot = base + n_hit - 1 - __popc(B[l] & H);
                        Entry e2 = make_entry(cw[l], tn[l]);
                        if (slot < CRT_SSTACK)
                            sst(slot, e2);
                        else if (slot < CRT_SSTACK + kLStack)
                            lstack[slot - CRT_SSTACK] = e2;
                        else
                            fits = false;
                    }
                    if (!fits) return abandon_ray();
                }
                sp = base + n_hit;
                return;
            }
#endif
            // Rays with a negative t_min (entry distances may be negative): ranks per class from float compares.
            const uint32_t H = hit_mask();
            uint32_t Lm = 0;
#pragma unroll
            for (int l = 0; l < 4; ++l) Lm |= (cw[l] & CRT_LEAF_BIT) ? (1u << l) : 0u;
            const uint32_t g01 = tn[0] > tn[1] ? 1u : 0u, g02 = tn[0] > tn[2] ? 1u : 0u, g03 = tn[0] > tn[3] ? 1u : 0u,
                           g12 = tn[1] > tn[2] ? 1u : 0u, g13 = tn[1] > tn[3] ? 1u : 0u, g23 = tn[2] > tn[3] ? 1u : 0u;
            uint32_t B[4];
            B[0] = (g01 << 1) | (g02 << 2) | (g03 << 3);
            B[1] = (g01 ^ 1u) | (g12 << 2) | (g13 << 3);
            B[2] = (g02 ^ 1u) | ((g12 ^ 1u) << 1) | (g23 << 3);
            B[3] = (g03 ^ 1u) | ((g13 ^ 1u) << 1) | ((g23 ^ 1u) << 2);
            const uint32_t HL = H & Lm, HI = H & ~Lm;
            const int n_int = __popc(HI), n_leaf = __popc(HL);
            bool fits = true;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                if (!hit[l]) continue;
                const bool lf = (Lm >> l) & 1u;
                const int rank = __popc(B[l] & (lf ? HL : HI));
                const int slot = base + (lf ? n_int + n_leaf - 1 : n_int - 1) - rank;
                Entry e2 = make_entry(cw[l], tn[l]);
                if (slot < CRT_SSTACK)
                    sst(slot, e2);
                else if (slot < CRT_SSTACK + kLStack)
                    lstack[slot - CRT_SSTACK] = e2;
                else
                    fits = false;
            }
            if (!fits) return abandon_ray();
            sp = base + n_int + n_leaf;
        }
    };

    // ---- one leaf record (bvh.rs:514-572 / 617-653) -------------------------------
    auto record_test = [&]() {
        const uint32_t at = rec;
        const float4* cur = leafs + at;
        // The record's first three quads in one round trip. Left alone, the compiler sinks the second and third load below
        // the kind / mask branches that need only the first, which serialises three L1 latencies per record; so the one
        // gate the visible-triangle path depends on reads all three (a primitive id is never all ones — the last term is
        // always true and only there to tie the middle quad in). Every record array is padded by one triangle record: the
        // over-read past a shorter record's tail stays inside the buffer.
        const float4 q0 = __ldg(cur), q1 = __ldg(cur + 1), q2 = __ldg(cur + 2);
        const uint32_t flags = __float_as_uint(q0.w);
        const uint32_t kind = flags & REC_KIND_MASK;
        const bool tri_visible = (kind == REC_TRI) & ((ray_mask & __float_as_uint(q2.w)) != 0u)  // triangle.rs:257-271 per-lane visibility
                                 & (__float_as_uint(q1.w) != CRT_INVALID_ID);
        if (tri_visible) {
            TriEval e = tri_eval(r, xyz(q0), xyz(q1), xyz(q2));
            if (e.on_edge) {
                pkt_deferred |= 1u << ((flags >> REC_LANE_SHIFT) & 3u);  // f64 tie-break after the packet's other lanes
            } else if (!mixed_signs(e) && e.det != 0.0f) {
                if constexpr (ANY_HIT) {
                    if (in_range_packet(e, t_min, closest)) {
                        any_found = true;
                        state = DONE;
                        rec = CRT_NO_REC;
                        return;
                    }
                } else {
                    if (in_range_packet(e, t_min, pkt_tmax)) {
                        float t = e.t_scaled * (1.0f / e.det);
                        if (!(t > closest) && !(flags & REC_DEGENERATE)) accept(at, t);  // bvh.rs:542, prim.rs:80-83
                    }
                }
            }
        }
        if (kind == REC_TRI) {
            rec = (flags & REC_LAST) ? CRT_NO_REC : at + REC_TRI_Q;
            if constexpr (COUNT) ++c_tris;
            if (flags & REC_PACKET_END) {
                // Lanes sitting exactly on an edge: scalar f64 path (bvh.rs:551-561 / 636-643). The packet's first record is
                // this one minus its lane number.
                const uint32_t pkt_first = at - REC_TRI_Q * ((flags >> REC_LANE_SHIFT) & 3u);
                while (pkt_deferred) {
                    int ln = __ffs(pkt_deferred) - 1;
                    pkt_deferred &= pkt_deferred - 1;
                    const uint32_t fat = pkt_first + REC_TRI_Q * ln;
                    const float4* fr = leafs + fat;
                    float4 f0 = __ldg(fr), f1 = __ldg(fr + 1), f2 = __ldg(fr + 2);
                    TriEval e = tri_eval_f64(r.o, r.sx, r.sy, r.sz, r.kz, xyz(f0), xyz(f1), xyz(f2));
                    if (mixed_signs(e) || e.det == 0.0f) continue;
                    if (!in_range_scalar(e, t_min, closest)) continue;
                    if constexpr (ANY_HIT) {
                        any_found = true;
                        state = DONE;
                        rec = CRT_NO_REC;
                        pkt_deferred = 0;
                        return;
                    } else {
                        if (__float_as_uint(f0.w) & REC_DEGENERATE) continue;
                        accept(fat, e.t_scaled * (1.0f / e.det));
                    }
                }
                pkt_tmax = closest;
            }
            return;
        }
        if constexpr ((FEAT & (FEAT_SPHERE | FEAT_CURVE | FEAT_INSTANCE)) != 0) {
            bool got = false;
            float t = 0.0f;
            F3 n = f3(0, 0, 0);
            if constexpr (COUNT) {
                if (kind == REC_INSTANCE) ++c_inst; else ++c_prims;
            }
            if (kind == REC_SPHERE) {
                rec = (flags & REC_LAST) ? CRT_NO_REC : at + REC_SPHERE_Q;
                if (ray_mask & __float_as_uint(q1.z)) got = sphere_hit(r, xyz(q0), q1.x, t_min, closest, t, n);
            } else if ((FEAT & FEAT_CURVE) && kind == REC_CURVE) {
                rec = (flags & REC_LAST) ? CRT_NO_REC : at + REC_CURVE_Q;
                if (ray_mask & __float_as_uint(q2.z))
                    got = rounded_cone_hit(r.o, r.d, xyz(q0), xyz(q1), q2.x, q2.y, t_min, closest, t, n);
            } else if ((FEAT & FEAT_CURVE) && kind == REC_CUBIC) {
                rec = (flags & REC_LAST) ? CRT_NO_REC : at + REC_CUBIC_Q;
                if (ray_mask & __float_as_uint(q2.w)) {
                    float4 q3 = __ldg(cur + 3), q4 = __ldg(cur + 4);
                    F3 cp[4] = {xyz(q0), xyz(q1), xyz(q2), xyz(q3)};
                    got = cubic_curve_hit(r.o, r.d, cp, q4.x, q4.y, t_min, closest, t, n);
                }
            } else if (INST && kind == REC_INSTANCE) {
                const uint32_t resume = (flags & REC_LAST) ? CRT_NO_REC : at + REC_INSTANCE_Q;
                rec = resume;
                if ((ray_mask & __float_as_uint(q2.w)) && depth < MAXD) {  // prim.rs:346-350 (commit refuses deeper nesting)
                    // The record carries the child's node and record arrays, so the descent is one round of independent
                    // loads (no dependent read of the child's descriptor).
                    const float4 q3 = __ldg(cur + 3);
                    const uint4 q5 = __ldg((const uint4*)(cur + 5));
                    Xf w2l;
                    if ((flags & REC_MOTION) && time > 0.0f) {
                        if constexpr (MAXD == 1) {
                            w2l = instance_w2l(p.scene, cur, time);
                        } else {
                            const DevScene& owner = cur_scene ? *cur_scene : p.scene;
                            w2l = instance_w2l(owner, cur, time);
                        }
                    } else {
                        w2l = Xf{xyz(q0), xyz(q1), xyz(q2), xyz(q3)};
                    }
                    if constexpr (MAXD == 1) {
                        ssave[0][tid] = __float_as_uint(r.o.x), ssave[1][tid] = __float_as_uint(r.o.y), ssave[2][tid] = __float_as_uint(r.o.z);
                        ssave[3][tid] = __float_as_uint(r.d.x), ssave[4][tid] = __float_as_uint(r.d.y), ssave[5][tid] = __float_as_uint(r.d.z);
                        ssave[6][tid] = __float_as_uint(r.inv.x), ssave[7][tid] = __float_as_uint(r.inv.y), ssave[8][tid] = __float_as_uint(r.inv.z);
                        ssave[9][tid] = __float_as_uint(r.sx), ssave[10][tid] = __float_as_uint(r.sy), ssave[11][tid] = __float_as_uint(r.sz);
                        ssave[12][tid] = r.pack_axes();
                        ssave[13][tid] = resume;
                        cur_inst = at;
                    } else {
                        LevelSave& s = saved[depth];
                        s.r = r;
                        s.nodes = nodes;
                        s.leafs = leafs;
                        s.scene = cur_scene;
                        s.resume = resume;
                        s.inst = at;
                        cur_scene = instance_child(cur);
                    }
                    ++depth;
                    nodes = (const DevNode*)(((uint64_t)q5.y << 32) | (uint64_t)q5.x);
                    leafs = (const float4*)(((uint64_t)q5.w << 32) | (uint64_t)q5.z);
                    const F3 lo = xf_point(w2l, r.o), ld = xf_vec(w2l, r.d);  // prim.rs:333-341
                    const bool child_tris = (flags & REC_CHILD_TRIS) != 0;
                    bool same_dir = false;
                    if constexpr (MAXD == 1) {
                        // An instance whose linear part is the identity (a translated copy — grids of instances) hands the
                        // child the parent's direction bit for bit; then everything derived from the direction alone —
                        // reciprocals, plane offsets and, if the top level computed them, the shear constants — is the
                        // child's too, and only the origin changes. Decided on the transformed values themselves, not on
                        // the matrix: any case where they differ (a signed zero, a NaN) takes the full setup.
                        same_dir = __float_as_uint(ld.x) == __float_as_uint(r.d.x) && __float_as_uint(ld.y) == __float_as_uint(r.d.y) &&
                                   __float_as_uint(ld.z) == __float_as_uint(r.d.z);
                    }
                    if (same_dir) {
                        r.o = lo;
                        if (child_tris && p.scene.shear_at_top == 0) r.set_shear(ld);
                    } else {
                        r.set(lo, ld, child_tris);
                    }
                    rec = CRT_NO_REC;
                    push(make_entry(CRT_RETURN_MARK, 0.0f));
                    push(make_entry(0u, t_min));  // committed child scenes are never empty (scene.rs:307-309)
                }
                return;
            } else {
                rec = CRT_NO_REC;  // unreachable: a record kind this variant was not built for
                return;
            }
            if (got) {
                if constexpr (ANY_HIT) {
                    any_found = true;
                    state = DONE;
                    rec = CRT_NO_REC;
                } else {
                    accept(at, t);
                    best_n = n;
                }
            }
        } else {
            rec = CRT_NO_REC;  // unreachable: a record kind this variant was not built for
        }
    };

    bool exhausted = false;  // warp-uniform: the ray counter has run past n
    for (;;) {
        // ---- warp-uniform housekeeping: write finished rays, refill idle lanes ----
        unsigned act = __ballot_sync(0xFFFFFFFFu, state == RUN);
        [[maybe_unused]] const int n_run = __popc(act);
        if (__popc(act) < CRT_REFILL || exhausted) {
            if (!exhausted || act == 0u) {
                if (state == DONE) {  // all finished lanes of the warp write their results together
                    finalize();
                    state = IDLE;
                }
            }
            if (exhausted) {
                if (act == 0u) break;
            } else {
                unsigned idle = ~act;
                unsigned base = 0;
                int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = atomicAdd(p.counter, (unsigned)__popc(idle));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (state == IDLE) {
                    uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < n_rays) {
                        float4 a = __ldg(p.rays + 2 * (size_t)idx), b = __ldg(p.rays + 2 * (size_t)idx + 1);
                        ray_idx = idx;
                        t_min = a.w;
                        tmin_nonneg = !(__float_as_uint(a.w) & 0x80000000u);  // not negative and not -0.0 (NaN: nothing hits anyway)
                        closest = b.w;
                        time = p.time ? __ldg(p.time + idx) : 0.0f;
                        ray_mask = p.mask ? __ldg(p.mask + idx) : p.mask_uniform;
                        r.set(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), p.scene.shear_at_top != 0);
                        nodes = p.scene.nodes;
                        leafs = (const float4*)p.scene.leafs;
                        cur_scene = nullptr;
                        cur_inst = CRT_NO_REC;
                        depth = 0;
                        best_rec = CRT_NO_REC;
                        best_inst = CRT_NO_REC;
                        best_depth = 0;
                        any_found = false;
                        rec = CRT_NO_REC;
                        sp = 0;
                        state = RUN;
                        if constexpr (COUNT) ++c_rays;
                        if (p.scene.n_nodes != 0) push(make_entry(0u, t_min));
                    }
                }
                if (base + (unsigned)__popc(idle) >= n_rays) exhausted = true;
                if (base >= n_rays && act == 0u) break;  // nothing left anywhere in this warp
            }
        }

        // ---- traversal burst (while-while): every running lane walks nodes until it
        // reaches a leaf, then tests that leaf's records; lanes reconverge after each phase.
        if (state == RUN) {
            // node phase
            while (rec == CRT_NO_REC) {
                if (sp == 0) {
                    state = DONE;
                    break;
                }
                Entry e = pop();
                uint32_t child;
                if constexpr (ANY_HIT) child = e; else child = e.x;
                if constexpr (INST) {
                    if (child == CRT_RETURN_MARK) {  // back to the instancing level
                        --depth;
                        if constexpr (MAXD == 1) {
                            nodes = p.scene.nodes;
                            leafs = (const float4*)p.scene.leafs;
                            r.o = f3(__uint_as_float(ssave[0][tid]), __uint_as_float(ssave[1][tid]), __uint_as_float(ssave[2][tid]));
                            r.d = f3(__uint_as_float(ssave[3][tid]), __uint_as_float(ssave[4][tid]), __uint_as_float(ssave[5][tid]));
                            r.inv = f3(__uint_as_float(ssave[6][tid]), __uint_as_float(ssave[7][tid]), __uint_as_float(ssave[8][tid]));
                            r.sx = __uint_as_float(ssave[9][tid]), r.sy = __uint_as_float(ssave[10][tid]), r.sz = __uint_as_float(ssave[11][tid]);
                            r.unpack_axes(ssave[12][tid]);
                            rec = ssave[13][tid];
                            cur_inst = CRT_NO_REC;
                        } else {
                            const LevelSave& s = saved[depth];
                            cur_scene = s.scene;
                            nodes = s.nodes;
                            leafs = s.leafs;
                            r = s.r;
                            rec = s.resume;
                        }
                        // the instancing leaf's packet state: instance records follow the leaf's triangles, so no packet is open
                        continue;
                    }
                }
                if (child & CRT_LEAF_BIT) {
                    if constexpr (COUNT) ++c_leaves;
                    rec = child & ~CRT_LEAF_BIT;
                    pkt_tmax = closest;
                    pkt_deferred = 0;
#if CRT_LEAF_PREFETCH
                    // the lane now waits for the rest of the warp to finish the node phase: start fetching its leaf's records
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(leafs + rec));
#if CRT_LEAF_PREFETCH > 1
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(leafs + rec + 6));
#endif
#endif
                    break;
                }
                if constexpr (!ANY_HIT) {
                    if (__uint_as_float(e.y) > closest) {  // see traverse.cuh: same result as visiting
                        if constexpr (COUNT) ++c_culled;
                        continue;
                    }
                }
                node_visit(child);
#if CRT_NODE_EXIT_NUM > 0
                // Phase balancing: when most running lanes already wait with a leaf, the stragglers stop searching and the
                // warp moves on to the leaf phase (their search resumes in the next round with the lanes that finished a leaf).
                // A scheduling heuristic only — every lane keeps its own stack and visiting order.
                if (__popc(__activemask()) * CRT_NODE_EXIT_DEN < n_run * CRT_NODE_EXIT_NUM) break;
#endif
            }
            // leaf phase
            while (state == RUN && rec != CRT_NO_REC) {
                record_test();
#if CRT_LEAF_EXIT_NUM > 0
                if (__popc(__activemask()) * CRT_LEAF_EXIT_DEN < n_run * CRT_LEAF_EXIT_NUM) break;
#endif
            }
        }
    }
    if (overflow) atomicOr(p.status, 1u);
    if constexpr (COUNT) {
        const uint32_t c[7] = {c_rays, c_nodes, c_culled, c_leaves, c_tris, c_prims, c_inst};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            uint32_t v = c[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0 && v) atomicAdd(p.stats + k, (unsigned long long)v);
        }
    }
}

template <bool ANY_HIT, uint32_t FEAT, int MAXD>
cudaError_t launch_variant(const TraceParams& p, int grid, cudaStream_t stream) {
    if (p.stats)
        trace_kernel<ANY_HIT, FEAT, MAXD, true><<<grid, CRT_BLOCK, 0, stream>>>(p);
    else
        trace_kernel<ANY_HIT, FEAT, MAXD, false><<<grid, CRT_BLOCK, 0, stream>>>(p);
    return cudaGetLastError();
}

template <bool ANY_HIT, uint32_t FEAT, int MAXD>
int blocks_per_sm_variant() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_kernel<ANY_HIT, FEAT, MAXD, false>, CRT_BLOCK, 0);
    return n > 0 ? n : 1;
}

// Feature sets actually instantiated; a scene uses the smallest superset. Instanced scenes whose nesting depth is 1 (an
// instance never contains another: every bundled sample but nested_instancing.usda) get their own variant.
constexpr uint32_t V_TRI = 0;
constexpr uint32_t V_SPH = FEAT_SPHERE | FEAT_NORMALS;
constexpr uint32_t V_INST = FEAT_SPHERE | FEAT_NORMALS | FEAT_INSTANCE;
constexpr uint32_t V_ALL = FEAT_SPHERE | FEAT_NORMALS | FEAT_INSTANCE | FEAT_CURVE;
constexpr int kVariants = 5;

int variant_of(uint32_t feat, uint32_t instance_depth) {
    if ((feat & ~V_TRI) == 0) return 0;
    if ((feat & ~V_SPH) == 0) return 1;
    if ((feat & ~V_INST) == 0) return instance_depth <= 1 ? 2 : 3;
    return 4;
}

// Calls f.template operator()<ANY_HIT, FEAT, MAXD>() for the variant (any_hit, v).
template <class F>
auto dispatch(bool any_hit, int v, F&& f) {
    if (any_hit) {
        switch (v) {
            case 0: return f.template operator()<true, V_TRI, 0>();
            case 1: return f.template operator()<true, V_SPH, 0>();
            case 2: return f.template operator()<true, V_INST, 1>();
            case 3: return f.template operator()<true, V_INST, CRT_MAX_INSTANCE_DEPTH>();
            default: return f.template operator()<true, V_ALL, CRT_MAX_INSTANCE_DEPTH>();
        }
    }
    switch (v) {
        case 0: return f.template operator()<false, V_TRI, 0>();
        case 1: return f.template operator()<false, V_SPH, 0>();
        case 2: return f.template operator()<false, V_INST, 1>();
        case 3: return f.template operator()<false, V_INST, CRT_MAX_INSTANCE_DEPTH>();
        default: return f.template operator()<false, V_ALL, CRT_MAX_INSTANCE_DEPTH>();
    }
}

struct BlocksFn {
    template <bool A, uint32_t F, int D>
    int operator()() const {
        return blocks_per_sm_variant<A, F, D>();
    }
};
struct LaunchFn {
    const TraceParams& p;
    int grid;
    cudaStream_t stream;
    template <bool A, uint32_t F, int D>
    cudaError_t operator()() const {
        return launch_variant<A, F, D>(p, grid, stream);
    }
};

__global__ void mark_empty_lanes_kernel(DevNode* nodes, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevNode& nd = nodes[i];
    const float ninf = __int_as_float(0xff800000), pinf = __int_as_float(0x7f800000);
#pragma unroll
    for (int l = 0; l < 4; ++l)
        if (nd.child[l] == CRT_EMPTY_LANE) {
            nd.bmin_x[l] = nd.bmin_y[l] = nd.bmin_z[l] = pinf;
            nd.bmax_x[l] = nd.bmax_y[l] = nd.bmax_z[l] = ninf;
        }
}

}  // namespace

cudaError_t mark_empty_lanes(DevNode* nodes, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    mark_empty_lanes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(nodes, n);
    return cudaGetLastError();
}

int trace_blocks_per_sm(bool any_hit, uint32_t feat, uint32_t instance_depth) {
    static int cache[2][kVariants] = {};
    int v = variant_of(feat, instance_depth);
    int& c = cache[any_hit ? 1 : 0][v];
    if (c == 0) c = dispatch(any_hit, v, BlocksFn{});
    return c;
}

cudaError_t launch_trace(bool any_hit, uint32_t feat, uint32_t instance_depth, const TraceParams& p, int grid, cudaStream_t stream) {
    return dispatch(any_hit, variant_of(feat, instance_depth), LaunchFn{p, grid, stream});
}

}  // namespace crt
