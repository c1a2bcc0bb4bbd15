// Persistent-warp BVH4 traversal kernels (closest-hit / any-hit) for sm_100a.
// See traverse.cuh for the semantics contract and reference citations.
#include "traverse.cuh"

namespace crt {

namespace {

// Stack entry: .x = child word (bit 31 = leaf record offset, else node index,
// 0xFFFFFFFF = "return to the instancing level"), .y = entry distance bits.
#define CRT_RETURN_MARK 0xFFFFFFFFu

template <bool ANY_HIT>
struct StackT {
    typedef uint2 type;
};
template <>
struct StackT<true> {
    typedef uint32_t type;
};

struct LevelSave {  // what an instance descent must restore (prim.rs:345-365 returns into the caller's leaf loop)
    LevelRay r;  // the caller's ray with its slab reciprocals and shear: restored, not recomputed (2 divisions + 4 reciprocals)
    const DevScene* scene;
    const float4* resume;  // next record of the instancing leaf, or nullptr
    const float4* inst;    // the instance record (for the hit chain)
};

template <bool ANY_HIT, uint32_t FEAT>
__global__ void __launch_bounds__(CRT_BLOCK, (FEAT & FEAT_INSTANCE) ? CRT_MINBLOCKS_INST : (ANY_HIT ? CRT_MINBLOCKS_ANY : CRT_MINBLOCKS_CLOSEST)) trace_kernel(const TraceParams p) {
    typedef typename StackT<ANY_HIT>::type Entry;
    __shared__ Entry sstack[CRT_SSTACK][CRT_BLOCK];
    Entry lstack[CRT_LSTACK];  // spill (local memory, touched only by deep trees)

    const unsigned lane = threadIdx.x & 31u;
    const unsigned tid = threadIdx.x;
    // wavefront queues pass their length by device pointer (written by the producing kernel)
    const uint32_t n_rays = p.n_dev ? min(__ldg(p.n_dev), p.n) : p.n;
    constexpr bool INST = (FEAT & FEAT_INSTANCE) != 0;

    // ---- per-ray state -------------------------------------------------------
    enum : int { IDLE = 0, RUN = 1, DONE = 2 };  // DONE = traversal finished, result not yet written
    int state = IDLE;
    uint32_t ray_idx = 0;
    LevelRay r;
    float t_min = 0.0f, closest = 0.0f, time = 0.0f;
    uint32_t ray_mask = 0;
    const DevNode* nodes = nullptr;
    const float4* leafs = nullptr;
    const DevScene* cur_scene = nullptr;  // nullptr = top level (p.scene)
    int sp = 0;
    const float4* rec = nullptr;  // leaf mode: next record to test
    // packet state (bvh.rs:530-562)
    float pkt_tmax = 0.0f;
    const float4* pkt_first = nullptr;
    uint32_t pkt_deferred = 0;
    // best hit
    const float4* best_rec = nullptr;
    F3 best_n = f3(0, 0, 0);
    bool any_found = false;
    bool overflow = false;  // a stack entry was dropped (reported once, at kernel exit)
    // instancing
    int depth = 0;
    int best_depth = 0;
    LevelSave saved[INST ? CRT_MAX_INSTANCE_DEPTH : 1];
    const float4* best_chain[INST ? CRT_MAX_INSTANCE_DEPTH : 1];
    F3 world_o = f3(0, 0, 0), world_d = f3(0, 0, 0);

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
    auto push = [&](Entry e) {
        if (sp < CRT_SSTACK)
            sst(sp, e);
        else if (sp < CRT_SSTACK + CRT_LSTACK)
            lstack[sp - CRT_SSTACK] = e;
        else
            overflow = true;  // entry dropped; reported through crt_device_status()
        ++sp;
    };
    auto pop = [&]() -> Entry {
        --sp;
        if (sp < CRT_SSTACK) return sld(sp);
        if (sp < CRT_SSTACK + CRT_LSTACK) return lstack[sp - CRT_SSTACK];
        Entry e;
        if constexpr (ANY_HIT) e = 0u; else e = make_uint2(0u, 0x7f800000u);
        return e;
    };
    auto make_entry = [&](uint32_t child, float tn) -> Entry {
        if constexpr (ANY_HIT) return child; else return make_uint2(child, __float_as_uint(tn));
    };
    auto accept = [&](const float4* at, float t) {
        closest = t;
        best_rec = at;
        if constexpr (INST) {
            best_depth = depth;
            for (int k = 0; k < depth; ++k) best_chain[k] = saved[k].inst;
        }
    };

    // ---- result of a finished ray (runs for all DONE lanes of a warp together) ----
    auto finalize = [&]() {
        if constexpr (ANY_HIT) {
            p.flag[ray_idx] = any_found ? 1 : 0;
        } else {
            float4 h0, h1;
            uint8_t ff = 0;
            if (best_rec == nullptr) {
                h0 = make_float4(__int_as_float(0x7f800000), 0.0f, 0.0f, __uint_as_float(CRT_INVALID_ID));
                h1 = make_float4(__uint_as_float(CRT_INVALID_ID), 0.0f, 0.0f, 0.0f);
            } else {
                // Re-derive the local ray at the hit level (same arithmetic as the descent).
                LevelRay lr;
                const DevScene* sc = nullptr;
                F3 o = world_o, d = world_d;
                Xf w2l[INST ? CRT_MAX_INSTANCE_DEPTH : 1];
                if constexpr (INST) {
                    for (int k = 0; k < best_depth; ++k) {
                        const DevScene& owner = sc ? *sc : p.scene;
                        w2l[k] = instance_w2l(owner, best_chain[k], time);
                        o = xf_point(w2l[k], o);
                        d = xf_vec(w2l[k], d);
                        sc = instance_child(best_chain[k]);
                    }
                }
                const DevScene& hs = sc ? *sc : p.scene;
                float4 q0 = __ldg(best_rec);
                uint32_t flags = __float_as_uint(q0.w);
                uint32_t kind = flags & REC_KIND_MASK;
                float u = 0.0f, v = 0.0f, t = closest;
                F3 n = best_n;
                uint32_t prim;
                if (kind == REC_TRI) {
                    float4 q1 = __ldg(best_rec + 1), q2 = __ldg(best_rec + 2);
                    prim = __float_as_uint(q1.w);
                    lr.set(o, d, true);
                    F3 v0 = xyz(q0), v1 = xyz(q1), v2 = xyz(q2);
                    TriEval e = tri_eval(lr, v0, v1, v2);
                    if (e.on_edge) e = tri_eval_f64(lr, v0, v1, v2);
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
                    prim = __float_as_uint(__ldg(best_rec + 1).y);
                } else {
                    prim = __float_as_uint(__ldg(best_rec + 1).w);
                }
                uint2 ids = __ldg((const uint2*)hs.prim_ids + prim);
                uint32_t geom_id = ids.x;
                if constexpr (INST) {
                    for (int k = best_depth - 1; k >= 0; --k) {  // prim.rs:359-363, innermost first
                        const Xf& m = w2l[k];                    // normal_mat = w2l.matrix3^T
                        F3 nx = f3(m.x.x, m.y.x, m.z.x), ny = f3(m.x.y, m.y.y, m.z.y), nz = f3(m.x.z, m.y.z, m.z.z);
                        n = normalize((nx * n.x + ny * n.y) + nz * n.z);
                        geom_id = __float_as_uint(__ldg(best_chain[k] + 4).z);
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
        const float4* np = (const float4*)(nodes + child);
        float4 lox = __ldg(np), loy = __ldg(np + 1), loz = __ldg(np + 2);
        float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
        uint4 ch = __ldg((const uint4*)(np + 6));
        float tn[4];
        uint32_t cw[4] = {ch.x, ch.y, ch.z, ch.w};
        bool hit[4];
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
            hit[l] = (near_ <= far_) & (cw[l] != CRT_EMPTY_LANE);
        }
        if constexpr (ANY_HIT) {
            // bvh.rs:598-608 — no ordering needed for a boolean answer; queue internal
            // lanes first so leaf lanes (the cheap early exits) pop first.
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
            uint32_t H = 0, Lm = 0;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                H |= hit[l] ? (1u << l) : 0u;
                Lm |= (cw[l] & CRT_LEAF_BIT) ? (1u << l) : 0u;
            }
            const uint32_t g01 = tn[0] > tn[1] ? 1u : 0u, g02 = tn[0] > tn[2] ? 1u : 0u, g03 = tn[0] > tn[3] ? 1u : 0u,
                           g12 = tn[1] > tn[2] ? 1u : 0u, g13 = tn[1] > tn[3] ? 1u : 0u, g23 = tn[2] > tn[3] ? 1u : 0u;
            uint32_t B[4];
            B[0] = (g01 << 1) | (g02 << 2) | (g03 << 3);
            B[1] = (g01 ^ 1u) | (g12 << 2) | (g13 << 3);
            B[2] = (g02 ^ 1u) | ((g12 ^ 1u) << 1) | (g23 << 3);
            B[3] = (g03 ^ 1u) | ((g13 ^ 1u) << 1) | ((g23 ^ 1u) << 2);
            const uint32_t HL = H & Lm, HI = H & ~Lm;
            const int n_int = __popc(HI), n_leaf = __popc(HL);
            const int base = sp;
            if (base + 4 <= CRT_SSTACK) {  // hot path: every slot is in shared memory
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const bool lf = (Lm >> l) & 1u;
                    const int rank = __popc(B[l] & (lf ? HL : HI));
                    const int slot = base + (lf ? n_int + n_leaf - 1 : n_int - 1) - rank;
                    if (hit[l]) sst(slot, make_entry(cw[l], tn[l]));
                }
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (!hit[l]) continue;
                    const bool lf = (Lm >> l) & 1u;
                    const int rank = __popc(B[l] & (lf ? HL : HI));
                    const int slot = base + (lf ? n_int + n_leaf - 1 : n_int - 1) - rank;
                    Entry e2 = make_entry(cw[l], tn[l]);
                    if (slot < CRT_SSTACK)
                        sst(slot, e2);
                    else if (slot < CRT_SSTACK + CRT_LSTACK)
                        lstack[slot - CRT_SSTACK] = e2;
                    else
                        overflow = true;
                }
            }
            sp = base + n_int + n_leaf;
        }
    };

    // ---- one leaf record (bvh.rs:514-572 / 617-653) -------------------------------
    auto record_test = [&]() {
        const float4* cur = rec;
        float4 q0 = __ldg(cur), q1 = __ldg(cur + 1), q2 = __ldg(cur + 2);
        const uint32_t flags = __float_as_uint(q0.w);
        const uint32_t kind = flags & REC_KIND_MASK;
        if (kind == REC_TRI) {
            rec = (flags & REC_LAST) ? nullptr : cur + REC_TRI_Q;
            if (ray_mask & __float_as_uint(q2.w)) {  // triangle.rs:257-271 per-lane visibility
                TriEval e = tri_eval(r, xyz(q0), xyz(q1), xyz(q2));
                if (e.on_edge) {
                    pkt_deferred |= 1u << ((flags >> REC_LANE_SHIFT) & 3u);  // f64 tie-break after the packet's other lanes
                } else if (!mixed_signs(e) && e.det != 0.0f) {
                    if constexpr (ANY_HIT) {
                        if (in_range_packet(e, t_min, closest)) {
                            any_found = true;
                            state = DONE;
                            rec = nullptr;
                            return;
                        }
                    } else {
                        if (in_range_packet(e, t_min, pkt_tmax)) {
                            float t = e.t_scaled * (1.0f / e.det);
                            if (!(t > closest) && !(flags & REC_DEGENERATE)) accept(cur, t);  // bvh.rs:542, prim.rs:80-83
                        }
                    }
                }
            }
            if (flags & REC_PACKET_END) {
                // Lanes sitting exactly on an edge: scalar f64 path (bvh.rs:551-561 / 636-643).
                while (pkt_deferred) {
                    int ln = __ffs(pkt_deferred) - 1;
                    pkt_deferred &= pkt_deferred - 1;
                    const float4* fr = pkt_first + REC_TRI_Q * ln;
                    float4 f0 = __ldg(fr), f1 = __ldg(fr + 1), f2 = __ldg(fr + 2);
                    TriEval e = tri_eval_f64(r, xyz(f0), xyz(f1), xyz(f2));
                    if (mixed_signs(e) || e.det == 0.0f) continue;
                    if (!in_range_scalar(e, t_min, closest)) continue;
                    if constexpr (ANY_HIT) {
                        any_found = true;
                        state = DONE;
                        rec = nullptr;
                        pkt_deferred = 0;
                        return;
                    } else {
                        if (__float_as_uint(f0.w) & REC_DEGENERATE) continue;
                        accept(fr, e.t_scaled * (1.0f / e.det));
                    }
                }
                pkt_tmax = closest;
                pkt_first = cur + REC_TRI_Q;
            }
            return;
        }
        if constexpr ((FEAT & (FEAT_SPHERE | FEAT_CURVE | FEAT_INSTANCE)) != 0) {
            bool got = false;
            float t = 0.0f;
            F3 n = f3(0, 0, 0);
            if (kind == REC_SPHERE) {
                rec = (flags & REC_LAST) ? nullptr : cur + REC_SPHERE_Q;
                if (ray_mask & __float_as_uint(q1.z)) got = sphere_hit(r, xyz(q0), q1.x, t_min, closest, t, n);
            } else if ((FEAT & FEAT_CURVE) && kind == REC_CURVE) {
                rec = (flags & REC_LAST) ? nullptr : cur + REC_CURVE_Q;
                if (ray_mask & __float_as_uint(q2.z))
                    got = rounded_cone_hit(r.o, r.d, xyz(q0), xyz(q1), q2.x, q2.y, t_min, closest, t, n);
            } else if ((FEAT & FEAT_CURVE) && kind == REC_CUBIC) {
                rec = (flags & REC_LAST) ? nullptr : cur + REC_CUBIC_Q;
                if (ray_mask & __float_as_uint(q2.w)) {
                    float4 q3 = __ldg(cur + 3), q4 = __ldg(cur + 4);
                    F3 cp[4] = {xyz(q0), xyz(q1), xyz(q2), xyz(q3)};
                    got = cubic_curve_hit(r.o, r.d, cp, q4.x, q4.y, t_min, closest, t, n);
                }
            } else if (INST && kind == REC_INSTANCE) {
                const float4* resume = (flags & REC_LAST) ? nullptr : cur + REC_INSTANCE_Q;
                rec = resume;
                if ((ray_mask & __float_as_uint(q2.w)) && depth < CRT_MAX_INSTANCE_DEPTH) {  // prim.rs:346-350
                    const DevScene& owner = cur_scene ? *cur_scene : p.scene;
                    Xf w2l = instance_w2l(owner, cur, time);
                    const DevScene* child = instance_child(cur);
                    LevelSave& s = saved[depth];
                    s.r = r;
                    s.scene = cur_scene;
                    s.resume = resume;
                    s.inst = cur;
                    ++depth;
                    cur_scene = child;
                    nodes = child->nodes;
                    leafs = (const float4*)child->leafs;
                    r.set(xf_point(w2l, r.o), xf_vec(w2l, r.d), child->has_triangles != 0);  // prim.rs:333-341
                    rec = nullptr;
                    push(make_entry(CRT_RETURN_MARK, 0.0f));
                    push(make_entry(0u, t_min));  // committed child scenes are never empty (scene.rs:307-309)
                }
                return;
            } else {
                rec = nullptr;  // unreachable: a record kind this variant was not built for
                return;
            }
            if (got) {
                if constexpr (ANY_HIT) {
                    any_found = true;
                    state = DONE;
                    rec = nullptr;
                } else {
                    accept(cur, t);
                    best_n = n;
                }
            }
        } else {
            rec = nullptr;  // unreachable: a record kind this variant was not built for
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
                        closest = b.w;
                        time = p.time ? __ldg(p.time + idx) : 0.0f;
                        ray_mask = p.mask ? __ldg(p.mask + idx) : p.mask_uniform;
                        world_o = f3(a.x, a.y, a.z);
                        world_d = f3(b.x, b.y, b.z);
                        r.set(world_o, world_d, p.scene.has_triangles != 0);
                        nodes = p.scene.nodes;
                        leafs = (const float4*)p.scene.leafs;
                        cur_scene = nullptr;
                        depth = 0;
                        best_rec = nullptr;
                        best_depth = 0;
                        any_found = false;
                        rec = nullptr;
                        sp = 0;
                        state = RUN;
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
            while (rec == nullptr) {
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
                        const LevelSave& s = saved[depth];
                        cur_scene = s.scene;
                        const DevScene& sc = cur_scene ? *cur_scene : p.scene;
                        nodes = sc.nodes;
                        leafs = (const float4*)sc.leafs;
                        r = s.r;
                        rec = s.resume;
                        continue;
                    }
                }
                if (child & CRT_LEAF_BIT) {
                    rec = leafs + (child & ~CRT_LEAF_BIT);
                    pkt_tmax = closest;
                    pkt_first = rec;
                    pkt_deferred = 0;
                    break;
                }
                if constexpr (!ANY_HIT) {
                    if (__uint_as_float(e.y) > closest) continue;  // see traverse.cuh: same result as visiting
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
            while (state == RUN && rec != nullptr) {
                record_test();
#if CRT_LEAF_EXIT_NUM > 0
                if (__popc(__activemask()) * CRT_LEAF_EXIT_DEN < n_run * CRT_LEAF_EXIT_NUM) break;
#endif
            }
        }
    }
    if (overflow) atomicOr(p.status, 1u);
}

template <bool ANY_HIT, uint32_t FEAT>
cudaError_t launch_variant(const TraceParams& p, int grid, cudaStream_t stream) {
    trace_kernel<ANY_HIT, FEAT><<<grid, CRT_BLOCK, 0, stream>>>(p);
    return cudaGetLastError();
}

template <bool ANY_HIT, uint32_t FEAT>
int blocks_per_sm_variant() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_kernel<ANY_HIT, FEAT>, CRT_BLOCK, 0);
    return n > 0 ? n : 1;
}

// Feature sets actually instantiated; a scene uses the smallest superset.
constexpr uint32_t V_TRI = 0;
constexpr uint32_t V_SPH = FEAT_SPHERE | FEAT_NORMALS;
constexpr uint32_t V_INST = FEAT_SPHERE | FEAT_NORMALS | FEAT_INSTANCE;
constexpr uint32_t V_ALL = FEAT_SPHERE | FEAT_NORMALS | FEAT_INSTANCE | FEAT_CURVE;

int variant_of(uint32_t feat) {
    if ((feat & ~V_TRI) == 0) return 0;
    if ((feat & ~V_SPH) == 0) return 1;
    if ((feat & ~V_INST) == 0) return 2;
    return 3;
}

}  // namespace

int trace_blocks_per_sm(bool any_hit, uint32_t feat) {
    static int cache[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    int v = variant_of(feat);
    int& c = cache[any_hit ? 1 : 0][v];
    if (c == 0) {
        if (any_hit)
            c = v == 0 ? blocks_per_sm_variant<true, V_TRI>()
                       : (v == 1 ? blocks_per_sm_variant<true, V_SPH>()
                                 : (v == 2 ? blocks_per_sm_variant<true, V_INST>() : blocks_per_sm_variant<true, V_ALL>()));
        else
            c = v == 0 ? blocks_per_sm_variant<false, V_TRI>()
                       : (v == 1 ? blocks_per_sm_variant<false, V_SPH>()
                                 : (v == 2 ? blocks_per_sm_variant<false, V_INST>() : blocks_per_sm_variant<false, V_ALL>()));
    }
    return c;
}

cudaError_t launch_trace(bool any_hit, uint32_t feat, const TraceParams& p, int grid, cudaStream_t stream) {
    int v = variant_of(feat);
    if (any_hit) {
        switch (v) {
            case 0: return launch_variant<true, V_TRI>(p, grid, stream);
            case 1: return launch_variant<true, V_SPH>(p, grid, stream);
            case 2: return launch_variant<true, V_INST>(p, grid, stream);
            default: return launch_variant<true, V_ALL>(p, grid, stream);
        }
    }
    switch (v) {
        case 0: return launch_variant<false, V_TRI>(p, grid, stream);
        case 1: return launch_variant<false, V_SPH>(p, grid, stream);
        case 2: return launch_variant<false, V_INST>(p, grid, stream);
        default: return launch_variant<false, V_ALL>(p, grid, stream);
    }
}

}  // namespace crt
