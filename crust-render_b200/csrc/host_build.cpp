// See host_build.h. Compiled with -ffp-contract=off: split decisions must be
// bit-reproducible (bvh.rs:22-24 "every split decision depends only on the input").
#include "host_build.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace crt {
namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();
// bvh.rs:142-167
constexpr size_t kMaxDepth = 60, kMinLeaf = 2, kMinLeafPacked = 4, kMaxLeaf = 8, kBins = 12, kParallelThreshold = 4096,
                 kSbvhMaxDepth = 32;
constexpr float kSbvhAlpha = 1e-5f;
constexpr float kPad = 1e-4f;  // aabb.rs:50, triangle.rs:421, prim.rs:311

inline Vec3 V(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 add(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 mul(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline float dot3(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }  // glam: (x + y) + z
inline Vec3 cross3(Vec3 a, Vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline Box unite(const Box& a, const Box& b) { return {vmin(a.lo, b.lo), vmax(a.hi, b.hi)}; }
inline float area(const Box& b) {  // bvh.rs:811-814
    Vec3 d = sub(b.hi, b.lo);
    return 2.0f * (d.x * d.y + d.y * d.z + d.z * d.x);
}
inline bool overlap_box(const Box& a, const Box& b, Box& out) {  // bvh.rs:823-831
    Vec3 lo = vmax(a.lo, b.lo), hi = vmin(a.hi, b.hi);
    if (lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z) {
        out = {lo, hi};
        return true;
    }
    return false;
}
inline void pad_flat_axes(Vec3& lo, Vec3& hi) {
    for (int a = 0; a < 3; ++a)
        if (hi[a] - lo[a] < kPad) {
            lo[a] -= kPad;
            hi[a] += kPad;
        }
}
inline int longest_axis(Vec3 e) { return (e.x >= e.y && e.x >= e.z) ? 0 : (e.y >= e.z ? 1 : 2); }
// Rust `as usize` on f32 saturates and maps NaN to 0 (bvh.rs:916, 988).
inline size_t to_bin(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= (float)kBins) return kBins - 1;
    return (size_t)f;
}

// glam Mat3A/Affine3A pieces used at commit (scene.rs:310, prim.rs:298-319).
inline Vec3 xform_vec(const Xform& m, Vec3 v) { return add(add(mul(m.x, v.x), mul(m.y, v.y)), mul(m.z, v.z)); }
inline Vec3 xform_point(const Xform& m, Vec3 p) { return add(xform_vec(m, p), m.t); }
Xform invert(const Xform& a) {
    Vec3 t0 = cross3(a.y, a.z), t1 = cross3(a.z, a.x), t2 = cross3(a.x, a.y);
    float inv_det = 1.0f / dot3(a.z, t2);
    Vec3 c0 = mul(t0, inv_det), c1 = mul(t1, inv_det), c2 = mul(t2, inv_det);
    Xform r;
    r.x = {c0.x, c1.x, c2.x};
    r.y = {c0.y, c1.y, c2.y};
    r.z = {c0.z, c1.z, c2.z};
    Vec3 mt = xform_vec(r, a.t);
    r.t = {-mt.x, -mt.y, -mt.z};
    return r;
}
Box transformed_box(const Box& local, const Xform& m) {
    Vec3 lo = V(kInf, kInf, kInf), hi = V(-kInf, -kInf, -kInf);
    for (int i = 0; i < 8; ++i) {
        Vec3 c{(i & 1) ? local.hi.x : local.lo.x, (i & 2) ? local.hi.y : local.lo.y, (i & 4) ? local.hi.z : local.lo.z};
        Vec3 p = xform_point(m, c);
        lo = vmin(lo, p);
        hi = vmax(hi, p);
    }
    pad_flat_axes(lo, hi);
    return {lo, hi};
}

struct Ref {  // bvh.rs:281-290 PrimRef
    Box b;
    uint32_t idx;
};
inline float centroid_axis(const Ref& r, int a) { return 0.5f * (r.b.lo[a] + r.b.hi[a]); }

// ---- primitive bounds -------------------------------------------------------
inline const float* tri_of(const PrimStore& ps, uint32_t pi) { return &ps.tri[9 * (size_t)ps.data[pi]]; }

Box prim_box(const PrimStore& ps, uint32_t pi) {
    switch (ps.kind[pi]) {
        case P_TRI: {  // aabb.rs:47-58
            const float* t = tri_of(ps, pi);
            Vec3 v0 = V(t[0], t[1], t[2]), v1 = V(t[3], t[4], t[5]), v2 = V(t[6], t[7], t[8]);
            Vec3 lo = vmin(vmin(v0, v1), v2), hi = vmax(vmax(v0, v1), v2);
            pad_flat_axes(lo, hi);
            return {lo, hi};
        }
        case P_SPHERE: {  // prim.rs:163-168
            const float* s = &ps.sphere[4 * (size_t)ps.data[pi]];
            return {V(s[0] - s[3], s[1] - s[3], s[2] - s[3]), V(s[0] + s[3], s[1] + s[3], s[2] + s[3])};
        }
        case P_CURVE: {  // prim.rs:202-206
            const float* c = &ps.curve[8 * (size_t)ps.data[pi]];
            Box a{V(c[0] - c[6], c[1] - c[6], c[2] - c[6]), V(c[0] + c[6], c[1] + c[6], c[2] + c[6])};
            Box b{V(c[3] - c[7], c[4] - c[7], c[5] - c[7]), V(c[3] + c[7], c[4] + c[7], c[5] + c[7])};
            return unite(a, b);
        }
        case P_CUBIC: {  // prim.rs:246-253
            const float* c = &ps.cubic[14 * (size_t)ps.data[pi]];
            float r = std::max(c[12], c[13]);
            Vec3 lo = V(c[0], c[1], c[2]), hi = lo;
            for (int k = 1; k < 4; ++k) {
                lo = vmin(lo, V(c[3 * k], c[3 * k + 1], c[3 * k + 2]));
                hi = vmax(hi, V(c[3 * k], c[3 * k + 1], c[3 * k + 2]));
            }
            return {V(lo.x - r, lo.y - r, lo.z - r), V(hi.x + r, hi.y + r, hi.z + r)};
        }
        default: return ps.inst[ps.data[pi]].bounds;  // prim.rs:380-382
    }
}

// triangle.rs:367-429 — Sutherland-Hodgman against the slab's two planes.
bool clip_triangle(const float* t, int axis, float mn, float mx, Box& out) {
    Vec3 poly[8], tmp[8];
    int n = 3;
    poly[0] = V(t[0], t[1], t[2]);
    poly[1] = V(t[3], t[4], t[5]);
    poly[2] = V(t[6], t[7], t[8]);
    for (int pass = 0; pass < 2; ++pass) {
        const float bound = pass == 0 ? mn : mx;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            Vec3 a = poly[i], b = poly[(i + 1) % n];
            float da = pass == 0 ? a[axis] - bound : bound - a[axis];
            float db = pass == 0 ? b[axis] - bound : bound - b[axis];
            if (da >= 0.0f) tmp[m++] = a;
            if ((da > 0.0f) != (db > 0.0f) && da != db) {
                float s = da / (da - db);
                tmp[m++] = add(a, mul(sub(b, a), s));
            }
        }
        n = m;
        if (n == 0) return false;
        for (int i = 0; i < n; ++i) poly[i] = tmp[i];
    }
    Vec3 lo = poly[0], hi = poly[0];
    for (int i = 1; i < n; ++i) {
        lo = vmin(lo, poly[i]);
        hi = vmax(hi, poly[i]);
    }
    lo[axis] = std::max(lo[axis], mn);
    hi[axis] = std::min(hi[axis], mx);
    pad_flat_axes(lo, hi);
    out = {lo, hi};
    return true;
}
// prim.rs:39-48 default + :116-118 triangle override, then ∩ ref box (bvh.rs:1013, 1116, 1122)
bool clipped_ref_box(const PrimStore& ps, const Ref& r, int axis, float mn, float mx, Box& out) {
    Box c;
    if (ps.kind[r.idx] == P_TRI) {
        if (!clip_triangle(tri_of(ps, r.idx), axis, mn, mx, c)) return false;
    } else {
        c = prim_box(ps, r.idx);
        if (c.lo[axis] > mx || c.hi[axis] < mn) return false;
        c.lo[axis] = std::max(c.lo[axis], mn);
        c.hi[axis] = std::min(c.hi[axis], mx);
    }
    return overlap_box(c, r.b, out);
}

// ---- binary build -----------------------------------------------------------
struct ObjSplit {
    int axis;
    float cmin, scale;
    size_t split_bin;
    float cost;
    Box left, right;
};
struct BinAcc {
    size_t n = 0;
    bool some = false;
    Box b;
    void grow(const Box& x) {
        b = some ? unite(b, x) : x;
        some = true;
    }
};

// Data-parallel passes over the references of one large node. [0, n) is cut into contiguous chunks, chunk k is processed by
// fn(k, lo, hi), and callers merge the per-chunk partial results IN CHUNK ORDER: min / max (ties keep the earlier operand, like
// the sequential fold), integer counts and order-preserving concatenation all come out exactly as one pass would produce
// them, so the tree does not depend on the number of threads (bvh.rs:1621-1638).
struct Par {
    int threads = 1;
    size_t min_refs = (size_t)1 << 19;
    int chunks_for(size_t n) const { return (threads > 1 && n >= min_refs) ? threads : 1; }
    template <class F>
    void run(size_t n, int nc, F&& fn) const {
        const size_t per = (n + (size_t)nc - 1) / (size_t)nc;
        auto lo = [&](int k) { return std::min(n, (size_t)k * per); };
        if (nc <= 1) return fn(0, (size_t)0, n);
        std::vector<std::thread> pool;
        for (int k = 1; k < nc; ++k) pool.emplace_back([&, k] { fn(k, lo(k), lo(k + 1)); });
        fn(0, lo(0), lo(1));
        for (auto& t : pool) t.join();
    }
};

// Sweep shared by both split kinds: candidate plane after bin `s` costs
// area(∪ bounds[0..=s]) * Σ lcount[0..=s] + area(∪ bounds[s+1..]) * Σ rcount[s+1..]
// (bvh.rs:931-967, 1020-1052); strict `<` keeps the first best.
template <class Emit>
void sweep(const BinAcc* bins, const size_t* lcount, const size_t* rcount, Emit&& emit) {
    Box rb[kBins];
    bool rsome[kBins];
    size_t rc[kBins];
    {
        BinAcc acc;
        size_t c = 0;
        for (size_t b = kBins; b-- > 1;) {
            if (bins[b].some) acc.grow(bins[b].b);
            c += rcount[b];
            rb[b] = acc.b;
            rsome[b] = acc.some;
            rc[b] = c;
        }
    }
    BinAcc lacc;
    size_t lc = 0;
    for (size_t s = 0; s + 1 < kBins; ++s) {
        if (bins[s].some) lacc.grow(bins[s].b);
        lc += lcount[s];
        if (lc == 0 || rc[s + 1] == 0) continue;
        if (!lacc.some || !rsome[s + 1]) continue;  // unreachable on finite input (reference would panic)
        float cost = area(lacc.b) * (float)lc + area(rb[s + 1]) * (float)rc[s + 1];
        emit(s, cost, lacc.b, rb[s + 1]);
    }
}

bool object_split(const std::vector<Ref>& refs, ObjSplit& out, const Par& par) {  // bvh.rs:894-969
    const int nc = par.chunks_for(refs.size());
    Vec3 cmin, cmax;
    {
        std::vector<Vec3> pmin(nc), pmax(nc);
        par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) {
            if (lo >= hi) return;
            Vec3 mn, mx;
            for (int a = 0; a < 3; ++a) mn[a] = mx[a] = centroid_axis(refs[lo], a);
            for (size_t i = lo + 1; i < hi; ++i)
                for (int a = 0; a < 3; ++a) {
                    float c = centroid_axis(refs[i], a);
                    mn[a] = std::min(mn[a], c);
                    mx[a] = std::max(mx[a], c);
                }
            pmin[k] = mn, pmax[k] = mx;
        });
        cmin = pmin[0], cmax = pmax[0];
        const size_t per = (refs.size() + (size_t)nc - 1) / (size_t)nc;
        for (int k = 1; k < nc; ++k) {
            if ((size_t)k * per >= refs.size()) break;  // empty trailing chunk
            for (int a = 0; a < 3; ++a) {
                cmin[a] = std::min(cmin[a], pmin[k][a]);
                cmax[a] = std::max(cmax[a], pmax[k][a]);
            }
        }
    }
    Vec3 extent = sub(cmax, cmin);
    int axis = longest_axis(extent);
    if (extent[axis] <= 1e-6f) return false;
    float scale = (float)kBins / extent[axis];
    float c0 = cmin[axis];
    BinAcc bins[kBins];
    size_t counts[kBins] = {};
    {
        struct Part {
            BinAcc bins[kBins];
            size_t counts[kBins] = {};
        };
        std::vector<Part> part(nc);
        par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) {
            Part& p = part[k];
            for (size_t i = lo; i < hi; ++i) {
                const Ref& r = refs[i];
                size_t b = to_bin((centroid_axis(r, axis) - c0) * scale);
                p.counts[b] += 1;
                p.bins[b].grow(r.b);
            }
        });
        for (int k = 0; k < nc; ++k)
            for (size_t b = 0; b < kBins; ++b) {
                counts[b] += part[k].counts[b];
                if (part[k].bins[b].some) bins[b].grow(part[k].bins[b].b);
            }
    }
    bool found = false;
    sweep(bins, counts, counts, [&](size_t s, float cost, const Box& lb, const Box& rb) {
        if (!found || cost < out.cost) {
            out = ObjSplit{axis, c0, scale, s, cost, lb, rb};
            found = true;
        }
    });
    return found;
}

struct SpatSplit {
    int axis;
    float pos, cost;
};
bool spatial_split(const PrimStore& ps, const std::vector<Ref>& refs, const Box& bbox, SpatSplit& out, const Par& par) {  // bvh.rs:974-1054
    Vec3 extent = sub(bbox.hi, bbox.lo);
    int axis = longest_axis(extent);
    if (extent[axis] <= 1e-6f) return false;
    float lo = bbox.lo[axis];
    float width = extent[axis] / (float)kBins;
    BinAcc bins[kBins];
    size_t entry[kBins] = {}, exit_[kBins] = {};
    {
        struct Part {
            BinAcc bins[kBins];
            size_t entry[kBins] = {}, exit_[kBins] = {};
        };
        const int nc = par.chunks_for(refs.size());
        std::vector<Part> part(nc);
        par.run(refs.size(), nc, [&](int k, size_t from, size_t to) {
            Part& p = part[k];
            for (size_t i = from; i < to; ++i) {
                const Ref& r = refs[i];
                size_t b0 = to_bin((r.b.lo[axis] - lo) / width), b1 = to_bin((r.b.hi[axis] - lo) / width);
                p.entry[b0] += 1;
                p.exit_[b1] += 1;
                if (b0 == b1) {
                    p.bins[b0].grow(r.b);
                    continue;
                }
                for (size_t b = b0; b <= b1; ++b) {
                    Box c;
                    if (clipped_ref_box(ps, r, axis, lo + (float)b * width, lo + (float)(b + 1) * width, c)) p.bins[b].grow(c);
                }
            }
        });
        for (int k = 0; k < nc; ++k)
            for (size_t b = 0; b < kBins; ++b) {
                entry[b] += part[k].entry[b];
                exit_[b] += part[k].exit_[b];
                if (part[k].bins[b].some) bins[b].grow(part[k].bins[b].b);
            }
    }
    bool found = false;
    sweep(bins, entry, exit_, [&](size_t s, float cost, const Box&, const Box&) {
        if (!found || cost < out.cost) {
            out = SpatSplit{axis, lo + (float)(s + 1) * width, cost};
            found = true;
        }
    });
    return found;
}

struct Out {  // one DFS-ordered piece of the binary tree
    std::vector<BinNode> nodes;
    std::vector<uint32_t> indices;
};
struct Builder {
    const PrimStore& ps;
    float root_area;
    std::atomic<int> spare_threads;
    Par all;  // every thread of the build; a node at depth d uses threads >> d of them for its own passes

    size_t min_leaf_for(const std::vector<Ref>& refs, const Par& par) const {  // bvh.rs:1201-1210
        const int nc = par.chunks_for(refs.size());
        if (nc > 1) {
            std::vector<char> other(nc, 0);
            par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i)
                    if (ps.kind[refs[i].idx] != P_TRI) {
                        other[k] = 1;
                        break;
                    }
            });
            for (char o : other)
                if (o) return kMinLeaf;
            return kMinLeafPacked;
        }
        for (const Ref& r : refs)
            if (ps.kind[r.idx] != P_TRI) return kMinLeaf;
        return kMinLeafPacked;
    }
    static void emit_leaf(Out& out, const Box& bbox, const std::vector<Ref>& refs) {  // bvh.rs:833-842
        out.nodes.push_back(BinNode{bbox, (uint32_t)out.indices.size(), (uint32_t)refs.size()});
        for (const Ref& r : refs) out.indices.push_back(r.idx);
    }
    void stable_split(std::vector<Ref>& refs, const ObjSplit& o, std::vector<Ref>& l, std::vector<Ref>& r, const Par& par) const {  // bvh.rs:1224-1237
        auto left = [&](const Ref& x) { return to_bin((centroid_axis(x, o.axis) - o.cmin) * o.scale) <= o.split_bin; };
        const int nc = par.chunks_for(refs.size());
        if (nc > 1) {  // per-chunk counts, offsets by prefix, every chunk writes its own slices: order preserved
            std::vector<size_t> nl(nc + 1, 0), nr(nc + 1, 0);
            par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) {
                size_t c = 0;
                for (size_t i = lo; i < hi; ++i) c += left(refs[i]);
                nl[k + 1] = c;
                nr[k + 1] = (hi - lo) - c;
            });
            for (int k = 0; k < nc; ++k) nl[k + 1] += nl[k], nr[k + 1] += nr[k];
            l.resize(nl[nc]);
            r.resize(nr[nc]);
            par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) {
                size_t a = nl[k], b = nr[k];
                for (size_t i = lo; i < hi; ++i) {
                    if (left(refs[i])) l[a++] = refs[i]; else r[b++] = refs[i];
                }
            });
            std::vector<Ref>().swap(refs);
            return;
        }
        size_t nl = 0;
        for (const Ref& x : refs) nl += left(x);
        l.reserve(nl);
        r.reserve(refs.size() - nl);
        for (const Ref& x : refs) (left(x) ? l : r).push_back(x);
        std::vector<Ref>().swap(refs);
    }
    // Appends `sub` (built standalone) behind `out`, rebasing like bvh.rs:846-872.
    static void splice(Out& out, Out&& sub) {
        uint32_t node_off = (uint32_t)out.nodes.size(), idx_off = (uint32_t)out.indices.size();
        out.nodes.reserve(out.nodes.size() + sub.nodes.size());
        for (BinNode n : sub.nodes) {
            n.first_or_right += n.count == 0 ? node_off : idx_off;
            out.nodes.push_back(n);
        }
        out.indices.insert(out.indices.end(), sub.indices.begin(), sub.indices.end());
    }

    void children(Out& out, size_t self, std::vector<Ref>& l, std::vector<Ref>& r, size_t depth) {
        bool big = std::max(l.size(), r.size()) > kParallelThreshold;  // bvh.rs:1156
        if (big && spare_threads.fetch_sub(1) > 0) {
            auto fut = std::async(std::launch::async, [this, depth, rr = std::move(r)]() mutable {
                Out sub;
                build(sub, rr, depth + 1);
                spare_threads.fetch_add(1);
                return sub;
            });
            build(out, l, depth + 1);
            Out sub = fut.get();
            out.nodes[self].first_or_right = (uint32_t)out.nodes.size();
            splice(out, std::move(sub));
            return;
        }
        if (big) spare_threads.fetch_add(1);
        build(out, l, depth + 1);
        out.nodes[self].first_or_right = (uint32_t)out.nodes.size();
        build(out, r, depth + 1);
    }

    // bvh.rs:1059-1169 — emits the subtree of `refs` in depth-first order.
    void build(Out& out, std::vector<Ref>& refs, size_t depth) {
        // up to 2^depth sibling subtrees are being built at the same time: each large node takes its share of the threads
        const Par par{std::max(1, all.threads >> (int)std::min<size_t>(depth, 20)), all.min_refs};
        Box bbox = refs[0].b;
        {
            const int nc = par.chunks_for(refs.size());
            if (nc > 1) {
                std::vector<Box> part(nc);
                std::vector<char> some(nc, 0);
                par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) {
                    if (lo >= hi) return;
                    Box b = refs[lo].b;
                    for (size_t i = lo + 1; i < hi; ++i) b = unite(b, refs[i].b);
                    part[k] = b, some[k] = 1;
                });
                bbox = part[0];
                for (int k = 1; k < nc; ++k)
                    if (some[k]) bbox = unite(bbox, part[k]);
            } else {
                for (size_t i = 1; i < refs.size(); ++i) bbox = unite(bbox, refs[i].b);
            }
        }
        const size_t count = refs.size();
        // (the kind scan behind min_leaf_for only matters for a node that small: it is a pass of scattered loads otherwise)
        if ((count <= std::max(kMinLeaf, kMinLeafPacked) && count <= min_leaf_for(refs, par)) || depth >= kMaxDepth) return emit_leaf(out, bbox, refs);

        ObjSplit obj{};
        const bool have_obj = object_split(refs, obj, par);
        SpatSplit spat{};
        bool have_spat = false;
        if (have_obj && depth < kSbvhMaxDepth) {
            Box ov;
            float overlap = overlap_box(obj.left, obj.right, ov) ? area(ov) : 0.0f;
            if (overlap / root_area > kSbvhAlpha) have_spat = spatial_split(ps, refs, bbox, spat, par) && spat.cost < obj.cost;
        }

        std::vector<Ref> l, r;
        if (have_spat) {  // bvh.rs:1089-1135
            auto chop = [&](size_t lo, size_t hi, std::vector<Ref>& ll, std::vector<Ref>& rr) {
                for (size_t i = lo; i < hi; ++i) {
                    const Ref& x = refs[i];
                    if (x.b.hi[spat.axis] <= spat.pos)
                        ll.push_back(x);
                    else if (x.b.lo[spat.axis] >= spat.pos)
                        rr.push_back(x);
                    else {
                        Box c;
                        if (clipped_ref_box(ps, x, spat.axis, -kInf, spat.pos, c)) ll.push_back(Ref{c, x.idx});
                        if (clipped_ref_box(ps, x, spat.axis, spat.pos, kInf, c)) rr.push_back(Ref{c, x.idx});
                    }
                }
            };
            const int nc = par.chunks_for(refs.size());
            if (nc > 1) {  // every chunk chops into its own pair of lists; concatenated in chunk order
                std::vector<std::vector<Ref>> pl(nc), pr(nc);
                par.run(refs.size(), nc, [&](int k, size_t lo, size_t hi) { chop(lo, hi, pl[k], pr[k]); });
                size_t tl = 0, tr = 0;
                for (int k = 0; k < nc; ++k) tl += pl[k].size(), tr += pr[k].size();
                l.reserve(tl), r.reserve(tr);
                for (int k = 0; k < nc; ++k) {
                    l.insert(l.end(), pl[k].begin(), pl[k].end());
                    r.insert(r.end(), pr[k].begin(), pr[k].end());
                }
            } else {
                chop(0, refs.size(), l, r);
            }
            std::vector<Ref>().swap(refs);
            if (l.empty() || r.empty()) {  // degenerate chop: bvh.rs:1130-1134, 1173-1196
                l.insert(l.end(), r.begin(), r.end());
                std::vector<Ref>().swap(r);
                if (have_obj && (l.size() > std::max(kMinLeaf, kMinLeafPacked) || l.size() > min_leaf_for(l, par))) {
                    std::vector<Ref> a, b;
                    stable_split(l, obj, a, b, par);
                    if (a.empty() || b.empty()) {
                        a.insert(a.end(), b.begin(), b.end());
                        return emit_leaf(out, bbox, a);
                    }
                    size_t self = out.nodes.size();
                    out.nodes.push_back(BinNode{bbox, 0, 0});
                    build(out, a, depth + 1);  // this path never forks (bvh.rs:1190-1191)
                    out.nodes[self].first_or_right = (uint32_t)out.nodes.size();
                    build(out, b, depth + 1);
                    return;
                }
                return emit_leaf(out, bbox, l);
            }
        } else if (have_obj) {
            if (count <= kMaxLeaf && obj.cost >= area(bbox) * (float)count) return emit_leaf(out, bbox, refs);  // bvh.rs:1140
            stable_split(refs, obj, l, r, par);
        } else {  // every centroid coincides: median by input order (bvh.rs:1148-1152)
            size_t mid = count / 2;
            r.assign(refs.begin() + mid, refs.end());
            refs.resize(mid);
            l.swap(refs);
        }
        size_t self = out.nodes.size();
        out.nodes.push_back(BinNode{bbox, 0, 0});
        children(out, self, l, r, depth);
    }
};

// The leaf record of a non-triangle primitive (crt_layout.h), appended to t.leafs; instances also append their motion pair
// and note where their record sits (child pointers are patched in at upload). `last`: REC_LAST or 0.
// Where a leaf record goes: appended to the tree's vectors (PushSink: device build's host-side preparation), or written at
// known positions of arrays sized beforehand (CursorSink: the flattener, whose subtrees are emitted in parallel).
struct PushSink {
    HostTree& t;
    void q(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        t.leafs.push_back(a);
        t.leafs.push_back(b);
        t.leafs.push_back(c);
        t.leafs.push_back(d);
    }
    uint32_t add_motion(const MotionXform& m) {
        t.motion.push_back(m);
        return (uint32_t)(t.motion.size() - 1);
    }
    void add_instance() { t.instance_record_q.push_back((uint32_t)(t.leafs.size() / 4)); }
};
struct Cur {  // positions in (or, in a dry run, sizes of) the flattened arrays
    size_t node = 0, quad = 0, motion = 0, inst = 0, leaves = 0, refs = 0;
};
template <bool DRY>
struct CursorSink {
    HostTree& t;
    Cur& c;
    void q(uint32_t a, uint32_t b, uint32_t c4, uint32_t d) {
        if (!DRY) {
            uint32_t* w = &t.leafs[4 * c.quad];
            w[0] = a, w[1] = b, w[2] = c4, w[3] = d;
        }
        ++c.quad;
    }
    uint32_t add_motion(const MotionXform& m) {
        if (!DRY) t.motion[c.motion] = m;
        return (uint32_t)c.motion++;
    }
    void add_instance() {
        if (!DRY) t.instance_record_q[c.inst] = (uint32_t)c.quad;
        ++c.inst;
    }
};
inline uint32_t fbits_of(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
template <class Sink>
void emit_other_record_to(const PrimStore& ps, uint32_t pi, uint32_t last, Sink& k) {
    auto qf = [&](float a, float b, float c, uint32_t d) { k.q(fbits_of(a), fbits_of(b), fbits_of(c), d); };
    switch (ps.kind[pi]) {
        case P_SPHERE: {
            const float* s = &ps.sphere[4 * (size_t)ps.data[pi]];
            qf(s[0], s[1], s[2], REC_SPHERE | last);
            k.q(fbits_of(s[3]), pi, ps.mask[pi], ps.geom_id[pi]);
            break;
        }
        case P_CURVE: {
            const float* c = &ps.curve[8 * (size_t)ps.data[pi]];
            qf(c[0], c[1], c[2], REC_CURVE | last);
            qf(c[3], c[4], c[5], pi);
            k.q(fbits_of(c[6]), fbits_of(c[7]), ps.mask[pi], 0);
            break;
        }
        case P_CUBIC: {
            const float* c = &ps.cubic[14 * (size_t)ps.data[pi]];
            qf(c[0], c[1], c[2], REC_CUBIC | last);
            qf(c[3], c[4], c[5], pi);
            qf(c[6], c[7], c[8], ps.mask[pi]);
            qf(c[9], c[10], c[11], 0);
            k.q(fbits_of(c[12]), fbits_of(c[13]), 0, 0);
            k.q(0, 0, 0, 0);
            break;
        }
        default: {
            const InstancePrim& I = ps.inst[ps.data[pi]];
            uint32_t flags = REC_INSTANCE | last;
            uint32_t motion = 0;
            if (I.has_end) {
                flags |= REC_MOTION;
                MotionXform m;
                const Vec3* a = &I.l2w.x;
                const Vec3* b = &I.l2w_end.x;
                for (int n = 0; n < 4; ++n) {
                    m.l2w[3 * n] = a[n].x, m.l2w[3 * n + 1] = a[n].y, m.l2w[3 * n + 2] = a[n].z;
                    m.l2w_end[3 * n] = b[n].x, m.l2w_end[3 * n + 1] = b[n].y, m.l2w_end[3 * n + 2] = b[n].z;
                }
                motion = k.add_motion(m);
            }
            k.add_instance();
            qf(I.w2l.x.x, I.w2l.x.y, I.w2l.x.z, flags);
            qf(I.w2l.y.x, I.w2l.y.y, I.w2l.y.z, pi);
            qf(I.w2l.z.x, I.w2l.z.y, I.w2l.z.z, ps.mask[pi]);
            qf(I.w2l.t.x, I.w2l.t.y, I.w2l.t.z, motion);
            k.q(0, 0, ps.geom_id[pi], ps.data[pi]);  // words 0-1: child descriptor pointer, patched at upload
            k.q(0, 0, 0, 0);                         // child node / leaf-record arrays, patched at upload
            break;
        }
    }
}

void emit_other_record(const PrimStore& ps, uint32_t pi, uint32_t last, HostTree& t) {
    PushSink k{t};
    emit_other_record_to(ps, pi, last, k);
}

// Runs fn(k) for k in [0, count) on up to `threads` workers (static interleave). Every k writes its own slice of the
// output arrays, so the result does not depend on the thread count.
template <class F>
void parallel_for_n(int threads, size_t count, F&& fn) {
    const int nt = (int)std::min<size_t>((size_t)threads, count);
    if (nt <= 1) {
        for (size_t k = 0; k < count; ++k) fn(k);
        return;
    }
    std::vector<std::thread> pool;
    for (int w = 1; w < nt; ++w)
        pool.emplace_back([&, w] {
            for (size_t k = (size_t)w; k < count; k += (size_t)nt) fn(k);
        });
    for (size_t k = 0; k < count; k += (size_t)nt) fn(k);
    for (auto& th : pool) th.join();
}

// ---- collapse to BVH4 + record emission --------------------------------------
// The flattened arrays are in the reference's depth-first order (bvh.rs:1299-1375): a node takes its slot when it is entered,
// a leaf's records are appended when its lane is reached. Every function below runs either as a dry run, which only advances
// the cursors (= measures a subtree), or for real, writing at the cursors into arrays sized by a dry run — so subtrees can
// be measured and then written in parallel, each at its own offsets, and the arrays do not depend on the thread count.
struct Flattener {
    const PrimStore& ps;
    const std::vector<BinNode>& bin;
    const std::vector<uint32_t>& bidx;
    HostTree& t;
    int threads;

    // bvh.rs:1256-1290: triangles of the range first (packet order), then the rest.
    template <bool DRY>
    uint32_t emit_leaf(const uint32_t* range, size_t n, Cur& c) {
        const uint32_t first_q = (uint32_t)c.quad;
        CursorSink<DRY> k{t, c};
        size_t n_tri = 0;
        for (size_t i = 0; i < n; ++i) n_tri += ps.kind[range[i]] == P_TRI;
        const size_t n_other = n - n_tri;
        if (DRY) {
            c.quad += (size_t)REC_TRI_Q * n_tri;
        } else {
            size_t ti = 0;
            for (size_t i = 0; i < n; ++i) {
                uint32_t pi = range[i];
                if (ps.kind[pi] != P_TRI) continue;
                const float* v = tri_of(ps, pi);
                uint32_t flags = REC_TRI | ((uint32_t)(ti & 3) << REC_LANE_SHIFT);
                if ((ti & 3) == 3 || ti + 1 == n_tri) flags |= REC_PACKET_END;
                if (ti + 1 == n_tri && n_other == 0) flags |= REC_LAST;
                if (ps.tri_normal_slot[ps.data[pi]] != 0xFFFFFFFFu) {
                    flags |= REC_HAS_NORMALS;
                } else {  // prim.rs:80-83
                    Vec3 a = V(v[0], v[1], v[2]);
                    Vec3 nrm = cross3(sub(V(v[3], v[4], v[5]), a), sub(V(v[6], v[7], v[8]), a));
                    if (nrm.x == 0.0f && nrm.y == 0.0f && nrm.z == 0.0f) flags |= REC_DEGENERATE;
                }
                k.q(fbits_of(v[0]), fbits_of(v[1]), fbits_of(v[2]), flags);
                k.q(fbits_of(v[3]), fbits_of(v[4]), fbits_of(v[5]), pi);
                k.q(fbits_of(v[6]), fbits_of(v[7]), fbits_of(v[8]), ps.mask[pi]);
                ++ti;
            }
        }
        size_t oi = 0;
        for (size_t i = 0; i < n && oi < n_other; ++i) {
            uint32_t pi = range[i];
            if (ps.kind[pi] == P_TRI) continue;
            const uint32_t last = (++oi == n_other) ? (uint32_t)REC_LAST : 0u;
            emit_other_record_to(ps, pi, last, k);
        }
        c.leaves += 1;
        c.refs += n;
        return CRT_LEAF_BIT | first_q;
    }

    static void set_lane(DevNode& w, int lane, const Box& b) {  // bvh.rs:243-251
        w.bmin_x[lane] = b.lo.x, w.bmin_y[lane] = b.lo.y, w.bmin_z[lane] = b.lo.z;
        w.bmax_x[lane] = b.hi.x, w.bmax_y[lane] = b.hi.y, w.bmax_z[lane] = b.hi.z;
        w.flags |= 1u << lane;
    }
    static DevNode empty_node() {  // bvh.rs:230-241
        DevNode w;
        for (int l = 0; l < 4; ++l) {
            w.bmin_x[l] = w.bmin_y[l] = w.bmin_z[l] = w.bmax_x[l] = w.bmax_y[l] = w.bmax_z[l] = kInf;
            w.child[l] = CRT_EMPTY_LANE;
        }
        w.flags = 0;
        w.pad[0] = w.pad[1] = w.pad[2] = 0;
        return w;
    }
    // The up-to-four lanes of the BVH4 node rooted at binary node b_idx: the internal child with the largest surface area is
    // expanded first, the first one winning ties (bvh.rs:1328-1356).
    int lanes_of(uint32_t b_idx, uint32_t kids[4]) const {
        kids[0] = b_idx + 1, kids[1] = bin[b_idx].first_or_right, kids[2] = kids[3] = 0;
        int n_kids = 2;
        while (n_kids < 4) {
            int best = -1;
            float best_a = 0;
            for (int i = 0; i < n_kids; ++i)
                if (bin[kids[i]].count == 0) {
                    float a = area(bin[kids[i]].bbox);
                    if (best < 0 || a > best_a) best = i, best_a = a;
                }
            if (best < 0) break;
            uint32_t k = kids[best];
            kids[best] = k + 1;
            kids[n_kids++] = bin[k].first_or_right;
        }
        return n_kids;
    }
    // A whole subtree, depth first (bvh.rs:1328-1375).
    template <bool DRY>
    uint32_t node(uint32_t b_idx, Cur& c) {
        const size_t slot = c.node++;
        uint32_t kids[4];
        const int n_kids = lanes_of(b_idx, kids);
        DevNode w;
        if (!DRY) w = empty_node();
        for (int lane = 0; lane < n_kids; ++lane) {
            const BinNode& k = bin[kids[lane]];
            uint32_t ch;
            if (k.count > 0) {
                ch = emit_leaf<DRY>(bidx.data() + k.first_or_right, k.count, c);
                if (!DRY) w.flags |= 1u << (4 + lane);
            } else {
                ch = node<DRY>(kids[lane], c);
            }
            if (!DRY) set_lane(w, lane, k.bbox), w.child[lane] = ch;
        }
        if (!DRY) t.nodes[slot] = w;
        return (uint32_t)slot;
    }

    // Subtrees handed to the workers: the internal lanes `kTaskDepth` BVH4 levels below the root, in depth-first order
    // (trees of 2^16 binary nodes or more; CRT_FLATTEN_TASK_DEPTH / CRT_FLATTEN_PAR_MIN move both for the tests).
    int kTaskDepth = getenv("CRT_FLATTEN_TASK_DEPTH") ? std::max(1, atoi(getenv("CRT_FLATTEN_TASK_DEPTH"))) : 5;
    size_t kParMin = getenv("CRT_FLATTEN_PAR_MIN") ? (size_t)atoll(getenv("CRT_FLATTEN_PAR_MIN")) : ((size_t)1 << 16);
    struct Task {
        uint32_t b_idx;
        Cur size, base;
    };
    void collect(uint32_t b_idx, int depth, std::vector<Task>& tasks) const {
        uint32_t kids[4];
        const int n_kids = lanes_of(b_idx, kids);
        for (int lane = 0; lane < n_kids; ++lane) {
            if (bin[kids[lane]].count > 0) continue;
            if (depth + 1 == kTaskDepth) tasks.push_back(Task{kids[lane], Cur{}, Cur{}});
            else collect(kids[lane], depth + 1, tasks);
        }
    }
    // The levels above the tasks: same walk as node(), except that a task's subtree is stepped over by its measured size
    // (dry run: to size the arrays; real run: the task's offsets are noted on the way).
    template <bool DRY>
    uint32_t top(uint32_t b_idx, int depth, Cur& c, std::vector<Task>& tasks, size_t& next_task) {
        const size_t slot = c.node++;
        uint32_t kids[4];
        const int n_kids = lanes_of(b_idx, kids);
        DevNode w;
        if (!DRY) w = empty_node();
        for (int lane = 0; lane < n_kids; ++lane) {
            const BinNode& k = bin[kids[lane]];
            uint32_t ch;
            if (k.count > 0) {
                ch = emit_leaf<DRY>(bidx.data() + k.first_or_right, k.count, c);
                if (!DRY) w.flags |= 1u << (4 + lane);
            } else if (depth + 1 == kTaskDepth) {
                Task& tk = tasks[next_task++];
                ch = (uint32_t)c.node;
                if (!DRY) tk.base = c;
                c.node += tk.size.node, c.quad += tk.size.quad, c.motion += tk.size.motion, c.inst += tk.size.inst;
                c.leaves += tk.size.leaves, c.refs += tk.size.refs;
            } else {
                ch = top<DRY>(kids[lane], depth + 1, c, tasks, next_task);
            }
            if (!DRY) set_lane(w, lane, k.bbox), w.child[lane] = ch;
        }
        if (!DRY) t.nodes[slot] = w;
        return (uint32_t)slot;
    }

    void size_arrays(const Cur& total) {
        t.nodes.resize(total.node);
        t.leafs.resize(4 * (total.quad + 4));  // + the over-read padding below
        t.motion.resize(total.motion);
        t.instance_record_q.resize(total.inst);
        t.n_leaves += total.leaves;
        t.n_leaf_refs += total.refs;
    }
    void run() {  // bvh.rs:1299-1320
        Cur total;
        if (bin[0].count > 0) {
            emit_leaf<true>(bidx.data() + bin[0].first_or_right, bin[0].count, total);
            total.node = 1;
            size_arrays(total);
            Cur c;
            c.node = 1;
            DevNode w = empty_node();
            set_lane(w, 0, bin[0].bbox);
            w.child[0] = emit_leaf<false>(bidx.data() + bin[0].first_or_right, bin[0].count, c);
            w.flags |= 1u << 4;
            t.nodes[0] = w;
        } else {
            std::vector<Task> tasks;
            if (threads > 1 && bin.size() >= kParMin) collect(0, 0, tasks);
            if (tasks.empty()) {
                node<true>(0, total);
                size_arrays(total);
                Cur c;
                node<false>(0, c);
            } else {
                parallel_for_n(threads, tasks.size(), [&](size_t i) { node<true>(tasks[i].b_idx, tasks[i].size); });
                size_t next = 0;
                top<true>(0, 0, total, tasks, next);
                size_arrays(total);
                Cur c;
                next = 0;
                top<false>(0, 0, c, tasks, next);
                parallel_for_n(threads, tasks.size(), [&](size_t i) {
                    Cur at = tasks[i].base;
                    node<false>(tasks[i].b_idx, at);
                });
            }
        }
        // Kernels may over-read up to one triangle record past a leaf's end.
        uint32_t* pad = &t.leafs[t.leafs.size() - 16];
        for (int i = 0; i < 4; ++i) pad[4 * i] = pad[4 * i + 1] = pad[4 * i + 2] = 0, pad[4 * i + 3] = REC_LAST | REC_SPHERE;
    }
};

inline Vec3 v3at(const std::vector<float>& a, size_t i) { return V(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }

}  // namespace

namespace {
// scene.rs:240-335: geometries -> primitives, ids, normals. `skip_meshes`: triangle meshes contribute nothing and keep their
// inputs (the device build expands them itself and only needs the other kinds from here); ids are unaffected.
bool expand_geoms(std::vector<Geom>& geoms, ChildInfoFn child_info, int threads, HostTree& out, const char** err, bool skip_meshes,
                  std::chrono::steady_clock::time_point t0) {
    PrimStore& ps = out.prims;
    auto parallel_for = [&](size_t count, auto&& fn) { parallel_for_n(threads, count, fn); };

    // scene.rs:240-335, in two passes so that meshes expand in parallel: (1) how many primitives / per-kind records each
    // geometry contributes (a triangle with an out-of-range index is skipped but still consumes its prim_id, :251-253),
    // (2) every geometry fills its own slices. Slot numbers are exactly those of a sequential expansion.
    const size_t ng = geoms.size();
    struct Slice {
        size_t prims = 0, tris = 0, tri_normals = 0, spheres = 0, curves = 0, cubics = 0, insts = 0;
    };
    std::vector<Slice> cnt(ng), off(ng);
    std::vector<ChildInfo> cinfo(ng);
    for (size_t gi = 0; gi < ng; ++gi) {  // instances first (sequential: they may fail, and they are few)
        const Geom& g = geoms[gi];
        if (g.type != G_INSTANCE) continue;
        ChildInfo ci = child_info(g.child);
        cinfo[gi] = ci;
        if (!ci.non_empty) continue;  // scene.rs:307-309
        if (ci.instance_depth + 1 > CRT_MAX_INSTANCE_DEPTH) {
            *err = "instance nesting deeper than CRT_MAX_INSTANCE_DEPTH";
            return false;
        }
        cnt[gi].prims = cnt[gi].insts = 1;
    }
    parallel_for(ng, [&](size_t gi) {
        const Geom& g = geoms[gi];
        Slice& c = cnt[gi];
        switch (g.type) {
            case G_MESH: {
                if (skip_meshes) break;
                const size_t nv = g.verts.size() / 3, nn = g.normals.size() / 3, nt = g.indices.size() / 3;
                for (size_t k = 0; k < nt; ++k) {
                    const size_t i0 = g.indices[3 * k], i1 = g.indices[3 * k + 1], i2 = g.indices[3 * k + 2];
                    if (i0 >= nv || i1 >= nv || i2 >= nv) continue;
                    ++c.tris;
                    if (g.has_normals && i0 < nn && i1 < nn && i2 < nn) ++c.tri_normals;
                }
                c.prims = c.tris;
                break;
            }
            case G_SPHERE: c.prims = c.spheres = 1; break;
            case G_CURVES: c.prims = c.curves = g.segs.size() / 8; break;
            case G_CUBICS: c.prims = c.cubics = g.segs.size() / 14; break;
            default: break;  // instances were counted above
        }
    });
    Slice total;
    for (size_t gi = 0; gi < ng; ++gi) {
        off[gi] = total;
        total.prims += cnt[gi].prims, total.tris += cnt[gi].tris, total.tri_normals += cnt[gi].tri_normals, total.spheres += cnt[gi].spheres;
        total.curves += cnt[gi].curves, total.cubics += cnt[gi].cubics, total.insts += cnt[gi].insts;
    }
    ps.kind.resize(total.prims), ps.geom_id.resize(total.prims), ps.prim_id.resize(total.prims), ps.mask.resize(total.prims), ps.data.resize(total.prims);
    ps.tri.resize(9 * total.tris), ps.tri_normal_slot.resize(total.tris), ps.tri_normals.resize(9 * total.tri_normals);
    ps.sphere.resize(4 * total.spheres), ps.curve.resize(8 * total.curves), ps.cubic.resize(14 * total.cubics), ps.inst.resize(total.insts);
    ps.any_normals = total.tri_normals != 0;
    parallel_for(ng, [&](size_t gi) {
        Geom& g = geoms[gi];
        const uint32_t gid = (uint32_t)gi;
        const Slice& o = off[gi];
        size_t pi = o.prims;
        auto push = [&](PrimKind k, uint32_t pid, uint32_t data) {
            ps.kind[pi] = k, ps.geom_id[pi] = gid, ps.prim_id[pi] = pid, ps.mask[pi] = g.mask, ps.data[pi] = data;
            ++pi;
        };
        switch (g.type) {
            case G_MESH: {
                if (skip_meshes) break;
                const size_t nv = g.verts.size() / 3, nn = g.normals.size() / 3, nt = g.indices.size() / 3;
                size_t slot = o.tris, nslot = o.tri_normals;
                for (size_t k = 0; k < nt; ++k) {
                    const size_t idx[3] = {g.indices[3 * k], g.indices[3 * k + 1], g.indices[3 * k + 2]};
                    if (idx[0] >= nv || idx[1] >= nv || idx[2] >= nv) continue;  // scene.rs:251-253: id still consumed
                    for (int v = 0; v < 3; ++v)
                        for (int c = 0; c < 3; ++c) ps.tri[9 * slot + 3 * v + c] = g.verts[3 * idx[v] + c];
                    if (g.has_normals && idx[0] < nn && idx[1] < nn && idx[2] < nn) {  // scene.rs:254-257
                        ps.tri_normal_slot[slot] = (uint32_t)nslot;
                        for (int v = 0; v < 3; ++v)
                            for (int c = 0; c < 3; ++c) ps.tri_normals[9 * nslot + 3 * v + c] = g.normals[3 * idx[v] + c];
                        ++nslot;
                    } else {
                        ps.tri_normal_slot[slot] = 0xFFFFFFFFu;
                    }
                    push(P_TRI, (uint32_t)k, (uint32_t)slot);
                    ++slot;
                }
                break;
            }
            case G_SPHERE:
                std::memcpy(&ps.sphere[4 * o.spheres], g.sphere, 4 * sizeof(float));
                push(P_SPHERE, 0, (uint32_t)o.spheres);
                break;
            case G_CURVES:
                for (size_t k = 0; k < g.segs.size() / 8; ++k) {
                    std::memcpy(&ps.curve[8 * (o.curves + k)], &g.segs[8 * k], 8 * sizeof(float));
                    push(P_CURVE, (uint32_t)k, (uint32_t)(o.curves + k));
                }
                break;
            case G_CUBICS:
                for (size_t k = 0; k < g.segs.size() / 14; ++k) {
                    std::memcpy(&ps.cubic[14 * (o.cubics + k)], &g.segs[14 * k], 14 * sizeof(float));
                    push(P_CUBIC, (uint32_t)k, (uint32_t)(o.cubics + k));
                }
                break;
            case G_INSTANCE: {
                if (!cnt[gi].insts) break;
                const ChildInfo& ci = cinfo[gi];
                InstancePrim I;
                I.child = g.child;
                I.l2w = g.l2w;
                I.w2l = invert(g.l2w);
                I.has_end = g.has_end;
                I.l2w_end = g.has_end ? g.l2w_end : g.l2w;
                I.bounds = g.has_end ? unite(transformed_box(ci.bounds, g.l2w), transformed_box(ci.bounds, g.l2w_end))
                                     : transformed_box(ci.bounds, g.l2w);
                ps.inst[o.insts] = I;
                push(P_INSTANCE, 0, (uint32_t)o.insts);
                break;
            }
        }
        // geometry inputs are dead once expanded
        if (skip_meshes && g.type == G_MESH) return;
        std::vector<float>().swap(g.verts);
        std::vector<uint32_t>().swap(g.indices);
        std::vector<float>().swap(g.normals);
        std::vector<float>().swap(g.segs);
    });
    for (size_t gi = 0; gi < ng; ++gi) {  // scene.rs:322 and the nesting depth, over the instances that were kept
        if (!cnt[gi].insts) continue;
        out.has_motion |= geoms[gi].has_end || cinfo[gi].has_motion;
        out.instance_depth = std::max(out.instance_depth, cinfo[gi].instance_depth + 1);
    }

    const bool timing = getenv("CRT_BUILD_TIMING") != nullptr;
    auto lap = [&](const char* what) {
        if (timing) fprintf(stderr, "[host build] %-28s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    lap("expand geometries");
    const size_t n = ps.size();
    out.prim_ids.resize(2 * n);
    const size_t kChunk = (size_t)1 << 16;
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    parallel_for(n_chunks, [&](size_t c) {
        for (size_t i = c * kChunk, e = std::min(n, (c + 1) * kChunk); i < e; ++i) {
            out.prim_ids[2 * i] = ps.geom_id[i];
            out.prim_ids[2 * i + 1] = ps.prim_id[i];
        }
    });
    if (ps.any_normals) {
        out.normals.assign(9 * n, 0.0f);
        for (size_t i = 0; i < n; ++i)
            if (ps.kind[i] == P_TRI && ps.tri_normal_slot[ps.data[i]] != 0xFFFFFFFFu)
                std::memcpy(&out.normals[9 * i], &ps.tri_normals[9 * (size_t)ps.tri_normal_slot[ps.data[i]]], 36);
    }
    return true;
}
}  // namespace

bool build_host_tree(std::vector<Geom>& geoms, ChildInfoFn child_info, int threads, HostTree& out, const char** err) {
    auto t0 = std::chrono::steady_clock::now();
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if (!expand_geoms(geoms, child_info, threads, out, err, false, t0)) return false;
    PrimStore& ps = out.prims;
    auto parallel_for = [&](size_t count, auto&& fn) { parallel_for_n(threads, count, fn); };
    const bool timing = getenv("CRT_BUILD_TIMING") != nullptr;
    auto lap = [&](const char* what) {
        if (timing) fprintf(stderr, "[host build] %-28s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    const size_t n = ps.size();
    const size_t kChunk = (size_t)1 << 16;
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    if (n == 0) {  // bvh.rs:310-311
        out.non_empty = false;
        for (int i = 0; i < 4; ++i) {
            out.leafs.push_back(0), out.leafs.push_back(0), out.leafs.push_back(0), out.leafs.push_back(REC_LAST | REC_SPHERE);
        }
        out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }

    std::vector<Ref> refs(n);
    std::vector<float> diag(n);
    parallel_for(n_chunks, [&](size_t c) {  // boxes in parallel ...
        for (size_t i = c * kChunk, e = std::min(n, (c + 1) * kChunk); i < e; ++i) {
            refs[i] = Ref{prim_box(ps, (uint32_t)i), (uint32_t)i};
            Vec3 ext = sub(refs[i].b.hi, refs[i].b.lo);
            diag[i] = std::sqrt(dot3(ext, ext));
        }
    });
    Box root;
    for (size_t i = 0; i < n; ++i) {  // ... the reductions in input order (a float sum depends on it)
        root = i == 0 ? refs[0].b : unite(root, refs[i].b);
        out.extent_sum += diag[i];
        out.extent_max = std::max(out.extent_max, diag[i]);
    }
    std::vector<float>().swap(diag);
    lap("ids, normals, references");
    Builder bld{ps, area(root), {threads - 1}, Par{threads, getenv("CRT_BUILD_PAR_MIN") ? (size_t)atoll(getenv("CRT_BUILD_PAR_MIN")) : ((size_t)1 << 19)}};
    Out bin;
    bld.build(bin, refs, 0);
    lap("binary SAH/SBVH build");
    out.n_binary = bin.nodes.size();
    out.non_empty = true;
    out.root = root;

    Flattener fl{ps, bin.nodes, bin.indices, out, threads};
    fl.run();
    lap("BVH4 collapse + leaf records");
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

// ---- the non-mesh primitives of a scene, for the device build ---------------------------------------------------------
bool expand_other_prims(std::vector<Geom>& geoms, ChildInfoFn child_info, uint32_t prim_index_base, OtherPrims& out, const char** err) {
    HostTree t;
    if (!expand_geoms(geoms, child_info, 1, t, err, true, std::chrono::steady_clock::now())) return false;
    const PrimStore& ps = t.prims;
    const size_t n = ps.size();
    out.n = n;
    out.lo.resize(4 * n), out.hi.resize(4 * n), out.ids.resize(2 * n), out.rec_off.resize(n + 1);
    out.has_motion = t.has_motion;
    out.instance_depth = t.instance_depth;
    for (size_t i = 0; i < n; ++i) {
        const Box b = prim_box(ps, (uint32_t)i);
        out.lo[4 * i] = b.lo.x, out.lo[4 * i + 1] = b.lo.y, out.lo[4 * i + 2] = b.lo.z, out.lo[4 * i + 3] = 1.0f;  // .w = valid
        out.hi[4 * i] = b.hi.x, out.hi[4 * i + 1] = b.hi.y, out.hi[4 * i + 2] = b.hi.z, out.hi[4 * i + 3] = 0.0f;
        out.ids[2 * i] = ps.geom_id[i], out.ids[2 * i + 1] = ps.prim_id[i];
        Vec3 ext = sub(b.hi, b.lo);
        const float diag = std::sqrt(dot3(ext, ext));
        out.extent_sum += diag;
        out.extent_max = std::max(out.extent_max, diag);
        const uint32_t at = (uint32_t)(t.leafs.size() / 4);
        out.rec_off[i] = at;
        emit_other_record(ps, (uint32_t)i, 0, t);
        // the record names its primitive by index into the scene's id table: the triangles come first there
        const size_t word = ps.kind[i] == P_SPHERE ? 4 * (size_t)(at + 1) + 1 : 4 * (size_t)(at + 1) + 3;
        t.leafs[word] += prim_index_base;
        switch (ps.kind[i]) {
            case P_SPHERE: out.counts.spheres++; break;
            case P_CURVE: out.counts.curve_segments++; break;
            case P_CUBIC: out.counts.cubic_curve_spans++; break;
            default:
                out.counts.instances++;
                out.inst_prim.push_back((uint32_t)i);
                out.inst_child.push_back(ps.inst[ps.data[i]].child);
                break;
        }
    }
    out.rec_off[n] = (uint32_t)(t.leafs.size() / 4);
    out.recs.swap(t.leafs);
    out.motion.swap(t.motion);
    return true;
}

// ---- canonical export ---------------------------------------------------------
namespace {
uint32_t rec_quads(uint32_t flags) {
    switch (flags & REC_KIND_MASK) {
        case REC_TRI: return REC_TRI_Q;
        case REC_SPHERE: return REC_SPHERE_Q;
        case REC_CURVE: return REC_CURVE_Q;
        case REC_CUBIC: return REC_CUBIC_Q;
        default: return REC_INSTANCE_Q;
    }
}
}  // namespace

void export_sizes(const HostTree& t, uint64_t out4[4]) {
    out4[0] = t.nodes.size();
    out4[1] = t.n_leaves;
    out4[2] = t.n_leaf_refs;
    out4[3] = t.n_binary;
}

void export_tree(const HostTree& t, void* nodes_out, uint32_t* tri_counts, uint32_t* other_counts, uint32_t* refs) {
    // Leaves were emitted in the same DFS order the nodes reference them, so
    // walking nodes/lanes in array order numbers them in creation order.
    uint32_t* w = (uint32_t*)nodes_out;
    uint32_t leaf_no = 0;
    size_t k = 0;
    // creation order == order of first_q; collect and sort to be safe
    std::vector<std::pair<uint32_t, std::pair<size_t, int>>> lanes;
    for (size_t ni = 0; ni < t.nodes.size(); ++ni)
        for (int l = 0; l < 4; ++l) {
            uint32_t c = t.nodes[ni].child[l];
            if (c != CRT_EMPTY_LANE && (c & CRT_LEAF_BIT)) lanes.push_back({c & ~CRT_LEAF_BIT, {ni, l}});
        }
    std::sort(lanes.begin(), lanes.end());
    std::vector<uint32_t> ordinal_of(t.nodes.size() * 4, 0);
    for (auto& e : lanes) {
        ordinal_of[e.second.first * 4 + e.second.second] = leaf_no;
        uint32_t qo = e.first, nt = 0, no = 0;
        for (;;) {
            uint32_t flags = t.leafs[4 * (size_t)qo + 3];
            uint32_t kind = flags & REC_KIND_MASK;
            uint32_t pi = kind == REC_SPHERE ? t.leafs[4 * (size_t)(qo + 1) + 1] : t.leafs[4 * (size_t)(qo + 1) + 3];
            refs[k++] = pi;
            (kind == REC_TRI ? nt : no)++;
            qo += rec_quads(flags);
            if (flags & REC_LAST) break;
        }
        tri_counts[leaf_no] = nt;
        other_counts[leaf_no] = no;
        leaf_no++;
    }
    for (size_t ni = 0; ni < t.nodes.size(); ++ni) {
        const DevNode& n = t.nodes[ni];
        uint32_t* o = w + 32 * ni;
        std::memcpy(o, n.bmin_x, 96);
        for (int l = 0; l < 4; ++l) {
            uint32_t c = n.child[l];
            o[24 + l] = (c != CRT_EMPTY_LANE && (c & CRT_LEAF_BIT)) ? ordinal_of[ni * 4 + l] : c;
        }
        o[28] = n.flags;
        o[29] = o[30] = o[31] = 0;
    }
}

}  // namespace crt
