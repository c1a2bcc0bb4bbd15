// Device layout of a committed scene (shared by the host flattener and the kernels).
//
// One committed crt_scene owns three device arrays plus a 64-byte descriptor:
//
//   nodes   : DevNode[ n_nodes ], 128 B each, 128-B aligned. BVH4 in the
//             reference's DFS pre-order (bvh.rs:196-215, 1328-1375). Lane k of
//             each bounds vector is child k. child[k]: 0xFFFFFFFF = empty lane,
//             bit 31 set = leaf (low 31 bits = offset of its first record in
//             `leafs`, in 16-byte units), else index of the child node.
//   leafs   : self-contained leaf records, 16-byte granular, triangles of a
//             leaf first (in the reference's packet order, bvh.rs:1256-1290)
//             then its other primitives; the last record of a leaf carries
//             REC_LAST. SBVH reference duplication duplicates records.
//   prim_ids: uint2 (geom_id, prim_id) per primitive, indexed by the record's
//             primitive index; read once per ray at the end.
//
// Records (w = flags word in q0.w; all float4 = 16 B):
//   triangle  48 B  q0 = v0.xyz, flags | q1 = v1.xyz, prim | q2 = v2.xyz, mask
//   sphere    32 B  q0 = c.xyz,  flags | q1 = radius, prim, mask, -
//   curve     48 B  q0 = p0.xyz, flags | q1 = p1.xyz, prim  | q2 = r0, r1, mask, -
//   cubic     96 B  q0..q3 = cp[k].xyz + {flags, prim, mask, -} | q4 = r0, r1, -, - | q5 pad
//   instance  96 B  q0..q3 = w2l columns x,y,z,translation + {flags, prim, mask, motion}
//                   q4 = child descriptor pointer (lo, hi), geom_id, instance slot
//                   q5 = child node array (lo, hi), child leaf-record array (lo, hi): the descent needs no descriptor read
#pragma once
#include <stdint.h>

#define CRT_EMPTY_LANE 0xFFFFFFFFu
#define CRT_LEAF_BIT 0x80000000u

enum : uint32_t {
    REC_KIND_MASK = 0x7u,
    REC_TRI = 0u,
    REC_SPHERE = 1u,
    REC_CURVE = 2u,
    REC_CUBIC = 3u,
    REC_INSTANCE = 4u,
    REC_LAST = 1u << 3,         // last record of its leaf
    REC_DEGENERATE = 1u << 4,   // triangle whose geometric normal is exactly zero (prim.rs:80-83)
    REC_HAS_NORMALS = 1u << 5,  // triangle with per-vertex shading normals (prim.rs:77-78)
    REC_PACKET_END = 1u << 6,   // 4th triangle of a reference packet, or the leaf's last triangle
    REC_MOTION = 1u << 7,       // instance with an end-of-shutter transform
    REC_LANE_SHIFT = 8,         // triangle: lane within its packet (0..3)
    REC_CHILD_TRIS = 1u << 10,  // instance: the child scene holds triangles at its own level (its rays need the Woop shear)
};

// record sizes in 16-byte units
#define REC_TRI_Q 3u
#define REC_SPHERE_Q 2u
#define REC_CURVE_Q 3u
#define REC_CUBIC_Q 6u
#define REC_INSTANCE_Q 6u

struct U4 {  // 16-byte unit of the leaf record stream (device code reads it as uint4 / float4)
    uint32_t x, y, z, w;
};
struct U2 {
    uint32_t x, y;
};

struct alignas(128) DevNode {
    float bmin_x[4], bmin_y[4], bmin_z[4];
    float bmax_x[4], bmax_y[4], bmax_z[4];
    uint32_t child[4];
    uint32_t flags;  // reference flags word (bits 0-3 valid, 4-7 leaf); kernels derive both from child[]
    uint32_t pad[3];
};

struct MotionXform {  // l2w at shutter 0 and 1, glam column order (prim.rs:264-276)
    float l2w[12];
    float l2w_end[12];
};

// Lives in device memory; instance records point at their child's descriptor.
struct alignas(64) DevScene {
    const DevNode* nodes;
    const U4* leafs;
    const U2* prim_ids;
    const float* normals;        // 9 floats per primitive index, or NULL when no mesh has normals
    const MotionXform* motion;   // indexed by instance record's motion word
    uint32_t n_nodes;
    uint32_t has_triangles;      // bvh.rs:422-425: shear only derived when packets exist
    uint32_t shear_at_top;       // has_triangles, or some instance is a translated copy (identity linear part) of a child with
                                 // triangles: such a child reuses the top level's shear constants (traverse.cu, depth-1 variant)
    uint32_t pad[3];
};

// Instance record (host side, 4 words per quad): true if its world-to-local linear part is the identity, i.e. the instance
// is a translated copy of its child.
inline bool rec_is_translated_copy(const uint32_t* rec) {
    auto is0 = [](uint32_t w) { return (w & 0x7FFFFFFFu) == 0u; };
    const uint32_t one = 0x3F800000u;
    return rec[0] == one && is0(rec[1]) && is0(rec[2]) && is0(rec[4]) && rec[5] == one && is0(rec[6]) && is0(rec[8]) && is0(rec[9]) && rec[10] == one;
}
