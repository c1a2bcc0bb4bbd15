// Launch interface of the traversal kernels (traverse.cu): parameters, feature
// bits and the host-callable launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crt.h"
#include "crt_layout.h"

namespace crt {

// feature bits: which record kinds can be met below a scene (chosen at commit)
enum : uint32_t { FEAT_SPHERE = 1, FEAT_CURVE = 2, FEAT_INSTANCE = 4, FEAT_NORMALS = 8 };

struct TraceParams {
    DevScene scene;  // top-level descriptor by value: no dependent load to start
    const float4* rays;
    const float* time;
    const uint32_t* mask;
    float4* hits;         // closest-hit: 2 x float4 per ray
    uint8_t* flag;        // closest-hit: front_face (optional); any-hit: occluded
    unsigned int* counter;  // ray counter for persistent fetch (zeroed before launch)
    unsigned int* status;   // sticky error bits (bit 0: traversal stack overflow)
    uint32_t n;               // ray count, or the capacity bound when n_dev is set
    const uint32_t* n_dev;    // optional: ray count in device memory (read at kernel start)
    uint32_t mask_uniform;    // ray mask of every ray when `mask` is null
    unsigned long long* stats;  // null, or 7 device counters the counting build of the kernel adds to (crt_trace_counters)
};

// Tunables (overridable with -D for the tuning harness, tools/tune_trace.py)
#ifndef CRT_BLOCK
#define CRT_BLOCK 128
#endif
#ifndef CRT_SSTACK
#define CRT_SSTACK 16  // shared-memory stack entries per thread
#endif
#ifndef CRT_MINBLOCKS_CLOSEST
#define CRT_MINBLOCKS_CLOSEST 7  // resident blocks per SM the closest-hit variants are compiled for: 72 registers without spills beat 64 with (measured 6/7/8 on B200, profiles/r01_tune_phase_balance.jsonl)
#endif
#ifndef CRT_MINBLOCKS_ANY
#define CRT_MINBLOCKS_ANY 8
#endif
#ifndef CRT_MINBLOCKS_INST
#define CRT_MINBLOCKS_INST 5  // variants with instance descent carry more live state (level saves, hit chain): 96 registers without spills beat 64 with (measured 3/4/5/6/8 on B200)
#endif
#ifndef CRT_MINBLOCKS_INST1
#define CRT_MINBLOCKS_INST1 6  // closest-hit in scenes whose instances nest one level deep (the parked top-level ray lives in shared memory): 80 registers / 6 blocks beat 72 / 7 and 64 / 8 since the triangle test dropped two axis registers (measured 6/7/8 on B200, profiles/r02_tune_after_triselect.jsonl; before that 7 won, r02_tune_signsel_d1.jsonl)
#endif
#ifndef CRT_MINBLOCKS_INST1_ANY
#define CRT_MINBLOCKS_INST1_ANY 7  // any-hit, same scenes: 72 registers, no spills (8 blocks = 64 registers spills 20 bytes since the record loads are issued together)
#endif
#ifndef CRT_REFILL
#define CRT_REFILL 16  // a warp refills its idle lanes when fewer than this many are active
#endif

// Phase balancing of the while-while loop: a phase ends early once fewer than NUM/DEN of the warp's running lanes are
// still in it (0 = run every lane to the end of its phase).
#ifndef CRT_NODE_EXIT_NUM
#define CRT_NODE_EXIT_NUM 1  // node phase ends once fewer than 1/4 of the running lanes still search (measured 0, 1/8, 3/16, 1/4, 1/2, 3/4)
#endif
#ifndef CRT_NODE_EXIT_DEN
#define CRT_NODE_EXIT_DEN 4
#endif
#ifndef CRT_LEAF_EXIT_NUM
#define CRT_LEAF_EXIT_NUM 0
#endif
#ifndef CRT_LEAF_EXIT_DEN
#define CRT_LEAF_EXIT_DEN 4
#endif

#ifndef CRT_SIGNSEL
#define CRT_SIGNSEL 1  // slab test with the near / far planes picked by the ray's direction signs (by address) instead of min / max
#endif

#ifndef CRT_KEY_ORDER
#define CRT_KEY_ORDER 1  // closest-hit push order from one unsigned key per lane (entry-distance bits | internal-lane bit) for rays with t_min >= 0
#endif
#ifndef CRT_KEY_ADDR
#define CRT_KEY_ADDR 1  // ...and, while the group stays in the shared window, the slots as shared byte addresses from six compares (no masks, no popcounts)
#endif

#ifndef CRT_ANY_ORDER
#define CRT_ANY_ORDER 4  // any-hit visiting order of a node's hit lanes (the answer does not depend on it): 0 lane order, 1 nearest entry first, 2 longest overlap first, 3 farthest first, 4 farthest first for a ray with a finite t_max and nearest first for an open-ended one (measured: profiles/r02_tune_any_hit_order.jsonl)
#endif

#ifndef CRT_LEAF_PREFETCH
#define CRT_LEAF_PREFETCH 0  // L1 prefetch of a popped leaf's first record(s) while the lane waits for the leaf phase (1: first 128-byte line, 2: two)
#endif

#ifndef CRT_SADDR
#define CRT_SADDR 1  // shared-stack accesses through a register-held 32-bit shared address (inline PTX): -56 instructions in the closest-hit kernel
#endif

// Resident blocks per SM of the kernel variant a scene with `feat` and instance nesting depth `instance_depth` uses.
int trace_blocks_per_sm(bool any_hit, uint32_t feat, uint32_t instance_depth);
// Enqueues one persistent launch of `grid` blocks x CRT_BLOCK threads.
cudaError_t launch_trace(bool any_hit, uint32_t feat, uint32_t instance_depth, const TraceParams& p, int grid, cudaStream_t stream);
// Rewrites the bounds of the empty lanes of `n` device nodes to lo = +inf, hi = -inf (what the sign-selected slab test needs
// to reject them without looking at the child word); the host tree keeps the reference's +inf / +inf (bvh.rs:230-241).
cudaError_t mark_empty_lanes(DevNode* nodes, size_t n, cudaStream_t stream);

}  // namespace crt
