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
    uint32_t n;
};

#define CRT_BLOCK 128
#define CRT_SSTACK 16  // shared-memory stack entries per thread
#define CRT_LSTACK 112  // local-memory spill entries per thread

// Resident blocks per SM of the kernel variant a scene with `feat` uses.
int trace_blocks_per_sm(bool any_hit, uint32_t feat);
// Enqueues one persistent launch of `grid` blocks x CRT_BLOCK threads.
cudaError_t launch_trace(bool any_hit, uint32_t feat, const TraceParams& p, int grid, cudaStream_t stream);

}  // namespace crt
