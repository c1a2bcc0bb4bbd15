// Library-internal declarations shared by scene.cu (crust-rt seam) and render.cu (integrator).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "crt.h"
#include "host_build.h"
#include "launch.h"

struct crt_scene {
    std::atomic<int> refs{1};
    crt::HostTree* host = nullptr;  // kept only with CRT_COMMIT_KEEP_HOST_TREE
    bool non_empty = false;
    crt::Box root{};
    bool has_motion = false;
    uint32_t n_geoms = 0;
    uint64_t n_prims = 0;
    crt_primitive_breakdown breakdown{};
    uint32_t instance_depth = 0;
    uint32_t features = 0;  // FEAT_* over this scene and everything below it
    float extent_sum = 0, extent_max = 0;
    std::vector<crt_scene*> children;  // distinct instanced scenes, retained
    crt_build_info info{};
    // device
    bool host_only = false;
    void *d_nodes = nullptr, *d_leafs = nullptr, *d_prim_ids = nullptr, *d_normals = nullptr, *d_motion = nullptr;
    DevScene desc{};
    DevScene* d_desc = nullptr;
    size_t top_bytes = 0;  // leading bytes of d_nodes that hold the tree's top levels (L2 access-policy window), 0 = none
    // element counts of the device arrays and where the instance records sit (crt_scene_export / crt_scene_import)
    uint64_t n_nodes = 0, n_leaf_quads = 0, n_prim_ids = 0, n_normal_floats = 0, n_motion = 0;
    std::vector<uint32_t> inst_q;      // offset (16-byte units) of every instance record in d_leafs
    std::vector<uint32_t> inst_child;  // index into `children` of the scene that record instances
    crt_memory_footprint footprint{};

    ~crt_scene();
};

namespace crt {
void scene_device_memory_released(bool had_window);
}

inline crt_scene::~crt_scene() {
    {
        delete host;
        cudaFree(d_nodes);
        cudaFree(d_leafs);
        cudaFree(d_prim_ids);
        cudaFree(d_normals);
        cudaFree(d_motion);
        cudaFree(d_desc);
        for (crt_scene* c : children) crt_scene_release(c);
    }
    crt::scene_device_memory_released(top_bytes != 0);
}

namespace crt {

// Sets the calling thread's crt_last_error() message and returns `code`.
int set_error(int code, const std::string& msg);
int set_cuda_error(cudaError_t e, const char* what);
// crt_init side effects (device properties, counter ring); CRT_ERR_NO_DEVICE without a GPU.
int require_device();
int device_sm_count();
unsigned int* device_status_word();
void count_launches(uint64_t n);

// One persistent traversal launch whose ray count lives in device memory (wavefront queues).
// `counter` must be zero when the kernel starts. No host synchronisation.
struct QueueTrace {
    const float4* rays;
    const float* time;         // or null (time 0)
    uint32_t mask_uniform;     // every ray of one queue has the same category mask
    float4* hits;              // closest-hit
    uint8_t* flag;             // closest-hit: front_face; any-hit: occluded
    const uint32_t* n_dev;     // queue length
    uint32_t capacity;         // upper bound of *n_dev
    unsigned int* counter;     // persistent-fetch ray counter
};
int enqueue_queue_trace(const crt_scene* s, bool any_hit, const QueueTrace& q, cudaStream_t stream);
// L2 access-policy window over the top levels of `s` for the launches that follow on `stream`; the caller clears it when its
// launches are enqueued (the stream belongs to the user). enqueue_queue_trace does not touch the attribute itself.
bool hint_scene_top_levels(const crt_scene* s, cudaStream_t stream);
void clear_scene_hint(cudaStream_t stream);
void scene_device_memory_released(bool had_window);

// device_build.cu
bool device_build_supported(const std::vector<Geom>& geoms);
int device_build(std::vector<Geom>& geoms, ChildInfoFn child_info, crt_scene* s);

}  // namespace crt
