// Shared helpers for the gencomm_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gencomm_b200.h"

namespace gc {

void set_error(const char *fmt, ...);

#define GC_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            gc::set_error(__VA_ARGS__);        \
            return (code);                     \
        }                                      \
    } while (0)

// Launch check: returns the (positive) cudaError_t of the launch, if any.
#define GC_LAUNCH_CHECK(name)                                              \
    do {                                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                           \
        if (e__ != cudaSuccess) {                                          \
            gc::set_error("%s: %s", name, cudaGetErrorString(e__));        \
            (void)cudaGetLastError();                                      \
            return (int)e__;                                               \
        }                                                                  \
    } while (0)

constexpr uint32_t kEmpty = 0xFFFFFFFFu;    // cell without points / free slot
constexpr uint32_t kDropped = 0xFFFFFFFEu;  // cell beyond max_voxels
constexpr uint32_t kPillarBit = 0x80000000u;

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Private layout of the voxelizer workspace.
struct VoxelWorkspace {
    uint32_t *cell_code;    // [A][ncell]  pass 1: min point index; afterwards kPillarBit|pillar, kDropped, kEmpty
    int32_t *point_cell;    // [total_points] cell id or -1
    uint32_t *slots;        // [A][max_voxels][32] ascending point indices (kEmpty = free)
    int32_t *pillar_cell;   // [A][max_voxels] cell id of each pillar
    int32_t *block_counts;  // [A][blocks_per_agent] first-point counts per block (look-back flags of k_pillar_build)
    int32_t *n_pillars;     // [A] copy of gc_voxelize's n_pillars output (read by the pillar-centric planes writer)
    size_t bytes;
};

inline VoxelWorkspace carve_workspace(void *base, const gcVoxelGeom &g, int n_agents, int total_points) {
    VoxelWorkspace w;
    const size_t ncell = (size_t)g.grid[0] * g.grid[1] * g.grid[2];
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) {
        char *p = b ? b + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.cell_code = (uint32_t *)take((size_t)n_agents * ncell * 4);
    w.point_cell = (int32_t *)take((size_t)(total_points > 0 ? total_points : 1) * 4);
    w.slots = (uint32_t *)take((size_t)n_agents * g.max_voxels * 32 * 4);
    w.pillar_cell = (int32_t *)take((size_t)n_agents * g.max_voxels * 4);
    // one count / look-back flag per block of >= 128 points; any agent has at most total_points points
    const size_t blocks = (size_t)(total_points + 127) / 128 + 2;
    w.block_counts = (int32_t *)take((size_t)n_agents * blocks * 4);
    w.n_pillars = (int32_t *)take((size_t)n_agents * 4);
    w.bytes = off;
    return w;
}

}  // namespace gc
