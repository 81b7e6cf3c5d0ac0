// PointPillars front end for sm_100a: voxelize -> PillarVFE -> PointPillarScatter.
//
// Replaces (paths relative to /root/reference/opencood):
//   data_utils/pre_processor/sp_voxel_preprocessor.py:62-85  (spconv Point2VoxelCPU3d, CPU, sequential)
//   models/sub_modules/pillar_vfe.py:105-155 (+ PFNLayer :31-53)
//   models/sub_modules/point_pillar_scatter.py:19-76
//
// Design: the spconv voxelizer is a sequential first-come algorithm; its
// result is reproduced exactly by an order-preserving parallel formulation:
//   1. k_cell_assign : cell id per point; cell_code[cell] = min point index (atomicMin).
//   2. k_pillar_count/k_pillar_assign : a point is the "first point" of its cell iff
//      cell_code[cell] == its index.  Pillar id = number of first-points with a smaller index
//      (block counts + in-block ballot scan) == spconv's creation order; ids >= max_voxels are
//      dropped exactly like spconv's `num_voxels >= max_voxels` test.
//   3. k_slot_insert : the first 32 points of a pillar in input order == the 32 smallest point
//      indices of the cell: a conserving atomicMin insertion chain over 32 slots per pillar
//      (order-independent final state -> deterministic).
//   4. k_canvas<...> : canvas-stationary writer.  One CTA owns 128 consecutive cells of one canvas
//      row, evaluates the PFN for the occupied ones (one warp per pillar, one lane per channel
//      pair, loop over the <=32 valid points only) and stores every canvas byte exactly once with
//      128-bit coalesced stores.  No memset, no pillar_features round trip.
// All of it is HBM-bound integer/byte work + ~1 kFLOP per pillar; no tensor cores.
#include "common.cuh"

namespace gc {

// ------------------------------------------------------------------------------------------------
// agent lookup: largest a with off[a] <= g
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const int32_t *__restrict__ off, int n, int g) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

struct GeomDev {
    float rmin[3];
    float voxel[3];
    int grid[3];
    int max_voxels;
    int ncell;
};

// ------------------------------------------------------------------------------------------------
// 1. cell assignment.  c_j = floor((p_j - min_j) / voxel_j) in IEEE fp32 (no FMA contraction, true
//    division) -- spconv's arithmetic, SURVEY.md App. A.1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cell_assign(const float4 *__restrict__ points, const int32_t *__restrict__ point_offsets, int n_agents,
              int total_points, GeomDev g, uint32_t *__restrict__ cell_code, int32_t *__restrict__ point_cell) {
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total_points) return;
    if (gi >= __ldg(point_offsets + n_agents)) { point_cell[gi] = -1; return; }
    const int a = find_segment(point_offsets, n_agents, gi);
    const float4 p = __ldg(points + gi);
    const float pv[3] = {p.x, p.y, p.z};
    int c[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float f = floorf(__fdiv_rn(__fsub_rn(pv[j], g.rmin[j]), g.voxel[j]));
        ok = ok && (f >= 0.0f) && (f < (float)g.grid[j]);   // false for NaN
        c[j] = (int)f;
    }
    int cell = -1;
    if (ok) {
        cell = (c[2] * g.grid[1] + c[1]) * g.grid[0] + c[0];
        atomicMin(cell_code + (size_t)a * g.ncell + cell, (uint32_t)(gi - __ldg(point_offsets + a)));
    }
    point_cell[gi] = cell;
}

// ------------------------------------------------------------------------------------------------
// 2. pillar ranking.  grid = (blocks_per_agent, n_agents), 256 threads x 4 items = 1024 points,
//    item-major order i = blk*1024 + item*256 + tid.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_first_point(const int32_t *__restrict__ point_cell, const uint32_t *codes,
                                               int base, int i, int n) {
    if (i >= n) return false;
    const int c = point_cell[base + i];
    if (c < 0) return false;
    return codes[c] == (uint32_t)i;   // plain load: codes are rewritten by other CTAs of k_pillar_assign
}

__global__ void __launch_bounds__(256)
k_pillar_count(const int32_t *__restrict__ point_offsets, const int32_t *__restrict__ point_cell,
               const uint32_t *__restrict__ cell_code, int ncell, int blocks_per_agent,
               int32_t *__restrict__ block_counts) {
    const int a = blockIdx.y;
    const int base = __ldg(point_offsets + a);
    const int n = __ldg(point_offsets + a + 1) - base;
    const int i0 = blockIdx.x * 1024;
    int cnt = 0;
    if (i0 < n) {
        const uint32_t *codes = cell_code + (size_t)a * ncell;
#pragma unroll
        for (int it = 0; it < 4; ++it)
            cnt += is_first_point(point_cell, codes, base, i0 + it * 256 + threadIdx.x, n) ? 1 : 0;
    }
    const int total = __syncthreads_count(cnt & 1) + 2 * __syncthreads_count(cnt & 2) + 4 * __syncthreads_count(cnt & 4);
    if (threadIdx.x == 0) block_counts[(size_t)a * blocks_per_agent + blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
k_pillar_assign(const int32_t *__restrict__ point_offsets, const int32_t *__restrict__ point_cell,
                uint32_t *__restrict__ cell_code, int ncell, int max_voxels, int blocks_per_agent,
                const int32_t *__restrict__ block_counts, uint32_t *__restrict__ slots,
                int32_t *__restrict__ pillar_cell, int32_t *__restrict__ n_pillars) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int a = blockIdx.y;
    const int base = __ldg(point_offsets + a);
    const int n = __ldg(point_offsets + a + 1) - base;
    const int i0 = blockIdx.x * 1024;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *bc = block_counts + (size_t)a * blocks_per_agent;

    if (blockIdx.x == 0 && warp == 1) {   // total number of occupied cells of this agent
        int t = 0;
        for (int k = lane; k < blocks_per_agent; k += 32) t += bc[k];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (lane == 0) n_pillars[a] = t < max_voxels ? t : max_voxels;
    }
    if (i0 >= n) return;   // uniform per block

    if (warp == 0) {       // exclusive prefix of the preceding blocks
        int t = 0;
        for (int k = lane; k < (int)blockIdx.x; k += 32) t += bc[k];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (lane == 0) s_base = t;
    }
    uint32_t *codes = cell_code + (size_t)a * ncell;
    bool flag[4];
    int cellv[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int i = i0 + it * 256 + threadIdx.x;
        flag[it] = is_first_point(point_cell, codes, base, i, n);
        cellv[it] = flag[it] ? point_cell[base + i] : -1;
    }
    __syncthreads();
    int running = s_base;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const unsigned bal = __ballot_sync(0xffffffffu, flag[it]);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int v = s_warp[w];
            before += (w < warp) ? v : 0;
            total += v;
        }
        if (flag[it]) {
            const int pid = running + before + __popc(bal & ((1u << lane) - 1u));
            if (pid < max_voxels) {
                codes[cellv[it]] = kPillarBit | (uint32_t)pid;
                const size_t gp = (size_t)a * max_voxels + pid;
                pillar_cell[gp] = cellv[it];
                uint4 *s = reinterpret_cast<uint4 *>(slots + gp * 32);
                const uint4 e = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] = e;
            } else {
                codes[cellv[it]] = kDropped;
            }
        }
        running += total;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// 3. slot insertion: conserving atomicMin chain.  Each atomicMin(slot, v) leaves min(slot, v) in the
//    slot and carries max(slot, v) on, so the multiset {slots} U {carried} is invariant and the final
//    slot contents are the 32 smallest indices in ascending order, whatever the interleaving.
// ------------------------------------------------------------------------------------------------
// grid = (n_agents, chunks): blockIdx.x (fastest-scheduled) is the agent, blockIdx.y the 256-point chunk, so
// the machine sweeps every agent's points in index order.  Late (large-index) points of a crowded cell then
// find 32 smaller indices already in place and leave without a single atomic.
__global__ void __launch_bounds__(256)
k_slot_insert(const int32_t *__restrict__ point_offsets, const int32_t *__restrict__ point_cell,
              const uint32_t *__restrict__ cell_code, int ncell, int max_voxels, uint32_t *__restrict__ slots) {
    const int a = blockIdx.x;
    const int base = __ldg(point_offsets + a);
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= __ldg(point_offsets + a + 1) - base) return;
    const int c = __ldg(point_cell + base + i);
    if (c < 0) return;
    const uint32_t code = __ldg(cell_code + (size_t)a * ncell + c);
    if (code >= kDropped) return;
    uint32_t *s = slots + ((size_t)a * max_voxels + (code & ~kPillarBit)) * 32;
    uint32_t v = (uint32_t)i;
    // The slot array is sorted ascending at every instant and its values only ever decrease, so a slot
    // observed below v stays below v and atomicMin(slot, v) there is a no-op: skip all of them.  The
    // snapshot is read through L2 (ld.cg); a stale (larger) value only skips less.
    int k0 = 0;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
    // ascending order: stop at the first quad that holds a slot >= v (most pillars hold a handful of points,
    // so one 16-byte read instead of the whole 128-byte row)
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const uint4 u = __ldcg(s4 + q);
        const int below = (u.x < v) + (u.y < v) + (u.z < v) + (u.w < v);
        k0 += below;
        if (below < 4) break;
    }
#pragma unroll 1
    for (int k = k0; k < 32; ++k) {
        const uint32_t old = atomicMin(s + k, v);
        if (old == kEmpty) break;
        v = old > v ? old : v;
        if ((k & 3) == 3 && __ldcg(s + 31) < v) break;   // 32 smaller indices already present
    }
}

// ------------------------------------------------------------------------------------------------
// PFN evaluation for one pillar by one warp ("kernel order", mirrored by oracle/pillar_ref.c).
// Lane l owns output channels l and l+32; lane s also holds point s of the pillar.
// ------------------------------------------------------------------------------------------------
struct PfnLane {     // packed row of the pfn table, see include/gencomm_b200.h
    float a0, a1, a2, a3, b0, b1, b2, d0, d1, d2, shift, pad;
};

__device__ __forceinline__ PfnLane load_pfn(const float *__restrict__ pfn, int ch) {
    const float4 *r = reinterpret_cast<const float4 *>(pfn + ch * 16);
    const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
    PfnLane w;
    w.a0 = a.x; w.a1 = a.y; w.a2 = a.z; w.a3 = a.w;
    w.b0 = b.x; w.b1 = b.y; w.b2 = b.z; w.d0 = b.w;
    w.d1 = c.x; w.d2 = c.y; w.shift = c.z; w.pad = c.w;
    return w;
}

__device__ __forceinline__ float tree_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

__device__ __forceinline__ float pfn_bias(const PfnLane &w, float cx, float cy, float cz, float mx, float my, float mz) {
    float b = __fmaf_rn(w.b0, cx, w.shift);
    b = __fmaf_rn(w.b1, cy, b);
    b = __fmaf_rn(w.b2, cz, b);
    b = __fmaf_rn(w.d0, mx, b);
    b = __fmaf_rn(w.d1, my, b);
    b = __fmaf_rn(w.d2, mz, b);
    return b;
}

__device__ __forceinline__ float pfn_point(const PfnLane &w, float xr, float yr, float zr, float pi) {
    float acc = __fmaf_rn(w.a2, zr, __fmul_rn(w.a3, pi));
    acc = __fmaf_rn(w.a1, yr, acc);
    return __fmaf_rn(w.a0, xr, acc);
}

__device__ __forceinline__ float pfn_finish(const PfnLane &w, float best, float b, int n) {
    float o = fmaxf(__fadd_rn(best, b), 0.0f);
    if (n < 32) o = fmaxf(o, w.pad);   // padded slots contribute relu(bn(0)) = max(shift, 0), pillar_vfe.py:46
    return o;
}

// p: this lane's point (ignored when lane >= n); returns the two channel maxima of the pillar.
// stage: 32 float4 of shared memory private to the calling warp.  The decorated points are exchanged through it
// (one broadcast LDS.128 per point instead of four shuffles); arithmetic and order are unchanged.
__device__ __forceinline__ float2 pfn_pillar(const PfnLane &wa, const PfnLane &wb, float4 p, int n, float cx,
                                             float cy, float cz, float4 *__restrict__ stage) {
    const int lane = threadIdx.x & 31;
    const bool valid = lane < n;
    const float xr = valid ? __fsub_rn(p.x, cx) : 0.0f;
    const float yr = valid ? __fsub_rn(p.y, cy) : 0.0f;
    const float zr = valid ? __fsub_rn(p.z, cz) : 0.0f;
    const float pi = valid ? p.w : 0.0f;
    __syncwarp();                                   // every lane is done reading the previous pillar's points
    stage[lane] = make_float4(xr, yr, zr, pi);
    const float inv_n = __frcp_rn((float)n);
    const float mx = __fmul_rn(tree_sum(xr), inv_n);
    const float my = __fmul_rn(tree_sum(yr), inv_n);
    const float mz = __fmul_rn(tree_sum(zr), inv_n);
    const float ba = pfn_bias(wa, cx, cy, cz, mx, my, mz);
    const float bb = pfn_bias(wb, cx, cy, cz, mx, my, mz);
    float best_a = -INFINITY, best_b = -INFINITY;
    __syncwarp();                                   // staged points visible to the whole warp
#pragma unroll 4
    for (int s = 0; s < n; ++s) {
        const float4 q = stage[s];
        best_a = fmaxf(best_a, pfn_point(wa, q.x, q.y, q.z, q.w));
        best_b = fmaxf(best_b, pfn_point(wb, q.x, q.y, q.z, q.w));
    }
    return make_float2(pfn_finish(wa, best_a, ba, n), pfn_finish(wb, best_b, bb, n));
}

// ------------------------------------------------------------------------------------------------
// Standalone PillarVFE on reference-shaped voxel tensors: one warp per pillar.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pillar_vfe(const float4 *__restrict__ voxels, const int32_t *__restrict__ num_points,
             const int4 *__restrict__ coords, int n_pillars, const float *__restrict__ pfn, float vx, float vy,
             float vz, float ox, float oy, float oz, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= n_pillars) return;   // warp-uniform
    const PfnLane wa = load_pfn(pfn, lane), wb = load_pfn(pfn, lane + 32);
    int n = __ldg(num_points + m);
    n = n < 0 ? 0 : (n > 32 ? 32 : n);
    const int4 c = __ldg(coords + m);   // (b, z, y, x)
    const float cx = __fadd_rn(__fmul_rn((float)c.w, vx), ox);
    const float cy = __fadd_rn(__fmul_rn((float)c.z, vy), oy);
    const float cz = __fadd_rn(__fmul_rn((float)c.y, vz), oz);
    __shared__ float4 s_stage[8][32];
    const float4 p = __ldg(voxels + (size_t)m * 32 + lane);
    const float2 r = pfn_pillar(wa, wb, p, n, cx, cy, cz, s_stage[threadIdx.x >> 5]);
    out[(size_t)m * 64 + lane] = r.x;
    out[(size_t)m * 64 + lane + 32] = r.y;
}

// ------------------------------------------------------------------------------------------------
// Reference-shaped voxel tensors out of the workspace (SpVoxelPreprocessor.preprocess + collate).
// One warp per (compact) pillar.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_voxel_gather(const float4 *__restrict__ points, const int32_t *__restrict__ point_offsets, int n_agents,
               GeomDev g, const uint32_t *__restrict__ slots, const int32_t *__restrict__ pillar_cell,
               const int32_t *__restrict__ pillar_offsets, int total_pillars, float4 *__restrict__ voxels,
               int4 *__restrict__ coords, int32_t *__restrict__ num_points) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= total_pillars) return;
    const int a = find_segment(pillar_offsets, n_agents, m);
    const size_t gp = (size_t)a * g.max_voxels + (m - __ldg(pillar_offsets + a));
    const uint32_t idx = __ldg(slots + gp * 32 + lane);
    const bool valid = idx != kEmpty;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) p = __ldg(points + __ldg(point_offsets + a) + idx);
    voxels[(size_t)m * 32 + lane] = p;
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) {
        const int cell = __ldg(pillar_cell + gp);
        const int x = cell % g.grid[0];
        const int y = (cell / g.grid[0]) % g.grid[1];
        const int z = cell / (g.grid[0] * g.grid[1]);
        coords[m] = make_int4(a, z, y, x);
        num_points[m] = __popc(bal);
    }
}

// ------------------------------------------------------------------------------------------------
// 4. canvas-stationary writer.  grid = (ceil(nx/128), ny, n_batch), 256 threads.
//    Src::occupied(code) / Src::features(...) select between
//      FeatSrc  : pillar_features rows looked up through a dense cell->row map   (PointPillarScatter)
//      FusedSrc : PFN evaluated from the voxelizer workspace                     (fused front end)
// ------------------------------------------------------------------------------------------------
constexpr int kTileX = 128;
constexpr int kTileStride = kTileX + 4;   // floats; keeps rows 16-byte aligned

struct FeatSrc {
    const float *feat;        // [M][C]
    const int32_t *cell_map;  // [n_batch][ny*nx], -1 = empty
    int C;
    static constexpr bool kFused = false;
    __device__ __forceinline__ int code(int b, int cell, int ncell) const {
        return __ldg(cell_map + (size_t)b * ncell + cell);
    }
    __device__ __forceinline__ static bool occupied(int code) { return code >= 0; }
};

struct FusedSrc {
    const float4 *points;
    const int32_t *point_offsets;
    const uint32_t *cell_code;
    const uint32_t *slots;
    const float *pfn;
    int max_voxels;
    float vx, vy, ox, oy, cz;   // cz = centre of z-cell 0
    static constexpr bool kFused = true;
    __device__ __forceinline__ int code(int b, int cell, int ncell) const {
        return (int)__ldg(cell_code + (size_t)b * ncell + cell);
    }
    // kPillarBit|pid is a negative int below kDropped (-2) and kEmpty (-1)
    __device__ __forceinline__ static bool occupied(int code) { return code < -2; }
};

template <class Src>
__global__ void __launch_bounds__(256, 4)
k_canvas(Src src, int nx, int ny, int C, float *__restrict__ canvas) {
    __shared__ __align__(16) float tile[64 * kTileStride];
    __shared__ int s_code[kTileX];
    __shared__ int s_list[kTileX];
    __shared__ unsigned s_mask[4];
    __shared__ float4 s_stage[8][32];   // per-warp point exchange of pfn_pillar

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int x0 = blockIdx.x * kTileX, y = blockIdx.y, b = blockIdx.z;
    const int ncell = nx * ny;

    if (t < kTileX) {
        const int x = x0 + t;
        const int code = (x < nx) ? src.code(b, y * nx + x, ncell) : -1;
        const bool occ = (x < nx) && Src::occupied(code);
        s_code[t] = code;
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        if (lane == 0) s_mask[warp] = bal;
    }
    __syncthreads();
    const unsigned m0 = s_mask[0], m1 = s_mask[1], m2 = s_mask[2], m3 = s_mask[3];
    const int n_occ = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
    if (t < kTileX) {
        const unsigned mine = warp == 0 ? m0 : warp == 1 ? m1 : warp == 2 ? m2 : m3;
        if ((mine >> lane) & 1u) {
            int pos = __popc(mine & ((1u << lane) - 1u));
            pos += (warp > 0 ? __popc(m0) : 0) + (warp > 1 ? __popc(m1) : 0) + (warp > 2 ? __popc(m2) : 0);
            s_list[pos] = t;
        }
    }
    __syncthreads();

    const bool vec_ok = (nx & 3) == 0;
    const size_t plane = (size_t)ny * nx;

    for (int c0 = 0; c0 < C; c0 += 64) {
        // the tile starts as zeros (empty cells); the load phase overwrites the occupied columns, so the store
        // phase is a plain shared -> global copy (no per-quad occupancy selects)
        {
            float4 *t4 = reinterpret_cast<float4 *>(tile);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = t; i < 64 * kTileStride / 4; i += 256) t4[i] = z;
        }
        __syncthreads();
        // ---- load phase: one warp per occupied cell -------------------------------------------
        if constexpr (!Src::kFused) {
            const Src &fs = src;
            for (int k = warp; k < n_occ; k += 8) {
                const int xc = s_list[k];
                const float *row = fs.feat + (size_t)s_code[xc] * C + c0;
                if (c0 + lane < C) tile[lane * kTileStride + xc] = __ldg(row + lane);
                if (c0 + lane + 32 < C) tile[(lane + 32) * kTileStride + xc] = __ldg(row + lane + 32);
            }
        } else {
            const Src &fs = src;
            if (warp < n_occ) {   // warp-uniform; skips the weight loads for empty tiles
                const PfnLane wa = load_pfn(fs.pfn, lane), wb = load_pfn(fs.pfn, lane + 32);
                const int pbase = __ldg(fs.point_offsets + b);
                const float cy = __fadd_rn(__fmul_rn((float)y, fs.vy), fs.oy);
                // two-stage software pipeline over this warp's pillars: slot indices two ahead, point one ahead
                auto slot_of = [&](int k) -> uint32_t {
                    const unsigned pid = (unsigned)s_code[s_list[k]] & ~kPillarBit;
                    return __ldg(fs.slots + ((size_t)b * fs.max_voxels + pid) * 32 + lane);
                };
                auto point_of = [&](uint32_t idx) -> float4 {
                    return idx != kEmpty ? __ldg(fs.points + pbase + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                };
                uint32_t idx_cur = slot_of(warp);
                uint32_t idx_nxt = (warp + 8 < n_occ) ? slot_of(warp + 8) : kEmpty;
                float4 p_cur = point_of(idx_cur);
                for (int k = warp; k < n_occ; k += 8) {
                    const float4 p_nxt = point_of(idx_nxt);
                    const uint32_t idx_nxt2 = (k + 16 < n_occ) ? slot_of(k + 16) : kEmpty;
                    const int xc = s_list[k];
                    const int n = __popc(__ballot_sync(0xffffffffu, idx_cur != kEmpty));   // slots fill from 0
                    const float cx = __fadd_rn(__fmul_rn((float)(x0 + xc), fs.vx), fs.ox);
                    const float2 r = pfn_pillar(wa, wb, p_cur, n, cx, cy, fs.cz, s_stage[warp]);
                    tile[lane * kTileStride + xc] = r.x;
                    tile[(lane + 32) * kTileStride + xc] = r.y;
                    idx_cur = idx_nxt; p_cur = p_nxt; idx_nxt = idx_nxt2;
                }
            }
        }
        __syncthreads();
        // ---- store phase: every canvas byte of the tile exactly once ---------------------------
        {
            const int x = x0 + 4 * lane;
            float *dst = canvas + ((size_t)b * C + c0 + warp) * plane + (size_t)y * nx + x;
            const float4 *src4 = reinterpret_cast<const float4 *>(tile + warp * kTileStride) + lane;
            if (vec_ok && x + 3 < nx) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    if (c0 + it * 8 + warp < C) *reinterpret_cast<float4 *>(dst + (size_t)it * 8 * plane) = src4[it * 8 * (kTileStride / 4)];
                }
            } else {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    if (c0 + it * 8 + warp >= C) continue;
                    const float4 v = src4[it * 8 * (kTileStride / 4)];
                    float *d = dst + (size_t)it * 8 * plane;
                    if (x < nx) d[0] = v.x;
                    if (x + 1 < nx) d[1] = v.y;
                    if (x + 2 < nx) d[2] = v.z;
                    if (x + 3 < nx) d[3] = v.w;
                }
            }
        }
        __syncthreads();
    }
}

// dense cell -> pillar-row map for the standalone scatter
__global__ void __launch_bounds__(256)
k_build_cell_map(const int4 *__restrict__ coords, int n_pillars, int nx, int ny, int n_batch,
                 int32_t *__restrict__ cell_map) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_pillars) return;
    const int4 c = __ldg(coords + m);   // (b, z, y, x); idx = z + y*nx + x  (point_pillar_scatter.py:58)
    const long long idx = (long long)c.y + (long long)c.z * nx + c.w;
    if (c.x < 0 || c.x >= n_batch || idx < 0 || idx >= (long long)nx * ny) return;
    cell_map[(size_t)c.x * nx * ny + idx] = m;
}

static int make_geom(const gcVoxelGeom *geom, GeomDev *g) {
    GC_REQUIRE(geom != nullptr, GC_EINVAL, "geom is null");
    GC_REQUIRE(geom->max_points == GC_MAX_POINTS_PER_PILLAR, GC_EUNSUPPORTED,
               "max_points_per_voxel must be 32 (got %d)", geom->max_points);
    GC_REQUIRE(geom->grid[0] > 0 && geom->grid[1] > 0 && geom->grid[2] > 0 && geom->max_voxels > 0, GC_EINVAL,
               "bad voxel grid");
    const long long ncell = (long long)geom->grid[0] * geom->grid[1] * geom->grid[2];
    GC_REQUIRE(ncell < (1ll << 30), GC_EUNSUPPORTED, "voxel grid too large");
    for (int j = 0; j < 3; ++j) {
        g->rmin[j] = geom->range_min[j];
        g->voxel[j] = geom->voxel[j];
        g->grid[j] = geom->grid[j];
    }
    g->max_voxels = geom->max_voxels;
    g->ncell = (int)ncell;
    return GC_OK;
}

static inline int blocks_per_agent(int max_agent_points) { return (max_agent_points + 1023) / 1024 + 1; }

}  // namespace gc

using namespace gc;

extern "C" size_t gc_voxelize_workspace_bytes(const gcVoxelGeom *geom, int n_agents, int total_points) {
    if (!geom || n_agents <= 0 || total_points < 0) return 0;
    return carve_workspace(nullptr, *geom, n_agents, total_points).bytes;
}

extern "C" int gc_voxelize(const float *points, const int32_t *point_offsets, int n_agents, int total_points,
                           int max_agent_points, const gcVoxelGeom *geom, void *workspace, int32_t *n_pillars,
                           void *stream) {
    GeomDev g;
    if (int rc = make_geom(geom, &g)) return rc;
    GC_REQUIRE(points && point_offsets && workspace && n_pillars, GC_EINVAL, "gc_voxelize: null pointer");
    GC_REQUIRE(n_agents > 0 && n_agents <= 65535 && total_points >= 0, GC_EINVAL, "gc_voxelize: bad sizes");
    if (max_agent_points <= 0 || max_agent_points > total_points) max_agent_points = total_points;
    cudaStream_t st = (cudaStream_t)stream;
    const VoxelWorkspace w = carve_workspace(workspace, *geom, n_agents, total_points);
    const int bpa = blocks_per_agent(max_agent_points);
    cudaMemsetAsync(w.cell_code, 0xFF, (size_t)n_agents * g.ncell * 4, st);
    if (total_points > 0) {
        k_cell_assign<<<(total_points + 255) / 256, 256, 0, st>>>((const float4 *)points, point_offsets, n_agents,
                                                                 total_points, g, w.cell_code, w.point_cell);
        GC_LAUNCH_CHECK("k_cell_assign");
    }
    k_pillar_count<<<dim3(bpa, n_agents), 256, 0, st>>>(point_offsets, w.point_cell, w.cell_code, g.ncell, bpa,
                                                        w.block_counts);
    GC_LAUNCH_CHECK("k_pillar_count");
    k_pillar_assign<<<dim3(bpa, n_agents), 256, 0, st>>>(point_offsets, w.point_cell, w.cell_code, g.ncell,
                                                         g.max_voxels, bpa, w.block_counts, w.slots, w.pillar_cell,
                                                         n_pillars);
    GC_LAUNCH_CHECK("k_pillar_assign");
    if (total_points > 0) {
        const int chunks = (max_agent_points + 255) / 256;
        GC_REQUIRE(chunks <= 65535, GC_EUNSUPPORTED, "gc_voxelize: more than 16.7M points per agent");
        k_slot_insert<<<dim3(n_agents, chunks), 256, 0, st>>>(point_offsets, w.point_cell, w.cell_code, g.ncell,
                                                            g.max_voxels, w.slots);
        GC_LAUNCH_CHECK("k_slot_insert");
    }
    return GC_OK;
}

extern "C" int gc_voxel_gather(const float *points, const int32_t *point_offsets, int n_agents, int total_points,
                               const gcVoxelGeom *geom, const void *workspace, const int32_t *pillar_offsets,
                               int total_pillars, float *voxels, int32_t *coords, int32_t *num_points,
                               void *stream) {
    GeomDev g;
    if (int rc = make_geom(geom, &g)) return rc;
    GC_REQUIRE(points && point_offsets && workspace && pillar_offsets, GC_EINVAL, "gc_voxel_gather: null pointer");
    GC_REQUIRE(n_agents > 0 && total_pillars >= 0, GC_EINVAL, "gc_voxel_gather: bad sizes");
    if (total_pillars == 0) return GC_OK;
    GC_REQUIRE(voxels && coords && num_points, GC_EINVAL, "gc_voxel_gather: null output");
    const VoxelWorkspace w = carve_workspace(const_cast<void *>(workspace), *geom, n_agents, total_points);
    k_voxel_gather<<<(total_pillars + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)points, point_offsets, n_agents, g, w.slots, w.pillar_cell, pillar_offsets, total_pillars,
        (float4 *)voxels, (int4 *)coords, num_points);
    GC_LAUNCH_CHECK("k_voxel_gather");
    return GC_OK;
}

extern "C" int gc_pillar_vfe(const float *voxels, const int32_t *num_points, const int32_t *coords, int n_pillars,
                             const float *pfn, const float voxel[3], const float centre_offset[3],
                             float *pillar_features, void *stream) {
    GC_REQUIRE(n_pillars >= 0, GC_EINVAL, "gc_pillar_vfe: negative pillar count");
    if (n_pillars == 0) return GC_OK;
    GC_REQUIRE(voxels && num_points && coords && pfn && voxel && centre_offset && pillar_features, GC_EINVAL,
               "gc_pillar_vfe: null pointer");
    k_pillar_vfe<<<(n_pillars + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)voxels, num_points, (const int4 *)coords, n_pillars, pfn, voxel[0], voxel[1], voxel[2],
        centre_offset[0], centre_offset[1], centre_offset[2], pillar_features);
    GC_LAUNCH_CHECK("k_pillar_vfe");
    return GC_OK;
}

extern "C" int gc_scatter_canvas(const float *pillar_features, const int32_t *coords, int n_pillars, int C, int nx,
                                 int ny, int n_batch, int32_t *cell_map, float *canvas, void *stream) {
    GC_REQUIRE(n_pillars >= 0 && C > 0 && nx > 0 && ny > 0 && n_batch >= 0, GC_EINVAL, "gc_scatter_canvas: bad sizes");
    GC_REQUIRE(ny <= 65535 && n_batch <= 65535, GC_EUNSUPPORTED, "gc_scatter_canvas: grid too large");
    if (n_batch == 0) return GC_OK;
    GC_REQUIRE(cell_map && canvas && (n_pillars == 0 || (pillar_features && coords)), GC_EINVAL,
               "gc_scatter_canvas: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(cell_map, 0xFF, (size_t)n_batch * nx * ny * 4, st);
    if (n_pillars > 0) {
        k_build_cell_map<<<(n_pillars + 255) / 256, 256, 0, st>>>((const int4 *)coords, n_pillars, nx, ny, n_batch,
                                                                 cell_map);
        GC_LAUNCH_CHECK("k_build_cell_map");
    }
    FeatSrc src{pillar_features, cell_map, C};
    k_canvas<FeatSrc><<<dim3((nx + kTileX - 1) / kTileX, ny, n_batch), 256, 0, st>>>(src, nx, ny, C, canvas);
    GC_LAUNCH_CHECK("k_canvas<FeatSrc>");
    return GC_OK;
}

extern "C" int gc_pillar_canvas(const float *points, const int32_t *point_offsets, int n_agents, int total_points,
                                const gcVoxelGeom *geom, const void *workspace, const float *pfn,
                                const float centre_offset[3], float *canvas, void *stream) {
    GeomDev g;
    if (int rc = make_geom(geom, &g)) return rc;
    GC_REQUIRE(g.grid[2] == 1, GC_EUNSUPPORTED, "PointPillarScatter requires nz == 1 (point_pillar_scatter.py:17)");
    GC_REQUIRE(points && point_offsets && workspace && pfn && centre_offset && canvas, GC_EINVAL,
               "gc_pillar_canvas: null pointer");
    GC_REQUIRE(n_agents > 0 && n_agents <= 65535 && g.grid[1] <= 65535, GC_EINVAL, "gc_pillar_canvas: bad sizes");
    const VoxelWorkspace w = carve_workspace(const_cast<void *>(workspace), *geom, n_agents, total_points);
    FusedSrc src;
    src.points = (const float4 *)points;
    src.point_offsets = point_offsets;
    src.cell_code = w.cell_code;
    src.slots = w.slots;
    src.pfn = pfn;
    src.max_voxels = g.max_voxels;
    src.vx = g.voxel[0];
    src.vy = g.voxel[1];
    src.ox = centre_offset[0];
    src.oy = centre_offset[1];
    src.cz = 0.0f * g.voxel[2] + centre_offset[2];
    k_canvas<FusedSrc><<<dim3((g.grid[0] + kTileX - 1) / kTileX, g.grid[1], n_agents), 256, 0,
                         (cudaStream_t)stream>>>(src, g.grid[0], g.grid[1], GC_PFN_OUT, canvas);
    GC_LAUNCH_CHECK("k_canvas<FusedSrc>");
    return GC_OK;
}
