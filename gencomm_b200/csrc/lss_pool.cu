// LSS voxel pooling ("splat"): LiftSplatShoot.voxel_pooling (opencood/models/heter_encoders.py:161-217)
// -- the camera agents' counterpart of the pillar scatter.  The reference flattens the B*N*D*H*W frustum points, computes
// an integer voxel per point, filters, argsorts by voxel rank, runs the "cumsum trick" (utils/camera_utils.py:209-217) and
// index_puts the per-voxel sums into a [B, C, Z, Y, X] grid that is then concatenated over Z.  Here one pass does it:
// a warp per frustum point computes the voxel with the reference's fp32 arithmetic (sub, div, truncation toward zero like
// .long()) and adds its C-vector into out[b][z*C + c][y][x] with fp32 reductions in L2 (red.global.add.f32) -- no sort, no
// prefix sum, no [Nprime, C] gather copies.  HBM-bound: Nprime*C*4 bytes read once, the grid written once (memset) plus the
// touched cells.  Summation order differs from the reference (whose cumsum differences carry ~1e-7 * |running sum| of
// cancellation noise): tolerance-bounded, indices exact.
#include "common.cuh"

namespace gc {

__global__ void __launch_bounds__(256)
k_lss_splat(const float *__restrict__ geom, const float *__restrict__ x, long long n_points, long long per_batch, int C,
            float3 lo /* bx - dx/2 (fp32, computed like the reference) */, float3 dx, int3 nx, float *__restrict__ out) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n_points) return;
    const float gx = __ldg(geom + 3 * p), gy = __ldg(geom + 3 * p + 1), gz = __ldg(geom + 3 * p + 2);
    // ((geom - (bx - dx/2)) / dx).long(): fp32 subtract, fp32 divide, truncation toward zero
    const float fx = __fdiv_rn(__fsub_rn(gx, lo.x), dx.x), fy = __fdiv_rn(__fsub_rn(gy, lo.y), dx.y), fz = __fdiv_rn(__fsub_rn(gz, lo.z), dx.z);
    if (!(fx > -1.0f && fx < (float)nx.x && fy > -1.0f && fy < (float)nx.y && fz > -1.0f && fz < (float)nx.z)) return;  // also NaN
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;          // trunc: (-1, 0) -> 0 like the reference
    if (ix < 0 || ix >= nx.x || iy < 0 || iy >= nx.y || iz < 0 || iz >= nx.z) return;
    const long long b = p / per_batch;
    const size_t plane = (size_t)nx.y * nx.x;
    float *dst = out + ((size_t)b * nx.z * C + (size_t)iz * C) * plane + (size_t)iy * nx.x + ix;
    const float *src = x + (size_t)p * C;
    for (int c = lane; c < C; c += 32) {
        const float v = __ldg(src + c);
        if (v != 0.0f) atomicAdd(dst + (size_t)c * plane, v);     // result unused -> RED.E.ADD.F32
    }
}


// Vector variant (C % 4 == 0, workspace given): the grid is accumulated channel-last ([B][nz][ny][nx][C]) so that a point
// issues C/4 128-bit reductions (red.global.add.v4.f32) instead of C scalar ones -- the scalar kernel is bound by the L2
// reduction rate (151 M RED per call at the OPV2V camera shape: 0.71 ms whatever the collision rate) -- and a tiled
// transpose writes the NCHW result.
__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256)
k_lss_splat_cl(const float *__restrict__ geom, const float4 *__restrict__ x, long long n_points, long long per_batch, int C4,
               float3 lo, float3 dx, int3 nx, float *__restrict__ grid_cl) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);     // 16 lanes per point
    const int l16 = threadIdx.x & 15;
    if (p >= n_points) return;
    const float gx = __ldg(geom + 3 * p), gy = __ldg(geom + 3 * p + 1), gz = __ldg(geom + 3 * p + 2);
    const float fx = __fdiv_rn(__fsub_rn(gx, lo.x), dx.x), fy = __fdiv_rn(__fsub_rn(gy, lo.y), dx.y), fz = __fdiv_rn(__fsub_rn(gz, lo.z), dx.z);
    if (!(fx > -1.0f && fx < (float)nx.x && fy > -1.0f && fy < (float)nx.y && fz > -1.0f && fz < (float)nx.z)) return;
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    if (ix < 0 || ix >= nx.x || iy < 0 || iy >= nx.y || iz < 0 || iz >= nx.z) return;
    const long long b = p / per_batch;
    const size_t cell = (((size_t)b * nx.z + iz) * nx.y + iy) * nx.x + ix;
    float *dst = grid_cl + cell * (size_t)(4 * C4);
    const float4 *src = x + (size_t)p * C4;
    for (int c = l16; c < C4; c += 16) red_add_v4(dst + 4 * c, __ldg(src + c));
}

// grid_cl [R = B*nz*ny][nx][C]  ->  out [B][nz*C][ny][nx]; one CTA per (row r, 32-cell x 32-channel tile)
__global__ void __launch_bounds__(256)
k_lss_to_nchw(const float *__restrict__ grid_cl, int C, int nz, int ny, int nxx, float *__restrict__ out) {
    __shared__ float t[32][33];
    const int r = blockIdx.z, x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int xx = x0 + k, c = c0 + tx;
        t[k][tx] = (xx < nxx && c < C) ? __ldg(grid_cl + ((size_t)r * nxx + xx) * C + c) : 0.0f;
    }
    __syncthreads();
    const int y = r % ny, z = (r / ny) % nz, b = r / (ny * nz);
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, xx = x0 + tx;
        if (c < C && xx < nxx) out[(((size_t)b * nz * C + (size_t)z * C + c) * ny + y) * nxx + xx] = t[tx][k];
    }
}

// Deterministic variant: the reference's sort + cumsum is deterministic, fp32 reductions in L2 are not (their order varies
// run to run).  Here every addend is converted to 40.24 fixed point and accumulated with 64-bit integer reductions, which
// commute exactly: bit-identical results on every run, each addend rounded to 2^-24 (the fp32 spacing at magnitude 1).
constexpr float kFixScale = 16777216.0f;          // 2^24
__global__ void __launch_bounds__(256)
k_lss_splat_det(const float *__restrict__ geom, const float *__restrict__ x, long long n_points, long long per_batch, int C,
                float3 lo, float3 dx, int3 nx, unsigned long long *__restrict__ grid_cl) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);     // 16 lanes per point
    const int l16 = threadIdx.x & 15;
    if (p >= n_points) return;
    const float gx = __ldg(geom + 3 * p), gy = __ldg(geom + 3 * p + 1), gz = __ldg(geom + 3 * p + 2);
    const float fx = __fdiv_rn(__fsub_rn(gx, lo.x), dx.x), fy = __fdiv_rn(__fsub_rn(gy, lo.y), dx.y), fz = __fdiv_rn(__fsub_rn(gz, lo.z), dx.z);
    if (!(fx > -1.0f && fx < (float)nx.x && fy > -1.0f && fy < (float)nx.y && fz > -1.0f && fz < (float)nx.z)) return;
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    if (ix < 0 || ix >= nx.x || iy < 0 || iy >= nx.y || iz < 0 || iz >= nx.z) return;
    const long long b = p / per_batch;
    const size_t cell = (((size_t)b * nx.z + iz) * nx.y + iy) * nx.x + ix;
    unsigned long long *dst = grid_cl + cell * (size_t)C;
    const float *src = x + (size_t)p * C;
    for (int c = l16; c < C; c += 16) {
        const long long q = __float2ll_rn(__fmul_rn(__ldg(src + c), kFixScale));   // power-of-two scale: exact
        if (q != 0) atomicAdd(dst + c, (unsigned long long)q);      // two's complement: signed sums wrap correctly
    }
}

__global__ void __launch_bounds__(256)
k_lss_det_to_nchw(const long long *__restrict__ grid_cl, int C, int nz, int ny, int nxx, float *__restrict__ out) {
    __shared__ float t[32][33];
    const int r = blockIdx.z, x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int xx = x0 + k, c = c0 + tx;
        t[k][tx] = (xx < nxx && c < C) ? (float)((double)grid_cl[((size_t)r * nxx + xx) * C + c] * (1.0 / (double)kFixScale)) : 0.0f;
    }
    __syncthreads();
    const int y = r % ny, z = (r / ny) % nz, b = r / (ny * nz);
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, xx = x0 + tx;
        if (c < C && xx < nxx) out[(((size_t)b * nz * C + (size_t)z * C + c) * ny + y) * nxx + xx] = t[tx][k];
    }
}

}  // namespace gc

using namespace gc;

extern "C" size_t gc_lss_pool_det_workspace_bytes(int n_batch, int C, const int *nx) {
    if (n_batch <= 0 || C <= 0 || !nx || nx[0] <= 0 || nx[1] <= 0 || nx[2] <= 0) return 0;
    return (size_t)n_batch * nx[2] * nx[1] * nx[0] * C * sizeof(long long);
}

extern "C" int gc_lss_voxel_pooling_det(const float *geom, const float *x, long long n_points, int n_batch, int C, const float *dx,
                                        const float *bx, const int *nx, void *workspace, float *out, void *stream) {
    GC_REQUIRE(n_points >= 0 && n_batch > 0 && C > 0 && dx && bx && nx && out && workspace, GC_EINVAL,
               "gc_lss_voxel_pooling_det: bad arguments");
    GC_REQUIRE(n_points % n_batch == 0, GC_EINVAL, "gc_lss_voxel_pooling_det: points must split evenly over the batch");
    GC_REQUIRE(nx[0] > 0 && nx[1] > 0 && nx[2] > 0 && (size_t)n_batch * nx[2] * nx[1] <= 65535, GC_EINVAL,
               "gc_lss_voxel_pooling_det: bad grid");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, gc_lss_pool_det_workspace_bytes(n_batch, C, nx), st);
    GC_REQUIRE(e == cudaSuccess, (int)e, "gc_lss_voxel_pooling_det: memset: %s", cudaGetErrorString(e));
    float3 lo, d;
    lo.x = bx[0] - dx[0] / 2.0f; lo.y = bx[1] - dx[1] / 2.0f; lo.z = bx[2] - dx[2] / 2.0f;
    d.x = dx[0]; d.y = dx[1]; d.z = dx[2];
    const int3 n3 = make_int3(nx[0], nx[1], nx[2]);
    if (n_points > 0) {
        GC_REQUIRE(geom && x, GC_EINVAL, "gc_lss_voxel_pooling_det: null pointer");
        const long long blocks = (n_points + 15) / 16;
        GC_REQUIRE(blocks < (1ll << 31), GC_EUNSUPPORTED, "gc_lss_voxel_pooling_det: too many points");
        k_lss_splat_det<<<(unsigned)blocks, 256, 0, st>>>(geom, x, n_points, n_points / n_batch, C, lo, d, n3,
                                                         (unsigned long long *)workspace);
        GC_LAUNCH_CHECK("k_lss_splat_det");
    }
    k_lss_det_to_nchw<<<dim3((nx[0] + 31) / 32, (C + 31) / 32, n_batch * nx[2] * nx[1]), 256, 0, st>>>((const long long *)workspace, C,
                                                                                                   nx[2], nx[1], nx[0], out);
    GC_LAUNCH_CHECK("k_lss_det_to_nchw");
    return GC_OK;
}

extern "C" size_t gc_lss_pool_workspace_bytes(int n_batch, int C, const int *nx) {
    if (n_batch <= 0 || C <= 0 || C % 4 != 0 || !nx || nx[0] <= 0 || nx[1] <= 0 || nx[2] <= 0) return 0;
    return (size_t)n_batch * nx[2] * nx[1] * nx[0] * C * sizeof(float);
}

extern "C" int gc_lss_voxel_pooling(const float *geom, const float *x, long long n_points, int n_batch, int C, const float *dx,
                                    const float *bx, const int *nx, void *workspace, float *out, void *stream) {
    GC_REQUIRE(n_points >= 0 && n_batch > 0 && C > 0 && dx && bx && nx && out, GC_EINVAL, "gc_lss_voxel_pooling: bad arguments");
    GC_REQUIRE(n_points % n_batch == 0, GC_EINVAL, "gc_lss_voxel_pooling: points must split evenly over the batch (B*N*D*H*W)");
    GC_REQUIRE(nx[0] > 0 && nx[1] > 0 && nx[2] > 0, GC_EINVAL, "gc_lss_voxel_pooling: bad grid");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t out_bytes = (size_t)n_batch * nx[2] * C * nx[1] * nx[0] * sizeof(float);
    const bool vec = workspace != nullptr && C % 4 == 0 && n_points > 0 && (size_t)n_batch * nx[2] * nx[1] <= 65535;
    cudaError_t e = cudaMemsetAsync(vec ? workspace : (void *)out, 0, out_bytes, st);
    GC_REQUIRE(e == cudaSuccess, (int)e, "gc_lss_voxel_pooling: memset: %s", cudaGetErrorString(e));
    if (n_points == 0) return GC_OK;
    GC_REQUIRE(geom && x, GC_EINVAL, "gc_lss_voxel_pooling: null pointer");
    // bx - dx / 2. in fp32, the reference's tensor arithmetic (heter_encoders.py:174)
    float3 lo, d;
    lo.x = bx[0] - dx[0] / 2.0f; lo.y = bx[1] - dx[1] / 2.0f; lo.z = bx[2] - dx[2] / 2.0f;
    d.x = dx[0]; d.y = dx[1]; d.z = dx[2];
    const int3 n3 = make_int3(nx[0], nx[1], nx[2]);
    if (vec) {
        const long long blocks = (n_points + 15) / 16;
        GC_REQUIRE(blocks < (1ll << 31), GC_EUNSUPPORTED, "gc_lss_voxel_pooling: too many points");
        k_lss_splat_cl<<<(unsigned)blocks, 256, 0, st>>>(geom, (const float4 *)x, n_points, n_points / n_batch, C / 4, lo, d, n3,
                                                        (float *)workspace);
        GC_LAUNCH_CHECK("k_lss_splat_cl");
        k_lss_to_nchw<<<dim3((nx[0] + 31) / 32, (C + 31) / 32, n_batch * nx[2] * nx[1]), 256, 0, st>>>((const float *)workspace, C, nx[2],
                                                                                                   nx[1], nx[0], out);
        GC_LAUNCH_CHECK("k_lss_to_nchw");
        return GC_OK;
    }
    const long long blocks = (n_points + 7) / 8;
    GC_REQUIRE(blocks < (1ll << 31), GC_EUNSUPPORTED, "gc_lss_voxel_pooling: too many points");
    k_lss_splat<<<(unsigned)blocks, 256, 0, st>>>(geom, x, n_points, n_points / n_batch, C, lo, d, n3, out);
    GC_LAUNCH_CHECK("k_lss_splat");
    return GC_OK;
}
