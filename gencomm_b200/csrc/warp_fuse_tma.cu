// TMA-staged warp + regroup + Max/Att fusion for sm_100a (the fast path of gc_warp_fuse).
//
// Same operator and numerics as warp_fuse.cu (reference: fusion_in_one.py:91-151,
// torch_transformation_utils.py:323-332); different data movement.  ncu on the gather version showed it
// bound by L1 tag/wavefront throughput: a warp of 32 output pixels maps onto a rotated line of the source,
// so every one of the 4 tap loads touches up to 32 cache lines.  Here:
//
//   * a CTA owns a 32x32 output tile of one frame (1024 threads, one pixel each, one warp per tile row);
//   * the source footprint of that tile under agent j's affine map is a rotated square that always fits a
//     52x48 box; ONE cp.async.bulk.tensor (TMA) per (agent, channel) copies that box of the NCHW plane into
//     shared memory.  The box origin may be negative / beyond the plane: TMA zero-fills out-of-bounds
//     elements, which is exactly grid_sample's padding_mode='zeros';
//   * boxes flow through a ring of shared-memory stages guarded by full/empty mbarriers (thread 0 is the TMA
//     producer, all 32 warps are consumers); the 4 bilinear taps are LDS from the box, the reduction across
//     agents (max / ego-row attention) stays in registers and each warp stores 128 contiguous bytes.
//
// The tensor map is built on the host per call (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint).
// Requirements: W % 4 == 0 (TMA global strides are multiples of 16 B) and at most kTmaMaxN agents per frame;
// otherwise gc_warp_fuse falls back to the gather kernels.  Affine maps that are not (near-)isometries can
// overflow the box: such an agent is detected per tile and sampled straight from global memory.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "warp_common.cuh"

namespace gc {

constexpr int kTile = 32;
// Footprint of a 32x32 tile under an isometry: 31*sqrt(2) + 3 = 46.9 -> 47 pixels per side.  TMA needs the
// box start 16-byte aligned in global memory, so the x origin is floored to a multiple of 4 floats (up to 3
// extra columns): 52 x 48.
constexpr int kBoxW = 52, kBoxH = 48;
constexpr int kBoxFloats = kBoxW * kBoxH;
constexpr int kBoxBytes = kBoxFloats * 4;        // 9984, a multiple of 128
constexpr int kTmaMaxN = 5;                      // register budget of a 1024-thread CTA (64 regs/thread)
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *map, int x, int y, int plane, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(plane), "r"(smem_u32(bar))
        : "memory");
}

struct TapS {
    float w_nw, w_ne, w_sw, w_se;
    int x0, y0;   // clamped to [-2, W] / [-2, H]: anything clamped has both taps out of bounds (zero)
};

__device__ __forceinline__ TapS make_tap_xy(const double *__restrict__ th, double xs, double ys, int H, int W) {
    const float gx = (float)(xs * th[0] + ys * th[1] + th[2]);
    const float gy = (float)(xs * th[3] + ys * th[4] + th[5]);
    const float ix = __fmaf_rn(gx + 1.0f, (float)W, -1.0f) * 0.5f;
    const float iy = __fmaf_rn(gy + 1.0f, (float)H, -1.0f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const float ex = (fx + 1.0f) - ix, wx = ix - fx;
    const float sy = (fy + 1.0f) - iy, ny_ = iy - fy;
    TapS t;
    t.w_nw = ex * sy; t.w_ne = wx * sy; t.w_sw = ex * ny_; t.w_se = wx * ny_;
    t.x0 = (int)fminf(fmaxf(fx, -2.0f), (float)W);
    t.y0 = (int)fminf(fmaxf(fy, -2.0f), (float)H);
    if (!(ix == ix) || !(iy == iy)) {   // NaN transform: contributes zeros
        t.w_nw = t.w_ne = t.w_sw = t.w_se = 0.0f;
        t.x0 = t.y0 = -2;
    }
    return t;
}

enum AgentPath { kPathTma = 0, kPathZero = 1, kPathGather = 2 };

struct TapR {                // per-thread, per-agent
    float w_nw, w_ne, w_sw, w_se;
    int off;                 // box-relative offset (TMA path) or y0*W+x0 (gather path)
    unsigned valid;          // gather path only
};

// MODE: GC_FUSE_WARP_ONLY (grid.z = agent), GC_FUSE_MAX, GC_FUSE_ATT (grid.z = frame)
template <int MODE, int NMAX>
__global__ void __launch_bounds__(1024, 1)
k_warp_fuse_tma(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ feat,
                const int32_t *__restrict__ agent_offsets, int n_frames, const double *__restrict__ theta, int L,
                int C, int H, int W, float sqrt_c, int stages, int stage_agents, float *__restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    // TMA destinations must be 128-byte aligned; the dynamic segment follows the static variables below
    float *const ring = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    __shared__ int s_bx[NMAX], s_by[NMAX], s_path[NMAX], s_slot[NMAX];
    __shared__ int s_ntma;

    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    const int w0 = blockIdx.x * kTile, h0 = blockIdx.y * kTile;
    const int w = w0 + lane, h = h0 + row;
    const bool active = w < W && h < H;

    int b, a0, n;
    if (MODE == GC_FUSE_WARP_ONLY) {
        const int a = blockIdx.z;
        int lo = 0, hi = n_frames;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(agent_offsets + mid) <= a) lo = mid; else hi = mid;
        }
        b = lo; a0 = a; n = 1;
    } else {
        b = blockIdx.z;
        a0 = __ldg(agent_offsets + b);
        n = min(__ldg(agent_offsets + b + 1) - a0, NMAX);
    }
    const int j0 = MODE == GC_FUSE_WARP_ONLY ? min(blockIdx.z - __ldg(agent_offsets + b), L - 1) : 0;
    const double *th_base = theta + ((size_t)b * L * L + j0) * 6;   // row [b][0][j0 + j]

    // ---- per-agent box of this tile (warp 0, lane j) ------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (row == 0 && lane < NMAX) {
        int path = kPathZero, bx = 0, by = 0;
        if (lane < n) {
            const int w1 = min(w0 + kTile - 1, W - 1), h1 = min(h0 + kTile - 1, H - 1);
            int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // the affine map is linear and rounding monotone: extremes are at corners
                const TapS t = make_tap_xy(th_base + lane * 6, base_coord((k & 1) ? w1 : w0, W),
                                           base_coord((k & 2) ? h1 : h0, H), H, W);
                minx = min(minx, t.x0); maxx = max(maxx, t.x0);
                miny = min(miny, t.y0); maxy = max(maxy, t.y0);
            }
            bx = minx & ~3;   // floor to a multiple of 4 (two's complement): 16-byte aligned box start
            by = miny;
            if (maxx + 1 < 0 || minx >= W || maxy + 1 < 0 || miny >= H) path = kPathZero;          // nothing in view
            else if (maxx - bx + 2 <= kBoxW && maxy - miny + 2 <= kBoxH) path = kPathTma;
            else path = kPathGather;
        }
        s_bx[lane] = bx; s_by[lane] = by; s_path[lane] = path;
        const unsigned tma_mask = __ballot_sync((1u << NMAX) - 1u, path == kPathTma);
        s_slot[lane] = __popc(tma_mask & ((1u << lane) - 1u));
        if (lane == 0) s_ntma = __popc(tma_mask);
    }
    __syncthreads();

    // ---- per-thread taps -----------------------------------------------------------------------------
    TapR tap[NMAX];
    const double xs = base_coord(active ? w : w0, W), ys = base_coord(active ? h : h0, H);
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        tap[j].w_nw = tap[j].w_ne = tap[j].w_sw = tap[j].w_se = 0.0f;
        tap[j].off = 0; tap[j].valid = 0;
        if (j < n && s_path[j] != kPathZero) {
            const TapS t = make_tap_xy(th_base + j * 6, xs, ys, H, W);
            tap[j].w_nw = t.w_nw; tap[j].w_ne = t.w_ne; tap[j].w_sw = t.w_sw; tap[j].w_se = t.w_se;
            if (s_path[j] == kPathTma) {
                tap[j].off = s_slot[j] * kBoxFloats + (t.y0 - s_by[j]) * kBoxW + (t.x0 - s_bx[j]);
            } else {
                const bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
                const bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
                tap[j].valid = ((xin0 && yin0) ? 1u : 0u) | ((xin1 && yin0) ? 2u : 0u) | ((xin0 && yin1) ? 4u : 0u) |
                               ((xin1 && yin1) ? 8u : 0u);
                tap[j].off = t.y0 * W + t.x0;
            }
        }
    }
    const int n_tma = s_ntma;
    const size_t plane = (size_t)H * W;
    const int passes = MODE == GC_FUSE_ATT ? 2 : 1;
    const int total = C * passes;
    const size_t stage_floats = (size_t)stage_agents * kBoxFloats;

    // producer (thread 0 only): arm the stage's full barrier and launch one box copy per TMA agent
#define GC_ISSUE(IT)                                                                                          \
    do {                                                                                                      \
        const int it_ = (IT), s_ = it_ % stages, c_ = it_ % C;                                                \
        if (it_ >= stages) mbar_wait(&empty_bar[s_], ((it_ / stages) - 1) & 1);                               \
        mbar_expect_tx(&full_bar[s_], (uint32_t)n_tma * kBoxBytes);                                           \
        float *dst_ = ring + (size_t)s_ * stage_floats;                                                       \
        for (int j_ = 0; j_ < n; ++j_) {                                                                      \
            if (s_path[j_] == kPathTma)                                                                       \
                tma_load_box(dst_ + s_slot[j_] * kBoxFloats, &tmap, s_bx[j_], s_by[j_], (a0 + j_) * C + c_, \
                             &full_bar[s_]);                                                                  \
        }                                                                                                     \
    } while (0)
    if (threadIdx.x == 0 && n_tma > 0) {
        for (int it = 0; it < stages - 1 && it < total; ++it) GC_ISSUE(it);
    }

    const float *src = feat + (size_t)a0 * C * plane;
    float *dst = out + ((size_t)(MODE == GC_FUSE_WARP_ONLY ? a0 : b) * C) * plane + (size_t)(active ? h : 0) * W +
                 (active ? w : 0);
    float score[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) score[j] = 0.0f;

    for (int it = 0; it < total; ++it) {
        const int s = it % stages, c = it % C;
        if (n_tma > 0) {
            if (threadIdx.x == 0 && it + stages - 1 < total) GC_ISSUE(it + stages - 1);
            mbar_wait(&full_bar[s], (it / stages) & 1);
        }
        const float *box = ring + (size_t)s * stage_floats;
        float v[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            v[j] = 0.0f;
            if (j < n) {
                const int path = s_path[j];
                if (path == kPathTma) {
                    const float *p = box + tap[j].off;
                    float acc = p[0] * tap[j].w_nw;
                    acc = __fmaf_rn(p[1], tap[j].w_ne, acc);
                    acc = __fmaf_rn(p[kBoxW], tap[j].w_sw, acc);
                    acc = __fmaf_rn(p[kBoxW + 1], tap[j].w_se, acc);
                    v[j] = acc;
                } else if (path == kPathGather && active) {
                    Tap g;
                    g.w_nw = tap[j].w_nw; g.w_ne = tap[j].w_ne; g.w_sw = tap[j].w_sw; g.w_se = tap[j].w_se;
                    g.off = tap[j].off; g.valid = tap[j].valid;
                    v[j] = sample(src + ((size_t)j * C + c) * plane, g, W);
                }
            }
        }
        if (n_tma > 0) {   // this warp is done with the stage
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        if (MODE == GC_FUSE_WARP_ONLY) {
            if (active) dst[(size_t)c * plane] = v[0];
        } else if (MODE == GC_FUSE_MAX) {
            float m = v[0];
#pragma unroll
            for (int j = 1; j < NMAX; ++j) if (j < n) m = fmaxf(m, v[j]);
            if (active) dst[(size_t)c * plane] = m;
        } else if (it < C) {        // pass 1: s_j = <w_0, w_j>
#pragma unroll
            for (int j = 0; j < NMAX; ++j) if (j < n) score[j] = __fmaf_rn(v[0], v[j], score[j]);
            if (it == C - 1) {      // softmax(score / sqrt(C)), fusion_in_one.py:42-43
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NMAX; ++j) if (j < n) { score[j] = __fdiv_rn(score[j], sqrt_c); mx = fmaxf(mx, score[j]); }
                float den = 0.0f;
#pragma unroll
                for (int j = 0; j < NMAX; ++j) if (j < n) { score[j] = expf(score[j] - mx); den += score[j]; }
#pragma unroll
                for (int j = 0; j < NMAX; ++j) score[j] = (j < n) ? __fdiv_rn(score[j], den) : 0.0f;
            }
        } else {                    // pass 2: out = sum_j a_j w_j
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) if (j < n) acc = __fmaf_rn(score[j], v[j], acc);
            if (active) dst[(size_t)c * plane] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
        (void)cudaGetLastError();
    }
    return fn;
}

template <int MODE, int NMAX>
static int launch_tma(const CUtensorMap &map, dim3 grid, size_t smem, cudaStream_t st, const float *feat,
                      const int32_t *off, int n_frames, const double *theta, int L, int C, int H, int W, float sqrt_c,
                      int stages, int stage_agents, float *out) {
    auto kern = k_warp_fuse_tma<MODE, NMAX>;
    static bool configured = false;   // one attribute call per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return (int)e; }
        configured = true;
    }
    kern<<<grid, 1024, smem, st>>>(map, feat, off, n_frames, theta, L, C, H, W, sqrt_c, stages, stage_agents, out);
    return GC_OK;
}

// Returns GC_OK when the TMA path was launched, 1 when the configuration is not eligible (caller falls
// back to the gather kernels), or an error code.
int warp_fuse_tma(const float *feat, const int32_t *agent_offsets, int n_frames, int total_agents,
                  const double *theta, int L, int C, int H, int W, int mode, int nmax, float *out,
                  cudaStream_t st) {
    if ((W & 3) != 0 || ((uintptr_t)feat & 15) != 0) return 1;
    if (mode != GC_FUSE_WARP_ONLY && nmax > kTmaMaxN) return 1;
    if ((long long)total_agents * C >= (1ll << 31)) return 1;
    PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
    if (!encode) return 1;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)total_agents * C};
    const cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    const cuuint32_t box[3] = {kBoxW, kBoxH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(feat), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;

    const int stage_agents = mode == GC_FUSE_WARP_ONLY ? 1 : nmax;
    int stages = (int)((225 * 1024) / ((size_t)stage_agents * kBoxBytes));
    stages = stages > kMaxStages ? kMaxStages : stages;
    if (stages < 2) return 1;
    const size_t smem = (size_t)stages * stage_agents * kBoxBytes + 128;   // + manual 128-B alignment slack
    const dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, mode == GC_FUSE_WARP_ONLY ? total_agents : n_frames);
    if (grid.y > 65535 || grid.z > 65535) return 1;
    const float sqrt_c = (float)sqrt((double)C);
    int rc;
#define GC_TMA_CASE(M, N)                                                                                             \
    rc = launch_tma<M, N>(map, grid, smem, st, feat, agent_offsets, n_frames, theta, L, C, H, W, sqrt_c, stages,      \
                          stage_agents, out);
    if (mode == GC_FUSE_WARP_ONLY) { GC_TMA_CASE(GC_FUSE_WARP_ONLY, 1) }
    else if (mode == GC_FUSE_MAX) {
        switch (nmax) {
            case 1: GC_TMA_CASE(GC_FUSE_MAX, 1) break;
            case 2: GC_TMA_CASE(GC_FUSE_MAX, 2) break;
            case 3: GC_TMA_CASE(GC_FUSE_MAX, 3) break;
            case 4: GC_TMA_CASE(GC_FUSE_MAX, 4) break;
            default: GC_TMA_CASE(GC_FUSE_MAX, 5) break;
        }
    } else {
        switch (nmax) {
            case 1: GC_TMA_CASE(GC_FUSE_ATT, 1) break;
            case 2: GC_TMA_CASE(GC_FUSE_ATT, 2) break;
            case 3: GC_TMA_CASE(GC_FUSE_ATT, 3) break;
            case 4: GC_TMA_CASE(GC_FUSE_ATT, 4) break;
            default: GC_TMA_CASE(GC_FUSE_ATT, 5) break;
        }
    }
#undef GC_TMA_CASE
    if (rc != GC_OK) { set_error("k_warp_fuse_tma: cudaFuncSetAttribute failed (%d)", rc); return rc; }
    GC_LAUNCH_CHECK("k_warp_fuse_tma");
    return GC_OK;
}

}  // namespace gc
