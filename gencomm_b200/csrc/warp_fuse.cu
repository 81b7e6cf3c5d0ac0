// Inter-agent BEV warp to the ego frame fused with regroup and Max / Att fusion, for sm_100a.
//
// Replaces (paths relative to /root/reference/opencood):
//   models/sub_modules/torch_transformation_utils.py:323-332  warp_affine_simple
//        = F.affine_grid(theta_f64, align_corners=False).to(f32) + F.grid_sample(bilinear, zeros)
//   models/fuse_modules/fusion_in_one.py:48-51 regroup, :91-124 MaxFusion, :131-151 AttFusion,
//        :41-45 ScaledDotProductAttention, :53-85 warp_feature
//   utils/transformation_utils.py:68-92 normalize_pairwise_tfm
//
// Output-stationary: one thread owns one output pixel of one frame, walks the channels and, per
// channel, gathers the 4 bilinear taps of each of the frame's N agents and reduces across agents in
// registers (max, or the ego row of the per-pixel N x N attention).  Nothing is materialised: no
// grid tensor, no warped copy, no [HW,N,C] permute, no D2H sync for regroup.  The kernel is bound
// by HBM/L2 gather bandwidth (< 1 flop/byte); see DESIGN.md section 4.
//
// Numerics: the base grid and the 2x3 affine are evaluated in float64 and
// only the resulting grid coordinate is rounded to float32 -- exactly what
// F.affine_grid(theta_f64).to(src) does; computing them in float32 moves results by up to 5e-5.
// Unnormalisation, floor, the four weights and the tap accumulation order follow ATen's
// grid_sampler_2d CUDA kernel (align_corners=False, padding zeros).
#include <stdlib.h>

#include "warp_common.cuh"

namespace gc {

// warp_fuse_tile.cu: 0 = launched, 1 = not eligible (use the gather kernels below), otherwise an error code
int warp_fuse_tile(const float *feat, const int32_t *agent_offsets, int n_frames, int total_agents,
                   const double *theta, int L, int C, int H, int W, int mode, int nmax, float *out, cudaStream_t st);
// warp_fuse_persist.cu (persistent CTAs, producer runs ahead across tiles): same contract
int warp_fuse_persist(const float *feat, const int32_t *agent_offsets, int n_frames, int total_agents,
                      const double *theta, int L, int C, int H, int W, int mode, int nmax, float *out, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Generic fused kernel: grid (ceil(W/32), ceil(H/8), n_frames), block 32x8, one pixel per thread.
// NMAX is the compile-time bound on agents per frame (register arrays); the actual N of the frame
// is read from agent_offsets, so ragged record_len is handled without a host sync.
// MODE 1: max over agents.  MODE 2: two passes over the channels (scores, then weighted sum); the
// second pass re-reads the taps through L1/L2.
// ------------------------------------------------------------------------------------------------
template <int MODE, int NMAX>
__global__ void __launch_bounds__(256)
k_warp_fuse(const float *__restrict__ feat, const int32_t *__restrict__ agent_offsets,
            const double *__restrict__ theta, int L, int C, int H, int W, float inv_unused, float sqrt_c,
            float *__restrict__ out) {
    const int w = blockIdx.x * 32 + threadIdx.x;
    const int h = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= W || h >= H) return;
    const int a0 = __ldg(agent_offsets + b);
    int n = __ldg(agent_offsets + b + 1) - a0;
    n = n > NMAX ? NMAX : n;
    const size_t plane = (size_t)H * W;
    const double xs = base_coord(w, W), ys = base_coord(h, H);

    Tap tap[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        if (j < n) {
            tap[j] = make_tap(theta + ((size_t)b * L * L + j) * 6, xs, ys, H, W);   // row [b][0][j]
        } else {
            tap[j].valid = 0; tap[j].off = 0;
            tap[j].w_nw = tap[j].w_ne = tap[j].w_sw = tap[j].w_se = 0.0f;
        }
    }
    const float *src = feat + (size_t)a0 * C * plane;
    float *dst = out + (size_t)b * C * plane + (size_t)h * W + w;

    if (MODE == GC_FUSE_MAX) {
        for (int c = 0; c < C; ++c) {
            float m = 0.0f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < n) {
                    const float v = sample(src + ((size_t)j * C + c) * plane, tap[j], W);
                    m = (j == 0) ? v : fmaxf(m, v);
                }
            }
            dst[(size_t)c * plane] = m;
        }
    } else {
        // pass 1: s_j = <w_0, w_j>
        float s[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) s[j] = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float v0 = sample(src + (size_t)c * plane, tap[0], W);
            s[0] = __fmaf_rn(v0, v0, s[0]);
#pragma unroll
            for (int j = 1; j < NMAX; ++j) {
                if (j < n) s[j] = __fmaf_rn(v0, sample(src + ((size_t)j * C + c) * plane, tap[j], W), s[j]);
            }
        }
        // softmax over the n scores (score / sqrt(C), fusion_in_one.py:42-43)
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) { s[j] = __fdiv_rn(s[j], sqrt_c); mx = fmaxf(mx, s[j]); }
        }
        float den = 0.0f;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) { s[j] = expf(s[j] - mx); den += s[j]; }
        }
#pragma unroll
        for (int j = 0; j < NMAX; ++j) s[j] = (j < n) ? __fdiv_rn(s[j], den) : 0.0f;
        // pass 2: out = sum_j a_j w_j
        for (int c = 0; c < C; ++c) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < n) acc = __fmaf_rn(s[j], sample(src + ((size_t)j * C + c) * plane, tap[j], W), acc);
            }
            dst[(size_t)c * plane] = acc;
        }
    }
}

// warp only (warp_affine_simple / warp_feature): grid z = agent.
__global__ void __launch_bounds__(256)
k_warp_only(const float *__restrict__ feat, const int32_t *__restrict__ agent_offsets, int n_frames,
            const double *__restrict__ theta, int L, int C, int H, int W, float *__restrict__ out) {
    const int w = blockIdx.x * 32 + threadIdx.x;
    const int h = blockIdx.y * 8 + threadIdx.y;
    const int a = blockIdx.z;
    if (w >= W || h >= H) return;
    int b = 0;   // frame of agent a: largest b with agent_offsets[b] <= a
    {
        int lo = 0, hi = n_frames;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(agent_offsets + mid) <= a) lo = mid; else hi = mid;
        }
        b = lo;
    }
    const int j = a - __ldg(agent_offsets + b);
    const size_t plane = (size_t)H * W;
    float *dst = out + (size_t)a * C * plane + (size_t)h * W + w;
    if (j >= L) {   // cannot happen for a well-formed record_len; define the result anyway
        for (int c = 0; c < C; ++c) dst[(size_t)c * plane] = 0.0f;
        return;
    }
    const Tap t = make_tap(theta + ((size_t)b * L * L + j) * 6, base_coord(w, W), base_coord(h, H), H, W);
    const float *src = feat + (size_t)a * C * plane;
    for (int c = 0; c < C; ++c) dst[(size_t)c * plane] = sample(src + (size_t)c * plane, t, W);
}

// normalize_pairwise_tfm (transformation_utils.py:86-90), float64, one thread per matrix.
__global__ void k_normalize_tfm(const double *__restrict__ pw, int n, double H, double W, double denom_x,
                                double denom_y, double *__restrict__ th) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *m = pw + (size_t)i * 16;
    double *o = th + (size_t)i * 6;
    o[0] = m[0];
    o[1] = __ddiv_rn(__dmul_rn(m[1], H), W);
    o[2] = __dmul_rn(__ddiv_rn(m[3], denom_x), 2.0);
    o[3] = __ddiv_rn(__dmul_rn(m[4], W), H);
    o[4] = m[5];
    o[5] = __dmul_rn(__ddiv_rn(m[7], denom_y), 2.0);
}

template <int MODE>
static int launch_fuse(int nmax, dim3 grid, dim3 block, cudaStream_t st, const float *feat, const int32_t *off,
                       const double *theta, int L, int C, int H, int W, float sqrt_c, float *out) {
#define GC_CASE(N)                                                                                         \
    case N:                                                                                                \
        k_warp_fuse<MODE, N><<<grid, block, 0, st>>>(feat, off, theta, L, C, H, W, 0.0f, sqrt_c, out);      \
        break;
    switch (nmax) {
        GC_CASE(1) GC_CASE(2) GC_CASE(3) GC_CASE(4) GC_CASE(5) GC_CASE(6) GC_CASE(7) GC_CASE(8)
        default:
            return GC_EUNSUPPORTED;
    }
#undef GC_CASE
    return GC_OK;
}

}  // namespace gc

using namespace gc;

extern "C" int gc_warp_fuse(const float *feat, const int32_t *agent_offsets, int n_frames, int total_agents,
                            int max_agents_per_frame, const double *theta, int L, int C, int H, int W, int mode, float *out, void *stream) {
    GC_REQUIRE(n_frames >= 0 && total_agents >= 0 && L > 0 && C > 0 && H > 0 && W > 0, GC_EINVAL,
               "gc_warp_fuse: bad sizes");
    GC_REQUIRE(mode == GC_FUSE_WARP_ONLY || mode == GC_FUSE_MAX || mode == GC_FUSE_ATT, GC_EINVAL,
               "gc_warp_fuse: unknown mode %d", mode);
    if (n_frames == 0 || total_agents == 0) return GC_OK;
    GC_REQUIRE(feat && agent_offsets && theta && out, GC_EINVAL, "gc_warp_fuse: null pointer");
    GC_REQUIRE((long long)H * W < (1ll << 31), GC_EUNSUPPORTED, "gc_warp_fuse: plane too large");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 block(32, 8);
    const unsigned gx = (W + 31) / 32, gy = (H + 7) / 8;
    GC_REQUIRE(gy <= 65535 && n_frames <= 65535 && total_agents <= 65535, GC_EUNSUPPORTED,
               "gc_warp_fuse: grid too large");
    // compile-time agent bound: every frame has >= 1 agent and at most L
    int nmax = total_agents - (n_frames - 1);
    if (nmax > L) nmax = L;
    if (max_agents_per_frame > 0 && max_agents_per_frame < nmax) nmax = max_agents_per_frame;   // caller's bound
    if (mode != GC_FUSE_WARP_ONLY)
        GC_REQUIRE(nmax >= 1 && nmax <= kMaxN, GC_EUNSUPPORTED,
                   "gc_warp_fuse: up to %d agents per frame supported (bound %d)", kMaxN, nmax);
    // fast path: TMA-staged tiles (warp_fuse_tile.cu); GC_WARP_FUSE_GATHER=1 forces the gather kernels
    const char *env = getenv("GC_WARP_FUSE_GATHER");
    const bool force_gather = env && env[0] == '1';
    if (!force_gather) {
        const char *impl = getenv("GC_FUSE_IMPL");   // "tile": the round-1b per-tile kernel (kept for A/B measurements)
        const int rc = (impl && impl[0] == 't')
            ? warp_fuse_tile(feat, agent_offsets, n_frames, total_agents, theta, L, C, H, W, mode, nmax, out, st)
            : warp_fuse_persist(feat, agent_offsets, n_frames, total_agents, theta, L, C, H, W, mode, nmax, out, st);
        if (rc != 1) return rc;
    }
    if (mode == GC_FUSE_WARP_ONLY) {
        k_warp_only<<<dim3(gx, gy, total_agents), block, 0, st>>>(feat, agent_offsets, n_frames, theta, L, C, H, W, out);
        GC_LAUNCH_CHECK("k_warp_only");
        return GC_OK;
    }
    const float sqrt_c = (float)sqrt((double)C);   // np.sqrt(dim) cast to the tensor dtype
    const dim3 grid(gx, gy, n_frames);
    int rc = (mode == GC_FUSE_MAX)
                 ? launch_fuse<GC_FUSE_MAX>(nmax, grid, block, st, feat, agent_offsets, theta, L, C, H, W, sqrt_c, out)
                 : launch_fuse<GC_FUSE_ATT>(nmax, grid, block, st, feat, agent_offsets, theta, L, C, H, W, sqrt_c, out);
    GC_REQUIRE(rc == GC_OK, rc, "gc_warp_fuse: unsupported agent bound %d", nmax);
    GC_LAUNCH_CHECK("k_warp_fuse");
    return GC_OK;
}

extern "C" int gc_normalize_pairwise_tfm(const double *pairwise, int n, double H, double W, double discrete_ratio,
                                         double downsample_rate, double *theta, void *stream) {
    GC_REQUIRE(n >= 0, GC_EINVAL, "gc_normalize_pairwise_tfm: negative count");
    if (n == 0) return GC_OK;
    GC_REQUIRE(pairwise && theta, GC_EINVAL, "gc_normalize_pairwise_tfm: null pointer");
    // python evaluates (downsample_rate * discrete_ratio * W) left to right in float64
    const double dx = downsample_rate * discrete_ratio * W;
    const double dy = downsample_rate * discrete_ratio * H;
    k_normalize_tfm<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pairwise, n, H, W, dx, dy, theta);
    GC_LAUNCH_CHECK("k_normalize_tfm");
    return GC_OK;
}
