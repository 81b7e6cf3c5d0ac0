// MessageExtractorv2 for sm_100a: the module that turns an agent's BEV feature into the
// 2-channel message GenComm's sampler is conditioned on.
//
// Replaces (paths relative to /root/reference/opencood):
//   models/gencomm_modules/message_extractor_v2.py:70-120  (BEVDeformableExtractor / MessageExtractorv2)
//     offset1 = Conv2d(C -> 18, 3x3)            :74,  :97
//     dcn1    = torchvision DeformConv2d(C -> 64, 3x3, padding 1)   :78, :101   (deformable im2col + GEMM)
//     attn    = AvgPool -> 1x1 (64 -> 32) -> ReLU -> 1x1 (32 -> 64) -> Sigmoid   :88-94, :108
//     fuse    = 1x1 (64 -> 64) -> ReLU -> 1x1 (64 -> 2)                            :82-86, :111
//
// Both 3x3 convolutions are real dense contractions (K = 9 C = 1152 at C = 128): they run as ONE tcgen05 implicit-GEMM
// kernel template, k_me_conv<NOUT, DEFORM>:
//   * A pre-pass writes the features once as channel-last bf16 (value + residual planes).
//   * CTA = 128 consecutive pixels (M = 128) of one agent, 256 threads.  Per tap the four bilinear corners + weights
//     of every pixel are evaluated once (torchvision's bilinear_interpolate, zero outside the image; the plain
//     convolution is the same code with one corner of weight 1).  Eight lanes then build one pixel's row of the A
//     operand: per corner ONE coalesced 16-byte load of 8 channels each, interpolation in fp32, rounding to bf16,
//     one 16-byte store into K-major no-swizzle core matrices -- the deformable im2col never exists in global memory.
//   * K is walked in stages of one tap x 64 channels (four K = 16 MMAs), double buffered: the gather of stage s+1
//     overlaps the MMAs of stage s; stage reuse is guarded by tcgen05.commit -> mbarrier.
//   * B operand: the weights, pre-packed once to bf16 core-matrix order (gc_me_pack_weights), 8 KB per stage from L2.
//   * D: fp32 accumulators in TMEM (NOUT columns), epilogue tcgen05.ld.32x32b (lane = pixel -> coalesced NCHW stores),
//     bias, and -- for the deformable layer -- the per-tile channel sums the average pool needs (deterministic order).
// The tail (global average pool -> squeeze/excite -> two 1x1 convolutions) is 1 % of the FLOPs: CUDA-core fp32 kernels,
// the excitation folded into the first 1x1's weights per agent (two threads per pixel with 32 accumulators each were
// measured slower: 117 -> 167 us, shared-memory weight reads lose their warp-wide broadcast).
//
// Arithmetic: bf16 operands, fp32 accumulation for the two 3x3 layers; everything else fp32.  Tolerance vs the fp32
// reference is written in tests/test_message_extractor_gpu.py.
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace gc {
namespace me {

using namespace umma;

constexpr int kPix = 128;            // pixels per CTA = M
constexpr int kThreads = 256;
// Input channels per stage (SC / 16 MMAs of K = 16): 64 for the deformable layer, 32 for the offset layer (its value +
// residual planes double the operand bytes; with 64-channel stages only two CTAs fit an SM and the layer was latency
// bound on its load -> store -> barrier -> MMA chain: 305 us at 28 % issue-active).
// A operand plane: [SC/8 channel groups][128 pixels][8 bf16], the group stride padded by 16 (32) bytes so that the lanes
// that build one pixel store to different bank quads (measured without the pad: 8-way conflicts, 32 wavefronts per
// STS.128, 70 M conflicts per launch).  UMMA no-swizzle K-major: LBO = group stride, SBO = 128 B.
__host__ __device__ constexpr int a_group_bytes(int SC) { return kPix * 16 + (SC == 64 ? 16 : 32); }
__host__ __device__ constexpr int a_plane_bytes(int SC) { return (SC / 8) * a_group_bytes(SC); }
__host__ __device__ constexpr int conv_smem_bytes(int NOUT, bool split, int SC) {
    return 2 * (split ? 2 : 1) * (a_plane_bytes(SC) + SC * NOUT * 2);
}
constexpr int kScOffset = 32, kScDeform = 64;

// ------------------------------------------------------------------------------------------------
// weights [NOUT_real][C][3][3] f32 -> bf16 B operand, per stage (tap, 64-channel chunk):
//   [k8 = SC/8 channel groups][n8 = NOUT/8][8 rows n][8 bf16 k]     (rows >= NOUT_real are zero)
// ------------------------------------------------------------------------------------------------
//   split: every stage is followed by its residual plane  lo = bf16(w - float(bf16(w)))  (the "bf16x3" offset layer)
__device__ __forceinline__ float bf16_residual(float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); }

__global__ void k_me_pack(const float *__restrict__ w, int n_real, int NOUT, int C, int SC, int split,
                          uint4 *__restrict__ out) {
    const int chunks = C / SC, groups = SC / 8;
    const int per_stage = groups * NOUT;       // uint4 (8 k values of one row n) per stage and plane
    const int total = 9 * chunks * per_stage;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int n = i % NOUT;                // i = ((stage * groups + k8) * (NOUT/8) + n8) * 8 + (n % 8)  with n = n8 * 8 + n % 8
    const int k8 = (i / NOUT) % groups;
    const int stage = i / per_stage;
    const int tap = stage / chunks, chunk = stage % chunks;
    uint32_t p[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = chunk * SC + k8 * 8 + 2 * j + e;
            v[e] = n < n_real ? w[((size_t)n * C + c) * 9 + tap] : 0.0f;
        }
        p[j] = pack_bf16(v[0], v[1]);
        q[j] = pack_bf16(bf16_residual(v[0]), bf16_residual(v[1]));
    }
    const int within = i - stage * per_stage;
    if (split) {
        out[(size_t)stage * 2 * per_stage + within] = make_uint4(p[0], p[1], p[2], p[3]);
        out[(size_t)stage * 2 * per_stage + per_stage + within] = make_uint4(q[0], q[1], q[2], q[3]);
    } else {
        out[i] = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// x [A][C][HW] f32 (NCHW)  ->  channel-last bf16 planes  xh = bf16(x),  xl = bf16(x - xh)   [A][HW][C]
// so that a bilinear corner of 8 channels is ONE 16-byte load and the 8 lanes that sample a pixel read 128 contiguous
// bytes.  (Measured before: sampling NCHW fp32 directly made the deformable layer L1-gather bound -- 4 loads per
// (pixel, tap, channel), l1tex 76 % busy, 0.64 ms of the 0.94 ms call.)   grid = (HW/64, C/64, A), 256 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_me_to_nhwc(const float *__restrict__ x, int C, int HW, uint4 *__restrict__ xh, uint4 *__restrict__ xl) {
    __shared__ float t[64][65];
    const int tid = threadIdx.x, p0 = blockIdx.x * 64, c0 = blockIdx.y * 64, a = blockIdx.z;
    const float *src = x + ((size_t)a * C + c0) * HW + p0;
    {
        const int p = tid & 63, c_base = tid >> 6;   // 16 independent loads in flight per thread
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (p0 + p < HW) ? __ldg(src + (size_t)(c_base + 4 * k) * HW + p) : 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) t[c_base + 4 * k][p] = v[k];
    }
    __syncthreads();
    for (int i = tid; i < 64 * 8; i += 256) {
        const int p = i >> 3, g = i & 7;
        if (p0 + p >= HW) continue;
        float v[8], r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = t[g * 8 + k][p]; r[k] = bf16_residual(v[k]); }
        const size_t o = ((size_t)a * HW + p0 + p) * (C / 8) + c0 / 8 + g;
        xh[o] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        xl[o] = make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]), pack_bf16(r[4], r[5]), pack_bf16(r[6], r[7]));
    }
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
// torchvision bilinear_interpolate: val = hh*hw*v1 + hh*lw*v2 + lh*hw*v3 + lh*lw*v4 (left to right), two channels
// (packed fp32: one FMUL2 + three FFMA2 for the two channels of a 32-bit word, same roundings as scalar fmaf)
__device__ __forceinline__ uint32_t lerp2(const float4 &w, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    float2 r = __fmul2_rn(make_float2(w.x, w.x), make_float2(bf_lo(a), bf_hi(a)));
    r = __ffma2_rn(make_float2(w.y, w.y), make_float2(bf_lo(b), bf_hi(b)), r);
    r = __ffma2_rn(make_float2(w.z, w.z), make_float2(bf_lo(c), bf_hi(c)), r);
    r = __ffma2_rn(make_float2(w.w, w.w), make_float2(bf_lo(d), bf_hi(d)), r);
    return pack_bf16(r.x, r.y);
}

// ------------------------------------------------------------------------------------------------
// 3x3 (deformable) convolution as a tcgen05 implicit GEMM.  grid = (H*W/128, n_agents), 256 threads.
//   xh, xl [A][HW][C] bf16 (k_me_to_nhwc); offset [A][18][H][W] f32 (DEFORM; channel 2k = dy, 2k+1 = dx of tap k);
//   wp: k_me_pack output;  out [A][n_store][H][W] f32 (+ bias);
//   tile_sums [A][tiles*4][NOUT] (DEFORM): channel sums over 32-pixel groups
// ------------------------------------------------------------------------------------------------
template <int NOUT, bool DEFORM, int SC>
__global__ void __launch_bounds__(kThreads)
k_me_conv(const uint4 *__restrict__ xh, const uint4 *__restrict__ xl, const float *__restrict__ offset,
          const uint4 *__restrict__ wp, const float *__restrict__ bias, int C, int H, int W, int n_store,
          float *__restrict__ out, float *__restrict__ tile_sums) {
    // The plain (offset) layer runs as "bf16x3": A and B are split into a bf16 value and a bf16 residual and three MMAs
    // (hi*hi + lo*hi + hi*lo) rebuild ~16 mantissa bits, because its output positions the deformable layer's taps:
    // a bf16-only offset (rel. error 4e-3) moves a tap by 0.02 px at 5 px, which on high-frequency features costs
    // more accuracy than the deformable layer's own bf16 rounding.
    constexpr bool SPLIT = !DEFORM;
    constexpr int kPlanes = SPLIT ? 2 : 1;
    constexpr int kAGroup = a_group_bytes(SC), kABytes = a_plane_bytes(SC), kGroups = SC / 8;
    constexpr int kBBytes = SC * NOUT * 2 * kPlanes;
    constexpr int kAStage = kABytes * kPlanes;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *a_s = smem;                      // [2][kPlanes][kABytes]
    uint8_t *b_s = smem + 2 * kAStage;        // [2][kPlanes][SC * NOUT * 2]
    __shared__ __align__(8) uint64_t s_empty[2], s_done;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) int4 s_o[kPix];      // pixel index (y*W + x) of the four bilinear corners of the tap
    __shared__ __align__(16) float4 s_w[kPix];    // their weights (0 for a corner outside the image)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int agent = blockIdx.y, tile = blockIdx.x;
    const int HW = H * W;
    const int chunks = C / SC, stages = 9 * chunks;
    const int C8 = C / 8;

    if (warp == 0) tmem_alloc<NOUT>(&s_tmem);
    if (tid == 32) {
        mbar_init(smem_u32(&s_empty[0]), 1); mbar_init(smem_u32(&s_empty[1]), 1); mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s);
    constexpr uint32_t idesc = make_idesc(128, NOUT);
    const uint4 *xh_a = xh + (size_t)agent * HW * C8;
    const uint4 *xl_a = xl + (size_t)agent * HW * C8;

    for (int s = 0; s < stages; ++s) {
        const int b = s & 1;
        const int tap = s / chunks, chunk = s - tap * chunks;
        if (chunk == 0) {
            // sampling position of every pixel of the tile for this tap (torchvision deformable_im2col /
            // bilinear_interpolate), once per tap, shared through s_o / s_w.  Every thread is past its staging of the
            // previous tap (barrier at the end of the previous stage), so the arrays can be overwritten.
            if (tid < kPix) {
                const int pix = tile * kPix + tid;
                const int py = pix / W, px = pix - py * W;
                const int ky = tap / 3, kx = tap - 3 * ky;
                int4 o = make_int4(0, 0, 0, 0);
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (DEFORM) {
                    const float *op = offset + ((size_t)agent * 18 + 2 * tap) * HW + pix;
                    const float hh_ = (float)(py - 1 + ky) + __ldg(op);
                    const float ww_ = (float)(px - 1 + kx) + __ldg(op + HW);
                    const bool inside = hh_ > -1.0f && hh_ < (float)H && ww_ > -1.0f && ww_ < (float)W;
                    const float hf = floorf(hh_), wf = floorf(ww_);
                    const int hl = (int)hf, wl = (int)wf, hh = hl + 1, wh = wl + 1;
                    const float lh = hh_ - hf, lw = ww_ - wf, uh = 1.0f - lh, uw = 1.0f - lw;
                    const bool t_ok = inside && hl >= 0, b_ok = inside && hh <= H - 1;
                    const bool l_ok = wl >= 0, r_ok = wh <= W - 1;
                    const int hlc = min(max(hl, 0), H - 1), hhc = min(max(hh, 0), H - 1);
                    const int wlc = min(max(wl, 0), W - 1), whc = min(max(wh, 0), W - 1);
                    o = make_int4(hlc * W + wlc, hlc * W + whc, hhc * W + wlc, hhc * W + whc);
                    w.x = (t_ok && l_ok) ? uh * uw : 0.0f;
                    w.y = (t_ok && r_ok) ? uh * lw : 0.0f;
                    w.z = (b_ok && l_ok) ? lh * uw : 0.0f;
                    w.w = (b_ok && r_ok) ? lh * lw : 0.0f;
                } else {
                    const int yy = py - 1 + ky, xx = px - 1 + kx;
                    const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
                    o.x = ok ? yy * W + xx : 0;
                    w.x = ok ? 1.0f : 0.0f;
                }
                s_o[tid] = o;
                s_w[tid] = w;
            }
            __syncthreads();
        }
        if (s >= 2) mbar_wait(smem_u32(&s_empty[b]), (uint32_t)((s >> 1) - 1) & 1u);   // MMAs of stage s-2 retired
        // ---- B stage: kBBytes contiguous bytes of the packed weights; the loads are issued here and stored after the
        // first batch of A loads is in flight (one round trip instead of two or three) ----
        constexpr int kBIter = (kBBytes / 16 + kThreads - 1) / kThreads;
        uint4 bq[kBIter];
        {
            const uint4 *src = wp + (size_t)s * (kBBytes / 16);
#pragma unroll
            for (int i = 0; i < kBIter; ++i)
                bq[i] = (i * kThreads + tid < kBBytes / 16) ? __ldg(src + i * kThreads + tid) : make_uint4(0u, 0u, 0u, 0u);
        }
        // ---- A stage: 128 pixels x 8 channel groups = 1024 16-byte rows, four per thread.  The eight lanes of a pixel
        // read 128 contiguous bytes per corner (one full line; spreading a warp over 8 pixels x 4 groups instead was
        // measured 1.9x slower: twice the L1 tags per request) ----
        {
            uint8_t *dst = a_s + b * kAStage;
            const int g_loc = tid & (kGroups - 1);
            const int cg = chunk * kGroups + g_loc;
#pragma unroll
            for (int pass = 0; pass < kGroups / 2; pass += 2) {
                uint4 q[2][4];
                float4 w[2];
                int p[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    p[u] = ((pass + u) * kThreads + tid) / kGroups;
                    const int4 o = s_o[p[u]];
                    w[u] = s_w[p[u]];
                    if (DEFORM) {
                        q[u][0] = __ldg(xh_a + (size_t)o.x * C8 + cg);
                        q[u][1] = __ldg(xh_a + (size_t)o.y * C8 + cg);
                        q[u][2] = __ldg(xh_a + (size_t)o.z * C8 + cg);
                        q[u][3] = __ldg(xh_a + (size_t)o.w * C8 + cg);
                    } else {
                        q[u][0] = __ldg(xh_a + (size_t)o.x * C8 + cg);
                        q[u][1] = __ldg(xl_a + (size_t)o.x * C8 + cg);
                    }
                }
                if (pass == 0) {
                    uint4 *bd = reinterpret_cast<uint4 *>(b_s + b * kBBytes);
#pragma unroll
                    for (int i = 0; i < kBIter; ++i)
                        if (i * kThreads + tid < kBBytes / 16) bd[i * kThreads + tid] = bq[i];
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    uint4 *d = reinterpret_cast<uint4 *>(dst + g_loc * kAGroup) + p[u];
                    if (DEFORM) {
                        d[0] = make_uint4(lerp2(w[u], q[u][0].x, q[u][1].x, q[u][2].x, q[u][3].x),
                                          lerp2(w[u], q[u][0].y, q[u][1].y, q[u][2].y, q[u][3].y),
                                          lerp2(w[u], q[u][0].z, q[u][1].z, q[u][2].z, q[u][3].z),
                                          lerp2(w[u], q[u][0].w, q[u][1].w, q[u][2].w, q[u][3].w));
                    } else {
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        const bool ok = w[u].x != 0.0f;
                        d[0] = ok ? q[u][0] : z;
                        d[kABytes / 16] = ok ? q[u][1] : z;
                    }
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_buf = a_base + (uint32_t)b * kAStage, b_buf = b_base + (uint32_t)b * kBBytes;
            constexpr uint32_t kBPlane = SC * NOUT * 2;
#pragma unroll
            for (int j = 0; j < SC / 16; ++j) {   // K = 16: channel groups 2j and 2j+1
                const uint32_t a_off = (uint32_t)(2 * j) * (uint32_t)kAGroup, b_off = (uint32_t)(2 * j) * (NOUT * 16u);
                const uint64_t a_hi = make_desc(a_buf + a_off, (uint32_t)kAGroup, 128u);
                const uint64_t b_hi = make_desc(b_buf + b_off, NOUT * 16u, 128u);
                mma_bf16(tmem, a_hi, b_hi, idesc, (s > 0 || j > 0) ? 1u : 0u);
                if (SPLIT) {
                    const uint64_t a_lo = make_desc(a_buf + kABytes + a_off, (uint32_t)kAGroup, 128u);
                    const uint64_t b_lo = make_desc(b_buf + kBPlane + b_off, NOUT * 16u, 128u);
                    mma_bf16(tmem, a_lo, b_hi, idesc, 1u);
                    mma_bf16(tmem, a_hi, b_lo, idesc, 1u);
                }
            }
            mma_commit(smem_u32(&s_empty[b]));
            if (s == stages - 1) mma_commit(smem_u32(&s_done));
        }
    }
    mbar_wait(smem_u32(&s_done), 0u);
    tc_fence_after();

    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (= pixels) and columns (w / 4) * NOUT/2 .. ----
    {
        const int q = warp & 3, ch0 = (warp >> 2) * (NOUT / 2);
        const int p_out = tile * kPix + q * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ch0;
#pragma unroll
        for (int c16 = 0; c16 < NOUT / 2; c16 += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)c16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int ch = ch0 + c16 + i;
                if (ch < n_store) {
                    const float r = v[i] + __ldg(bias + ch);
                    out[((size_t)agent * n_store + ch) * HW + p_out] = r;
                    if (DEFORM) {
                        float t = r;
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
                        if (lane == 0) tile_sums[((size_t)agent * gridDim.x * 4 + tile * 4 + q) * NOUT + ch] = t;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free<NOUT>(tmem);
}

// ------------------------------------------------------------------------------------------------
// offset1 fast path for W % 128 == 0 (every shipped configuration: 64 x 128 maps): a tile is 128 pixels of ONE image
// row, so the three taps kx = 0, 1, 2 of a kernel row are the same staged pixels shifted by one 16-byte operand row.
// Stage (ky, 32-channel chunk) = 130 pixels x0-1 .. x0+128 of input row y-1+ky, copied once from the channel-last
// planes (zero outside the image); the taps are addressed by moving the descriptor start by kx * 16 bytes (the conv_in
// scheme of denoiser_tc.cu).  3 x C/32 stages of 18 MMAs instead of 9 x C/32 stages of 6: a third of the loads, stores
// and barriers of the generic kernel (ncu: that one spends the offset layer waiting on its per-stage
// load -> store -> barrier -> MMA chain).  bf16x3 like the generic path.  grid = (H*W/128, n_agents), 256 threads.
// ------------------------------------------------------------------------------------------------
constexpr int kR3Pix = 130;
constexpr int kR3Group = kR3Pix * 16;                  // 2080 B = 32 (mod 128): the 4-lane pixel rows store conflict-free
constexpr int kR3Plane = 4 * kR3Group;
constexpr int kR3BBlk = 32 * 32 * 2 * 2;               // one tap's B: value + residual planes (= one packed stage, SC = 32)
constexpr int kR3Stage = 2 * kR3Plane + 3 * kR3BBlk;
constexpr int kR3Smem = 2 * kR3Stage;

__global__ void __launch_bounds__(kThreads)
k_me_offset_row3(const uint4 *__restrict__ xh, const uint4 *__restrict__ xl, const uint4 *__restrict__ wp,
                 const float *__restrict__ bias, int C, int H, int W, float *__restrict__ out) {
    constexpr int NOUT = 32, n_store = 18;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_empty[2], s_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int agent = blockIdx.y, tile = blockIdx.x;
    const int HW = H * W, C8 = C / 8, chunks = C / 32, stages = 3 * chunks;
    const int pix0 = tile * kPix, py = pix0 / W, x0 = pix0 - py * W;

    if (warp == 0) tmem_alloc<NOUT>(&s_tmem);
    if (tid == 32) {
        mbar_init(smem_u32(&s_empty[0]), 1); mbar_init(smem_u32(&s_empty[1]), 1); mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t s_base = smem_u32(smem);
    constexpr uint32_t idesc = make_idesc(128, NOUT);
    const uint4 *xh_a = xh + (size_t)agent * HW * C8;
    const uint4 *xl_a = xl + (size_t)agent * HW * C8;

    for (int s = 0; s < stages; ++s) {
        const int b = s & 1;
        const int ky = s / chunks, chunk = s - ky * chunks;
        if (s >= 2) mbar_wait(smem_u32(&s_empty[b]), (uint32_t)((s >> 1) - 1) & 1u);   // MMAs of stage s-2 retired
        uint8_t *a_buf = smem + b * kR3Stage, *b_buf = a_buf + 2 * kR3Plane;
        // ---- all nine 16-byte loads of this thread are issued before the first store (one L2 round trip per stage;
        // load -> store pairs in sequence made the layer wait five round trips per stage: ncu, 57 % of the stall
        // samples on the STS instructions) ----
        {
            // B: the three taps of kernel row ky (4 KB each, value + residual): 768 rows, three per thread
            uint4 bq[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
                bq[kx] = __ldg(wp + (size_t)((ky * 3 + kx) * chunks + chunk) * (kR3BBlk / 16) + tid);
            // A: 130 pixels x 4 channel groups, value and residual planes
            const int yy = py - 1 + ky;
            const bool row_ok = yy >= 0 && yy < H;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 hi[3], lo[3];
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int item = it * kThreads + tid;
                const int p = item >> 2, g = item & 3, xx = x0 - 1 + p;
                const bool ok = item < kR3Pix * 4 && row_ok && xx >= 0 && xx < W;
                const size_t idx = ok ? (size_t)(yy * W + xx) * C8 + chunk * 4 + g : 0;
                hi[it] = ok ? __ldg(xh_a + idx) : z;
                lo[it] = ok ? __ldg(xl_a + idx) : z;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) reinterpret_cast<uint4 *>(b_buf)[kx * (kR3BBlk / 16) + tid] = bq[kx];
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int item = it * kThreads + tid;
                if (item < kR3Pix * 4) {
                    uint4 *d = reinterpret_cast<uint4 *>(a_buf + (item & 3) * kR3Group) + (item >> 2);
                    d[0] = hi[it];
                    d[kR3Plane / 16] = lo[it];
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = s_base + (uint32_t)b * kR3Stage, b_addr = a_addr + 2u * kR3Plane;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {   // K = 16: channel groups 2j and 2j+1
                    const uint32_t a_off = (uint32_t)(2 * j) * kR3Group + (uint32_t)kx * 16u;
                    const uint32_t b_off = (uint32_t)kx * kR3BBlk + (uint32_t)(2 * j) * (NOUT * 16u);
                    const uint64_t a_hi = make_desc(a_addr + a_off, kR3Group, 128u);
                    const uint64_t a_lo = make_desc(a_addr + kR3Plane + a_off, kR3Group, 128u);
                    const uint64_t b_hi = make_desc(b_addr + b_off, NOUT * 16u, 128u);
                    const uint64_t b_lo = make_desc(b_addr + kR3BBlk / 2 + b_off, NOUT * 16u, 128u);
                    mma_bf16(tmem, a_hi, b_hi, idesc, (s > 0 || kx > 0 || j > 0) ? 1u : 0u);
                    mma_bf16(tmem, a_lo, b_hi, idesc, 1u);
                    mma_bf16(tmem, a_hi, b_lo, idesc, 1u);
                }
            }
            mma_commit(smem_u32(&s_empty[b]));
            if (s == stages - 1) mma_commit(smem_u32(&s_done));
        }
    }
    mbar_wait(smem_u32(&s_done), 0u);
    tc_fence_after();
    {
        const int q = warp & 3, ch0 = (warp >> 2) * (NOUT / 2);
        const int p_out = tile * kPix + q * 32 + lane;
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ch0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int ch = ch0 + i;
            if (ch < n_store) out[((size_t)agent * n_store + ch) * HW + p_out] = v[i] + __ldg(bias + ch);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free<NOUT>(tmem);
}

// ------------------------------------------------------------------------------------------------
// tail parameters (device blob, floats), see include/gencomm_b200.h
// ------------------------------------------------------------------------------------------------
constexpr int kOffBias = 0;                      // offset1.bias [18]
constexpr int kDcnBias = 32;                     // dcn1.bias [64]
constexpr int kAttn1W = kDcnBias + 64;           // attn.1.weight [32][64]
constexpr int kAttn1B = kAttn1W + 32 * 64;       // attn.1.bias [32]
constexpr int kAttn2W = kAttn1B + 32;            // attn.3.weight [64][32]
constexpr int kAttn2B = kAttn2W + 64 * 32;       // attn.3.bias [64]
constexpr int kFuse1W = kAttn2B + 64;            // fuse.0.weight [64][64]
constexpr int kFuse1B = kFuse1W + 64 * 64;       // fuse.0.bias [64]
constexpr int kFuse2W = kFuse1B + 64;            // fuse.2.weight [2][64]
constexpr int kFuse2B = kFuse2W + 2 * 64;        // fuse.2.bias [2]
constexpr int kTailFloats = kFuse2B + 2 + 6;     // padded to a multiple of 8

// global average pool (fixed summation order) + squeeze/excite.  grid = n_agents, 64 threads.
__global__ void __launch_bounds__(64)
k_me_attn(const float *__restrict__ tile_sums, int groups, int HW, const float *__restrict__ prm, float *__restrict__ attn) {
    __shared__ float s_mean[64], s_hid[32];
    const int a = blockIdx.x, c = threadIdx.x;
    const float *ts = tile_sums + (size_t)a * groups * 64;
    float s = 0.0f;
    for (int g = 0; g < groups; ++g) s += ts[(size_t)g * 64 + c];
    s_mean[c] = s / (float)HW;
    __syncthreads();
    if (c < 32) {
        float h = prm[kAttn1B + c];
        for (int k = 0; k < 64; ++k) h = fmaf(prm[kAttn1W + c * 64 + k], s_mean[k], h);
        s_hid[c] = fmaxf(h, 0.0f);
    }
    __syncthreads();
    float z = prm[kAttn2B + c];
    for (int k = 0; k < 32; ++k) z = fmaf(prm[kAttn2W + c * 32 + k], s_hid[k], z);
    attn[a * 64 + c] = 1.0f / (1.0f + expf(-z));
}

// enhanced = b1 * attn; out = fuse.2(relu(fuse.0(enhanced))).  grid = (H*W/128, n_agents), 128 threads (thread = pixel).
// The excitation is folded into fuse.0's weights per agent: W'[o][c] = W[o][c] * attn[c].
__global__ void __launch_bounds__(128)
k_me_tail(const float *__restrict__ b1, const float *__restrict__ attn, const float *__restrict__ prm, int HW,
          float *__restrict__ out) {
    __shared__ __align__(16) float s_w[64][64];   // [c][o]
    __shared__ float s_b[64], s_w2[2][64];
    const int a = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < 64 * 64; i += 128) {
        const int o = i >> 6, c = i & 63;
        s_w[c][o] = prm[kFuse1W + i] * attn[a * 64 + c];
    }
    if (tid < 64) { s_b[tid] = prm[kFuse1B + tid]; s_w2[0][tid] = prm[kFuse2W + tid]; s_w2[1][tid] = prm[kFuse2W + 64 + tid]; }
    __syncthreads();
    const int p = blockIdx.x * 128 + tid;
    if (p >= HW) return;
    float2 acc[32];   // packed fp32 (FFMA2): two output channels per instruction, same roundings as scalar fmaf
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = make_float2(0.0f, 0.0f);
    const float *src = b1 + (size_t)a * 64 * HW + p;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 8) {
        float v8[8];   // eight plane-strided loads in flight (two at a time left the kernel latency bound: 24 % issue-active)
#pragma unroll
        for (int k = 0; k < 8; ++k) v8[k] = __ldg(src + (size_t)(c0 + k) * HW);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 vv = make_float2(v8[k], v8[k]);
            const float4 *wr = reinterpret_cast<const float4 *>(&s_w[c0 + k][0]);
#pragma unroll
            for (int o4 = 0; o4 < 16; ++o4) {
                const float4 w = wr[o4];
                acc[2 * o4 + 0] = __ffma2_rn(make_float2(w.x, w.y), vv, acc[2 * o4 + 0]);
                acc[2 * o4 + 1] = __ffma2_rn(make_float2(w.z, w.w), vv, acc[2 * o4 + 1]);
            }
        }
    }
    float r0 = prm[kFuse2B], r1 = prm[kFuse2B + 1];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const float h0 = fmaxf(acc[o].x + s_b[2 * o], 0.0f), h1 = fmaxf(acc[o].y + s_b[2 * o + 1], 0.0f);
        r0 = fmaf(s_w2[0][2 * o], h0, r0);
        r1 = fmaf(s_w2[1][2 * o], h0, r1);
        r0 = fmaf(s_w2[0][2 * o + 1], h1, r0);
        r1 = fmaf(s_w2[1][2 * o + 1], h1, r1);
    }
    out[((size_t)a * 2 + 0) * HW + p] = r0;
    out[((size_t)a * 2 + 1) * HW + p] = r1;
}

// The same tail with fuse.0 (64 -> 64) on the tensor cores: one M = 128 pixels, N = 64, K = 64 GEMM per CTA in
// "bf16x3" (value + residual planes of both operands, three MMAs per K = 16 step, ~16 mantissa bits), the excitation
// folded into the A operand, ReLU + fuse.2 (64 -> 2) in the TMEM epilogue.  Replaces 2048 FFMA2 + 1024 LDS per pixel
// (k_me_tail: 117 us, latency bound) by 12 MMAs per 128 pixels.  grid = (H*W/128, n_agents), 128 threads (thread = pixel).
constexpr int kTtGroup = kPix * 16 + 16;        // [8 channel groups][128 pixels][16 B], padded group stride
constexpr int kTtPlane = 8 * kTtGroup;
constexpr int kTtBPlane = 64 * 64 * 2;          // [8 k groups][64 n][16 B]
constexpr int kTtSmem = 2 * kTtPlane + 2 * kTtBPlane;

__global__ void __launch_bounds__(128)
k_me_tail_tc(const float *__restrict__ b1, const float *__restrict__ attn, const float *__restrict__ prm, int HW,
             float *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *a_hi = smem, *a_lo = smem + kTtPlane, *b_hi = smem + 2 * kTtPlane, *b_lo = b_hi + kTtBPlane;
    __shared__ __align__(8) uint64_t s_done;
    __shared__ uint32_t s_tmem;
    __shared__ float s_attn[64], s_b[64], s_w2[2][64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a = blockIdx.y, p = blockIdx.x * kPix + tid;

    if (warp == 0) tmem_alloc<64>(&s_tmem);
    if (tid == 32) {
        mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) {
        s_attn[tid] = attn[a * 64 + tid];
        s_b[tid] = prm[kFuse1B + tid];
        s_w2[0][tid] = prm[kFuse2W + tid];
        s_w2[1][tid] = prm[kFuse2W + 64 + tid];
    }
    // B: fuse.0.weight [n][k] -> value / residual planes in core-matrix order [k8][n][8 k]
    for (int i = tid; i < 64 * 8; i += 128) {
        const int n = i & 63, k8 = i >> 6;
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(prm + kFuse1W + n * 64 + k8 * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(prm + kFuse1W + n * 64 + k8 * 8 + 4));
        const float v[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
            l[j] = pack_bf16(bf16_residual(v[2 * j]), bf16_residual(v[2 * j + 1]));
        }
        reinterpret_cast<uint4 *>(b_hi)[k8 * 64 + n] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4 *>(b_lo)[k8 * 64 + n] = make_uint4(l[0], l[1], l[2], l[3]);
    }
    // A: this thread's pixel, 64 plane-strided loads in flight
    float x[64];
    {
        const float *src = b1 + (size_t)a * 64 * HW + (p < HW ? p : 0);
#pragma unroll
        for (int c = 0; c < 64; ++c) x[c] = __ldg(src + (size_t)c * HW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v0 = x[8 * g + 2 * j] * s_attn[8 * g + 2 * j], v1 = x[8 * g + 2 * j + 1] * s_attn[8 * g + 2 * j + 1];
            h[j] = pack_bf16(v0, v1);
            l[j] = pack_bf16(bf16_residual(v0), bf16_residual(v1));
        }
        reinterpret_cast<uint4 *>(a_hi + g * kTtGroup)[tid] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4 *>(a_lo + g * kTtGroup)[tid] = make_uint4(l[0], l[1], l[2], l[3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc(128, 64);
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // K = 16: channel groups 2j and 2j+1
            const uint32_t a_off = (uint32_t)(2 * j) * kTtGroup, b_off = (uint32_t)(2 * j) * (64 * 16u);
            const uint64_t d_ah = make_desc(ah + a_off, kTtGroup, 128u), d_al = make_desc(al + a_off, kTtGroup, 128u);
            const uint64_t d_bh = make_desc(bh + b_off, 64 * 16u, 128u), d_bl = make_desc(bl + b_off, 64 * 16u, 128u);
            mma_bf16(tmem, d_ah, d_bh, idesc, j > 0 ? 1u : 0u);
            mma_bf16(tmem, d_al, d_bh, idesc, 1u);
            mma_bf16(tmem, d_ah, d_bl, idesc, 1u);
        }
        mma_commit(smem_u32(&s_done));
    }
    mbar_wait(smem_u32(&s_done), 0u);
    tc_fence_after();
    {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        float r0 = prm[kFuse2B], r1 = prm[kFuse2B + 1];
#pragma unroll
        for (int c16 = 0; c16 < 64; c16 += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)c16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float h = fmaxf(v[i] + s_b[c16 + i], 0.0f);
                r0 = fmaf(s_w2[0][c16 + i], h, r0);
                r1 = fmaf(s_w2[1][c16 + i], h, r1);
            }
        }
        if (p < HW) {
            out[((size_t)a * 2 + 0) * HW + p] = r0;
            out[((size_t)a * 2 + 1) * HW + p] = r1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free<64>(tmem);
}

struct Workspace {
    float *offset;      // [A][18][HW]
    float *b1;          // [A][64][HW]
    float *tile_sums;   // [A][HW/32][64]
    float *attn;        // [A][64]
    uint4 *xh, *xl;     // [A][HW][C] bf16: value and residual planes of the input, channel-last
    size_t bytes;
};
static Workspace carve(void *base, int A, int C, int HW) {
    Workspace w;
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) {
        char *p = b ? b + off : nullptr;
        off += align_up(bytes, 256);
        return (float *)p;
    };
    w.offset = take((size_t)A * 18 * HW * 4);
    w.b1 = take((size_t)A * 64 * HW * 4);
    w.tile_sums = take((size_t)A * (HW / 32) * 64 * 4);
    w.attn = take((size_t)A * 64 * 4);
    w.xh = (uint4 *)take((size_t)A * HW * C * 2);
    w.xl = (uint4 *)take((size_t)A * HW * C * 2);
    w.bytes = off;
    return w;
}
static inline size_t packed_off_bytes(int C) { return (size_t)9 * C * 32 * 2 * 2; }   // value + residual planes
static inline size_t packed_dcn_bytes(int C) { return (size_t)9 * C * 64 * 2; }

}  // namespace me
}  // namespace gc

using namespace gc;

extern "C" size_t gc_me_param_floats(void) { return me::kTailFloats; }
extern "C" size_t gc_me_packed_bytes(int C) {
    return C > 0 ? align_up(me::packed_off_bytes(C), 256) + me::packed_dcn_bytes(C) : 0;
}
extern "C" size_t gc_me_workspace_bytes(int total_agents, int C, int H, int W) {
    if (total_agents <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return me::carve(nullptr, total_agents, C, H * W).bytes;
}

extern "C" int gc_me_pack_weights(const float *w_offset, const float *w_dcn, int C, void *packed, void *stream) {
    GC_REQUIRE(w_offset && w_dcn && packed, GC_EINVAL, "gc_me_pack_weights: null pointer");
    GC_REQUIRE(C > 0 && C % 64 == 0, GC_EUNSUPPORTED, "gc_me_pack_weights: C must be a multiple of 64 (got %d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    uint4 *p_off = (uint4 *)packed;
    uint4 *p_dcn = (uint4 *)((char *)packed + align_up(me::packed_off_bytes(C), 256));
    const int n_off = 9 * (C / 8) * 32, n_dcn = 9 * (C / 8) * 64;
    me::k_me_pack<<<(n_off + 255) / 256, 256, 0, st>>>(w_offset, 18, 32, C, me::kScOffset, 1, p_off);
    GC_LAUNCH_CHECK("k_me_pack(offset1)");
    me::k_me_pack<<<(n_dcn + 255) / 256, 256, 0, st>>>(w_dcn, 64, 64, C, me::kScDeform, 0, p_dcn);
    GC_LAUNCH_CHECK("k_me_pack(dcn1)");
    return GC_OK;
}

extern "C" int gc_message_extractor(const float *x, int total_agents, int C, int H, int W, const void *packed,
                                    const float *params, void *workspace, float *message, void *stream) {
    GC_REQUIRE(total_agents >= 0 && total_agents <= 65535, GC_EINVAL, "gc_message_extractor: bad agent count");
    if (total_agents == 0) return GC_OK;
    GC_REQUIRE(x && packed && params && workspace && message, GC_EINVAL, "gc_message_extractor: null pointer");
    GC_REQUIRE(C > 0 && C % 64 == 0, GC_EUNSUPPORTED, "gc_message_extractor: C must be a multiple of 64 (got %d)", C);
    GC_REQUIRE(H > 0 && W > 0 && (H * W) % me::kPix == 0, GC_EUNSUPPORTED,
               "gc_message_extractor: H*W must be a multiple of 128 (got %dx%d)", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W, tiles = HW / me::kPix;
    const me::Workspace ws = me::carve(workspace, total_agents, C, HW);
    const uint4 *p_off = (const uint4 *)packed;
    const uint4 *p_dcn = (const uint4 *)((const char *)packed + align_up(me::packed_off_bytes(C), 256));
    static bool attr_done = false;
    constexpr int kSmemOff = me::conv_smem_bytes(32, true, me::kScOffset), kSmemDcn = me::conv_smem_bytes(64, false, me::kScDeform);
    if (!attr_done) {
        cudaFuncSetAttribute(me::k_me_conv<32, false, me::kScOffset>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOff);
        cudaFuncSetAttribute(me::k_me_conv<64, true, me::kScDeform>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDcn);
        cudaFuncSetAttribute(me::k_me_offset_row3, cudaFuncAttributeMaxDynamicSharedMemorySize, me::kR3Smem);
        cudaFuncSetAttribute(me::k_me_tail_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, me::kTtSmem);
        attr_done = true;
    }
    const dim3 grid(tiles, total_agents);
    me::k_me_to_nhwc<<<dim3((HW + 63) / 64, C / 64, total_agents), 256, 0, st>>>(x, C, HW, ws.xh, ws.xl);
    GC_LAUNCH_CHECK("k_me_to_nhwc");
    if (W % me::kPix == 0 && !getenv("GC_ME_GENERIC_OFFSET")) {   // a tile is 128 pixels of one image row
        me::k_me_offset_row3<<<grid, me::kThreads, me::kR3Smem, st>>>(ws.xh, ws.xl, p_off, params + me::kOffBias, C, H, W,
                                                                      ws.offset);
        GC_LAUNCH_CHECK("k_me_offset_row3");
    } else {
        me::k_me_conv<32, false, me::kScOffset><<<grid, me::kThreads, kSmemOff, st>>>(ws.xh, ws.xl, nullptr, p_off,
                                                                                   params + me::kOffBias, C, H, W, 18,
                                                                                   ws.offset, nullptr);
        GC_LAUNCH_CHECK("k_me_conv<offset1>");
    }
    me::k_me_conv<64, true, me::kScDeform><<<grid, me::kThreads, kSmemDcn, st>>>(ws.xh, ws.xl, ws.offset, p_dcn, params + me::kDcnBias, C, H,
                                                                  W, 64, ws.b1, ws.tile_sums);
    GC_LAUNCH_CHECK("k_me_conv<dcn1>");
    me::k_me_attn<<<total_agents, 64, 0, st>>>(ws.tile_sums, tiles * 4, HW, params, ws.attn);
    GC_LAUNCH_CHECK("k_me_attn");
    if (!getenv("GC_ME_FP32_TAIL")) {
        me::k_me_tail_tc<<<grid, 128, me::kTtSmem, st>>>(ws.b1, ws.attn, params, HW, message);
        GC_LAUNCH_CHECK("k_me_tail_tc");
    } else {   // CUDA-core fp32 tail (A/B)
        me::k_me_tail<<<grid, 128, 0, st>>>(ws.b1, ws.attn, params, HW, message);
        GC_LAUNCH_CHECK("k_me_tail");
    }
    return GC_OK;
}
