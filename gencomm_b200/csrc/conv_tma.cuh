// Persistent TMA-fed version of the plain path of k_me_conv (implicit_gemm.cuh): stride-1 3x3 (padding 1) or 1x1 convolution over
// channel-last bf16 value + residual planes as a tcgen05 implicit GEMM in bf16x3.
//
// Why (profiles/r02ae_conv_stagers.txt): in k_me_conv 256 threads gather every [128 px x 32 ch] operand stage through registers
// (LDG -> STS -> fence.proxy.async -> 256 mbarrier arrivals).  The ncu source page of the 256-channel layer shows those threads
// busy, not waiting: STS + FENCE.VIEW.ASYNC + SYNCS.ARRIVE take as many stall samples as the wait for a free stage, L2 -> SM
// traffic sits at 28 % of peak, the tensor pipe idles half the time.  A cp.async variant of the same gather with decoupled
// epilogue warps was measured slower still (7.2 vs 6.0 ms per backbone call).  The gather is what TMA is for:
//   * the activation planes are a 4-D tensor (C, W, H, agent); the operand of tap (ky, kx) for a tile of BH rows x BW pixels
//     (BH * BW = 128) is the box {32 ch, BW, BH, 1} at (chunk * 32, x0 + kx - 1, y0 + ky - 1, agent) -- padding is TMA's
//     out-of-bounds zero fill -- written with the 64-byte swizzle, which is exactly UMMA's K-major SWIZZLE_64B operand layout
//     (row = pixel, 64 bytes = 32 channels): no thread touches the operand;
//   * one producer lane issues, per stage, two tensor loads (value + residual plane) and one bulk copy of packed weights on one
//     mbarrier; one feeder lane issues the MMAs; four epilogue warps drain one of two TMEM accumulator sets while the next
//     tile accumulates into the other.  A CTA owns an SM for the whole launch and walks tiles b, b + gridDim.x, ...
// Same packed weights (k_me_pack, split), same arithmetic and epilogues as k_me_conv: results are bit-identical.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "implicit_gemm.cuh"

namespace gc {
namespace ct {

using namespace umma;
using me::bf16_residual;
using me::gelu_erf;

constexpr int kEpi = 256;                            // warps 0-7: epilogue (warp w: TMEM lanes 32 (w % 4) .., column half w / 4)
constexpr int kThreads = kEpi + 64;                  // warp 8: feeder (MMA issue), warp 9: producer (TMA)
constexpr int kSc = 32;                              // channels per stage = one 64-byte swizzle row
constexpr int kAPlane = 128 * kSc * 2, kAStage = 2 * kAPlane;   // value + residual

__host__ __device__ constexpr int b_stage_bytes(int NOUT) { return kSc * NOUT * 2 * 2; }
__host__ __device__ constexpr int ring_depth(int NOUT) {
    return (200 * 1024) / (kAStage + b_stage_bytes(NOUT)) > 8 ? 8 : (200 * 1024) / (kAStage + b_stage_bytes(NOUT));
}
__host__ __device__ constexpr int smem_bytes(int NOUT) { return ring_depth(NOUT) * (kAStage + b_stage_bytes(NOUT)) + 1024; }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c, int x, int y, int a, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(a), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// K-major SWIZZLE_64B operand: rows of 64 bytes, 8-row groups 512 bytes apart (LBO unused), layout type 4 in bits 61-63
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) { return make_desc(saddr, 16u, 512u) | (4ull << 61); }

// H x W: the output grid; tiles of BH rows x BW pixels, BW = min(W, 128), BH = 128 / BW.  stride 2: the tensor maps traverse the
// input with element strides {1, 2, 2, 1} (box {32, 2 BW, 2 BH, 1} -> BW x BH pixels in shared memory), start (2 x0 + kx - 1, ..).
// EPI as in k_me_conv: 0 bias, 1 bias + GELU, 3 bias + ReLU -> fp32 [A][out_ch_total][H*up][W*up] at (y*up + up_dy, x*up + up_dx);
// 2 bias + GELU -> fp32 channel-last [A][HW][out_ch_total]; 5 bias + ReLU -> bf16 value + residual planes [A][HW*up*up][out_ch_total]
template <int NOUT, int TAPS, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tma(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l, const uint4 *__restrict__ wp,
           const float *__restrict__ bias, int n_tiles, int c_in, int H, int W, int BW, int stride, int n_store, int out_ch_total, int out_ch_off,
           float *__restrict__ out, uint4 *__restrict__ oh, uint4 *__restrict__ ol, int up, int up_dy, int up_dx) {
    constexpr int NS = ring_depth(NOUT);
    constexpr int kBStage = b_stage_bytes(NOUT), kBPlane = kSc * NOUT * 2;
    static_assert(NS >= 3, "ring too shallow");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[NS], empty[NS], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[NOUT];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // swizzle atoms: 512-byte aligned tiles
    const uint32_t b_base = a_base + NS * kAStage;
    const int BH = me::kPix / BW, tiles_x = W / BW, tiles_agent = tiles_x * (H / BH);
    const int chunks = c_in / kSc, stages = TAPS * chunks, HW = H * W;

    if (warp == 0) tmem_alloc<2 * NOUT>(&s_tmem);
    if (tid == 32) {
        for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), kEpi); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NOUT; i += kThreads) s_bias[i] = (bias && i < n_store) ? bias[i] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == kEpi / 32) {
        // ---- feeder ----
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(128, NOUT);
            int g = 0;
            for (int k = 0; k < my_tiles; ++k) {
                const int set = k & 1;
                if (k >= 2) { mbar_wait(smem_u32(&acc_empty[set]), (uint32_t)((k >> 1) - 1) & 1u); tc_fence_after(); }
                const uint32_t acc = tmem + (uint32_t)(set * NOUT);
                for (int s = 0; s < stages; ++s, ++g) {
                    const int sb = g % NS;
                    mbar_wait(smem_u32(&full[sb]), (uint32_t)(g / NS) & 1u);
                    tc_fence_after();
                    const uint64_t a_d = make_desc_sw64(a_base + (uint32_t)sb * kAStage);
                    const uint64_t b_d = make_desc(b_base + (uint32_t)sb * kBStage, NOUT * 16u, 128u);
#pragma unroll
                    for (int j = 0; j < kSc / 16; ++j) {   // +1 in the address field = 16 bytes; K = 16 bf16 = 32 bytes of the row
                        const uint64_t a_hi = a_d + (uint64_t)(2 * j), a_lo = a_hi + (uint64_t)(kAPlane / 16);
                        const uint64_t b_hi = b_d + (uint64_t)(2 * j * NOUT), b_lo = b_hi + (uint64_t)(kBPlane / 16);
                        mma_bf16(acc, a_hi, b_hi, idesc, (s > 0 || j > 0) ? 1u : 0u);
                        mma_bf16(acc, a_lo, b_hi, idesc, 1u);
                        mma_bf16(acc, a_hi, b_lo, idesc, 1u);
                    }
                    mma_commit(smem_u32(&empty[sb]));
                }
                mma_commit(smem_u32(&acc_full[set]));
            }
        }
        __syncwarp();
    } else if (warp == kEpi / 32 + 1) {
        // ---- producer: per stage two tensor loads (value, residual plane) + the packed weights, one mbarrier ----
        if (elect_one()) {
            int g = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int agent = t / tiles_agent, rem = t - agent * tiles_agent;
                const int x0 = (rem % tiles_x) * BW, y0 = (rem / tiles_x) * BH;
                for (int s = 0; s < stages; ++s, ++g) {
                    const int sb = g % NS;
                    if (g >= NS) mbar_wait(smem_u32(&empty[sb]), (uint32_t)((g / NS) - 1) & 1u);   // MMAs of stage g - NS retired
                    const int tap = s / chunks, chunk = s - tap * chunks;
                    const int ky = TAPS == 1 ? 1 : tap / 3, kx = TAPS == 1 ? 1 : tap - 3 * ky;
                    const uint32_t bar = smem_u32(&full[sb]), a_dst = a_base + (uint32_t)sb * kAStage;
                    mbar_expect_tx(bar, (uint32_t)(kAStage + kBStage));
                    tma_load_4d(a_dst, &map_h, chunk * kSc, x0 * stride + kx - 1, y0 * stride + ky - 1, agent, bar);
                    tma_load_4d(a_dst + kAPlane, &map_l, chunk * kSc, x0 * stride + kx - 1, y0 * stride + ky - 1, agent, bar);
                    bulk_copy(b_base + (uint32_t)sb * kBStage, wp + (size_t)s * (kBStage / 16), kBStage, bar);
                }
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: lane = pixel (row bh, column bw of the tile); warps 0-3 take the first half of the columns, 4-7 the second ----
        const size_t hw_store = (size_t)HW * up * up;
        const int q4 = warp & 3, m = q4 * 32 + lane, col0 = (warp >> 2) * (NOUT / 2);
        int k = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
            const int agent = t / tiles_agent, rem = t - agent * tiles_agent;
            const int pyo = (rem / tiles_x) * BH + m / BW, pxo = (rem % tiles_x) * BW + m % BW;
            const int p_out = pyo * W + pxo;
            const size_t p_store = (size_t)(pyo * up + up_dy) * (W * up) + pxo * up + up_dx;
            const int set = k & 1;
            mbar_wait(smem_u32(&acc_full[set]), (uint32_t)(k >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(set * NOUT);
#pragma unroll 2
            for (int c16 = col0; c16 < col0 + NOUT / 2; c16 += 16) {
                if (c16 >= n_store) break;
                float v[16];
                tmem_ld16(taddr + (uint32_t)c16, v);
                if (EPI == 5) {
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float a = fmaxf(v[2 * i] + s_bias[c16 + 2 * i], 0.0f);
                        const float b = fmaxf(v[2 * i + 1] + s_bias[c16 + 2 * i + 1], 0.0f);
                        h[i] = pack_bf16(a, b);
                        l[i] = pack_bf16(bf16_residual(a), bf16_residual(b));
                    }
                    const size_t o = ((size_t)agent * hw_store + p_store) * (out_ch_total >> 3) + ((out_ch_off + c16) >> 3);
                    oh[o] = make_uint4(h[0], h[1], h[2], h[3]); oh[o + 1] = make_uint4(h[4], h[5], h[6], h[7]);
                    ol[o] = make_uint4(l[0], l[1], l[2], l[3]); ol[o + 1] = make_uint4(l[4], l[5], l[6], l[7]);
                } else if (EPI == 2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i] + s_bias[c16 + i]);
                    float4 *dst = reinterpret_cast<float4 *>(out + ((size_t)agent * HW + p_out) * out_ch_total + out_ch_off + c16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int ch = c16 + i;
                        if (ch < n_store) {
                            float r = v[i] + s_bias[ch];
                            if (EPI == 1) r = gelu_erf(r);
                            if (EPI == 3) r = fmaxf(r, 0.0f);
                            out[((size_t)agent * out_ch_total + out_ch_off + ch) * hw_store + p_store] = r;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&acc_empty[set]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free<2 * NOUT>(tmem);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
        (void)cudaGetLastError();
    }
    return fn;
}
// planes [A][H][W][C] bf16 as the 4-D tensor (C, W, H, A); box = {32 channels, BW, BH, 1} output pixels, 64-byte swizzle, zero fill
// outside; H x W is the INPUT grid, every stride-th pixel is loaded
inline bool encode_plane_map(CUtensorMap *map, const void *plane, int A, int H, int W, int C, int BW, int BH, int stride) {
    PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
    if (!encode) return false;
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)A};
    const cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kSc, (cuuint32_t)(BW * stride), (cuuint32_t)(BH * stride), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(plane), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// layers whose output map tiles into BH x BW = 128-pixel boxes; GC_CONV_TMA=0 disables, =1 restricts to stride 1 (A/B)
inline bool conv_tma_eligible(int stride, int C, int c_in, int H, int W, int up) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("GC_CONV_TMA"); on = e ? atoi(e) : 2; }
    if (!on || (stride != 1 && !(stride == 2 && on >= 2)) || C % 8 != 0 || c_in % kSc != 0 || c_in > C) return false;
    const int BW = W < 128 ? W : 128;
    if (BW < 8 || (BW & (BW - 1)) != 0 || W % BW != 0) return false;
    return H % (128 / BW) == 0 && up >= 1;
}

template <int NOUT, int TAPS, int EPI>
static int launch_conv_tma(cudaStream_t st, int A, const uint4 *xh, const uint4 *xl, const uint4 *wp, const float *bias, int C, int c_in,
                           int H, int W, int H_in, int W_in, int stride, int n_store, int out_ch_total, int out_ch_off, float *out,
                           uint4 *oh, uint4 *ol, int up, int up_dy, int up_dx) {
    constexpr int kSmem = smem_bytes(NOUT);
    static_assert(kSmem <= 227 * 1024, "k_conv_tma: shared memory");
    static int sms = 0;
    if (!sms) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_tma<NOUT, TAPS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        int dev = 0, n = 0;
        if (e == cudaSuccess) e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || n <= 0) {
            (void)cudaGetLastError();
            set_error("k_conv_tma: launch set-up failed (%d)", (int)e);
            return e != cudaSuccess ? (int)e : (int)cudaErrorUnknown;
        }
        sms = n;
    }
    const int BW = W < 128 ? W : 128, BH = 128 / BW;
    CUtensorMap mh, ml;
    const int Hi = H_in > 0 ? H_in : H, Wi = W_in > 0 ? W_in : W;
    if (!encode_plane_map(&mh, xh, A, Hi, Wi, C, BW, BH, stride) || !encode_plane_map(&ml, xl, A, Hi, Wi, C, BW, BH, stride)) {
        set_error("k_conv_tma: cuTensorMapEncodeTiled failed (A=%d H=%d W=%d C=%d stride=%d)", A, Hi, Wi, C, stride);
        return (int)cudaErrorInvalidValue;
    }
    const int n_tiles = A * (H * W / me::kPix);
    k_conv_tma<NOUT, TAPS, EPI><<<n_tiles < sms ? n_tiles : sms, kThreads, kSmem, st>>>(mh, ml, wp, bias, n_tiles, c_in, H, W, BW, stride, n_store,
                                                                                       out_ch_total, out_ch_off, out, oh, ol, up, up_dy, up_dx);
    GC_LAUNCH_CHECK("k_conv_tma");
    return GC_OK;
}

}  // namespace ct
}  // namespace gc
