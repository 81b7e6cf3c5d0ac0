// Cluster-resident DiffusionUNet for sm_100a: the 26 width-8 layers of one denoiser evaluation (and, when fused,
// conv_in in front and norm_out + conv_out + the sampler update behind them) run inside ONE thread-block cluster
// per agent.  The agent's activations never leave the SMs: every 8-channel tensor lives in shared memory, split
// into 8 row bands (one per CTA of the cluster); 3x3 halos are read from the neighbouring CTA through
// distributed shared memory, GroupNorm statistics are exchanged by remote shared-memory stores, and the only
// synchronisation between layers is one cluster barrier.
//
// Replaces (paths relative to /root/reference/opencood):
//   models/gencomm_modules/unet.py:307-344 DiffusionUNet.forward, :81-138 ResnetBlock, :59-78 Downsample,
//        :40-56 Upsample, :36-37 GroupNorm(4, eps 1e-6), :31-33 swish
//   models/gencomm_modules/cond_diff.py:321-329 p_sample_loop (the T evaluations of one agent run back to back in the
//        same cluster, so x_t stays in L2 between steps)
// for the shipped denoiser shape (ch=8, ch_mult=[1,1], num_res_blocks=2, no attention) at H x W = 64 x 128.
//
// Tensor-core formulation (tcgen05, tf32 operands, fp32 accumulation in TMEM) - "input-row stationary":
//   D[128 pixels of input row i, 32] += A_i[128, 8 cin] * B_kx[8 cin, 4 blocks x 8 cout]      (one MMA per tap kx)
// where block j of B holds the weights of tap ky = 2 - j (block 3 = 0).  The 32 accumulator columns of input row i
// start at TMEM column 8 i, i.e. block j lands in the accumulator of output row i - 2 + j (kept at column
// 8 (r + 2)): one staged row feeds all three output rows it contributes to, so a 3x3 layer costs 3 MMAs per input
// row and channel group instead of 9, and the staged row is read in place for the three kx taps (descriptor start
// address shifted by one 16-byte pixel).  TMEM columns are zeroed with tcgen05.st before the first MMA.
#include <cooperative_groups.h>
#include <stddef.h>
#include <stdlib.h>

#include "common.cuh"
#include "denoiser_cluster.cuh"
#include "denoiser_tc.cuh"
#include "umma.cuh"

namespace gc {
namespace cl {

namespace cg = cooperative_groups;
using namespace umma;

constexpr int kThreads = 512;                      // 16 warps: halves every per-thread latency chain (r02c: 19 % issue-active at 8 warps)
constexpr int kWarps = kThreads / 32;
constexpr int kRowPx = 130;                        // staged pixels per operand row: x = -1 .. 128
constexpr int kFBytes = 8 * 2 * 128 * 16;          // one full-resolution raw tensor band: [8 rows][2 halves][128 px] float4
constexpr int kQBytes = 4 * 2 * 64 * 16;           // half resolution: [4 rows][2 halves][64 px] float4
constexpr int kOperBytes = 10 * 2 * kRowPx * 16;   // tf32 operand ring: cin 8: [2 planes][10 rows][130]; cin 16: [4][5][130]
constexpr int kNumBufs = 13;                       // F0..F4, Q0..Q7 (Q0-3 alias F3, Q4-7 alias F4)

// ---- shared memory carve-up (bytes from the 1024-aligned dynamic base) ----
constexpr int kOffF = 0;
constexpr int kOffOper = kOffF + 5 * kFBytes;                    // 163840
constexpr int kOffRec = kOffOper + kOperBytes;                   // 205440
constexpr int kOffStats = kOffRec + 2 * kClRecBytes;             // 219136
constexpr int kOffMisc = kOffStats + kNumBufs * kCl * 8 * 4;     // 222464
constexpr int kSmemBytes = kOffMisc + 1024;

struct Misc {
    float ga[16], gb[16];
    float part[kWarps][8];
    uint64_t mma_bar[2], done_bar[2], rec_bar[2];
    uint32_t tmem;
};
static_assert(sizeof(Misc) <= 1024, "Misc does not fit its slot");
static_assert(kSmemBytes <= 227 * 1024, "cluster UNet kernel exceeds the shared memory of one SM");
static_assert(kClRecBytes % 16 == 0, "layer records are bulk-copied");

enum Kind { kConv = 0, kDownK = 1, kUpK = 2 };
struct LayerCfg { int8_t kind, cin, gn, res, in_a, in_b, res_a, res_b, out, half; };
enum Buf { F0 = 0, F1, F2, F3, F4, Q0, Q1, Q2, Q3, Q4, Q5, Q6, Q7 };

// Execution order of unet.py:315-340 with the buffer each tensor lives in (hs[] = skip stack):
//   F0 = hs0 (conv_in), F1 = hs1, F2 = hs2, Q0 = hs3, Q1 = hs4, Q2 = hs5; F3/F4/Q4.. = temporaries.
__constant__ LayerCfg c_layers[kClLayers] = {
    {kConv, 8, 1, kNone, F0, F0, F0, F0, F3, 0},    //  0 down.0.block.0 conv1
    {kConv, 8, 1, kIdent, F3, F3, F0, F0, F1, 0},   //  1                conv2 + x        -> hs1
    {kConv, 8, 1, kNone, F1, F1, F1, F1, F3, 0},    //  2 down.0.block.1 conv1
    {kConv, 8, 1, kIdent, F3, F3, F1, F1, F2, 0},   //  3                conv2 + x        -> hs2
    {kDownK, 8, 0, kNone, F2, F2, F2, F2, Q0, 1},   //  4 down.0.downsample               -> hs3
    {kConv, 8, 1, kNone, Q0, Q0, Q0, Q0, Q4, 1},    //  5 down.1.block.0 conv1
    {kConv, 8, 1, kIdent, Q4, Q4, Q0, Q0, Q1, 1},   //  6                conv2 + x        -> hs4
    {kConv, 8, 1, kNone, Q1, Q1, Q1, Q1, Q4, 1},    //  7 down.1.block.1 conv1
    {kConv, 8, 1, kIdent, Q4, Q4, Q1, Q1, Q2, 1},   //  8                conv2 + x        -> hs5
    {kConv, 8, 1, kNone, Q2, Q2, Q2, Q2, Q4, 1},    //  9 mid.block_1 conv1
    {kConv, 8, 1, kIdent, Q4, Q4, Q2, Q2, Q3, 1},   // 10             conv2 + x
    {kConv, 8, 1, kNone, Q3, Q3, Q3, Q3, Q4, 1},    // 11 mid.block_2 conv1
    {kConv, 8, 1, kIdent, Q4, Q4, Q3, Q3, Q5, 1},   // 12             conv2 + x
    {kConv, 16, 1, kNone, Q5, Q2, Q5, Q2, Q4, 1},   // 13 up.1.block.0 conv1 on cat(h, hs5)
    {kConv, 8, 1, kNin, Q4, Q4, Q5, Q2, Q6, 1},     // 14              conv2 + nin(cat)
    {kConv, 16, 1, kNone, Q6, Q1, Q6, Q1, Q4, 1},   // 15 up.1.block.1 conv1 on cat(h, hs4)
    {kConv, 8, 1, kNin, Q4, Q4, Q6, Q1, Q5, 1},     // 16              conv2 + nin(cat)
    {kConv, 16, 1, kNone, Q5, Q0, Q5, Q0, Q4, 1},   // 17 up.1.block.2 conv1 on cat(h, hs3)
    {kConv, 8, 1, kNin, Q4, Q4, Q5, Q0, Q6, 1},     // 18              conv2 + nin(cat)
    {kUpK, 8, 0, kNone, Q6, Q6, Q6, Q6, F3, 0},     // 19 up.1.upsample (nearest x2 + 3x3)
    {kConv, 16, 1, kNone, F3, F2, F3, F2, F4, 0},   // 20 up.0.block.0 conv1 on cat(h, hs2)
    {kConv, 8, 1, kNin, F4, F4, F3, F2, F2, 0},     // 21              conv2 + nin(cat)   (in place over hs2)
    {kConv, 16, 1, kNone, F2, F1, F2, F1, F4, 0},   // 22 up.0.block.1 conv1 on cat(h, hs1)
    {kConv, 8, 1, kNin, F4, F4, F2, F1, F3, 0},     // 23              conv2 + nin(cat)
    {kConv, 16, 1, kNone, F3, F0, F3, F0, F4, 0},   // 24 up.0.block.2 conv1 on cat(h, hs0)
    {kConv, 8, 1, kNin, F4, F4, F3, F0, F2, 0},     // 25              conv2 + nin(cat)   -> input of norm_out
};

__device__ __forceinline__ uint32_t buf_off(int id) {
    return id < 5 ? (uint32_t)(id * kFBytes)
                  : (id < 9 ? (uint32_t)(3 * kFBytes + (id - 5) * kQBytes) : (uint32_t)(4 * kFBytes + (id - 9) * kQBytes));
}

// ---- distributed shared memory ----
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_cluster4(uint32_t caddr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(caddr));
    return v;
}
__device__ __forceinline__ void st_cluster(uint32_t caddr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tensor memory ----
__device__ __forceinline__ void tmem_alloc512(uint32_t *slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_zero8(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float lds1(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// tcgen05.mma / commit issued by ONE elected lane of a converged warp.  Every operand is warp-uniform, so ptxas keeps
// the descriptors in uniform registers; a `tid == 0` branch instead makes it wrap each MMA in a lane-serialising
// R2UR loop (measured on the B200: 88 cycles per MMA, profiles/r02b_*).
__device__ __forceinline__ void mma_tf32_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pa;\n\t"
        "setp.eq.b32 pa, 0, 0;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pa;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(bar) : "memory");
}

// x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU per element (tanh.approx, rel. error 2^-11 = the tf32 operand grid)
__device__ __forceinline__ float swish_tanh(float u) {
    const float h = 0.5f * u;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

struct Ctx {
    uint32_t smem;        // shared::cta address of the carve-up base
    uint32_t rank;        // CTA rank in the cluster = row band
    uint32_t tmem;
    uint32_t ph_mma[2], ph_done[2], ph_rec[2];   // mbarrier phase parities (uniform across the CTA)
    int tid, warp, lane;
    long long *trace;     // GC_CL_DEBUG & 32: per-layer clock64 timeline of CTA 0 / thread 0: [layer][16]
    int trace_row;
    int dbg;              // timing experiments (GC_CL_DEBUG): 1 skip MMAs, 2 skip staging, 4 skip epilogue math, 8 skip cluster barriers
};
#define GC_TRACE(c, k) do { if ((c).trace != nullptr && (c).tid == 0) (c).trace[(c).trace_row * 16 + (k)] = clock64(); } while (0)
#define MISC_ADDR(c, member) ((c).smem + kOffMisc + (uint32_t)offsetof(Misc, member))

// ------------------------------------------------------------------------------------------------
// Stage the rows [i0, i1) of a tensor-core layer's operand: GroupNorm + swish of the raw input (own band or the
// neighbour's halo row through DSMEM), zero x-halo columns, fp32 (tf32) pixel halves into the operand planes.
// KMAX = ceil(rows * (PXW + 2) / 256) items per thread.
// ------------------------------------------------------------------------------------------------
template <int CG, bool HALF, bool UP, bool GN, int KMAX>
__device__ __forceinline__ void stage_rows(const Ctx &c, const LayerCfg &L, int i0, int i1) {
    constexpr int PXW = HALF ? 64 : 128, RW = PXW + 2, R = HALF ? 4 : 8, HRES = HALF ? 32 : 64;
    constexpr int SPXW = UP ? 64 : PXW, SR = UP ? 4 : R;             // source tensor geometry
    constexpr int PROWS = CG == 1 ? 10 : 5;                          // slot rows per operand plane
    const int nitems = (i1 - i0) * RW;
    const int y0 = (int)c.rank * R;
    float4 v[KMAX][CG][2];
    int slot_px[KMAX];
    bool inside[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int it = c.tid + k * kThreads;
        slot_px[k] = -1;
        inside[k] = false;
#pragma unroll
        for (int g = 0; g < CG; ++g) v[k][g][0] = v[k][g][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it >= nitems) continue;
        const int ri = it / RW, px = it - ri * RW, i = i0 + ri, gy = y0 - 1 + i;
        if (gy < 0 || gy >= HRES) continue;                       // row outside the image: never read by an MMA
        const int slot = CG == 1 ? i : (((i >> 1) & 1) * 2 + (i & 1));
        slot_px[k] = slot * kRowPx + px;
        if (px < 1 || px > PXW) continue;                         // x halo: zeros (padding applies after the activation)
        inside[k] = true;
        const int sx = UP ? ((px - 1) >> 1) : (px - 1);
        int sy = UP ? ((gy >> 1) - (int)c.rank * SR) : (i - 1);   // source row relative to this CTA's band
        uint32_t srank = c.rank;
        if (sy < 0) { srank = c.rank - 1; sy += SR; } else if (sy >= SR) { srank = c.rank + 1; sy -= SR; }
#pragma unroll
        for (int g = 0; g < CG; ++g) {
            const uint32_t a = mapa(c.smem + kOffF + buf_off(g == 0 ? L.in_a : L.in_b) + (uint32_t)((sy * 2) * SPXW + sx) * 16u, srank);
            v[k][g][0] = ld_cluster4(a);
            v[k][g][1] = ld_cluster4(a + SPXW * 16u);
        }
    }
    float4 ga[CG][2], gb[CG][2];
    if (GN) {
#pragma unroll
        for (int g = 0; g < CG; ++g) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                ga[g][h] = lds4(MISC_ADDR(c, ga) + (uint32_t)(g * 8 + h * 4) * 4u);
                gb[g][h] = lds4(MISC_ADDR(c, gb) + (uint32_t)(g * 8 + h * 4) * 4u);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (slot_px[k] < 0) continue;
#pragma unroll
        for (int g = 0; g < CG; ++g) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 u = v[k][g][h];
                if (GN && inside[k]) {
                    u.x = swish_tanh(fmaf(u.x, ga[g][h].x, gb[g][h].x));
                    u.y = swish_tanh(fmaf(u.y, ga[g][h].y, gb[g][h].y));
                    u.z = swish_tanh(fmaf(u.z, ga[g][h].z, gb[g][h].z));
                    u.w = swish_tanh(fmaf(u.w, ga[g][h].w, gb[g][h].w));
                }
                sts4(c.smem + kOffOper + (uint32_t)(((g * 2 + h) * PROWS) * kRowPx + slot_px[k]) * 16u, u);
            }
        }
    }
}

// Per-CTA partial GroupNorm sums of the tile just produced -> slot `out` of every CTA of the cluster.
__device__ __forceinline__ void push_stats(const Ctx &c, float (&q8)[8], int out_buf, float *gstats) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) q8[i] += __shfl_xor_sync(0xffffffffu, q8[i], m);
    }
    if (c.lane == 0) {
        sts4(MISC_ADDR(c, part) + (uint32_t)c.warp * 32u, make_float4(q8[0], q8[1], q8[2], q8[3]));
        sts4(MISC_ADDR(c, part) + (uint32_t)c.warp * 32u + 16u, make_float4(q8[4], q8[5], q8[6], q8[7]));
    }
    __syncthreads();
    if (c.tid < 64) {
        const int vi = c.tid & 7, peer = c.tid >> 3;
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += lds1(MISC_ADDR(c, part) + (uint32_t)(w * 8 + vi) * 4u);
        st_cluster(mapa(c.smem + kOffStats + (uint32_t)(((out_buf * kCl) + (int)c.rank) * 8 + vi) * 4u, (uint32_t)peer), t);
        if (gstats != nullptr && peer == 0) gstats[c.rank * 8 + vi] = t;
    }
}

// GroupNorm(4 groups) coefficients of the layer input from the cluster-wide partial sums (fixed order -> every CTA
// of the cluster computes identical values).  unet.py:36-37: eps 1e-6, biased variance.
__device__ __forceinline__ void gn_coeffs(const Ctx &c, const LayerCfg &L, uint32_t rec_saddr) {
    if (c.tid < L.cin) {
        const int ch = c.tid, cc = ch & 7;
        const uint32_t st = c.smem + kOffStats + (uint32_t)((ch < 8 ? L.in_a : L.in_b) * kCl * 8) * 4u;
        float s = 0.0f, ss = 0.0f;
        if (L.cin == 8) {
            const int p = cc >> 1;
#pragma unroll
            for (int r = 0; r < kCl; ++r) { s += lds1(st + (uint32_t)(r * 8 + 2 * p) * 4u); ss += lds1(st + (uint32_t)(r * 8 + 2 * p + 1) * 4u); }
        } else {
            const int p = (cc >> 2) * 2;
#pragma unroll
            for (int r = 0; r < kCl; ++r) {
                s += lds1(st + (uint32_t)(r * 8 + 2 * p) * 4u) + lds1(st + (uint32_t)(r * 8 + 2 * p + 2) * 4u);
                ss += lds1(st + (uint32_t)(r * 8 + 2 * p + 1) * 4u) + lds1(st + (uint32_t)(r * 8 + 2 * p + 3) * 4u);
            }
        }
        const float inv_cnt = 1.0f / ((L.half ? 32.0f * 64.0f : 64.0f * 128.0f) * (L.cin == 8 ? 2.0f : 4.0f));
        const float mean = s * inv_cnt;
        const float var = fmaxf(ss * inv_cnt - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + 1e-6f);
        const float gamma = lds1(rec_saddr + (uint32_t)(kClRecGamma + ch) * 4u), beta = lds1(rec_saddr + (uint32_t)(kClRecBeta + ch) * 4u);
        sts1(MISC_ADDR(c, ga) + (uint32_t)ch * 4u, gamma * rstd);
        sts1(MISC_ADDR(c, gb) + (uint32_t)ch * 4u, beta - mean * gamma * rstd);
    }
}

// ------------------------------------------------------------------------------------------------
// One tensor-core layer as a two-phase pipeline over the band's output rows (A = upper half, B = lower half):
//   stage rows of A -> [MMAs of A | stage the remaining rows] -> [MMAs of B | epilogue A] -> epilogue B
// Epilogue: bias, residual / nin_shortcut, raw output into this CTA's band, GroupNorm partial sums to the cluster.
// ------------------------------------------------------------------------------------------------
template <int CG, bool HALF, bool UP, bool GN>
__device__ __forceinline__ void conv_layer(Ctx &c, const LayerCfg &L, uint32_t rec_saddr, float *gout, float *gstats) {
    constexpr int PXW = HALF ? 64 : 128, RW = PXW + 2, R = HALF ? 4 : 8, NR = R + 2, HRES = HALF ? 32 : 64;
    constexpr int NRA = R / 2 + 2;                                    // input rows feeding the output rows of phase A
    constexpr int PROWS = CG == 1 ? 10 : 5;
    constexpr int NG = CG == 1 ? 2 : NR / 2;                          // staging groups: cin 8: A | B; cin 16: pairs of rows
    constexpr int GA = CG == 1 ? 0 : NRA / 2 - 1;                     // group whose completion finishes phase A
    constexpr uint32_t kPlane = PROWS * kRowPx * 16u;
    const int y0 = (int)c.rank * R;

#pragma unroll 1
    for (int g = 0; g < NG; ++g) {      // not unrolled: the kernel is instruction-cache bound (r02c: 14 % no-instruction stalls)
        const int i0 = CG == 1 ? (g == 0 ? 0 : NRA) : 2 * g, i1 = CG == 1 ? (g == 0 ? NRA : NR) : 2 * g + 2;
        if (CG == 2 && g >= 2) {   // the slots of group g were read by the MMAs of group g - 2
            mbar_wait(MISC_ADDR(c, mma_bar) + (uint32_t)(g & 1) * 8u, c.ph_mma[g & 1]);
            c.ph_mma[g & 1] ^= 1u;
        }
        if (!(c.dbg & 2)) {
            if (CG == 1) {
                if (g == 0) stage_rows<CG, HALF, UP, GN, (NRA * RW + kThreads - 1) / kThreads>(c, L, i0, i1);
                else stage_rows<CG, HALF, UP, GN, ((NR - NRA) * RW + kThreads - 1) / kThreads>(c, L, i0, i1);
            } else {
                stage_rows<CG, HALF, UP, GN, (2 * RW + kThreads - 1) / kThreads>(c, L, i0, i1);
            }
        }
        if (g == 0) GC_TRACE(c, 4);
        fence_async_smem();
        if (g == 0) GC_TRACE(c, 5);
        tc_fence_before();
        __syncthreads();
        if (g == 0) GC_TRACE(c, 6);
        if (g == NG - 1) GC_TRACE(c, 7);
        if (c.warp == 0) {         // warp-uniform branch; one elected lane issues
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc_tf32(128, 32);
            const int iend = (c.dbg & 1) ? i0 : i1;
            for (int i = i0; i < iend; ++i) {
                const int gy = y0 - 1 + i;
                if (gy < 0 || gy >= HRES) continue;
                const int slot = CG == 1 ? i : (((i >> 1) & 1) * 2 + (i & 1));
#pragma unroll
                for (int cgi = 0; cgi < CG; ++cgi) {
                    const uint64_t a0 = make_desc(c.smem + kOffOper + (uint32_t)(cgi * 2) * kPlane + (uint32_t)(slot * kRowPx) * 16u, kPlane, 128u);
                    const uint64_t b0 = make_desc(rec_saddr + (uint32_t)(cgi * 2 * 4 * 8) * 16u, 512u, 128u);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)   // +1 in the descriptor's address field = one 16-byte pixel; B: 2048 B per kx
                        mma_tf32_elect(c.tmem + 8u * i, a0 + (uint64_t)kx, b0 + (uint64_t)(kx * 128), idesc);
                }
            }
            if (CG == 2) commit_elect(MISC_ADDR(c, mma_bar) + (uint32_t)(g & 1) * 8u);
            if (g == GA) commit_elect(MISC_ADDR(c, done_bar));
            if (g == NG - 1) commit_elect(MISC_ADDR(c, done_bar) + 8u);
            if (g == NG - 1) GC_TRACE(c, 8);
        }
    }
    if (CG == 2) {   // consume the two ring commits nobody waited for (keeps the phase bookkeeping in step)
#pragma unroll
        for (int g = NG - 2; g < NG; ++g) {
            mbar_wait(MISC_ADDR(c, mma_bar) + (uint32_t)(g & 1) * 8u, c.ph_mma[g & 1]);
            c.ph_mma[g & 1] ^= 1u;
        }
    }

    // ---- epilogue, phase by phase: thread = pixel x of NROW consecutive output rows ----
    constexpr int NROW = 1;                            // one row per thread and phase: 4 (full) / 2 (half) row groups of 4 warps
    const int q = c.warp & 3, hsel = c.warp >> 2;      // hsel 0..3
    const int px = q * 32 + c.lane;
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q8[i] = 0.0f;
    float bias[8];
    {
        const float4 b0 = lds4(rec_saddr + kClRecBias * 4u), b1 = lds4(rec_saddr + kClRecBias * 4u + 16u);
        bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w; bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    }
    const uint32_t out_base = c.smem + kOffF + buf_off(L.out);
    const uint32_t ra_base = c.smem + kOffF + buf_off(L.res_a), rb_base = c.smem + kOffF + buf_off(L.res_b);
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
        mbar_wait(MISC_ADDR(c, done_bar) + (uint32_t)ph * 8u, c.ph_done[ph]);
        c.ph_done[ph] ^= 1u;
        tc_fence_after();
        GC_TRACE(c, 9 + 2 * ph);
        if (px >= PXW || hsel >= R / 2 || (c.dbg & 4)) continue;   // warp-uniform (half resolution: 4 of the 16 warps work)
        const int r0 = ph * (R / 2) + hsel;
        float acc[NROW * 8];
        const uint32_t ta = c.tmem + ((uint32_t)(q * 32) << 16) + 8u * (r0 + 2);
        {
            uint32_t rr[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                         : "r"(ta));
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(rr[i]);
        }
        tmem_wait_ld();
#pragma unroll
        for (int r = 0; r < NROW; ++r) {
            float *a = acc + r * 8;
            const uint32_t poff = (uint32_t)(((r0 + r) * 2) * PXW + px) * 16u;
#pragma unroll
            for (int o = 0; o < 8; ++o) a[o] += bias[o];
            if (L.res != kNone) {
                const float4 a0 = lds4(ra_base + poff), a1 = lds4(ra_base + poff + PXW * 16u);
                if (L.res == kIdent) {
                    a[0] += a0.x; a[1] += a0.y; a[2] += a0.z; a[3] += a0.w;
                    a[4] += a1.x; a[5] += a1.y; a[6] += a1.z; a[7] += a1.w;
                } else {
                    const float4 b0 = lds4(rb_base + poff), b1 = lds4(rb_base + poff + PXW * 16u);
                    const float xr[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                                          b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    const float4 n0 = lds4(rec_saddr + kClRecNinB * 4u), n1 = lds4(rec_saddr + kClRecNinB * 4u + 16u);
                    float sh[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
                    for (int ci = 0; ci < 16; ++ci) {
                        const float4 w0 = lds4(rec_saddr + (uint32_t)(kClRecNinW + ci * 8) * 4u);
                        const float4 w1 = lds4(rec_saddr + (uint32_t)(kClRecNinW + ci * 8 + 4) * 4u);
                        sh[0] = fmaf(xr[ci], w0.x, sh[0]); sh[1] = fmaf(xr[ci], w0.y, sh[1]);
                        sh[2] = fmaf(xr[ci], w0.z, sh[2]); sh[3] = fmaf(xr[ci], w0.w, sh[3]);
                        sh[4] = fmaf(xr[ci], w1.x, sh[4]); sh[5] = fmaf(xr[ci], w1.y, sh[5]);
                        sh[6] = fmaf(xr[ci], w1.z, sh[6]); sh[7] = fmaf(xr[ci], w1.w, sh[7]);
                    }
#pragma unroll
                    for (int o = 0; o < 8; ++o) a[o] += sh[o];
                }
            }
            const float4 o0 = make_float4(a[0], a[1], a[2], a[3]), o1 = make_float4(a[4], a[5], a[6], a[7]);
            sts4(out_base + poff, o0);
            sts4(out_base + poff + PXW * 16u, o1);
            if (gout != nullptr) {   // last layer of the unfused variant: the raw tensor also goes to global memory (NHWC8)
                float4 *dst = reinterpret_cast<float4 *>(gout + ((size_t)(y0 + r0 + r) * PXW + px) * 8);
                dst[0] = o0;
                dst[1] = o1;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                q8[2 * p] += a[2 * p] + a[2 * p + 1];
                q8[2 * p + 1] += a[2 * p] * a[2 * p] + a[2 * p + 1] * a[2 * p + 1];
            }
        }
    }
    GC_TRACE(c, 12);
    tc_fence_before();
    push_stats(c, q8, L.out, gstats);
    GC_TRACE(c, 13);
}

// down.0.downsample (unet.py:59-78): pad (0,1,0,1) + 3x3 stride 2, no GroupNorm, on CUDA cores (thread = output pixel).
__device__ __forceinline__ void down_layer(Ctx &c, const LayerCfg &L, uint32_t rec_saddr) {
    const bool act = c.tid < 256;                      // 4 rows x 64 output pixels; the other warps only join the reduction
    const int r = (c.tid >> 6) & 3, ox = c.tid & 63;
    float acc[8];
    {
        const float4 b0 = lds4(rec_saddr + kClRecBias * 4u), b1 = lds4(rec_saddr + kClRecBias * 4u + 16u);
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        int sy = 2 * r + ky;                              // full-resolution row relative to this CTA's band (0..8)
        uint32_t srank = c.rank;
        if (sy >= 8) { srank = c.rank + 1; sy -= 8; }
        const bool row_ok = srank < (uint32_t)kCl;        // row 64 is the zero padding
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int sx = 2 * ox + kx;
            if (!act || !row_ok || sx >= 128) continue;
            const uint32_t a = mapa(c.smem + kOffF + buf_off(L.in_a) + (uint32_t)((sy * 2) * 128 + sx) * 16u, srank);
            const float4 v0 = ld_cluster4(a), v1 = ld_cluster4(a + 128u * 16u);
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            const uint32_t w = rec_saddr + (uint32_t)((ky * 3 + kx) * 64) * 4u;   // [tap][cin][cout]
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const float4 w0 = lds4(w + (uint32_t)(ci * 8) * 4u), w1 = lds4(w + (uint32_t)(ci * 8 + 4) * 4u);
                acc[0] = fmaf(vv[ci], w0.x, acc[0]); acc[1] = fmaf(vv[ci], w0.y, acc[1]);
                acc[2] = fmaf(vv[ci], w0.z, acc[2]); acc[3] = fmaf(vv[ci], w0.w, acc[3]);
                acc[4] = fmaf(vv[ci], w1.x, acc[4]); acc[5] = fmaf(vv[ci], w1.y, acc[5]);
                acc[6] = fmaf(vv[ci], w1.z, acc[6]); acc[7] = fmaf(vv[ci], w1.w, acc[7]);
            }
        }
    }
    const uint32_t poff = (uint32_t)((r * 2) * 64 + ox) * 16u, out_base = c.smem + kOffF + buf_off(L.out);
    float q8[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) q8[p] = 0.0f;
    if (act) {
        sts4(out_base + poff, make_float4(acc[0], acc[1], acc[2], acc[3]));
        sts4(out_base + poff + 64u * 16u, make_float4(acc[4], acc[5], acc[6], acc[7]));
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            q8[2 * p] = acc[2 * p] + acc[2 * p + 1];
            q8[2 * p + 1] = acc[2 * p] * acc[2 * p] + acc[2 * p + 1] * acc[2 * p + 1];
        }
    }
    push_stats(c, q8, L.out, nullptr);
}

// The 26 middle layers of one UNet evaluation.  rec_g: this step's layer records in global memory.  On entry every
// thread has ARRIVED at the cluster barrier that publishes the input tensor (F0 + its statistics); on exit likewise for
// the output of layer 25.
__device__ __forceinline__ void middle_layers(Ctx &c, const float *rec_g, float *gout, float *gstats) {
    if (c.tid == 0) bulk_load(c.smem + kOffRec, rec_g, kClRecBytes, MISC_ADDR(c, rec_bar));
#pragma unroll 1
    for (int l = 0; l < kClLayers; ++l) {
        const LayerCfg L = c_layers[l];
        const int rb = l & 1;
        c.trace_row = l;
        GC_TRACE(c, 0);
        // record of the next layer: its buffer was last read during layer l - 1, which this CTA has finished
        if (c.tid == 0 && l + 1 < kClLayers)
            bulk_load(c.smem + kOffRec + (uint32_t)((rb ^ 1) * kClRecBytes), rec_g + (size_t)(l + 1) * kClRecFloats, kClRecBytes,
                      MISC_ADDR(c, rec_bar) + (uint32_t)(rb ^ 1) * 8u);
        // zero the accumulator columns (every thread finished reading TMEM before it arrived at the cluster barrier)
        if (L.kind != kDownK && c.warp < 4) {
            const uint32_t ta = c.tmem + ((uint32_t)(c.warp * 32) << 16);
#pragma unroll
            for (int j = 0; j < 13; ++j) tmem_zero8(ta + 8u * j);
            tmem_wait_st();
        }
        GC_TRACE(c, 1);
        mbar_wait(MISC_ADDR(c, rec_bar) + (uint32_t)rb * 8u, c.ph_rec[rb]);
        c.ph_rec[rb] ^= 1u;
        GC_TRACE(c, 2);
        // one cluster barrier per layer: the producer's raw rows + statistics are visible, and every CTA is done
        // reading what this layer is about to overwrite
        if (!(c.dbg & 8)) cluster_wait(); else __syncthreads();
        GC_TRACE(c, 3);
        const uint32_t rec_saddr = c.smem + kOffRec + (uint32_t)(rb * kClRecBytes);
        if (L.gn) {
            gn_coeffs(c, L, rec_saddr);
            __syncthreads();
        }
        GC_TRACE(c, 14);
        const bool last = l == kClLayers - 1;
        if (L.kind == kDownK) {
            down_layer(c, L, rec_saddr);
        } else if (L.kind == kUpK) {
            conv_layer<1, false, true, false>(c, L, rec_saddr, nullptr, nullptr);
        } else if (L.half) {
            if (L.cin == 8) conv_layer<1, true, false, true>(c, L, rec_saddr, nullptr, nullptr);
            else conv_layer<2, true, false, true>(c, L, rec_saddr, nullptr, nullptr);
        } else {
            if (L.cin == 8) conv_layer<1, false, false, true>(c, L, rec_saddr, last ? gout : nullptr, last ? gstats : nullptr);
            else conv_layer<2, false, false, true>(c, L, rec_saddr, nullptr, nullptr);
        }
        if (!(c.dbg & 8)) cluster_arrive();
        GC_TRACE(c, 15);
    }
}

// ------------------------------------------------------------------------------------------------
// Unfused variant: h0 = conv_in output (NHWC8, global) -> 26 layers -> input of norm_out (NHWC8, global) + its
// per-band GroupNorm partial sums [A][8 bands][8] for k_conv_out_tc.  One cluster per agent (grid-stride).
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads, 1)
k_unet_middle_cluster(const float *__restrict__ h0, const float *__restrict__ rec_g, float *__restrict__ out,
                      float *__restrict__ stats_out, int n_agents, int dbg, long long *trace) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Ctx c;
    c.smem = smem_u32(smem_raw);
    c.tid = threadIdx.x;
    c.warp = __shfl_sync(0xffffffffu, c.tid >> 5, 0);   // provably warp-uniform (the MMA warp branches on it)
    c.lane = c.tid & 31;
    c.rank = blockIdx.x % kCl;
    c.dbg = dbg;
    c.trace = (blockIdx.x == 0) ? trace : nullptr;
    c.trace_row = 0;
    c.ph_mma[0] = c.ph_mma[1] = c.ph_done[0] = c.ph_done[1] = c.ph_rec[0] = c.ph_rec[1] = 0u;
    const int n_clusters = gridDim.x / kCl, cluster_id = blockIdx.x / kCl;
    Misc *misc = reinterpret_cast<Misc *>(smem_raw + kOffMisc);

    if (c.warp == 0) tmem_alloc512(&misc->tmem);
    if (c.tid == 32) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(MISC_ADDR(c, mma_bar) + 8u * i, 1); mbar_init(MISC_ADDR(c, done_bar) + 8u * i, 1);
            mbar_init(MISC_ADDR(c, rec_bar) + 8u * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = __shfl_sync(0xffffffffu, misc->tmem, 0);
    cluster_arrive();
    cluster_wait();

    for (int agent = cluster_id; agent < n_agents; agent += n_clusters) {
        // ---- load this CTA's band of h0 into F0 and publish its GroupNorm partial sums ----
        {
            const int q = c.warp & 3, hsel = c.warp >> 2, px = q * 32 + c.lane, r0 = hsel * 2;
            float q8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) q8[i] = 0.0f;
            float4 v[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float4 *src = reinterpret_cast<const float4 *>(h0 + (((size_t)agent * 64 + c.rank * 8 + r0 + r) * 128 + px) * 8);
                v[r][0] = __ldg(src);
                v[r][1] = __ldg(src + 1);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t poff = (uint32_t)(((r0 + r) * 2) * 128 + px) * 16u;
                sts4(c.smem + kOffF + buf_off(F0) + poff, v[r][0]);
                sts4(c.smem + kOffF + buf_off(F0) + poff + 128u * 16u, v[r][1]);
                const float a[8] = {v[r][0].x, v[r][0].y, v[r][0].z, v[r][0].w, v[r][1].x, v[r][1].y, v[r][1].z, v[r][1].w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    q8[2 * p] += a[2 * p] + a[2 * p + 1];
                    q8[2 * p + 1] += a[2 * p] * a[2 * p] + a[2 * p + 1] * a[2 * p + 1];
                }
            }
            push_stats(c, q8, F0, nullptr);
        }
        cluster_arrive();
        middle_layers(c, rec_g, out + (size_t)agent * 64 * 128 * 8, stats_out + (size_t)agent * kCl * 8);
        // the next agent's load overwrites F0, which the neighbours read as a halo during layer 24
        cluster_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (c.warp == 0) tmem_free512(c.tmem);
}

}  // namespace cl

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool unet_cluster_eligible(int C, int H, int W) {
    (void)C;
    return H == 64 && W == 128;
}

static int cluster_grid(const void *kernel, size_t smem, int n_agents, int *n_clusters) {
    static int cached = -1;
    if (cached < 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_error("unet cluster kernel: cudaFuncSetAttribute failed (%d)", (int)e); return (int)e; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCl, 1, 1);
        cfg.blockDim = dim3(cl::kThreads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
        if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); set_error("unet cluster kernel: no active cluster fits (%d)", (int)e); return e != cudaSuccess ? (int)e : GC_EUNSUPPORTED; }
        cached = n;
    }
    *n_clusters = n_agents < cached ? n_agents : cached;
    return GC_OK;
}

int unet_middle_cluster(cudaStream_t st, int A, const float *h0, const float *rec_dev, float *out, float *stats_out) {
    int n_clusters = 0;
    if (int rc = cluster_grid((const void *)cl::k_unet_middle_cluster, cl::kSmemBytes, A, &n_clusters)) return rc;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("GC_CL_DEBUG"); dbg = e ? atoi(e) : 0; }
    static long long *trace = nullptr;
    if ((dbg & 32) && trace == nullptr) {
        cudaMalloc(&trace, kClLayers * 16 * sizeof(long long));
        cudaMemset(trace, 0, kClLayers * 16 * sizeof(long long));
    }
    cl::k_unet_middle_cluster<<<n_clusters * kCl, cl::kThreads, cl::kSmemBytes, st>>>(h0, rec_dev, out, stats_out, A, dbg, trace);
    if (trace != nullptr) {   // debug only: dump the timeline of the last agent CTA 0 processed
        cudaStreamSynchronize(st);
        long long h[kClLayers * 16];
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        for (int l = 0; l < kClLayers; ++l) {
            fprintf(stderr, "trace layer %2d:", l);
            for (int k = 1; k < 16; ++k) fprintf(stderr, " %lld", h[l * 16 + k] ? h[l * 16 + k] - h[l * 16] : -1);
            fprintf(stderr, "\n");
        }
    }
    GC_LAUNCH_CHECK("k_unet_middle_cluster");
    return GC_OK;
}

}  // namespace gc
