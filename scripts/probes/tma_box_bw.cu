// Probe: sustained TMA box-load throughput per box shape (f32, 3-D tile mode, 1 persistent CTA per SM, S-stage ring,
// consumers only release the stage).  Answers: is a small-row box (96 B rows) slower than a wide one?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_box_bw tma_box_bw.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t par) {
    asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}" ::"r"(bar), "r"(par) : "memory");
}

__global__ void __launch_bounds__(160, 1)
k(const __grid_constant__ CUtensorMap map, int W, int H, int planes, int bw, int bh, int bc, int stages, int box_bytes,
  int boxes_per_stage, int iters, float *sink, int tx_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < 16; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t stage_bytes = (uint32_t)box_bytes * boxes_per_stage;
    if (tid >= 128) {
        if (tid == 128) {
            uint32_t par = 1, rng = blockIdx.x * 2654435761u + 12345u;
            int s = 0;
            for (int it = 0; it < iters; ++it) {
                mwait(s32(&empty[s]), par);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"((uint32_t)tx_bytes * boxes_per_stage) : "memory");
                for (int b = 0; b < boxes_per_stage; ++b) {
                    rng = rng * 1664525u + 1013904223u;
                    const int x = (int)((rng >> 8) % (uint32_t)(W - bw)) & ~3;
                    rng = rng * 1664525u + 1013904223u;
                    const int y = (int)((rng >> 8) % (uint32_t)(H - bh));
                    rng = rng * 1664525u + 1013904223u;
                    const int c = (int)((rng >> 8) % (uint32_t)(planes - bc));
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(s32(smem) + (uint32_t)s * stage_bytes + (uint32_t)b * box_bytes), "l"(reinterpret_cast<uint64_t>(&map)),
                                   "r"(x), "r"(y), "r"(c), "r"(s32(&full[s])) : "memory");
                }
                if (++s == stages) { s = 0; par ^= 1; }
            }
        }
        return;
    }
    uint32_t par = 0;
    int s = 0;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        mwait(s32(&full[s]), par);
        acc += reinterpret_cast<float *>(smem + (size_t)s * stage_bytes)[tid];
        __syncwarp();
        if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        if (++s == stages) { s = 0; par ^= 1; }
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char **argv) {
    const int W = 256, H = 256;
    const int P = argc > 1 ? atoi(argv[1]) : 8192;   // 8192 planes = 2 GiB (HBM resident); 128 planes = 32 MiB (L2 resident)
    printf("tensor %d x %d x %d f32 = %.0f MiB\n", W, H, P, (double)W * H * P * 4 / 1048576.0);
    float *d, *sink;
    cudaMalloc(&d, (size_t)W * H * P * 4); cudaMalloc(&sink, 4);
    cudaMemset(d, 0, (size_t)W * H * P * 4);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Shape { int bw, bh, bc, per_stage; } shapes[] = {
        {24, 19, 4, 3}, {24, 19, 8, 3}, {24, 19, 4, 1}, {16, 8, 8, 1}, {16, 12, 8, 3}, {40, 37, 2, 3}, {52, 46, 2, 3}, {52, 46, 1, 3},
        {28, 24, 4, 3}, {64, 16, 4, 1}, {128, 8, 4, 1}, {256, 4, 4, 1}, {32, 32, 4, 1}, {64, 64, 1, 1}};
    for (auto sh : shapes) {
        CUtensorMap map;
        cuuint64_t gdim[3] = {W, H, P}, gstr[2] = {W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t box[3] = {(cuuint32_t)sh.bw, (cuuint32_t)sh.bh, (cuuint32_t)sh.bc}, es[3] = {1, 1, 1};
        if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("encode failed for %dx%dx%d\n", sh.bw, sh.bh, sh.bc); continue;
        }
        const int box_bytes = (sh.bw * sh.bh * sh.bc * 4 + 127) & ~127;   // smem slot per box (128-byte aligned)
        const int tx_bytes = sh.bw * sh.bh * sh.bc * 4;
        const int stage_bytes = box_bytes * sh.per_stage;
        for (int stages : {4, 8}) {
            if ((long long)stages * stage_bytes > 200 * 1024) continue;
            const int iters = (int)(64ll * 1024 * 1024 / stage_bytes);   // 64 MiB per CTA
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k<<<148, 160, 220 * 1024>>>(map, W, H, P, sh.bw, sh.bh, sh.bc, stages, box_bytes, sh.per_stage, iters / 8, sink, tx_bytes);
            cudaEventRecord(e0);
            k<<<148, 160, 220 * 1024>>>(map, W, H, P, sh.bw, sh.bh, sh.bc, stages, box_bytes, sh.per_stage, iters, sink, tx_bytes);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = 148.0 * iters * (double)tx_bytes * sh.per_stage;
            printf("box %3dx%3dx%d (row %4d B, %6d B/box) x%d per stage, %2d stages (%6.1f KB in flight): %7.1f GB/s  %6.2f Mrows/s/SM\n",
                   sh.bw, sh.bh, sh.bc, sh.bw * 4, box_bytes, sh.per_stage, stages, stages * stage_bytes / 1024.0, bytes / ms / 1e6,
                   148.0 * iters * sh.per_stage * sh.bh * sh.bc / ms / 1e3 / 148.0);
        }
    }
    return 0;
}
