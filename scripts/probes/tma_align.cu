// Probe: does cp.async.bulk.tensor (tile mode, f32, no swizzle) accept an innermost start coordinate that is not
// a multiple of 4 elements (16 bytes)?   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_align tma_align.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

constexpr int BW = 20, BH = 19, BC = 2;

__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int c, float *out) {
    __shared__ __align__(128) float box[BC * BH * BW];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(box);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)sizeof(box)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(d), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(c), "r"(b) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W;\n}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < BC * BH * BW; i += blockDim.x) out[i] = box[i];
}

int main() {
    const int W = 64, H = 48, P = 4;
    std::vector<float> h(W * H * P);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BC * BH * BW * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
    CUtensorMap map;
    cuuint64_t gdim[3] = {W, H, P}, gstr[2] = {W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {BW, BH, BC}, es[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    const int xs[] = {0, 4, 1, 2, 3, 5, -1, -3, 50, 61}, ys[] = {0, 3, -2, 40};
    std::vector<float> res(BC * BH * BW);
    for (int x : xs) for (int y : ys) {
        k<<<1, 128>>>(map, x, y, 1, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("x=%d y=%d: CUDA error %s\n", x, y, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < BC; ++c) for (int j = 0; j < BH; ++j) for (int i = 0; i < BW; ++i) {
            const int gx = x + i, gy = y + j, gc = 1 + c;
            const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H && gc < P) ? h[((size_t)gc * H + gy) * W + gx] : 0.0f;
            if (res[(c * BH + j) * BW + i] != want) ++bad;
        }
        printf("x=%3d y=%3d: %s (%d mismatches)\n", x, y, bad ? "MISMATCH" : "ok", bad);
    }
    return 0;
}
