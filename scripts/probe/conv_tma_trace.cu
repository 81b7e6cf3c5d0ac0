// Where a CTA of k_conv_tma spends its cycles (feeder waiting for operands / for a free accumulator, producer waiting for a free
// ring slot, epilogue waiting for a finished tile) at the backbone's shapes.  Arbitrary operand contents.
#define CT_TRACE 1
#include <cstdarg>
#include <cstdio>
#include "../../gencomm_b200/csrc/conv_tma.cuh"
namespace gc { void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); } }
using namespace gc;

template <int NOUT, int TAPS, int EPI>
static void run(const char *what, int A, int C, int H, int W, int stride, int up) {
    const int Hi = H * stride, Wi = W * stride;
    const size_t plane = (size_t)A * Hi * Wi * C * 2, wbytes = (size_t)TAPS * C * NOUT * 4 + 65536;
    const size_t obytes = (size_t)A * H * W * up * up * 384 * 2;
    void *xh, *xl, *w, *oh, *ol; float *bias;
    cudaMalloc(&xh, plane); cudaMalloc(&xl, plane); cudaMalloc(&w, wbytes); cudaMalloc(&oh, obytes); cudaMalloc(&ol, obytes);
    cudaMalloc(&bias, 4096);
    cudaMemset(xh, 0x3c, plane); cudaMemset(xl, 0x30, plane); cudaMemset(w, 0x38, wbytes); cudaMemset(bias, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 3; ++it) {
        long long z[16] = {0};
        cudaMemcpyToSymbol(ct::g_trace, z, sizeof(z));
        cudaEventRecord(e0);
        int rc = ct::launch_conv_tma<NOUT, TAPS, EPI>(0, A, (const uint4 *)xh, (const uint4 *)xl, (const uint4 *)w, bias, C, C, H, W, Hi, Wi,
                                                      stride, NOUT, up > 1 ? 384 : NOUT, 0, nullptr, (uint4 *)oh, (uint4 *)ol, up, 0, 0);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (rc || e != cudaSuccess) { printf("failed rc=%d %s\n", rc, cudaGetErrorString(e)); exit(1); }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long t[16];
    cudaMemcpyFromSymbol(t, ct::g_trace, sizeof(t));
    printf("%-44s %7.1f us | CTA 8: %lld cycles, %lld stages (%.0f / stage); feeder waits: operands %lld, accumulator %lld; producer waits "
           "for a slot %lld; epilogue waits for a tile %lld\n", what, ms * 1e3f, t[4], t[5], (double)t[4] / (double)(t[5] ? t[5] : 1), t[0], t[1],
           t[2], t[3]);
    cudaFree(xh); cudaFree(xl); cudaFree(w); cudaFree(oh); cudaFree(ol); cudaFree(bias);
}
int main() {
    run<256, 9, 5>("level 3: 256->256 3x3 at 32x64", 32, 256, 32, 64, 1, 1);
    run<128, 9, 5>("level 2 first: 64->128 3x3 s2 -> 64x128", 32, 64, 64, 128, 2, 1);
    run<128, 9, 5>("shrink conv1: 384->128 3x3 s2 -> 64x128", 32, 384, 64, 128, 2, 1);
    run<64, 9, 5>("level 1 first: 64->64 3x3 s2 -> 128x256", 32, 64, 128, 256, 2, 1);
    run<128, 1, 5>("deblock 3 phase: 256->128 1x1 up 4", 32, 256, 32, 64, 1, 4);
    run<128, 1, 5>("deblock 2 phase: 128->128 1x1 up 2", 32, 128, 64, 128, 1, 2);
    run<128, 1, 5>("deblock 1: 64->128 1x1", 32, 64, 128, 256, 1, 1);
    return 0;
}
