// clock64 timeline of one CTA of k_conv_rows at the backbone's level-1 / level-2 shapes (arbitrary operand contents).
#define CR_TRACE 1
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../gencomm_b200/csrc/conv_rows.cu"
namespace gc { void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); } }

static void run(int A, int C, int NOUT, int H, int W) {
    const size_t plane = (size_t)A * H * W * C * 2, wbytes = (size_t)9 * C * NOUT * 4 + 65536, obytes = (size_t)A * H * W * NOUT * 2;
    void *xh, *xl, *w, *oh, *ol; float *bias;
    cudaMalloc(&xh, plane); cudaMalloc(&xl, plane); cudaMalloc(&w, wbytes); cudaMalloc(&oh, obytes); cudaMalloc(&ol, obytes);
    cudaMalloc(&bias, 4096);
    cudaMemset(xh, 0x3c, plane); cudaMemset(xl, 0x30, plane); cudaMemset(w, 0x38, wbytes); cudaMemset(bias, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        int rc = gc::conv_rows(0, A, xh, xl, w, bias, C, NOUT, H, W, NOUT, 0, nullptr, oh, ol);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (rc || e != cudaSuccess) { printf("failed rc=%d %s\n", rc, cudaGetErrorString(e)); exit(1); }
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long t[64];
    cudaMemcpyFromSymbol(t, gc::cr::g_trace, sizeof(t));
    printf("A=%d C=%d NOUT=%d %dx%d: %.1f us\n", A, C, NOUT, H, W, ms * 1e3f);
    // 1 + g: chunk g staged; 16 + 2g / 17 + 2g: MMAs of chunk g start / issued; 40 + 2k / 41 + 2k: epilogue of tile k start / end
    for (int g = 0; g < 12; ++g) if (t[1 + g]) printf("  chunk %2d staged   %8lld\n", g, t[1 + g] - t[0]);
    for (int g = 0; g < 12; ++g) if (t[16 + 2 * g]) printf("  chunk %2d mma      %8lld .. %8lld\n", g, t[16 + 2 * g] - t[0], t[17 + 2 * g] - t[0]);
    for (int k = 0; k < 12; ++k) if (t[40 + 2 * k]) printf("  tile  %2d epilogue %8lld .. %8lld\n", k, t[40 + 2 * k] - t[0], t[41 + 2 * k] - t[0]);
    cudaMemset(nullptr, 0, 0);
    { long long z[64] = {0}; cudaMemcpyToSymbol(gc::cr::g_trace, z, sizeof(z)); }
    cudaFree(xh); cudaFree(xl); cudaFree(w); cudaFree(oh); cudaFree(ol); cudaFree(bias);
}
int main() {
    run(32, 64, 64, 128, 256);
    cudaMemset(nullptr, 0, 0);
    run(32, 128, 128, 64, 128);
    return 0;
}
