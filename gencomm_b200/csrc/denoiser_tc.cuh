// Interface between denoiser.cu (sampler orchestration, fp32 kernels) and denoiser_tc.cu (tcgen05 kernels).
#pragma once
#include <cuda_runtime.h>

namespace gc {

struct Bias8 { float b[8]; };
struct Affine8 { float gamma[8], beta[8]; };

// conv_in on tensor cores (unet.py:315).  Eligible when W % 128 == 0 and C % 8 == 0.
bool conv_in_tc_eligible(int C, int H, int W);
int conv_in_tc_tiles(int H, int W);   // number of GroupNorm partial-sum tiles per agent it writes
int conv_in_tc(cudaStream_t st, int A, const float *cond, const float *x, const float *w, const Bias8 &bias, int C, int H,
               int W, float *out, float *stats_out);

// norm_out + swish + conv_out (+ sampler update) on tensor cores (unet.py:341-343).  Eligible when
// W % 128 == 0 and C % 64 == 0.
bool conv_out_tc_eligible(int C, int H, int W);
int conv_out_tc(cudaStream_t st, int A, const float *in, const float *st_in, int tiles_in, const float *w, const float *bias,
                const Affine8 &aff, int C, int H, int W, int mode, float c1, float c2, float sigma, const float *noise,
                float *x, float *pred, int materialize);

}  // namespace gc
