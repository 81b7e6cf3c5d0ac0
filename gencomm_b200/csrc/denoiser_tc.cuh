// Interface between denoiser.cu (sampler orchestration, fp32 kernels) and denoiser_tc.cu (tcgen05 kernels).
#pragma once
#include <cuda_runtime.h>

namespace gc {

struct C8Params {                  // passed by value as a __grid_constant__ kernel parameter
    float w[9][16][8];             // [tap][cin][cout]
    float bias[8];                 // conv bias (+ temb projection for conv1 of a ResnetBlock, per step)
    float gamma[16], beta[16];     // GroupNorm affine of the (concatenated) input
    float nin_w[16][8];            // 1x1 nin_shortcut [cin][cout]
    float nin_b[8];
};
static_assert(sizeof(C8Params) == 1328 * 4, "C8Params layout is part of the packed weight format");

enum Geom { kSame = 0, kDown = 1, kUp = 2 };
enum Res { kNone = 0, kIdent = 1, kNin = 2 };

struct Bias8 { float b[8]; };
struct Affine8 { float gamma[8], beta[8]; };

// bf16 B operands of both convolutions, repacked once per call into `packed` (conv_tc_packed_bytes(C) bytes).
size_t conv_tc_packed_bytes(int C);
int conv_tc_pack_weights(cudaStream_t st, const float *w_in, const float *w_out, int C, void *packed);

// conv_in on tensor cores (unet.py:315).  Eligible when W % 128 == 0 and C % 8 == 0.
bool conv_in_tc_eligible(int C, int H, int W);
int conv_in_tc_tiles(int H, int W);   // number of GroupNorm partial-sum tiles per agent it writes
int conv_in_tc(cudaStream_t st, int A, const float *cond, const float *x, const void *packed, const Bias8 &bias, int C, int H,
               int W, float *out, float *stats_out);

// conv_in, input-row-stationary version (8 rows per CTA, 3 MMAs per staged row and chunk): W == 128, H % 8 == 0.
bool conv_in_tc2_eligible(int C, int H, int W);
int conv_in_tc2_tiles(int H);
int conv_in_tc2(cudaStream_t st, int A, const float *cond, const float *x, const void *packed, const Bias8 &bias, int C, int H,
                int W, float *out, float *stats_out);

// norm_out + swish + conv_out (+ sampler update) on tensor cores (unet.py:341-343).  Eligible when
// W % 128 == 0 and C % 64 == 0.
bool conv_out_tc_eligible(int C, int H, int W);
int conv_out_tc(cudaStream_t st, int A, const float *in, const float *st_in, int tiles_in, const void *packed, const float *bias,
                const Affine8 &aff, int C, int H, int W, int mode, float c1, float c2, float sigma, const float *noise,
                float *x, float *pred, int materialize);

// Width-8 middle layers (3x3, CIN in {8,16} -> 8, NHWC8 fp32 tensors) as tf32 tcgen05 implicit GEMMs, for
// full-resolution layers with W % 128 == 0 and geometry kSame / kUp.  Returns the number of GroupNorm partial-sum
// tiles per agent it writes through *tiles_out.
bool conv_c8_tc_eligible(int geom, int H, int W);
int conv_c8_tc(cudaStream_t st, int A, int cin, bool pre_gn, int geom, int res, const float *in_a, const float *in_b,
               const float *st_a, const float *st_b, int tiles_a, int tiles_b, const float *res_a, const float *res_b,
               float *out, float *stats_out, int H, int W, const C8Params &prm, int *tiles_out);

}  // namespace gc
