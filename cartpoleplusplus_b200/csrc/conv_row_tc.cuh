// Row-sweep tcgen05 convolution of the 10 -> 10 channel layers (conv_row_tc.cu): conv2 / conv3 forward and input gradient on the
// 24-channel fp16 piece layout, input strips by TMA tensor maps.  Same scratch / phase contract as conv_tc.cuh.
#pragma once
#include "common.cuh"

namespace cpp {
namespace tcr {

void set_conv_row(int on);            // cpp_set_option("conv_row")
int conv_row_enabled();
bool shape_ok(int H, int W, int KS);  // even H, W; W + KS - 1 <= 128
bool supported(int H, int W, int KS); // shape_ok and enabled
bool fused_unpool(int H, int W, int KS);   // supported, and the input gradient builds its strips from d(pooled) + arg-max itself
int64_t scratch_bytes(int H, int W, int KS);
// between begin and flush every launch(..., kPhasePrep) is collected; flush packs the weights of all of them in ONE kernel on `s`
void prep_batch_begin();
int prep_batch_flush(cudaStream_t s);
int piece_overflow_count(int reset, unsigned int* out);
// dgrad = 0: out = pooled fp32 [B][H/2][W/2][10], amax, optional out_hl (piece copy); y = maxpool2x2(relu(conv_same(x, w) + bias))
// dgrad = 1: out = dense fp32 [B][H][W][10] = conv_same(x, flip(w)^T) * *out_scale; optional max|out| into *out_absmax (atomic max)
// x_pieces fp16 [B][H][W][24]; phase as tc::kPhase*
int launch(const void* x_pieces, const float* w, const float* bias, int B, int H, int W, int KS, int dgrad, float* out, uint8_t* amax,
           __half* out_hl, const float* out_scale, float* out_absmax, void* scratch, cudaStream_t s, int phase,
           // fused un-pool (dgrad only; x_pieces and out_scale unused): d(pooled) fp32 [B][H/2][W/2][10], the arg-max side band, max|d(pooled)|
           // (device float, must be final before the launch), optional 1/scale output
           const float* unpool_gp = nullptr, const uint8_t* unpool_amax = nullptr, const float* unpool_gmax = nullptr,
           float* unpool_inv_scale = nullptr);

}  // namespace tcr
}  // namespace cpp
