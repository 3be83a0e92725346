// Row-sweep tcgen05 weight / bias gradient of the 10 -> 10 channel conv layers on the 24-channel fp16 piece layout
// (conv_wgrad_row_tc.cu): TMA tensor-map strips as the B operand, un-pooled output-gradient windows built on the fly as the A operand.
#pragma once
#include "common.cuh"

namespace cpp {
namespace wgr {

bool shape_ok(int H, int W, int KS);        // even H, W; W + KS - 1 <= 128; KS 5 or 3
int64_t scratch_bytes(int H, int W, int KS);      // per-CTA partials; 0 when the shape is not covered
// dw [KS][KS][10][10], db [10] of one network; x_pieces fp16 [B][H][W][24]; d_pooled fp32 [B][H/2][W/2][10] + arg-max side band;
// gmax: device float holding max |d_pooled| (final before the launch)
int launch(const void* x_pieces, const float* d_pooled, const uint8_t* amax, const float* gmax, int B, int H, int W, int KS, float* dw, float* db,
           void* scratch, cudaStream_t s);

}  // namespace wgr
}  // namespace cpp
