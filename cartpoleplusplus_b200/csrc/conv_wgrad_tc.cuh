// Weight / bias gradients of the conv layers of the reference trunk (Network.simple_conv_net_on, base_network.py:103-123, differentiated by
// tf.gradients in ddpg_cartpole.py:111,213 / naf_cartpole.py:233) on the 5th-generation tensor cores: tcgen05.mma with BOTH
// operands MN-major (the reduction dimension is the pixel index), accumulators in TMEM.  See conv_wgrad_tc.cu for the design.
#pragma once
#include "common.cuh"

namespace cpp {
namespace wgtc {

constexpr int kMaxNets = 2;        // sibling networks side by side along N (2 nets x 2 pieces x 10 filters = 40 -> N = 48)

// pieces = 0: conv1 on raw fp16 pixels - 5x5 layers whose five-pixel window (5 * C channel values + one block of inside-the-image
// flags) fits 64 rows of the MMA tile (C <= 11), 16 | W <= 64, 16-byte image rows: conv1 of BASELINE config c3 (64x64x9).
// pieces = 1: conv2 / conv3 on the 24-channel fp16 piece layout of conv_tc.cuh (C == 24, one network, KS 5 or 3, 16 | W <= 64).
bool supported(int nets, int H, int W, int C, int KS, int pieces = 0);
int64_t scratch_bytes(int nets, int H, int W, int C, int KS, int pieces = 0);      // 0 when unsupported
// gmax[n]: device pointer to max |d_pooled[n]| (wg::launch_conv_wgrad_mma computes it or takes it from the dgrad epilogue)
int launch(const void* x_f16, const float* mean_inv, int nets, const float* const* d_pooled, const uint8_t* const* amax, int B, int H,
           int W, int C, int KS, float* const* dw, float* const* db, const float* const* gmax, void* scratch, cudaStream_t s,
           int pieces = 0);

}  // namespace wgtc
}  // namespace cpp
