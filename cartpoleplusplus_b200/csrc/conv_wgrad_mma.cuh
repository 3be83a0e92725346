// Weight/bias gradient of one conv layer of the reference trunk (Network.simple_conv_net_on, base_network.py:73-127) for up
// to 3 sibling networks that read the same input, on the tensor cores (warp-level mma.sync m16n8k16, fp16 x fp16 -> fp32).
// See conv_wgrad_mma.cu for the design.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"

namespace cpp {
namespace wg {

constexpr int kMaxNets = 3;
constexpr int kMaxSlabs = 96;      // K8 slabs of (tap, channel group) rows of the gradient matrix
constexpr int kMaxWarps = 8;
constexpr int kBandRows = 4;       // conv-output rows staged per work unit (2 for very wide images)

struct Slab {
  int8_t kind;    // 0: 8 channels of one tap (ky, kx, group); 1: remainder channels packed along kx (ky, slab j); 2: unused
  int8_t set;     // channel group g / packed slab j
  int8_t ky, kx;
};

struct Plan {
  // ---- problem
  const __half* x;                 // fp16 NHWC input [B][H][W][C]
  const float* mean_inv;           // [mean(C) | inv(C)] or NULL: whitening of x (base_network.py:95-99) folded into the result
  const float* g[kMaxNets];        // d(pooled) fp32 [B][PH][PW][10]
  const uint8_t* amax[kMaxNets];   // argmax side band of the forward pass
  float* dw[kMaxNets];             // HWIO [KS][KS][Cout_in][10]
  float* db[kMaxNets];             // [10]
  const float* gmax[kMaxNets];     // max |g| per network (device; filled by the absmax kernel here or handed in by the caller)
  float* gmax_own;                 // [nets] scratch the absmax kernel writes when the caller has no precomputed maxima
  float* partials;                 // [grid][part_floats] per-CTA partial sums, thread-native order
  float* gsum;                     // [part_floats] reduced over CTAs
  int B, H, W, C, KS, PAD, PH, PW, nets;
  int dup;                         // 1: x holds [hi(C/2) | lo(C/2)] fp16 pieces of an fp32 activation: dw[c] = G[c] + G[c + C/2]
                                   // 2: x is in the 24-channel piece layout of conv_tc.cuh (aligned vectors, constant-one channel
                                   //    included): rows go global -> shared planes with cp.async, no re-layout at all
  // ---- geometry
  int CE;                          // channels incl. the constant-one channel appended at index C
  int G8, R, nR, nvec;             // full 8-channel groups, remainder channels, packed slabs per ky, smem vectors per pixel
  int Wp, pitch, rows_in, band_rows;   // W rounded up to 16; plane row pitch (pixels); input / output rows per band
  int n_slabs, m_tiles, MT, NW;    // M tiles of 16 rows (2 slabs); M tiles per warp; warps
  int NT, NTp;                     // N tiles of 8 columns (nets * 2 pieces * 10 filters); vectors per pixel of the dY staging (odd)
  int bands_per_image, total_bands, grid, flush_every;
  int part_floats;
  int raw_bytes, plane_bytes, smem_bytes;
  Slab slab[kMaxSlabs];
  int32_t slab_off[kMaxSlabs];     // byte offset of the slab's row (band row 0, output column 0) inside the x planes
};

bool conv_wgrad_mma_supported(int nets, int H, int W, int C, int KS, int dup = 0);
int64_t conv_wgrad_mma_scratch_bytes(int nets, int H, int W, int C, int KS, int dup = 0);
// dw_n, db_n of the layer for every sibling network; x fp16 (exact pixels, or hi|lo pieces when dup); scratch 256-byte aligned
int launch_conv_wgrad_mma(const void* x_f16, const float* mean_inv, int dup, int nets, const float* const* d_pooled,
                          const uint8_t* const* amax, int B, int H, int W, int C, int KS, float* const* dw, float* const* db,
                          void* scratch, cudaStream_t s, const float* const* gmax_pre = nullptr);

}  // namespace wg
}  // namespace cpp
