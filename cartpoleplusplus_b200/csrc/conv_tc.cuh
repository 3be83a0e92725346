// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for the reference conv trunk
// (Network.simple_conv_net_on, base_network.py:73-127).  See conv_tc.cu for the design.
#pragma once
#include "common.cuh"

namespace cpp {
namespace tc {

constexpr int kMaxNets = 3;     // sibling networks that read the same input (actor+critic; NAF value/mu/l)
constexpr int kMaxPairs = 64;   // K16 MMAs per accumulator
constexpr int kPieces = 2;      // fp16 pieces per fp32 weight (hi + lo): 22 mantissa bits
// 24-channel piece layout of a 10-channel fp32 activation: [hi0..7 | lo0..7 | hi8 hi9 lo8 lo9 1.0 0 0 0] - every 8-channel
// group is one aligned 16-byte vector, and the constant-one channel the weight-gradient kernel needs is already in place
constexpr int kC24 = 24;
__host__ __device__ inline int c24_weight_channel(int ch) { return ch < 8 ? ch : ch < 16 ? ch - 8 : ch < 18 ? ch - 8 : ch < 20 ? ch - 10 : -1; }
__host__ __device__ inline int c24_hi(int c) { return c < 8 ? c : 8 + c; }
__host__ __device__ inline int c24_lo(int c) { return c < 8 ? 8 + c : 10 + c; }
constexpr int kC24One = 20;

// one K8 slab of the implicit GEMM: 8 consecutive K entries that sit in ONE 16-byte shared-memory row
struct Slab {
  int8_t kind;   // 0: 8 channels of one tap (set = channel group); 1: remainder channels packed along kx; 2: zero
  int8_t set;    // channel group g (kind 0) or packed slab j (kind 1)
  int8_t ky, kx; // tap (kx unused for kind 1)
};

struct FwdPlan {
  // ---- problem
  const __half* x;          // fp16 NHWC images
  const int32_t* rows;      // optional: image b of the batch is x + rows[b] * H*W*C (fused replay gather)
  const __half* bpack;      // packed weights, canonical K-major UMMA layout (built by the prep kernel)
  const float* corr;        // [(2*PAD+1)^2][nets][10] bias - border-aware mean term, then [1] = 2^-S
  float* pooled[kMaxNets];
  uint8_t* amax[kMaxNets];
  __half* pooled_hl[kMaxNets];   // optional: the pooled output again as fp16 pieces in the 24-channel layout below, for the next layer
  int B, H, W, C, PH, PW, Pq, KS, PAD;
  int Cw;                   // weight input channels: C, or 10 when x holds fp16 pieces of an fp32 activation
  int in_layout;            // 0: plain channels; 1: [hi(10) | lo(10)]; 2: kPieceLayout24 (three aligned 16-byte vectors per pixel)
  int dgrad;                // 1: input-gradient mode - x = un-pooled output gradient pieces, taps flipped and channels transposed,
                            //    no bias / ReLU / pool: every conv position is written to the dense fp32 output pooled[n] [B][H][W][10]
  const float* out_scale;   // dgrad: device scalar the result is multiplied with (undoes the power-of-two scaling of the pieces)
  float* out_absmax;        // dgrad, optional: max |dx| is atomically max-ed into this device float (the next weight gradient's scale)
  int nets, N;              // MMA N = round_up(nets * kPieces * 10, 16)
  // ---- shared-memory geometry
  int G8, R, nR;            // full 8-channel groups, remainder channels, packed slabs per ky
  int n_planes, rows_alloc, plane_bytes, unit_bytes;
  int crh, stage_bytes, use_bulk;   // parity rows per staging group, staging buffer bytes, TMA bulk copy usable
  int stage_slots;                  // staging ring depth: raw rows are fetched this many groups ahead, across units
  int tiles_per_image, tiles_per_unit, units_per_image, n_units;
  int n_pairs;
  int smem_bytes;
  Slab slab[kMaxPairs][2];
  uint32_t adesc_lo[4][kMaxPairs];   // low word of every A descriptor (pool position a, instruction i), buffer 0, q = 0
};

struct PrepArgs {
  const float* w[kMaxNets];      // HWIO (KS,KS,C,10)
  const float* bias[kMaxNets];
  const float* mean_inv;         // [mean(C) | inv(C)] or NULL (no whitening fold)
  __half* bpack;
  float* corr;
};

// number of activations > 131 008 (= 2 x the fp16 maximum) that the hi + lo piece copy for the next layer could not represent
int piece_overflow_count(int reset, unsigned int* out);
int64_t conv_tc_scratch_bytes(int nets, int H, int W, int C, int KS);
bool conv_tc_supported(int nets, int H, int W, int C, int KS);
// y_n = maxpool2x2(relu(conv_same(whiten(x), w_n) + b_n)) for n < nets sibling networks in ONE pass over x.
// x_is_pieces: 1 = x holds [hi(10) | lo(10)] fp16 pieces of an fp32 activation, 2 = the 24-channel piece layout above
// (conv2/conv3); w has 10 input channels, mean_inv must be NULL.  pooled_hl (optional, may be NULL or hold NULLs): piece
// copy of the output in the 24-channel layout.
// phase: the launch is a weight-prep kernel (fp16 weight pieces + border table into `scratch`; needs only w, bias, mean_inv)
// followed by the main kernel.  kPhasePrep / kPhaseMain launch one of the two, so that a caller can prepare the weights
// of many layers on a side stream at the start of a step (the activations pointers may be NULL for kPhasePrep).
enum { kPhaseBoth = 0, kPhasePrep = 1, kPhaseMain = 2 };
int launch_conv_fwd_tc(const void* x_f16, const int32_t* rows, const float* mean_inv, int nets,
                       const float* const* w, const float* const* bias, int B, int H, int W, int C, int KS,
                       float* const* pooled, uint8_t* const* amax, void* scratch, cudaStream_t s,
                       int x_is_pieces = 0, __half* const* pooled_hl = nullptr, int phase = kPhaseBoth);
// gradient wrt the input of a 10 -> 10 channel layer (conv2 / conv3): dx = conv_same(dY, flip(w)^T) on the tensor cores.
// dy_pieces fp16 [B][H][W][24] (piece layout above, constant channel 0) = the un-pooled output gradient times *inv_scale^-1
// (launch_unpool_split);
// dx fp32 [B][H][W][10].
int launch_conv_dgrad_tc(const void* dy_pieces, const float* inv_scale, const float* w, int B, int H, int W, int KS, float* dx,
                         void* scratch, cudaStream_t s, float* out_absmax = nullptr, int phase = kPhaseBoth);
// the same input gradient with the un-pool / split pass fused into the row-sweep kernel's producer warps (conv_row_tc.cu): no piece
// tensor is written.  gmax: device float, receives max|d_pooled| here unless gmax_ready; inv_scale (optional) receives 1 / scale
bool conv_dgrad_fused_supported(int H, int W, int KS);
int launch_absmax(const float* g, int64_t n, float* gmax, cudaStream_t s);      // *gmax = max |g[i]| (zeroed here)
int launch_conv_dgrad_tc_fused(const float* d_pooled, const uint8_t* amax, float* gmax, float* inv_scale, int gmax_ready, const float* w,
                               int B, int H, int W, int KS, float* dx, void* scratch, cudaStream_t s, float* out_absmax = nullptr,
                               int phase = kPhaseBoth);
// un-pool + split: d_pooled fp32 [B][H/2][W/2][10] and the arg-max side band -> dy_pieces fp16 [B][H][W][24], scaled by
// a power of two from max|d_pooled| (gmax: device float, zeroed and filled here); inv_scale receives 1/scale
// gmax_ready: *gmax already holds max|d_pooled| (left there by the kernel that produced d_pooled)
int launch_unpool_split(const float* d_pooled, const uint8_t* amax, int B, int H, int W, float* gmax, float* inv_scale,
                        __half* dy_pieces, cudaStream_t s, int gmax_ready = 0);

}  // namespace tc
}  // namespace cpp
