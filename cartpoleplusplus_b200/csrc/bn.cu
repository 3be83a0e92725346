// --use-batch-norm (base_network.py:74-79): slim.conv2d(normalizer_fn=slim.batch_norm) = raw conv (no bias) -> batch_norm
// -> ReLU, then slim.max_pool2d.  slim.batch_norm defaults: center (beta) but no scale, epsilon 1e-3;
// IS_TRAINING = True: per-channel batch mean and POPULATION variance over (B, H, W), gradients flow through both;
// IS_TRAINING = False: the moving statistics, which the reference never updates (UPDATE_OPS are never run), so 0 / 1 unless a
// target copy wrote something else.  The raw conv output comes from conv.cu (launch_conv_raw); here:
//   forward : channel moments of the raw output (fp64 partials, fixed order) -> y = (x - mean) * inv + beta -> ReLU -> 2x2 max-pool
//             with the same arg-max side band as the fused conv kernels (0..3 winner, 4 = ReLU closed)
//   backward: dY (un-pooled through the side band) -> S1 = sum dY, S2 = sum dY * xhat per channel (fp64 partials) ->
//             d(conv) = inv * (dY - S1 / n - xhat * S2 / n) DENSE over every position of the layer, d(beta) = S1
// Every cross-CTA reduction is partials + a fixed-order second pass (no floating-point atomics), like the rest of the library.
#include "net.cuh"

namespace cpp {

constexpr int CO = kConvCout;
constexpr int kBnThreads = 32 * CO;          // thread t: channel t % 10, position slot t / 10
constexpr int kBnMaxParts = 2 * kNumSMs;
constexpr double kBnEps = 1e-3;

// scratch of one layer: double partials[kBnMaxParts][20] | float stats[32] = mean[10], inv[10] | float sums[32] = S1/n[10], S2/n[10]
int64_t bn_scratch_bytes() { return (int64_t)kBnMaxParts * 2 * CO * sizeof(double) + 64 * sizeof(float); }
static inline double* bn_partials(void* scratch) { return reinterpret_cast<double*>(scratch); }
static inline float* bn_stats(void* scratch) { return reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + (size_t)kBnMaxParts * 2 * CO * sizeof(double)); }
static inline float* bn_sums(void* scratch) { return bn_stats(scratch) + 32; }

// block-level fixed-order reduction of (a, b) per channel over the 32 position slots; thread o < 10 returns the sums
__device__ __forceinline__ void bn_block_reduce(double a, double b, double* out2 /* [20] of this block */) {
  __shared__ double sh[2][kBnThreads];
  const int t = threadIdx.x;
  sh[0][t] = a; sh[1][t] = b;
  __syncthreads();
  if (t < CO) {
    double sa = 0.0, sb = 0.0;
    for (int k = 0; k < 32; ++k) { sa += sh[0][k * CO + t]; sb += sh[1][k * CO + t]; }
    out2[t] = sa; out2[CO + t] = sb;
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_moments_partial_kernel(const float* __restrict__ raw, int64_t n_pix, double* __restrict__ partials) {
  const int o = threadIdx.x % CO, slot = threadIdx.x / CO;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * 32 + slot; p < n_pix; p += (int64_t)gridDim.x * 32) {
    const double v = (double)raw[p * CO + o];
    s1 += v; s2 += v * v;
  }
  bn_block_reduce(s1, s2, partials + (size_t)blockIdx.x * 2 * CO);
}

// one block of 32 threads: mean / inv from the partials (training) or from the moving statistics (inference)
__global__ void bn_stats_kernel(const double* __restrict__ partials, int nparts, int64_t n_pix, const float* __restrict__ moving_mean,
                                const float* __restrict__ moving_var, int training, float* __restrict__ stats) {
  const int o = threadIdx.x;
  if (o >= CO) return;
  double mean, var;
  if (training) {
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < nparts; ++k) { s1 += partials[(size_t)k * 2 * CO + o]; s2 += partials[(size_t)k * 2 * CO + CO + o]; }
    mean = s1 / (double)n_pix;
    var = s2 / (double)n_pix - mean * mean;
    if (var < 0.0) var = 0.0;
  } else {
    mean = (double)moving_mean[o]; var = (double)moving_var[o];
  }
  stats[o] = (float)mean;
  stats[CO + o] = (float)(1.0 / sqrt(var + kBnEps));
}

__global__ void __launch_bounds__(256) bn_relu_pool_kernel(const float* __restrict__ raw, const float* __restrict__ stats,
                                                           const float* __restrict__ beta, int B, int H, int W, int PH, int PW,
                                                           float* __restrict__ pooled, uint8_t* __restrict__ amax) {
  const int64_t total = (int64_t)B * PH * PW * CO;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % CO);
    const int64_t q = i / CO;
    const int px = (int)(q % PW), py = (int)((q / PW) % PH), b = (int)(q / ((int64_t)PW * PH));
    const float mean = stats[o], inv = stats[CO + o], be = beta[o];
    const float* r0 = raw + (((size_t)b * H + 2 * py) * W + 2 * px) * CO + o;
    float best = fmaf(r0[0] - mean, inv, be);
    int arg = 0;
    const float v01 = fmaf(r0[CO] - mean, inv, be), v10 = fmaf(r0[(size_t)W * CO] - mean, inv, be), v11 = fmaf(r0[(size_t)W * CO + CO] - mean, inv, be);
    if (v01 > best) { best = v01; arg = 1; }
    if (v10 > best) { best = v10; arg = 2; }
    if (v11 > best) { best = v11; arg = 3; }
    pooled[i] = fmaxf(best, 0.f);
    amax[i] = best > 0.f ? (uint8_t)arg : (uint8_t)4;
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_bwd_partial_kernel(const float* __restrict__ d_pooled, const uint8_t* __restrict__ amax,
                                                                    const float* __restrict__ raw, const float* __restrict__ stats,
                                                                    int B, int H, int W, int PH, int PW, double* __restrict__ partials) {
  const int o = threadIdx.x % CO, slot = threadIdx.x / CO;
  const float mean = stats[o], inv = stats[CO + o];
  const int64_t nq = (int64_t)B * PH * PW;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * 32 + slot; q < nq; q += (int64_t)gridDim.x * 32) {
    const int a = amax[q * CO + o];
    if (a < 4) {
      const int px = (int)(q % PW), py = (int)((q / PW) % PH), b = (int)(q / ((int64_t)PW * PH));
      const float g = d_pooled[q * CO + o];
      const float xh = (raw[(((size_t)b * H + 2 * py + (a >> 1)) * W + 2 * px + (a & 1)) * CO + o] - mean) * inv;
      s1 += (double)g; s2 += (double)g * (double)xh;
    }
  }
  bn_block_reduce(s1, s2, partials + (size_t)blockIdx.x * 2 * CO);
}

__global__ void bn_bwd_sums_kernel(const double* __restrict__ partials, int nparts, int64_t n_pix, float* __restrict__ sums,
                                   float* __restrict__ dbeta) {
  const int o = threadIdx.x;
  if (o >= CO) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nparts; ++k) { s1 += partials[(size_t)k * 2 * CO + o]; s2 += partials[(size_t)k * 2 * CO + CO + o]; }
  sums[o] = (float)(s1 / (double)n_pix);
  sums[CO + o] = (float)(s2 / (double)n_pix);
  if (dbeta != nullptr) dbeta[o] = (float)s1;
}

__global__ void __launch_bounds__(256) bn_bwd_dconv_kernel(const float* __restrict__ d_pooled, const uint8_t* __restrict__ amax,
                                                           const float* __restrict__ raw, const float* __restrict__ stats,
                                                           const float* __restrict__ sums, int B, int H, int W, int PH, int PW,
                                                           float* __restrict__ dconv) {
  const int64_t total = (int64_t)B * H * W * CO;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % CO);
    const int64_t p = i / CO;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((int64_t)W * H));
    const int py = y >> 1, px = x >> 1;
    float dy = 0.f;
    if (py < PH && px < PW) {
      const size_t q = (((size_t)b * PH + py) * PW + px) * CO + o;
      if (amax[q] == (((y & 1) << 1) | (x & 1))) dy = d_pooled[q];
    }
    const float inv = stats[CO + o];
    const float xh = (raw[i] - stats[o]) * inv;
    dconv[i] = inv * (dy - sums[o] - xh * sums[CO + o]);
  }
}

static int bn_grid(int64_t n_items32) { return (int)std::max<int64_t>(1, std::min<int64_t>(kBnMaxParts, ceil_div(n_items32, 32 * 8))); }

int launch_bn_forward(const float* raw, const float* beta, const float* moving_mean, const float* moving_var, int training, int B,
                      int H, int W, void* scratch, float* pooled, uint8_t* amax, cudaStream_t s) {
  if (B <= 0) return CPP_OK;
  const int64_t n_pix = (int64_t)B * H * W;
  int nparts = 0;
  if (training) {
    nparts = bn_grid(n_pix);
    bn_moments_partial_kernel<<<nparts, kBnThreads, 0, s>>>(raw, n_pix, bn_partials(scratch));
    CPP_CHECK_LAUNCH();
  }
  bn_stats_kernel<<<1, 32, 0, s>>>(bn_partials(scratch), nparts, n_pix, moving_mean, moving_var, training, bn_stats(scratch));
  CPP_CHECK_LAUNCH();
  const int PH = H / 2, PW = W / 2;
  const int64_t total = (int64_t)B * PH * PW * CO;
  bn_relu_pool_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), 8 * kNumSMs)), 256, 0, s>>>(
      raw, bn_stats(scratch), beta, B, H, W, PH, PW, pooled, amax);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// needs the statistics launch_bn_forward(training = 1) left in `scratch`
int launch_bn_backward(const float* d_pooled, const uint8_t* amax, const float* raw, int B, int H, int W, void* scratch,
                       float* dconv, float* dbeta, cudaStream_t s) {
  if (B <= 0) return CPP_OK;
  const int PH = H / 2, PW = W / 2;
  const int64_t n_pix = (int64_t)B * H * W, nq = (int64_t)B * PH * PW;
  const int nparts = bn_grid(nq);
  bn_bwd_partial_kernel<<<nparts, kBnThreads, 0, s>>>(d_pooled, amax, raw, bn_stats(scratch), B, H, W, PH, PW, bn_partials(scratch));
  CPP_CHECK_LAUNCH();
  bn_bwd_sums_kernel<<<1, 32, 0, s>>>(bn_partials(scratch), nparts, n_pix, bn_sums(scratch), dbeta);
  CPP_CHECK_LAUNCH();
  const int64_t total = n_pix * CO;
  bn_bwd_dconv_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), 16 * kNumSMs)), 256, 0, s>>>(
      d_pooled, amax, raw, bn_stats(scratch), bn_sums(scratch), B, H, W, PH, PW, dconv);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
