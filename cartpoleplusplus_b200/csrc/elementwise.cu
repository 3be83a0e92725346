// HBM-bound / latency-bound kernels of the hot path: replay gather, whitening moments, TD target +
// squared error, NAF L.L^T head, LRPG loss, global-norm clip, optimiser apply, tau-soft target copy.
// 128-bit coalesced accesses, warp-shuffle reductions, fixed-order (deterministic) cross-block sums.
#include "common.cuh"

namespace cpp {

// ---------------------------------------------------------------------------- small utilities
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in thread 0 (fixed order => deterministic)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh /* [32] */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  T r = 0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) r += sh[i];
  }
  return r;
}

// ---------------------------------------------------------------------------- a2: replay gather
// one CTA per (row, which): copies a whole fp16 state row with 128-bit loads/stores
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ slab, const int32_t* __restrict__ s1_idx,
                                                     const int32_t* __restrict__ s2_idx, const float* __restrict__ action,
                                                     const float* __restrict__ reward, const float* __restrict__ mask,
                                                     const int64_t* __restrict__ idxs, int64_t row_vec, int64_t row_elems,
                                                     int A, uint4* __restrict__ o1, uint4* __restrict__ o2,
                                                     float* __restrict__ oa, float* __restrict__ orw, float* __restrict__ om) {
  const int b = blockIdx.x, which = blockIdx.y;
  const int64_t idx = idxs[b];
  const int64_t slot = which == 0 ? s1_idx[idx] : s2_idx[idx];
  if (row_vec * 8 == row_elems) {
    const uint4* src = slab + slot * row_vec;
    uint4* dst = (which == 0 ? o1 : o2) + (int64_t)b * row_vec;
    for (int64_t i = threadIdx.x; i < row_vec; i += blockDim.x) dst[i] = __ldg(src + i);
  } else {   // rows that are not a multiple of 16 bytes (e.g. low-dim 28 x fp16): element copy
    const __half* src = reinterpret_cast<const __half*>(slab) + slot * row_elems;
    __half* dst = reinterpret_cast<__half*>(which == 0 ? o1 : o2) + (int64_t)b * row_elems;
    for (int64_t i = threadIdx.x; i < row_elems; i += blockDim.x) dst[i] = src[i];
  }
  if (which == 0 && threadIdx.x < 32) {
    if (threadIdx.x < A) oa[b * A + threadIdx.x] = action[idx * A + threadIdx.x];
    if (threadIdx.x == 0) { orw[b] = reward[idx]; om[b] = mask[idx]; }
  }
}

// ---------------------------------------------------------------------------- whitening moments
// Deterministic per-channel reduction of per-thread lane sums.  Thread t (global index t0 + tid) holds 8
// lane sums whose channels are (8*(t0+tid) + k) % C; threads tid and tid + C share the pattern, so
// (A) thread r < min(C, T) folds tid = r, r+C, ... in order, (B) thread c gathers its 8*min(C,T)/C folds.
// buf: double[T][16] shared.  out[c] / out[C+c] receive the block totals (fixed order).
__device__ void channel_reduce(const double (&s1)[8], const double (&s2)[8], int64_t t0, int C, double* buf, double* out) {
  const int T = blockDim.x, tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) { buf[tid * 16 + k] = s1[k]; buf[tid * 16 + 8 + k] = s2[k]; }
  __syncthreads();
  const int R = C < T ? C : T;
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 0.0;
  if (tid < R)
    for (int t = tid; t < T; t += C)
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] += buf[t * 16 + k];
  __syncthreads();
  if (tid < R)
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[tid * 16 + k] = a[k];
  __syncthreads();
  for (int c = tid; c < C; c += T) {
    double x1 = 0.0, x2 = 0.0;
    for (int r = 0; r < R; ++r) {
      const int base = (int)(((t0 + r) * 8) % C);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int ch = base + k; if (ch >= C) ch %= C;
        if (ch == c) { x1 += buf[r * 16 + k]; x2 += buf[r * 16 + 8 + k]; }
      }
    }
    out[c] = x1; out[C + c] = x2;
  }
}

// vectors of 8 halfs (or 8 floats); the grid-stride (in vectors) is a multiple of C so a thread's 8
// lanes always see the same 8 channels -> per-thread fp64 accumulators, no atomics, fixed summation order.
template <bool F16>
__global__ void __launch_bounds__(256) moments_partial_kernel(const void* __restrict__ x, int64_t n_elems, int C,
                                                              int64_t stride_vec, double* __restrict__ partial /*[grid][2C]*/) {
  extern __shared__ double sh[];   // [256*16]
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
  const int64_t n_vec = n_elems / 8;
  for (int64_t v = t; v < n_vec && t < stride_vec; v += stride_vec) {
    float f[8];
    if (F16) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x) + v);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 p = __half22float2(h[k]); f[2 * k] = p.x; f[2 * k + 1] = p.y; }
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * v), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * v + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { const double d = (double)f[k]; s1[k] += d; s2[k] += d * d; }
  }
  // tail elements (n_elems % 8 < 8) are folded in by the thread that owns the matching lane pattern:
  // thread 0 of block 0 adds them to a private copy after the reduction (fixed order)
  double* outp = partial + (size_t)blockIdx.x * 2 * C;
  channel_reduce(s1, s2, (int64_t)blockIdx.x * blockDim.x, C, sh, outp);
  if (blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int64_t e = n_vec * 8; e < n_elems; ++e) {
        const double d = F16 ? (double)__half2float(reinterpret_cast<const __half*>(x)[e]) : (double)reinterpret_cast<const float*>(x)[e];
        outp[(int)(e % C)] += d; outp[C + (int)(e % C)] += d * d;
      }
    }
  }
}

// one warp per channel: lanes split the per-block partials (fp64), fixed-order shuffle reduction
__global__ void __launch_bounds__(256) moments_finalize_kernel(const double* __restrict__ partial, int nparts, int C, double n_per_channel,
                                                               float* __restrict__ mean_inv) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = lane; k < nparts; k += 32) { s1 += partial[(size_t)k * 2 * C + c]; s2 += partial[(size_t)k * 2 * C + C + c]; }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane != 0) return;
  const double mean = s1 / n_per_channel;
  double var = s2 / n_per_channel - mean * mean;      // population variance (tf.nn.moments)
  if (var < 0.0) var = 0.0;
  mean_inv[c] = (float)mean;
  mean_inv[C + c] = (float)(1.0 / sqrt(var + 1e-6));   // rsqrt(var + 1e-6), base_network.py:97-99
}

constexpr int kMomentBlocks = 296;

int64_t moments_scratch_doubles(int C) { return (int64_t)kMomentBlocks * 2 * C; }

int launch_channel_moments(const void* x, int is_f16, int64_t n_pix_total, int C, double* scratch, float* mean_inv, cudaStream_t s) {
  CPP_REQUIRE(C >= 1 && C <= 256, "moments: C=%d out of range", C);
  const int64_t n = n_pix_total * C;
  int blocks = (int)ceil_div(ceil_div(n, 8), 256 * 4);
  if (blocks > kMomentBlocks) blocks = kMomentBlocks;
  if (blocks < (int)ceil_div(C, 256)) blocks = (int)ceil_div(C, 256);
  if (blocks < 1) blocks = 1;
  // vector stride = a multiple of C (<= thread count) so every thread keeps fixed channels
  const int64_t st = ((int64_t)blocks * 256 / C) * C;
  if (is_f16) moments_partial_kernel<true><<<blocks, 256, 256 * 16 * sizeof(double), s>>>(x, n, C, st, scratch);
  else moments_partial_kernel<false><<<blocks, 256, 256 * 16 * sizeof(double), s>>>(x, n, C, st, scratch);
  CPP_CHECK_LAUNCH();
  moments_finalize_kernel<<<(unsigned)ceil_div(C, 8), 256, 0, s>>>(scratch, blocks, C, (double)n_pix_total, mean_inv);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// per-slot sums (computed once when a state is stored), one CTA per slot; same deterministic reduction
__global__ void __launch_bounds__(256) slot_stats_kernel(const __half* __restrict__ slab, const int32_t* __restrict__ slots,
                                                         int64_t n_pix, int C, int64_t stride_vec, double* __restrict__ stats) {
  extern __shared__ double sh[];
  const int slot = slots[blockIdx.x];
  const __half* row = slab + (size_t)slot * n_pix * C;
  const int64_t n = n_pix * C, n_vec = n / 8;
  const int64_t t = threadIdx.x;
  double s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
  const bool vec_ok = ((n * 2) % 16 == 0);    // rows stay 16-byte aligned inside the slab
  if (vec_ok) {
    for (int64_t v = t; v < n_vec && t < stride_vec; v += stride_vec) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(row) + v);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 p = __half22float2(h[k]);
        s1[2 * k] += (double)p.x; s2[2 * k] += (double)p.x * (double)p.x;
        s1[2 * k + 1] += (double)p.y; s2[2 * k + 1] += (double)p.y * (double)p.y;
      }
    }
  } else {
    for (int64_t v = t; v < n_vec && t < stride_vec; v += stride_vec)
#pragma unroll
      for (int k = 0; k < 8; ++k) { const double d = (double)__half2float(row[v * 8 + k]); s1[k] += d; s2[k] += d * d; }
  }
  double* outp = stats + (size_t)slot * 2 * C;
  channel_reduce(s1, s2, 0, C, sh, outp);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int64_t e = n_vec * 8; e < n; ++e) {
      const double d = (double)__half2float(row[e]);
      outp[(int)(e % C)] += d; outp[C + (int)(e % C)] += d * d;
    }
  }
}

// one warp per channel: lane l sums slots l, l+32, ... in fp64, then a fixed-order shuffle reduction (deterministic)
__global__ void __launch_bounds__(256) moments_from_slots_kernel(const double* __restrict__ stats, const int32_t* __restrict__ slot_table,
                                                                 const int64_t* __restrict__ idxs, int B, double n_per_channel, int C,
                                                                 float* __restrict__ mean_inv) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < B; b += 32) {
    const size_t slot = (size_t)slot_table[idxs[b]];
    s1 += stats[slot * 2 * C + c]; s2 += stats[slot * 2 * C + C + c];
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane != 0) return;
  const double mean = s1 / n_per_channel;
  double var = s2 / n_per_channel - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_inv[c] = (float)mean;
  mean_inv[C + c] = (float)(1.0 / sqrt(var + 1e-6));
}

// ---------------------------------------------------------------------------- layout helpers
__global__ void state_to_f32_kernel(const void* __restrict__ src, int is_f16, int B, int dim, float* __restrict__ dst, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * dim) return;
  const int b = (int)(i / dim), d = (int)(i - (int64_t)b * dim);
  dst[(size_t)b * ld + d] = is_f16 ? __half2float(reinterpret_cast<const __half*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
}
int launch_state_to_f32(const void* state, int is_f16, int B, int dim, float* dst, int ld, cudaStream_t s) {
  const int64_t n = (int64_t)B * dim;
  if (n == 0) return CPP_OK;
  state_to_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(state, is_f16, B, dim, dst, ld);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// fp32 -> fp16 copy of an environment state whose values are fp16 numbers stored as fp32 (bullet_cartpole.py:239-242 renders
// fp16(k) / 255 in fp16 into a float32 array); *inexact is set to 1 when an element does not survive the round trip
__global__ void f32_to_f16_exact_kernel(const float* __restrict__ src, int64_t n, __half* __restrict__ dst, float* __restrict__ inexact) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const __half h = __float2half_rn(v);
  dst[i] = h;
  if (!(__half2float(h) == v)) *inexact = 1.f;     // (NaN lands here too); racing writers all store the same value
}
int launch_f32_to_f16_exact(const float* src, int64_t n, __half* dst, float* inexact, cudaStream_t s) {
  if (n == 0) return CPP_OK;
  CPP_CHECK_CUDA(cudaMemsetAsync(inexact, 0, sizeof(float), s));
  f32_to_f16_exact_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(src, n, dst, inexact);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int src_ld, int B, int cols, float* __restrict__ dst, int dst_ld, int c0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * cols) return;
  const int b = i / cols, c = i - b * cols;
  dst[(size_t)b * dst_ld + c0 + c] = src[(size_t)b * src_ld + c];
}
int launch_copy_cols(const float* src, int src_ld, int B, int cols, float* dst, int dst_ld, int dst_col0, cudaStream_t s) {
  if (B * cols == 0) return CPP_OK;
  copy_cols_kernel<<<(unsigned)ceil_div(B * cols, 256), 256, 0, s>>>(src, src_ld, B, cols, dst, dst_ld, dst_col0);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// d_pre = d_out * act'(out)
__global__ void act_grad_kernel(const float* __restrict__ d_out, int d_ld, const float* __restrict__ out, int out_ld, int act,
                                int B, int n, float* __restrict__ d_pre, int p_ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int b = i / n, c = i - b * n;
  float g = d_out[(size_t)b * d_ld + c];
  const float y = out[(size_t)b * out_ld + c];
  if (act == 1) g = (y > 0.f) ? g : 0.f;
  else if (act == 2) g = g * (1.f - y * y);
  d_pre[(size_t)b * p_ld + c] = g;
}
int launch_act_grad(const float* d_out, int d_ld, const float* out, int out_ld, int act, int B, int n, float* d_pre, int p_ld, cudaStream_t s) {
  if (B * n == 0) return CPP_OK;
  act_grad_kernel<<<(unsigned)ceil_div(B * n, 256), 256, 0, s>>>(d_out, d_ld, out, out_ld, act, B, n, d_pre, p_ld);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// NAF --share-input-state-representation: input gradients of the single-layer mu and l heads, summed
//   out[b][j] = sum_k dmu_pre[b][k] * Wmu[j][k] + sum_k dl_pre[b][k] * Wl[j][k]          (W stored [in][out])
__global__ void heads_dgrad_kernel(const float* __restrict__ dmu, const float* __restrict__ Wmu, int A, const float* __restrict__ dl,
                                   const float* __restrict__ Wl, int NL, int B, int D, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, j = i - b * D;
  float s = 0.f;
  for (int k = 0; k < A; ++k) s = fmaf(dmu[b * A + k], Wmu[j * A + k], s);
  float t = 0.f;
  for (int k = 0; k < NL; ++k) t = fmaf(dl[b * NL + k], Wl[j * NL + k], t);
  out[i] = s + t;
}
int launch_heads_dgrad(const float* dmu_pre, const float* Wmu, int A, const float* dl_pre, const float* Wl, int NL, int B, int D,
                       float* out, cudaStream_t s) {
  heads_dgrad_kernel<<<(unsigned)ceil_div(B * D, 256), 256, 0, s>>>(dmu_pre, Wmu, A, dl_pre, Wl, NL, B, D, out);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// d[b][j] += extra[b][j], gated by the ReLU below (x = that layer's output; NULL: no gate)
__global__ void add_gated_kernel(float* __restrict__ d, int d_ld, const float* __restrict__ extra, const float* __restrict__ x, int x_ld,
                                 int B, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int b = i / n, j = i - b * n;
  const bool open = x == nullptr || x[(size_t)b * x_ld + j] > 0.f;
  if (open) d[(size_t)b * d_ld + j] += extra[i] * scale;
}
int launch_add_gated(float* d, int d_ld, const float* extra, const float* x, int x_ld, int B, int n, cudaStream_t s, float scale) {
  add_gated_kernel<<<(unsigned)ceil_div(B * n, 256), 256, 0, s>>>(d, d_ld, extra, x, x_ld, B, n, scale);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// --use-dropout (base_network.py:69-70): slim.dropout(layer, keep_prob 0.5, is_training=IS_TRAINING) = inverted dropout
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void dropout_kernel(float* __restrict__ h, int ld, int B, int n, uint8_t* __restrict__ mask, const unsigned long long* __restrict__ counter,
                               unsigned long long key, int external) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int b = i / n, j = i - b * n;
  uint8_t m;
  if (external) m = mask[i];
  else {
    m = (uint8_t)(splitmix64(splitmix64(key ^ *counter) + (unsigned long long)i) >> 63);     // Bernoulli(0.5)
    mask[i] = m;
  }
  float* p = h + (size_t)b * ld + j;
  *p = m ? *p * 2.f : 0.f;
}
__global__ void dropout_tick_kernel(unsigned long long* counter) { *counter += 1; }
int launch_dropout(float* h, int ld, int B, int n, uint8_t* mask, const unsigned long long* counter, int layer, cudaStream_t s) {
  // distinct streams per seed, layer and network (the counter lives in that network's workspace: its address tells them apart)
  const unsigned long long key = ((unsigned long long)(unsigned)g_dropout_seed << 32) ^ ((unsigned long long)(unsigned)layer * 0x9E3779B1ull) ^
                                 ((unsigned long long)reinterpret_cast<uintptr_t>(counter) * 0xD6E8FEB86659FD93ull);
  dropout_kernel<<<(unsigned)ceil_div(B * n, 256), 256, 0, s>>>(h, ld, B, n, mask, counter, key, g_dropout_external);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}
int launch_dropout_tick(unsigned long long* counter, cudaStream_t s) {
  dropout_tick_kernel<<<1, 1, 0, s>>>(counter);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// bias gradient: out[c] = sum_b x[b][c]; one warp per column, fixed order
__global__ void colsum_kernel(const float* __restrict__ x, int ld, int B, int n, float* __restrict__ out) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= n) return;
  float s = 0.f;
  for (int b = lane; b < B; b += 32) s += x[(size_t)b * ld + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
}
int launch_colsum(const float* x, int ld, int B, int n, float* out, cudaStream_t s) {
  if (n == 0) return CPP_OK;
  colsum_kernel<<<(unsigned)ceil_div(n, 4), 128, 0, s>>>(x, ld, B, n, out);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

__global__ void scale_copy_kernel(const float* __restrict__ src, float scale, int64_t n, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * scale;
}
int launch_scale_copy(const float* src, float scale, int64_t n, float* dst, cudaStream_t s) {
  if (n == 0) return CPP_OK;
  scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(src, scale, n, dst);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}
__global__ void fill_kernel(float* __restrict__ dst, float v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
int launch_fill(float* dst, float v, int64_t n, cudaStream_t s) {
  if (n == 0) return CPP_OK;
  fill_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(dst, v, n);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// ---------------------------------------------------------------------------- a8/a9: TD target + MSE
// y = r + mask*gamma*q2 ; td = q - y ; loss = sum(td^2)/B_global ; dq = 2 td / B_global
// single CTA (B <= a few thousand): fixed-order reduction.  loss_flag[0] = loss, [1] = non-finite count.
__global__ void __launch_bounds__(1024) td_mse_kernel(const float* __restrict__ q, const float* __restrict__ q2,
                                                      const float* __restrict__ reward, const float* __restrict__ mask,
                                                      float gamma, int B, float inv_bglobal,
                                                      float* __restrict__ td_out, float* __restrict__ dq, float* __restrict__ loss_flag) {
  __shared__ float sh[32];
  float s = 0.f, bad = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float y = reward[b] + mask[b] * gamma * q2[b];     // ddpg_cartpole.py:199 / naf_cartpole.py:225
    const float td = q[b] - y;
    if (td_out) td_out[b] = td;
    if (dq) dq[b] = 2.f * td * inv_bglobal;
    s += td * td;
    if (!isfinite(td)) bad += 1.f;
  }
  const float tot = block_sum(s, sh);
  const float nb = block_sum(bad, sh);
  if (threadIdx.x == 0 && loss_flag) { loss_flag[0] = tot * inv_bglobal; loss_flag[1] = nb; }
}
int launch_td_mse(const float* q, const float* q2, const float* reward, const float* mask, float gamma, int B, int B_global,
                  float* td, float* dq, float* loss_flag, cudaStream_t s) {
  td_mse_kernel<<<1, 1024, 0, s>>>(q, q2, reward, mask, gamma, B, 1.f / (float)B_global, td, dq, loss_flag);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// ---------------------------------------------------------------------------- a10: NAF head
// naf_cartpole.py:163-230 for a general action_dim A (<= 8): L rows = [lower.., exp(diag), 0..],
// z = L^T d, A = -1/2 z.z, Q = V + A, delta = 2 (Q - y)/B_global; closed-form backward (SURVEY A-7).
constexpr int kMaxA = 8;
__global__ void __launch_bounds__(1024) naf_head_kernel(const float* __restrict__ V, const float* __restrict__ mu,
                                                        const float* __restrict__ lv, const float* __restrict__ u,
                                                        const float* __restrict__ reward, const float* __restrict__ mask,
                                                        const float* __restrict__ V2, float gamma, int B, int A, float inv_bglobal,
                                                        float* __restrict__ dV, float* __restrict__ dmu, float* __restrict__ dl,
                                                        float* __restrict__ adv_out, float* __restrict__ loss_flag) {
  __shared__ float sh[32];
  const int NL = A * (A + 1) / 2;
  float s = 0.f, bad = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float L[kMaxA][kMaxA], d[kMaxA], z[kMaxA];
    bool finite = true;
#pragma unroll
    for (int r = 0; r < kMaxA; ++r) {
      if (r >= A) break;
      const int off = r * (r + 1) / 2;
      for (int c = 0; c < r; ++c) { L[r][c] = lv[b * NL + off + c]; finite = finite && isfinite(L[r][c]); }
      L[r][r] = expf(lv[b * NL + off + r]);
      finite = finite && isfinite(L[r][r]) && isfinite(lv[b * NL + off + r]);
      d[r] = u[b * A + r] - mu[b * A + r];
    }
    float zz = 0.f;
    for (int c = 0; c < A; ++c) {
      float t = 0.f;
      for (int r = c; r < A; ++r) t = fmaf(L[r][c], d[r], t);
      z[c] = t; zz = fmaf(t, t, zz);
    }
    const float adv = -0.5f * zz;
    const float y = reward[b] + mask[b] * gamma * V2[b];
    const float td = (V[b] + adv) - y;
    const float delta = 2.f * td * inv_bglobal;
    if (adv_out) adv_out[b] = adv;
    if (dV) {
      dV[b] = delta;
      for (int r = 0; r < A; ++r) {       // d/dmu = delta * (L z)
        float t = 0.f;
        for (int c = 0; c <= r; ++c) t = fmaf(L[r][c], z[c], t);
        dmu[b * A + r] = delta * t;
        const int off = r * (r + 1) / 2;
        for (int c = 0; c < r; ++c) dl[b * NL + off + c] = -delta * d[r] * z[c];
        dl[b * NL + off + r] = -delta * d[r] * z[r] * L[r][r];
      }
    }
    s += td * td;
    if (!finite || !isfinite(td)) bad += 1.f;
  }
  const float tot = block_sum(s, sh);
  const float nb = block_sum(bad, sh);
  if (threadIdx.x == 0 && loss_flag) { loss_flag[0] = tot * inv_bglobal; loss_flag[1] = nb; }
}
int launch_naf_head(const float* V, const float* mu, const float* lv, const float* u, const float* reward, const float* mask,
                    const float* V2, float gamma, int B, int A, int B_global, float* dV, float* dmu, float* dl,
                    float* adv, float* loss_flag, cudaStream_t s) {
  CPP_REQUIRE(A >= 1 && A <= kMaxA, "naf: action_dim %d unsupported (max %d)", A, kMaxA);
  naf_head_kernel<<<1, 1024, 0, s>>>(V, mu, lv, u, reward, mask, V2, gamma, B, A, 1.f / (float)B_global, dV, dmu, dl, adv, loss_flag);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// ---------------------------------------------------------------------------- a14: LRPG loss
// loss = -sum_i log_softmax(logits_i)[a_i] * (adv_i - mean)/std ; dlogits_i = -advhat_i (onehot - softmax)
__global__ void __launch_bounds__(1024) lrpg_loss_kernel(const float* __restrict__ logits, const int32_t* __restrict__ actions,
                                                         const float* __restrict__ adv, int N, int K,
                                                         float* __restrict__ dlogits, float* __restrict__ loss_out) {
  __shared__ float sh[32];
  __shared__ float s_mean, s_std;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += adv[i];
  float tot = block_sum(s, sh);
  if (threadIdx.x == 0) s_mean = tot / (float)N;
  __syncthreads();
  const float mean = s_mean;
  s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const float d = adv[i] - mean; s += d * d; }
  tot = block_sum(s, sh);
  if (threadIdx.x == 0) s_std = sqrtf(tot / (float)N);       // util.py:39-42, population std, no epsilon
  __syncthreads();
  const float stdv = s_std;
  float l = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float* z = logits + (size_t)i * K;
    float m = z[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, z[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(z[k] - m);
    const float lse = m + logf(se);
    const float ah = (adv[i] - mean) / stdv;
    const int a = actions[i];
    l -= (z[a] - lse) * ah;
    for (int k = 0; k < K; ++k) {
      const float pk = expf(z[k] - lse);
      dlogits[(size_t)i * K + k] = -ah * ((k == a ? 1.f : 0.f) - pk);
    }
  }
  tot = block_sum(l, sh);
  if (threadIdx.x == 0) loss_out[0] = tot;
}
int launch_lrpg_loss(const float* logits, const int32_t* actions, const float* adv, int N, int K, float* dlogits, float* loss, cudaStream_t s) {
  lrpg_loss_kernel<<<1, 1024, 0, s>>>(logits, actions, adv, N, K, dlogits, loss);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// ---------------------------------------------------------------------------- a11: global norm clip
constexpr int kNormBlocks = 148;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ double sh[32];
  double s = 0.0;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0) for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += (double)g[i] * g[i];
  const double tot = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void norm_finalize_kernel(const double* __restrict__ partial, int nparts, float clip, float* __restrict__ out2) {
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < nparts; ++i) s += partial[i];
  const float norm = (float)sqrt(s);
  float scale = 1.f;
  if (clip > 0.f) scale = clip * fminf(1.f / norm, 1.f / clip);      // tf.clip_by_global_norm
  out2[0] = scale; out2[1] = norm;
}
int64_t norm_scratch_doubles() { return kNormBlocks; }
int launch_global_norm_scale(const float* grads, int64_t n, float clip, double* scratch, float* out2, cudaStream_t s) {
  CPP_REQUIRE(((uintptr_t)grads & 15) == 0, "grads must be 16-byte aligned");
  int blocks = (int)ceil_div(ceil_div(n, 4), 256);
  if (blocks > kNormBlocks) blocks = kNormBlocks;
  if (blocks < 1) blocks = 1;
  sumsq_partial_kernel<<<blocks, 256, 0, s>>>(grads, n, scratch);
  CPP_CHECK_LAUNCH();
  norm_finalize_kernel<<<1, 32, 0, s>>>(scratch, blocks, clip, out2);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// DDPG applies two independent (clip, SGD) pairs - actor and critic (ddpg_cartpole.py:116-119,213-218) - back to back:
// the same three kernels handle both segments in one launch each (blockIdx.y = segment), same arithmetic and order
struct DualSeg { const float* g[2]; float* p[2]; int64_t n[2]; float lr[2]; };
__global__ void __launch_bounds__(256) sumsq_partial_dual_kernel(DualSeg d, double* __restrict__ partial /*[2][kNormBlocks]*/) {
  __shared__ double sh[32];
  const int seg = blockIdx.y;
  const float* g = d.g[seg];
  const int64_t n = d.n[seg], n4 = n / 4;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0) for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += (double)g[i] * g[i];
  const double tot = block_sum(s, sh);
  if (threadIdx.x == 0) partial[seg * kNormBlocks + blockIdx.x] = tot;
}
__global__ void norm_finalize_dual_kernel(const double* __restrict__ partial, int nparts, float clip, float* __restrict__ out4) {
  const int seg = threadIdx.x;
  if (seg >= 2) return;
  double s = 0.0;
  for (int i = 0; i < nparts; ++i) s += partial[seg * kNormBlocks + i];
  const float norm = (float)sqrt(s);
  float scale = 1.f;
  if (clip > 0.f) scale = clip * fminf(1.f / norm, 1.f / clip);
  out4[2 * seg] = scale; out4[2 * seg + 1] = norm;
}
__global__ void __launch_bounds__(256) sgd_dual_kernel(DualSeg d, const float* __restrict__ scale4) {
  const int seg = blockIdx.y;
  const int64_t n = d.n[seg], n4 = n / 4;
  const float scale = scale4[2 * seg], lr = d.lr[seg];
  float* p = d.p[seg]; const float* g = d.g[seg];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    pv.x -= lr * (gv.x * scale); pv.y -= lr * (gv.y * scale); pv.z -= lr * (gv.z * scale); pv.w -= lr * (gv.w * scale);
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  if (i == 0) for (int64_t e = n4 * 4; e < n; ++e) p[e] -= lr * (g[e] * scale);
}
int launch_clip_sgd_dual(float* p0, const float* g0, int64_t n0, float lr0, float* p1, const float* g1, int64_t n1, float lr1,
                         float clip, double* scratch /*[2 * norm_scratch_doubles()]*/, float* out4, cudaStream_t s) {
  DualSeg d;
  d.g[0] = g0; d.g[1] = g1; d.p[0] = p0; d.p[1] = p1; d.n[0] = n0; d.n[1] = n1; d.lr[0] = lr0; d.lr[1] = lr1;
  const int64_t nmax = n0 > n1 ? n0 : n1;
  int blocks = (int)ceil_div(ceil_div(nmax, 4), 256);
  if (blocks > kNormBlocks) blocks = kNormBlocks;
  if (blocks < 1) blocks = 1;
  sumsq_partial_dual_kernel<<<dim3(blocks, 2), 256, 0, s>>>(d, scratch);
  CPP_CHECK_LAUNCH();
  norm_finalize_dual_kernel<<<1, 32, 0, s>>>(scratch, blocks, clip, out4);
  CPP_CHECK_LAUNCH();
  sgd_dual_kernel<<<dim3((unsigned)ceil_div(ceil_div(nmax, 4) + 1, 256), 2), 256, 0, s>>>(d, out4);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// ---------------------------------------------------------------------------- a12: optimisers
// kind 0: p -= lr*g ; 1: acc = m*acc + g, p -= lr*acc ; 2: Adam with epsilon-hat (SURVEY A-9)
template <int KIND>
__global__ void __launch_bounds__(256) optimiser_kernel(float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ scale_ptr,
                                                        int64_t n, float lr, float mom, float b1, float b2, float eps,
                                                        float* __restrict__ slots, const float* __restrict__ opt_state,
                                                        const float* __restrict__ skip) {
  if (skip && skip[0] != 0.f) return;     // non-finite step: leave parameters and slots untouched
  const float scale = scale_ptr ? scale_ptr[0] : 1.f;
  float lr_t = lr;
  if (KIND == 2) {
    const float b1p = opt_state[0] * b1, b2p = opt_state[1] * b2;    // powers AFTER this step
    lr_t = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  auto upd = [&](float& pv, float gv, float& s0, float& s1) {
    gv *= scale;
    if (KIND == 0) pv -= lr * gv;
    else if (KIND == 1) { s0 = mom * s0 + gv; pv -= lr * s0; }
    else { s0 += (1.f - b1) * (gv - s0); s1 += (1.f - b2) * (gv * gv - s1); pv -= lr_t * s0 / (sqrtf(s1) + eps); }
  };
  if (i < n4) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
    if (KIND >= 1) a = reinterpret_cast<float4*>(slots)[i];
    if (KIND == 2) b = reinterpret_cast<float4*>(slots + n)[i];     // requires n % 4 == 0 for alignment, checked on host
    upd(pv.x, gv.x, a.x, b.x); upd(pv.y, gv.y, a.y, b.y); upd(pv.z, gv.z, a.z, b.z); upd(pv.w, gv.w, a.w, b.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    if (KIND >= 1) reinterpret_cast<float4*>(slots)[i] = a;
    if (KIND == 2) reinterpret_cast<float4*>(slots + n)[i] = b;
  }
  if (i == 0) {
    for (int64_t e = n4 * 4; e < n; ++e) {
      float a = 0, b = 0;
      if (KIND >= 1) a = slots[e];
      if (KIND == 2) b = slots[n + e];
      float pv = p[e];
      upd(pv, g[e], a, b);
      p[e] = pv;
      if (KIND >= 1) slots[e] = a;
      if (KIND == 2) slots[n + e] = b;
    }
  }
}
__global__ void adam_advance_kernel(float* opt_state, float b1, float b2, const float* skip) {
  if (skip && skip[0] != 0.f) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) { opt_state[0] *= b1; opt_state[1] *= b2; }
}
int launch_optimiser(int kind, float* params, const float* grads, const float* scale, int64_t n, float lr, float momentum,
                     float beta1, float beta2, float eps, float* slots, float* opt_state, const float* skip, cudaStream_t s) {
  CPP_REQUIRE(kind >= 0 && kind <= 2, "optimiser kind %d unknown", kind);
  CPP_REQUIRE(((uintptr_t)params & 15) == 0 && ((uintptr_t)grads & 15) == 0, "params/grads must be 16-byte aligned");
  CPP_REQUIRE(kind == 0 || (slots != nullptr && ((uintptr_t)slots & 15) == 0), "optimiser slots missing/unaligned");
  CPP_REQUIRE(kind != 2 || (opt_state != nullptr && n % 4 == 0), "Adam needs opt_state and n %% 4 == 0 (pad the flat buffer)");
  if (n == 0) return CPP_OK;
  const unsigned blocks = (unsigned)ceil_div(ceil_div(n, 4), 256);
  if (kind == 0) optimiser_kernel<0><<<blocks, 256, 0, s>>>(params, grads, scale, n, lr, momentum, beta1, beta2, eps, slots, opt_state, skip);
  else if (kind == 1) optimiser_kernel<1><<<blocks, 256, 0, s>>>(params, grads, scale, n, lr, momentum, beta1, beta2, eps, slots, opt_state, skip);
  else optimiser_kernel<2><<<blocks, 256, 0, s>>>(params, grads, scale, n, lr, momentum, beta1, beta2, eps, slots, opt_state, skip);
  CPP_CHECK_LAUNCH();
  if (kind == 2) { adam_advance_kernel<<<1, 32, 0, s>>>(opt_state, beta1, beta2, skip); CPP_CHECK_LAUNCH(); }
  return CPP_OK;
}

// ---------------------------------------------------------------------------- a13: target copy
__global__ void __launch_bounds__(256) soft_update_kernel(float* __restrict__ t, const float* __restrict__ src, float c, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  if (i < n4) {
    float4 tv = reinterpret_cast<float4*>(t)[i];
    const float4 sv = __ldg(reinterpret_cast<const float4*>(src) + i);
    tv.x = __fsub_rn(tv.x, __fmul_rn(c, __fsub_rn(tv.x, sv.x)));       // t - c*(t - s), base_network.py:31
    tv.y = __fsub_rn(tv.y, __fmul_rn(c, __fsub_rn(tv.y, sv.y)));
    tv.z = __fsub_rn(tv.z, __fmul_rn(c, __fsub_rn(tv.z, sv.z)));
    tv.w = __fsub_rn(tv.w, __fmul_rn(c, __fsub_rn(tv.w, sv.w)));
    reinterpret_cast<float4*>(t)[i] = tv;
  }
  if (i == 0) for (int64_t e = n4 * 4; e < n; ++e) t[e] = __fsub_rn(t[e], __fmul_rn(c, __fsub_rn(t[e], src[e])));
}
int launch_soft_update(float* target, const float* source, float coeff, int64_t n, cudaStream_t s) {
  CPP_REQUIRE(coeff >= 0.f && coeff <= 1.f, "affine_combo_coeff %f outside [0,1]", coeff);   // base_network.py:22
  CPP_REQUIRE(((uintptr_t)target & 15) == 0 && ((uintptr_t)source & 15) == 0, "target/source must be 16-byte aligned");
  if (n == 0) return CPP_OK;
  soft_update_kernel<<<(unsigned)ceil_div(n / 4 + 1, 256), 256, 0, s>>>(target, source, coeff, n);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

int launch_gather(const void* slab, const int32_t* s1_idx, const int32_t* s2_idx, const float* action, const float* reward,
                  const float* mask, const int64_t* idxs, int B, int64_t row_elems, int A, void* o1, void* o2, float* oa,
                  float* orw, float* om, cudaStream_t s) {
  CPP_REQUIRE(A <= 32, "action_dim %d too large", A);
  if (B <= 0) return CPP_OK;
  const bool vec = (row_elems % 8 == 0) && (((uintptr_t)slab | (uintptr_t)o1 | (uintptr_t)o2) & 15) == 0;
  gather_kernel<<<dim3(B, 2), 256, 0, s>>>(reinterpret_cast<const uint4*>(slab), s1_idx, s2_idx, action, reward, mask, idxs,
                                          vec ? row_elems / 8 : 0, row_elems, A, reinterpret_cast<uint4*>(o1),
                                          reinterpret_cast<uint4*>(o2), oa, orw, om);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

int launch_slot_stats(const void* slab, const int32_t* slots, int n, int64_t n_pix, int C, double* stats, cudaStream_t s) {
  if (n <= 0) return CPP_OK;
  CPP_REQUIRE(C >= 1 && C <= 256, "slot stats: C=%d out of range", C);
  slot_stats_kernel<<<n, 256, 256 * 16 * sizeof(double), s>>>(reinterpret_cast<const __half*>(slab), slots, n_pix, C,
                                                             (int64_t)(256 / C) * C, stats);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}
int launch_moments_from_slots(const double* stats, const int32_t* slot_table, const int64_t* idxs, int B, int64_t n_pix, int C,
                              float* mean_inv, cudaStream_t s) {
  moments_from_slots_kernel<<<(unsigned)ceil_div(C, 8), 256, 0, s>>>(stats, slot_table, idxs, B, (double)B * (double)n_pix, C, mean_inv);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
