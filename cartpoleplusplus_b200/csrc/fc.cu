// FP32 GEMM for the fully connected layers (slim.fully_connected, base_network.py:58-71;
// ddpg_cartpole.py:95-100,168-184; naf_cartpole.py:105-109,156-184): forward, dgrad and wgrad are the
// same kernel with different operand orientations.  The layers are tiny (K <= 2560, N <= 200) and
// launch/latency bound; exact fp32 FFMA keeps the 1e-5 parity budget for the conv path.
#include "common.cuh"

namespace cpp {

constexpr int BM = 32, BN = 32, BK = 32;     // small tiles: the layers are tiny, parallelism comes from the CTA count

// 128 threads, thread (tm, tn) owns a 2 x 4 block of C.  The next K tile is fetched into registers while the current one is
// multiplied out of shared memory (the layers are latency bound: K <= 2560, M or N <= 256).
template <bool TA, bool TB>
__global__ void __launch_bounds__(128) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tm = tid / 8, tn = tid % 8;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  constexpr int NA = (BM * BK) / 128, NB = (BN * BK) / 128;
  float ra[NA], rb[NB];

  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * 128;
      int m, k;
      if (TA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < g.K) ? (TA ? g.A[(size_t)gk * g.lda + gm] : g.A[(size_t)gm * g.lda + gk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * 128;
      int n, k;
      if (TB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < g.N && gk < g.K) ? (TB ? g.B[(size_t)gn * g.ldb + gk] : g.B[(size_t)gk * g.ldb + gn]) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * 128;
      int m, k;
      if (TA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * 128;
      int n, k;
      if (TB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      Bs[k][n] = rb[i];
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < g.K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float2 a = *reinterpret_cast<const float2*>(&As[kk][2 * tm]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * tn]);
      const float av[2] = {a.x, a.y}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float amx = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + 2 * tm + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tn + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.epi == EPI_BIAS_ACT) {
        v += g.bias[n];
        if (g.act == 1) v = fmaxf(v, 0.f);
        else if (g.act == 2) v = tanhf(v);
      } else if (g.epi == EPI_RELU_MASK) {
        if (n < g.mask_cols) v = (g.aux[(size_t)m * g.aux_ld + n] > 0.f) ? (g.mask_scale != 0.f ? v * g.mask_scale : v) : 0.f;
      }
      g.C[(size_t)m * g.ldc + n] = v;
      amx = fmaxf(amx, fabsf(v));
    }
  }
  if (g.absmax != nullptr) {                                       // max is order independent: an atomic keeps the result deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
    if ((tid & 31) == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(g.absmax), __float_as_int(amx));
  }
}

int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return CPP_OK;
  if (gemm_tc_wanted(g)) return launch_gemm_tc(g, s);
  dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM));
  if (g.transA && g.transB) gemm_kernel<true, true><<<grid, 128, 0, s>>>(g);
  else if (g.transA) gemm_kernel<true, false><<<grid, 128, 0, s>>>(g);
  else if (g.transB) gemm_kernel<false, true><<<grid, 128, 0, s>>>(g);
  else gemm_kernel<false, false><<<grid, 128, 0, s>>>(g);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
