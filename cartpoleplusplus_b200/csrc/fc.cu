// FP32 GEMM for the fully connected layers (slim.fully_connected, base_network.py:58-71;
// ddpg_cartpole.py:95-100,168-184; naf_cartpole.py:105-109,156-184): forward, dgrad and wgrad are the
// same kernel with different operand orientations.  The layers are tiny (K <= 2560, N <= 200) and
// launch/latency bound; exact fp32 FFMA keeps the 1e-5 parity budget for the conv path.
#include "common.cuh"

namespace cpp {

constexpr int BM = 32, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(128) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tm = tid / 16, tn = tid % 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + i * 128;
      int m, k;
      if (TA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < g.K) v = TA ? g.A[(size_t)gk * g.lda + gm] : g.A[(size_t)gm * g.lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 128; ++i) {
      const int e = tid + i * 128;
      int n, k;
      if (TB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < g.K) v = TB ? g.B[(size_t)gn * g.ldb + gk] : g.B[(size_t)gk * g.ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][4 * tm]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + 4 * tm + i;
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
        if (n < g.mask_cols && !(g.aux[(size_t)m * g.aux_ld + n] > 0.f)) v = 0.f;
      }
      g.C[(size_t)m * g.ldc + n] = v;
    }
  }
}

int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return CPP_OK;
  dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM));
  if (g.transA && g.transB) gemm_kernel<true, true><<<grid, 128, 0, s>>>(g);
  else if (g.transA) gemm_kernel<true, false><<<grid, 128, 0, s>>>(g);
  else if (g.transB) gemm_kernel<false, true><<<grid, 128, 0, s>>>(g);
  else gemm_kernel<false, false><<<grid, 128, 0, s>>>(g);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
