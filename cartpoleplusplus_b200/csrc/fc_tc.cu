// The fully connected layers (slim.fully_connected, base_network.py:58-71; ddpg_cartpole.py:95-100,168-184;
// naf_cartpole.py:105-109,156-184) on the 5th-generation tensor cores: one tcgen05 GEMM kernel for all three passes
//   forward  h = act(x . W + b)          dgrad  dX = (dPre . W^T) * relu'(x)          wgrad  dW = x^T . dPre
// i.e. C[M,N] = opA(A)[M,K] . opB(B)[K,N] with the orientations and epilogues of fc.cu's GemmArgs.
//  * fp32 in, fp32 out, inside the 1e-5 parity budget: every fp32 operand is split into THREE bf16 pieces (8 + 8 + 8 mantissa
//    bits, fp32 exponent range: no scaling pass, no overflow flag); the six piece products whose weight is >= 2^-16 are
//    accumulated in fp32 in TMEM (kind::f16 with bf16 inputs; each product is exact), the dropped ones are <= 2^-24.
//  * no transposes: a source whose reduction index is contiguous is staged K-major, one whose row index is contiguous is
//    staged MN-major (8 consecutive rows at one k per 16-byte vector) - the instruction descriptor takes either per operand,
//    and both layouts share LBO / SBO / K-step advance.
//  * one CTA per 128 x 128 tile of C; K runs in chunks of 32 through two shared-memory buffers: all eight warps convert the
//    next chunk (global fp32 -> bf16 piece vectors) while the tensor core multiplies the previous one; an elected lane of
//    warp 0 issues the 12 instructions of a chunk and commits them to the buffer's mbarrier.
#include <cuda_bf16.h>
#include "common.cuh"
#include "umma.cuh"

namespace cpp {
namespace fctc {

using namespace umma;

constexpr int kThreads = 256;
constexpr int kKc = 32;                     // K per chunk: two instructions (K = 16) per piece product
constexpr int kTileM = 128, kTileN = 128;
constexpr int kAccs = 3;                    // accumulators per tile: leading products of even / odd K steps, all smaller products
constexpr int kPieces = 3;
constexpr int kABytes = kTileM * kKc * 2, kBBytes = kTileN * kKc * 2;        // one piece of a chunk
constexpr int kBufBytes = kPieces * (kABytes + kBBytes);
constexpr int kSmemBytes = 2 * kBufBytes + 64;

// 8 consecutive source floats -> three 16-byte vectors of bf16 pieces (x = p0 + p1 + p2 to 2^-24)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& v0, uint4& v1, uint4& v2) {
  uint32_t w0[4], w1[4], w2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t p[3][2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float f = x[2 * i + e];
      const __nv_bfloat16 b0 = __float2bfloat16_rn(f);
      const float r1 = f - __bfloat162float(b0);
      const __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(b1);
      const __nv_bfloat16 b2 = __float2bfloat16_rn(r2);
      p[0][e] = __bfloat16_as_ushort(b0); p[1][e] = __bfloat16_as_ushort(b1); p[2][e] = __bfloat16_as_ushort(b2);
    }
    w0[i] = p[0][0] | (p[0][1] << 16); w1[i] = p[1][0] | (p[1][1] << 16); w2[i] = p[2][0] | (p[2][1] << 16);
  }
  v0 = make_uint4(w0[0], w0[1], w0[2], w0[3]); v1 = make_uint4(w1[0], w1[1], w1[2], w1[3]); v2 = make_uint4(w2[0], w2[1], w2[2], w2[3]);
}

// 8 floats src[0..7], of which the first `valid` exist (the rest read as zero); vector loads when the address allows
__device__ __forceinline__ void load8(const float* src, int valid, float (&x)[8]) {
  if (valid >= 8 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = e < valid ? __ldg(src + e) : 0.f;
  }
}

// Stage one operand chunk: logical matrix T[r][k], r in [0, R) (R a multiple of 8, <= rows_valid real rows from r0), k in
// [k0, k0 + 32).  contig_k: the source is row-major in k (src[(r0 + r) * ld + k]) -> K-major core matrices; otherwise the
// source is row-major in r (src[k * ld + r0 + r]) -> MN-major vectors.  Both layouts: LBO = 128, SBO = 512, K step = 256 bytes.
__device__ __forceinline__ void stage_operand(const float* __restrict__ src, int ld, bool contig_k, int r0, int rows_valid, int R,
                                              int k0, int K, uint8_t* p0, uint8_t* p1, uint8_t* p2, int tid, int ones_row = -1) {
  if (contig_k) {
    // vectors (r, kb): 8 consecutive k of row r; consecutive threads take consecutive rows (128 contiguous bytes of shared memory)
    for (int v = tid; v < R * (kKc / 8); v += kThreads) {
      const int r = v % R, kb = v / R;
      const int k = k0 + 8 * kb;
      float x[8];
      const int valid = (r < rows_valid) ? min(8, K - k) : 0;
      if (valid > 0) load8(src + (size_t)(r0 + r) * ld + k, valid, x);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = 0.f;
      }
      uint4 a, b, c;
      split8(x, a, b, c);
      const uint32_t off = (uint32_t)(((r >> 3) * (kKc / 8) + kb) * 128 + (r & 7) * 16);
      *reinterpret_cast<uint4*>(p0 + off) = a; *reinterpret_cast<uint4*>(p1 + off) = b; *reinterpret_cast<uint4*>(p2 + off) = c;
    }
  } else {
    // vectors (rb, k): rows 8 rb .. 8 rb + 7 at one k; consecutive threads take consecutive k (contiguous shared memory)
    for (int v = tid; v < (R / 8) * kKc; v += kThreads) {
      const int kk = v % kKc, rb = v / kKc;
      const int k = k0 + kk, r = 8 * rb;
      float x[8];
      const int valid = (k < K) ? min(8, rows_valid - r) : 0;
      if (valid > 0) load8(src + (size_t)k * ld + r0 + r, valid, x);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = 0.f;
      }
      if (ones_row >= r0 + r && ones_row < r0 + r + 8 && k < K) {            // the appended all-ones row (bias gradient)
#pragma unroll
        for (int e = 0; e < 8; ++e) if (r0 + r + e == ones_row) x[e] = 1.f;
      }
      uint4 a, b, c;
      split8(x, a, b, c);
      const uint32_t off = (uint32_t)((rb * kKc + kk) * 16);
      *reinterpret_cast<uint4*>(p0 + off) = a; *reinterpret_cast<uint4*>(p1 + off) = b; *reinterpret_cast<uint4*>(p2 + off) = c;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kBufBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * kBufBytes + 32);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int m0 = blockIdx.y * kTileM, n0 = blockIdx.x * kTileN;
  const int Mt = max(0, min(kTileM, g.M - m0)), Nt = min(kTileN, g.N - n0);
  const int ones_row = (g.colsum != nullptr && g.transA) ? g.M : -1;
  const int Npad = (Nt + 15) & ~15;

  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // A: [M][K] (transA: stored [K][M]); the UMMA B operand is B^T: [N][K] (GemmArgs B is [K][N]; transB: stored [N][K])
  const bool a_contig_k = g.transA == 0, b_contig_k = g.transB != 0;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_contig_k ? 0u : 1u) << 15) | ((b_contig_k ? 0u : 1u) << 16) |
                         ((uint32_t)(Npad >> 3) << 17) | ((128u >> 4) << 24);          // fp32 D, bf16 A / B, M = 128
  const int nchunks = (g.K + kKc - 1) / kKc;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    uint8_t* base = smem + buf * kBufBytes;
    if (c >= 2) mbar_wait(&bars[buf], (uint32_t)((c >> 1) - 1) & 1);       // the instructions that read this buffer two chunks ago are done
    stage_operand(g.A, g.lda, a_contig_k, m0, Mt, kTileM, c * kKc, g.K, base, base + kABytes, base + 2 * kABytes, tid, ones_row);
    uint8_t* bb = base + kPieces * kABytes;
    stage_operand(g.B, g.ldb, b_contig_k, n0, Nt, Npad, c * kKc, g.K, bb, bb + kBBytes, bb + 2 * kBBytes, tid);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint32_t a_u = smem_u32(base), b_u = smem_u32(bb);
#pragma unroll
      for (int ks = 0; ks < kKc / 16; ++ks) {
        // piece products in decreasing weight: (0,0) (0,1) (1,0) (0,2) (1,1) (2,0).  The tensor-core accumulator truncates, so
        // the leading product alternates between two accumulators (chain length K / 32) and the five small ones share a third
        constexpr int pa[6] = {0, 0, 1, 0, 1, 2}, pb[6] = {0, 1, 0, 2, 1, 0};
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const uint64_t ad = make_desc(a_u + (uint32_t)(pa[q] * kABytes + ks * 256), 128, 512);
          const uint64_t bd = make_desc(b_u + (uint32_t)(pb[q] * kBBytes + ks * 256), 128, 512);
          const int acc = q == 0 ? ks : 2;                                   // (kKc / 16 == 2: ks is the K-step parity)
          const uint32_t accum = (c > 0 || (acc == 2 && (ks > 0 || q > 1))) ? 1u : 0u;
          if (elect_one()) umma_f16(tmem_base + (uint32_t)(acc * kTileN), ad, bd, idesc, accum);
        }
      }
      if (elect_one()) umma_commit(&bars[buf]);
      __syncwarp();
    }
  }
  // the last commit covers every earlier instruction
  {
    const int last = nchunks - 1;
    mbar_wait(&bars[last & 1], (uint32_t)(last >> 1) & 1);
    tc_fence_after();
  }
  // epilogue: warp w drains TMEM lanes 32 (w % 4) .. + 31 (= rows of the tile), columns of half w / 4
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int row = 32 * quarter + lane, m = m0 + row;
    const int cols_half = ((Npad / 8 + 1) / 2) * 8;                          // column split in whole 8-column groups
    const int c_begin = half * cols_half, c_end = min(Npad, c_begin + cols_half);
    for (int c0 = c_begin; c0 < c_end; c0 += 8) {
      float v8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v8[e] = 0.f;
#pragma unroll
      for (int a = kAccs - 1; a >= 0; --a) {                                 // small products first
        uint32_t r[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(a * kTileN + c0);
        tmem_ld4(taddr, r); tmem_ld4(taddr + 4, r + 4);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) v8[e] += __uint_as_float(r[e]);
      }
      if (m < g.M || m == ones_row) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int n = n0 + c0 + e;
          if (n < g.N) {
            float v = v8[e];
            if (m == ones_row) { g.colsum[n] = v; continue; }
            if (g.epi == EPI_BIAS_ACT) {
              v += g.bias[n];
              if (g.act == 1) v = fmaxf(v, 0.f);
              else if (g.act == 2) v = tanhf(v);
            } else if (g.epi == EPI_RELU_MASK) {
              if (n < g.mask_cols) v = (g.aux[(size_t)m * g.aux_ld + n] > 0.f) ? (g.mask_scale != 0.f ? v * g.mask_scale : v) : 0.f;
            }
            g.C[(size_t)m * g.ldc + n] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace fctc

// shapes worth a 128-row tensor-core tile; the heads (N = 1, 2) and single-sample calls stay on the FFMA kernel of fc.cu
// g_fc_tc: mask over the passes - 1 forward (x . W), 2 input gradient (dPre . W^T), 4 weight gradient (x^T . dPre) - plus
// 8 = "also the small ones".  Without bit 3 a GEMM goes to the tensor cores only when it is large enough to pay for a 128-row
// tile per CTA (TMEM allocation, piece conversion of both operands by the CTA's own threads).  Default 0: measured on B200
// inside the fused c3 step (batch 256, K <= 640: 16 M MACs per GEMM) the tcgen05 kernel costs +24 us per step on the weight
// gradients (side streams) and +110 us on the input-gradient chain, and at the single-GPU c5 size (batch 1024 x 2560 x 200) it
// is 2 % behind as well (profiles/r4/fc_tc.md): these GEMMs are latency bound, and a CTA that owns all of an SM's TMEM
// cannot slip in next to the persistent conv kernels the way a 128-thread FFMA block does.
constexpr int64_t kAutoMinMacs = 64ll << 20;
bool gemm_tc_wanted(const GemmArgs& g) {
  const int pass = g.transA ? 4 : (g.transB ? 2 : 1);
  if ((g_fc_tc & pass) == 0 || g.M < 64 || g.N < 8 || g.K < 16) return false;
  return (g_fc_tc & 8) != 0 || (int64_t)g.M * g.N * g.K >= kAutoMinMacs;
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return CPP_OK;
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(fctc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fctc::kSmemBytes));
    configured = true;
  }
  const int m_ext = g.M + ((g.colsum != nullptr && g.transA) ? 1 : 0);
  dim3 grid((unsigned)ceil_div(g.N, fctc::kTileN), (unsigned)ceil_div(m_ext, fctc::kTileM));
  fctc::gemm_tc_kernel<<<grid, fctc::kThreads, fctc::kSmemBytes, s>>>(g);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
