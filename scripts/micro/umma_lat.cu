// (variant of umma_mn.cu that also times issue -> tcgen05.commit arrival of a short dependent chain: the latency of ONE hand-over)
// Micro test: one tcgen05.mma kind::f16 with BOTH operands MN-major, no swizzle, read from "plane" layouts
//   A[m][k] = Aplane[m / 8][k][m % 8]   (16-byte vector per (8-row block, k); k = pixel index, contiguous vectors)
//   B[n][k] = Bplane[n / 8][k][n % 8]
// which is what the tcgen05 conv weight-gradient kernel needs (reduction over pixels).  Checks the result against the
// host for the two possible readings of the descriptor's LBO / SBO fields and prints which one is right.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_mn umma_mn.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr int M = 128, KP = 32;      // KP pixels per plane row = 2 MMAs of K = 16

__global__ void __launch_bounds__(128, 1) k(int N, int variant, const __half* Ag, const __half* Bg, float* out, long long* lat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  const int a_bytes = (M / 8) * KP * 16, b_bytes = (N / 8) * KP * 16;
  for (int i = tid; i < a_bytes / 2; i += 128) reinterpret_cast<__half*>(smem)[i] = Ag[i];
  for (int i = tid; i < b_bytes / 2; i += 128) reinterpret_cast<__half*>(smem + a_bytes)[i] = Bg[i];
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  // D fp32, A/B fp16, A and B MN-major (bits 15, 16)
  const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const long long t0 = clock64();
  if (tid == 0) {
    const uint32_t plane = KP * 16;                                // byte stride between 8-row blocks
    for (int ks = 0; ks < KP / 16; ++ks) {
      const uint32_t a = smem_u32(smem) + ks * 256, b = smem_u32(smem + a_bytes) + ks * 256;
      const uint64_t ad = variant == 0 ? make_desc(a, 128, plane) : make_desc(a, plane, 128);
      const uint64_t bd = variant == 0 ? make_desc(b, 128, plane) : make_desc(b, plane, 128);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 32) lat[0] = clock64() - t0;      // another warp's view: issue of 2 MMAs -> commit -> barrier observed
  const int warp = tid >> 5;
  for (int c = 0; c < N; c += 4) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tmem + ((uint32_t)(32 * warp) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 4; ++j) out[tid * N + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;     // run one variant per process: a bad descriptor may poison the context
  for (int N : {48, 32, 64, 128, 240}) {
    std::vector<__half> A((M / 8) * KP * 8), B((N / 8) * KP * 8);
    std::vector<float> Af(M * KP), Bf(N * KP);
    srand(1);
    for (int m = 0; m < M; ++m) for (int p = 0; p < KP; ++p) {
      const float v = (float)(rand() % 7 - 3); Af[m * KP + p] = v; A[((m / 8) * KP + p) * 8 + m % 8] = __float2half(v);
    }
    for (int n = 0; n < N; ++n) for (int p = 0; p < KP; ++p) {
      const float v = (float)(rand() % 5 - 2); Bf[n * KP + p] = v; B[((n / 8) * KP + p) * 8 + n % 8] = __float2half(v);
    }
    __half *dA, *dB; float* dO; long long* dL; cudaMalloc(&dL, 8);
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant) {
      if (only >= 0 && variant != only) continue;
      cudaMemset(dO, 0, M * N * 4);
      k<<<1, 128, 32 * 1024>>>(N, variant, dA, dB, dO, dL);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> O(M * N);
      cudaMemcpy(O.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float ref = 0; for (int p = 0; p < KP; ++p) ref += Af[m * KP + p] * Bf[n * KP + p];
        if (O[m * N + n] != ref) ++bad;
      }
      long long hl = 0; cudaMemcpy(&hl, dL, 8, cudaMemcpyDeviceToHost);
      printf("[2 MMAs issue -> commit observed by another warp: %lld clk] ", hl);
      printf("N=%d variant %d (%s): %s, %d / %d mismatches\n", N, variant, variant == 0 ? "LBO=k-block 128B, SBO=mn-block plane" : "LBO=mn-block plane, SBO=k-block 128B",
             cudaGetErrorString(e), bad, M * N);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  return 0;
}
