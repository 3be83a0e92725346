// (MN-major variant of umma_rate.cu: both operands in the pixel-major plane layout of conv_wgrad_tc.cu)
// Microbenchmark: cycles per tcgen05.mma (kind::f16, cta_group::1, M=128/64) as a function of N, of the
// shared-memory layout (no swizzle vs 128B swizzle) and of the A start-address alignment.  One CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__global__ void __launch_bounds__(128, 1) k(int M, int N, int iters, int layout, int sbo_sw, int a_shift, int a_step, int n_acc, long long* out, int lbo = 2064) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  const bool mn = layout >= 100; if (mn) layout = 0;
  const uint32_t idesc = (1u << 4) | (mn ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  long long t0 = 0, t1 = 0;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == 0) {
    const uint32_t a_base = smem_u32(smem) + a_shift, b_base = smem_u32(smem) + 128 * 1024;
    const uint32_t hi_mn_a = ((uint32_t)(sbo_sw >> 4) | (1u << 14)), hi_mn_b = ((1024u >> 4) | (1u << 14));
    const uint32_t hi_a = layout == 0 ? ((128u >> 4) | (1u << 14)) : ((uint32_t)(sbo_sw >> 4) | (1u << 14) | ((uint32_t)layout << 29));
    const uint32_t lo_a0 = layout == 0 ? (((a_base & 0x3FFFF) >> 4) | (((uint32_t)lbo >> 4) << 16)) : (((a_base & 0x3FFFF) >> 4) | (1u << 16));
    const uint32_t lo_b = layout == 0 ? (((b_base & 0x3FFFF) >> 4) | ((uint32_t)N << 16)) : (((b_base & 0x3FFFF) >> 4) | (1u << 16));
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t ad = mn ? (((uint64_t)hi_mn_a << 32) | (uint64_t)((((a_base & 0x3FFFF) >> 4) | (8u << 16)) + (uint32_t)((i & 15) * (a_step >> 4)))) : (((uint64_t)hi_a << 32) | (uint64_t)(lo_a0 + (uint32_t)((i & 15) * (a_step >> 4))));
      const uint64_t bd = mn ? (((uint64_t)hi_mn_b << 32) | (uint64_t)(((b_base & 0x3FFFF) >> 4) | (8u << 16))) : (((uint64_t)hi_a << 32) | (uint64_t)lo_b);
      const uint32_t d = tmem + (uint32_t)((i & (n_acc - 1)) * N);
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i >= n_acc ? 1u : 0u) : "memory");
    }
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    } while (!ok);
    t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
  long long* out; cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  auto run = [&](int M, int N, int layout, int sbo, int step, int nacc) {
    k<<<148, 128, 200 * 1024>>>(M, N, iters, layout, sbo, 0, step, nacc, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("M=%3d N=%3d %s A-SBO=%5d step=%4d nacc=%d : %7.1f cyc/mma  (%s)\n", M, N, layout >= 100 ? "MN-major" : "K-major ", sbo, step, nacc, (double)mx / iters, cudaGetErrorString(e));
  };
  for (int N : {16, 32, 48, 64, 128}) run(128, N, 0, 0, 0, 4);
  for (int N : {16, 32, 48, 64, 128}) run(128, N, 100, 8192, 0, 4);
  for (int N : {32, 48}) run(128, N, 100, 8192, 256, 4);
  for (int N : {32, 48}) run(128, N, 100, 8192, 1024, 2);
  for (int N : {32, 48}) run(128, N, 100, 2048, 256, 4);
  for (int N : {32, 48}) run(128, N, 100, 8192 + 128, 256, 4);
  for (int N : {32, 48, 64}) run(64, N, 100, 8192, 256, 4);
  return 0;
}
