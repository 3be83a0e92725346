// tcgen05 implicit-GEMM forward convolution of the reference conv trunk (base_network.py:73-127):
//   y_n = maxpool2x2(relu(conv_same(whiten(x), w_n) + b_n))   for up to 3 sibling networks n in one pass.
//
// Why it looks the way it does (DESIGN.md "conv on tensor cores"):
//  * Every conv of the reference has 10 filters, so as a GEMM it is M = pixels, N = 10, K = taps x Cin.  With the
//    pixels as the UMMA "A" operand (M = 128 rows per instruction) the tensor pipe is limited by the shared-memory
//    read of A, not by N: widening N up to 64 is free.  N is therefore filled with (a) sibling networks that read
//    the same input (actor+critic on state_1, the two targets on state_2, NAF's value/mu/l) and (b) two fp16
//    "pieces" (hi + lo, 22 mantissa bits) of every fp32 weight, which keeps the result inside the 1e-5 parity
//    budget: the fp16 replay pixels are exact fp16 operands, products are exact in the fp32 accumulator.
//  * The whitening x^ = (x - mean_c) * inv_c (base_network.py:95-99) is folded into the weights
//    (w' = w * inv_c) and into a border-aware additive term (zero padding applies to x^, not to x): the kernel
//    multiplies RAW pixels, the epilogue adds  b - sum_{taps inside the image} mean_c inv_c w.
//  * im2col-free: a unit of the image is staged ONCE in shared memory as "parity planes" -
//    plane(yp, xp)[yh][xh] = the 8-channel (16 byte) vector of pixel (2*yh + yp, 2*xh + xp) - so that for every
//    filter tap the A operand of 128 consecutive pooled positions is the same plane shifted by a whole number of
//    16-byte rows: a no-swizzle K-major UMMA descriptor pointing INTO the plane.  The four positions of every 2x2
//    max-pool window get their own TMEM accumulator, so the pool is a per-lane max in the epilogue.
//    Channels beyond a multiple of 8 (9 = 8 + 1, 18 = 16 + 2, 10 = 8 + 2) are packed along kx into extra planes
//    (5 taps x r channels per 16-byte row) instead of padding the channel group with zeros.
//  * Raw image rows come in through the TMA engine (cp.async.bulk global -> shared, mbarrier completion) and are
//    re-laid into the planes by all warps; one elected thread issues the tcgen05.mma stream; all warps drain
//    TMEM (tcgen05.ld), apply scale / bias / ReLU / max-pool and write pooled values + argmax bytes.
#include <vector>
#include <algorithm>
#include "conv_tc.cuh"

namespace cpp {
namespace tc {

constexpr int CO = kConvCout;
constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();   // a lost arrival must fail loudly, never hang the GPU
  } while (!ok);
}
// TMA engine, non-tensor form: contiguous global -> shared bulk copy completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// shared-memory matrix descriptor, no swizzle ("interleave"), K-major: 8 rows x 16 bytes core matrices,
// LBO = byte distance between the two K8 halves of a K16 instruction, SBO = between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  return d;
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 x fp16 -> fp32), issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// instruction descriptor: D fp32, A/B fp16, both K-major, M = 128, N
__host__ __device__ inline uint32_t make_idesc(int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ weight packing
// One block.  (1) S: power of two that brings max |w * inv| just under 2^15; (2) B operand in canonical K-major
// layout [pair][half][n][8], n = (net * pieces + piece) * 10 + o; (3) corr[yc][xc][net][o] = bias - sum over the
// taps that fall inside the image of mean_c * inv_c * w (fp64), and 2^-S.
__global__ void __launch_bounds__(256) conv_tc_prep_kernel(const __grid_constant__ FwdPlan P, const __grid_constant__ PrepArgs A) {
  __shared__ float red[256];
  __shared__ float s_scale;
  const int tid = threadIdx.x;
  const int KS = P.KS, C = P.C, nw = KS * KS * C * CO;
  float mx = 0.f;
  for (int n = 0; n < P.nets; ++n)
    for (int i = tid; i < nw; i += blockDim.x) {
      const int ch = (i / CO) % C;
      const float inv = A.mean_inv ? A.mean_inv[C + ch] : 1.f;
      mx = fmaxf(mx, fabsf(A.w[n][i] * inv));
    }
  red[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]); __syncthreads(); }
  if (tid == 0) {
    int e = 0;
    const float m = red[0];
    if (m > 0.f && isfinite(m)) frexpf(m, &e);      // m < 2^e
    else e = 15;
    s_scale = ldexpf(1.f, 15 - e);
    A.corr[(2 * P.PAD + 1) * (2 * P.PAD + 1) * P.nets * CO] = ldexpf(1.f, e - 15);
  }
  __syncthreads();
  const float scale = s_scale;
  const int N = P.N, used = P.nets * kPieces * CO, total = P.n_pairs * 2 * N * 8;
  for (int idx = tid; idx < total; idx += blockDim.x) {
    const int e = idx & 7, n = (idx >> 3) % N, h = ((idx >> 3) / N) & 1, i = (idx >> 3) / (2 * N);
    const Slab sl = P.slab[i][h];
    float v = 0.f;
    bool valid = n < used && sl.kind != 2;
    int ky = sl.ky, kx = sl.kx, ch = 0;
    if (valid) {
      if (sl.kind == 0) { ch = 8 * sl.set + e; valid = ch < C; }
      else { const int E = 8 * sl.set + e; valid = E < KS * P.R; kx = E / P.R; ch = 8 * P.G8 + E % P.R; }
    }
    __half out = __float2half_rn(0.f);
    if (valid) {
      const int net = n / (kPieces * CO), piece = (n / CO) % kPieces, o = n % CO;
      const float inv = A.mean_inv ? A.mean_inv[C + ch] : 1.f;
      v = A.w[net][((ky * KS + kx) * C + ch) * CO + o] * inv * scale;
      const __half hi = __float2half_rn(v);
      out = (piece == 0) ? hi : __float2half_rn(v - __half2float(hi));
    }
    A.bpack[idx] = out;
  }
  const int ncls = 2 * P.PAD + 1;
  for (int idx = tid; idx < ncls * ncls * P.nets * CO; idx += blockDim.x) {
    const int o = idx % CO, net = (idx / CO) % P.nets, xc = (idx / (CO * P.nets)) % ncls, yc = idx / (CO * P.nets * ncls);
    double acc = (double)A.bias[net][o];
    if (A.mean_inv) {
      const int ky0 = yc < P.PAD ? P.PAD - yc : 0, ky1 = yc > P.PAD ? KS - 1 - (yc - P.PAD) : KS - 1;
      const int kx0 = xc < P.PAD ? P.PAD - xc : 0, kx1 = xc > P.PAD ? KS - 1 - (xc - P.PAD) : KS - 1;
      for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx)
          for (int ch = 0; ch < C; ++ch)
            acc -= (double)A.mean_inv[ch] * (double)A.mean_inv[C + ch] * (double)A.w[net][((ky * KS + kx) * C + ch) * CO + o];
    }
    A.corr[idx] = (float)acc;
  }
}

// ------------------------------------------------------------------------------------------ main kernel
struct SmemLayout { uint32_t planes, bsm, stage, corr, bars, tmem, total; };

__host__ __device__ inline SmemLayout smem_layout(const FwdPlan& P) {
  SmemLayout L;
  uint32_t off = 0;
  L.planes = off; off += (uint32_t)P.n_planes * P.plane_bytes;
  L.bsm = off; off += (uint32_t)P.n_pairs * 2 * P.N * 16;
  L.stage = off; off += (uint32_t)P.stage_bytes;
  const int ncls = 2 * P.PAD + 1;
  L.corr = off; off += (uint32_t)((ncls * ncls * P.nets * CO + 4) * 4);
  off = (off + 15) & ~15u;
  L.bars = off; off += 32;
  L.tmem = off; off += 16;
  L.total = off;
  return L;
}

__global__ void __launch_bounds__(kThreads, 2) conv_fwd_tc_kernel(const __grid_constant__ FwdPlan P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout L = smem_layout(P);
  uint8_t* planes = smem + L.planes;
  uint8_t* bsm = smem + L.bsm;
  __half* stage = reinterpret_cast<__half*>(smem + L.stage);
  float* corr_s = reinterpret_cast<float*>(smem + L.corr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);     // [0] staging full, [1] accumulators ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = P.H, W = P.W, C = P.C, PH = P.PH, PW = P.PW, Pq = P.Pq, KS = P.KS, PAD = P.PAD, N = P.N;
  const int ncls = 2 * PAD + 1, ncorr = ncls * ncls * P.nets * CO;

  // ---- one-time setup: barriers, TMEM, weights, correction table
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.bpack);
    uint4* dst = reinterpret_cast<uint4*>(bsm);
    for (int i = tid; i < P.n_pairs * 2 * N; i += kThreads) dst[i] = src[i];
    for (int i = tid; i < ncorr + 1; i += kThreads) corr_s[i] = P.corr[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float scale_inv = corr_s[ncorr];
  const uint32_t idesc = make_idesc(N);
  const uint32_t planes_addr = smem_u32(planes), bsm_addr = smem_u32(bsm);
  uint32_t ph_stage = 0, ph_mma = 0;
  const size_t img_elems = (size_t)H * W * C;
  const int n_groups = (P.rows_alloc + P.crh - 1) / P.crh;

  for (int unit = blockIdx.x; unit < P.n_units; unit += gridDim.x) {
    const int b = unit / P.units_per_image, u = unit - b * P.units_per_image;
    const int t0 = u * P.tiles_per_unit, t1 = min(t0 + P.tiles_per_unit, P.tiles_per_image);
    const int py_first = (128 * t0) / Pq, yh0 = py_first - 1;
    const __half* img = P.x + (size_t)(P.rows ? P.rows[b] : b) * img_elems;

    // ---- fill the parity planes of this unit, one staging group (crh parity rows = 2*crh raw rows) at a time
    for (int g = 0; g < n_groups; ++g) {
      const int rho0 = g * P.crh, rho1 = min(rho0 + P.crh, P.rows_alloc);
      const int yc0 = max(0, 2 * (yh0 + rho0)), yc1 = min(H, 2 * (yh0 + rho1));
      const bool have = yc1 > yc0;
      if (have) {
        const uint32_t bytes = (uint32_t)(yc1 - yc0) * W * C * 2;
        if (P.use_bulk) {
          if (tid == 0) {
            fence_proxy_async();                       // staging was read through the generic proxy just before
            mbar_expect_tx(&bars[0], bytes);
            bulk_g2s(stage, img + (size_t)yc0 * W * C, bytes, &bars[0]);
          }
          mbar_wait(&bars[0], ph_stage);
          ph_stage ^= 1;
        } else {
          const __half* src = img + (size_t)yc0 * W * C;
          for (int i = tid; i < (yc1 - yc0) * W * C; i += kThreads) stage[i] = src[i];
          __syncthreads();
        }
      }
      const int per_plane = (rho1 - rho0) * Pq;
      for (int it = tid; it < P.n_planes * per_plane; it += kThreads) {
        const int pl = it / per_plane, rem = it - pl * per_plane;
        const int rho = rho0 + rem / Pq, kap = rem % Pq;
        const int yh = yh0 + rho;
        __align__(16) __half v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __float2half_rn(0.f);
        if (pl < 4 * P.G8) {                           // 8 channels of pixel (2*yh + yp, 2*(kap-1) + xp)
          const int gq = pl >> 2, yp = (pl >> 1) & 1, xp = pl & 1;
          const int y = 2 * yh + yp, x = 2 * (kap - 1) + xp;
          if (have && y >= yc0 && y < yc1 && x >= 0 && x < W) {
            const __half* sp = stage + ((size_t)(y - yc0) * W + x) * C + 8 * gq;
            const int nch = min(8, C - 8 * gq);
#pragma unroll
            for (int e = 0; e < 8; ++e) if (e < nch) v[e] = sp[e];
          }
        } else {                                       // remainder channels packed along kx
          const int q = pl - 4 * P.G8, j = q >> 2, yp = (q >> 1) & 1, dx = q & 1;
          const int y = 2 * yh + yp, px = kap - 1;
          if (have && y >= yc0 && y < yc1 && px >= 0 && px < PW) {
            const __half* sp = stage + (size_t)(y - yc0) * W * C + 8 * P.G8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int E = 8 * j + e;
              if (E < KS * P.R) {
                const int kx = E / P.R, cr = E - kx * P.R;
                const int x = 2 * px + dx + kx - PAD;
                if (x >= 0 && x < W) v[e] = sp[(size_t)x * C + cr];
              }
            }
          }
        }
        *reinterpret_cast<uint4*>(planes + (size_t)pl * P.plane_bytes + ((size_t)rho * Pq + kap) * 16) = *reinterpret_cast<const uint4*>(v);
      }
      __syncthreads();                                 // staging is free again
    }
    fence_proxy_async();                               // plane writes (generic proxy) -> visible to the tensor core (async proxy)
    __syncthreads();

    for (int t = t0; t < t1; ++t) {
      // ---- MMA: 4 accumulators (one per position of the 2x2 pool window) x n_pairs K16 instructions
      if (warp == 0) {
        if (lane == 0) {
          tc_fence_after();
          const int q_off = 128 * t - py_first * Pq;
          for (int a = 0; a < 4; ++a) {
            const int dy = a >> 1, dx = a & 1;
            for (int i = 0; i < P.n_pairs; ++i) {
              uint32_t addr[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const Slab sl = P.slab[i][h];
                const int ty = dy + sl.ky - PAD, yp = ty & 1, oy = (ty >> 1) + 1;
                if (sl.kind == 0) {
                  const int tx = dx + sl.kx - PAD, xp = tx & 1, ox = (tx >> 1) + 1;
                  addr[h] = planes_addr + (uint32_t)((sl.set * 4 + yp * 2 + xp) * P.plane_bytes + (q_off + oy * Pq + ox) * 16);
                } else if (sl.kind == 1) {
                  addr[h] = planes_addr + (uint32_t)((4 * P.G8 + sl.set * 4 + yp * 2 + dx) * P.plane_bytes + (q_off + oy * Pq + 1) * 16);
                } else {
                  addr[h] = addr[0] + 16;              // zero weights: any initialised shared memory will do
                }
              }
              const uint64_t adesc = make_desc(addr[0], addr[1] - addr[0], 128);
              const uint64_t bdesc = make_desc(bsm_addr + (uint32_t)i * 2 * N * 16, (uint32_t)N * 16, 128);
              umma_f16(tmem_base + (uint32_t)(a * N), adesc, bdesc, idesc, i > 0 ? 1u : 0u);
            }
          }
          umma_commit(&bars[1]);
        }
        __syncwarp();
      }
      mbar_wait(&bars[1], ph_mma);
      ph_mma ^= 1;
      tc_fence_after();

      // ---- epilogue: lane = pooled position; sum the weight pieces, rescale, add the border-aware bias, max-pool
      const int quarter = warp & 3;
      const int q = 128 * t + 32 * quarter + lane;
      const int py = q / Pq, px = q - py * Pq;
      const bool valid = py < PH && px < PW;
      for (int net = warp >> 2; net < P.nets; net += kThreads / 128) {
        float best[CO];
        int arg[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) { best[o] = 0.f; arg[o] = 0; }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          uint32_t r[20];
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(a * N + net * kPieces * CO);
          tmem_ld16(taddr, r);
          tmem_ld4(taddr + 16, r + 16);
          tmem_ld_wait();
          const int y = 2 * py + (a >> 1), x = 2 * px + (a & 1);
          const int yc = y < PAD ? y : (y >= H - PAD ? 2 * PAD - (H - 1 - y) : PAD);
          const int xc = x < PAD ? x : (x >= W - PAD ? 2 * PAD - (W - 1 - x) : PAD);
          const float* cr = corr_s + ((valid ? (yc * ncls + xc) : 0) * P.nets + net) * CO;
#pragma unroll
          for (int o = 0; o < CO; ++o) {
            const float v = fmaf(__uint_as_float(r[o]) + __uint_as_float(r[CO + o]), scale_inv, cr[o]);
            if (a == 0) { best[o] = v; arg[o] = 0; }
            else if (v > best[o]) { best[o] = v; arg[o] = a; }
          }
        }
        if (valid) {
          const size_t base = (((size_t)b * PH + py) * PW + px) * CO;
          float* op = P.pooled[net] + base;
          uint8_t* ap = P.amax[net] + base;
#pragma unroll
          for (int o = 0; o < CO; o += 2) {
            *reinterpret_cast<float2*>(op + o) = make_float2(fmaxf(best[o], 0.f), fmaxf(best[o + 1], 0.f));
            const uint16_t pk = (uint16_t)((best[o] > 0.f ? arg[o] : 4) | ((best[o + 1] > 0.f ? arg[o + 1] : 4) << 8));
            *reinterpret_cast<uint16_t*>(ap + o) = pk;
          }
        }
      }
      tc_fence_before();
      __syncthreads();                                 // TMEM and (after the last tile) the planes are free again
    }
  }
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------ host: plan
static int build_plan(int nets, int B, int H, int W, int C, int KS, FwdPlan* P) {
  CPP_REQUIRE(KS == 5 || KS == 3, "conv_tc: kernel size %d", KS);
  CPP_REQUIRE(nets >= 1 && nets <= kMaxNets, "conv_tc: %d sibling networks", nets);
  const int PAD = KS / 2;
  CPP_REQUIRE(H >= 2 * PAD && W >= 2 * PAD && H >= 2 && W >= 2, "conv_tc: input %dx%d too small", H, W);
  P->B = B; P->H = H; P->W = W; P->C = C; P->KS = KS; P->PAD = PAD;
  P->PH = H / 2; P->PW = W / 2;
  P->Pq = (W + 1) / 2 + 1;      // parity-plane pitch: ceil(W/2) pixels + one zero column shared by neighbouring rows
  P->nets = nets; P->N = (int)round_up(nets * kPieces * CO, 16);
  CPP_REQUIRE(4 * P->N <= 256, "conv_tc: N=%d does not fit the TMEM allocation", P->N);
  const int rem = C % 8;
  if (rem == 1 || rem == 2) { P->G8 = C / 8; P->R = rem; P->nR = (KS * rem + 7) / 8; }
  else { P->G8 = (C + 7) / 8; P->R = 0; P->nR = 0; }
  P->n_planes = 4 * (P->G8 + P->nR);

  // slab pairs: both K8 halves of an instruction live in the same plane set with the second at a higher address
  // for every pool position (same tap parity class, later tap), or in a later plane set
  std::vector<Slab> left;
  int np = 0;
  auto emit = [&](const Slab& a, const Slab& b) -> bool {
    if (np >= kMaxPairs) return false;
    P->slab[np][0] = a; P->slab[np][1] = b; ++np;
    return true;
  };
  bool ok = true;
  for (int g = 0; g < P->G8; ++g)
    for (int cls = 0; cls < 4; ++cls) {
      std::vector<Slab> v;
      for (int ky = 0; ky < KS; ++ky)
        for (int kx = 0; kx < KS; ++kx)
          if ((ky & 1) == (cls >> 1) && (kx & 1) == (cls & 1)) v.push_back(Slab{0, (int8_t)g, (int8_t)ky, (int8_t)kx});
      size_t i = 0;
      for (; i + 1 < v.size(); i += 2) ok = ok && emit(v[i], v[i + 1]);
      if (i < v.size()) left.push_back(v[i]);
    }
  for (int j = 0; j < P->nR; ++j)
    for (int cls = 0; cls < 2; ++cls) {
      std::vector<Slab> v;
      for (int ky = 0; ky < KS; ++ky)
        if ((ky & 1) == cls) v.push_back(Slab{1, (int8_t)j, (int8_t)ky, 0});
      size_t i = 0;
      for (; i + 1 < v.size(); i += 2) ok = ok && emit(v[i], v[i + 1]);
      if (i < v.size()) left.push_back(v[i]);
    }
  // leftovers come from distinct plane sets only if each set left at most one; sort by set order and verify
  std::stable_sort(left.begin(), left.end(), [](const Slab& a, const Slab& b) {
    return a.kind != b.kind ? a.kind < b.kind : a.set < b.set;
  });
  for (size_t i = 0; i + 1 < left.size(); ++i)
    CPP_REQUIRE(left[i].kind != left[i + 1].kind || left[i].set != left[i + 1].set, "conv_tc: slab pairing failed");
  {
    size_t i = 0;
    for (; i + 1 < left.size(); i += 2) ok = ok && emit(left[i], left[i + 1]);
    if (i < left.size()) ok = ok && emit(left[i], Slab{2, 0, 0, 0});
  }
  CPP_REQUIRE(ok, "conv_tc: too many K slabs (C=%d)", C);
  P->n_pairs = np;

  P->tiles_per_image = (int)ceil_div((int64_t)P->PH * P->Pq, 128);
  const int row_bytes = W * C * 2;
  P->crh = std::max(1, 8192 / (2 * row_bytes));
  P->stage_bytes = (int)round_up((int64_t)2 * P->crh * row_bytes, 16);
  P->use_bulk = (row_bytes % 16 == 0) && (((size_t)H * W * C * 2) % 16 == 0);
  int best_tpu = 0;
  for (int pass = 0; pass < 2 && best_tpu == 0; ++pass) {
    const int limit = pass == 0 ? 110 * 1024 : 220 * 1024;      // two CTAs per SM if possible
    for (int tpu = std::min(P->tiles_per_image, 8); tpu >= 1; --tpu) {
      P->tiles_per_unit = tpu;
      P->rows_alloc = (P->Pq - 1 + 128 * tpu - 1 + 2 * P->Pq + 2) / P->Pq + 1;
      P->plane_bytes = P->rows_alloc * P->Pq * 16;
      if ((int)smem_layout(*P).total <= limit) { best_tpu = tpu; break; }
    }
  }
  CPP_REQUIRE(best_tpu > 0, "conv_tc: %dx%dx%d does not fit shared memory", H, W, C);
  P->units_per_image = (int)ceil_div(P->tiles_per_image, best_tpu);
  P->tiles_per_unit = (int)ceil_div(P->tiles_per_image, P->units_per_image);
  P->rows_alloc = (P->Pq - 1 + 128 * P->tiles_per_unit - 1 + 2 * P->Pq + 2) / P->Pq + 1;
  P->plane_bytes = P->rows_alloc * P->Pq * 16;
  P->n_units = B * P->units_per_image;
  P->smem_bytes = (int)smem_layout(*P).total;
  return CPP_OK;
}

static inline size_t bpack_bytes(const FwdPlan& P) { return (size_t)round_up((int64_t)P.n_pairs * 2 * P.N * 16, 256); }

bool conv_tc_supported(int nets, int H, int W, int C, int KS) {
  FwdPlan P{};
  const bool ok = build_plan(nets, 1, H, W, C, KS, &P) == CPP_OK;
  return ok;
}

int64_t conv_tc_scratch_bytes(int nets, int H, int W, int C, int KS) {
  FwdPlan P{};
  if (build_plan(nets, 1, H, W, C, KS, &P) != CPP_OK) return -1;
  const int ncls = 2 * P.PAD + 1;
  return (int64_t)bpack_bytes(P) + (int64_t)round_up((int64_t)(ncls * ncls * nets * CO + 4) * 4, 256);
}

int launch_conv_fwd_tc(const void* x_f16, const int32_t* rows, const float* mean_inv, int nets,
                       const float* const* w, const float* const* bias, int B, int H, int W, int C, int KS,
                       float* const* pooled, uint8_t* const* amax, void* scratch, cudaStream_t s) {
  if (B <= 0) return CPP_OK;
  FwdPlan P{};
  CPP_TRY(build_plan(nets, B, H, W, C, KS, &P));
  CPP_REQUIRE(((uintptr_t)x_f16 & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "conv_tc: unaligned buffers");
  P.x = reinterpret_cast<const __half*>(x_f16); P.rows = rows;
  P.bpack = reinterpret_cast<const __half*>(scratch);
  P.corr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(scratch) + bpack_bytes(P));
  PrepArgs A{};
  for (int n = 0; n < nets; ++n) {
    CPP_REQUIRE(w[n] && bias[n] && pooled[n] && amax[n], "conv_tc: null pointer for network %d", n);
    A.w[n] = w[n]; A.bias[n] = bias[n]; P.pooled[n] = pooled[n]; P.amax[n] = amax[n];
  }
  A.mean_inv = mean_inv;
  A.bpack = const_cast<__half*>(P.bpack); A.corr = const_cast<float*>(P.corr);
  conv_tc_prep_kernel<<<1, 256, 0, s>>>(P, A);
  CPP_CHECK_LAUNCH();
  static int configured = 0;
  if (P.smem_bytes > configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = 220 * 1024;
  }
  const int ctas_per_sm = P.smem_bytes <= 110 * 1024 ? 2 : 1;
  const int grid = std::min(P.n_units, ctas_per_sm * kNumSMs);
  conv_fwd_tc_kernel<<<grid, kThreads, P.smem_bytes, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace tc
}  // namespace cpp
