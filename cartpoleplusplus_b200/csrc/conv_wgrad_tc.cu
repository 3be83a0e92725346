// conv1 weight / bias gradient on tcgen05 (base_network.py:103-107 backwards; see conv_wgrad_mma.cu for the algebra):
//
//   G[(ky,kx,c), (n,piece,o)] = sum over output pixels p of  X[p + (ky,kx)][c] * dY_n[p][o]
//
// The reduction runs over PIXELS, so both UMMA operands are MN-major: a 16-byte shared-memory vector holds 8 consecutive
// rows of the operand at ONE pixel, and consecutive pixels of an image row are consecutive vectors (K = 16 pixels per
// instruction = two 128-byte core matrices, LBO = 128).  What makes the operand expressible by ONE descriptor:
//  * A (M = 128): the five-pixel window of output column x - [X[x-2][0..C), ..., X[x+2][0..C)] = 5C contiguous halfs of the
//    zero-padded raw row, plus one block of "tap inside the image" flags (the constant-one channel of the whitening fold) - is
//    cut into 8-row blocks j = 0..nbx and stored as planes E[j][row][x] (one vector per pixel), i.e. the kx taps are expanded
//    ONCE per input row by a sliding 16-byte window copy (no per-tap replication by the MMA loop, no im2col in HBM).  The
//    block stride (SBO) is the plane stride, so rows 0..63 of the tile are the window of input row r.  Rows 64..127 are the
//    same window one input row further down, from a second set of planes (8 + j) that holds every row shifted by one: one
//    instruction covers the tap rows (ky, ky + 1), three instructions (ky = 0|1, 2|3, 4|-) cover the 5x5 filter.
//  * B (N = 48): dY rebuilt from d(pooled) and the arg-max side band as fp16 hi + lo pieces of both sibling networks,
//    planes [n / 8][x].
//  * D: three [128 x N] fp32 accumulators in TMEM, two sets: the MMAs of flush period i + 1 overlap the drain of period i.
//    The tensor-core accumulator truncates, so a period is g_wgrad_flush_steps K-steps; 16 epilogue warps add every period
//    into fp32 registers (fixed order -> deterministic) and write one partial per CTA; reduce + finalize kernels follow.
// Input rows live in a ring of kRing row slots, each row is expanded once and used by the five output rows around it.
// Roles: 16 epilogue warps | 8 fill warps | 1 MMA warp (one elected lane issues) | 1 producer warp (TMA bulk copies of the
// raw pixel row and of the d(pooled) / arg-max rows into a staging ring, kRing steps ahead).
#include <algorithm>
#include "conv_wgrad_tc.cuh"
#include "umma.cuh"

namespace cpp {
namespace wgtc {

using namespace umma;

constexpr int CO = kConvCout;
constexpr int kRing = 8;
constexpr int kEpiWarps = 16, kFillWarps = 8;
constexpr int kMmaWarp = kEpiWarps + kFillWarps, kProdWarp = kMmaWarp + 1;
constexpr int kThreads = 32 * (kProdWarp + 1);
constexpr int kHaloPx = 8;          // zero pixels on either side of a staged raw row (keeps the TMA destination 16-byte aligned)
constexpr int kMaxColsPerThread = 36;
enum { BAR_FULL = 0, BAR_FREE = kRing, BAR_STAGE = 2 * kRing, BAR_FULL_ACC = 3 * kRing, BAR_EMPTY_ACC = 3 * kRing + 2, BAR_COUNT = 3 * kRing + 4 };

struct Plan {
  const __half* x; const float* mean_inv;
  const float* g[kMaxNets]; const uint8_t* amax[kMaxNets]; const float* gmax[kMaxNets];
  float* dw[kMaxNets]; float* db[kMaxNets];
  float* partials; float* gsum;
  int B, H, W, C, PH, PW, nets, N, NB, nbx;
  int row_bytes, g_row_bytes, a_row_bytes, stage_g, stage_a, stage_bytes;
  int plane_stride, dy_slot_bytes;
  int flush_rows, grid, cols, part_floats;
  uint32_t off_dy, off_stage, off_bars, off_tmem, smem_bytes;
};

// power of two that brings max|g| just under 2^15 (1 when the tensor is all zero or not finite) - same rule as conv_wgrad_mma.cu
__device__ __forceinline__ float scale_for(float mx) {
  if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
  int e;
  frexpf(mx, &e);
  return ldexpf(1.f, 15 - e);
}

// 8 halfs from a 2-byte aligned shared-memory address: five aligned words, funnel-shifted by the misalignment
__device__ __forceinline__ uint4 load8h_any(const unsigned short* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t sh = ((uint32_t)a & 2u) << 3;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

// The CTA's output rows [g0, g1) of the batch (image-major) as segments that never cross an image.  A segment of output rows
// [ya, yb) is processed as steps pr = ya .. yb + 3: step pr stages padded input row pr (image row pr - 2) and, once pr >= ya + 4,
// output row y = pr - 4.  Every role walks the same sequence.
struct Seq {
  int g, g_end, H;
  int b, ya, yb;
  __device__ Seq(int g0, int g1, int H_) : g(g0), g_end(g1), H(H_), b(0), ya(0), yb(0) {}
  __device__ bool next() {
    if (g >= g_end) return false;
    b = g / H; ya = g - b * H;
    yb = min(H, ya + (g_end - g));
    g += yb - ya;
    return true;
  }
};

// Pipeline diagnosis build (nvcc -DWGTC_PROF, scripts/prof_wgrad_tc.py): per CTA, the cycles each role spends waiting on each
// hand-over barrier.  Slots: 0 kernel, 1 MMA loop, 2 MMA waits FULL, 3 MMA waits EMPTY_ACC, 4 fill loop, 5 fill waits FREE,
// 6 fill waits STAGE, 7 epilogue loop, 8 epilogue waits FULL_ACC, 9 producer waits FULL, 10 steps, 11 set-up
#ifdef WGTC_PROF
__device__ unsigned long long g_wprof[160][12];
#define WPROF_WAIT(acc, stmt) { const long long pf_a = clock64(); stmt; acc += (unsigned long long)(clock64() - pf_a); }
#define WPROF_PUT(slot, v) { if (lane == 0) g_wprof[blockIdx.x][slot] = (unsigned long long)(v); }
extern "C" __attribute__((visibility("default"))) int cpp_debug_wgrad_tc_prof(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_wprof, sizeof(unsigned long long) * 160 * 12);
}
#else
#define WPROF_WAIT(acc, stmt) { stmt; }
#define WPROF_PUT(slot, v)
#endif

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_tc_kernel(const __grid_constant__ Plan P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* E = smem;
  uint8_t* dyb = smem + P.off_dy;
  uint8_t* stage = smem + P.off_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.off_tmem);
  __shared__ float s_scale[kMaxNets];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int H = P.H, W = P.W, C = P.C, N = P.N;
  const long long G = (long long)P.B * H;
  const int g0 = (int)(G * blockIdx.x / gridDim.x), g1 = (int)(G * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(&bars[BAR_FULL + i], kFillWarps); mbar_init(&bars[BAR_FREE + i], 1); mbar_init(&bars[BAR_STAGE + i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[BAR_FULL_ACC + i], 1); mbar_init(&bars[BAR_EMPTY_ACC + i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  // the dY columns beyond nets * 20 and the pixel halos of the staged rows are never written again
  for (uint32_t i = tid; i < (P.off_bars - P.off_dy) / 16; i += kThreads) reinterpret_cast<uint4*>(dyb)[i] = make_uint4(0, 0, 0, 0);
  if (tid < kMaxNets) s_scale[tid] = tid < P.nets ? scale_for(P.gmax[tid][0]) : 1.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long pf_w0 = 0, pf_w1 = 0;
  const long long pf_start = clock64();
  (void)pf_w0; (void)pf_w1; (void)pf_start;

  if (warp < kEpiWarps) {
    // =========================================================================== epilogue: TMEM -> fp32 register partial sums
    const int quarter = warp & 3, grp = warp >> 2;
    const int CG = P.cols / 4;
    float acc[kMaxColsPerThread];
#pragma unroll
    for (int c = 0; c < kMaxColsPerThread; ++c) acc[c] = 0.f;
    const int rows = g1 - g0, nper = (rows + P.flush_rows - 1) / P.flush_rows;
    for (int p = 0; p < nper; ++p) {
      const uint32_t set = p & 1;
      WPROF_WAIT(pf_w0, mbar_wait_sleep(&bars[BAR_FULL_ACC + set], (p >> 1) & 1, 256));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + set * (uint32_t)P.cols + (uint32_t)(grp * CG);
#pragma unroll
      for (int c0 = 0; c0 < kMaxColsPerThread; c0 += 12) {
        uint32_t r[12];
#pragma unroll
        for (int c = 0; c < 12; c += 4)
          if (c0 + c < CG) tmem_ld4(taddr + c0 + c, r + c);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 12; ++c)
          if (c0 + c < CG) acc[c0 + c] += __uint_as_float(r[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_EMPTY_ACC + set]);
    }
    if (warp == 0) { WPROF_PUT(7, clock64() - pf_start); WPROF_PUT(8, pf_w0); }
    float* part = P.partials + (size_t)blockIdx.x * P.part_floats;
#pragma unroll
    for (int c = 0; c < kMaxColsPerThread; ++c)
      if (c < CG) part[(size_t)(grp * CG + c) * 128 + 32 * quarter + lane] = acc[c];
  } else if (warp < kEpiWarps + kFillWarps) {
    // =========================================================================== fill: staged rows -> window planes + dY pieces
    const int ftid = tid - 32 * kEpiWarps, nfill = 32 * kFillWarps;
    const int nX = (P.nbx + 1) * W, nDY = P.PW * P.nets;
    const uint32_t PS = (uint32_t)P.plane_stride;
    const int j_first = ftid / W, x_first = ftid - j_first * W, dj = nfill / W, dx = nfill - dj * W;   // item -> (block j, column x) without a division per item
    int n = 0;
    for (Seq sq(g0, g1, H); sq.next();) {
      for (int pr = sq.ya; pr < sq.yb + 4; ++pr, ++n) {
        const uint32_t slot = (uint32_t)n % kRing, k = (uint32_t)n / kRing;
        WPROF_WAIT(pf_w0, mbar_wait_sleep(&bars[BAR_FREE + slot], (k & 1) ^ 1, 32));   // the MMAs that read this slot one ring turn ago are done
        WPROF_WAIT(pf_w1, mbar_wait(&bars[BAR_STAGE + slot], k & 1));                  // the producer's copies for this step have landed
        const int iy = pr - 2, y = pr - 4;
        const bool has_x = iy >= 0 && iy < H, has_dy = y >= sq.ya, dy_data = has_dy && (y >> 1) < P.PH;
        const uint8_t* st = stage + (size_t)slot * P.stage_bytes;
        const uint32_t prev = (slot + kRing - 1) % kRing;
        int j = j_first, x = x_first;
        for (int item = ftid; item < nX + (has_dy ? nDY : 0); item += nfill, j += dj, x += dx) {
          if (x >= W) { x -= W; ++j; }
          if (item < nX) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (has_x) {
              if (j < P.nbx) {
                v = load8h_any(reinterpret_cast<const unsigned short*>(st) + (x + kHaloPx - 2) * C + 8 * j);
              } else {                                            // flags: tap kx of output column x reads a pixel inside the image
                uint32_t f[5];
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) f[kx] = (x - 2 + kx >= 0 && x - 2 + kx < W) ? 0x3C00u : 0u;
                v = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4], 0u);
              }
            }
            *reinterpret_cast<uint4*>(E + (size_t)j * PS + (size_t)(slot * W + x) * 16) = v;
            if (n > 0) *reinterpret_cast<uint4*>(E + (size_t)(8 + j) * PS + (size_t)(prev * W + x) * 16) = v;   // "one row down" copy
          } else {
            const int idx = item - nX, net = idx / P.PW, px = idx - net * P.PW;
            const float sc = s_scale[net];
            const float2* gp = reinterpret_cast<const float2*>(st + P.stage_g + net * P.g_row_bytes + px * (CO * 4));
            const unsigned short* ap = reinterpret_cast<const unsigned short*>(st + P.stage_a + net * P.a_row_bytes + px * CO);
            uint32_t hi2[5], lo2[5], a0[5], a1[5];
#pragma unroll
            for (int v = 0; v < 5; ++v) {
              const float2 gq = dy_data ? gp[v] : make_float2(0.f, 0.f);
              const uint32_t aq = dy_data ? (uint32_t)ap[v] : 0x0404u;
              const float q0 = gq.x * sc, q1 = gq.y * sc;
              const __half h0 = __float2half_rn(q0), h1 = __float2half_rn(q1);
              const __half l0 = __float2half_rn(q0 - __half2float(h0)), l1 = __float2half_rn(q1 - __half2float(h1));
              hi2[v] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lo2[v] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
              a0[v] = aq & 0xffu; a1[v] = aq >> 8;
            }
            uint8_t* drow = dyb + (size_t)slot * P.dy_slot_bytes;
#pragma unroll
            for (int xp = 0; xp < 2; ++xp) {
              const uint32_t pa = (uint32_t)((y & 1) * 2 + xp);
              uint32_t w[10];
#pragma unroll
              for (int v = 0; v < 5; ++v) {
                const uint32_t m = (a0[v] == pa ? 0x0000ffffu : 0u) | (a1[v] == pa ? 0xffff0000u : 0u);
                w[v] = hi2[v] & m; w[5 + v] = lo2[v] & m;
              }
              // words net * 10 .. + 9 along N (column n = 2 * word): 16-byte stores where a block is complete
              uint8_t* dpx = drow + (size_t)((net * 10) >> 2) * (W * 16) + (size_t)(2 * px + xp) * 16;
              const size_t bs = (size_t)W * 16;
              if ((net & 1) == 0) {
                *reinterpret_cast<uint4*>(dpx) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dpx + bs) = make_uint4(w[4], w[5], w[6], w[7]);
                *reinterpret_cast<uint2*>(dpx + 2 * bs) = make_uint2(w[8], w[9]);
              } else {
                *reinterpret_cast<uint2*>(dpx + 8) = make_uint2(w[0], w[1]);
                *reinterpret_cast<uint4*>(dpx + bs) = make_uint4(w[2], w[3], w[4], w[5]);
                *reinterpret_cast<uint4*>(dpx + 2 * bs) = make_uint4(w[6], w[7], w[8], w[9]);
              }
            }
          }
        }
        fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_FULL + slot]);
      }
    }
    if (warp == kEpiWarps) { WPROF_PUT(4, clock64() - pf_start); WPROF_PUT(5, pf_w0); WPROF_PUT(6, pf_w1); WPROF_PUT(10, n); }
  } else if (warp == kMmaWarp) {
    // =========================================================================== MMA issue
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);   // fp32 D, fp16 A/B, both MN-major
    const uint32_t e_u = smem_u32(E), d_u = smem_u32(dyb);
    int n = 0, rows_in_period = 0;
    uint32_t period = 0;
    for (Seq sq(g0, g1, H); sq.next();) {
      for (int pr = sq.ya; pr < sq.yb + 4; ++pr, ++n) {
        const uint32_t slot = (uint32_t)n % kRing, k = (uint32_t)n / kRing;
        WPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_FULL + slot], k & 1));
        tc_fence_after();
        if (pr - 4 >= sq.ya) {
          const uint32_t set = period & 1;
          if (rows_in_period == 0) { WPROF_WAIT(pf_w1, mbar_wait(&bars[BAR_EMPTY_ACC + set], ((period >> 1) & 1) ^ 1)); tc_fence_after(); }
          for (int x0 = 0; x0 < W; x0 += 16) {
            const uint64_t bdesc = make_desc(d_u + slot * (uint32_t)P.dy_slot_bytes + (uint32_t)x0 * 16, 128, (uint32_t)W * 16);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const uint32_t sa = (uint32_t)(n - 4 + 2 * a) % kRing;     // input row y + 2a (and y + 2a + 1 from the shifted planes)
              const uint64_t adesc = make_desc(e_u + (sa * (uint32_t)W + (uint32_t)x0) * 16, 128, (uint32_t)P.plane_stride);
              if (elect_one()) umma_f16(tmem_base + set * (uint32_t)P.cols + (uint32_t)(a * N), adesc, bdesc, idesc, (rows_in_period > 0 || x0 > 0) ? 1u : 0u);
            }
          }
          if (elect_one()) umma_commit(&bars[BAR_FREE + (uint32_t)(n - 4) % kRing]);      // input row y is dead
          if (++rows_in_period == P.flush_rows) {
            if (elect_one()) umma_commit(&bars[BAR_FULL_ACC + set]);
            ++period; rows_in_period = 0;
          }
        }
        if (pr == sq.yb + 3 && elect_one())                        // the segment's last four input rows have no output row of their own
          for (int t = 3; t >= 0; --t) umma_commit(&bars[BAR_FREE + (uint32_t)(n - t) % kRing]);
        __syncwarp();
      }
    }
    if (rows_in_period > 0 && elect_one()) umma_commit(&bars[BAR_FULL_ACC + (period & 1)]);
    __syncwarp();
    WPROF_PUT(1, clock64() - pf_start); WPROF_PUT(2, pf_w0); WPROF_PUT(3, pf_w1);
  } else if (lane == 0) {
    // =========================================================================== producer: TMA bulk copies, kRing steps ahead
    int n = 0;
    for (Seq sq(g0, g1, H); sq.next();) {
      for (int pr = sq.ya; pr < sq.yb + 4; ++pr, ++n) {
        const uint32_t slot = (uint32_t)n % kRing, k = (uint32_t)n / kRing;
        if (n >= kRing) WPROF_WAIT(pf_w0, mbar_wait_sleep(&bars[BAR_FULL + slot], (k - 1) & 1, 128));   // the fill warps are done with this staging slot
        const int iy = pr - 2, y = pr - 4;
        const bool has_x = iy >= 0 && iy < H, dy_data = y >= sq.ya && (y >> 1) < P.PH;
        uint8_t* st = stage + (size_t)slot * P.stage_bytes;
        const uint32_t bytes = (has_x ? (uint32_t)P.row_bytes : 0u) + (dy_data ? (uint32_t)(P.nets * (P.g_row_bytes + P.a_row_bytes)) : 0u);
        if (bytes == 0) { mbar_arrive(&bars[BAR_STAGE + slot]); continue; }
        fence_proxy_async();
        mbar_expect_tx(&bars[BAR_STAGE + slot], bytes);
        if (has_x) bulk_g2s(st + kHaloPx * C * 2, P.x + ((size_t)sq.b * H + iy) * W * C, (uint32_t)P.row_bytes, &bars[BAR_STAGE + slot]);
        if (dy_data) {
          const size_t q = ((size_t)sq.b * P.PH + (y >> 1)) * P.PW * CO;
          for (int net = 0; net < P.nets; ++net) {
            bulk_g2s(st + P.stage_g + net * P.g_row_bytes, P.g[net] + q, (uint32_t)P.g_row_bytes, &bars[BAR_STAGE + slot]);
            bulk_g2s(st + P.stage_a + net * P.a_row_bytes, P.amax[net] + q, (uint32_t)P.a_row_bytes, &bars[BAR_STAGE + slot]);
          }
        }
      }
    }
  }
#ifdef WGTC_PROF
  if (warp == kProdWarp) WPROF_PUT(9, pf_w0);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
#ifdef WGTC_PROF
  if (tid == 0) g_wprof[blockIdx.x][0] = (unsigned long long)(clock64() - pf_start);
#endif
}

// gsum[i] = sum over CTAs of partials[cta][i] in fixed order; block (32, 8)
__global__ void __launch_bounds__(256) wgtc_reduce_kernel(const float* __restrict__ partials, int nparts, int part_floats, float* __restrict__ gsum) {
  __shared__ float sh[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x, y = threadIdx.y;
  const int per = (nparts + 7) / 8, k0 = y * per, k1 = min(nparts, k0 + per);
  float s = 0.f;
  if (i < part_floats) for (int k = k0; k < k1; ++k) s += partials[(size_t)k * part_floats + i];
  sh[y][threadIdx.x] = s;
  __syncthreads();
  if (y == 0 && i < part_floats) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sh[r][threadIdx.x];
    gsum[i] = t;
  }
}

// G of tap row ky, window row `wrow` (kx * C + c, or 8 * nbx + kx for the flag block), column n
__device__ __forceinline__ float g_at(const Plan& P, int ky, int wrow, int n) {
  return P.gsum[(size_t)((ky >> 1) * P.N + n) * 128 + 64 * (ky & 1) + wrow];
}
__global__ void __launch_bounds__(256) wgtc_finalize_kernel(const __grid_constant__ Plan P) {
  const int C = P.C, nw = 25 * C * CO;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nets * (nw + CO)) return;
  const int net = i / (nw + CO), j = i - net * (nw + CO);
  const float inv_scale = 1.f / scale_for(P.gmax[net][0]);
  if (j >= nw) {                                                          // bias gradient: flag block, centre tap
    const int o = j - nw;
    const float s = g_at(P, 2, 8 * P.nbx + 2, net * 2 * CO + o) + g_at(P, 2, 8 * P.nbx + 2, (net * 2 + 1) * CO + o);
    P.db[net][o] = s * inv_scale;
    return;
  }
  const int o = j % CO, c = (j / CO) % C, kx = (j / (CO * C)) % 5, ky = j / (CO * C * 5);
  const int n0 = net * 2 * CO + o, n1 = n0 + CO;
  const float gsum = g_at(P, ky, kx * C + c, n0) + g_at(P, ky, kx * C + c, n1);
  float v = gsum * inv_scale;
  if (P.mean_inv) {
    const float ssum = (g_at(P, ky, 8 * P.nbx + kx, n0) + g_at(P, ky, 8 * P.nbx + kx, n1)) * inv_scale;
    v = P.mean_inv[C + c] * (v - P.mean_inv[c] * ssum);
  }
  P.dw[net][j] = v;
}

// ------------------------------------------------------------------------------------------ host
static inline size_t al256(size_t b) { return (size_t)round_up((int64_t)b, 256); }

static bool build_plan(int nets, int B, int H, int W, int C, int KS, Plan* P) {
  if (KS != 5 || nets < 1 || nets > kMaxNets || H < 2 || W < 16 || W > 64 || (W % 16) != 0 || C < 1) return false;
  if (((W * C * 2) % 16) != 0 || ((kHaloPx * C * 2) % 16) != 0) return false;
  const int nbx = (5 * C + 7) / 8;
  if (nbx + 1 > 8) return false;
  P->B = B; P->H = H; P->W = W; P->C = C; P->PH = H / 2; P->PW = W / 2; P->nets = nets;
  P->N = (int)round_up(nets * 2 * CO, 16); P->NB = P->N / 8; P->nbx = nbx;
  P->cols = 3 * P->N;
  if (P->cols / 4 > kMaxColsPerThread || (P->cols % 16) != 0) return false;
  P->row_bytes = W * C * 2;
  P->g_row_bytes = P->PW * CO * 4; P->a_row_bytes = P->PW * CO;
  const int xbuf = (int)round_up((W + 2 * kHaloPx) * C * 2 + 4, 16);          // + 4: load8h_any reads one word past its window
  P->stage_g = xbuf; P->stage_a = xbuf + nets * P->g_row_bytes;
  P->stage_bytes = (int)round_up(P->stage_a + nets * P->a_row_bytes, 16);
  P->plane_stride = kRing * W * 16;
  P->dy_slot_bytes = P->NB * W * 16;
  P->off_dy = 16u * (uint32_t)P->plane_stride;
  P->off_stage = P->off_dy + (uint32_t)(kRing * P->dy_slot_bytes);
  P->off_bars = P->off_stage + (uint32_t)(kRing * P->stage_bytes);
  P->off_tmem = P->off_bars + BAR_COUNT * 8;
  P->smem_bytes = P->off_tmem + 16;
  if (P->smem_bytes > 225 * 1024) return false;
  P->flush_rows = std::max(1, g_wgrad_flush_steps / (W / 16));
  P->grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)B * H, sm_budget()));
  P->part_floats = P->cols * 128;
  return true;
}

bool supported(int nets, int H, int W, int C, int KS) {
  Plan P{};
  return build_plan(nets, 1, H, W, C, KS, &P);
}

int64_t scratch_bytes(int nets, int H, int W, int C, int KS) {
  Plan P{};
  if (!build_plan(nets, 1, H, W, C, KS, &P)) return 0;
  return (int64_t)(al256((size_t)kNumSMs * P.part_floats * 4) + al256((size_t)P.part_floats * 4));
}

int launch(const void* x_f16, const float* mean_inv, int nets, const float* const* d_pooled, const uint8_t* const* amax, int B, int H,
           int W, int C, int KS, float* const* dw, float* const* db, const float* const* gmax, void* scratch, cudaStream_t s) {
  if (B <= 0) return CPP_OK;
  Plan P{};
  CPP_REQUIRE(build_plan(nets, B, H, W, C, KS, &P), "wgrad_tc: unsupported layer %dx%dx%d k%d, %d networks", H, W, C, KS, nets);
  CPP_REQUIRE(((uintptr_t)x_f16 & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "wgrad_tc: unaligned input or scratch");
  P.x = reinterpret_cast<const __half*>(x_f16); P.mean_inv = mean_inv;
  for (int n = 0; n < nets; ++n) {
    CPP_REQUIRE(d_pooled[n] && amax[n] && dw[n] && db[n] && gmax[n], "wgrad_tc: null pointer for network %d", n);
    CPP_REQUIRE(((uintptr_t)d_pooled[n] & 15) == 0 && ((uintptr_t)amax[n] & 15) == 0, "wgrad_tc: unaligned gradient / arg-max of network %d", n);
    P.g[n] = d_pooled[n]; P.amax[n] = amax[n]; P.dw[n] = dw[n]; P.db[n] = db[n]; P.gmax[n] = gmax[n];
  }
  char* sc = reinterpret_cast<char*>(scratch);
  P.partials = reinterpret_cast<float*>(sc); sc += al256((size_t)kNumSMs * P.part_floats * 4);
  P.gsum = reinterpret_cast<float*>(sc);
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    configured = true;
  }
  conv_wgrad_tc_kernel<<<P.grid, kThreads, P.smem_bytes, s>>>(P);
  CPP_CHECK_LAUNCH();
  wgtc_reduce_kernel<<<(unsigned)ceil_div(P.part_floats, 32), dim3(32, 8), 0, s>>>(P.partials, P.grid, P.part_floats, P.gsum);
  CPP_CHECK_LAUNCH();
  const int total = nets * (25 * C * CO + CO);
  wgtc_finalize_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace wgtc
}  // namespace cpp
