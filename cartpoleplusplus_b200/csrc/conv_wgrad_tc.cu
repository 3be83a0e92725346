// conv1 weight / bias gradient on tcgen05 (base_network.py:103-107 backwards; see conv_wgrad_mma.cu for the algebra):
//
//   G[(ky,kx,c), (n,piece,o)] = sum over output pixels p of  X[p + (ky,kx)][c] * dY_n[p][o]
//
// The reduction runs over PIXELS, so both UMMA operands are MN-major: a 16-byte shared-memory vector holds 8 consecutive
// rows of the operand at ONE pixel, consecutive pixels of an image row are consecutive vectors (K = 16 pixels per
// instruction = two 128-byte core matrices, LBO = 128), and the 8-row blocks of an operand sit at ONE uniform stride (SBO).
// Both filter-tap directions are turned into that form without replicating data per tap:
//  * kx -> rows of A.  The five-pixel window of column x, [X[x-2][0..C), ..., X[x+2][0..C)] = 5C contiguous halfs of the
//    zero-padded raw row, plus one block of "tap inside the image" flags (the constant-one channel of the whitening fold), is
//    cut into 8-row blocks j and stored ONCE per input row as planes (a sliding 16-byte window copy of the staged row).
//    A tile = an input row PAIR (r, r + 1), r even: block 2j + parity, SBO = one plane row, M = 128.
//  * ky -> columns of B.  With the input row fixed, tap ky meets output row y = r + 2 - ky, so the B operand of ONE
//    instruction is the SIX consecutive dY rows r - 2 .. r + 3 (each: both networks x hi/lo fp16 pieces x 10 filters = 40
//    columns), N = 240: column block t of the even input row is tap 4 - t, of the odd input row tap 5 - t.  dY rows live in a
//    ring of row pairs laid out [row][column block][x], so the six rows are one descriptor (two when the window wraps).
//  One instruction per (input row pair, 16 pixels) replaces the 2 x 3 x 5/2 instructions of a per-tap formulation and reads
//  4 KB (A) + 7.5 KB (B) of shared memory for 2 x 25 taps.
//  * D: [128 x 240] fp32 in TMEM, two sets: the MMAs of flush period i + 1 overlap the drain of period i.  The tensor-core
//    accumulator truncates, so a period is g_wgrad_flush_steps K-steps per element; 24 warps add every period into
//    fp32 registers (fixed order -> deterministic) and write one partial per CTA; reduce + finalize kernels follow.
// Roles: 24 general warps (the six fill items of a step rotate over them; every warp also drains its TMEM columns) | 1 MMA warp
// (one elected lane issues) | 1 producer warp (one TMA bulk copy of the raw pixel row pair per step into a staging ring, kStage steps
// ahead; the d(pooled) / arg-max rows are read straight from L2 by the dY items).
#include <algorithm>
#include "conv_wgrad_tc.cuh"
#include "umma.cuh"

namespace cpp {
namespace wgtc {

using namespace umma;

constexpr int CO = kConvCout;
constexpr int kESlots = 4;          // ring of expanded input row pairs
constexpr int kDSlots = 8;          // ring of dY row pairs (three consecutive ones per instruction)
constexpr int kStage = 8;           // staging ring (raw rows, pooled gradients, arg-max bytes), one slot per step
constexpr int kGenWarps = 24;        // fill + epilogue warps: 6 groups of 4 (one warp per TMEM lane quarter), group t drains dY row t
constexpr int kItems = 6;           // fill work items of a step (one warp each): 4 x (input row, column parity) + 2 x dY
constexpr int kMmaWarp = kGenWarps, kProdWarp = kMmaWarp + 1;
constexpr int kThreads = 32 * (kProdWarp + 1);
constexpr int kFrontSlack = 48, kBackSlack = 64;   // bytes around the two staged raw rows: windows of the border columns start / end outside them
constexpr int kMaxCB = 40;          // columns of one dY row: nets * 2 pieces * 10 filters rounded up to 8
enum { BAR_FULL = 0, BAR_FREE = BAR_FULL + kESlots, BAR_STAGE_FULL = BAR_FREE + kESlots, BAR_STAGE_FREE = BAR_STAGE_FULL + kStage,
       BAR_FULL_ACC = BAR_STAGE_FREE + kStage, BAR_EMPTY_ACC = BAR_FULL_ACC + 2, BAR_COUNT = BAR_EMPTY_ACC + 2 };

struct Plan {
  const __half* x; const float* mean_inv;
  const float* g[kMaxNets]; const uint8_t* amax[kMaxNets]; const float* gmax[kMaxNets];
  float* dw[kMaxNets]; float* db[kMaxNets];
  float* partials; float* gsum;
  int B, H, W, C, HP, PH, PW, nets, CB, NBR, N, nbx;
  int row_bytes, stage_bytes;
  int e_slot_bytes, d_slot_bytes;
  int flush_steps, grid, part_floats;
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

// The CTA's input row pairs [g0, g1) of the batch (image-major) as segments that never cross an image.  A segment of pairs
// [pa, pb) is processed as steps t = 0 .. pb - pa + 1: step t stages dY row pair q = pa - 1 + t (zero outside the image) and,
// once t >= 2, input row pair p = pa + t - 2, whose instruction reads the three dY pairs staged last.  Every role walks the
// same sequence.
struct Seq {
  int g, g_end, HP;
  int b, pa, pb;
  __device__ Seq(int g0, int g1, int HP_) : g(g0), g_end(g1), HP(HP_), b(0), pa(0), pb(0) {}
  __device__ bool next() {
    if (g >= g_end) return false;
    b = g / HP; pa = g - b * HP;
    pb = min(HP, pa + (g_end - g));
    g += pb - pa;
    return true;
  }
};

// Pipeline diagnosis build (nvcc -DWGTC_PROF, scripts/prof_wgrad_tc.py): per CTA, the cycles each role spends waiting on each
// hand-over barrier.  Slots: 0 kernel, 1 MMA loop, 2 MMA waits FULL, 3 MMA waits EMPTY_ACC, 4 fill loop, 5 fill waits FREE,
// 6 fill waits STAGE, 7 epilogue loop, 8 epilogue waits FULL_ACC, 9 producer waits STAGE_FREE, 10 steps
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

__device__ __forceinline__ uint32_t idesc_mn(int N) {      // fp32 D, fp16 A/B, both MN-major, M = 128
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

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
  const int H = P.H, W = P.W, C = P.C, CB = P.CB;
  const long long G = (long long)P.B * P.HP;
  const int g0 = (int)(G * blockIdx.x / gridDim.x), g1 = (int)(G * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < kESlots; ++i) { mbar_init(&bars[BAR_FULL + i], kItems); mbar_init(&bars[BAR_FREE + i], 1); }
    for (int i = 0; i < kStage; ++i) { mbar_init(&bars[BAR_STAGE_FULL + i], 1); mbar_init(&bars[BAR_STAGE_FREE + i], kItems); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[BAR_FULL_ACC + i], 1); mbar_init(&bars[BAR_EMPTY_ACC + i], kGenWarps); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  // dY columns beyond nets * 20 are never written again
  for (uint32_t i = tid; i < (P.off_bars - P.off_dy) / 16; i += kThreads) reinterpret_cast<uint4*>(dyb)[i] = make_uint4(0, 0, 0, 0);
  if (tid < kMaxNets) s_scale[tid] = tid < P.nets ? scale_for(P.gmax[tid][0]) : 1.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long pf_w0 = 0, pf_w1 = 0, pf_w2 = 0, pf_w3 = 0;
  const long long pf_start = clock64();
  (void)pf_w0; (void)pf_w1; (void)pf_w2; (void)pf_w3; (void)pf_start;

  if (warp < kGenWarps) {
    // =========================================================================== general warps: fill items + epilogue
    // The six fill items of step n go to warps (6 n + i) mod 24, so four consecutive steps are staged concurrently by
    // disjoint warps (the latency of one item - shared-memory loads behind the tensor core's operand fetches - is hidden
    // behind three other steps).  Between items every warp polls for a finished flush period and drains its 40 columns.
    const int quarter = warp & 3, grp = warp >> 2;                 // group t drains the columns of dY row t of the window
    float acc[kMaxCB];
#pragma unroll
    for (int c = 0; c < kMaxCB; ++c) acc[c] = 0.f;
    const int pairs = g1 - g0, nper = (pairs + P.flush_steps - 1) / P.flush_steps;
    int pd = 0;                                                    // next flush period to drain
    auto drain = [&](bool blocking) {
      while (pd < nper) {
        const uint32_t set = pd & 1, par = (pd >> 1) & 1;
        if (blocking) { WPROF_WAIT(pf_w0, mbar_wait_park(&bars[BAR_FULL_ACC + set], par)); }
        else if (!mbar_test(&bars[BAR_FULL_ACC + set], par)) break;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + set * (uint32_t)P.N + (uint32_t)(grp * CB);
#pragma unroll
        for (int c0 = 0; c0 < kMaxCB; c0 += 8) {
          uint32_t r[8];
          if (c0 < CB) { tmem_ld4(taddr + c0, r); tmem_ld4(taddr + c0 + 4, r + 4); }
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c0 < CB) acc[c0 + c] += __uint_as_float(r[c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_EMPTY_ACC + set]);
        ++pd;
      }
    };
    const uint32_t bsW = (uint32_t)W * 16;
    int n = 0;
    for (Seq sq(g0, g1, P.HP); sq.next();) {
      const int nsteps = sq.pb - sq.pa + 2;
      for (int t = 0; t < nsteps; ++t, ++n) {
        drain(false);
        const int item = (warp + kGenWarps - (kItems * n) % kGenWarps) % kGenWarps;
        if (item >= kItems) continue;
        const uint32_t es = (uint32_t)n % kESlots, ds = (uint32_t)n % kDSlots, ss = (uint32_t)n % kStage;
        WPROF_WAIT(pf_w1, mbar_wait_park(&bars[BAR_FREE + es], (((uint32_t)n / kESlots) & 1) ^ 1));   // the MMAs that read this E slot one ring turn ago are done
        WPROF_WAIT(pf_w2, mbar_wait_park(&bars[BAR_STAGE_FULL + ss], ((uint32_t)n / kStage) & 1));    // the producer's copies for this step have landed
        const int q = sq.pa - 1 + t, p = sq.pa + t - 2;
        const bool has_x = t >= 2, dy_data = q >= 0 && q < P.PH;
        const uint8_t* st = stage + (size_t)ss * P.stage_bytes;
#ifdef WGTC_PROF
        const long long pf_i0 = clock64();
#endif
        if (item < 4) {
          // input row `rpar` of the pair, columns of parity `xpar`, one lane per column (for an odd channel count the words of
          // 32 windows two pixels apart sit in 32 distinct banks); ALL window blocks of the column from one run of loads.
          // The pixels of a row are stored even columns first, then odd ones (the reduction index of the MMA may be permuted
          // as long as A and B agree): the lanes of a warp write consecutive vectors.
          const int rpar = item >> 1, xpar = item & 1;
          const int xcol = 2 * lane + xpar, xpos = lane + (W / 2) * xpar;
          if (has_x && xcol < W) {
            uint8_t* dst = E + (size_t)es * P.e_slot_bytes + (size_t)(rpar * W + xpos) * 16;        // block j at + 2 j W 16
            if (2 * p + rpar < H) {
              const uintptr_t a = reinterpret_cast<uintptr_t>(st + kFrontSlack + rpar * P.row_bytes) + (uintptr_t)((xcol - 2) * C * 2);
              const uint32_t sh = ((uint32_t)a & 2u) << 3;
              const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
              // SAME padding: window halfs [lo, hi) come from pixels inside the row, the rest (border columns only) are zero
              const int lo = max(0, 2 - xcol) * C, hi = min(5, W + 2 - xcol) * C;
              const bool border = lo > 0 || hi < 5 * C;
#pragma unroll
              for (int j0 = 0; j0 < 8; j0 += 2) {                 // two blocks per run: 9 words in flight, then 2 stores
                if (j0 < P.nbx) {
                  uint32_t w[9];
#pragma unroll
                  for (int i = 0; i < 9; ++i) w[i] = (4 * j0 + i <= 4 * P.nbx) ? wp[4 * j0 + i] : 0u;
#pragma unroll
                  for (int jj = 0; jj < 2; ++jj)
                    if (j0 + jj < P.nbx) {
                      uint32_t o[4];
#pragma unroll
                      for (int k = 0; k < 4; ++k) o[k] = __funnelshift_r(w[4 * jj + k], w[4 * jj + k + 1], sh);
                      if (border) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                          const int h0 = 8 * (j0 + jj) + 2 * k;
                          o[k] &= ((h0 >= lo && h0 < hi) ? 0x0000ffffu : 0u) | ((h0 + 1 >= lo && h0 + 1 < hi) ? 0xffff0000u : 0u);
                        }
                      }
                      *reinterpret_cast<uint4*>(dst + (size_t)(2 * (j0 + jj)) * bsW) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
              }
              uint32_t f[5];                                        // flags: tap kx of output column x reads a pixel inside the image
#pragma unroll
              for (int kx = 0; kx < 5; ++kx) f[kx] = (xcol - 2 + kx >= 0 && xcol - 2 + kx < W) ? 0x3C00u : 0u;
              *reinterpret_cast<uint4*>(dst + (size_t)(2 * P.nbx) * bsW) = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4], 0u);
            } else {
              for (int j = 0; j <= P.nbx; ++j) *reinterpret_cast<uint4*>(dst + (size_t)(2 * j) * bsW) = make_uint4(0, 0, 0, 0);
            }
          }
        } else {
          uint8_t* dslot = dyb + (size_t)ds * P.d_slot_bytes;
          for (int idx = (item - 4) * 32 + lane; idx < P.PW * P.nets; idx += 64) {
            const int net = idx / P.PW, px = idx - net * P.PW;
            const float sc = s_scale[net];
            // pooled gradients and arg-max bytes straight from global memory (L2): four steps are in flight on different warps
            const size_t qo = (((size_t)sq.b * P.PH + (dy_data ? q : 0)) * P.PW + px) * CO;
            const float2* gp = reinterpret_cast<const float2*>(P.g[net] + qo);
            const unsigned short* ap = reinterpret_cast<const unsigned short*>(P.amax[net] + qo);
            uint32_t hi2[5], lo2[5], aq[5];
#pragma unroll
            for (int v = 0; v < 5; ++v) {
              const float2 gq = dy_data ? __ldg(gp + v) : make_float2(0.f, 0.f);
              aq[v] = dy_data ? (uint32_t)__ldg(ap + v) : 0x0404u;
              const float q0 = gq.x * sc, q1 = gq.y * sc;
              const __half2 h = __floats2half2_rn(q0, q1);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(q0 - hf.x, q1 - hf.y);
              hi2[v] = *reinterpret_cast<const uint32_t*>(&h);
              lo2[v] = *reinterpret_cast<const uint32_t*>(&l);
            }
#pragma unroll
            for (int pa4 = 0; pa4 < 4; ++pa4) {                   // the four pixels of the 2x2 window
              // word i of this pixel's 10 (5 hi + 5 lo half2 pairs), masked by "the arg-max of the window is this pixel"
              auto word = [&](int i) -> uint32_t {
                const int v = i < 5 ? i : i - 5;
                const uint32_t eq = __vcmpeq4(aq[v], (uint32_t)pa4 * 0x0101u);      // 0xff in the bytes whose arg-max is this pixel
                return (i < 5 ? hi2[v] : lo2[v]) & __byte_perm(eq, 0u, 0x1100);     // byte 0 -> low half, byte 1 -> high half
              };
              // words net * 10 .. + 9 of the row's columns (column = 2 * word): 16-byte stores where a block is complete
              uint8_t* dpx = dslot + (size_t)((pa4 >> 1) * P.NBR + ((net * 10) >> 2)) * bsW + (size_t)(px + (W / 2) * (pa4 & 1)) * 16;
              if ((net & 1) == 0) {
                *reinterpret_cast<uint4*>(dpx) = make_uint4(word(0), word(1), word(2), word(3));
                *reinterpret_cast<uint4*>(dpx + bsW) = make_uint4(word(4), word(5), word(6), word(7));
                *reinterpret_cast<uint2*>(dpx + 2 * bsW) = make_uint2(word(8), word(9));
              } else {
                *reinterpret_cast<uint2*>(dpx + 8) = make_uint2(word(0), word(1));
                *reinterpret_cast<uint4*>(dpx + bsW) = make_uint4(word(2), word(3), word(4), word(5));
                *reinterpret_cast<uint4*>(dpx + 2 * bsW) = make_uint4(word(6), word(7), word(8), word(9));
              }
            }
          }
        }
        fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bars[BAR_STAGE_FREE + ss]); mbar_arrive(&bars[BAR_FULL + es]); }
#ifdef WGTC_PROF
        pf_w3 += (unsigned long long)(clock64() - pf_i0);
#endif
      }
    }
    drain(true);
    if (warp == 0) { WPROF_PUT(4, clock64() - pf_start); WPROF_PUT(5, pf_w1); WPROF_PUT(6, pf_w2); WPROF_PUT(7, pf_w3); }
    if (warp == 4) { WPROF_PUT(8, pf_w1); WPROF_PUT(11, pf_w2); WPROF_PUT(10, pf_w3); }
    float* part = P.partials + (size_t)blockIdx.x * P.part_floats;
#pragma unroll
    for (int c = 0; c < kMaxCB; ++c)
      if (c < CB) part[(size_t)(grp * CB + c) * 128 + 32 * quarter + lane] = acc[c];
  } else if (warp == kMmaWarp) {
    // =========================================================================== MMA issue
    const uint32_t e_u = smem_u32(E), d_u = smem_u32(dyb);
    const uint32_t sbo = (uint32_t)W * 16;
    int n = 0, steps_in_period = 0;
    uint32_t period = 0;
    for (Seq sq(g0, g1, P.HP); sq.next();) {
      const int nsteps = sq.pb - sq.pa + 2;
      for (int t = 0; t < nsteps; ++t, ++n) {
        const uint32_t es = (uint32_t)n % kESlots;
        WPROF_WAIT(pf_w0, mbar_wait_park(&bars[BAR_FULL + es], ((uint32_t)n / kESlots) & 1));
        tc_fence_after();
        if (t >= 2) {
          const uint32_t set = period & 1;
          if (steps_in_period == 0) { WPROF_WAIT(pf_w1, mbar_wait_park(&bars[BAR_EMPTY_ACC + set], ((period >> 1) & 1) ^ 1)); tc_fence_after(); }
          // the three dY pairs staged last: ring slots (n - 2, n - 1, n) mod kDSlots, contiguous unless the window wraps
          const uint32_t s0 = (uint32_t)(n - 2) % kDSlots;
          const uint32_t c1 = min(3u, (uint32_t)kDSlots - s0);
          const int N1 = (int)c1 * 2 * CB, N2 = (3 - (int)c1) * 2 * CB;
          const uint32_t d_tmem = tmem_base + set * (uint32_t)P.N;
          for (int x0 = 0; x0 < W; x0 += 16) {
            const uint64_t adesc = make_desc(e_u + es * (uint32_t)P.e_slot_bytes + (uint32_t)x0 * 16, 128, sbo);
            const uint32_t accum = (steps_in_period > 0 || x0 > 0) ? 1u : 0u;
            const uint64_t b1 = make_desc(d_u + s0 * (uint32_t)P.d_slot_bytes + (uint32_t)x0 * 16, 128, sbo);
            if (elect_one()) umma_f16(d_tmem, adesc, b1, idesc_mn(N1), accum);
            if (N2 > 0) {
              const uint64_t b2 = make_desc(d_u + (uint32_t)x0 * 16, 128, sbo);
              if (elect_one()) umma_f16(d_tmem + (uint32_t)N1, adesc, b2, idesc_mn(N2), accum);
            }
          }
          if (++steps_in_period == P.flush_steps) {
            if (elect_one()) umma_commit(&bars[BAR_FULL_ACC + set]);
            ++period; steps_in_period = 0;
          }
        }
        if (elect_one()) umma_commit(&bars[BAR_FREE + es]);        // (a warm-up step has no instruction: the slot is free at once)
        __syncwarp();
      }
    }
    if (steps_in_period > 0 && elect_one()) umma_commit(&bars[BAR_FULL_ACC + (period & 1)]);
    __syncwarp();
    WPROF_PUT(1, clock64() - pf_start); WPROF_PUT(2, pf_w0); WPROF_PUT(3, pf_w1);
  } else if (lane == 0) {
    // =========================================================================== producer: TMA bulk copies, kStage steps ahead
    int n = 0;
    for (Seq sq(g0, g1, P.HP); sq.next();) {
      const int nsteps = sq.pb - sq.pa + 2;
      for (int t = 0; t < nsteps; ++t, ++n) {
        const uint32_t ss = (uint32_t)n % kStage;
        WPROF_WAIT(pf_w0, mbar_wait_park(&bars[BAR_STAGE_FREE + ss], (((uint32_t)n / kStage) & 1) ^ 1));   // the fill warps are done with this staging slot
        const int p = sq.pa + t - 2;
        const int nrows = t >= 2 ? min(2, H - 2 * p) : 0;
        if (nrows == 0) { mbar_arrive(&bars[BAR_STAGE_FULL + ss]); continue; }
        // the two rows of the pair are contiguous in the image: ONE bulk copy per step
        mbar_expect_tx(&bars[BAR_STAGE_FULL + ss], (uint32_t)(nrows * P.row_bytes));
        bulk_g2s(stage + (size_t)ss * P.stage_bytes + kFrontSlack, P.x + ((size_t)sq.b * H + 2 * p) * W * C, (uint32_t)(nrows * P.row_bytes), &bars[BAR_STAGE_FULL + ss]);
      }
    }
    WPROF_PUT(9, pf_w0);
  }
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

// G of tap row ky, window row v (kx * C + c, or 8 * nbx + kx for the flag block), column n of a dY row: the even input rows
// (lanes 16 j + e, column block 4 - ky) plus the odd input rows (lanes 16 j + 8 + e, column block 5 - ky)
__device__ __forceinline__ float g_at(const Plan& P, int ky, int v, int n) {
  const int lane = 16 * (v >> 3) + (v & 7);
  return P.gsum[(size_t)((4 - ky) * P.CB + n) * 128 + lane] + P.gsum[(size_t)((5 - ky) * P.CB + n) * 128 + lane + 8];
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
  if (((W * C * 2) % 16) != 0) return false;
  const int nbx = (5 * C + 7) / 8;
  if (nbx + 1 > 8) return false;
  P->B = B; P->H = H; P->W = W; P->C = C; P->HP = (H + 1) / 2; P->PH = H / 2; P->PW = W / 2; P->nets = nets;
  P->CB = (int)round_up(nets * 2 * CO, 8); P->NBR = P->CB / 8; P->N = 6 * P->CB; P->nbx = nbx;
  if (P->CB > kMaxCB || ((2 * P->CB) % 16) != 0 || 2 * P->N > 512) return false;
  P->row_bytes = W * C * 2;
  P->stage_bytes = (int)round_up(kFrontSlack + 2 * P->row_bytes + kBackSlack, 16);
  P->e_slot_bytes = 16 * W * 16;
  P->d_slot_bytes = 2 * P->NBR * W * 16;
  P->off_dy = (uint32_t)(kESlots * P->e_slot_bytes);
  P->off_stage = P->off_dy + (uint32_t)(kDSlots * P->d_slot_bytes);
  P->off_bars = P->off_stage + (uint32_t)(kStage * P->stage_bytes);
  P->off_tmem = P->off_bars + BAR_COUNT * 8;
  P->smem_bytes = P->off_tmem + 16;
  if (P->smem_bytes > 225 * 1024) return false;
  P->flush_steps = std::max(1, g_wgrad_flush_steps / (W / 16));              // every accumulator element takes W / 16 instructions per step
  P->grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)B * P->HP, sm_budget()));
  P->part_floats = P->N * 128;
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
