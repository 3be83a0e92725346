// Conv weight / bias gradients on tcgen05 (base_network.py:103-123 backwards; see conv_wgrad_mma.cu for the algebra):
//
//   G[(ky,kx,c), (n,piece,o)] = sum over output pixels p of  X[p + (ky,kx)][c] * dY_n[p][o]
//
// The reduction runs over PIXELS, so both UMMA operands are MN-major: a 16-byte shared-memory vector holds 8 consecutive
// rows of the operand at ONE pixel, consecutive pixels of an image row are consecutive vectors (K = 16 pixels per
// instruction = two 128-byte core matrices, LBO = 128), and the 8-row blocks of an operand sit at ONE uniform stride (SBO).
// Both filter-tap directions are turned into that form without replicating data per tap:
//  * kx -> rows of A.  The KS-pixel window of column x, [X[x-PAD][0..C), ..., X[x+PAD][0..C)] = KS*C contiguous halfs of the
//    (zero-padded) row, is cut into 8-row blocks and stored ONCE per input row as planes.
//  * ky -> columns of B.  With the input row fixed, tap ky meets output row y = r + PAD - ky, so the B operand of ONE
//    instruction is the KS (conv1: six, see below) consecutive dY rows around r; dY rows live in a ring laid out
//    [row][column block][x] (the first slots mirrored behind the last so that a window never wraps): one descriptor.
// Two layer classes:
//  mode 0, conv1 on raw pixels (C <= 11: c3): the window (5 C halfs, a sliding 16-byte copy of the staged raw row) plus one
//    block of "tap inside the image" flags (the constant-one channel of the whitening fold) is <= 8 blocks, so a tile is an
//    input row PAIR (block 2j + row parity, M = 128) against SIX dY rows (both networks x hi/lo x 10 filters = 40 columns
//    each, N = 240): column block t is tap 4 - t for the even input row, 5 - t for the odd one.  One instruction per
//    (row pair, 16 pixels) covers 2 x 25 taps; 4 KB (A) + 7.5 KB (B) of shared-memory reads.
//  mode 1, conv2 / conv3 on the 24-channel piece layout of conv_tc.cuh (48 bytes per pixel, constant-one channel included):
//    the window is KS x 3 whole 16-byte vectors (15 or 9 blocks), copied from the staged row by aligned 16-byte moves (no
//    re-layout arithmetic); a tile is ONE input row against KS dY rows of 24 columns (N = 128 / 80).
//  * D: [128 x N] fp32 in TMEM, two sets: the MMAs of flush period i + 1 overlap the drain of period i.  The tensor-core
//    accumulator truncates, so a period is g_wgrad_flush_steps K-steps per element; 24 warps add every period into
//    fp32 registers (fixed order -> deterministic) and write one partial per CTA; reduce + finalize kernels follow.
// Roles: 24 general warps (the fill items of a step rotate over them; every warp also drains its TMEM columns) | 1 MMA warp
// (one elected lane issues) | 1 producer warp (one TMA bulk copy of the unit's raw rows per step into a staging ring,
// kStage steps ahead; the d(pooled) / arg-max rows are read straight from L2 by the dY items).
#include <algorithm>
#include "conv_wgrad_tc.cuh"
#include "conv_tc.cuh"
#include "umma.cuh"

namespace cpp {
namespace wgtc {

using namespace umma;

constexpr int CO = kConvCout;
constexpr int kESlots = 4;          // ring of expanded input units (row pairs / rows)
constexpr int kDSlots = 8;          // ring of dY units; the first nwin - 1 are mirrored behind the last
constexpr int kStage = 8;           // staging ring of raw rows (mode 0), one slot per step
constexpr int kGenWarps = 24;       // fill + epilogue warps: 6 groups of 4 (one warp per TMEM lane quarter)
constexpr int kMmaWarp = kGenWarps, kProdWarp = kMmaWarp + 1;
constexpr int kThreads = 32 * (kProdWarp + 1);
constexpr int kFrontSlack = 48, kBackSlack = 64;   // bytes around the two staged raw rows: windows of the border columns start / end outside them
constexpr int kMaxCG = 40;          // TMEM columns one epilogue warp accumulates
enum { BAR_FULL = 0, BAR_FREE = BAR_FULL + kESlots, BAR_STAGE_FULL = BAR_FREE + kESlots, BAR_STAGE_FREE = BAR_STAGE_FULL + kStage,
       BAR_FULL_ACC = BAR_STAGE_FREE + kStage, BAR_EMPTY_ACC = BAR_FULL_ACC + 2, BAR_COUNT = BAR_EMPTY_ACC + 2 };

struct Plan {
  const __half* x; const float* mean_inv;
  const float* g[kMaxNets]; const uint8_t* amax[kMaxNets]; const float* gmax[kMaxNets];
  float* dw[kMaxNets]; float* db[kMaxNets];
  float* partials; float* gsum;
  int mode;                       // 0: raw pixels, input row pairs; 1: 24-channel pieces, single input rows
  int B, H, W, C, KS, PAD, UP, PH, PW, nets;
  int U, HU, nwin;                // rows per unit, halo units on either side, units per dY window (2 HU + 1)
  int CB, NBR, Npad, CG, nbx;     // columns / 8-column blocks of one dY row, MMA N, columns per epilogue group, window blocks
  int items;                      // fill items (warps) per step
  int row_bytes, stage_bytes;
  int e_slot_bytes, d_slot_bytes;
  int flush_steps, grid, part_floats, tmem_cols;
  uint32_t off_dy, off_stage, off_bars, off_tmem, smem_bytes;
};

// power of two that brings max|g| just under 2^15 (1 when the tensor is all zero or not finite) - same rule as conv_wgrad_mma.cu
__device__ __forceinline__ float scale_for(float mx) {
  if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
  int e;
  frexpf(mx, &e);
  return ldexpf(1.f, 15 - e);
}

// The CTA's input units [g0, g1) of the batch (image-major; a unit = U rows) as segments that never cross an image.  A segment
// of units [ua, ub) is processed as steps t = 0 .. ub - ua + 2 HU - 1: step t stages dY unit q = ua - HU + t (zero outside the
// image) and, once t >= 2 HU, input unit p = ua + t - 2 HU, whose instruction reads the nwin dY units staged last.  Every role
// walks the same sequence.
struct Seq {
  int g, g_end, UP;
  int b, ua, ub;
  __device__ Seq(int g0, int g1, int UP_) : g(g0), g_end(g1), UP(UP_), b(0), ua(0), ub(0) {}
  __device__ bool next() {
    if (g >= g_end) return false;
    b = g / UP; ua = g - b * UP;
    ub = min(UP, ua + (g_end - g));
    g += ub - ua;
    return true;
  }
};

// Pipeline diagnosis build (nvcc -DWGTC_PROF, scripts/prof_wgrad_tc.py): per CTA, the cycles each role spends waiting on each
// hand-over barrier.  Slots: 0 kernel, 1 MMA loop, 2 MMA waits FULL, 3 MMA waits EMPTY_ACC, 4 warp 0 loop, 5 warp 0 waits FREE,
// 6 warp 0 waits STAGE_FULL, 7 warp 0 item work, 8 warp 4 waits FREE, 9 producer waits STAGE_FREE, 10 warp 4 item work,
// 11 warp 4 waits STAGE_FULL
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
  const int H = P.H, W = P.W, C = P.C, HU = P.HU;
  const long long G = (long long)P.B * P.UP;
  const int g0 = (int)(G * blockIdx.x / gridDim.x), g1 = (int)(G * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < kESlots; ++i) { mbar_init(&bars[BAR_FULL + i], (uint32_t)P.items); mbar_init(&bars[BAR_FREE + i], 1); }
    for (int i = 0; i < kStage; ++i) { mbar_init(&bars[BAR_STAGE_FULL + i], 1); mbar_init(&bars[BAR_STAGE_FREE + i], (uint32_t)P.items); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[BAR_FULL_ACC + i], 1); mbar_init(&bars[BAR_EMPTY_ACC + i], kGenWarps); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  // dY columns beyond the networks' own (and the slack slot behind the ring) are never written again
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
    // The fill items of step n go to warps (items n + i) mod 24, so consecutive steps are staged concurrently by disjoint warps
    // (the latency of one item - loads behind the tensor core's operand fetches - hides behind the other steps).  Between
    // items every warp polls for a finished flush period and drains its TMEM columns.
    const int quarter = warp & 3, grp = warp >> 2;
    const int c_lo = grp * P.CG, ncol = max(0, min(P.CG, P.Npad - c_lo));      // this warp's columns [c_lo, c_lo + ncol)
    float acc[kMaxCG];
#pragma unroll
    for (int c = 0; c < kMaxCG; ++c) acc[c] = 0.f;
    const int units = g1 - g0, nper = (units + P.flush_steps - 1) / P.flush_steps;
    int pd = 0;                                                    // next flush period to drain
    auto drain = [&](bool blocking) {
      while (pd < nper) {
        const uint32_t set = pd & 1, par = (pd >> 1) & 1;
        if (blocking) { WPROF_WAIT(pf_w0, mbar_wait_park(&bars[BAR_FULL_ACC + set], par)); }
        else if (!mbar_test(&bars[BAR_FULL_ACC + set], par)) break;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + set * (uint32_t)P.Npad + (uint32_t)c_lo;
#pragma unroll
        for (int c0 = 0; c0 < kMaxCG; c0 += 8) {
          uint32_t r[8];
          if (c0 < ncol) { tmem_ld4(taddr + c0, r); tmem_ld4(taddr + c0 + 4, r + 4); }
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c0 < ncol) acc[c0 + c] += __uint_as_float(r[c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_EMPTY_ACC + set]);
        ++pd;
      }
    };
    const uint32_t bsW = (uint32_t)W * 16;
    const int my_phase = warp / P.items, item = warp - my_phase * P.items;          // my_phase >= kESlots: no fill work, epilogue only
    int n = 0;
    for (Seq sq(g0, g1, P.UP); sq.next();) {
      const int nsteps = sq.ub - sq.ua + 2 * HU;
      for (int t = 0; t < nsteps; ++t, ++n) {
        // rotation over items * kESlots warps: warp w owns item w % items of the steps n = w / items (mod kESlots).  It thereby
        // meets EVERY phase of the ring barriers it waits on (a parity wait that skipped a phase could pass one ring turn early)
        if (((uint32_t)n & (kESlots - 1)) != (uint32_t)my_phase) continue;
        drain(false);
        const uint32_t es = (uint32_t)n % kESlots, ds = (uint32_t)n % kDSlots, ss = (uint32_t)n % kStage;
        WPROF_WAIT(pf_w1, mbar_wait_park(&bars[BAR_FREE + es], (((uint32_t)n / kESlots) & 1) ^ 1));   // the MMAs that read this E slot one ring turn ago are done
        WPROF_WAIT(pf_w2, mbar_wait_park(&bars[BAR_STAGE_FULL + ss], ((uint32_t)n / kStage) & 1));   // the producer's copy for this step has landed
        const int q = sq.ua - HU + t, p = sq.ua + t - 2 * HU;
        const bool has_x = t >= 2 * HU;
        const int py = P.U == 2 ? q : (q >> 1);                            // pooled row that holds dY unit q
        const bool dy_data = q >= 0 && py < P.PH;
        const uint8_t* st = stage + (size_t)ss * P.stage_bytes;
        uint8_t* eslot = E + (size_t)es * P.e_slot_bytes;
#ifdef WGTC_PROF
        const long long pf_i0 = clock64();
#endif
        const int n_x_items = P.mode == 0 ? 4 : P.KS;
        if (item < n_x_items && P.mode == 0) {
          // input row `rpar` of the pair, columns of parity `xpar`, one lane per column (for an odd channel count the words of
          // 32 windows two pixels apart sit in 32 distinct banks); ALL window blocks of the column from one run of loads.
          // The pixels of a row are stored even columns first, then odd ones (the reduction index of the MMA may be permuted
          // as long as A and B agree): the lanes of a warp write consecutive vectors.
          const int rpar = item >> 1, xpar = item & 1;
          const int xcol = 2 * lane + xpar, xpos = lane + (W / 2) * xpar;
          if (has_x && xcol < W) {
            uint8_t* dst = eslot + (size_t)(rpar * W + xpos) * 16;                                  // block j at + 2 j W 16
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
        } else if (item < n_x_items) {
          // mode 1, tap column kx = item: the three 16-byte vectors of pixel x - PAD + kx (staged row, 48 bytes per pixel) go to
          // blocks 3 kx + g; pixels outside the row are the zero padding
          if (has_x) {
            const int kx = item;
            for (int x = lane; x < W; x += 32) {
              const int xs = x - P.PAD + kx;
              const bool ok = xs >= 0 && xs < W;
              const uint4* src = reinterpret_cast<const uint4*>(st + (size_t)(ok ? xs : 0) * (tc::kC24 * 2));
              uint4 v[3];
#pragma unroll
              for (int gq = 0; gq < 3; ++gq) v[gq] = ok ? src[gq] : make_uint4(0, 0, 0, 0);
#pragma unroll
              for (int gq = 0; gq < 3; ++gq) *reinterpret_cast<uint4*>(eslot + (size_t)((3 * kx + gq) * W + x) * 16) = v[gq];
            }
          }
        } else {
          // dY unit q: one lane per (window column, network); the unit's rows from pooled row py (both rows of a pair, or the
          // row of parity q & 1), as hi + lo fp16 pieces scaled by the network's power of two
          uint8_t* dslot = dyb + (size_t)ds * P.d_slot_bytes;
          const int first = (item - n_x_items) * 32, step = (P.items - n_x_items) * 32;
          for (int idx = first + lane; idx < P.PW * P.nets; idx += step) {
            const int net = idx / P.PW, px = idx - net * P.PW;
            const float sc = s_scale[net];
            // pooled gradients and arg-max bytes straight from global memory (L2): several steps are in flight on different warps
            const size_t qo = (((size_t)sq.b * P.PH + (dy_data ? py : 0)) * P.PW + px) * CO;
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
              if (P.U == 1 && (pa4 >> 1) != (q & 1)) continue;    // single-row units: only the window row this dY row is
              // word i of this pixel's 10 (5 hi + 5 lo half2 pairs), masked by "the arg-max of the window is this pixel"
              auto word = [&](int i) -> uint32_t {
                const int v = i < 5 ? i : i - 5;
                const uint32_t eq = __vcmpeq4(aq[v], (uint32_t)pa4 * 0x0101u);      // 0xff in the bytes whose arg-max is this pixel
                return (i < 5 ? hi2[v] : lo2[v]) & __byte_perm(eq, 0u, 0x1100);     // byte 0 -> low half, byte 1 -> high half
              };
              const int rp = P.U == 2 ? (pa4 >> 1) : 0;
              const int xpos = P.mode == 0 ? px + (W / 2) * (pa4 & 1) : 2 * px + (pa4 & 1);     // the pixel order of the A planes
              // words net * 10 .. + 9 of the row's columns (column = 2 * word): 16-byte stores where a block is complete
              const size_t off = (size_t)(rp * P.NBR + ((net * 10) >> 2)) * bsW + (size_t)xpos * 16;
#pragma unroll
              for (int mir = 0; mir < 2; ++mir) {                 // the first nwin - 1 ring slots are mirrored behind the last
                if (mir == 1 && (int)ds >= P.nwin - 1) break;
                uint8_t* dpx = dslot + (mir ? (size_t)kDSlots * P.d_slot_bytes : 0) + off;
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
        }
        fence_proxy_async();                                       // generic-proxy / cp.async writes -> visible to the tensor core
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
    for (int c = 0; c < kMaxCG; ++c)
      if (c < ncol) part[(size_t)(c_lo + c) * 128 + 32 * quarter + lane] = acc[c];
  } else if (warp == kMmaWarp) {
    // =========================================================================== MMA issue
    const uint32_t e_u = smem_u32(E), d_u = smem_u32(dyb);
    const uint32_t sbo = (uint32_t)W * 16;
    const uint32_t idesc = idesc_mn(P.Npad);
    int n = 0, steps_in_period = 0;
    uint32_t period = 0;
    for (Seq sq(g0, g1, P.UP); sq.next();) {
      const int nsteps = sq.ub - sq.ua + 2 * HU;
      for (int t = 0; t < nsteps; ++t, ++n) {
        const uint32_t es = (uint32_t)n % kESlots;
        WPROF_WAIT(pf_w0, mbar_wait_park(&bars[BAR_FULL + es], ((uint32_t)n / kESlots) & 1));
        tc_fence_after();
        if (t >= 2 * HU) {
          const uint32_t set = period & 1;
          if (steps_in_period == 0) { WPROF_WAIT(pf_w1, mbar_wait_park(&bars[BAR_EMPTY_ACC + set], ((period >> 1) & 1) ^ 1)); tc_fence_after(); }
          // the nwin dY units staged last: ring slots n - nwin + 1 .. n (mod kDSlots), contiguous thanks to the mirror slots
          const uint32_t s0 = (uint32_t)(n - (P.nwin - 1)) % kDSlots;
          const uint32_t d_tmem = tmem_base + set * (uint32_t)P.Npad;
          for (int x0 = 0; x0 < W; x0 += 16) {
            const uint64_t adesc = make_desc(e_u + es * (uint32_t)P.e_slot_bytes + (uint32_t)x0 * 16, 128, sbo);
            const uint64_t bdesc = make_desc(d_u + s0 * (uint32_t)P.d_slot_bytes + (uint32_t)x0 * 16, 128, sbo);
            if (elect_one()) umma_f16(d_tmem, adesc, bdesc, idesc, (steps_in_period > 0 || x0 > 0) ? 1u : 0u);
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
    for (Seq sq(g0, g1, P.UP); sq.next();) {
      const int nsteps = sq.ub - sq.ua + 2 * HU;
      for (int t = 0; t < nsteps; ++t, ++n) {
        const uint32_t ss = (uint32_t)n % kStage;
        WPROF_WAIT(pf_w0, mbar_wait_park(&bars[BAR_STAGE_FREE + ss], (((uint32_t)n / kStage) & 1) ^ 1));   // the fill warps are done with this staging slot
        const int p = sq.ua + t - 2 * HU;
        const int nrows = t >= 2 * HU ? min(P.U, H - P.U * p) : 0;
        if (nrows == 0) { mbar_arrive(&bars[BAR_STAGE_FULL + ss]); continue; }
        // the rows of a unit are contiguous in the image: ONE bulk copy per step
        mbar_expect_tx(&bars[BAR_STAGE_FULL + ss], (uint32_t)(nrows * P.row_bytes));
        bulk_g2s(stage + (size_t)ss * P.stage_bytes + (P.mode == 0 ? kFrontSlack : 0), P.x + ((size_t)sq.b * H + (size_t)P.U * p) * W * C,
                 (uint32_t)(nrows * P.row_bytes), &bars[BAR_STAGE_FULL + ss]);
      }
    }
    WPROF_PUT(9, pf_w0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
#ifdef WGTC_PROF
  if (tid == 0) g_wprof[blockIdx.x][0] = (unsigned long long)(clock64() - pf_start);
#endif
}

// gsum[i] = sum over CTAs of partials[cta][i] in fixed order; block (32, 8), four consecutive floats per thread (part_floats is a
// multiple of 128, the buffers are 256-byte aligned): a warp reads 512 contiguous bytes of every partial
__global__ void __launch_bounds__(256) wgtc_reduce_kernel(const float* __restrict__ partials, int nparts, int part_floats, float* __restrict__ gsum) {
  __shared__ float4 sh[8][33];
  const int i = (blockIdx.x * 32 + threadIdx.x) * 4, y = threadIdx.y;
  const int per = (nparts + 7) / 8, k0 = y * per, k1 = min(nparts, k0 + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < part_floats)
    for (int k = k0; k < k1; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(partials + (size_t)k * part_floats + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  sh[y][threadIdx.x] = s;
  __syncthreads();
  if (y == 0 && i < part_floats) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) { const float4 v = sh[r][threadIdx.x]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    *reinterpret_cast<float4*>(gsum + i) = t;
  }
}

// mode 0: G of tap row ky, window row v (kx * C + c, or 8 * nbx + kx for the flag block), column n of a dY row: the even input
// rows (lanes 16 j + e, column block 4 - ky) plus the odd input rows (lanes 16 j + 8 + e, column block 5 - ky)
__device__ __forceinline__ float g_at0(const Plan& P, int ky, int v, int n) {
  const int lane = 16 * (v >> 3) + (v & 7);
  return P.gsum[(size_t)((4 - ky) * P.CB + n) * 128 + lane] + P.gsum[(size_t)((5 - ky) * P.CB + n) * 128 + lane + 8];
}
// mode 1: window row v = kx * 24 + piece channel = the lane; tap ky sits in column block KS - 1 - ky
__device__ __forceinline__ float g_at1(const Plan& P, int ky, int v, int n) {
  return P.gsum[(size_t)((P.KS - 1 - ky) * P.CB + n) * 128 + v];
}
__global__ void __launch_bounds__(256) wgtc_finalize_kernel(const __grid_constant__ Plan P) {
  const int KS = P.KS, Cr = P.mode == 0 ? P.C : CO, nw = KS * KS * Cr * CO;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nets * (nw + CO)) return;
  const int net = i / (nw + CO), j = i - net * (nw + CO);
  const float inv_scale = 1.f / scale_for(P.gmax[net][0]);
  const int nb0 = net * 2 * CO;                                           // columns of this network: hi pieces, then lo pieces
  if (j >= nw) {                                                          // bias gradient: constant-one channel, centre tap
    const int o = j - nw;
    float s;
    if (P.mode == 0) s = g_at0(P, 2, 8 * P.nbx + 2, nb0 + o) + g_at0(P, 2, 8 * P.nbx + 2, nb0 + CO + o);
    else s = g_at1(P, P.PAD, P.PAD * tc::kC24 + tc::kC24One, nb0 + o) + g_at1(P, P.PAD, P.PAD * tc::kC24 + tc::kC24One, nb0 + CO + o);
    P.db[net][o] = s * inv_scale;
    return;
  }
  const int o = j % CO, c = (j / CO) % Cr, kx = (j / (CO * Cr)) % KS, ky = j / (CO * Cr * KS);
  const int n0 = nb0 + o, n1 = n0 + CO;
  if (P.mode == 1) {                                                      // dW[c] = G[hi(c)] + G[lo(c)] (pieces of the activation)
    const int vh = kx * tc::kC24 + tc::c24_hi(c), vl = kx * tc::kC24 + tc::c24_lo(c);
    const float gsum = g_at1(P, ky, vh, n0) + g_at1(P, ky, vh, n1) + g_at1(P, ky, vl, n0) + g_at1(P, ky, vl, n1);
    P.dw[net][j] = gsum * inv_scale;
    return;
  }
  const float gsum = g_at0(P, ky, kx * P.C + c, n0) + g_at0(P, ky, kx * P.C + c, n1);
  float v = gsum * inv_scale;
  if (P.mean_inv) {
    const float ssum = (g_at0(P, ky, 8 * P.nbx + kx, n0) + g_at0(P, ky, 8 * P.nbx + kx, n1)) * inv_scale;
    v = P.mean_inv[P.C + c] * (v - P.mean_inv[c] * ssum);
  }
  P.dw[net][j] = v;
}

// ------------------------------------------------------------------------------------------ host
static inline size_t al256(size_t b) { return (size_t)round_up((int64_t)b, 256); }

static bool build_plan(int pieces, int nets, int B, int H, int W, int C, int KS, Plan* P) {
  if (nets < 1 || nets > kMaxNets || H < 2 || W < 16 || W > 64 || (W % 16) != 0 || C < 1) return false;
  P->mode = pieces ? 1 : 0;
  P->B = B; P->H = H; P->W = W; P->C = C; P->KS = KS; P->PAD = KS / 2; P->PH = H / 2; P->PW = W / 2; P->nets = nets;
  P->CB = (int)round_up(nets * 2 * CO, 8); P->NBR = P->CB / 8;
  if (P->mode == 0) {
    if (KS != 5 || ((W * C * 2) % 16) != 0) return false;
    P->nbx = (5 * C + 7) / 8;
    if (P->nbx + 1 > 8) return false;
    P->U = 2; P->HU = 1; P->items = 6;
    P->row_bytes = W * C * 2;
    P->stage_bytes = (int)round_up(kFrontSlack + 2 * P->row_bytes + kBackSlack, 16);
  } else {
    if ((KS != 5 && KS != 3) || C != tc::kC24 || nets != 1) return false;
    P->nbx = 3 * KS;
    P->U = 1; P->HU = P->PAD; P->items = KS + 1;
    P->row_bytes = W * tc::kC24 * 2; P->stage_bytes = P->row_bytes;
  }
  P->nwin = 2 * P->HU + 1;
  P->UP = (H + P->U - 1) / P->U;
  P->Npad = (int)round_up(P->nwin * P->U * P->CB, 16);
  P->CG = (int)round_up(ceil_div(P->Npad, kGenWarps / 4), 8);
  if (P->CG > kMaxCG || P->Npad > 256 || P->items * kESlots > kGenWarps) return false;
  P->tmem_cols = 2 * P->Npad <= 256 ? 256 : 512;
  if (2 * P->Npad > 512) return false;
  P->e_slot_bytes = 16 * W * 16;
  P->d_slot_bytes = P->U * P->NBR * W * 16;
  P->off_dy = (uint32_t)(kESlots * P->e_slot_bytes);
  P->off_stage = P->off_dy + (uint32_t)((kDSlots + P->nwin) * P->d_slot_bytes);      // ring + mirror of the first nwin - 1 slots + one slack slot
  P->off_bars = P->off_stage + (uint32_t)(kStage * P->stage_bytes);
  P->off_tmem = P->off_bars + BAR_COUNT * 8;
  P->smem_bytes = P->off_tmem + 16;
  if (P->smem_bytes > 225 * 1024) return false;
  P->flush_steps = std::max(1, g_wgrad_flush_steps / (W / 16));              // every accumulator element takes W / 16 instructions per step
  P->grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)B * P->UP, sm_budget()));
  P->part_floats = P->Npad * 128;
  return true;
}

bool supported(int nets, int H, int W, int C, int KS, int pieces) {
  Plan P{};
  return build_plan(pieces, nets, 1, H, W, C, KS, &P);
}

int64_t scratch_bytes(int nets, int H, int W, int C, int KS, int pieces) {
  Plan P{};
  if (!build_plan(pieces, nets, 1, H, W, C, KS, &P)) return 0;
  return (int64_t)(al256((size_t)kNumSMs * P.part_floats * 4) + al256((size_t)P.part_floats * 4));
}

int launch(const void* x_f16, const float* mean_inv, int nets, const float* const* d_pooled, const uint8_t* const* amax, int B, int H,
           int W, int C, int KS, float* const* dw, float* const* db, const float* const* gmax, void* scratch, cudaStream_t s, int pieces) {
  if (B <= 0) return CPP_OK;
  Plan P{};
  CPP_REQUIRE(build_plan(pieces, nets, B, H, W, C, KS, &P), "wgrad_tc: unsupported layer %dx%dx%d k%d, %d networks", H, W, C, KS, nets);
  CPP_REQUIRE(((uintptr_t)x_f16 & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "wgrad_tc: unaligned input or scratch");
  CPP_REQUIRE(!pieces || mean_inv == nullptr, "wgrad_tc: the piece input has no whitening");
  P.x = reinterpret_cast<const __half*>(x_f16); P.mean_inv = mean_inv;
  for (int n = 0; n < nets; ++n) {
    CPP_REQUIRE(d_pooled[n] && amax[n] && dw[n] && db[n] && gmax[n], "wgrad_tc: null pointer for network %d", n);
    CPP_REQUIRE(((uintptr_t)d_pooled[n] & 7) == 0 && ((uintptr_t)amax[n] & 1) == 0, "wgrad_tc: unaligned gradient / arg-max of network %d", n);
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
  wgtc_reduce_kernel<<<(unsigned)ceil_div(P.part_floats, 128), dim3(32, 8), 0, s>>>(P.partials, P.grid, P.part_floats, P.gsum);
  CPP_CHECK_LAUNCH();
  const int Cr = pieces ? CO : C;
  const int total = nets * (KS * KS * Cr * CO + CO);
  wgtc_finalize_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace wgtc
}  // namespace cpp
