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
#include "conv_row_tc.cuh"
#include "umma.cuh"

namespace cpp {
namespace tc {

constexpr int CO = kConvCout;
constexpr int kSmemLimit = 224 * 1024;   // dynamic shared memory one CTA may use (227 KB max on sm_100)

using namespace umma;

// activations whose fp16 piece pair saturated since the last reset (cpp_piece_overflow_count)
__device__ unsigned int g_piece_overflow = 0;
int piece_overflow_count(int reset, unsigned int* out) {
  CPP_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_piece_overflow, sizeof(unsigned int)));
  if (reset) { const unsigned int z = 0; CPP_CHECK_CUDA(cudaMemcpyToSymbol(g_piece_overflow, &z, sizeof(unsigned int))); }
  unsigned int row = 0;
  CPP_TRY(tcr::piece_overflow_count(reset, &row));               // the row-sweep kernel of conv2 / conv3 keeps its own counter
  *out += row;
  return CPP_OK;
}

// instruction descriptor: D fp32, A/B fp16, both K-major, M = 128, N
__host__ __device__ inline uint32_t make_idesc(int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ weight packing
// Every block recomputes (1), the blocks share (2) and (3).  (1) S: power of two that brings max |w * inv| just under 2^15; (2) B operand in canonical K-major
// layout [pair][half][n][8], n = (net * pieces + piece) * 10 + o; (3) corr[yc][xc][net][o] = bias - sum over the
// taps that fall inside the image of mean_c * inv_c * w (fp64), and 2^-S.
__global__ void __launch_bounds__(256) conv_tc_prep_kernel(const __grid_constant__ FwdPlan P, const __grid_constant__ PrepArgs A) {
  __shared__ float red[256];
  __shared__ float s_scale;
  const int tid = threadIdx.x;
  const int KS = P.KS, C = P.C, Cw = P.Cw, nw = KS * KS * Cw * CO;
  float mx = 0.f;
  for (int n = 0; n < P.nets; ++n)
    for (int i = tid; i < nw; i += blockDim.x) {
      const int ch = (i / CO) % Cw;
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
    if (blockIdx.x == 0) A.corr[(2 * P.PAD + 1) * (2 * P.PAD + 1) * P.nets * CO] = ldexpf(1.f, e - 15);
  }
  __syncthreads();
  const float scale = s_scale;
  const int N = P.N, used = P.nets * kPieces * CO, total = P.n_pairs * 2 * N * 8;
  const int gtid = blockIdx.x * blockDim.x + tid, gthreads = gridDim.x * blockDim.x;
  for (int idx = gtid; idx < total; idx += gthreads) {
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
      const int chw = P.in_layout == 2 ? c24_weight_channel(ch) : (ch < Cw ? ch : ch - Cw);
      float wv = 0.f;
      if (chw >= 0)
        wv = P.dgrad ? A.w[net][(((KS - 1 - ky) * KS + (KS - 1 - kx)) * CO + o) * CO + chw]      // flipped taps, in/out swapped
                     : A.w[net][((ky * KS + kx) * Cw + chw) * CO + o];
      v = wv * inv * scale;
      const __half hi = __float2half_rn(v);
      out = (piece == 0) ? hi : __float2half_rn(v - __half2float(hi));
    }
    A.bpack[idx] = out;
  }
  // (3) one warp per table entry: the lanes split the taps x channels of the border-aware mean term (fp64, fixed order)
  const int ncls = 2 * P.PAD + 1;
  const int lane = tid & 31, gwarp = gtid >> 5, gwarps = gthreads >> 5;
  for (int idx = gwarp; idx < ncls * ncls * P.nets * CO; idx += gwarps) {
    const int o = idx % CO, net = (idx / CO) % P.nets, xc = (idx / (CO * P.nets)) % ncls, yc = idx / (CO * P.nets * ncls);
    double acc = 0.0;
    if (A.mean_inv) {
      const int ky0 = yc < P.PAD ? P.PAD - yc : 0, ky1 = yc > P.PAD ? KS - 1 - (yc - P.PAD) : KS - 1;
      const int kx0 = xc < P.PAD ? P.PAD - xc : 0, kx1 = xc > P.PAD ? KS - 1 - (xc - P.PAD) : KS - 1;
      const int nky = ky1 - ky0 + 1, nkx = kx1 - kx0 + 1, terms = nky * nkx * Cw;
      for (int t = lane; t < terms; t += 32) {
        const int ch = t % Cw, kx = kx0 + (t / Cw) % nkx, ky = ky0 + t / (Cw * nkx);
        acc -= (double)A.mean_inv[ch] * (double)A.mean_inv[C + ch] * (double)A.w[net][((ky * KS + kx) * Cw + ch) * CO + o];
      }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (lane == 0) A.corr[idx] = (float)((A.bias[net] ? (double)A.bias[net][o] : 0.0) + acc);
  }
}

// ------------------------------------------------------------------------------------------ main kernel
// Warp-specialised, persistent (one CTA per SM):
//   warps 0-7   epilogue : TMEM -> registers (tcgen05.ld), pieces summed, scale, border-aware bias, ReLU, 2x2 max-pool;
//                          two groups of four (one warp per TMEM lane quarter), group e drains accumulator set e
//   warps 8-19  fill     : TMA bulk copies of raw image rows into a staging ring, re-laid into the parity planes
//   warp  20    MMA      : one elected lane issues the tcgen05.mma stream
// Two plane buffers (fill of unit u+1 overlaps the MMAs of unit u) and two sets of TMEM accumulators (MMAs of
// tile t+1 overlap the epilogue of tile t), all handed over through mbarriers.
constexpr int kEpiWarps = 8, kEpiPerAcc = 4, kFillWarps = 12;
constexpr int kThreads2 = 32 * (kEpiWarps + kFillWarps + 1);
constexpr int kMaxStageSlots = 4;
enum { BAR_FULL_PL = 0, BAR_EMPTY_PL = 2, BAR_FULL_ACC = 4, BAR_EMPTY_ACC = 6, BAR_STAGE = 8, BAR_COUNT = 8 + kMaxStageSlots };

struct SmemLayout { uint32_t planes, bsm, stage, corr, bars, tmem, total; };

__host__ __device__ inline SmemLayout smem_layout(const FwdPlan& P) {
  SmemLayout L;
  uint32_t off = 0;
  L.planes = off; off += 2u * (uint32_t)P.unit_bytes;
  L.bsm = off; off += (uint32_t)P.n_pairs * 2 * P.N * 16;
  L.stage = off; off += (uint32_t)P.stage_slots * (uint32_t)P.stage_bytes;
  const int ncls = 2 * P.PAD + 1;
  L.corr = off; off += (uint32_t)((ncls * ncls * P.nets * CO + 4) * 4);
  off = (off + 15) & ~15u;
  L.bars = off; off += BAR_COUNT * 8;
  L.tmem = off; off += 16;
  L.total = off;
  return L;
}

// up to 8 halfs (zero padded) from a 2-byte aligned shared-memory address
__device__ __forceinline__ uint4 load8h(const unsigned short* p, int nch) {
  if (nch >= 8) {
    if ((reinterpret_cast<uintptr_t>(p) & 2) == 0) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
      return make_uint4(q[0], q[1], q[2], q[3]);
    }
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p + 1);
    const uint32_t a = p[0], b = q[0], c = q[1], d = q[2], e = p[7];
    return make_uint4(a | (b << 16), (b >> 16) | (c << 16), (c >> 16) | (d << 16), (d >> 16) | (e << 16));
  }
  uint32_t h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) h[e] = e < nch ? (uint32_t)p[e] : 0u;
  return make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}

// 8 halfs from a 2-byte aligned shared-memory address, branch free: five aligned words, funnel-shifted by the misalignment
// (reads up to 4 bytes past the 8th half: callers keep that inside the shared-memory allocation and mask what they use)
__device__ __forceinline__ uint4 load8h_any(const unsigned short* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t sh = ((uint32_t)a & 2u) << 3;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
constexpr int kMaxG8 = 3;       // 8-channel groups of the input (C <= 24)

// Work distribution: the tiles of the whole batch (image-major) are split into one contiguous range per CTA
// (near-perfect balance); a CTA walks its range in units of <= tiles_per_unit tiles that never cross an image.
struct UnitIter {
  int g, g_end, tpi, tpu;
  int b, t0, t1;
  __device__ UnitIter(const FwdPlan& P) {
    const long long total = (long long)P.B * P.tiles_per_image;
    g = (int)(total * blockIdx.x / gridDim.x); g_end = (int)(total * (blockIdx.x + 1) / gridDim.x);
    tpi = P.tiles_per_image; tpu = P.tiles_per_unit;
  }
  __device__ bool next() {
    if (g >= g_end) return false;
    b = g / tpi; t0 = g - b * tpi;
    t1 = min(min(tpi, t0 + tpu), t0 + (g_end - g));
    g += t1 - t0;
    return true;
  }
};

// Pipeline diagnosis build (nvcc -DCONV_TC_PROF, scripts/prof_conv_tc.py): per CTA, the cycles each role spends waiting on
// each hand-over barrier.  Slots: 0 prologue, 1 MMA total, 2 MMA waits FULL_PL, 3 MMA waits EMPTY_ACC, 4 fill total,
// 5 fill waits EMPTY_PL, 6 fill waits STAGE, 7 epilogue total, 8 epilogue waits FULL_ACC, 9 whole kernel, 10 tiles, 11 units
#ifdef CONV_TC_PROF
__device__ unsigned long long g_prof[160][12];
__device__ int g_dbg_skip;     // timing experiments only (results are wrong): 1 no MMAs, 2 no re-layout, 4 no epilogue math
#define PROF_DECL unsigned long long pf_t = 0, pf_w0 = 0, pf_w1 = 0; const long long pf_start = clock64();
#define PROF_WAIT(acc, stmt) { const long long pf_a = clock64(); stmt; acc += (unsigned long long)(clock64() - pf_a); }
#define PROF_PUT(slot, v) { if (lane == 0) g_prof[blockIdx.x][slot] = (unsigned long long)(v); }
#else
#define PROF_DECL
#define PROF_WAIT(acc, stmt) { stmt; }
#define PROF_PUT(slot, v)
#endif

template <int KS, int R>
__global__ void __launch_bounds__(kThreads2, 1) conv_fwd_tc_kernel(const __grid_constant__ FwdPlan P) {
  extern __shared__ __align__(128) uint8_t smem[];
#ifdef CONV_TC_PROF
  const long long pf_k0 = clock64();
  const int pf_dbg = g_dbg_skip;
#endif
  const SmemLayout L = smem_layout(P);
  uint8_t* planes_base = smem + L.planes;
  uint8_t* bsm = smem + L.bsm;
  uint8_t* stage_base = smem + L.stage;
  float* corr_s = reinterpret_cast<float*>(smem + L.corr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler
  constexpr int PAD = KS / 2;
  const int H = P.H, W = P.W, C = P.C, PH = P.PH, PW = P.PW, Pq = P.Pq, N = P.N;
  const int ncls = 2 * PAD + 1, ncorr = ncls * ncls * P.nets * CO;

  // ---- one-time setup: barriers, TMEM, weights, correction table
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[BAR_FULL_PL + i], kFillWarps); mbar_init(&bars[BAR_EMPTY_PL + i], 1);
      mbar_init(&bars[BAR_FULL_ACC + i], 1); mbar_init(&bars[BAR_EMPTY_ACC + i], kEpiPerAcc);
    }
    for (int i = 0; i < kMaxStageSlots; ++i) mbar_init(&bars[BAR_STAGE + i], 1);
    fence_mbar_init();
  }
  if (warp == kEpiWarps + kFillWarps) tmem_alloc(tmem_slot, 512);
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.bpack);
    uint4* dst = reinterpret_cast<uint4*>(bsm);
    for (int i = tid; i < P.n_pairs * 2 * N; i += kThreads2) dst[i] = src[i];
    for (int i = tid; i < ncorr + 1; i += kThreads2) corr_s[i] = P.corr[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PROF_DECL
#ifdef CONV_TC_PROF
  if (tid == 0) g_prof[blockIdx.x][0] = (unsigned long long)(pf_start - pf_k0);
#endif

  if (warp < kEpiWarps) {
    // =========================================================================== epilogue warps
    const float scale_inv = corr_s[ncorr];
    const int quarter = warp & 3;                                  // TMEM lanes 32*quarter .. +31
    const uint32_t egroup = (uint32_t)warp >> 2;                   // this group of four drains accumulator set `egroup`
    uint32_t tc = 0;
    for (UnitIter ui(P); ui.next();) {
      const int b = ui.b;
      for (int t = ui.t0; t < ui.t1; ++t, ++tc) {
        const uint32_t ab = tc & 1;
        if (ab != egroup) continue;
        PROF_WAIT(pf_w0, mbar_wait(&bars[BAR_FULL_ACC + ab], (tc >> 1) & 1));
        tc_fence_after();
        const int q = 128 * t + 32 * quarter + lane;
        const int py = q / Pq, px = q - py * Pq;
        const bool valid = py < PH && px < PW;
        if (P.dgrad) {
          // input-gradient mode: no bias / ReLU / pool, the four window positions are four output pixels
          const float osc = scale_inv * P.out_scale[0];
          float amx = 0.f;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            uint32_t r[20];
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)((ab * 4 + a) * N);
            tmem_ld16(taddr, r);
            tmem_ld4(taddr + 16, r + 16);
            tmem_ld_wait();
            const int y = 2 * py + (a >> 1), x = 2 * px + (a & 1);
            if (valid && y < H && x < W) {
              float* op = P.pooled[0] + (((size_t)b * H + y) * W + x) * CO;
#pragma unroll
              for (int o = 0; o < CO; o += 2) {
                const float v0 = (__uint_as_float(r[o]) + __uint_as_float(r[CO + o])) * osc;
                const float v1 = (__uint_as_float(r[o + 1]) + __uint_as_float(r[CO + o + 1])) * osc;
                *reinterpret_cast<float2*>(op + o) = make_float2(v0, v1);
                amx = fmaxf(amx, fmaxf(fabsf(v0), fabsf(v1)));
              }
            }
          }
          if (P.out_absmax != nullptr) {               // max is order independent: an atomic keeps the result deterministic
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, sft));
            if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(P.out_absmax), __float_as_int(amx));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_EMPTY_ACC + ab]);
          continue;
        }
        for (int net = 0; net < P.nets; ++net) {
#ifdef CONV_TC_PROF
          if (pf_dbg & 4) break;
#endif
          float best[CO];
          int arg[CO];
#pragma unroll
          for (int o = 0; o < CO; ++o) { best[o] = 0.f; arg[o] = 0; }
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            uint32_t r[20];
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)((ab * 4 + a) * N + net * kPieces * CO);
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
            if (P.pooled_hl[net] != nullptr) {                     // the same values as fp16 pieces (24-channel layout) for the next layer
              uint32_t hi[CO / 2], lo[CO / 2];
#pragma unroll
              for (int o = 0; o < CO; o += 2) {
                const float v0 = fmaxf(best[o], 0.f), v1 = fmaxf(best[o + 1], 0.f);
                const __half h0 = __float2half_rn(fminf(v0, 65504.f)), h1 = __float2half_rn(fminf(v1, 65504.f));
                const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
                // hi + lo represent activations up to 65504 + 65504: beyond that the piece copy (not the fp32 output) saturates -
                // never silently: the counter is read by cpp_piece_overflow_count (the engines raise on it)
                if (fmaxf(v0, v1) > 131000.f) atomicAdd(&g_piece_overflow, 1u);
                hi[o >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                lo[o >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
              }
              uint4* hp = reinterpret_cast<uint4*>(P.pooled_hl[net] + (base / CO) * kC24);
              hp[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              hp[1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              hp[2] = make_uint4(hi[4], lo[4], 0x00003C00u, 0u);   // hi8 hi9 | lo8 lo9 | 1.0 0 | 0 0
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_EMPTY_ACC + ab]);     // this warp's quarter of the accumulators is drained
      }
    }
#ifdef CONV_TC_PROF
    if (warp == 0) { PROF_PUT(7, clock64() - pf_start); PROF_PUT(8, pf_w0); PROF_PUT(10, tc); }
#endif
  } else if (warp < kEpiWarps + kFillWarps) {
    // =========================================================================== fill warps
    const int ftid = tid - 32 * kEpiWarps, nfill = 32 * kFillWarps;
    const size_t img_elems = (size_t)H * W * C;
    const int n_groups = (P.rows_alloc + P.crh - 1) / P.crh;
    const int rowC = W * C;
    // channels of the last 8-group that exist (a zero-padded group when C is not 8*G8 + R)
    const int nch_last = P.R > 0 ? 8 : C - 8 * (P.G8 - 1);
    uint4 last_mask;
    {
      uint32_t m[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = nch_last >= 2 * e + 2 ? 0xffffffffu : (nch_last == 2 * e + 1 ? 0x0000ffffu : 0u);
      last_mask = make_uint4(m[0], m[1], m[2], m[3]);
    }
    uint32_t it = 0, icnt = 0, wcnt = 0;                           // units done; staging copies issued / consumed
    // producer cursor (fill thread 0): raw rows are requested up to stage_slots groups ahead of the re-layout, across
    // unit boundaries, so the TMA latency is off the fill warps' critical path
    const uint32_t NS = (uint32_t)P.stage_slots;
    UnitIter pui(P);
    bool p_ok = pui.next();
    int p_g = 0;
    auto issue_next = [&]() {                                      // thread ftid == 0 only
      while (p_ok) {
        const int pyh0 = (128 * pui.t0) / Pq - 1;
        const int rho0 = p_g * P.crh, rho1 = min(rho0 + P.crh, P.rows_alloc);
        const int yc0 = max(0, 2 * (pyh0 + rho0)), yc1 = min(H, 2 * (pyh0 + rho1));
        const __half* pimg = P.x + (size_t)(P.rows ? P.rows[pui.b] : pui.b) * img_elems;
        if (++p_g >= n_groups) { p_g = 0; p_ok = pui.next(); }
        if (yc1 > yc0) {
          const uint32_t sb = icnt % NS, bytes = (uint32_t)(yc1 - yc0) * rowC * 2;
          fence_proxy_async();                                     // the buffer was read through the generic proxy before
          mbar_expect_tx(&bars[BAR_STAGE + sb], bytes);
          bulk_g2s(stage_base + (size_t)sb * P.stage_bytes, pimg + (size_t)yc0 * rowC, bytes, &bars[BAR_STAGE + sb]);
          ++icnt;
          return;
        }
      }
    };
    if (P.use_bulk && P.in_layout != 2 && ftid == 0)
      for (uint32_t i = 0; i < NS; ++i) issue_next();
    for (UnitIter ui(P); ui.next(); ++it) {
      const int b = ui.b;
      const int py_first = (128 * ui.t0) / Pq, yh0 = py_first - 1;
      const __half* img = P.x + (size_t)(P.rows ? P.rows[b] : b) * img_elems;
      const uint32_t buf = it & 1;
      PROF_WAIT(pf_w0, mbar_wait(&bars[BAR_EMPTY_PL + buf], ((it >> 1) & 1) ^ 1));   // the MMAs that read this buffer two units ago are done
      uint8_t* planes = planes_base + (size_t)buf * P.unit_bytes;

      if (P.in_layout == 2) {
        // aligned 24-channel pieces: every plane vector is ONE 16-byte global vector -> cp.async straight into the parity
        // planes (zero-size copy = zero padding); no staging ring, no re-layout, one round trip per unit
        const uint32_t pl = smem_u32(planes);
        const int fw = ftid >> 5, nfw = kFillWarps;
        for (int rho = fw; rho < P.rows_alloc; rho += nfw) {
          const int yh = yh0 + rho;
          for (int kap = lane; kap < Pq; kap += 32) {
            const int xh = kap - 1;
            const uint32_t d0 = pl + (uint32_t)((rho * Pq + kap) * 16);
#pragma unroll
            for (int yp = 0; yp < 2; ++yp) {
              const int y = 2 * yh + yp;
#pragma unroll
              for (int xp = 0; xp < 2; ++xp) {
                const int x = 2 * xh + xp;
                const bool ok = y >= 0 && y < H && x >= 0 && x < W;
                const __half* g = img + ((size_t)(ok ? y : 0) * W + (ok ? x : 0)) * kC24;
                const int nbytes = ok ? 16 : 0;
#pragma unroll
                for (int gq = 0; gq < 3; ++gq)
                  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                               ::"r"(d0 + (uint32_t)((gq * 4 + yp * 2 + xp) * P.plane_bytes)), "l"(g + 8 * gq), "r"(nbytes) : "memory");
              }
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();                                       // plane writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_FULL_PL + buf]);
        continue;
      }

      auto group_rows = [&](int g, int& rho0, int& rho1, int& yc0, int& yc1) {
        rho0 = g * P.crh; rho1 = min(rho0 + P.crh, P.rows_alloc);
        yc0 = max(0, 2 * (yh0 + rho0)); yc1 = min(H, 2 * (yh0 + rho1));
      };
      for (int g = 0; g < n_groups; ++g) {
        int rho0, rho1, yc0, yc1;
        group_rows(g, rho0, rho1, yc0, yc1);
        const bool have = yc1 > yc0;
        const unsigned short* stage = reinterpret_cast<const unsigned short*>(stage_base);   // (never read when !have, but addressed)
        if (P.use_bulk) {
          if (have) {
            const uint32_t sb = wcnt % NS;
            PROF_WAIT(pf_w1, mbar_wait(&bars[BAR_STAGE + sb], (wcnt / NS) & 1));
            stage = reinterpret_cast<const unsigned short*>(stage_base + (size_t)sb * P.stage_bytes);
            ++wcnt;
          }
        } else if (have) {                                         // rows not 16-byte granular: plain loads
          unsigned short* st = reinterpret_cast<unsigned short*>(stage_base);
          const unsigned short* src = reinterpret_cast<const unsigned short*>(img) + (size_t)yc0 * rowC;
          for (int i = ftid; i < (yc1 - yc0) * rowC; i += nfill) st[i] = src[i];
          named_bar_sync(1, nfill);
          stage = st;
        }
        // one work item = one plane position x one row parity (short items: the re-layout is latency bound, not issue bound)
        int npos = (rho1 - rho0) * Pq;
#ifdef CONV_TC_PROF
        if (pf_dbg & 2) npos = 0;
#endif
        for (int item = ftid; item < 2 * npos; item += nfill) {
          const int yp = item >= npos ? 1 : 0;
          const int pos = item - yp * npos;
          const int rl = pos / Pq, kap = pos - rl * Pq;
          const int rho = rho0 + rl, yh = yh0 + rho, xh = kap - 1;
          uint8_t* dst = planes + ((size_t)rho * Pq + kap) * 16;
          {
            // all shared-memory loads of the item first (branch free, clamped addresses), then the selects, then all stores:
            // stores into the planes may alias the staging rows as far as the compiler knows, so interleaving them with the
            // loads would serialise every load behind the previous store
            const int y = 2 * yh + yp;
            const bool yok = have && y >= yc0 && y < yc1;
            const unsigned short* rowp = stage + (yok ? (y - yc0) : 0) * rowC;
            uint4 v[2][kMaxG8];
            const bool ok0 = yok && xh >= 0 && 2 * xh < W, ok1 = yok && xh >= 0 && 2 * xh + 1 < W;
            const unsigned short* px0 = rowp + (ok0 ? 2 * xh : 0) * C;
            const unsigned short* px1 = rowp + (ok1 ? 2 * xh + 1 : 0) * C;
#pragma unroll
            for (int gq = 0; gq < kMaxG8; ++gq)
              if (gq < P.G8) {                                     // (uniform)
                uint4 t0 = load8h_any(px0 + 8 * gq), t1 = load8h_any(px1 + 8 * gq);
                if (gq == P.G8 - 1) {
                  t0.x &= last_mask.x; t0.y &= last_mask.y; t0.z &= last_mask.z; t0.w &= last_mask.w;
                  t1.x &= last_mask.x; t1.y &= last_mask.y; t1.z &= last_mask.z; t1.w &= last_mask.w;
                }
                v[0][gq] = ok0 ? t0 : make_uint4(0, 0, 0, 0);
                v[1][gq] = ok1 ? t1 : make_uint4(0, 0, 0, 0);
              }
            constexpr int NPIX = KS + 1, NE = KS * R, NJ = (NE + 7) / 8;
            uint32_t pix[NPIX][R > 0 ? R : 1];
            if constexpr (R > 0) {
              // remainder channels: the KS taps along x of the R channels share 16-byte rows, one per output column.
              // Output columns 2*xh (dx = 0) and 2*xh + 1 (dx = 1) read pixels 2*xh - PAD .. 2*xh + 1 + PAD.
              const bool cok = yok && xh >= 0 && xh < PW;
#pragma unroll
              for (int i = 0; i < NPIX; ++i) {
                const int x = 2 * xh - PAD + i;
                const bool ok = cok && x >= 0 && x < W;
#pragma unroll
                for (int c = 0; c < R; ++c) {
                  const uint32_t t = (uint32_t)rowp[(ok ? x : 0) * C + 8 * P.G8 + c];
                  pix[i][c] = ok ? t : 0u;
                }
              }
            }
#pragma unroll
            for (int xp = 0; xp < 2; ++xp)
#pragma unroll
              for (int gq = 0; gq < kMaxG8; ++gq)
                if (gq < P.G8) *reinterpret_cast<uint4*>(dst + (size_t)(gq * 4 + yp * 2 + xp) * P.plane_bytes) = v[xp][gq];
            if constexpr (R > 0) {
#pragma unroll
              for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                  uint32_t h[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const int E = 8 * j + e;                       // compile time: kx = E / R, channel = E % R
                    h[e] = E < NE ? pix[dx + E / R][E % R] : 0u;
                  }
                  *reinterpret_cast<uint4*>(dst + (size_t)(4 * P.G8 + j * 4 + yp * 2 + dx) * P.plane_bytes) =
                      make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                }
            }
          }
        }
        named_bar_sync(1, nfill);                                  // every fill thread is done with this staging buffer
        if (P.use_bulk && have && ftid == 0) issue_next();         // ... which the producer cursor refills right away
      }
      fence_proxy_async();                                         // generic-proxy plane writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_FULL_PL + buf]);
    }
#ifdef CONV_TC_PROF
    if (warp == kEpiWarps) { PROF_PUT(4, clock64() - pf_start); PROF_PUT(5, pf_w0); PROF_PUT(6, pf_w1); PROF_PUT(11, it); }
#endif
  } else {
    // =========================================================================== MMA warp (one elected lane issues)
    const uint32_t idesc = make_idesc(N);
    const uint32_t hi = (128u >> 4) | (1u << 14);                  // SBO = 128 bytes, descriptor version 1, no swizzle
    const uint32_t b_lo0 = ((smem_u32(bsm) & 0x3FFFFu) >> 4) | ((uint32_t)N << 16);   // LBO = N * 16 bytes
    uint32_t it = 0, tc = 0;
    for (UnitIter ui(P); ui.next(); ++it) {
      const int t0 = ui.t0, t1 = ui.t1;
      const int py_first = (128 * t0) / Pq;
      const uint32_t buf = it & 1;
      PROF_WAIT(pf_w0, mbar_wait(&bars[BAR_FULL_PL + buf], (it >> 1) & 1));
      tc_fence_after();
      const uint32_t pl16 = (smem_u32(planes_base + (size_t)buf * P.unit_bytes) & 0x3FFFFu) >> 4;
      for (int t = t0; t < t1; ++t, ++tc) {
        const uint32_t ab = tc & 1;
        PROF_WAIT(pf_w1, mbar_wait(&bars[BAR_EMPTY_ACC + ab], ((tc >> 1) & 1) ^ 1)); // the epilogue drained these accumulators
        tc_fence_after();
        const uint32_t a_add = pl16 + (uint32_t)(128 * t - py_first * Pq);
        // K step outermost, the four pool positions (independent accumulators) innermost.  (Round 5 measured that consecutive MMAs
        // into the SAME accumulator run at the full rate too - profiles/umma_rate_nacc_r5.txt - so the order is a matter of taste.)
        for (int i = 0; i < P.n_pairs; ++i) {
          const uint64_t bdesc = ((uint64_t)hi << 32) | (uint64_t)(b_lo0 + (uint32_t)(i * 2 * N));
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((ab * 4 + a) * N);
            const uint64_t adesc = ((uint64_t)hi << 32) | (uint64_t)(P.adesc_lo[a][i] + a_add);
#ifdef CONV_TC_PROF
            if (pf_dbg & 1) continue;
#endif
            if (elect_one()) umma_f16(d_tmem, adesc, bdesc, idesc, i > 0 ? 1u : 0u);
          }
        }
        if (elect_one()) umma_commit(&bars[BAR_FULL_ACC + ab]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bars[BAR_EMPTY_PL + buf]);     // all MMAs that read this plane buffer have completed
      __syncwarp();
    }
    PROF_PUT(1, clock64() - pf_start); PROF_PUT(2, pf_w0); PROF_PUT(3, pf_w1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kFillWarps) tmem_dealloc(tmem_base, 512);
#ifdef CONV_TC_PROF
  if (tid == 0) g_prof[blockIdx.x][9] = (unsigned long long)(clock64() - pf_k0);
#endif
}

#ifdef CONV_TC_PROF
extern "C" __attribute__((visibility("default"))) int cpp_debug_conv_tc_skip(int mask) {
  return (int)cudaMemcpyToSymbol(g_dbg_skip, &mask, sizeof(int));
}
extern "C" __attribute__((visibility("default"))) int cpp_debug_conv_tc_prof(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_prof, sizeof(unsigned long long) * 160 * 12);
}
#endif

// ------------------------------------------------------------------------------------------ host: plan
static int build_plan(int nets, int B, int H, int W, int C, int KS, FwdPlan* P, int dgrad = 0) {
  CPP_REQUIRE(KS == 5 || KS == 3, "conv_tc: kernel size %d", KS);
  CPP_REQUIRE(nets >= 1 && nets <= kMaxNets, "conv_tc: %d sibling networks", nets);
  const int PAD = KS / 2;
  CPP_REQUIRE(H >= 2 * PAD && W >= 2 * PAD && H >= 2 && W >= 2, "conv_tc: input %dx%d too small", H, W);
  P->B = B; P->H = H; P->W = W; P->C = C; P->KS = KS; P->PAD = PAD;
  P->PH = dgrad ? (H + 1) / 2 : H / 2; P->PW = dgrad ? (W + 1) / 2 : W / 2;      // dgrad covers every input pixel, VALID pooling drops odd edges
  P->dgrad = dgrad;
  P->Pq = (W + 1) / 2 + 1;      // parity-plane pitch: ceil(W/2) pixels + one zero column shared by neighbouring rows
  P->nets = nets; P->N = (int)round_up(nets * kPieces * CO, 16);
  CPP_REQUIRE(8 * P->N <= 512, "conv_tc: N=%d does not fit two sets of TMEM accumulators", P->N);
  const int rem = C % 8;
  if (rem == 1 || rem == 2) { P->G8 = C / 8; P->R = rem; P->nR = (KS * rem + 7) / 8; }
  else { P->G8 = (C + 7) / 8; P->R = 0; P->nR = 0; }
  CPP_REQUIRE(P->G8 <= kMaxG8, "conv_tc: %d input channels (at most %d groups of 8)", C, kMaxG8);
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
  P->use_bulk = (row_bytes % 16 == 0) && (((size_t)H * W * C * 2) % 16 == 0);
  // shared-memory budget: two staging buffers of up to 16 KB first (few, long TMA copies), then the largest unit
  // (fewer halo rows re-staged per tile) whose two plane buffers still fit
  int stage_budget = 16384;
  auto size_unit = [&](int tpu) {
    P->tiles_per_unit = tpu;
    P->rows_alloc = (P->Pq - 1 + 128 * tpu - 1 + 2 * P->Pq + 2) / P->Pq + 1;
    P->plane_bytes = P->rows_alloc * P->Pq * 16;
    P->unit_bytes = P->n_planes * P->plane_bytes;
    P->crh = std::max(1, std::min(P->rows_alloc, stage_budget / (2 * row_bytes)));
    P->crh = (int)ceil_div(P->rows_alloc, ceil_div(P->rows_alloc, P->crh));      // equal groups
    P->stage_bytes = (int)round_up((int64_t)2 * P->crh * row_bytes, 16);
    return (int)smem_layout(*P).total;
  };
  // pipelining needs several units per CTA (fill of unit u+1 overlaps the MMAs of unit u): with few tiles per CTA, small
  // units beat the halo rows they re-stage
  const int64_t tiles_per_cta = ceil_div((int64_t)B * P->tiles_per_image, sm_budget());
  const int want_tpu = (int)std::max<int64_t>(1, std::min<int64_t>(8, tiles_per_cta / 6));
  int best_tpu = 0;
  for (int ns = kMaxStageSlots; ns >= 2 && best_tpu == 0; ns -= 2) {      // deepest staging ring that leaves room for a unit
    P->stage_slots = ns;
    for (int tpu = std::min(P->tiles_per_image, want_tpu); tpu >= 1; --tpu)
      if (size_unit(tpu) <= kSmemLimit) { best_tpu = tpu; break; }
  }
  if (best_tpu > 0 && P->crh < P->rows_alloc) {
    // room left: stage a whole unit's rows per TMA copy, so the re-layout of a unit is ONE pass of the fill warps with one
    // barrier (it is latency bound: 2 passes of half the rows cost twice as much as one pass of all of them)
    const int ns_keep = P->stage_slots;
    stage_budget = 2 * P->rows_alloc * row_bytes;
    bool fits = false;
    for (int ns = ns_keep; ns >= 2 && !fits; ns -= 2) {
      P->stage_slots = ns;
      fits = size_unit(best_tpu) <= kSmemLimit;
    }
    if (!fits) { stage_budget = 16384; P->stage_slots = ns_keep; }
  }
  CPP_REQUIRE(best_tpu > 0, "conv_tc: %dx%dx%d does not fit shared memory", H, W, C);
  P->units_per_image = (int)ceil_div(P->tiles_per_image, best_tpu);
  P->smem_bytes = size_unit(best_tpu);
  P->n_units = B * P->units_per_image;

  // low words of the A descriptors for plane buffer 0 at pooled offset 0: (address >> 4) | (LBO >> 4) << 16
  for (int a = 0; a < 4; ++a) {
    const int dy = a >> 1, dx = a & 1;
    for (int i = 0; i < P->n_pairs; ++i) {
      int64_t addr[2];
      for (int h = 0; h < 2; ++h) {
        const Slab sl = P->slab[i][h];
        const int ty = dy + sl.ky - PAD, yp = ty & 1, oy = (ty >> 1) + 1;
        if (sl.kind == 0) {
          const int tx = dx + sl.kx - PAD, xp = tx & 1, ox = (tx >> 1) + 1;
          addr[h] = (int64_t)(sl.set * 4 + yp * 2 + xp) * P->plane_bytes + (int64_t)(oy * P->Pq + ox) * 16;
        } else if (sl.kind == 1) {
          addr[h] = (int64_t)(4 * P->G8 + sl.set * 4 + yp * 2 + dx) * P->plane_bytes + (int64_t)(oy * P->Pq + 1) * 16;
        } else {
          addr[h] = addr[0] + 16;      // zero weights: any initialised shared memory will do
        }
      }
      const int64_t lbo = addr[1] - addr[0];
      CPP_REQUIRE(lbo > 0 && lbo < (1 << 18) && addr[0] % 16 == 0, "conv_tc: bad slab pair");
      P->adesc_lo[a][i] = (uint32_t)(addr[0] >> 4) | ((uint32_t)(lbo >> 4) << 16);
    }
  }
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
  const int64_t b = (int64_t)bpack_bytes(P) + (int64_t)round_up((int64_t)(ncls * ncls * nets * CO + 4) * 4, 256);
  // the same scratch serves the row-sweep kernel (conv_row_tc.cu) when the layer is a piece-layout one it covers
  return C == kC24 && nets == 1 ? std::max(b, tcr::scratch_bytes(H, W, KS)) : b;
}

template <int KS, int R>
static int launch_main(const FwdPlan& P, int grid, cudaStream_t s) {
  auto k = conv_fwd_tc_kernel<KS, R>;
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    configured = true;
  }
  k<<<grid, kThreads2, P.smem_bytes, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

int launch_conv_fwd_tc(const void* x_f16, const int32_t* rows, const float* mean_inv, int nets,
                       const float* const* w, const float* const* bias, int B, int H, int W, int C, int KS,
                       float* const* pooled, uint8_t* const* amax, void* scratch, cudaStream_t s,
                       int x_is_pieces, __half* const* pooled_hl, int phase) {
  if (B <= 0) return CPP_OK;
  if (x_is_pieces == 2 && nets == 1 && rows == nullptr && mean_inv == nullptr && C == kC24 && tcr::supported(H, W, KS)) {
    CPP_REQUIRE(w[0] && bias[0], "conv_tc: null pointer for network 0");
    return tcr::launch(x_f16, w[0], bias[0], B, H, W, KS, 0, phase == kPhasePrep ? nullptr : pooled[0], phase == kPhasePrep ? nullptr : amax[0],
                       phase != kPhasePrep && pooled_hl ? pooled_hl[0] : nullptr, nullptr, nullptr, scratch, s, phase);
  }
  FwdPlan P{};
  CPP_TRY(build_plan(nets, B, H, W, C, KS, &P));
  CPP_REQUIRE(x_is_pieces == 0 || (mean_inv == nullptr && C == (x_is_pieces == 2 ? kC24 : 2 * CO)), "conv_tc: piece input has 20 or 24 channels and no whitening");
  P.Cw = x_is_pieces ? CO : C;
  P.in_layout = x_is_pieces;
  CPP_REQUIRE(((uintptr_t)x_f16 & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "conv_tc: unaligned buffers");
  const bool prep_only = phase == kPhasePrep;
  P.x = reinterpret_cast<const __half*>(x_f16); P.rows = rows;
  P.bpack = reinterpret_cast<const __half*>(scratch);
  P.corr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(scratch) + bpack_bytes(P));
  PrepArgs A{};
  for (int n = 0; n < nets; ++n) {
    CPP_REQUIRE(w[n] && bias[n] && (prep_only || (pooled[n] && amax[n])), "conv_tc: null pointer for network %d", n);
    A.w[n] = w[n]; A.bias[n] = bias[n];
    if (!prep_only) { P.pooled[n] = pooled[n]; P.amax[n] = amax[n]; P.pooled_hl[n] = pooled_hl ? pooled_hl[n] : nullptr; }
  }
  A.mean_inv = mean_inv;
  A.bpack = const_cast<__half*>(P.bpack); A.corr = const_cast<float*>(P.corr);
  if (phase != kPhaseMain) {
    conv_tc_prep_kernel<<<64, 256, 0, s>>>(P, A);
    CPP_CHECK_LAUNCH();
  }
  if (prep_only) return CPP_OK;
  const int grid = (int)std::min<int64_t>((int64_t)B * P.tiles_per_image, sm_budget());   // persistent: one CTA per SM (it owns all 512 TMEM columns)
  if (KS == 5) {
    if (P.R == 0) return launch_main<5, 0>(P, grid, s);
    if (P.R == 1) return launch_main<5, 1>(P, grid, s);
    return launch_main<5, 2>(P, grid, s);
  }
  if (P.R == 0) return launch_main<3, 0>(P, grid, s);
  if (P.R == 1) return launch_main<3, 1>(P, grid, s);
  return launch_main<3, 2>(P, grid, s);
}

int launch_conv_dgrad_tc(const void* dy_pieces, const float* inv_scale, const float* w, int B, int H, int W, int KS, float* dx,
                         void* scratch, cudaStream_t s, float* out_absmax, int phase) {
  if (B <= 0) return CPP_OK;
  if (out_absmax != nullptr && phase != kPhasePrep) CPP_CHECK_CUDA(cudaMemsetAsync(out_absmax, 0, sizeof(float), s));
  if (tcr::supported(H, W, KS))
    return tcr::launch(dy_pieces, w, nullptr, B, H, W, KS, 1, dx, nullptr, nullptr, inv_scale, out_absmax, scratch, s, phase);
  FwdPlan P{};
  CPP_TRY(build_plan(1, B, H, W, kC24, KS, &P, 1));
  P.out_absmax = out_absmax;
  CPP_REQUIRE(((uintptr_t)dy_pieces & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "conv_tc: unaligned buffers");
  P.Cw = CO; P.in_layout = 2;
  P.x = reinterpret_cast<const __half*>(dy_pieces); P.rows = nullptr;
  P.bpack = reinterpret_cast<const __half*>(scratch);
  P.corr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(scratch) + bpack_bytes(P));
  P.out_scale = inv_scale;
  P.pooled[0] = dx; P.amax[0] = nullptr; P.pooled_hl[0] = nullptr;
  PrepArgs A{};
  A.w[0] = w; A.bias[0] = nullptr; A.mean_inv = nullptr;
  A.bpack = const_cast<__half*>(P.bpack); A.corr = const_cast<float*>(P.corr);
  if (phase != kPhaseMain) {
    conv_tc_prep_kernel<<<64, 256, 0, s>>>(P, A);
    CPP_CHECK_LAUNCH();
  }
  if (phase == kPhasePrep) return CPP_OK;
  const int grid = (int)std::min<int64_t>((int64_t)B * P.tiles_per_image, sm_budget());
  return KS == 5 ? launch_main<5, 0>(P, grid, s) : launch_main<3, 0>(P, grid, s);
}

// one thread per (pixel, 16-byte vector of the 24-channel piece layout): coalesced stores; the window's gradient and
// arg-max bytes are re-read by its 4 pixels x 3 vectors from L1
__global__ void __launch_bounds__(256) unpool_split_kernel(const float* __restrict__ g, const uint8_t* __restrict__ amax, int B, int H, int W,
                                                           const float* __restrict__ gmax, float* __restrict__ inv_scale,
                                                           uint4* __restrict__ out) {
  const int PH = H / 2, PW = W / 2;
  float scale = 1.f;
  {
    const float mx = gmax[0];
    if (mx > 0.f && isfinite(mx)) { int e; frexpf(mx, &e); scale = ldexpf(1.f, 15 - e); }
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) inv_scale[0] = 1.f / scale;
  if (i >= (int64_t)B * H * W * 3) return;
  const int v = (int)(i % 3);
  const int64_t pix = i / 3;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
  const int py = y >> 1, px = x >> 1, pa = ((y & 1) << 1) | (x & 1);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (py < PH && px < PW) {
    const size_t idx = (((size_t)b * PH + py) * PW + px) * CO;
    auto piece = [&](int o, bool lo) -> uint32_t {
      const float val = amax[idx + o] == pa ? g[idx + o] * scale : 0.f;
      const __half h = __float2half_rn(val);
      return (uint32_t)__half_as_ushort(lo ? __float2half_rn(val - __half2float(h)) : h);
    };
    if (v < 2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = piece(2 * k, v == 1) | (piece(2 * k + 1, v == 1) << 16);
    } else {
      w[0] = piece(8, false) | (piece(9, false) << 16);
      w[1] = piece(8, true) | (piece(9, true) << 16);
    }
  }
  out[i] = make_uint4(w[0], w[1], w[2], w[3]);                          // constant channel and padding stay 0 in a gradient
}

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  __shared__ float sh[8];
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(g[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]);
    atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));             // non-negative floats order like ints: deterministic
  }
}

int launch_absmax(const float* g, int64_t n, float* gmax, cudaStream_t s) {
  CPP_CHECK_CUDA(cudaMemsetAsync(gmax, 0, sizeof(float), s));
  if (n > 0) {
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, ceil_div(n, 256 * 8)));
    absmax_kernel<<<blocks, 256, 0, s>>>(g, n, gmax);
    CPP_CHECK_LAUNCH();
  }
  return CPP_OK;
}

bool conv_dgrad_fused_supported(int H, int W, int KS) { return tcr::fused_unpool(H, W, KS); }

int launch_conv_dgrad_tc_fused(const float* d_pooled, const uint8_t* amax, float* gmax, float* inv_scale, int gmax_ready, const float* w,
                               int B, int H, int W, int KS, float* dx, void* scratch, cudaStream_t s, float* out_absmax, int phase) {
  if (B <= 0) return CPP_OK;
  CPP_REQUIRE(conv_dgrad_fused_supported(H, W, KS), "conv_tc: fused un-pool input gradient does not cover %dx%d k%d", H, W, KS);
  if (phase != kPhasePrep) {
    if (!gmax_ready) CPP_TRY(launch_absmax(d_pooled, (int64_t)B * (H / 2) * (W / 2) * CO, gmax, s));
    if (out_absmax != nullptr) CPP_CHECK_CUDA(cudaMemsetAsync(out_absmax, 0, sizeof(float), s));
  }
  return tcr::launch(nullptr, w, nullptr, B, H, W, KS, 1, dx, nullptr, nullptr, nullptr, out_absmax, scratch, s, phase, d_pooled, amax, gmax,
                     inv_scale);
}

int launch_unpool_split(const float* d_pooled, const uint8_t* amax, int B, int H, int W, float* gmax, float* inv_scale,
                        __half* dy_pieces, cudaStream_t s, int gmax_ready) {
  if (B <= 0) return CPP_OK;
  if (!gmax_ready) CPP_TRY(launch_absmax(d_pooled, (int64_t)B * (H / 2) * (W / 2) * CO, gmax, s));
  const int64_t items = (int64_t)B * H * W * 3;
  unpool_split_kernel<<<(unsigned)ceil_div(items, 256), 256, 0, s>>>(d_pooled, amax, B, H, W, gmax, inv_scale,
                                                                    reinterpret_cast<uint4*>(dy_pieces));
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace tc
}  // namespace cpp
