// Row-sweep tcgen05 convolution for the 10 -> 10 channel layers of the reference trunk (conv2 / conv3 forward and their input
// gradients; base_network.py:111-127) - round 5.  Same contract as conv_fwd_tc_kernel with in_layout = 2 (conv_tc.cu), other GEMM
// view:
//  * The parity-plane kernel puts ONE tap per K step: with only 10 filters the instruction is bound by the 4 KB shared-memory
//    read of its A operand (40 clk) and every output pixel pays 25 taps x 3 channel groups of them.
//  * Here an instruction covers one INPUT row r and one tap column kx, and the FIVE ky taps sit side by side along N:
//    column block t holds w[ky = KS-1-t] and lands in the accumulator of output row y = r - PAD + t.  Output rows live in a ring
//    of 12 TMEM slots (32 columns each: 10 filters x hi/lo weight piece, padded) plus 4 shadow slots behind it, so the N = KS*32
//    columns of the instruction are KS consecutive slots (a window that starts in the last ring slots runs on into the shadows
//    instead of wrapping: always ONE instruction; the epilogue adds the two halves): the tensor core accumulates every output row
//    over its KS input rows IN TENSOR MEMORY, nothing is summed in registers.  A read of A is now amortised over 160 columns:
//    the instruction is tensor-rate bound (N/2 clk), and a row of 3 images costs 8 instructions (5 kx x 3 channel groups = 15
//    K8 halves) instead of 3 x 38 x 4 / 2.
//  * M = 128 lanes = a strip of `ipt` whole image rows including their zero halo: entry e of image i is pixel e - PAD, lane
//    (i, x) reads entries x .. x + KS - 1, so tap kx is the SAME strip shifted by kx 16-byte entries - a no-swizzle K-major
//    descriptor pointing into the strip.  The strip of a row is three TMA tensor-map boxes (one per 8-channel group of the
//    24-channel piece layout) {8 channels, W + KS - 1 pixels from x = -PAD, 1 row, ipt images}: the SAME padding and the
//    images past the end of the batch are the TMA unit's out-of-bounds zero fill - no fill warps, no re-layout instructions.
//  * The epilogue drains two finished output rows at a time (tcgen05.ld), zeroes their slots (tcgen05.st) for the rows that
//    reuse them, and does bias + ReLU + 2x2 max-pool (x partner = the neighbouring lane: one shuffle) + arg-max byte + fp16
//    piece copy, or (input-gradient mode) writes the dense fp32 rows and max|dx|.
//  * Input-gradient mode builds its strips itself: eight producer warps un-pool d(pooled) through the arg-max side band, scale by a
//    power of two and split into hi / lo pieces straight into shared memory (no piece tensor, no separate pass).
//  * Work = the pooled rows of all tiles, cut into one contiguous range per CTA (segments never cross a tile; a segment that starts
//    inside an image re-reads PAD input rows).  Roles: 8 epilogue warps, 8 producer warps (TMA: one lane), 1 MMA warp that walks
//    its loop with warp-uniform values only - one elected lane issues, the descriptors stay in uniform registers.
// Deterministic: every output accumulates its taps in a fixed order; no atomics on sums.
// Measured, role cycle accounting and the versions that did not work: DESIGN.md 3.3, profiles/r5/.
#include <cuda.h>
#include <algorithm>
#include <string.h>
#include "conv_tc.cuh"
#include "conv_row_tc.cuh"
#include "umma.cuh"

namespace cpp {
namespace tcr {

using namespace umma;
using tc::kC24;
using tc::c24_weight_channel;

constexpr int CO = kConvCout;
constexpr int kSlotCols = 32;                 // TMEM columns of one output row: [hi piece x 10 | lo piece x 10 | 12 zero]
constexpr int kRing = 12;                     // output rows in the TMEM ring ...
constexpr int kPhysRing = 16;                 // ... plus 4 shadow slots behind it (16 x 32 = all 512 columns): a slot window that starts in
                                              // the last ring slots runs on into the shadows instead of wrapping around
constexpr int kPairs = kRing / 2;             // rows are drained two at a time (one pooled row)
constexpr int kShadowPairs = (kPhysRing - kRing) / 2;
constexpr int kStages = 8;                    // input-row strips in flight
constexpr int kAhead = 6;                     // cp.async route: row strips whose copies are in flight before the oldest is awaited
constexpr int kMaxSteps = 8;                  // K16 instructions per input row (KS = 5: 15 K8 halves)
constexpr int kEpiWarps = 8;                  // two groups of four (one warp per TMEM lane quarter); group e drains pairs of parity e
constexpr int kProdWarps = 8;                 // producer warps: TMA / cp.async use the first; the fused un-pool route all of them (two groups
                                              // of four, group q builds the strips of rows n = q mod 2)
constexpr int kThreads = 32 * (kEpiWarps + kProdWarps + 1);   // + MMA warp
constexpr int kMinSmem = 116 * 1024;          // more than half an SM: two CTAs of concurrent chains must never share an SM - the
                                              // second would sit in tcgen05.alloc until the first is done while other SMs idle

struct RowPlan {
  const __half* bpack;      // [step][half][N = KS*32][8] fp16, canonical K-major B operand
  const float* tab;         // [10] bias, [10] = 2^-S
  float* out;               // forward: pooled fp32 [B][H/2][W/2][10]; dgrad: dense fp32 [B][H][W][10]
  uint8_t* amax;
  __half* out_hl;           // optional piece copy of the pooled output (24-channel layout)
  const float* out_scale;   // dgrad
  float* out_absmax;        // dgrad, optional
  int B, H, W, KS, PAD, E, ipt, n_tiles, dgrad;
  int n_steps, N;
  int use_tma;              // 1: input strips by TMA tensor-map boxes (OOB zero fill, default); 0: 16-byte cp.async into pre-zeroed strips
  const __half* x;          // fp16 pieces [B][H][W][24] (cp.async route)
  // fused un-pool route (dgrad): the strips are built from the pooled-output gradient and the arg-max side band, no piece tensor
  int unpool;
  const float* gp;          // d(pooled) fp32 [B][H/2][W/2][10]
  const uint8_t* gamax;     // arg-max side band of the layer's forward pass
  const float* gmax;        // device float: max |gp| (power-of-two scaling of the fp16 pieces)
  float* inv_scale_out;     // optional: receives 1 / scale
  int plane_bytes, stage_bytes;
  int8_t kx[kMaxSteps][2], g[kMaxSteps][2];   // (tap column, channel group) of the two K8 halves of a step; g = -1: zero half
  uint32_t a_lo[kMaxSteps];                   // low word of the A descriptor of a step relative to the stage base: (offset >> 4) | (LBO >> 4) << 16
};

struct PrepArgs { const float* w; const float* bias; __half* bpack; float* tab; };

// activations whose fp16 piece pair saturated since the last reset (added to tc::piece_overflow_count)
__device__ unsigned int g_piece_overflow_row = 0;
int piece_overflow_count(int reset, unsigned int* out) {
  CPP_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_piece_overflow_row, sizeof(unsigned int)));
  if (reset) { const unsigned int z = 0; CPP_CHECK_CUDA(cudaMemcpyToSymbol(g_piece_overflow_row, &z, sizeof(unsigned int))); }
  return CPP_OK;
}

// ------------------------------------------------------------------------------------------ weight packing
__global__ void __launch_bounds__(256) conv_row_prep_kernel(const __grid_constant__ RowPlan P, const __grid_constant__ PrepArgs A) {
  __shared__ float red[256];
  __shared__ float s_scale;
  const int tid = threadIdx.x, KS = P.KS, nw = KS * KS * CO * CO;
  float mx = 0.f;
  for (int i = tid; i < nw; i += blockDim.x) mx = fmaxf(mx, fabsf(A.w[i]));
  red[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]); __syncthreads(); }
  if (tid == 0) {
    int e = 0;
    const float m = red[0];
    if (m > 0.f && isfinite(m)) frexpf(m, &e); else e = 15;      // m < 2^e
    s_scale = ldexpf(1.f, 15 - e);
    if (blockIdx.x == 0) A.tab[CO] = ldexpf(1.f, e - 15);
  }
  __syncthreads();
  const float scale = s_scale;
  const int N = P.N, total = P.n_steps * 2 * N * 8;
  for (int idx = blockIdx.x * blockDim.x + tid; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7, n = (idx >> 3) % N, h = ((idx >> 3) / N) & 1, i = (idx >> 3) / (2 * N);
    const int g = P.g[i][h], kx = P.kx[i][h], t = n / kSlotCols, j = n % kSlotCols;
    __half out = __float2half_rn(0.f);
    if (g >= 0 && j < 2 * CO) {
      const int piece = j / CO, o = j % CO, ky = KS - 1 - t, chw = c24_weight_channel(8 * g + e);
      if (chw >= 0) {
        const float wv = P.dgrad ? A.w[(((KS - 1 - ky) * KS + (KS - 1 - kx)) * CO + o) * CO + chw]      // flipped taps, in/out swapped
                                 : A.w[((ky * KS + kx) * CO + chw) * CO + o];
        const float v = wv * scale;
        const __half hi = __float2half_rn(v);
        out = piece == 0 ? hi : __float2half_rn(v - __half2float(hi));
      }
    }
    A.bpack[idx] = out;
  }
  if (blockIdx.x == 0 && tid < CO) A.tab[tid] = A.bias ? A.bias[tid] : 0.f;
}

// (kx, channel group) of K8 half h of step i; g = -1: the zero-weight half in front of the odd tap left over (build_plan)
__host__ __device__ inline void step_half(int KS, int i, int h, int* kx, int* g) {
  if (i < KS) { *kx = i; *g = h; return; }
  const int k0 = 2 * (i - KS);
  if (k0 + 1 < KS) { *kx = k0 + h; *g = 2; }
  else if (h == 0) { *kx = k0 - 1; *g = -1; }
  else { *kx = k0; *g = 2; }
}

// The weight packs of MANY layer passes in one launch (blockIdx.y = job): the fused steps prepare every conv2 / conv3 forward and
// input-gradient pass of a step at t = 0 (Net::prep_trunk_tc under tcr::prep_batch_begin / prep_batch_flush).
constexpr int kMaxPrepJobs = 16;
struct PrepJob { const float* w; const float* bias; __half* bpack; float* tab; int KS, dgrad; };
struct PrepBatch { int n; PrepJob job[kMaxPrepJobs]; };

__global__ void __launch_bounds__(256) conv_row_prep_batch_kernel(const __grid_constant__ PrepBatch Bt) {
  __shared__ float red[256];
  __shared__ float s_scale;
  const PrepJob& J = Bt.job[blockIdx.y];
  const int tid = threadIdx.x, KS = J.KS, nw = KS * KS * CO * CO;
  float mx = 0.f;
  for (int i = tid; i < nw; i += blockDim.x) mx = fmaxf(mx, fabsf(J.w[i]));
  red[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]); __syncthreads(); }
  if (tid == 0) {
    int e = 0;
    const float m = red[0];
    if (m > 0.f && isfinite(m)) frexpf(m, &e); else e = 15;
    s_scale = ldexpf(1.f, 15 - e);
    if (blockIdx.x == 0) J.tab[CO] = ldexpf(1.f, e - 15);
  }
  __syncthreads();
  const float scale = s_scale;
  const int N = KS * kSlotCols, n_steps = KS == 5 ? 8 : 5, total = n_steps * 2 * N * 8;
  for (int idx = blockIdx.x * blockDim.x + tid; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7, n = (idx >> 3) % N, h = ((idx >> 3) / N) & 1, i = (idx >> 3) / (2 * N);
    int kx, g;
    step_half(KS, i, h, &kx, &g);
    const int t = n / kSlotCols, j = n % kSlotCols;
    __half out = __float2half_rn(0.f);
    if (g >= 0 && j < 2 * CO) {
      const int piece = j / CO, o = j % CO, ky = KS - 1 - t, chw = c24_weight_channel(8 * g + e);
      if (chw >= 0) {
        const float wv = J.dgrad ? J.w[(((KS - 1 - ky) * KS + (KS - 1 - kx)) * CO + o) * CO + chw]
                                 : J.w[((ky * KS + kx) * CO + chw) * CO + o];
        const float v = wv * scale;
        const __half hi = __float2half_rn(v);
        out = piece == 0 ? hi : __float2half_rn(v - __half2float(hi));
      }
    }
    J.bpack[idx] = out;
  }
  if (blockIdx.x == 0 && tid < CO) J.tab[tid] = J.bias ? J.bias[tid] : 0.f;
}

// Pipeline diagnosis build (nvcc -DCONV_ROW_PROF, scripts/prof_conv_row.py): cycles per CTA.  Slots: 0 prologue, 1 MMA total,
// 2 MMA waits FULL_A, 3 MMA waits ACC_FREE, 4 producer total, 5 producer waits EMPTY_A, 6 producer waits copies, 7 epilogue
// total (group 0 quarter 0), 8 epilogue waits ACC_FULL, 9 whole kernel, 10 rows, 11 tiles
#ifdef CONV_ROW_PROF
__device__ unsigned long long g_rprof[160][12];
#define RPROF_DECL unsigned long long pf_w0 = 0, pf_w1 = 0; const long long pf_start = clock64();
#define RPROF_WAIT(acc, stmt) { const long long pf_a = clock64(); stmt; acc += (unsigned long long)(clock64() - pf_a); }
#define RPROF_PUT(slot, v) { if (lane == 0) g_rprof[blockIdx.x][slot] = (unsigned long long)(v); }
#else
#define RPROF_DECL
#define RPROF_WAIT(acc, stmt) { stmt; }
#define RPROF_PUT(slot, v)
#endif

// ------------------------------------------------------------------------------------------ main kernel
struct SmemLayout { uint32_t bsm, stages, tab, bars, tmem, total; };
enum { BAR_FULL_A = 0, BAR_EMPTY_A = kStages, BAR_ACC_FULL = 2 * kStages, BAR_ACC_FREE = 2 * kStages + kPairs, BAR_W = 2 * kStages + 2 * kPairs,
       BAR_COUNT = 2 * kStages + 2 * kPairs + 1 };

__host__ __device__ inline SmemLayout smem_layout(const RowPlan& P) {
  SmemLayout L;
  uint32_t off = 0;
  L.stages = off; off += (uint32_t)kStages * (uint32_t)P.stage_bytes;          // 128-byte aligned TMA destinations first
  L.bsm = off; off += (uint32_t)P.n_steps * 2 * P.N * 16;
  L.tab = off; off += 64;
  L.bars = off; off += BAR_COUNT * 8;
  L.tmem = off; off += 16;
  L.total = off;
  return L;
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] += A[smem] * B[smem], kind::f16, always accumulating (the slots start at zero)
__device__ __forceinline__ void umma_acc(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 p, 1, 1;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc) : "memory");
}

__host__ __device__ inline uint32_t make_idesc(int N) {      // D fp32, A/B fp16, both K-major, M = 128
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// power of two that brings max |g| just under 2^15 (the un-pool / split pass of conv_tc.cu uses the same rule)
__device__ __forceinline__ float piece_scale(float mx) {
  float scale = 1.f;
  if (mx > 0.f && isfinite(mx)) { int e; frexpf(mx, &e); scale = ldexpf(1.f, 15 - e); }
  return scale;
}

// Work distribution: the pooled rows (output row pairs) of all tiles, tile-major, are cut into one contiguous range per CTA; a
// CTA walks its range in segments (tile, output rows [ya, yb)) that never cross a tile.  A segment that starts or ends inside an
// image re-reads PAD input rows of its neighbour - the price of an even split when there are fewer tiles than 3 per SM.
struct SegIter {
  long long g, g1;
  int HP;
  __device__ SegIter(const RowPlan& P) {
    HP = P.H / 2;
    const long long total = (long long)P.n_tiles * HP;
    g = total * blockIdx.x / gridDim.x; g1 = total * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ bool next(int& tile, int& ya, int& yb) {
    if (g >= g1) return false;
    tile = (int)(g / HP);
    const int pa = (int)(g - (long long)tile * HP), pb = (int)min((long long)HP, pa + (g1 - g));
    g += pb - pa;
    ya = 2 * pa; yb = 2 * pb;
    return true;
  }
};

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) conv_row_tc_kernel(const __grid_constant__ RowPlan P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint8_t smem[];
#ifdef CONV_ROW_PROF
  const long long pf_k0 = clock64();
#endif
  constexpr int PAD = KS / 2, NSTEPS = KS == 5 ? 8 : 5, N = KS * kSlotCols;
  const SmemLayout L = smem_layout(P);
  uint8_t* stage_base = smem + L.stages;
  uint8_t* bsm = smem + L.bsm;
  float* tab_s = reinterpret_cast<float*>(smem + L.tab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int H = P.H, W = P.W, E = P.E, HP = H / 2;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars[BAR_FULL_A + i], P.unpool ? 4 : 1); mbar_init(&bars[BAR_EMPTY_A + i], 1); }
    for (int i = 0; i < kPairs; ++i) { mbar_init(&bars[BAR_ACC_FULL + i], 1); mbar_init(&bars[BAR_ACC_FREE + i], 4); }
    mbar_init(&bars[BAR_W], 1);
    fence_mbar_init();
  }
  if (warp == kEpiWarps + kProdWarps) tmem_alloc(tmem_slot, 512);
  if (tid < CO + 1) tab_s[tid] = P.tab[tid];
  if (!P.use_tma && !P.unpool) {
    // cp.async route: only image pixels are ever copied, the SAME-padding halo entries (and the tail past the last image) stay zero
    uint4* st = reinterpret_cast<uint4*>(stage_base);
    for (int i = tid; i < kStages * P.stage_bytes / 16; i += kThreads) st[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  RPROF_DECL
#ifdef CONV_ROW_PROF
  if (tid == 0) g_rprof[blockIdx.x][0] = (unsigned long long)(pf_start - pf_k0);
#endif

  if (warp < kEpiWarps) {
    // =========================================================================== epilogue warps
    const int quarter = warp & 3, eg = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * quarter) << 16);
    // every slot (shadow slots included) starts at zero: the MMAs only ever accumulate
    for (int rp = eg; rp < kPhysRing / 2; rp += 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16_zero(lane_base + (uint32_t)(rp * 2 * kSlotCols + 16 * c));
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) for (int rp = eg; rp < kPairs; rp += 2) mbar_arrive(&bars[BAR_ACC_FREE + rp]);

    const float scale_inv = tab_s[CO];
    const int p = 32 * quarter + lane, img_l = p / E, xe = p - img_l * E;
    const bool lane_ok = img_l < P.ipt && xe < W;
    float bias[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) bias[o] = tab_s[o];
    const float osc = P.dgrad ? scale_inv * (P.unpool ? 1.f / piece_scale(P.gmax[0]) : P.out_scale[0]) : 0.f;
    float amx = 0.f;
    int gp = 0;
    int tile, ya, yb;
    for (SegIter it(P); it.next(tile, ya, yb);) {
      const int b = tile * P.ipt + img_l;
      const bool valid = lane_ok && b < P.B;
      for (int yh = ya >> 1; yh < (yb >> 1); ++yh, ++gp) {
        if ((gp & 1) != eg) continue;
        const int rp = gp % kPairs;
        RPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_ACC_FULL + rp], (uint32_t)(gp / kPairs) & 1u));
        tc_fence_after();
        uint32_t r0[20], r1[20];
        const uint32_t taddr = lane_base + (uint32_t)(rp * 2 * kSlotCols);
        tmem_ld16(taddr, r0); tmem_ld4(taddr + 16, r0 + 16);
        tmem_ld16(taddr + kSlotCols, r1); tmem_ld4(taddr + kSlotCols + 16, r1 + 16);
        if (rp < kShadowPairs) {
          // the slot windows of the input rows that started near the end of the ring ran on into the shadow slots: the rest of
          // these two output rows sits there
          uint32_t s0[20], s1[20];
          const uint32_t saddr = taddr + (uint32_t)(kRing * kSlotCols);
          tmem_ld16(saddr, s0); tmem_ld4(saddr + 16, s0 + 16);
          tmem_ld16(saddr + kSlotCols, s1); tmem_ld4(saddr + kSlotCols + 16, s1 + 16);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 20; ++j) {
            r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __uint_as_float(s0[j]));
            r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __uint_as_float(s1[j]));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_st16_zero(saddr + (uint32_t)(16 * c));
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st16_zero(taddr + (uint32_t)(16 * c));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_ACC_FREE + rp]);        // this warp's quarter of the two slots is drained and zero again

        if (P.dgrad) {
          // input-gradient mode: no bias / ReLU / pool - two dense output rows
          if (valid) {
            float* op0 = P.out + (((size_t)b * H + 2 * yh) * W + xe) * CO;
            float* op1 = op0 + (size_t)W * CO;
#pragma unroll
            for (int o = 0; o < CO; o += 2) {
              const float a0 = (__uint_as_float(r0[o]) + __uint_as_float(r0[CO + o])) * osc;
              const float a1 = (__uint_as_float(r0[o + 1]) + __uint_as_float(r0[CO + o + 1])) * osc;
              const float c0 = (__uint_as_float(r1[o]) + __uint_as_float(r1[CO + o])) * osc;
              const float c1 = (__uint_as_float(r1[o + 1]) + __uint_as_float(r1[CO + o + 1])) * osc;
              *reinterpret_cast<float2*>(op0 + o) = make_float2(a0, a1);
              *reinterpret_cast<float2*>(op1 + o) = make_float2(c0, c1);
              amx = fmaxf(amx, fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fmaxf(fabsf(c0), fabsf(c1))));
            }
          }
          continue;
        }
        // forward: the 2x2 window of pooled position (yh, xe / 2) = rows 2yh, 2yh+1 (this lane's two slots) x columns xe, xe ^ 1
        // (this lane and its neighbour).  Both lanes of a pair compute the same winner; the even one stores fp32 + arg-max, the
        // odd one the piece copy.
        const bool odd = (xe & 1) != 0;
        float best[CO];
        int arg[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const float m0 = fmaf(__uint_as_float(r0[o]) + __uint_as_float(r0[CO + o]), scale_inv, bias[o]);
          const float m1 = fmaf(__uint_as_float(r1[o]) + __uint_as_float(r1[CO + o]), scale_inv, bias[o]);
          const float q0 = __shfl_xor_sync(0xffffffffu, m0, 1), q1 = __shfl_xor_sync(0xffffffffu, m1, 1);
          const float v00 = odd ? q0 : m0, v01 = odd ? m0 : q0, v10 = odd ? q1 : m1, v11 = odd ? m1 : q1;
          float bv = v00; int ba = 0;
          if (v01 > bv) { bv = v01; ba = 1; }
          if (v10 > bv) { bv = v10; ba = 2; }
          if (v11 > bv) { bv = v11; ba = 3; }
          best[o] = bv; arg[o] = ba;
        }
        if (valid) {
          const size_t pbase = (((size_t)b * HP + yh) * (W / 2) + (xe >> 1)) * CO;
          if (!odd) {
            float* op = P.out + pbase;
            uint8_t* ap = P.amax + pbase;
#pragma unroll
            for (int o = 0; o < CO; o += 2) {
              *reinterpret_cast<float2*>(op + o) = make_float2(fmaxf(best[o], 0.f), fmaxf(best[o + 1], 0.f));
              *reinterpret_cast<uint16_t*>(ap + o) = (uint16_t)((best[o] > 0.f ? arg[o] : 4) | ((best[o + 1] > 0.f ? arg[o + 1] : 4) << 8));
            }
          } else if (P.out_hl != nullptr) {
            uint32_t hi[CO / 2], lo[CO / 2];
#pragma unroll
            for (int o = 0; o < CO; o += 2) {
              const float v0 = fmaxf(best[o], 0.f), v1 = fmaxf(best[o + 1], 0.f);
              const __half h0 = __float2half_rn(fminf(v0, 65504.f)), h1 = __float2half_rn(fminf(v1, 65504.f));
              const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
              if (fmaxf(v0, v1) > 131000.f) atomicAdd(&g_piece_overflow_row, 1u);   // never silent: cpp_piece_overflow_count
              hi[o >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lo[o >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            uint4* hp = reinterpret_cast<uint4*>(P.out_hl + (pbase / CO) * kC24);
            hp[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            hp[1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            hp[2] = make_uint4(hi[4], lo[4], 0x00003C00u, 0u);       // hi8 hi9 | lo8 lo9 | 1.0 0 | 0 0
          }
        }
      }
    }
    if (P.dgrad && P.out_absmax != nullptr) {                    // max is order independent: an atomic keeps the result deterministic
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, sft));
      if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(P.out_absmax), __float_as_int(amx));
    }
#ifdef CONV_ROW_PROF
    if (warp == 0) { RPROF_PUT(7, clock64() - pf_start); RPROF_PUT(8, pf_w0); }
#endif
  } else if (warp < kEpiWarps + kProdWarps && P.unpool) {
    // =========================================================================== producer warps, fused un-pool route (dgrad)
    // The A strips are built on the fly: entry (image, x) of input row r = the un-pooled output gradient at (r, x) - d(pooled) of the
    // window where the arg-max byte points at this position, else 0 - times a power of two, as hi / lo fp16 pieces.  Thread j of a
    // group owns strip entry j; group q takes the rows with n = q mod 2, so that the L2 latency of one row's loads hides behind
    // the other group's row.
    const int pw = warp - kEpiWarps, grp = pw >> 2, j = 32 * (pw & 3) + lane;
    if (pw == 0 && lane == 0) {
      const uint32_t wbytes = (uint32_t)(NSTEPS * 2 * N * 16);
      mbar_expect_tx(&bars[BAR_W], wbytes);
      bulk_g2s(bsm, P.bpack, wbytes, &bars[BAR_W]);
    }
    const float scale = piece_scale(P.gmax[0]);
    if (blockIdx.x == 0 && pw == 0 && lane == 0 && P.inv_scale_out != nullptr) P.inv_scale_out[0] = 1.f / scale;
    const int img_l = j / E, xe = j - img_l * E, x = xe - PAD, PW = W / 2;
    const bool col_ok = img_l < P.ipt && x >= 0 && x < W;
    const bool entry_ok = j < P.ipt * E;
    uint32_t n = 0;
    int tile, ya, yb;
    for (SegIter it(P); it.next(tile, ya, yb);) {
      const int ra = max(0, ya - PAD), rb = min(H - 1, yb - 1 + PAD);
      const int b = tile * P.ipt + img_l;
      const bool ok = col_ok && b < P.B;
      for (int r = ra; r <= rb; ++r, ++n) {
        if ((int)(n & 1u) != grp) continue;
        uint32_t hi[5] = {0u, 0u, 0u, 0u, 0u}, lo[5] = {0u, 0u, 0u, 0u, 0u};
        if (ok) {
          const size_t idx = (((size_t)b * HP + (r >> 1)) * PW + (x >> 1)) * CO;
          const uint32_t pa = (uint32_t)(((r & 1) << 1) | (x & 1));
          float2 g[5];
          uint32_t am[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            g[k] = *reinterpret_cast<const float2*>(P.gp + idx + 2 * k);
            am[k] = *reinterpret_cast<const uint16_t*>(P.gamax + idx + 2 * k);
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const float v0 = (am[k] & 0xffu) == pa ? g[k].x * scale : 0.f, v1 = (am[k] >> 8) == pa ? g[k].y * scale : 0.f;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
            hi[k] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            lo[k] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          }
        }
        const uint32_t s = n % kStages;
        RPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_EMPTY_A + s], ((n / kStages) & 1u) ^ 1u));
        if (entry_ok) {                                            // halo entries and images past the batch: zeros
          uint8_t* dst = stage_base + (size_t)s * P.stage_bytes + (size_t)j * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + P.plane_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(dst + 2 * P.plane_bytes) = make_uint4(hi[4], lo[4], 0u, 0u);    // the constant channel stays 0 in a gradient
        }
        fence_proxy_async();                                       // generic-proxy strip writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_FULL_A + s]);
      }
    }
#ifdef CONV_ROW_PROF
    if (pw == 0) { RPROF_PUT(4, clock64() - pf_start); RPROF_PUT(5, pf_w0); }
#endif
  } else if (warp == kEpiWarps) {
    // =========================================================================== producer warp
    if (lane == 0) {                                             // packed weights: one bulk copy, awaited by the MMA thread only
      const uint32_t wbytes = (uint32_t)(NSTEPS * 2 * N * 16);
      mbar_expect_tx(&bars[BAR_W], wbytes);
      bulk_g2s(bsm, P.bpack, wbytes, &bars[BAR_W]);
    }
    int tile, ya, yb;
    if (!P.use_tma) {
      // 16-byte cp.async straight into the strips: lane v -> (image, pixel, channel group), consecutive lanes read consecutive
      // 16-byte vectors of the row; kAhead rows in flight, the oldest is awaited, made visible to the tensor core, handed over
      const int per_img = 3 * W, nvec = P.ipt * per_img;
      uint32_t n = 0;
      for (SegIter it(P); it.next(tile, ya, yb);) {
        const int ra = max(0, ya - PAD), rb = min(H - 1, yb - 1 + PAD);
        for (int r = ra; r <= rb; ++r, ++n) {
          const uint32_t s = n % kStages;
          RPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_EMPTY_A + s], ((n / kStages) & 1u) ^ 1u));
          const uint32_t dst = smem_u32(stage_base + (size_t)s * P.stage_bytes);
          for (int v = lane; v < nvec; v += 32) {
            const int i = v / per_img, rem = v - i * per_img, x = rem / 3, g = rem - 3 * x, b = tile * P.ipt + i;
            if (b < P.B)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                           ::"r"(dst + (uint32_t)(g * P.plane_bytes + (i * E + PAD + x) * 16)),
                             "l"(P.x + (((size_t)b * H + r) * W + x) * kC24 + 8 * g) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (n >= (uint32_t)kAhead) {
            RPROF_WAIT(pf_w1, asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory"));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_FULL_A + (n - kAhead) % kStages]);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async();
      __syncwarp();
      if (lane == 0)
        for (uint32_t m = n > (uint32_t)kAhead ? n - kAhead : 0u; m < n; ++m) mbar_arrive(&bars[BAR_FULL_A + m % kStages]);
      RPROF_PUT(4, clock64() - pf_start); RPROF_PUT(5, pf_w0); RPROF_PUT(6, pf_w1);
    } else if (lane == 0) {
      // TMA tensor-map boxes: {8 channels of group g, E pixels from x = -PAD, 1 row, ipt images}; out-of-bounds = zero fill
      const uint32_t bytes = 3u * (uint32_t)(P.ipt * E * 16);
      uint32_t n = 0;
      for (SegIter it(P); it.next(tile, ya, yb);) {
        const int ra = max(0, ya - PAD), rb = min(H - 1, yb - 1 + PAD);
        for (int r = ra; r <= rb; ++r, ++n) {
          const uint32_t s = n % kStages;
          RPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_EMPTY_A + s], ((n / kStages) & 1u) ^ 1u));        // the MMAs that read this stage have completed
          mbar_expect_tx(&bars[BAR_FULL_A + s], bytes);
          uint8_t* dst = stage_base + (size_t)s * P.stage_bytes;
#pragma unroll
          for (int g = 0; g < 3; ++g) tma_load_4d(dst + (size_t)g * P.plane_bytes, &tmap, 8 * g, -PAD, r, tile * P.ipt, &bars[BAR_FULL_A + s]);
        }
      }
      RPROF_PUT(4, clock64() - pf_start); RPROF_PUT(5, pf_w0);
    }
  } else if (warp == kEpiWarps + kProdWarps) {
    // =========================================================================== MMA warp (one elected lane issues)
    // One instruction per (input row, K16 step): N = nt * 32 columns = the nt output rows this input row contributes to, which are
    // consecutive ring slots; windows that start in the last slots run on into the shadow slots behind the ring instead of
    // wrapping (the epilogue adds the two halves), so it is always ONE instruction.  The whole warp walks the loop (everything in
    // it is warp-uniform, so the descriptors live in uniform registers); one elected lane issues.
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);             // SBO = 128 bytes, descriptor version 1, no swizzle
    const uint32_t st16 = (smem_u32(stage_base) & 0x3FFFFu) >> 4;
    const uint32_t b_lo0 = ((smem_u32(bsm) & 0x3FFFFu) >> 4) | ((uint32_t)N << 16);   // LBO = N * 16 bytes
    const uint32_t stage16 = (uint32_t)P.stage_bytes >> 4;
    RPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_W], 0));
    uint32_t n = 0;
    int gseg = 0;                                                  // running index of this CTA's output rows (ring slot = index % kRing)
    int tile, ya, yb;
    for (SegIter it(P); it.next(tile, ya, yb); gseg += yb - ya) {
      const int ra = max(0, ya - PAD), rb = min(H - 1, yb - 1 + PAD);
      for (int r = ra; r <= rb; ++r, ++n) {
        // output rows this input row touches for the first time: their slots must have been drained and zeroed
        {
          const int y_lo = r == ra ? ya : r + PAD, y_hi = min(yb - 1, r + PAD);
          for (int y = y_lo; y <= y_hi; ++y)
            if ((y & 1) == 0) {
              const int gp = (gseg + y - ya) >> 1;
              RPROF_WAIT(pf_w1, mbar_wait(&bars[BAR_ACC_FREE + gp % kPairs], (uint32_t)(gp / kPairs) & 1u));
            }
        }
        const uint32_t s = n % kStages;
        RPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_FULL_A + s], (n / kStages) & 1u));
        tc_fence_after();
        const int t0 = max(0, ya - r + PAD), t1 = min(KS, yb - r + PAD);
        const uint32_t d_tmem = tmem_base + (uint32_t)(((gseg + r - PAD + t0 - ya) % kRing) * kSlotCols);
        const uint32_t idesc = make_idesc((t1 - t0) * kSlotCols);
        const uint32_t a_off = st16 + s * stage16, b_off = b_lo0 + (uint32_t)(t0 * kSlotCols);
        const int y_lo = r == rb ? max(ya, r - PAD) : r - PAD, y_hi = r == rb ? yb - 1 : r - PAD;
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < NSTEPS; ++i) umma_acc(d_tmem, P.a_lo[i] + a_off, b_off + (uint32_t)(i * 2 * N), desc_hi, idesc);
          umma_commit(&bars[BAR_EMPTY_A + s]);                     // the strip may be overwritten once these MMAs have read it
          // output rows that received their last tap: hand finished pairs to the epilogue
          for (int y = max(ya, y_lo); y <= y_hi; ++y)
            if (y & 1) umma_commit(&bars[BAR_ACC_FULL + ((gseg + y - ya) >> 1) % kPairs]);
        }
        __syncwarp();
      }
    }
    RPROF_PUT(1, clock64() - pf_start); RPROF_PUT(2, pf_w0); RPROF_PUT(3, pf_w1); RPROF_PUT(10, n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kProdWarps) tmem_dealloc(tmem_base, 512);
#ifdef CONV_ROW_PROF
  if (tid == 0) g_rprof[blockIdx.x][9] = (unsigned long long)(clock64() - pf_k0);
#endif
}

#ifdef CONV_ROW_PROF
extern "C" __attribute__((visibility("default"))) int cpp_debug_conv_row_prof(unsigned long long* host_out) {
  const int rc = (int)cudaMemcpyFromSymbol(host_out, g_rprof, sizeof(unsigned long long) * 160 * 12);
  static unsigned long long zeros[160 * 12];
  cudaMemcpyToSymbol(g_rprof, zeros, sizeof(zeros));
  return rc;
}
#endif

// ------------------------------------------------------------------------------------------ host
static int g_conv_row = 1;      // bit 0: route on; bit 1: input strips by 16-byte cp.async instead of TMA tensor-map boxes; bit 2: input
                                // gradient from a piece tensor written by a separate un-pool / split pass instead of the fused producer
void set_conv_row(int on) { g_conv_row = on & 7; }
int conv_row_enabled() { return g_conv_row; }

static bool g_batching = false;
static PrepBatch g_batch;
void prep_batch_begin() { g_batching = true; g_batch.n = 0; }
int prep_batch_flush(cudaStream_t s) {
  g_batching = false;
  if (g_batch.n == 0) return CPP_OK;
  conv_row_prep_batch_kernel<<<dim3(8, g_batch.n), 256, 0, s>>>(g_batch);
  CPP_CHECK_LAUNCH();
  g_batch.n = 0;
  return CPP_OK;
}

bool shape_ok(int H, int W, int KS) {
  return (KS == 5 || KS == 3) && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && W + KS - 1 <= 128;
}
bool supported(int H, int W, int KS) { return (g_conv_row & 1) && shape_ok(H, W, KS); }
bool fused_unpool(int H, int W, int KS) { return supported(H, W, KS) && !(g_conv_row & 4); }

static int build_plan(int B, int H, int W, int KS, int dgrad, RowPlan* P) {
  CPP_REQUIRE(shape_ok(H, W, KS), "conv_row: %dx%d k%d not supported", H, W, KS);
  const int PAD = KS / 2;
  P->B = B; P->H = H; P->W = W; P->KS = KS; P->PAD = PAD; P->dgrad = dgrad;
  P->E = W + KS - 1;
  P->ipt = 128 / P->E;
  P->n_tiles = (int)ceil_div(B, P->ipt);
  P->N = KS * kSlotCols;
  P->plane_bytes = (int)round_up((int64_t)(128 + KS) * 16, 128);
  P->stage_bytes = 3 * P->plane_bytes;
  // K8 halves: (kx, channel group); a step pairs two of them with the second at the higher shared-memory address.  The odd tap left
  // over gets a zero-weight half (g = -1) in FRONT of it that re-reads the previous tap's entries - every entry a valid lane reads
  // must be one the TMA unit wrote (0 x an uninitialised NaN pattern would poison the lane)
  const int ns = KS + (KS + 1) / 2;
  for (int i = 0; i < ns && i < kMaxSteps; ++i)
    for (int h = 0; h < 2; ++h) {
      int kx, g;
      step_half(KS, i, h, &kx, &g);
      P->kx[i][h] = (int8_t)kx; P->g[i][h] = (int8_t)g;
    }
  CPP_REQUIRE(ns <= kMaxSteps, "conv_row: %d K steps", ns);
  P->n_steps = ns;
  for (int i = 0; i < ns; ++i) {
    int64_t addr[2];
    for (int h = 0; h < 2; ++h) addr[h] = (int64_t)(P->g[i][h] >= 0 ? P->g[i][h] : 2) * P->plane_bytes + (int64_t)P->kx[i][h] * 16;
    const int64_t lbo = addr[1] - addr[0];
    CPP_REQUIRE(lbo > 0 && lbo < (1 << 18) && addr[0] % 16 == 0, "conv_row: bad K8 pair");
    P->a_lo[i] = (uint32_t)(addr[0] >> 4) | ((uint32_t)(lbo >> 4) << 16);
  }
  return CPP_OK;
}

static inline size_t bpack_bytes(const RowPlan& P) { return (size_t)round_up((int64_t)P.n_steps * 2 * P.N * 16, 256); }

int64_t scratch_bytes(int H, int W, int KS) {
  RowPlan P{};
  if (!shape_ok(H, W, KS) || build_plan(1, H, W, KS, 0, &P) != CPP_OK) return 0;
  return (int64_t)bpack_bytes(P) + 256;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int launch(const void* x_pieces, const float* w, const float* bias, int B, int H, int W, int KS, int dgrad, float* out, uint8_t* amax,
           __half* out_hl, const float* out_scale, float* out_absmax, void* scratch, cudaStream_t s, int phase,
           const float* unpool_gp, const uint8_t* unpool_amax, const float* unpool_gmax, float* unpool_inv_scale) {
  if (B <= 0) return CPP_OK;
  RowPlan P{};
  CPP_TRY(build_plan(B, H, W, KS, dgrad, &P));
  CPP_REQUIRE(((uintptr_t)scratch & 255) == 0, "conv_row: unaligned scratch");
  P.bpack = reinterpret_cast<const __half*>(scratch);
  P.tab = reinterpret_cast<const float*>(reinterpret_cast<const char*>(scratch) + bpack_bytes(P));
  if (phase != tc::kPhaseMain) {
    CPP_REQUIRE(w != nullptr, "conv_row: null weights");
    if (g_batching && phase == tc::kPhasePrep && g_batch.n < kMaxPrepJobs) {
      g_batch.job[g_batch.n++] = PrepJob{w, bias, const_cast<__half*>(P.bpack), const_cast<float*>(P.tab), KS, dgrad};
      return CPP_OK;
    }
    PrepArgs A{w, bias, const_cast<__half*>(P.bpack), const_cast<float*>(P.tab)};
    conv_row_prep_kernel<<<32, 256, 0, s>>>(P, A);
    CPP_CHECK_LAUNCH();
  }
  if (phase == tc::kPhasePrep) return CPP_OK;
  P.unpool = unpool_gp != nullptr ? 1 : 0;
  if (P.unpool) {
    CPP_REQUIRE(dgrad && unpool_amax != nullptr && unpool_gmax != nullptr && out != nullptr, "conv_row: fused un-pool needs dgrad mode, the arg-max band and max|g|");
    P.gp = unpool_gp; P.gamax = unpool_amax; P.gmax = unpool_gmax; P.inv_scale_out = unpool_inv_scale;
  } else {
    CPP_REQUIRE(x_pieces != nullptr && out != nullptr && (dgrad ? out_scale != nullptr : amax != nullptr), "conv_row: null pointer");
    CPP_REQUIRE(((uintptr_t)x_pieces & 15) == 0, "conv_row: unaligned input");
  }
  P.out = out; P.amax = amax; P.out_hl = out_hl; P.out_scale = out_scale; P.out_absmax = out_absmax;
  P.x = reinterpret_cast<const __half*>(x_pieces);
  // (a driver without cuTensorMapEncodeTiled: the cp.async producer carries the strips - same results)
  P.use_tma = ((g_conv_row & 2) || P.unpool || encode_fn() == nullptr) ? 0 : 1;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (P.use_tma) {
  EncodeTiledFn enc = encode_fn();
  CPP_REQUIRE(enc != nullptr, "conv_row: cuTensorMapEncodeTiled not available from this driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)kC24, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)kC24 * 2, (cuuint64_t)W * kC24 * 2, (cuuint64_t)H * W * kC24 * 2};
  const cuuint32_t box[4] = {8u, (cuuint32_t)P.E, 1u, (cuuint32_t)P.ipt};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x_pieces), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CPP_REQUIRE(cr == CUDA_SUCCESS, "conv_row: cuTensorMapEncodeTiled failed (%d) for %dx%dx%d", (int)cr, B, H, W);
  }
  const SmemLayout L = smem_layout(P);
  const int smem_bytes = std::max<int>((int)L.total, kMinSmem);
  CPP_REQUIRE(smem_bytes <= 200 * 1024 && P.n_steps == (KS == 5 ? 8 : 5), "conv_row: plan does not match the kernel");
  // persistent: one CTA per SM (it owns all 512 TMEM columns); the pooled rows are split evenly over the CTAs
  const int grid = (int)std::min<int64_t>((int64_t)P.n_tiles * (H / 2), sm_budget());
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(conv_row_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CPP_CHECK_CUDA(cudaFuncSetAttribute(conv_row_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (KS == 5) conv_row_tc_kernel<5><<<grid, kThreads, smem_bytes, s>>>(P, tm);
  else conv_row_tc_kernel<3><<<grid, kThreads, smem_bytes, s>>>(P, tm);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace tcr
}  // namespace cpp
