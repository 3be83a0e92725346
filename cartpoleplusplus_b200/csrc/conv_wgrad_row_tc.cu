// Weight / bias gradient of the 10 -> 10 channel conv layers (conv2 / conv3; base_network.py:111-127 backwards) on tcgen05, row-sweep
// formulation - round 5.  Same contract as conv_wgrad_mma.cu with dup = 2 (input = the 24-channel fp16 piece copy of the layer below):
//
//   G[(t, n), (kx, c)] = sum over pixels q of  dY[y = r - PAD + t][q][n] * X[r][q + kx][c]          (ky = KS - 1 - t)
//
//  * The reduction runs over PIXELS, so both UMMA operands are MN-major (a 16-byte vector = 8 consecutive rows of the operand at
//    ONE pixel, consecutive pixels = consecutive vectors, LBO = 128 between the K8 halves, one uniform stride SBO between 8-row
//    blocks; conventions as in conv_wgrad_tc.cu).
//  * The pixel index q walks a STRIP of `ipt` whole image rows including their zero halo - exactly the strips of the forward kernel
//    (conv_row_tc.cu), brought in by the same TMA tensor-map boxes, one per 8-channel group g.  With the strip as the B operand,
//    tap kx is the strip shifted by kx vectors: N block kx starts 16 bytes after block kx - 1 (SBO = 16), so the window is never
//    materialised.  One accumulator per channel group (3 x [128 x 48] fp32).
//  * The A operand is the window of KS un-pooled output-gradient rows around input row r, built on the fly from d(pooled) + the
//    arg-max side band (power-of-two scaled hi / lo fp16 pieces, 24 columns per row) into a ring [row][column block][q] whose
//    first rows are mirrored behind the last, so a window is always ONE descriptor: M = KS x 3 blocks (120 of 128 rows for 5x5).
//  * Per (3 images x 1 input row): 3 groups x 7 K16 steps, M 128 N 48 - 44 clk each, A re-read per group.  Accumulators are
//    flushed into fp32 registers of 12 epilogue warps every <= 32 K steps (the tensor-core accumulator truncates), two TMEM sets.
//  * Deterministic: every CTA owns a contiguous range of input rows and writes one partial; one finalize kernel sums the partials
//    in a fixed order and maps them to dW (sum over the hi / lo pieces of X and of dY) and db (constant-one channel, centre tap).
#include <cuda.h>
#include <algorithm>
#include <string.h>
#include "conv_tc.cuh"
#include "conv_wgrad_row_tc.cuh"
#include "umma.cuh"

namespace cpp {
namespace wgr {

using namespace umma;
using tc::kC24;

constexpr int CO = kConvCout;
constexpr int kRing = 8;                      // dY rows / input strips in flight (one ring position per step)
constexpr int kDSlots = 13;                   // dY ring + KS - 1 mirror slots + the slot the junk M block of a window may reach
constexpr int kNCols = 48;                    // MMA N: 5 (3) tap columns x 8 channels of a group, padded to a multiple of 16
constexpr int kEpiWarps = 12;                 // 3 channel groups x 4 TMEM lane quarters
constexpr int kDyGroup = 4;                   // dY producers: thread q of a group owns strip position q ...
constexpr int kDyGroups = 2;                  // ... and group j builds the rows of the steps n = j mod 2, one own step prefetched
constexpr int kDyWarps = kDyGroup * kDyGroups;
constexpr int kXWarp = kEpiWarps + kDyWarps, kMmaWarp = kXWarp + 1;
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr int kMinSmem = 116 * 1024;          // more than half an SM: never two TMEM-owning CTAs on one SM
enum { BAR_DFULL = 0, BAR_DONE = kRing, BAR_XFULL = 2 * kRing, BAR_XFREE = 3 * kRing, BAR_ACC_FULL = 4 * kRing, BAR_ACC_EMPTY = 4 * kRing + 2,
       BAR_COUNT = 4 * kRing + 4 };

struct Plan {
  const float* gp; const uint8_t* amax; const float* gmax;
  float* dw; float* db; float* partials;
  int B, H, W, KS, PAD, E, ipt, n_tiles, Kq, ksteps;
  int plane_bytes, xstage_bytes, dblk_bytes, dslot_bytes;
  int flush_steps, grid, part_floats;
  uint32_t off_dy, off_bars, off_tmem, smem_bytes;
};

__device__ __forceinline__ float scale_for(float mx) {      // same rule as conv_wgrad_mma.cu / conv_tc.cu
  if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
  int e;
  frexpf(mx, &e);
  return ldexpf(1.f, 15 - e);
}
__device__ __forceinline__ uint32_t idesc_mn(int N) {      // fp32 D, fp16 A/B, both MN-major, M = 128
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16; descriptors handed over as 32-bit halves (the high words never change)
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

// The CTA's input rows [g0, g1) of all tiles (tile-major) as segments (tile, rows [ra, rb)) that never cross a tile.  A segment is
// processed as steps t = 0 .. rb - ra + 2 PAD - 1: step t stages dY row ra - PAD + t (zero outside the image) and, once t >= 2 PAD,
// input row ra + t - 2 PAD, whose instructions read the KS dY rows staged last.  Every role walks the same sequence.
struct Seq {
  long long g, g1;
  int H;
  int tile, ra, rb;
  __device__ Seq(const Plan& P) : H(P.H), tile(0), ra(0), rb(0) {
    const long long total = (long long)P.n_tiles * P.H;
    g = total * blockIdx.x / gridDim.x; g1 = total * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ bool next() {
    if (g >= g1) return false;
    tile = (int)(g / H); ra = (int)(g - (long long)tile * H);
    rb = (int)min((long long)H, ra + (g1 - g));
    g += rb - ra;
    return true;
  }
};

// Pipeline diagnosis build (nvcc -DWGROW_PROF, scripts/prof_conv_row.py wgrad): cycles per CTA.  Slots: 0 prologue, 1 MMA total,
// 2 MMA waits DFULL, 3 MMA waits XFULL, 4 MMA waits ACC_EMPTY, 5 dY group 0 total, 6 dY group 0 waits DONE, 7 epilogue total,
// 8 epilogue waits ACC_FULL, 9 whole kernel, 10 steps, 11 X steps
#ifdef WGROW_PROF
__device__ unsigned long long g_wrprof[160][12];
#define WRPROF_WAIT(acc, stmt) { const long long pf_a = clock64(); stmt; acc += (unsigned long long)(clock64() - pf_a); }
#define WRPROF_PUT(slot, v) { if (lane == 0) g_wrprof[blockIdx.x][slot] = (unsigned long long)(v); }
extern "C" __attribute__((visibility("default"))) int cpp_debug_wgrad_row_prof(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_wrprof, sizeof(unsigned long long) * 160 * 12);
}
#else
#define WRPROF_WAIT(acc, stmt) { stmt; }
#define WRPROF_PUT(slot, v)
#endif

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_row_kernel(const __grid_constant__ Plan P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int PAD = KS / 2;
  const long long pf_k0 = clock64();
  unsigned long long pf_w0 = 0, pf_w1 = 0, pf_w2 = 0;
  uint8_t* xst = smem;
  uint8_t* dyr = smem + P.off_dy;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.off_tmem);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int H = P.H, W = P.W, E = P.E;

  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&bars[BAR_DFULL + i], kDyGroup); mbar_init(&bars[BAR_DONE + i], 1);
      mbar_init(&bars[BAR_XFULL + i], 1); mbar_init(&bars[BAR_XFREE + i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[BAR_ACC_FULL + i], 1); mbar_init(&bars[BAR_ACC_EMPTY + i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  // strips and ring start at zero: the tail of a strip past its last image, the mirror / slack slots and every vector a junk row or
  // column of an instruction may touch must be finite (0 x NaN would poison real rows)
  for (uint32_t i = tid; i < P.off_bars / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long pf_start = clock64();
  if (tid == 0) { WRPROF_PUT(0, pf_start - pf_k0); }

  if (warp < kEpiWarps) {
    // =========================================================================== epilogue: accumulate the flushed periods in fp32
    const int quarter = warp & 3, ge = warp >> 2;
    float acc[kNCols];
#pragma unroll
    for (int c = 0; c < kNCols; ++c) acc[c] = 0.f;
    long long rows = 0;
    { Seq sq(P); rows = sq.g1 - sq.g; }
    const int periods = (int)((rows + P.flush_steps - 1) / P.flush_steps);
    for (int p = 0; p < periods; ++p) {
      const uint32_t set = (uint32_t)p & 1u;
      WRPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_ACC_FULL + set], ((uint32_t)p >> 1) & 1u));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + set * (uint32_t)(3 * kNCols) + (uint32_t)(ge * kNCols);
#pragma unroll
      for (int c0 = 0; c0 < kNCols; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c0 + c] += __uint_as_float(r[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_ACC_EMPTY + set]);
    }
    float* part = P.partials + (size_t)blockIdx.x * P.part_floats + (size_t)ge * kNCols * 128 + 32 * quarter + lane;
#pragma unroll
    for (int c = 0; c < kNCols; ++c) part[(size_t)c * 128] = acc[c];
    if (warp == 0) { WRPROF_PUT(7, clock64() - pf_start); WRPROF_PUT(8, pf_w0); }
  } else if (warp < kEpiWarps + kDyWarps) {
    // =========================================================================== dY producers
    // strip position q = (image i, column x): the un-pooled output gradient of row y at (i, x) as hi / lo pieces in three vectors
    // [hi0..7 | lo0..7 | hi8 hi9 lo8 lo9 0 0 0 0]; halo positions, rows outside the image and images past the batch are zero.
    // Group `grp` owns the steps n = grp mod kDyGroups; the global loads of its next own step are in flight while it converts and
    // stores the current one.
    const int pw = warp - kEpiWarps, grp = pw / kDyGroup, q = 32 * (pw % kDyGroup) + lane;
    const float scale = scale_for(P.gmax[0]);
    const int img_l = q / E, xe = q - img_l * E, PH = H / 2, PW = W / 2;
    const bool col_ok = img_l < P.ipt && xe < W;
    struct Raw { float2 g[5]; uint32_t am[5]; uint32_t pa; bool ok; };
    // flat walk over the steps of this CTA
    Seq sq(P);
    bool live = sq.next();
    int t = 0;
    uint32_t n = 0;
    auto advance = [&]() {                                         // to the next step of the sequence
      ++n;
      if (++t >= sq.rb - sq.ra + 2 * PAD) { live = sq.next(); t = 0; }
    };
    auto fetch = [&](Raw& w) {                                      // the loads of the current step
      const int b = sq.tile * P.ipt + img_l, y = sq.ra - PAD + t;
      w.ok = live && col_ok && b < P.B && y >= 0 && y < H;
      if (w.ok) {
        const size_t idx = (((size_t)b * PH + (y >> 1)) * PW + (xe >> 1)) * CO;
        w.pa = (uint32_t)(((y & 1) << 1) | (xe & 1));
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          w.g[k] = *reinterpret_cast<const float2*>(P.gp + idx + 2 * k);
          w.am[k] = *reinterpret_cast<const uint16_t*>(P.amax + idx + 2 * k);
        }
      }
    };
    for (int i = 0; i < grp && live; ++i) advance();               // first own step
    Raw cur, nxt;
    fetch(cur);
    while (live) {
      const uint32_t n_cur = n;
      for (int i = 0; i < kDyGroups && live; ++i) advance();       // next own step: its loads go out now
      fetch(nxt);
      uint32_t hi[5] = {0u, 0u, 0u, 0u, 0u}, lo[5] = {0u, 0u, 0u, 0u, 0u};
      if (cur.ok) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float v0 = (cur.am[k] & 0xffu) == cur.pa ? cur.g[k].x * scale : 0.f, v1 = (cur.am[k] >> 8) == cur.pa ? cur.g[k].y * scale : 0.f;
          const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
          hi[k] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          lo[k] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
      }
      // the row this slot held was read by the steps up to (its own step) + 2 PAD
      if (n_cur + 2 * PAD >= (uint32_t)kRing) {
        const uint32_t sdone = n_cur + 2 * PAD - kRing;
        WRPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_DONE + sdone % kRing], (sdone / kRing) & 1u));
      }
      const uint32_t slot = n_cur % kRing;
      if (q < P.Kq) {
        uint8_t* dst = dyr + (size_t)slot * P.dslot_bytes + (size_t)q * 16;
        const uint4 v0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), v1 = make_uint4(lo[0], lo[1], lo[2], lo[3]), v2 = make_uint4(hi[4], lo[4], 0u, 0u);
        *reinterpret_cast<uint4*>(dst) = v0;
        *reinterpret_cast<uint4*>(dst + P.dblk_bytes) = v1;
        *reinterpret_cast<uint4*>(dst + 2 * P.dblk_bytes) = v2;
        if (slot < (uint32_t)(KS - 1)) {                             // mirror behind the ring: a window never wraps
          uint8_t* d2 = dst + (size_t)kRing * P.dslot_bytes;
          *reinterpret_cast<uint4*>(d2) = v0;
          *reinterpret_cast<uint4*>(d2 + P.dblk_bytes) = v1;
          *reinterpret_cast<uint4*>(d2 + 2 * P.dblk_bytes) = v2;
        }
      }
      fence_proxy_async();                                           // generic-proxy ring writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_DFULL + slot]);
      cur = nxt;
    }
    if (pw == 0) { WRPROF_PUT(5, clock64() - pf_start); WRPROF_PUT(6, pf_w0); }
  } else if (warp == kXWarp) {
    // =========================================================================== X strips: TMA tensor-map boxes (one lane)
    if (lane == 0) {
      const uint32_t bytes = 3u * (uint32_t)(P.ipt * E * 16);
      uint32_t nx = 0;
      for (Seq sq(P); sq.next();)
        for (int r = sq.ra; r < sq.rb; ++r, ++nx) {
          const uint32_t s = nx % kRing;
          mbar_wait(&bars[BAR_XFREE + s], ((nx / kRing) & 1u) ^ 1u);
          mbar_expect_tx(&bars[BAR_XFULL + s], bytes);
          uint8_t* dst = xst + (size_t)s * P.xstage_bytes;
#pragma unroll
          for (int g = 0; g < 3; ++g) tma_load_4d(dst + (size_t)g * P.plane_bytes, &tmap, 8 * g, -PAD, r, sq.tile * P.ipt, &bars[BAR_XFULL + s]);
        }
    }
  } else {
    // =========================================================================== MMA warp (one elected lane issues)
    const uint32_t x_u = smem_u32(xst), d_u = smem_u32(dyr);
    const uint32_t idesc = idesc_mn(kNCols);
    // descriptor halves: LBO = 128 bytes between the K8 halves; SBO = one column block of a dY row (A) / ONE 16-byte vector (B: tap
    // column kx + 1 is the strip shifted by one pixel); version 1, no swizzle
    const uint32_t lbo16 = (128u >> 4) << 16;
    const uint32_t hi_a = ((uint32_t)P.dblk_bytes >> 4) | (1u << 14), hi_b = (16u >> 4) | (1u << 14);
    const int ksteps = P.ksteps;
    uint32_t n = 0, nx = 0, period = 0;
    int in_period = 0;
    for (Seq sq(P); sq.next();) {
      const int nsteps = sq.rb - sq.ra + 2 * PAD;
      for (int t = 0; t < nsteps; ++t, ++n) {
        WRPROF_WAIT(pf_w0, mbar_wait(&bars[BAR_DFULL + n % kRing], (n / kRing) & 1u));
        if (t >= 2 * PAD) {
          const uint32_t xs = nx % kRing;
          WRPROF_WAIT(pf_w1, mbar_wait(&bars[BAR_XFULL + xs], (nx / kRing) & 1u));
          const uint32_t set = period & 1u;
          if (in_period == 0) WRPROF_WAIT(pf_w2, mbar_wait(&bars[BAR_ACC_EMPTY + set], ((period >> 1) & 1u) ^ 1u));
          tc_fence_after();
          // the KS dY rows staged last: ring slots n - 2 PAD .. n (mod kRing), contiguous thanks to the mirror slots
          const uint32_t a_base = d_u + ((n - 2 * PAD) % kRing) * (uint32_t)P.dslot_bytes;
          const uint32_t b_base = x_u + xs * (uint32_t)P.xstage_bytes;
          if (elect_one()) {
            const uint32_t a_lo0 = ((a_base & 0x3FFFFu) >> 4) | lbo16;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              const uint32_t d_tmem = tmem_base + set * (uint32_t)(3 * kNCols) + (uint32_t)(g * kNCols);
              const uint32_t b_lo0 = (((b_base + (uint32_t)g * (uint32_t)P.plane_bytes) & 0x3FFFFu) >> 4) | lbo16;
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k)                      // 16 pixels further = 256 bytes = 16 in the address field
                umma_lohi(d_tmem, a_lo0 + 16u * (uint32_t)k, hi_a, b_lo0 + 16u * (uint32_t)k, hi_b, idesc, (in_period > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&bars[BAR_XFREE + xs]);
          }
          ++nx;
          if (++in_period == P.flush_steps) {
            if (elect_one()) umma_commit(&bars[BAR_ACC_FULL + set]);
            ++period; in_period = 0;
          }
        }
        if (elect_one()) umma_commit(&bars[BAR_DONE + n % kRing]);    // (a warm-up step has no instruction: done at once)
        __syncwarp();
      }
    }
    if (in_period > 0 && elect_one()) umma_commit(&bars[BAR_ACC_FULL + (period & 1u)]);
    __syncwarp();
    WRPROF_PUT(1, clock64() - pf_start); WRPROF_PUT(2, pf_w0); WRPROF_PUT(3, pf_w1); WRPROF_PUT(4, pf_w2); WRPROF_PUT(10, n); WRPROF_PUT(11, nx);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
  if (tid == 0) { WRPROF_PUT(9, clock64() - pf_k0); }
  (void)pf_w0; (void)pf_w1; (void)pf_w2; (void)pf_k0; (void)pf_start;
}

// one warp per output: dw[ky][kx][c][o] = 1/scale * sum over CTAs and over the hi / lo pieces of X (rows of the strip planes) and of dY
// (lanes); db[o] from the constant-one channel at the centre tap.  Fixed order -> deterministic.
__global__ void __launch_bounds__(256) wgrad_row_finalize_kernel(const __grid_constant__ Plan P) {
  const int KS = P.KS, nw = KS * KS * CO * CO;
  const int out = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (out >= nw + CO) return;
  int ky, kx, c, o;
  const bool is_b = out >= nw;
  if (is_b) { o = out - nw; ky = P.PAD; kx = P.PAD; c = 0; }
  else { o = out % CO; c = (out / CO) % CO; kx = (out / (CO * CO)) % KS; ky = out / (CO * CO * KS); }
  const int t = KS - 1 - ky;
  int idx[4];
#pragma unroll
  for (int xp = 0; xp < 2; ++xp) {
    const int ch = is_b ? tc::kC24One : (xp ? tc::c24_lo(c) : tc::c24_hi(c));
    const int g = ch >> 3, col = kx * 8 + (ch & 7);
#pragma unroll
    for (int yp = 0; yp < 2; ++yp) {
      const int code = yp ? tc::c24_lo(o) : tc::c24_hi(o);
      idx[xp * 2 + yp] = (g * kNCols + col) * 128 + (t * 3 + (code >> 3)) * 8 + (code & 7);
    }
  }
  const int ncomb = is_b ? 2 : 4;                                  // the bias reads ONE constant channel
  float s = 0.f;
  for (int cta = lane; cta < P.grid; cta += 32) {
    const float* part = P.partials + (size_t)cta * P.part_floats;
    float v = 0.f;
    for (int k = 0; k < ncomb; ++k) v += part[idx[k]];
    s += v;
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
  if (lane == 0) {
    const float v = s / scale_for(P.gmax[0]);
    if (is_b) P.db[o] = v; else P.dw[out] = v;
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn();
bool shape_ok(int H, int W, int KS) {
  if (encode_fn() == nullptr) return false;      // the strips come in through TMA tensor maps only: without the driver entry point the mma.sync kernel takes the layer
  return (KS == 5 || KS == 3) && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && W + KS - 1 <= 128;
}

static int build_plan(int B, int H, int W, int KS, Plan* P) {
  CPP_REQUIRE(shape_ok(H, W, KS), "wgrad_row: %dx%d k%d not supported", H, W, KS);
  P->B = B; P->H = H; P->W = W; P->KS = KS; P->PAD = KS / 2;
  P->E = W + KS - 1;
  P->ipt = 128 / P->E;
  P->n_tiles = (int)ceil_div(B, P->ipt);
  P->Kq = (int)round_up(P->ipt * P->E, 16);
  P->ksteps = P->Kq / 16;
  P->plane_bytes = (int)round_up((int64_t)(P->Kq + 8) * 16, 128);
  P->xstage_bytes = 3 * P->plane_bytes;
  P->dblk_bytes = P->Kq * 16;
  P->dslot_bytes = 3 * P->dblk_bytes;
  P->flush_steps = std::max(1, std::min(g_wgrad_flush_steps, 32) / P->ksteps);
  P->part_floats = 3 * kNCols * 128;
  uint32_t off = (uint32_t)kRing * (uint32_t)P->xstage_bytes;
  P->off_dy = off; off += (uint32_t)kDSlots * (uint32_t)P->dslot_bytes;
  off = (off + 15u) & ~15u;
  P->off_bars = off; off += BAR_COUNT * 8;
  P->off_tmem = off; off += 16;
  P->smem_bytes = std::max<uint32_t>(off, kMinSmem);
  CPP_REQUIRE(P->smem_bytes <= 200u * 1024u, "wgrad_row: shared memory");
  return CPP_OK;
}

int64_t scratch_bytes(int H, int W, int KS) {
  if (!shape_ok(H, W, KS)) return 0;
  return (int64_t)round_up((int64_t)kNumSMs * 3 * kNCols * 128 * 4, 256);
}

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int launch(const void* x_pieces, const float* d_pooled, const uint8_t* amax, const float* gmax, int B, int H, int W, int KS, float* dw, float* db,
           void* scratch, cudaStream_t s) {
  if (B <= 0) return CPP_OK;
  Plan P{};
  CPP_TRY(build_plan(B, H, W, KS, &P));
  CPP_REQUIRE(x_pieces && d_pooled && amax && gmax && dw && db && scratch, "wgrad_row: null pointer");
  CPP_REQUIRE(((uintptr_t)x_pieces & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "wgrad_row: unaligned buffers");
  P.gp = d_pooled; P.amax = amax; P.gmax = gmax; P.dw = dw; P.db = db;
  P.partials = reinterpret_cast<float*>(scratch);
  P.grid = (int)std::min<int64_t>((int64_t)P.n_tiles * H, sm_budget());
  EncodeTiledFn enc = encode_fn();
  CPP_REQUIRE(enc != nullptr, "wgrad_row: cuTensorMapEncodeTiled not available from this driver");
  CUtensorMap tm;
  const cuuint64_t gdim[4] = {(cuuint64_t)kC24, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)kC24 * 2, (cuuint64_t)W * kC24 * 2, (cuuint64_t)H * W * kC24 * 2};
  const cuuint32_t box[4] = {8u, (cuuint32_t)P.E, 1u, (cuuint32_t)P.ipt};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x_pieces), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CPP_REQUIRE(cr == CUDA_SUCCESS, "wgrad_row: cuTensorMapEncodeTiled failed (%d) for %dx%dx%d", (int)cr, B, H, W);
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_row_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CPP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_row_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (KS == 5) conv_wgrad_row_kernel<5><<<P.grid, kThreads, P.smem_bytes, s>>>(P, tm);
  else conv_wgrad_row_kernel<3><<<P.grid, kThreads, P.smem_bytes, s>>>(P, tm);
  CPP_CHECK_LAUNCH();
  wgrad_row_finalize_kernel<<<(unsigned)ceil_div(KS * KS * CO * CO + CO, 8), 256, 0, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace wgr
}  // namespace cpp
