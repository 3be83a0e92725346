// Weight / bias gradient of one conv layer of the reference trunk (slim.conv2d + relu + max_pool2d, base_network.py:103-123,
// differentiated by tf.gradients in ddpg_cartpole.py:111,213 / naf_cartpole.py:233) on the tensor cores:
//
//   dW_n[ky][kx][c][o] = sum_{b,y,x} xhat[b][y+ky-P][x+kx-P][c] * dY_n[b][y][x][o]          db_n[o] = sum dY_n
//   dY_n[b][y][x][o]   = d_pooled_n[b][y/2][x/2][o] if (y,x) is the arg-max of its 2x2 window and the ReLU is open, else 0
//
// as ONE GEMM  G[(ky,kx,c), (n,piece,o)] = X^T . dY  whose reduction dimension is the pixel index:
//  * the sibling networks n (actor+critic on state_1; NAF value/mu/l) share X, so they sit side by side along N, and every
//    fp32 gradient enters as two fp16 pieces (hi + lo, 22 mantissa bits, scaled by a power of two taken from max|d_pooled|) so
//    that the fp16 x fp16 -> fp32 MMA stays inside the 1e-5 parity budget; the replay pixels are exact fp16 operands.
//  * the whitening xhat = (x - mean_c) inv_c with ZERO padding of xhat (base_network.py:95-99 then SAME conv) is folded out:
//    the kernel multiplies RAW pixels plus one constant-one channel (1 inside the image, 0 in the padding), giving
//    S[(ky,kx),o] = sum over the positions whose tap lies inside the image of dY, and the finalize kernel forms
//    dW = inv_c (G - mean_c S), db = S[centre tap].
//  * im2col-free: a band of input rows is staged ONCE in shared memory as 16-byte channel-group vectors per pixel
//    (plane layout [group][row][col]); the A fragment of tap (ky,kx) for 16 consecutive output pixels is then
//    ldmatrix.trans on 16 consecutive vectors of the same plane shifted by (ky,kx) - no data is replicated per tap.
//    Channels beyond a multiple of 8 (9+1 = 8 + 2, 18+1 = 16 + 3, 24+1 = 24 + 1) are packed along kx into extra planes.
//  * why mma.sync and not tcgen05 here: the reduction runs over pixels, so both operands are "MN-major" with a per-tap
//    address pattern that a single UMMA shared-memory descriptor cannot express for M = 128 rows (wgrad).
//  * deterministic: every CTA owns a contiguous range of bands and a private partial (flushed with plain fp32 adds every
//    few bands, which also bounds the length of the tensor-core accumulation chains); partials are reduced in fixed order.
#include <algorithm>
#include "conv_wgrad_mma.cuh"
#include "conv_wgrad_tc.cuh"
#include "conv_wgrad_row_tc.cuh"

namespace cpp {
namespace wg {

constexpr int CO = kConvCout;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
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

// power of two that brings max|g| just under 2^15 (1 when the tensor is all zero or not finite)
__device__ __forceinline__ float scale_for(float mx) {
  if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
  int e;
  frexpf(mx, &e);                       // mx < 2^e
  return ldexpf(1.f, 15 - e);
}

// ------------------------------------------------------------------------------------------ max |d_pooled| per network
__global__ void __launch_bounds__(256) wgrad_absmax_kernel(const __grid_constant__ Plan P) {
  __shared__ float sh[8];
  const int net = blockIdx.y;
  const float* g = P.g[net];
  const int64_t n = (int64_t)P.B * P.PH * P.PW * CO;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(g[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]);
    atomicMax(reinterpret_cast<int*>(P.gmax_own + net), __float_as_int(m));      // non-negative floats order like ints
  }
}

// ------------------------------------------------------------------------------------------ main kernel
// Shared memory: two staging buffers (x planes + dY pieces) and one raw-row buffer.  While the MMAs of band i run out of
// buffer i&1, the raw rows of band i+1 arrive through cp.async and its pooled gradients / arg-max bytes sit in registers;
// afterwards they are re-laid into buffer (i+1)&1.  Two CTAs per SM overlap one CTA's re-layout with the other's MMAs.
struct Smem { uint32_t planes, dy, raw, buf_stride, total; };
__host__ __device__ inline Smem smem_layout(const Plan& P) {
  Smem L;
  const uint32_t planes_b = (uint32_t)P.nvec * P.plane_bytes, dy_b = (uint32_t)P.band_rows * P.Wp * P.NTp * 16;
  L.planes = 0; L.dy = planes_b; L.buf_stride = planes_b + dy_b;
  L.raw = 2 * L.buf_stride;
  L.total = L.raw + (uint32_t)P.raw_bytes + 16;          // + 16: the B-fragment load of an odd tile count reads one vector past the end
  return L;
}

constexpr int kDyItems = 1;        // (2x2 window, network) items a thread owns per band (the plan shrinks the band if there are more)

template <int MT, int NT>
__global__ void __launch_bounds__(32 * kMaxWarps, (MT * NT <= 16) ? 2 : 1) conv_wgrad_mma_kernel(const __grid_constant__ Plan P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const Smem L = smem_layout(P);
  const unsigned short* raw = reinterpret_cast<const unsigned short*>(smem + L.raw);

  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int H = P.H, W = P.W, C = P.C, KS = P.KS, PAD = P.PAD, PH = P.PH, PW = P.PW;
  const int Wp = P.Wp, pitch = P.pitch, rows_in = P.rows_in, NTp = P.NTp, nets = P.nets, BR = P.band_rows;
  const int rowC = W * C;
  const int raw_mode = (((rowC * 2) & 15) == 0 && ((((size_t)H * rowC * 2) & 15) == 0)) ? 16
                     : ((((rowC * 2) & 3) == 0 && ((((size_t)H * rowC * 2) & 3) == 0)) ? 4 : 0);

  // zero both dY buffers once: the columns between nets*20 and NT*8 are never written again
  for (int b2 = 0; b2 < 2; ++b2) {
    uint32_t* d = reinterpret_cast<uint32_t*>(smem + b2 * L.buf_stride + L.dy);
    for (int i = tid; i < BR * Wp * NTp * 4; i += nthr) d[i] = 0u;
  }

  __shared__ float s_scale[kMaxNets];
  __shared__ int s_pk[8][8];                    // packed planes: entry (j, e) -> kx | channel << 8 (kx = 255: unused entry)
  if (tid < kMaxNets) s_scale[tid] = tid < nets ? scale_for(P.gmax[tid][0]) : 1.f;
  if (tid < 64) {
    const int E = tid, kx = P.R > 0 ? E / P.R : 255;
    s_pk[tid >> 3][tid & 7] = (P.R > 0 && kx < KS) ? (kx | ((8 * P.G8 + E - kx * P.R) << 8)) : 255;
  }

  // per-lane fragment address bases.  A (x4.trans): matrix id = lane / 8 -> (slab = id & 1, K half = id >> 1), row = lane % 8
  uint32_t a_base[MT];
  const uint32_t smem_u = smem_u32(smem);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int T = warp * MT + mt;
    int s = 2 * T + ((lane >> 3) & 1);
    if (s >= P.n_slabs) s = 0;                                           // padding rows: any valid address, result ignored
    a_base[mt] = smem_u + L.planes + (uint32_t)P.slab_off[s] + (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * 16);
  }
  // B (x4.trans) for the n-tile pair (2j, 2j+1): id -> (K half = id & 1, tile = 2j + (id >> 1))
  const uint32_t b_lane = smem_u + L.dy + (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * NTp * 16 + (lane >> 4) * 16);

  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;

  float* part = P.partials + (size_t)blockIdx.x * P.part_floats;
  bool first_flush = true;
  auto flush = [&]() {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float* p = part + ((size_t)((warp * MT + mt) * NT + nt) * 4) * 32 + lane;
        float old[4] = {0.f, 0.f, 0.f, 0.f};
        if (!first_flush) {                              // four independent loads in flight, then add + store
#pragma unroll
          for (int r = 0; r < 4; ++r) old[r] = __ldcg(p + r * 32);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) { __stcg(p + r * 32, old[r] + acc[mt][nt][r]); acc[mt][nt][r] = 0.f; }
      }
    first_flush = false;
  };

  // ---- staging pieces.  dY item = (window row of the band, window column, network): the 10 pooled gradients and arg-max
  // bytes of one 2x2 window, prefetched into registers one band ahead, converted to fp16 pieces once and scattered to the
  // window's four pixels.
  const int wp2 = Wp / 2, hp = BR / 2, dy_items = hp * wp2 * nets;
  float2 gq[kDyItems][5];
  unsigned short aq[kDyItems][5];
  int it_pack[kDyItems];                                               // this thread's items (net | window col << 4 | window row << 20): the same for every band
#pragma unroll
  for (int k = 0; k < kDyItems; ++k) {
    const int it = tid + k * nthr;
    it_pack[k] = (it % nets) | (((it / nets) % wp2) << 4) | ((it / (nets * wp2)) << 20);
  }
  int cur_b = 0, cur_bi = 0;                                            // (image, band-in-image) of the band being multiplied
  auto band_rows_of = [&](bool next, int& b, int& y0, int& ylo, int& yhi) {   // next: the band after the current one (no division)
    int bb = cur_b, bi = cur_bi;
    if (next && ++bi == P.bands_per_image) { bi = 0; ++bb; }
    b = bb; y0 = bi * BR;
    ylo = max(0, y0 - PAD); yhi = min(H, y0 + BR + PAD);
  };
  // (1) global -> shared / registers, asynchronous
  auto issue_loads = [&](bool next, int buf) {
    int b, y0, ylo, yhi;
    band_rows_of(next, b, y0, ylo, yhi);
    const __half* src = P.x + ((size_t)b * H + ylo) * rowC;
    const int n_bytes = (yhi - ylo) * rowC * 2;
    const uint32_t dst = smem_u + L.raw;
    if (P.dup == 2) {
      // 24-channel piece layout: every plane vector is one aligned 16-byte global vector; the zero padding is a zero-size copy
      const int nwarps = nthr >> 5;
      const uint32_t pl = smem_u + (uint32_t)buf * L.buf_stride + L.planes;
      for (int lr = warp; lr < rows_in; lr += nwarps) {
        const int y = y0 - PAD + lr;
        const bool yok = y >= 0 && y < H;
        const __half* rowg = P.x + ((size_t)b * H + (yok ? y : 0)) * rowC;
        for (int lc = lane; lc < pitch; lc += 32) {
          const int xin = lc - PAD;
          const bool ok = yok && xin >= 0 && xin < W;
          const __half* g = rowg + (ok ? xin : 0) * tc::kC24;
          const uint32_t d = pl + (uint32_t)((lr * pitch + lc) * 16);
          const int nbytes = ok ? 16 : 0;
#pragma unroll
          for (int v = 0; v < 3; ++v)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d + (uint32_t)(v * P.plane_bytes)), "l"(g + 8 * v), "r"(nbytes) : "memory");
        }
      }
    } else if (raw_mode == 16) {
      for (int i = tid * 16; i < n_bytes; i += nthr * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i), "l"(reinterpret_cast<const char*>(src) + i) : "memory");
    } else if (raw_mode == 4) {
      for (int i = tid * 4; i < n_bytes; i += nthr * 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + i), "l"(reinterpret_cast<const char*>(src) + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
    for (int k = 0; k < kDyItems; ++k) {
      const int it = tid + k * nthr;
#pragma unroll
      for (int v = 0; v < 5; ++v) { gq[k][v] = make_float2(0.f, 0.f); aq[k][v] = 0x0404; }
      if (it < dy_items) {
        const int net = it_pack[k] & 15, pxl = (it_pack[k] >> 4) & 0xffff, pyl = it_pack[k] >> 20;
        const int py = (y0 >> 1) + pyl;
        if (py < PH && pxl < PW) {
          const size_t idx = (((size_t)b * PH + py) * PW + pxl) * CO;
          const float2* gp = reinterpret_cast<const float2*>(P.g[net] + idx);
          const unsigned short* ap = reinterpret_cast<const unsigned short*>(P.amax[net] + idx);
#pragma unroll
          for (int v = 0; v < 5; ++v) { gq[k][v] = __ldg(gp + v); aq[k][v] = __ldg(ap + v); }
        }
      }
    }
  };
  // (2) registers -> dY pieces of buffer `buf`; raw rows that cp.async cannot move (2-byte granular) are copied here
  auto stage_dy = [&](bool next, int buf) {
    int b, y0, ylo, yhi;
    band_rows_of(next, b, y0, ylo, yhi);
    if (raw_mode == 0 && P.dup != 2) {
      unsigned short* d = reinterpret_cast<unsigned short*>(smem + L.raw);
      const unsigned short* s2 = reinterpret_cast<const unsigned short*>(P.x + ((size_t)b * H + ylo) * rowC);
      for (int i = tid; i < (yhi - ylo) * rowC; i += nthr) d[i] = s2[i];
    }
    __half* dys = reinterpret_cast<__half*>(smem + buf * L.buf_stride + L.dy);
#pragma unroll
    for (int k = 0; k < kDyItems; ++k) {
      const int it = tid + k * nthr;
      if (it < dy_items) {
        const int net = it_pack[k] & 15, pxl = (it_pack[k] >> 4) & 0xffff, pyl = it_pack[k] >> 20;
        const float sc = s_scale[net];
        uint32_t hi2[5], lo2[5], a0[5], a1[5];                         // 10 filters as 5 half2 words + their arg-max bytes
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          const float g0 = gq[k][v].x * sc, g1 = gq[k][v].y * sc;
          const __half h0 = __float2half_rn(g0), h1 = __float2half_rn(g1);
          const __half l0 = __float2half_rn(g0 - __half2float(h0)), l1 = __float2half_rn(g1 - __half2float(h1));
          hi2[v] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          lo2[v] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          a0[v] = aq[k][v] & 0xffu; a1[v] = aq[k][v] >> 8;
        }
#pragma unroll
        for (int pa = 0; pa < 4; ++pa) {
          uint32_t w[10];                                              // [hi(10) | lo(10)] halves of this pixel and network
#pragma unroll
          for (int v = 0; v < 5; ++v) {
            const uint32_t m = (a0[v] == (uint32_t)pa ? 0x0000ffffu : 0u) | (a1[v] == (uint32_t)pa ? 0xffff0000u : 0u);
            w[v] = hi2[v] & m; w[5 + v] = lo2[v] & m;
          }
          uint2* d = reinterpret_cast<uint2*>(dys + ((size_t)((2 * pyl + (pa >> 1)) * Wp + 2 * pxl + (pa & 1)) * NTp) * 8 + net * 2 * CO);
#pragma unroll
          for (int v = 0; v < 5; ++v) d[v] = make_uint2(w[2 * v], w[2 * v + 1]);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  };
  // (3) raw rows -> planes of buffer `buf`: 16-byte channel-group vectors per pixel (zero padding; constant-one channel at C).
  // One warp per (plane, input row), lanes along the columns: no per-vector index arithmetic.
  auto stage_planes = [&](bool next, int buf) {
    int b, y0, ylo, yhi;
    band_rows_of(next, b, y0, ylo, yhi);
    uint8_t* planes = smem + buf * L.buf_stride + L.planes;
    const uint32_t ONE = 0x3C00u;                                          // fp16 1.0
    const int nwarps = nthr >> 5;
    const bool fast_r2 = P.R == 2 && KS == 5 && C == 8 * P.G8 + 1;          // c3: 9 pixels channels + 1
    for (int item = warp; item < P.nvec * rows_in; item += nwarps) {
      const int v = item / rows_in, lr = item - v * rows_in;
      const int y = y0 - PAD + lr;
      uint4* dst = reinterpret_cast<uint4*>(planes + (size_t)v * P.plane_bytes + (size_t)lr * pitch * 16);
      if (y < ylo || y >= yhi) {
        for (int lc = lane; lc < pitch; lc += 32) dst[lc] = make_uint4(0, 0, 0, 0);
        continue;
      }
      const unsigned short* rowp = raw + (size_t)(y - ylo) * rowC;
      if (v < P.G8) {
        const int nch = min(8, C - 8 * v), one_at = C - 8 * v;            // one_at in [0, 8): the constant-one channel sits in this group
        for (int lc = lane; lc < pitch; lc += 32) {
          const int xin = lc - PAD;
          uint4 val = make_uint4(0, 0, 0, 0);
          if (xin >= 0 && xin < W) {
            val = load8h(rowp + xin * C + 8 * v, nch);
            if (one_at >= 0 && one_at < 8) {
              const uint32_t o1 = ONE << ((one_at & 1) * 16);
              if ((one_at >> 1) == 0) val.x |= o1; else if ((one_at >> 1) == 1) val.y |= o1; else if ((one_at >> 1) == 2) val.z |= o1; else val.w |= o1;
            }
          }
          dst[lc] = val;
        }
      } else if (fast_r2) {
        // one real remainder channel + the constant-one channel: entry pair kx = (pixel[x + kx - PAD][8 G8], inside ? 1 : 0);
        // both packed planes of this row in one pass (plane G8: kx 0..3, plane G8 + 1: kx 4)
        if (v != P.G8) continue;
        uint4* dst1 = reinterpret_cast<uint4*>(planes + (size_t)(v + 1) * P.plane_bytes + (size_t)lr * pitch * 16);
        const unsigned short* rp = rowp + 8 * P.G8;
        for (int lc = lane; lc < pitch; lc += 32) {
          uint32_t w[5];
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const int xin = lc + kx - PAD;
            w[kx] = (xin >= 0 && xin < W && lc < Wp) ? ((uint32_t)rp[xin * C] | (ONE << 16)) : 0u;
          }
          dst[lc] = make_uint4(w[0], w[1], w[2], w[3]);
          dst1[lc] = make_uint4(w[4], 0u, 0u, 0u);
        }
      } else {
        const int j = v - P.G8;                                            // packed: entry e <-> (kx, channel) from the table
        int kxs[8], chs[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { kxs[e] = s_pk[j][e] & 0xff; chs[e] = s_pk[j][e] >> 8; }
        for (int lc = lane; lc < pitch; lc += 32) {
          uint32_t h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int xin = lc + kxs[e] - PAD;
            const bool ok = kxs[e] < KS && xin >= 0 && xin < W && lc < Wp;
            h[e] = !ok ? 0u : (chs[e] < C ? (uint32_t)rowp[xin * C + chs[e]] : ONE);
          }
          dst[lc] = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
        }
      }
    }
  };

  const int band0 = (int)((long long)P.total_bands * blockIdx.x / gridDim.x);
  const int band1 = (int)((long long)P.total_bands * (blockIdx.x + 1) / gridDim.x);
  __syncthreads();                                                         // s_scale, zeroed dY buffers
  if (band0 < band1) {
    cur_b = band0 / P.bands_per_image; cur_bi = band0 - cur_b * P.bands_per_image;
    issue_loads(false, 0);
    stage_dy(false, 0);
    __syncthreads();
    if (P.dup != 2) stage_planes(false, 0);
  }
  __syncthreads();
  int since_flush = 0;
  for (int band = band0; band < band1; ++band) {
    const int buf = (band - band0) & 1;
    const bool has_next = band + 1 < band1;
    if (has_next) issue_loads(true, buf ^ 1);
    // ---- MMAs: K runs over the band's output pixels, 16 per step
    {
      const int y0 = cur_bi * BR;
      const int ly_end = min(BR, H - y0);
      const uint32_t boff = (uint32_t)buf * L.buf_stride;
      for (int ly = 0; ly < ly_end; ++ly)
        for (int x0 = 0; x0 < Wp; x0 += 16) {
          uint32_t bf[(NT + 1) / 2][4];
          const uint32_t b_addr = b_lane + boff + (uint32_t)((ly * Wp + x0) * NTp * 16);
#pragma unroll
          for (int j = 0; j < (NT + 1) / 2; ++j) ldsm_x4_t(b_addr + (uint32_t)(j * 32), bf[j]);
          const uint32_t a_off = boff + (uint32_t)((ly * pitch + x0) * 16);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t af[4];
            ldsm_x4_t(a_base[mt] + a_off, af);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_16816(acc[mt][nt], af, bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
          }
        }
    }
    if (has_next) {
      stage_dy(true, buf ^ 1);                                         // buffer buf^1 was last read one iteration ago
      __syncthreads();                                                     // raw rows of band+1 are complete in shared memory
      if (P.dup != 2) stage_planes(true, buf ^ 1);
    }
    __syncthreads();                                                       // band+1 staged; every warp is done reading `buf` and raw
    if (++since_flush >= P.flush_every) { flush(); since_flush = 0; }
    if (++cur_bi == P.bands_per_image) { cur_bi = 0; ++cur_b; }             // advance to the next band
  }
  if (since_flush > 0 || first_flush) flush();
}

// ------------------------------------------------------------------------------------------ reduce + finalize
// gsum[i] = sum over CTAs of partials[cta][i] in fixed order; block (32, 8)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ Plan P, int nparts) {
  __shared__ float sh[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x, y = threadIdx.y;
  const int per = (nparts + 7) / 8, k0 = y * per, k1 = min(nparts, k0 + per);
  float s = 0.f;
  if (i < P.part_floats) for (int k = k0; k < k1; ++k) s += P.partials[(size_t)k * P.part_floats + i];
  sh[y][threadIdx.x] = s;
  __syncthreads();
  if (y == 0 && i < P.part_floats) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sh[r][threadIdx.x];
    P.gsum[i] = t;
  }
}

__device__ __forceinline__ int row_of(const Plan& P, int ky, int kx, int ch) {       // -> slab * 8 + m
  if (ch < 8 * P.G8) return (((ch >> 3) * P.KS + ky) * P.KS + kx) * 8 + (ch & 7);
  const int E = kx * P.R + (ch - 8 * P.G8);
  return (P.KS * P.KS * P.G8 + (E >> 3) * P.KS + ky) * 8 + (E & 7);
}
__device__ __forceinline__ float g_at(const Plan& P, int row, int n) {
  const int s = row >> 3, m = row & 7, T = s >> 1, half = s & 1;
  const int nt = n >> 3, cn = n & 7;
  const int idx = ((T * P.NT + nt) * 4 + half * 2 + (cn & 1)) * 32 + m * 4 + (cn >> 1);     // T = warp * MT + mt
  return P.gsum[idx];
}
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const __grid_constant__ Plan P) {
  const int KS = P.KS, Cr = P.dup == 2 ? CO : (P.dup ? P.C / 2 : P.C);
  const int one_ch = P.dup == 2 ? tc::kC24One : P.C;
  const int nw = KS * KS * Cr * CO;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nets * (nw + CO)) return;
  const int net = i / (nw + CO), j = i - net * (nw + CO);
  const float inv_scale = 1.f / scale_for(P.gmax[net][0]);
  if (j >= nw) {                                                          // bias gradient: constant-one channel, centre tap
    const int o = j - nw, row = row_of(P, P.PAD, P.PAD, one_ch);
    const float s = g_at(P, row, net * 2 * CO + o) + g_at(P, row, (net * 2 + 1) * CO + o);
    P.db[net][o] = s * inv_scale;
    return;
  }
  const int o = j % CO, c = (j / CO) % Cr, kx = (j / (CO * Cr)) % KS, ky = j / (CO * Cr * KS);
  const int n0 = net * 2 * CO + o, n1 = n0 + CO;
  const int c_hi = P.dup == 2 ? tc::c24_hi(c) : c, c_lo = P.dup == 2 ? tc::c24_lo(c) : c + Cr;
  float gsum = g_at(P, row_of(P, ky, kx, c_hi), n0) + g_at(P, row_of(P, ky, kx, c_hi), n1);
  if (P.dup) gsum += g_at(P, row_of(P, ky, kx, c_lo), n0) + g_at(P, row_of(P, ky, kx, c_lo), n1);
  float v = gsum * inv_scale;
  if (P.mean_inv) {
    const int rs = row_of(P, ky, kx, one_ch);
    const float ssum = (g_at(P, rs, n0) + g_at(P, rs, n1)) * inv_scale;
    v = P.mean_inv[P.C + c] * (v - P.mean_inv[c] * ssum);
  }
  P.dw[net][j] = v;
}

// ------------------------------------------------------------------------------------------ host
static int build_plan(int nets, int B, int H, int W, int C, int KS, Plan* P, int dup = 0) {
  CPP_REQUIRE(KS == 5 || KS == 3, "wgrad_mma: kernel size %d", KS);
  CPP_REQUIRE(nets >= 1 && nets <= kMaxNets, "wgrad_mma: %d sibling networks", nets);
  CPP_REQUIRE(H >= 2 && W >= 2 && C >= 1, "wgrad_mma: input %dx%dx%d", H, W, C);
  P->B = B; P->H = H; P->W = W; P->C = C; P->KS = KS; P->PAD = KS / 2; P->PH = H / 2; P->PW = W / 2; P->nets = nets;
  P->dup = dup;
  CPP_REQUIRE(dup != 2 || C == tc::kC24, "wgrad_mma: the aligned piece layout has %d channels", tc::kC24);
  P->CE = dup == 2 ? C : C + 1;               // the 24-channel layout already carries its constant-one channel
  P->G8 = P->CE / 8; P->R = P->CE % 8;
  P->nR = (KS * P->R + 7) / 8;
  if (P->R > 0 && P->nR >= KS) { P->G8 += 1; P->R = 0; P->nR = 0; }          // packing along kx would not save slabs
  P->nvec = P->G8 + P->nR;
  P->Wp = (int)round_up(W, 16);
  P->pitch = P->Wp + 2 * P->PAD;
  P->band_rows = kBandRows;
  P->rows_in = P->band_rows + 2 * P->PAD;
  P->plane_bytes = P->rows_in * P->pitch * 16;
  int ns = 0;
  for (int g = 0; g < P->G8; ++g)
    for (int ky = 0; ky < KS; ++ky)
      for (int kx = 0; kx < KS; ++kx) {
        CPP_REQUIRE(ns < kMaxSlabs, "wgrad_mma: too many channel groups (C=%d)", C);
        P->slab[ns] = Slab{0, (int8_t)g, (int8_t)ky, (int8_t)kx};
        P->slab_off[ns] = g * P->plane_bytes + (ky * P->pitch + kx) * 16;
        ++ns;
      }
  for (int j = 0; j < P->nR; ++j)
    for (int ky = 0; ky < KS; ++ky) {
      CPP_REQUIRE(ns < kMaxSlabs, "wgrad_mma: too many channel groups (C=%d)", C);
      P->slab[ns] = Slab{1, (int8_t)j, (int8_t)ky, 0};
      P->slab_off[ns] = (P->G8 + j) * P->plane_bytes + (ky * P->pitch) * 16;
      ++ns;
    }
  P->n_slabs = ns;
  P->m_tiles = (ns + 1) / 2;
  P->MT = (int)ceil_div(P->m_tiles, kMaxWarps);
  CPP_REQUIRE(P->MT <= 5, "wgrad_mma: %d gradient rows do not fit the register tile", ns * 8);
  P->NW = (int)ceil_div(P->m_tiles, P->MT);
  const int N = nets * 2 * CO;
  P->NT = (N + 7) / 8;
  if (P->NT == 4) P->NT = 5;                                                  // instantiated widths: 3, 5, 8
  if (P->NT == 6 || P->NT == 7) P->NT = 8;
  P->NTp = P->NT | 1;
  P->raw_bytes = dup == 2 ? 0 : (int)round_up((int64_t)P->rows_in * W * C * 2, 16);
  P->smem_bytes = (int)smem_layout(*P).total;
  if (P->smem_bytes > 220 * 1024 || (P->band_rows / 2) * (P->Wp / 2) * nets > kDyItems * 32 * P->NW) {        // wide images: two-row bands keep both staging buffers inside one SM's shared memory
    P->band_rows = 2;
    P->rows_in = P->band_rows + 2 * P->PAD;
    P->plane_bytes = P->rows_in * P->pitch * 16;
    for (int i = 0; i < ns; ++i) {
      const Slab& sl = P->slab[i];
      P->slab_off[i] = sl.kind == 0 ? sl.set * P->plane_bytes + (sl.ky * P->pitch + sl.kx) * 16
                                    : (P->G8 + sl.set) * P->plane_bytes + (sl.ky * P->pitch) * 16;
    }
    P->raw_bytes = dup == 2 ? 0 : (int)round_up((int64_t)P->rows_in * W * C * 2, 16);
    P->smem_bytes = (int)smem_layout(*P).total;
  }
  CPP_REQUIRE(P->smem_bytes <= 220 * 1024, "wgrad_mma: %dx%dx%d does not fit shared memory", H, W, C);
  P->bands_per_image = (int)ceil_div(H, P->band_rows);
  P->total_bands = B * P->bands_per_image;
  CPP_REQUIRE((P->band_rows / 2) * (P->Wp / 2) * nets <= kDyItems * 32 * P->NW, "wgrad_mma: image too wide for the dY staging (W=%d)", W);
  CPP_REQUIRE(P->nR <= 8, "wgrad_mma: too many packed planes");

  const int occ = (P->MT * P->NT <= 16 && P->smem_bytes <= 110 * 1024) ? 2 : 1;
  P->grid = std::max(1, std::min(P->total_bands, sm_budget() * occ));
  // bound the tensor-core accumulation chains to ~128 MMA steps between fp32 flushes
  const int steps_per_band = P->band_rows * (P->Wp / 16);
  P->flush_every = std::max(1, g_wgrad_flush_steps / steps_per_band);
  P->part_floats = P->NW * P->MT * P->NT * 128;
  return CPP_OK;
}

static inline size_t al256(size_t b) { return (size_t)round_up((int64_t)b, 256); }

bool conv_wgrad_mma_supported(int nets, int H, int W, int C, int KS, int dup) {
  Plan P{};
  return build_plan(nets, 1, H, W, C, KS, &P, dup) == CPP_OK;
}

int64_t conv_wgrad_mma_scratch_bytes(int nets, int H, int W, int C, int KS, int dup) {
  Plan P{};
  if (build_plan(nets, 1, H, W, C, KS, &P, dup) != CPP_OK) return -1;
  const int64_t own = (int64_t)(al256(16) + al256((size_t)2 * kNumSMs * P.part_floats * 4) + al256((size_t)P.part_floats * 4));
  const int64_t tcb = (dup == 0 || dup == 2) ? (int64_t)al256(16) + wgtc::scratch_bytes(nets, H, W, C, KS, dup == 2) : 0;     // the tcgen05 route shares the scratch
  const int64_t rowb = dup == 2 ? (int64_t)al256(16) + wgr::scratch_bytes(H, W, KS) : 0;      // ... and the row-sweep route
  return std::max(std::max(own, tcb), rowb);
}

template <int MT, int NT>
static int launch_main(const Plan& P, cudaStream_t s) {
  auto k = conv_wgrad_mma_kernel<MT, NT>;
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  k<<<P.grid, 32 * P.NW, P.smem_bytes, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

template <int MT>
static int launch_nt(const Plan& P, cudaStream_t s) {
  if (P.NT == 3) return launch_main<MT, 3>(P, s);
  if (P.NT == 5) return launch_main<MT, 5>(P, s);
  if (P.NT == 8) return launch_main<MT, 8>(P, s);
  set_error("wgrad_mma: N tile count %d not instantiated", P.NT);
  return CPP_ERR_INVALID;
}

int launch_conv_wgrad_mma(const void* x_f16, const float* mean_inv, int dup, int nets, const float* const* d_pooled,
                          const uint8_t* const* amax, int B, int H, int W, int C, int KS, float* const* dw, float* const* db,
                          void* scratch, cudaStream_t s, const float* const* gmax_pre) {
  if (B <= 0) return CPP_OK;
  Plan P{};
  CPP_TRY(build_plan(nets, B, H, W, C, KS, &P, dup));
  CPP_REQUIRE(!dup || (C % 2 == 0 && mean_inv == nullptr), "wgrad_mma: piece input needs an even channel count and no whitening");
  CPP_REQUIRE(dup != 2 || ((uintptr_t)x_f16 & 15) == 0, "wgrad_mma: unaligned piece input");
  CPP_REQUIRE(((uintptr_t)scratch & 255) == 0, "wgrad_mma: unaligned scratch");
  P.x = reinterpret_cast<const __half*>(x_f16); P.mean_inv = mean_inv; P.dup = dup;
  for (int n = 0; n < nets; ++n) {
    CPP_REQUIRE(d_pooled[n] && amax[n] && dw[n] && db[n], "wgrad_mma: null pointer for network %d", n);
    P.g[n] = d_pooled[n]; P.amax[n] = amax[n]; P.dw[n] = dw[n]; P.db[n] = db[n];
  }
  char* sc = reinterpret_cast<char*>(scratch);
  P.gmax_own = reinterpret_cast<float*>(sc); sc += al256(16);
  for (int n = 0; n < nets; ++n) P.gmax[n] = gmax_pre ? gmax_pre[n] : P.gmax_own + n;
  P.partials = reinterpret_cast<float*>(sc); sc += al256((size_t)2 * kNumSMs * P.part_floats * 4);
  P.gsum = reinterpret_cast<float*>(sc);
  if (gmax_pre == nullptr) {
    CPP_CHECK_CUDA(cudaMemsetAsync(P.gmax_own, 0, 16, s));
    const int64_t n = (int64_t)B * P.PH * P.PW * CO;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, ceil_div(n, 256 * 8)));
    wgrad_absmax_kernel<<<dim3(blocks, nets), 256, 0, s>>>(P);
    CPP_CHECK_LAUNCH();
  }
  // row-sweep tcgen05 route (conv_wgrad_row_tc.cu, wgrad_tc bit 2): conv2 / conv3 on the 24-channel pieces, one network
  if (dup == 2 && nets == 1 && (g_wgrad_tc & 4) && wgr::shape_ok(H, W, KS))
    return wgr::launch(x_f16, d_pooled[0], amax[0], P.gmax[0], B, H, W, KS, dw[0], db[0], reinterpret_cast<char*>(scratch) + al256(16), s);
  // tcgen05 route (conv_wgrad_tc.cu): conv1 of c3-class inputs (wgrad_tc bit 0), conv2 / conv3 on the 24-channel pieces (bit 1)
  if ((dup == 0 && (g_wgrad_tc & 1) && wgtc::supported(nets, H, W, C, KS, 0)) || (dup == 2 && (g_wgrad_tc & 2) && wgtc::supported(nets, H, W, C, KS, 1)))
    return wgtc::launch(x_f16, mean_inv, nets, d_pooled, amax, B, H, W, C, KS, dw, db, P.gmax, reinterpret_cast<char*>(scratch) + al256(16), s,
                        dup == 2);
  int st;
  switch (P.MT) {
    case 1: st = launch_nt<1>(P, s); break;
    case 2: st = launch_nt<2>(P, s); break;
    case 3: st = launch_nt<3>(P, s); break;
    case 4: st = launch_nt<4>(P, s); break;
    default: st = launch_nt<5>(P, s); break;
  }
  CPP_TRY(st);
  wgrad_reduce_kernel<<<(unsigned)ceil_div(P.part_floats, 32), dim3(32, 8), 0, s>>>(P, P.grid);
  CPP_CHECK_LAUNCH();
  const int Cr = dup == 2 ? CO : (dup ? C / 2 : C);
  const int total = nets * (KS * KS * Cr * CO + CO);
  wgrad_finalize_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(P);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace wg
}  // namespace cpp
