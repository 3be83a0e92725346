// FP32 direct convolution kernels for the reference conv trunk (base_network.py:73-127):
//   conv KSxKS SAME (cross-correlation) + bias + ReLU + 2x2/2 VALID max-pool, its dgrad and wgrad.
// This is the exact-fp32 CUDA-core path (parity reference for the tcgen05 path in conv_tc.cu).
//
// Layouts: activations NHWC; weights HWIO (kh,kw,Cin,10).  The pooled output carries a side band
// `amax` (u8 per pooled element): 0..3 = position of the max inside the 2x2 window in scan order,
// 4 = the pooled value was <= 0 (ReLU closed, no gradient).
#include "common.cuh"

namespace cpp {

constexpr int CO = kConvCout;   // 10 output channels
constexpr int CP = 12;          // padded to 3 x float4 in shared memory

struct ConvParams {
  const void* x;            // IN 0: fp16 NHWC ; IN 1: fp32 NHWC ; IN 2: pooled gradient f32 (B,PH,PW,10) ; IN 3: dense gradient (B,H,W,10)
  const uint8_t* gamax;     // IN 2
  const float* mean_inv;    // IN 0: [mean(Cin) | inv(Cin)]
  const float* w;           // HWIO
  const float* bias;
  float* out;               // OUT 0: pooled (B,PH,PW,10) ; OUT 1: dense (B,H,W,10)
  uint8_t* amax;            // OUT 0
  int B, H, W, Cin, PH, PW;
  int CC;                   // input channels staged per pass
  int chpitch;              // shared-memory floats per staged channel plane
};

// One thread: 2 rows x 4 cols of conv outputs x 10 channels (80 fp32 accumulators).
// Block (TX,TY,TZ): tile of 2*TY rows x 4*TX cols for TZ images.  Input planes are staged
// channel-planar in shared memory so each thread reads its 8-wide row segment with two LDS.128
// and re-uses it for every kx (sliding window); weights are read as warp-wide broadcasts.
template <int KS, int IN_MODE, int OUT_MODE>
__global__ void __launch_bounds__(128, 4) conv_kernel(ConvParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PAD = KS / 2;
  const int TX = blockDim.x, TY = blockDim.y, TZ = blockDim.z;
  const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
  const int tid = (tz * TY + ty) * TX + tx, nthr = TX * TY * TZ;
  const int rows = 2 * TY + KS - 1, TWP = 4 * TX + 4, tile_w = 4 * TX + KS - 1;
  float* in_s = smem;
  float* w_s = smem + TZ * p.CC * p.chpitch;
  const int x0 = blockIdx.x * 4 * TX, y0 = blockIdx.y * 2 * TY, b0 = blockIdx.z * TZ;

  float acc[2][4][CO];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 0; o < CO; ++o) acc[r][j][o] = 0.f;

  for (int c0 = 0; c0 < p.Cin; c0 += p.CC) {
    __syncthreads();
    for (int i = tid; i < KS * KS * p.CC * CO; i += nthr) {
      const int o = i % CO, c = (i / CO) % p.CC, tap = i / (CO * p.CC);
      float v = 0.f;
      if (c0 + c < p.Cin) {
        if (IN_MODE >= 2) v = p.w[((KS * KS - 1 - tap) * CO + o) * CO + (c0 + c)];   // flipped taps, in/out swapped
        else v = p.w[(tap * p.Cin + c0 + c) * CO + o];
      }
      w_s[(tap * p.CC + c) * CP + o] = v;
    }
    const int per_img = rows * tile_w * p.CC;
    for (int e = tid; e < TZ * per_img; e += nthr) {
      const int z = e / per_img, r2 = e - z * per_img;
      const int c = r2 % p.CC, ix = (r2 / p.CC) % tile_w, iy = r2 / (p.CC * tile_w);
      const int gy = y0 + iy - PAD, gx = x0 + ix - PAD, b = b0 + z, ch = c0 + c;
      float v = 0.f;
      if (b < p.B && ch < p.Cin && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
        if (IN_MODE == 0) {
          const __half* xh = reinterpret_cast<const __half*>(p.x);
          const float xv = __half2float(xh[(((size_t)b * p.H + gy) * p.W + gx) * p.Cin + ch]);
          const float inv = p.mean_inv[p.Cin + ch];
          v = __fsub_rn(__fmul_rn(xv, inv), __fmul_rn(p.mean_inv[ch], inv));   // x*inv - mean*inv, base_network.py:97
        } else if (IN_MODE == 1) {
          v = reinterpret_cast<const float*>(p.x)[(((size_t)b * p.H + gy) * p.W + gx) * p.Cin + ch];
          if (p.mean_inv) {   // fp32 state fed straight from the env (action_given): same whitening
            const float inv = p.mean_inv[p.Cin + ch];
            v = __fsub_rn(__fmul_rn(v, inv), __fmul_rn(p.mean_inv[ch], inv));
          }
        } else if (IN_MODE == 2) {
          const int py = gy >> 1, px = gx >> 1;
          if (py < p.PH && px < p.PW) {
            const size_t idx = (((size_t)b * p.PH + py) * p.PW + px) * CO + ch;
            if (p.gamax[idx] == (((gy & 1) << 1) | (gx & 1))) v = reinterpret_cast<const float*>(p.x)[idx];
          }
        } else {                                                  // IN 3: dense gradient wrt the conv output (batch-norm route)
          v = reinterpret_cast<const float*>(p.x)[(((size_t)b * p.H + gy) * p.W + gx) * CO + ch];
        }
      }
      in_s[(z * p.CC + c) * p.chpitch + iy * TWP + ix] = v;
    }
    __syncthreads();

    const float* my_in = in_s + tz * p.CC * p.chpitch + (2 * ty) * TWP + 4 * tx;
    for (int c = 0; c < p.CC; ++c) {
      const float* inc = my_in + c * p.chpitch;
      const float* wc = w_s + c * CP;
#pragma unroll
      for (int iy = 0; iy < KS + 1; ++iy) {
        const float4 va = *reinterpret_cast<const float4*>(inc + iy * TWP);
        const float4 vb = *reinterpret_cast<const float4*>(inc + iy * TWP + 4);
        const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int ky = iy - r;
          if (ky < 0 || ky >= KS) continue;
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float* wp = wc + (ky * KS + kx) * p.CC * CP;
            const float4 w0 = *reinterpret_cast<const float4*>(wp);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
            const float2 w2 = *reinterpret_cast<const float2*>(wp + 8);
            const float w[CO] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int o = 0; o < CO; ++o) acc[r][j][o] = fmaf(v[j + kx], w[o], acc[r][j][o]);
          }
        }
      }
    }
  }

  const int b = b0 + tz;
  if (b >= p.B) return;
  if (OUT_MODE == 0) {
    const int py = (y0 >> 1) + ty;
    if (py >= p.PH) return;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int px = (x0 >> 1) + 2 * tx + jj;
      if (px >= p.PW) continue;
      const size_t base = (((size_t)b * p.PH + py) * p.PW + px) * CO;
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        const float bo = p.bias[o];
        float m = acc[0][2 * jj][o] + bo;
        int a = 0;
        const float v01 = acc[0][2 * jj + 1][o] + bo, v10 = acc[1][2 * jj][o] + bo, v11 = acc[1][2 * jj + 1][o] + bo;
        if (v01 > m) { m = v01; a = 1; }
        if (v10 > m) { m = v10; a = 2; }
        if (v11 > m) { m = v11; a = 3; }
        p.out[base + o] = fmaxf(m, 0.f);
        p.amax[base + o] = (m > 0.f) ? (uint8_t)a : (uint8_t)4;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int y = y0 + 2 * ty + r;
      if (y >= p.H) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + 4 * tx + j;
        if (x >= p.W) continue;
        float* o_ = p.out + (((size_t)b * p.H + y) * p.W + x) * CO;
#pragma unroll
        for (int o = 0; o < CO; ++o) o_[o] = acc[r][j][o];
      }
    }
  }
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct ConvTile { dim3 grid, block; int CC, chpitch; size_t smem; };

static ConvTile pick_tile(int H, int W, int Cin, int KS, int B) {
  ConvTile t;
  const int TX = (W <= 16) ? 4 : (W <= 32) ? 8 : 16;
  int TY = 128 / TX;
  const int hp = pow2ceil((H + 1) / 2);
  if (hp < TY) TY = hp;
  int TZ = 128 / (TX * TY);
  const int rows = 2 * TY + KS - 1, TWP = 4 * TX + 4;
  int chp = rows * TWP;
  while (chp % 32 != 4) chp += 4;
  int nch = (int)ceil_div(Cin, 12);
  int CC = (int)ceil_div(Cin, nch);
  auto smem_of = [&](int cc, int tz) { return (size_t)(tz * cc * chp + KS * KS * cc * CP) * sizeof(float); };
  while (smem_of(CC, TZ) > 100 * 1024 && TZ > 1) TZ /= 2;
  while (smem_of(CC, TZ) > 100 * 1024 && CC > 1) CC = (CC + 1) / 2;
  t.block = dim3(TX, TY, TZ);
  t.grid = dim3((unsigned)ceil_div(W, 4 * TX), (unsigned)ceil_div(H, 2 * TY), (unsigned)ceil_div(B, TZ));
  t.CC = CC; t.chpitch = chp; t.smem = smem_of(CC, TZ);
  return t;
}

template <int KS, int IN_MODE, int OUT_MODE>
static int launch_conv_t(const ConvParams& p, const ConvTile& t, cudaStream_t s) {
  auto k = conv_kernel<KS, IN_MODE, OUT_MODE>;
  static size_t configured = 0;   // per instantiation
  if (t.smem > configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)));
    configured = 100 * 1024;
  }
  k<<<t.grid, t.block, t.smem, s>>>(p);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

int launch_conv_fwd(const ConvLayer& L, const void* x, int x_is_f16, const float* mean_inv,
                    const float* w, const float* b, int B, float* pooled, uint8_t* amax, cudaStream_t s) {
  CPP_REQUIRE(L.KS == 5 || L.KS == 3, "conv: kernel size %d unsupported", L.KS);
  CPP_REQUIRE(!x_is_f16 || mean_inv != nullptr, "conv: fp16 input needs whitening stats");
  CPP_REQUIRE(L.PH() >= 1 && L.PW() >= 1, "conv: input %dx%d too small to pool", L.H, L.W);
  if (B <= 0) return CPP_OK;
  ConvParams p{};
  p.x = x; p.mean_inv = mean_inv; p.w = w; p.bias = b; p.out = pooled; p.amax = amax;
  p.B = B; p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.PH = L.PH(); p.PW = L.PW();
  ConvTile t = pick_tile(L.H, L.W, L.Cin, L.KS, B);
  p.CC = t.CC; p.chpitch = t.chpitch;
  if (x_is_f16) return L.KS == 5 ? launch_conv_t<5, 0, 0>(p, t, s) : launch_conv_t<3, 0, 0>(p, t, s);
  return L.KS == 5 ? launch_conv_t<5, 1, 0>(p, t, s) : launch_conv_t<3, 1, 0>(p, t, s);
}

int launch_conv_dgrad(const ConvLayer& L, const float* d_pooled, const uint8_t* amax, const float* w,
                      int B, float* dx, cudaStream_t s) {
  CPP_REQUIRE(L.Cin == CO, "conv dgrad is only needed for conv2/conv3 (Cin == 10), got %d", L.Cin);
  if (B <= 0) return CPP_OK;
  ConvParams p{};
  p.x = d_pooled; p.gamax = amax; p.w = w; p.out = dx;
  p.B = B; p.H = L.H; p.W = L.W; p.Cin = CO; p.PH = L.PH(); p.PW = L.PW();
  ConvTile t = pick_tile(L.H, L.W, CO, L.KS, B);
  p.CC = t.CC; p.chpitch = t.chpitch;
  return L.KS == 5 ? launch_conv_t<5, 2, 1>(p, t, s) : launch_conv_t<3, 2, 1>(p, t, s);
}

// --use-batch-norm route (base_network.py:74-79): slim.conv2d with a normalizer_fn has no bias and hands its raw output to
// slim.batch_norm, so the layer is split: raw SAME conv here, statistics / normalise / ReLU / pool in bn.cu
int launch_conv_raw(const ConvLayer& L, const void* x, int x_is_f16, const float* mean_inv, const float* w, int B, float* raw,
                    cudaStream_t s) {
  CPP_REQUIRE(L.KS == 5 || L.KS == 3, "conv: kernel size %d unsupported", L.KS);
  CPP_REQUIRE(!x_is_f16 || mean_inv != nullptr, "conv: fp16 input needs whitening stats");
  if (B <= 0) return CPP_OK;
  ConvParams p{};
  p.x = x; p.mean_inv = mean_inv; p.w = w; p.out = raw;
  p.B = B; p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.PH = L.PH(); p.PW = L.PW();
  ConvTile t = pick_tile(L.H, L.W, L.Cin, L.KS, B);
  p.CC = t.CC; p.chpitch = t.chpitch;
  if (x_is_f16) return L.KS == 5 ? launch_conv_t<5, 0, 1>(p, t, s) : launch_conv_t<3, 0, 1>(p, t, s);
  return L.KS == 5 ? launch_conv_t<5, 1, 1>(p, t, s) : launch_conv_t<3, 1, 1>(p, t, s);
}

// d(x) from the DENSE gradient wrt the conv output (B,H,W,10)
int launch_conv_dgrad_dense(const ConvLayer& L, const float* d_conv, const float* w, int B, float* dx, cudaStream_t s) {
  CPP_REQUIRE(L.Cin == CO, "conv dgrad is only needed for conv2/conv3 (Cin == 10), got %d", L.Cin);
  if (B <= 0) return CPP_OK;
  ConvParams p{};
  p.x = d_conv; p.w = w; p.out = dx;
  p.B = B; p.H = L.H; p.W = L.W; p.Cin = CO; p.PH = L.PH(); p.PW = L.PW();
  ConvTile t = pick_tile(L.H, L.W, CO, L.KS, B);
  p.CC = t.CC; p.chpitch = t.chpitch;
  return L.KS == 5 ? launch_conv_t<5, 3, 1>(p, t, s) : launch_conv_t<3, 3, 1>(p, t, s);
}

// ------------------------------------------------------------------------------------------ wgrad

struct WgradParams {
  const void* x; const float* mean_inv;
  const float* gp; const uint8_t* gamax;
  float* partials;           // [gridDim.x][nout]
  int B, H, W, Cin, PH, PW;
  int G, TWt, chpitch, tiles_x, tiles_y, total_tiles, nout;
};

// dW[ky][kx][c][o] = sum_{b,y,x} in[b][y+ky-P][x+kx-P][c] * g[b][y][x][o] ;  db[o] = sum g.
// Thread (group, ky, c) owns the KS x 10 accumulators of its (ky, c); a group covers 2 output rows of
// the tile; the input row segment slides along x in registers.  Each CTA walks a strided list of tiles
// and writes ONE partial, reduced afterwards in fixed order (deterministic, needed for DP replicas).
template <int KS, int IN_MODE>
__global__ void __launch_bounds__(256, 2) conv_wgrad_kernel(WgradParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PAD = KS / 2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int per_group = KS * p.Cin;
  const int grp = tid / per_group, l = tid - grp * per_group;
  const bool active = grp < p.G;
  const int ky = l / p.Cin, c = l - ky * p.Cin;
  const int TH = 2 * p.G, TWt = p.TWt, TWP = TWt + 4, rows = TH + KS - 1, tile_w = TWt + KS - 1;
  float* in_s = smem;                                 // [Cin][chpitch]
  float* g_s = smem + p.Cin * p.chpitch;              // [TH][TWt][CP]

  float acc[KS][CO];
#pragma unroll
  for (int k = 0; k < KS; ++k)
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[k][o] = 0.f;
  float bacc = 0.f;

  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int b = tile / (p.tiles_x * p.tiles_y), t2 = tile - b * (p.tiles_x * p.tiles_y);
    const int y0 = (t2 / p.tiles_x) * TH, x0 = (t2 % p.tiles_x) * TWt;
    __syncthreads();
    for (int e = tid; e < rows * tile_w * p.Cin; e += nthr) {
      const int ch = e % p.Cin, ix = (e / p.Cin) % tile_w, iy = e / (p.Cin * tile_w);
      const int gy = y0 + iy - PAD, gx = x0 + ix - PAD;
      float v = 0.f;
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
        if (IN_MODE == 0) {
          const float xv = __half2float(reinterpret_cast<const __half*>(p.x)[(((size_t)b * p.H + gy) * p.W + gx) * p.Cin + ch]);
          const float inv = p.mean_inv[p.Cin + ch];
          v = __fsub_rn(__fmul_rn(xv, inv), __fmul_rn(p.mean_inv[ch], inv));
        } else {
          v = reinterpret_cast<const float*>(p.x)[(((size_t)b * p.H + gy) * p.W + gx) * p.Cin + ch];
          if (p.mean_inv) {
            const float inv = p.mean_inv[p.Cin + ch];
            v = __fsub_rn(__fmul_rn(v, inv), __fmul_rn(p.mean_inv[ch], inv));
          }
        }
      }
      in_s[ch * p.chpitch + iy * TWP + ix] = v;
    }
    // the 4 floats after tile_w in each row are read (never used) by the sliding float4 prefetch
    for (int e = tid; e < TH * TWt * CO; e += nthr) {
      const int o = e % CO, x = (e / CO) % TWt, y = e / (CO * TWt);
      const int gy = y0 + y, gx = x0 + x, py = gy >> 1, px = gx >> 1;
      float v = 0.f;
      if (p.gamax == nullptr) {                                    // dense gradient wrt the conv output (batch-norm route)
        if (gy < p.H && gx < p.W) v = p.gp[(((size_t)b * p.H + gy) * p.W + gx) * CO + o];
      } else if (gy < p.H && gx < p.W && py < p.PH && px < p.PW) {
        const size_t idx = (((size_t)b * p.PH + py) * p.PW + px) * CO + o;
        if (p.gamax[idx] == (((gy & 1) << 1) | (gx & 1))) v = p.gp[idx];
      }
      g_s[(y * TWt + x) * CP + o] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int y = grp * 2 + rr;
        const float* inrow = in_s + c * p.chpitch + (y + ky) * TWP;
        const float* grow = g_s + (size_t)y * TWt * CP;
        float4 cur = *reinterpret_cast<const float4*>(inrow);
        for (int x4 = 0; x4 < TWt; x4 += 4) {
          const float4 nxt = *reinterpret_cast<const float4*>(inrow + x4 + 4);
          const float v[8] = {cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* gpx = grow + (x4 + j) * CP;
            const float4 g0 = *reinterpret_cast<const float4*>(gpx);
            const float4 g1 = *reinterpret_cast<const float4*>(gpx + 4);
            const float2 g2 = *reinterpret_cast<const float2*>(gpx + 8);
            const float g[CO] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y};
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
#pragma unroll
              for (int o = 0; o < CO; ++o) acc[kx][o] = fmaf(v[j + kx], g[o], acc[kx][o]);
          }
          cur = nxt;
        }
        if (l < CO) {   // bias gradient: thread l of each group sums channel l of its rows
          for (int x = 0; x < TWt; ++x) bacc += grow[x * CP + l];
        }
      }
    }
  }

  // fixed-order reduction over the groups, then one partial per CTA
  __syncthreads();
  float* red = smem;   // [nout]
  for (int gi = 0; gi < p.G; ++gi) {
    if (active && grp == gi) {
#pragma unroll
      for (int kx = 0; kx < KS; ++kx)
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const int i = ((ky * KS + kx) * p.Cin + c) * CO + o;
          red[i] = (gi == 0) ? acc[kx][o] : red[i] + acc[kx][o];
        }
      if (l < CO) red[KS * KS * p.Cin * CO + l] = (gi == 0) ? bacc : red[KS * KS * p.Cin * CO + l] + bacc;
    }
    __syncthreads();
  }
  float* out = p.partials + (size_t)blockIdx.x * p.nout;
  for (int i = tid; i < p.nout; i += nthr) out[i] = red[i];
}

// block (32, 8): column x sums the partials of output i, row y covers a contiguous range of CTAs; fixed order
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nparts, int nout, int nw,
                                                              float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sh[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x, y = threadIdx.y;
  const int per = (nparts + 7) / 8, k0 = y * per, k1 = min(nparts, k0 + per);
  float s = 0.f;
  if (i < nout) for (int k = k0; k < k1; ++k) s += partials[(size_t)k * nout + i];
  sh[y][threadIdx.x] = s;
  __syncthreads();
  if (y == 0 && i < nout) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sh[r][threadIdx.x];
    if (i < nw) dw[i] = t; else db[i - nw] = t;
  }
}

struct WgradPlan { int G, TWt, chpitch, tiles_x, tiles_y, nout, threads, grid_cap; size_t smem; };

static WgradPlan wgrad_plan(const ConvLayer& L) {
  WgradPlan w;
  const int per_group = L.KS * L.Cin;
  int gmax = 256 / per_group; if (gmax > 8) gmax = 8; if (gmax < 1) gmax = 1;
  int G = gmax;
  for (int g = gmax; g >= (gmax + 1) / 2; --g) if (L.H % (2 * g) == 0) { G = g; break; }
  w.G = G;
  w.TWt = (int)round_up(L.W < 64 ? L.W : 64, 4);
  const int rows = 2 * G + L.KS - 1, TWP = w.TWt + 4;
  int chp = rows * TWP; while (chp % 32 != 4) chp += 4;
  w.chpitch = chp;
  w.tiles_x = (int)ceil_div(L.W, w.TWt); w.tiles_y = (int)ceil_div(L.H, 2 * G);
  w.nout = L.KS * L.KS * L.Cin * CO + CO;
  w.threads = (int)round_up(G * per_group, 32);
  size_t tile_b = (size_t)(L.Cin * chp + 2 * G * w.TWt * CP) * sizeof(float);
  size_t red_b = (size_t)w.nout * sizeof(float);
  w.smem = tile_b > red_b ? tile_b : red_b;
  w.grid_cap = 2 * kNumSMs;
  return w;
}

int64_t conv_wgrad_scratch_floats(const ConvLayer& L) {
  WgradPlan w = wgrad_plan(L);
  return (int64_t)w.grid_cap * w.nout;
}

template <int KS, int IN_MODE>
static int launch_wgrad_t(const WgradParams& p, const WgradPlan& w, int grid, cudaStream_t s) {
  auto k = conv_wgrad_kernel<KS, IN_MODE>;
  static bool configured = false;
  if (!configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(110 * 1024)));
    configured = true;
  }
  k<<<grid, w.threads, w.smem, s>>>(p);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

int launch_conv_wgrad(const ConvLayer& L, const void* x, int x_is_f16, const float* mean_inv,
                      const float* d_pooled, const uint8_t* amax, int B,
                      float* dw, float* db, float* scratch, cudaStream_t s) {
  CPP_REQUIRE(L.KS == 5 || L.KS == 3, "conv: kernel size %d unsupported", L.KS);
  CPP_REQUIRE(L.KS * L.Cin <= 256, "conv wgrad: Cin=%d too large", L.Cin);
  WgradPlan w = wgrad_plan(L);
  CPP_REQUIRE(w.smem <= 110 * 1024, "conv wgrad: tile does not fit shared memory (Cin=%d)", L.Cin);
  WgradParams p{};
  p.x = x; p.mean_inv = mean_inv; p.gp = d_pooled; p.gamax = amax; p.partials = scratch;
  p.B = B; p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.PH = L.PH(); p.PW = L.PW();
  p.G = w.G; p.TWt = w.TWt; p.chpitch = w.chpitch; p.tiles_x = w.tiles_x; p.tiles_y = w.tiles_y;
  p.total_tiles = B * w.tiles_x * w.tiles_y; p.nout = w.nout;
  int grid = p.total_tiles < w.grid_cap ? p.total_tiles : w.grid_cap;
  if (grid < 1) grid = 1;
  if (x_is_f16) { CPP_TRY((L.KS == 5 ? launch_wgrad_t<5, 0>(p, w, grid, s) : launch_wgrad_t<3, 0>(p, w, grid, s))); }
  else { CPP_TRY((L.KS == 5 ? launch_wgrad_t<5, 1>(p, w, grid, s) : launch_wgrad_t<3, 1>(p, w, grid, s))); }
  reduce_partials_kernel<<<(unsigned)ceil_div(w.nout, 32), dim3(32, 8), 0, s>>>(scratch, grid, w.nout, w.nout - CO, dw, db);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
