// Fused fully connected stacks (slim.fully_connected chains, base_network.py:58-71; ddpg_cartpole.py:95-100,168-184;
// naf_cartpole.py:105-109,156-184).  The layers are tiny (K <= 2560, N <= 200, batch 256): one GEMM launch per layer is
// pure launch / dependency latency on the critical path of a step, so
//   mlp_forward_kernel   runs layers [first, end) of one network for 8 batch rows per CTA, activations in shared memory,
//   mlp_dgrad_kernel     runs the chain of input gradients dX_i = (dPre_i . W_i^T) * relu'(h_{i-1}) from the top down,
// both writing every intermediate (activations h_i, gradients dPre_i) to the workspace, because the weight gradients
// (x^T . dPre, the plain GEMM of fc.cu on a side stream) need them.  Exact fp32 FFMA with a fixed summation order.
#include "net.cuh"
#include "umma.cuh"

namespace cpp {

constexpr int kRows = 8;          // batch rows per CTA
constexpr int kMlpThreads = 256;

struct MlpLayer {
  const float* W; const float* b;   // [in][out], [out]
  float* h;                         // activation output in the workspace [B][h_ld]
  float* dX;                        // backward: gradient wrt this layer's input [B][in]
  int in, out, act, h_ld;
  int action_in;                    // 1: the last action_dim inputs of this layer are the action (tf.concat, ddpg_cartpole.py:170,175)
};

struct MlpArgs {
  MlpLayer L[CPP_MAX_FC];
  int first, end;                   // forward: layers [first, end);  backward: layers end-1 down to first
  int B, action_dim;
  const float* x; int x_ld, x_cols; // forward: input of layer `first` without the action columns
  float* x_tail;                    // forward: where the action columns of that input live in the workspace (or NULL)
  const float* action;              // forward: [B][action_dim] or NULL
  // backward
  const float* d_out;               // [B][out(end-1)] gradient wrt the post-activation output of layer end-1
  float* dTop;                      // [B][out(end-1)] gradient wrt its pre-activation (kept for the weight gradient)
  float* d_action;                  // [B][action_dim] or NULL
  int tiled;                        // forward: register-tiled inner loop (4 columns x 8 rows per thread) - CARTPOLEPP_MLP_TILED
  int rotate;                       // forward: CTA b starts its tile sequence at tile b mod ntiles (mlp_fast bit 2)
  int bulk;                         // forward: weight tiles by bulk copies (TMA engine) instead of per-thread cp.async - CARTPOLEPP_MLP_BULK
  int need_dx_first;                // 1: also produce dX of layer `first`
  float* absmax_first;              // optional: max |dX of layer `first`| is atomically max-ed into this device float (zeroed by the caller):
                                    // the power-of-two scale of conv3's input / weight gradient pieces without a pass of its own
  int maxw;
};

__device__ __forceinline__ float act_fwd(float v, int act) {
  return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? tanhf(v) : v);
}

// ------------------------------------------------------------------------------------------ weight streaming
// The weight matrix of a layer is streamed global -> shared in tiles of KT input rows x N outputs through a 4-stage
// cp.async ring, so that the L2 latency of ~hundreds of rows is paid once, not per row (one CTA alone cannot keep enough
// plain loads in flight).  W tiles are contiguous in memory (row-major [in][out]).
constexpr int kStages = 4;
constexpr int kTileFloats = 8192;          // 32 KB per stage: ~100 KB of weights in flight per CTA hide the L2 latency

__device__ __forceinline__ int tile_rows(int K, int N) {
  int kt = (kTileFloats / N) & ~3;
  if (kt < 4) kt = 4;
  return kt < K ? kt : K;
}
// issue the copy of tile t (rows [t*KT, min(K, (t+1)*KT))) into `dst`; every thread commits exactly one group
__device__ __forceinline__ void issue_tile(const float* __restrict__ W, int K, int N, int KT, int t, float* dst, int tid, bool vec4) {
  const int k0 = t * KT;
  if (k0 < K) {
    const int rows = min(KT, K - k0), n_el = rows * N;
    const float* src = W + (size_t)k0 * N;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    if (vec4) {
      for (int i = tid * 4; i < n_el; i += kMlpThreads * 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
    } else {
      for (int i = tid; i < n_el; i += kMlpThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + i * 4), "l"(src + i) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// The same tile as ONE bulk copy (TMA engine, cp.async.bulk global -> shared, mbarrier completion) issued by thread 0: a CTA's
// 256 threads x 16-byte cp.async keep only ~11 bytes / clk in flight from L2 (3 k clk per 32 KB tile whatever the FMA loop costs,
// scripts/prof_mlp.py); the bulk engine moves the tile without occupying load/store slots.  Needs 16-byte sizes and addresses.
__device__ __forceinline__ bool bulk_ok(const float* W, int K, int N) {
  return ((reinterpret_cast<uintptr_t>(W) & 15) == 0) && ((((K & 3) * N) & 3) == 0);
}
__device__ __forceinline__ void issue_tile_bulk(const float* __restrict__ W, int K, int N, int KT, int t, float* dst, uint64_t* bar, int tid) {
  const int k0 = t * KT;
  if (tid == 0 && k0 < K) {
    const uint32_t bytes = (uint32_t)(min(KT, K - k0) * N) * 4u;
    umma::fence_proxy_async();                                    // the stage was read through the generic proxy before
    umma::mbar_expect_tx(bar, bytes);
    umma::bulk_g2s(dst, W + (size_t)k0 * N, bytes, bar);
  }
}

// ------------------------------------------------------------------------------------------ forward
// shared: xa / xb [maxw][kRows] (k-major so that one k is two float4 loads for all 8 rows), red [kMlpThreads][kRows],
// wt [kStages][kTileFloats]
#ifdef MLP_PROF
__device__ long long g_mlp_prof[64][8];
extern "C" __attribute__((visibility("default"))) int cpp_debug_mlp_prof(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_mlp_prof, sizeof(long long) * 64 * 8);
}
#define MLP_STAMP(i) { if (threadIdx.x == 0 && blockIdx.x < 64) g_mlp_prof[blockIdx.x][i] = clock64(); }
#else
#define MLP_STAMP(i)
#endif
__global__ void __launch_bounds__(kMlpThreads) mlp_forward_kernel(const __grid_constant__ MlpArgs A) {
  extern __shared__ __align__(16) float sm[];
  MLP_STAMP(0)
  float* xa = sm;
  float* xb = sm + (size_t)A.maxw * kRows;
  float* red = xb + (size_t)A.maxw * kRows;
  float* wt = red + (size_t)kMlpThreads * kRows;
  const int tid = threadIdx.x, row0 = blockIdx.x * kRows;
  const int nrow = min(kRows, A.B - row0);
  __shared__ uint64_t wbar[kStages];                             // bulk route: one mbarrier per weight stage
  uint32_t wphase = 0;                                           // bit s: parity the next wait on stage s expects
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) umma::mbar_init(&wbar[i], 1);
    umma::fence_mbar_init();
  }
  __syncthreads();
  // tile t of a layer -> stage t % kStages, by bulk copy when the layer allows it, else by per-thread cp.async
  // Position t of a layer's tile sequence goes to stage t % kStages; the tile it holds is (t + rot) % ntiles with rot = this CTA's
  // index (A.rotate): the CTAs then pull DIFFERENT weight tiles at any moment instead of all hitting the same L2 lines.  The order
  // in which a row's K range is summed depends on the CTA that owns the row - a fixed function of the row, so still deterministic.
  auto tile_at = [&](int t, int ntl) { return t < ntl ? (A.rotate ? (t + (int)blockIdx.x) % ntl : t) : ntl; };
  auto issue = [&](const float* W, int K, int N, int KT, int t, bool bulk, bool v4) {
    float* dst = wt + (t % kStages) * kTileFloats;
    const int tt = tile_at(t, (K + KT - 1) / KT);                // (past the end: nothing to copy, an empty group keeps the counts)
    if (bulk) { issue_tile_bulk(W, K, N, KT, tt, dst, &wbar[t % kStages], tid); asm volatile("cp.async.commit_group;" ::: "memory"); }
    else issue_tile(W, K, N, KT, tt, dst, tid, v4);
  };
  auto wait_tile = [&](int t, bool bulk) {
    if (bulk) { const int st = t % kStages; umma::mbar_wait(&wbar[st], (wphase >> st) & 1u); wphase ^= 1u << st; }
    else asm volatile("cp.async.wait_group 2;" ::: "memory");
  };
  // start streaming the first layer's weights, then fetch its input (coalesced along k) and its action columns
  {
    const MlpLayer& L0 = A.L[A.first];
    const bool vec4 = ((reinterpret_cast<uintptr_t>(L0.W) & 15) == 0) && (L0.out % 4 == 0);
    const int KT = tile_rows(L0.in, L0.out);
    const bool bulk0 = A.bulk && bulk_ok(L0.W, L0.in, L0.out);
    for (int t = 0; t < kStages - 1; ++t) issue(L0.W, L0.in, L0.out, KT, t, bulk0, vec4);
  }
  for (int k = tid; k < A.x_cols; k += kMlpThreads) {
    float v[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) v[r] = r < nrow ? __ldg(A.x + (size_t)(row0 + r) * A.x_ld + k) : 0.f;
    *reinterpret_cast<float4*>(xa + k * kRows) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(xa + k * kRows + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (A.L[A.first].action_in) {
    for (int i = tid; i < kRows * A.action_dim; i += kMlpThreads) {
      const int r = i / A.action_dim, j = i - r * A.action_dim;
      const float v = r < nrow ? A.action[(size_t)(row0 + r) * A.action_dim + j] : 0.f;
      xa[(A.x_cols + j) * kRows + r] = v;
      if (r < nrow && A.x_tail) A.x_tail[(size_t)(row0 + r) * A.x_ld + j] = v;
    }
  }
  MLP_STAMP(1)
  for (int l = A.first; l < A.end; ++l) {
    MLP_STAMP(2 + (l - A.first))
    const MlpLayer& Ly = A.L[l];
    const int N = Ly.out, K = Ly.in;
    const bool vec4 = ((reinterpret_cast<uintptr_t>(Ly.W) & 15) == 0) && (N % 4 == 0);
    const int KT = tile_rows(K, N), ntiles = (K + KT - 1) / KT;
    const bool bulk = A.bulk && bulk_ok(Ly.W, K, N);
    const bool more = l + 1 < A.end;
    const int ngrp = (N + 3) >> 2;
    if (ngrp <= 64 && A.tiled && K >= 256) {
      // Register-tiled: a thread owns 4 output columns (and 4 more, 128 columns further, when N > 128) x the 8 rows; the 8 K
      // slices of a column group sit in ONE warp (lane = slice * 4 + group), so the reduction over them is three shuffles.  Per k:
      // two 16-byte x loads + one 16-byte w load for 32 FMAs (the one-column mapping below needs 3 loads for 8).
      const int warp = tid >> 5, lane = tid & 31, slice = lane >> 2;
      const int c0 = 4 * (warp * 4 + (lane & 3)), c1 = c0 + 128;
      const bool a0 = c0 < N, a1 = c1 < N;
      float acc0[4][kRows], acc1[4][kRows];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < kRows; ++r) { acc0[c][r] = 0.f; acc1[c][r] = 0.f; }
      auto load4 = [&](const float* p, int avail) -> float4 {     // 4 weights of one k; columns past N read as 0
        if (vec4) return *reinterpret_cast<const float4*>(p);
        return make_float4(p[0], avail > 1 ? p[1] : 0.f, avail > 2 ? p[2] : 0.f, avail > 3 ? p[3] : 0.f);
      };
      for (int t = 0; t < ntiles; ++t) {
        wait_tile(t, bulk);                                       // tile t has landed (kStages - 2 younger tiles may be in flight)
        __syncthreads();                                         // ... for every thread; the stage of tile t-1 is free again; xa is complete
        issue(Ly.W, K, N, KT, t + kStages - 1, bulk, vec4);
        const float* wtile = wt + (t % kStages) * kTileFloats;
        const int k0 = tile_at(t, ntiles) * KT, rows = min(KT, K - k0);
        if (a0) {
#pragma unroll 2
          for (int kk = slice; kk < rows; kk += 8) {
            const float* xp = xa + (k0 + kk) * kRows;
            const float4 x0 = *reinterpret_cast<const float4*>(xp), x1 = *reinterpret_cast<const float4*>(xp + 4);
            const float xv[kRows] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const float4 w0 = load4(wtile + kk * N + c0, N - c0);
            const float wv[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int r = 0; r < kRows; ++r) acc0[c][r] = fmaf(xv[r], wv[c], acc0[c][r]);
            if (a1) {
              const float4 w1 = load4(wtile + kk * N + c1, N - c1);
              const float wu[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
              for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < kRows; ++r) acc1[c][r] = fmaf(xv[r], wu[c], acc1[c][r]);
            }
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                           // all weight stages idle: the next layer may start streaming
      if (more) {
        const MlpLayer& Ln = A.L[l + 1];
        const bool v4 = ((reinterpret_cast<uintptr_t>(Ln.W) & 15) == 0) && (Ln.out % 4 == 0);
        const int KTn = tile_rows(Ln.in, Ln.out);
        { const bool bn = A.bulk && bulk_ok(Ln.W, Ln.in, Ln.out); for (int t = 0; t < kStages - 1; ++t) issue(Ln.W, Ln.in, Ln.out, KTn, t, bn, v4); }
      }
      // fixed-order butterfly over the 8 K slices: afterwards every lane holds the sums; lane `slice` finishes row r = slice
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          float v = acc0[c][r];
          v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
          acc0[c][r] = v;
          float u = acc1[c][r];
          u += __shfl_xor_sync(0xffffffffu, u, 4); u += __shfl_xor_sync(0xffffffffu, u, 8); u += __shfl_xor_sync(0xffffffffu, u, 16);
          acc1[c][r] = u;
        }
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int cb = blk ? c1 : c0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int n = cb + c;
          if (n < N) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < kRows; ++r) if (r == slice) sum = blk ? acc1[c][r] : acc0[c][r];
            const float v = act_fwd(sum + Ly.b[n], Ly.act);
            if (more) xb[n * kRows + slice] = v;
            if (slice < nrow) Ly.h[(size_t)(row0 + slice) * Ly.h_ld + n] = v;
          }
        }
      }
    } else {
      const int npad = (N + 31) & ~31;
      const int ks = max(1, min(8, kMlpThreads / npad));          // K slices: threads beyond one column set split the reduction
      const int slice = tid / npad, n = tid - slice * npad;
      const bool active = slice < ks && n < N;
      float acc[kRows];
  #pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
      for (int t = 0; t < ntiles; ++t) {
        wait_tile(t, bulk);                                         // tile t has landed (kStages - 2 younger tiles may be in flight)
        __syncthreads();                                           // ... for every thread; the stage of tile t-1 is free again; xa is complete
        issue(Ly.W, K, N, KT, t + kStages - 1, bulk, vec4);
        if (active) {
          const float* w = wt + (t % kStages) * kTileFloats + n;
          const int k0 = tile_at(t, ntiles) * KT, rows = min(KT, K - k0);
  #pragma unroll 4
          for (int kk = slice; kk < rows; kk += ks) {
            const float wv = w[kk * N];
            const float* xp = xa + (k0 + kk) * kRows;
            const float4 x0 = *reinterpret_cast<const float4*>(xp), x1 = *reinterpret_cast<const float4*>(xp + 4);
            acc[0] = fmaf(x0.x, wv, acc[0]); acc[1] = fmaf(x0.y, wv, acc[1]); acc[2] = fmaf(x0.z, wv, acc[2]); acc[3] = fmaf(x0.w, wv, acc[3]);
            acc[4] = fmaf(x1.x, wv, acc[4]); acc[5] = fmaf(x1.y, wv, acc[5]); acc[6] = fmaf(x1.z, wv, acc[6]); acc[7] = fmaf(x1.w, wv, acc[7]);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                             // all weight stages idle: the next layer may start streaming
      if (more) {
        const MlpLayer& Ln = A.L[l + 1];
        const bool v4 = ((reinterpret_cast<uintptr_t>(Ln.W) & 15) == 0) && (Ln.out % 4 == 0);
        const int KTn = tile_rows(Ln.in, Ln.out);
        { const bool bn = A.bulk && bulk_ok(Ln.W, Ln.in, Ln.out); for (int t = 0; t < kStages - 1; ++t) issue(Ln.W, Ln.in, Ln.out, KTn, t, bn, v4); }
      }
      if (ks > 1) {                                                // fixed-order reduction over the K slices
        if (slice > 0 && active) {
  #pragma unroll
          for (int r = 0; r < kRows; ++r) red[(size_t)tid * kRows + r] = acc[r];
        }
        __syncthreads();
        if (slice == 0 && n < N)
          for (int s2 = 1; s2 < ks; ++s2)
  #pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r] += red[(size_t)(s2 * npad + n) * kRows + r];
      }
      if (slice == 0 && n < N) {
        const float bv = Ly.b[n];
  #pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const float v = act_fwd(acc[r] + bv, Ly.act);
          if (more) xb[n * kRows + r] = v;
          if (r < nrow) Ly.h[(size_t)(row0 + r) * Ly.h_ld + n] = v;
        }
      }
    }
    if (more && A.L[l + 1].action_in) {                          // the next layer reads [h | action]
      for (int i = tid; i < kRows * A.action_dim; i += kMlpThreads) {
        const int r = i / A.action_dim, j = i - r * A.action_dim;
        const float v = r < nrow ? A.action[(size_t)(row0 + r) * A.action_dim + j] : 0.f;
        xb[(N + j) * kRows + r] = v;
        if (r < nrow) Ly.h[(size_t)(row0 + r) * Ly.h_ld + N + j] = v;
      }
    }
    float* tsw = xa; xa = xb; xb = tsw;                          // the first sync of the next layer's tile loop publishes xb
  }
  MLP_STAMP(7)
}

// ------------------------------------------------------------------------------------------ input-gradient chain
// One warp per input column k of the current weight tile, lanes over the output columns n, shuffle reduction.
// shared: da / db [maxw][kRows], wt [kStages][kTileFloats]
__global__ void __launch_bounds__(kMlpThreads) mlp_dgrad_kernel(const __grid_constant__ MlpArgs A) {
  extern __shared__ __align__(16) float sm[];
  float* da = sm;
  float* db = sm + (size_t)A.maxw * kRows;
  float* wt = db + (size_t)A.maxw * kRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = blockIdx.x * kRows;
  const int nrow = min(kRows, A.B - row0);
  const int last = A.end - 1;
  const int lowest = A.need_dx_first ? A.first : A.first + 1;      // lowest layer whose input gradient is produced
  float amx = 0.f;
  if (last >= lowest) {
    const MlpLayer& L0 = A.L[last];
    const bool vec4 = ((reinterpret_cast<uintptr_t>(L0.W) & 15) == 0) && (L0.out % 4 == 0);
    const int KT = tile_rows(L0.in, L0.out);
    for (int t = 0; t < kStages - 1; ++t) issue_tile(L0.W, L0.in, L0.out, KT, t, wt + t * kTileFloats, tid, vec4);
  }
  {   // gradient wrt the last pre-activation: d_out * act'(out)
    const MlpLayer& Ly = A.L[last];
    for (int i = tid; i < kRows * Ly.out; i += kMlpThreads) {
      const int r = i / Ly.out, n = i - r * Ly.out;
      float g = 0.f;
      if (r < nrow) {
        g = A.d_out[(size_t)(row0 + r) * Ly.out + n];
        const float y = Ly.h[(size_t)(row0 + r) * Ly.h_ld + n];
        if (Ly.act == 1) g = y > 0.f ? g : 0.f;
        else if (Ly.act == 2) g = g * (1.f - y * y);
        A.dTop[(size_t)(row0 + r) * Ly.out + n] = g;
      }
      da[n * kRows + r] = g;
    }
  }
  for (int l = last; l >= lowest; --l) {
    const MlpLayer& Ly = A.L[l];
    const int N = Ly.out, K = Ly.in;
    const bool vec4 = ((reinterpret_cast<uintptr_t>(Ly.W) & 15) == 0) && (N % 4 == 0);
    const int KT = tile_rows(K, N), ntiles = (K + KT - 1) / KT;
    const int relu_cols = l > 0 ? A.L[l - 1].out : 0;               // the first relu_cols inputs are the ReLU output of layer l-1
    const float* hprev = l > 0 ? A.L[l - 1].h : nullptr;
    const int hprev_ld = l > 0 ? A.L[l - 1].h_ld : 0;
    // ReLU mask of the layer below: its activations go into db up front (coalesced, all loads in flight); the epilogue of
    // column k reads db[k][r] and overwrites it with the gradient - no global load on the per-column critical path
    for (int k = tid; k < relu_cols; k += kMlpThreads) {
      float v[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) v[r] = r < nrow ? __ldg(hprev + (size_t)(row0 + r) * hprev_ld + k) : 0.f;
      *reinterpret_cast<float4*>(db + k * kRows) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(db + k * kRows + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int t = 0; t < ntiles; ++t) {
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();                                             // tile t visible to all; da complete; the stage of tile t-1 is free
      issue_tile(Ly.W, K, N, KT, t + kStages - 1, wt + ((t + kStages - 1) % kStages) * kTileFloats, tid, vec4);
      const float* wtile = wt + (t % kStages) * kTileFloats;
      const int k0 = t * KT, rows = min(KT, K - k0);
      for (int kk = warp; kk < rows; kk += kMlpThreads / 32) {
        const int k = k0 + kk;
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
        const float* w = wtile + kk * N;
        for (int n = lane; n < N; n += 32) {
          const float wv = w[n];
          const float4 d0 = *reinterpret_cast<const float4*>(da + n * kRows), d1 = *reinterpret_cast<const float4*>(da + n * kRows + 4);
          acc[0] = fmaf(d0.x, wv, acc[0]); acc[1] = fmaf(d0.y, wv, acc[1]); acc[2] = fmaf(d0.z, wv, acc[2]); acc[3] = fmaf(d0.w, wv, acc[3]);
          acc[4] = fmaf(d1.x, wv, acc[4]); acc[5] = fmaf(d1.y, wv, acc[5]); acc[6] = fmaf(d1.z, wv, acc[6]); acc[7] = fmaf(d1.w, wv, acc[7]);
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
        if (lane < kRows) {                                          // lane r finishes row r
          float v = 0.f;
#pragma unroll
          for (int r = 0; r < kRows; ++r) if (lane == r) v = acc[r];
          const int r = lane;
          if (k < relu_cols && !(db[k * kRows + r] > 0.f)) v = 0.f;
          if (r >= nrow) v = 0.f;
          db[k * kRows + r] = v;
          if (r < nrow) {
            Ly.dX[(size_t)(row0 + r) * K + k] = v;
            if (l == A.first) amx = fmaxf(amx, fabsf(v));
            if (Ly.action_in && A.d_action && k >= K - A.action_dim) A.d_action[(size_t)(row0 + r) * A.action_dim + (k - (K - A.action_dim))] = v;
          }
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                               // all stages idle, db complete
    if (l - 1 >= lowest) {
      const MlpLayer& Ln = A.L[l - 1];
      const bool v4 = ((reinterpret_cast<uintptr_t>(Ln.W) & 15) == 0) && (Ln.out % 4 == 0);
      const int KTn = tile_rows(Ln.in, Ln.out);
      for (int t = 0; t < kStages - 1; ++t) issue_tile(Ln.W, Ln.in, Ln.out, KTn, t, wt + t * kTileFloats, tid, v4);
    }
    float* tsw = da; da = db; db = tsw;
  }
  if (A.absmax_first != nullptr) {                                 // max is order independent: an atomic keeps the result deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
    if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(A.absmax_first), __float_as_int(amx));
  }
}

// cpp_set_option("mlp_fast"): bit 0 register-tiled inner loop for layers with >= 256 inputs, bit 1 weight tiles by bulk copies
int g_mlp_fast = [] { const char* e = getenv("CARTPOLEPP_MLP_FAST"); return e ? (atoi(e) & 7) : 3; }();
// 0: one GEMM per layer; 1: fused forward stacks only; 2: fused forward and fused input-gradient chain
static int g_use_fused_mlp = -1;
void set_fused_mlp(int on) { g_use_fused_mlp = on; }
int fused_mlp_level() {
  static const int d = [] { const char* e = getenv("CARTPOLEPP_FUSED_MLP"); return e ? atoi(e) : 1; }();
  return g_use_fused_mlp < 0 ? d : g_use_fused_mlp;
}
bool fused_mlp_enabled() { return fused_mlp_level() >= 1; }

static void fill_layers(const Net& net, const float* params, char* ws, const Net::Layout& L, MlpArgs* A) {
  int maxw = net.feat + net.action_dim;
  for (int i = 0; i < net.n_fc; ++i) {
    MlpLayer& y = A->L[i];
    y.W = params + net.off_fc_w[i]; y.b = params + net.off_fc_b[i];
    y.h = reinterpret_cast<float*>(ws + L.h[i]); y.dX = reinterpret_cast<float*>(ws + L.dX[i]);
    y.in = net.in_dim[i]; y.out = net.out_dim[i]; y.act = net.act[i]; y.h_ld = net.out_ld[i];
    y.action_in = net.concat_at == i ? 1 : 0;
    maxw = std::max(maxw, std::max(y.in, y.out + net.action_dim));
  }
  A->maxw = maxw;
  A->action_dim = net.action_dim;
}

bool mlp_fits(const Net& net) {
  int maxw = net.feat + net.action_dim;
  for (int i = 0; i < net.n_fc; ++i) maxw = std::max(maxw, std::max(net.in_dim[i], net.out_dim[i] + net.action_dim));
  return (size_t)(2 * maxw + kMlpThreads) * kRows * sizeof(float) + (size_t)kStages * kTileFloats * sizeof(float) <= 200 * 1024;
}

int launch_mlp_forward(const Net& net, const float* params, const float* action, int B, void* ws_, float* out, cudaStream_t s,
                       int first_fc, int end_fc) {
  char* ws = reinterpret_cast<char*>(ws_);
  const Net::Layout L = net.layout(B);
  if (end_fc < 0 || end_fc > net.n_fc) end_fc = net.n_fc;
  if (first_fc >= end_fc) return CPP_OK;
  MlpArgs A{};
  fill_layers(net, params, ws, L, &A);
  A.first = first_fc; A.end = end_fc; A.B = B; A.action = action;
  A.tiled = g_mlp_fast & 1; A.bulk = (g_mlp_fast >> 1) & 1; A.rotate = (g_mlp_fast >> 2) & 1;
  int ld;
  const float* x = net.fc_input(L, ws, first_fc, &ld);
  A.x = x; A.x_ld = ld;
  A.x_cols = net.in_dim[first_fc] - (net.concat_at == first_fc ? net.action_dim : 0);
  A.x_tail = net.concat_at == first_fc ? const_cast<float*>(x) + A.x_cols : nullptr;
  const size_t smem = (size_t)(2 * A.maxw + kMlpThreads) * kRows * sizeof(float) + (size_t)kStages * kTileFloats * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  mlp_forward_kernel<<<(unsigned)ceil_div(B, kRows), kMlpThreads, smem, s>>>(A);
  CPP_CHECK_LAUNCH();
  if (out != nullptr && end_fc == net.n_fc) {
    const int n = net.out_dim[net.n_fc - 1];
    CPP_CHECK_CUDA(cudaMemcpyAsync(out, ws + L.h[net.n_fc - 1], (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  return CPP_OK;
}

// gradients wrt the inputs of layers last .. stop_at (dTop and every dX[i] land in the workspace); d_action optional
int launch_mlp_dgrad(const Net& net, const float* params, int B, void* ws_, const float* d_out, int stop_at, int need_dx_first,
                     float* d_action, cudaStream_t s, float* absmax_first) {
  char* ws = reinterpret_cast<char*>(ws_);
  const Net::Layout L = net.layout(B);
  MlpArgs A{};
  fill_layers(net, params, ws, L, &A);
  A.first = stop_at; A.end = net.n_fc; A.B = B;
  A.d_out = d_out; A.dTop = reinterpret_cast<float*>(ws + L.dTop); A.d_action = d_action; A.need_dx_first = need_dx_first;
  A.absmax_first = need_dx_first ? absmax_first : nullptr;
  const size_t smem = (size_t)2 * A.maxw * kRows * sizeof(float) + (size_t)kStages * kTileFloats * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    CPP_CHECK_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  mlp_dgrad_kernel<<<(unsigned)ceil_div(B, kRows), kMlpThreads, smem, s>>>(A);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

// ------------------------------------------------------------------------------------------ critic tail
// The top of the pixel critic, hidden3 = relu([hidden2, action] . W3 + b3), q = hidden3 . wq + bq (ddpg_cartpole.py:168-171,
// 180-184), is evaluated twice per step (at mu(s1) for dQ/da, at the batch action for the TD error) and sits on the step's
// critical path between the actor's output and both backward chains.  As per-layer launches it is 4-8 dependent tiny kernels;
// here ONE kernel does a whole evaluation (one warp per batch row, W3 in shared memory):
//   forward : writes the action columns of the concat buffer, hidden3, q into the workspace (the weight gradients read them),
//             q_out, and optionally dQ/da = W3[action rows] . (wq * relu'(hidden3)) and its negation (tf.neg, :113)
//   backward: from dq, dTop = dq, dPre3 = dq * wq * relu'(hidden3), dXcat = (dPre3 . W3^T) * relu'(hidden2) in one launch
constexpr int kTailThreads = 256;
constexpr int kTailMaxH = 128, kTailMaxIn = 256;

struct TailArgs {
  float* xcat; int x_ld;            // [B][x_ld]: columns [0, D) = hidden2 output, [D, D + A) = action
  const float* action;              // forward: [B][A] copied into xcat (NULL: already there)
  const float* W3; const float* b3; // [D + A][H], [H]
  const float* wq; const float* bq; // [H][1], [1]
  float* h3; int h3_ld;             // [B][h3_ld] relu output
  float* hq;                        // [B] q in the workspace
  float* q_out; float* dqda; float* neg_dqda;   // optional
  // backward
  const float* dq;                  // [B]
  float* dTop; float* dPre3; float* dXcat;      // [B], [B][H], [B][D + A]
  int B, D, A, H;
};

__global__ void __launch_bounds__(kTailThreads) critic_tail_fwd_kernel(const __grid_constant__ TailArgs T) {
  extern __shared__ float tsm[];
  const int In = T.D + T.A, H = T.H, Hp = H | 1;
  float* W = tsm;                                   // [In][Hp]
  float* xs = W + (size_t)In * Hp;                  // [warps][In]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kTailThreads >> 5;
  for (int i = tid; i < In * H; i += kTailThreads) W[(i / H) * Hp + (i % H)] = __ldg(T.W3 + i);
  __syncthreads();
  for (int b = blockIdx.x * nwarp + warp; b < T.B; b += gridDim.x * nwarp) {
    float* x = xs + warp * In;
    for (int i = lane; i < In; i += 32) {
      float v;
      if (i >= T.D && T.action != nullptr) { v = T.action[(size_t)b * T.A + (i - T.D)]; T.xcat[(size_t)b * T.x_ld + i] = v; }
      else v = T.xcat[(size_t)b * T.x_ld + i];
      x[i] = v;
    }
    __syncwarp();
    float qacc = 0.f, da[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) da[k] = 0.f;
    for (int j = lane; j < H; j += 32) {
      float z = __ldg(T.b3 + j);
      for (int i = 0; i < In; ++i) z = fmaf(x[i], W[i * Hp + j], z);
      const float h = fmaxf(z, 0.f), wq = __ldg(T.wq + j);
      T.h3[(size_t)b * T.h3_ld + j] = h;
      qacc = fmaf(h, wq, qacc);
      if (T.dqda != nullptr && z > 0.f) {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < T.A) da[k] = fmaf(W[(T.D + k) * Hp + j], wq, da[k]);
      }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      qacc += __shfl_xor_sync(0xffffffffu, qacc, sft);
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < T.A) da[k] += __shfl_xor_sync(0xffffffffu, da[k], sft);
    }
    if (lane == 0) {
      const float q = qacc + __ldg(T.bq);
      T.hq[b] = q;
      if (T.q_out != nullptr) T.q_out[b] = q;
    }
    if (T.dqda != nullptr && lane < T.A) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k == lane) v = da[k];
      T.dqda[(size_t)b * T.A + lane] = v;
      if (T.neg_dqda != nullptr) T.neg_dqda[(size_t)b * T.A + lane] = -v;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kTailThreads) critic_tail_bwd_kernel(const __grid_constant__ TailArgs T) {
  extern __shared__ float tsm[];
  const int In = T.D + T.A, H = T.H, Hp = H | 1;
  float* W = tsm;                                   // [In][Hp]
  float* ds = W + (size_t)In * Hp;                  // [warps][H]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kTailThreads >> 5;
  for (int i = tid; i < In * H; i += kTailThreads) W[(i / H) * Hp + (i % H)] = __ldg(T.W3 + i);
  __syncthreads();
  for (int b = blockIdx.x * nwarp + warp; b < T.B; b += gridDim.x * nwarp) {
    float* d3 = ds + warp * H;
    const float dq = T.dq[b];
    if (lane == 0) T.dTop[b] = dq;                                   // q is linear
    for (int j = lane; j < H; j += 32) {
      const float g = T.h3[(size_t)b * T.h3_ld + j] > 0.f ? dq * __ldg(T.wq + j) : 0.f;
      d3[j] = g;
      T.dPre3[(size_t)b * H + j] = g;
    }
    __syncwarp();
    for (int i = lane; i < In; i += 32) {
      float acc = 0.f;
      for (int j = 0; j < H; ++j) acc = fmaf(d3[j], W[i * Hp + j], acc);
      if (i < T.D && !(T.xcat[(size_t)b * T.x_ld + i] > 0.f)) acc = 0.f;      // ReLU of hidden2; the action columns pass
      T.dXcat[(size_t)b * In + i] = acc;
    }
    __syncwarp();
  }
}

bool critic_tail_ok(const Net& net) {
  const int ca = net.concat_at, last = net.n_fc - 1;
  return g_critic_tail && fused_mlp_enabled() && ca >= 1 && last == ca + 1 && net.act[ca] == 1 && net.act[last] == 0 && net.out_dim[last] == 1 &&
         net.act[ca - 1] == 1 && net.out_dim[ca] <= kTailMaxH && net.in_dim[ca] <= kTailMaxIn && net.action_dim <= 8;
}

static void tail_common(const Net& net, const float* params, int B, char* ws, const Net::Layout& L, TailArgs* T) {
  const int ca = net.concat_at, last = net.n_fc - 1;
  int ld;
  T->xcat = const_cast<float*>(net.fc_input(L, ws, ca, &ld)); T->x_ld = ld;
  T->W3 = params + net.off_fc_w[ca]; T->b3 = params + net.off_fc_b[ca];
  T->wq = params + net.off_fc_w[last]; T->bq = params + net.off_fc_b[last];
  T->h3 = reinterpret_cast<float*>(ws + L.h[ca]); T->h3_ld = net.out_ld[ca];
  T->hq = reinterpret_cast<float*>(ws + L.h[last]);
  T->B = B; T->D = net.in_dim[ca] - net.action_dim; T->A = net.action_dim; T->H = net.out_dim[ca];
}

static int tail_grid(int B) { return (int)std::min<int64_t>(ceil_div(B, kTailThreads / 32), 4 * sm_budget()); }

int launch_critic_tail_fwd(const Net& net, const float* params, const float* action, int B, void* ws_, float* q_out, float* dqda,
                           float* neg_dqda, cudaStream_t s) {
  CPP_REQUIRE(critic_tail_ok(net), "critic tail kernel: unsupported head");
  char* ws = reinterpret_cast<char*>(ws_);
  const Net::Layout L = net.layout(B);
  TailArgs T{};
  tail_common(net, params, B, ws, L, &T);
  T.action = action; T.q_out = q_out; T.dqda = dqda; T.neg_dqda = neg_dqda;
  const size_t smem = ((size_t)(T.D + T.A) * (T.H | 1) + (size_t)(kTailThreads / 32) * (T.D + T.A)) * sizeof(float);
  static bool configured = false;
  if (!configured) { CPP_CHECK_CUDA(cudaFuncSetAttribute(critic_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); configured = true; }
  critic_tail_fwd_kernel<<<tail_grid(B), kTailThreads, smem, s>>>(T);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

int launch_critic_tail_bwd(const Net& net, const float* params, const float* dq, int B, void* ws_, cudaStream_t s) {
  CPP_REQUIRE(critic_tail_ok(net), "critic tail kernel: unsupported head");
  char* ws = reinterpret_cast<char*>(ws_);
  const Net::Layout L = net.layout(B);
  const int ca = net.concat_at, last = net.n_fc - 1;
  TailArgs T{};
  tail_common(net, params, B, ws, L, &T);
  T.dq = dq;
  T.dTop = reinterpret_cast<float*>(ws + L.dTop); T.dPre3 = reinterpret_cast<float*>(ws + L.dX[last]);
  T.dXcat = reinterpret_cast<float*>(ws + L.dX[ca]);
  const size_t smem = ((size_t)(T.D + T.A) * (T.H | 1) + (size_t)(kTailThreads / 32) * T.H) * sizeof(float);
  static bool configured = false;
  if (!configured) { CPP_CHECK_CUDA(cudaFuncSetAttribute(critic_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); configured = true; }
  critic_tail_bwd_kernel<<<tail_grid(B), kTailThreads, smem, s>>>(T);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
