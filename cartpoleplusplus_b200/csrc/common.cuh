// Shared helpers for libcartpolepp (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/cartpolepp.h"

namespace cpp {

void set_error(const char* fmt, ...);
const char* get_error();

#define CPP_CHECK_CUDA(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      cpp::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                     cudaGetErrorString(_e));                                           \
      return CPP_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

extern long long g_launch_count;    // kernels launched by this library (capi.cu)
extern int g_tc_prepped;           // 1 while the agent has already run Net::prep_trunk_tc for this step (the trunk launches skip their weight prep)
extern int g_prep_hoist;
extern int g_critic_tail;
extern int g_bwd_critic_sms;
extern int g_fwd_actor_sms;
extern int g_conv1_split;
extern int g_wgrad_flush_steps;   // mma.sync weight gradient: MMA K-steps accumulated on the tensor cores between two fp32 flushes
extern int g_wgrad_tc;            // conv weight gradients on tcgen05 (conv_wgrad_tc.cu) where supported: bit 0 conv1 on raw pixels, bit 1 conv2 / conv3 on pieces; 0: always mma.sync
extern int g_mlp_fast;         // mlp_forward_kernel variants (mlp.cu): bit 0 register-tiled inner loop, bit 1 bulk-copy weight tiles
extern int g_fc_tc;               // FC passes on tcgen05 (fc_tc.cu): 1 forward, 2 input gradient, 4 weight gradient, 8 = also GEMMs below the size where it pays
extern int g_is_training;         // base_network.py:11 IS_TRAINING: batch statistics (1) or moving statistics (0) in slim.batch_norm; dropout on / off
extern int g_dropout_seed, g_dropout_external;   // cpp_set_option: mask generator seed; 1 = masks are supplied by the caller (tests)
extern int g_cta_cap;               // SMs a persistent kernel may occupy (agents lower it while independent chains share the GPU)
static inline int sm_budget() { return g_cta_cap < 1 ? 1 : (g_cta_cap > 148 ? 148 : g_cta_cap); }
#define CPP_CHECK_LAUNCH() do { ++cpp::g_launch_count; CPP_CHECK_CUDA(cudaGetLastError()); } while (0)

#define CPP_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      cpp::set_error(__VA_ARGS__);            \
      return CPP_ERR_INVALID;                 \
    }                                         \
  } while (0)

#define CPP_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != CPP_OK) return _s; \
  } while (0)

// CARTPOLEPP_TRACE=1: eager steps record a timing event at every chain milestone and print the timeline (us since the
// start of the step) to stderr - the poor man's nsys for the fork/join schedule (capi.cu)
bool trace_enabled();
int trace_level();                  // CARTPOLEPP_TRACE: 1 = eager steps, 2 = also the timeline inside the captured graph
void trace_dump_graph();            // after a graph replay: prints the event-record nodes captured with the step
void trace_begin();
void trace_mark(const char* label, cudaStream_t st);     // no-op unless trace_begin() was called
void trace_dump();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

constexpr int kNumSMs = 148;      // B200
constexpr int kConvCout = 10;     // every conv of the reference trunk has 10 filters (base_network.py:103,111,119)

// ---- kernels launched from more than one translation unit (declared here, defined in the .cu named)

// elementwise.cu
int launch_state_to_f32(const void* state, int is_f16, int B, int dim, float* dst, int ld, cudaStream_t s);
int launch_f32_to_f16_exact(const float* src, int64_t n, __half* dst, float* inexact, cudaStream_t s);
int launch_copy_cols(const float* src, int src_ld, int B, int cols, float* dst, int dst_ld, int dst_col0, cudaStream_t s);
int launch_act_grad(const float* d_out, int d_ld, const float* out, int out_ld, int act, int B, int n, float* d_pre, int p_ld, cudaStream_t s);
int launch_heads_dgrad(const float* dmu_pre, const float* Wmu, int A, const float* dl_pre, const float* Wl, int NL, int B, int D,
                       float* out, cudaStream_t s);
int launch_add_gated(float* d, int d_ld, const float* extra, const float* x, int x_ld, int B, int n, cudaStream_t s, float scale = 1.f);
// slim.dropout(keep_prob 0.5) in training mode: h *= 2 * mask; mask u8 [B][n] is drawn here (counter-based hash of seed, *counter,
// layer, element) unless external; launch_dropout_tick advances the device-side counter once per forward
int launch_dropout(float* h, int ld, int B, int n, uint8_t* mask, const unsigned long long* counter, int layer, cudaStream_t s);
int launch_dropout_tick(unsigned long long* counter, cudaStream_t s);
int launch_colsum(const float* x, int ld, int B, int n, float* out, cudaStream_t s);
int launch_scale_copy(const float* src, float scale, int64_t n, float* dst, cudaStream_t s);
int launch_fill(float* dst, float v, int64_t n, cudaStream_t s);

// fc.cu : C[M,N] = opA(A)[M,K] * opB(B)[K,N]  (+ epilogue)
enum { EPI_NONE = 0, EPI_BIAS_ACT = 1, EPI_RELU_MASK = 2 };
struct GemmArgs {
  const float* A; int lda; int transA;      // transA: A stored [K][M]
  const float* B; int ldb; int transB;      // transB: B stored [N][K]
  float* C; int ldc;
  int M, N, K;
  int epi; const float* bias; int act;      // EPI_BIAS_ACT
  const float* aux; int aux_ld; int mask_cols;   // EPI_RELU_MASK: C[m][n] *= (aux[m][n] > 0) for n < mask_cols
  float mask_scale;                         // EPI_RELU_MASK: factor of the open positions (0 reads as 1; 2 behind a dropout layer: d(2 m relu(z)))
  float* absmax;                            // FFMA route, optional: max |C[m][n]| is atomically max-ed into this device float (the caller zeroes
                                            // it; callers that set it must check !gemm_tc_wanted first)
  float* colsum;                            // tensor-core route, transA only: also out[n] = sum_k B[k][n] (an all-ones row appended to A);
                                            // the FFMA route ignores it (callers check gemm_tc_wanted and launch colsum themselves)
};
int launch_gemm(const GemmArgs& g, cudaStream_t s);
// fc_tc.cu: the same contract on tcgen05 (three bf16 pieces per fp32 operand)
bool gemm_tc_wanted(const GemmArgs& g);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t s);

// conv.cu
struct ConvLayer {
  int H, W, Cin, KS;                // input spatial size / channels, kernel size (5 or 3); Cout = 10
  int PH() const { return H / 2; }
  int PW() const { return W / 2; }
};
// y = maxpool2x2(relu(conv_same(x, w) + b)); x is fp16 + whitening (mean_inv != NULL) or fp32
int launch_conv_fwd(const ConvLayer& L, const void* x, int x_is_f16, const float* mean_inv,
                    const float* w, const float* b, int B, float* pooled, uint8_t* amax, cudaStream_t s);
// d(x) of the layer, from the pooled-output gradient: dx f32 [B][H][W][Cin==10]
int launch_conv_dgrad(const ConvLayer& L, const float* d_pooled, const uint8_t* amax, const float* w,
                      int B, float* dx, cudaStream_t s);
// --use-batch-norm route: raw SAME conv without bias (B,H,W,10), and d(x) from a dense gradient wrt the conv output
int launch_conv_raw(const ConvLayer& L, const void* x, int x_is_f16, const float* mean_inv, const float* w, int B, float* raw,
                    cudaStream_t s);
int launch_conv_dgrad_dense(const ConvLayer& L, const float* d_conv, const float* w, int B, float* dx, cudaStream_t s);
// d(w), d(b); partials scratch f32[conv_wgrad_scratch_floats(L)]; amax == NULL: d_pooled is the DENSE gradient (B,H,W,10)
int64_t conv_wgrad_scratch_floats(const ConvLayer& L);
int launch_conv_wgrad(const ConvLayer& L, const void* x, int x_is_f16, const float* mean_inv,
                      const float* d_pooled, const uint8_t* amax, int B,
                      float* dw, float* db, float* scratch, cudaStream_t s);

}  // namespace cpp
