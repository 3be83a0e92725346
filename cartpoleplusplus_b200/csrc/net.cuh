// One reference network (base_network.Network subclass): optional conv trunk -> flatten -> FC stack.
#pragma once
#include <vector>
#include <algorithm>
#include "common.cuh"

namespace cpp {

struct VarInfo { int64_t offset; int ndim; int64_t shape[4]; };

// Side stream for the work of a backward pass that nothing downstream in the same pass waits for (every weight / bias
// gradient): the chain of input gradients stays short on the main stream, the caller joins `done` before it reads grads.
struct BackwardAux {
  cudaStream_t stream = nullptr;
  cudaEvent_t ready[CPP_MAX_FC + 3] = {};   // "the gradient this weight gradient consumes is ready", one per layer
  cudaEvent_t done = nullptr;               // recorded on `stream` after the last weight gradient
  bool used = false;
  int cta_cap = 0;                          // SM budget of the persistent kernels launched on the side stream (0: unchanged)
};

struct Net {
  cpp_net_spec spec;
  bool pixels = false;
  bool bn = false;                    // --use-batch-norm: conv layers are raw conv -> slim.batch_norm -> ReLU -> pool (bn.cu), no conv bias
  ConvLayer conv[3];
  int feat = 0;                       // flattened trunk features, or input_dim
  int n_fc = 0;
  int in_dim[CPP_MAX_FC], out_dim[CPP_MAX_FC], act[CPP_MAX_FC], out_ld[CPP_MAX_FC];
  int concat_at = -1, action_dim = 0;
  int drop[CPP_MAX_FC] = {};          // slim.dropout after this FC layer (--use-dropout)
  bool any_drop = false;
  int64_t off_conv_w[3], off_conv_b[3], off_fc_w[CPP_MAX_FC], off_fc_b[CPP_MAX_FC];   // off_conv_b: biases, or BatchNorm/beta
  int64_t off_bn_mean[3], off_bn_var[3];                                              // BatchNorm/moving_mean, moving_variance
  int64_t nparams = 0;
  std::vector<VarInfo> vars;

  struct Layout {
    size_t pooled[3], amax[3], hl[2], x0, h[CPP_MAX_FC], dX[CPP_MAX_FC], dTop, dpool[2], wgrad, dyp, gsc, total;
    size_t mask[CPP_MAX_FC], dropctr;           // dropout: u8 masks [B][out] of the last training forward, device-side draw counter
    size_t raw[3], bnscr[3], dconv, bnjunk;     // batch norm: raw conv outputs, statistics scratch, dense d(conv), sink for the unused bias gradient
  };

  int init(const cpp_net_spec& s);
  Layout layout(int B) const;
  size_t workspace_bytes(int B) const { return layout(B).total; }
  int out_width() const { return out_dim[n_fc - 1]; }

  // pointer to the input of FC layer i inside ws, and its leading dimension
  const float* fc_input(const Layout& L, char* ws, int i, int* ld) const;

  // first_fc == 0: full forward.  first_fc == k > 0: activations below FC layer k are reused from the
  // previous forward held in ws (k must be <= concat_at or the action unchanged); action is re-copied.
  int forward(const float* params, const void* state, int is_f16, const float* mean_inv, const float* action,
              int B, void* ws, float* out, cudaStream_t s, int first_fc = 0) const;
  // the two halves of forward(): conv trunk (first_conv == 1: conv1 is already in ws, written by the tensor-core
  // kernel) or the fp16/fp32 -> fp32 state copy of a low-dim network; then the FC stack from layer first_fc
  // tc_scratch != NULL with first_conv == 1: conv1 came from the tensor-core kernel together with its fp16 piece copy
  // (Layout::hl[0]); conv2/conv3 then run on the tensor cores from the pieces as well.
  int forward_trunk(const float* params, const void* state, int is_f16, const float* mean_inv, int B, void* ws,
                    cudaStream_t s, int first_conv = 0, void* tc_scratch = nullptr) const;
  // does this network take the tensor-core route for a state of this dtype (same answer in forward and backward)
  bool tc_route(int is_f16) const;
  // weight-prep kernels of conv2 / conv3 forward (and, with_dgrad, of their input-gradient passes) into their slots of
  // tc_scratch; forward_trunk / backward then launch main kernels only while g_tc_prepped is set.  The weights must not change
  // in between (they do not inside one step: the optimiser runs last).
  int prep_trunk_tc(const float* params, int B, void* tc_scratch, bool with_dgrad, cudaStream_t s) const;
  // layers [first_fc, end_fc) (end_fc < 0: to the last one; `out` is only written when the last layer is included)
  int forward_fc(const float* params, const float* action, int B, void* ws, float* out, cudaStream_t s, int first_fc = 0,
                 int end_fc = -1) const;
  // grads == nullptr: only d_action is produced (stops at the concat layer).  defer_conv1: stop in front of conv1's
  // weight gradient (its input gradient d(pooled1) stays in ws) so that conv1_wgrad_group can do it for all siblings.
  // d_rep_extra [B][in_dim[last]]: gradient that other heads sharing the input of the LAST FC layer send into it (NAF's
  // --share-input-state-representation); added to this network's own before the ReLU gate of the layer below.
  // wg_scratch != NULL on the tensor-core route: conv2/conv3 weight gradients from the fp16 piece copies (conv_wgrad_mma.cu);
  // tc_scratch != NULL: conv3/conv2 input gradients through the tcgen05 kernel in dgrad mode (conv_tc.cu).
  int backward(const float* params, const void* state, int is_f16, const float* mean_inv, int B, void* ws,
               const float* d_out, float* grads, float* d_action, cudaStream_t s, int defer_conv1 = 0,
               void* wg_scratch = nullptr, void* tc_scratch = nullptr, BackwardAux* aux = nullptr,
               const float* d_rep_extra = nullptr) const;
};

// Conv trunks of n (<= 3) sibling networks that read the SAME state (actor+critic on state_1, the two targets on
// state_2, ddpg_cartpole.py:270-273; NAF's value/mu/l, naf_cartpole.py:104,150,175): conv1 of all of them in one
// tcgen05 pass over the pixels (conv_tc.cu) when the state is fp16 and tc_scratch is given, conv2/conv3 per net.
// Falls back to the exact-fp32 CUDA-core conv1 for fp32 states (action_given from the env) or CARTPOLEPP_CONV1=ffma.
// tc_scratch = kTcSlots slots of trunk_slot_bytes(net): 0 conv1 (all siblings), 1 conv2 fwd, 2 conv3 fwd, 3 conv3 dgrad, 4 conv2 dgrad
constexpr int kTcSlots = 5;
int64_t trunk_slot_bytes(const Net& net);
int64_t trunk_group_scratch_bytes(int n, const Net& net);
int trunk_forward_group(int n, const Net* const* nets, const float* const* params, char* const* ws, const void* state,
                        int is_f16, const float* mean_inv, int B, void* tc_scratch, cudaStream_t s);
// the two halves of trunk_forward_group, for callers that run the per-network tails on separate streams:
// conv1 of all siblings (*took_tc = 1 when the tensor-core kernel ran), then Net::forward_trunk(first_conv = *took_tc)
int conv1_forward_group(int n, const Net* const* nets, const float* const* params, char* const* ws, const void* state,
                        int is_f16, const float* mean_inv, int B, void* tc_scratch, cudaStream_t s, int* took_tc);
bool conv1_tc_enabled();
void set_conv1_tc_enabled(int on);     // -1: back to the CARTPOLEPP_CONV1 environment default
// conv1 weight/bias gradients of n sibling networks whose backward passes were run with defer_conv1 (tensor cores,
// conv_wgrad_mma.cu, when the state is fp16 and scratch is given; the exact-fp32 CUDA-core kernel otherwise)
int64_t conv1_wgrad_group_scratch_bytes(int n, const Net& net);
// gmax_from_dgrad: the backward passes ran conv2's input gradient on the tensor cores (tc_scratch given), which left
// max|d(pooled1)| of every network in its workspace - no separate max pass is needed
int conv1_wgrad_group(int n, const Net* const* nets, char* const* ws, float* const* grads, const void* state, int is_f16,
                      const float* mean_inv, int B, void* scratch, cudaStream_t s, int gmax_from_dgrad = 0);

// mlp.cu: fused FC stacks (one launch per network and direction instead of one GEMM per layer)
bool fused_mlp_enabled();
int fused_mlp_level();                // 0: GEMM per layer, 1: fused forward stacks, 2: + fused input-gradient chain
void set_fused_mlp(int on);           // -1: CARTPOLEPP_FUSED_MLP environment default
bool mlp_fits(const Net& net);
int launch_mlp_forward(const Net& net, const float* params, const float* action, int B, void* ws, float* out, cudaStream_t s,
                       int first_fc, int end_fc);
int launch_mlp_dgrad(const Net& net, const float* params, int B, void* ws, const float* d_out, int stop_at, int need_dx_first,
                     float* d_action, cudaStream_t s, float* absmax_first = nullptr);

// critic tail (mlp.cu): [hidden2, action] -> hidden3 -> q of the pixel critic in one launch per evaluation / per backward
bool critic_tail_ok(const Net& net);
int launch_critic_tail_fwd(const Net& net, const float* params, const float* action, int B, void* ws, float* q_out, float* dqda,
                           float* neg_dqda, cudaStream_t s);
// needs d_out [B] (= dq); leaves dTop, dX[last], dX[concat_at] in the workspace exactly like the per-layer path
int launch_critic_tail_bwd(const Net& net, const float* params, const float* dq, int B, void* ws, cudaStream_t s);

// bn.cu (--use-batch-norm, base_network.py:74-79)
int64_t bn_scratch_bytes();
int launch_bn_forward(const float* raw, const float* beta, const float* moving_mean, const float* moving_var, int training, int B,
                      int H, int W, void* scratch, float* pooled, uint8_t* amax, cudaStream_t s);
int launch_bn_backward(const float* d_pooled, const uint8_t* amax, const float* raw, int B, int H, int W, void* scratch,
                       float* dconv, float* dbeta, cudaStream_t s);

// elementwise.cu
int64_t moments_scratch_doubles(int C);
int launch_channel_moments(const void* x, int is_f16, int64_t n_pix_total, int C, double* scratch, float* mean_inv, cudaStream_t s);
int launch_td_mse(const float* q, const float* q2, const float* reward, const float* mask, float gamma, int B, int B_global,
                  float* td, float* dq, float* loss_flag, cudaStream_t s);
int launch_naf_head(const float* V, const float* mu, const float* lv, const float* u, const float* reward, const float* mask,
                    const float* V2, float gamma, int B, int A, int B_global, float* dV, float* dmu, float* dl,
                    float* adv, float* loss_flag, cudaStream_t s);
int launch_lrpg_loss(const float* logits, const int32_t* actions, const float* adv, int N, int K, float* dlogits, float* loss, cudaStream_t s);
int64_t norm_scratch_doubles();
int launch_global_norm_scale(const float* grads, int64_t n, float clip, double* scratch, float* out2, cudaStream_t s);
int launch_optimiser(int kind, float* params, const float* grads, const float* scale, int64_t n, float lr, float momentum,
                     float beta1, float beta2, float eps, float* slots, float* opt_state, const float* skip, cudaStream_t s);
int launch_soft_update(float* target, const float* source, float coeff, int64_t n, cudaStream_t s);
int launch_clip_sgd_dual(float* p0, const float* g0, int64_t n0, float lr0, float* p1, const float* g1, int64_t n1, float lr1,
                         float clip, double* scratch, float* out4, cudaStream_t s);
int launch_gather(const void* slab, const int32_t* s1_idx, const int32_t* s2_idx, const float* action, const float* reward,
                  const float* mask, const int64_t* idxs, int B, int64_t row_elems, int A, void* o1, void* o2, float* oa,
                  float* orw, float* om, cudaStream_t s);
int launch_slot_stats(const void* slab, const int32_t* slots, int n, int64_t n_pix, int C, double* stats, cudaStream_t s);
int launch_moments_from_slots(const double* stats, const int32_t* slot_table, const int64_t* idxs, int B, int64_t n_pix, int C,
                              float* mean_inv, cudaStream_t s);

}  // namespace cpp
