/*
 * cartpolepp.h - C ABI of libcartpolepp.so: the cartpole++ RL-training hot path on B200 (sm_100a).
 *
 * The reference (matpalm/cartpoleplusplus) has no FFI: its seam is "a Python method that performs one
 * tf.Session.run(op, feed_dict)" plus ReplayMemory.batch.  Each entry point
 * below names the reference call it replaces (file:line relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer documented "dev" is CUDA device memory owned by
 *     the caller (the Python host allocates it with torch), "host" is host memory.
 *   - every function returns 0 on success or a negative cpp_status; cpp_last_error() gives the text
 *     (thread local).  Nothing throws across the ABI, nothing allocates device memory.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless documented.
 *   - single caller thread per object (the reference is single threaded).
 */
#ifndef CARTPOLEPP_H_
#define CARTPOLEPP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define CPP_ABI_VERSION 4
#define CPP_MAX_FC 8

typedef enum {
  CPP_OK = 0,
  CPP_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  CPP_ERR_CUDA = -2,         /* CUDA runtime error */
  CPP_ERR_STATE = -3,        /* object not bound / wrong call order (e.g. update_weights on a non target net) */
  CPP_ERR_NUMERICS = -4,     /* non-finite value where the reference's tf.check_numerics would raise */
  CPP_ERR_NCCL = -5
} cpp_status;

int cpp_version(void);
const char* cpp_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py's gpu_launches) */
int64_t cpp_launch_count(void);
/* The tensor-core route hands the fp32 activation between conv layers on as two fp16 pieces (hi + lo); a value above
 * 2 x 65504 saturates that copy (the fp32 output itself is exact).  Number of such values since the last reset, synchronises
 * the device; -1 on a CUDA error.  The Python engines check it after a step when CARTPOLEPP_CHECK_OVERFLOW=1. */
int64_t cpp_piece_overflow_count(int32_t reset);

/* runtime switches (tests and A/B timing): "conv1_tc" = 1 routes conv1 forward / weight gradient of fp16 states through the
 * tensor-core kernels (default, also CARTPOLEPP_CONV1=tc), 0 through the exact-fp32 CUDA-core kernels, -1 = environment default;
 * "streams" = 1 forks the independent chains of a fused DDPG step onto side streams (CARTPOLEPP_STREAMS), "graphs" = 1
 * replays a fused step as one CUDA graph (CARTPOLEPP_GRAPHS); "fused_mlp" = 1 runs every FC stack as one forward and one
 * input-gradient launch instead of one GEMM per layer (CARTPOLEPP_FUSED_MLP; 1 = forward stacks (default), 2 = also the
 * input-gradient chain); schedule of the fused DDPG step, all for A/B timing, results are unchanged up to summation order:
 * "prep_hoist" = 1 (default) runs the conv2/conv3 weight-prep kernels of the whole step at its start on idle streams,
 * "conv1_split" = 1 (default) runs the two conv1 passes side by side on half of the SMs each, "critic_tail" = 1 (default) evaluates
 * the pixel critic's [hidden2, action] -> hidden3 -> q head (incl. dQ/da) as one kernel, "bwd_critic_sms" (default 92) and
 * "fwd_actor_sms" (default 37) set the SM budgets of the critic's backward and the actor's forward chain;
 * "wgrad_tc" = mask (default 5) of the conv weight gradients that run on tcgen05 where the layer shape allows: 1 conv1 on raw
 * pixels (conv_wgrad_tc.cu), 4 conv2 / conv3 on the fp16 piece layout through the row-sweep kernel (conv_wgrad_row_tc.cu: TMA
 * tensor-map strips, dY windows built on the fly), 2 conv2 / conv3 through the older piece mode of conv_wgrad_tc.cu; 0: always the
 * mma.sync kernel; "conv_row" = mask (default 1) for the conv2 / conv3 forward and input-gradient passes: 1 the row-sweep tcgen05
 * kernel (conv_row_tc.cu: ky taps along N into a ring of TMEM slots, input strips by TMA tensor-map boxes), + 2 strips by 16-byte
 * cp.async instead of TMA, + 4 input gradient from a piece tensor written by a separate un-pool / split pass instead of the fused
 * producer warps; 0: the parity-plane kernel of conv_tc.cu everywhere (it also takes the shapes the row-sweep kernel does not
 * cover: odd sizes, rows wider than 124 pixels); "wgrad_flush_steps" (default 32) = tensor-core K-steps between two fp32 flushes of the
 * weight-gradient accumulators; "fc_tc" = mask of the fully connected passes that run on tcgen05 (fc_tc.cu): 1 forward,
 * 2 input gradient, 4 weight gradient (with the bias gradient folded in), + 8 to include GEMMs below 64 M MACs (default 0: measured
 * slower than the FFMA kernels at the BASELINE sizes, profiles/r4/fc_tc.md; CARTPOLEPP_FC_TC sets the start-up value).
 * "mlp_fast" = mask (default 3, CARTPOLEPP_MLP_FAST) for the fused FC forward kernel (mlp.cu): 1 register-tiled inner loop (4 output
 * columns x 8 rows per thread) for layers with >= 256 inputs, 2 weight tiles by bulk copies (TMA engine) instead of per-thread
 * cp.async, 4 every CTA starts its tile sequence at a different tile (measured slower, off); results differ from 0 only in
 * summation order;
 * "dropout_seed" = seed of the library's own counter-based mask generator (TensorFlow's random stream cannot be reproduced;
 * every training forward of a dropout network advances a device-side counter, so graph replays draw fresh masks),
 * "dropout_external" = 1: masks are NOT generated - the caller has written 0/1 bytes into the mask buffers
 * (cpp_*_debug_view kind 3) - for parity tests against an oracle that takes the masks as inputs;
 * "is_training" = the reference's global IS_TRAINING placeholder (base_network.py:11) for the raw cpp_net_* calls: batch (1,
 * default) or moving (0) statistics in slim.batch_norm; the agent entry points set it themselves (train ops: 1; action_given,
 * check_loss, debug_values, value_given: 0).
 * Process-wide state: set options from the (single) caller thread, between steps. */
int cpp_set_option(const char* name, int32_t value);

/* ------------------------------------------------------------------ a1: index sampling (host)
 * replaces np.random.randint(0, size, n) in ReplayMemory.random_indexes, replay_memory.py:123-129.
 * MT19937 with numpy-legacy seeding and masked rejection; bit exact with numpy's RandomState. */
typedef struct cpp_mt19937 cpp_mt19937;
int cpp_mt_create(cpp_mt19937** out);
int cpp_mt_destroy(cpp_mt19937* mt);
int cpp_mt_seed(cpp_mt19937* mt, uint32_t seed);                         /* == np.random.seed(int) */
int cpp_mt_set_state(cpp_mt19937* mt, const uint32_t* key624, int32_t pos);   /* np.random.get_state()[1:3] */
int cpp_mt_get_state(const cpp_mt19937* mt, uint32_t* key624, int32_t* pos);
int cpp_mt_randint(cpp_mt19937* mt, int64_t high, int64_t n, int64_t* out_host);

/* ------------------------------------------------------------------ a2: replay gather (device)
 * replaces the five fancy-index gathers of ReplayMemory.batch, replay_memory.py:131-138.
 * state slab fp16 [n_slots][row_elems]; s1_idx/s2_idx int32[N]; action f32[N][A]; reward/mask f32[N].
 * idxs int64[B] (dev).  Outputs: s1/s2 fp16 [B][row_elems], action [B][A], reward/mask [B]. */
int cpp_replay_gather(const void* state_slab, const int32_t* s1_idx, const int32_t* s2_idx,
                      const float* action, const float* reward, const float* mask,
                      const int64_t* idxs, int32_t B, int64_t row_elems, int32_t action_dim,
                      void* out_s1, void* out_s2, float* out_action, float* out_reward, float* out_mask,
                      void* stream);
/* per-slot per-channel (sum x, sum x^2) in fp64, computed when a state is stored (8f row 1 / 8e):
 * slot_stats f64[n_slots][2*C]; rows int32[n] lists the slots to (re)compute. */
int cpp_slot_stats(const void* state_slab, const int32_t* slots, int32_t n, int64_t n_pix, int32_t C,
                   double* slot_stats, void* stream);
/* whitening moments of a batch from the per-slot sums of the B selected slots -> mean_inv f32[2*C] */
int cpp_moments_from_slots(const double* slot_stats, const int32_t* slot_table, const int64_t* idxs,
                           int32_t B, int64_t n_pix, int32_t C, float* mean_inv, void* stream);
/* whitening moments straight from a batch (tf.nn.moments, base_network.py:95-96): x is fp16 (is_f16=1)
 * or fp32 [n_pix_total][C]; scratch f64[cpp_moments_scratch_doubles(C)]; mean_inv f32[2*C] =
 * {mean_c}, {rsqrt(var_c + 1e-6)} */
int64_t cpp_moments_scratch_doubles(int32_t C);
int cpp_channel_moments(const void* x, int32_t is_f16, int64_t n_pix_total, int32_t C,
                        double* scratch, float* mean_inv, void* stream);

/* ------------------------------------------------------------------ a3-a6: networks
 * One reference network = optional conv trunk (Network.simple_conv_net_on, base_network.py:73-127)
 * -> flatten -> FC stack (hidden_layers_starting_at :58-71) with an optional concat of the action
 * in front of FC layer `concat_at` (CriticNetwork, ddpg_cartpole.py:161-184).
 * Parameters live in ONE flat fp32 buffer in TF variable-creation order: conv{1,2,3}/{weights HWIO,
 * biases}, then per FC layer {weights [in][out], biases}. */
typedef struct {
  int32_t pixels;              /* 1: state is (H,W,Cin) image, Cin = 3*cameras*repeats; 0: flat vector */
  int32_t H, W, Cin;
  int32_t input_dim;           /* pixels==0: prod(state_shape) */
  int32_t n_fc;
  int32_t fc_out[CPP_MAX_FC];
  int32_t fc_act[CPP_MAX_FC];  /* 0 linear, 1 relu, 2 tanh */
  int32_t concat_at;           /* -1: no action input */
  int32_t action_dim;
  int32_t fc_dropout[CPP_MAX_FC];  /* 1: slim.dropout(keep_prob 0.5) after this FC layer (--use-dropout, base_network.py:69-70): under
                                * IS_TRAINING the output is multiplied by a Bernoulli(0.5) mask and by 2, else the identity */
  int32_t batch_norm;          /* --use-batch-norm (base_network.py:74-79): every conv layer is conv (no bias) -> slim.batch_norm
                                * (center, no scale, eps 1e-3) -> ReLU -> pool; variables per conv layer: weights,
                                * BatchNorm/beta, BatchNorm/moving_mean, BatchNorm/moving_variance (the last two are never
                                * updated by the reference and only copied by the target update) */
} cpp_net_spec;

typedef struct cpp_net cpp_net;
int cpp_net_create(const cpp_net_spec* spec, cpp_net** out);
int cpp_net_destroy(cpp_net* net);
int64_t cpp_net_num_params(const cpp_net* net);
int32_t cpp_net_num_vars(const cpp_net* net);
int cpp_net_var_info(const cpp_net* net, int32_t i, int64_t* offset, int32_t* ndim, int64_t* shape4);
int32_t cpp_net_feature_dim(const cpp_net* net);
/* bytes of activation workspace one forward(+backward) at batch B needs */
int64_t cpp_net_workspace_bytes(const cpp_net* net, int32_t B);

/* forward: state fp16/fp32 [B][...]; mean_inv f32[2*Cin] (pixels only; whitening stats of the batch);
 * action f32[B][A] or NULL; out f32[B][fc_out[last]].  ws is kept for a following backward. */
int cpp_net_forward(const cpp_net* net, const float* params, const void* state, int32_t state_is_f16,
                    const float* mean_inv, const float* action, int32_t B, void* ws, float* out, void* stream);
/* backward of the last forward held in ws: d_out f32[B][out] (gradient wrt the post-activation output);
 * grads f32[num_params] is OVERWRITTEN (may be NULL: only d_action wanted); d_action f32[B][A] or NULL. */
int cpp_net_backward(const cpp_net* net, const float* params, const void* state, int32_t state_is_f16,
                     const float* mean_inv, int32_t B, void* ws, const float* d_out,
                     float* grads, float* d_action, void* stream);

/* single conv layer of the trunk, for kernel-level parity tests and the roofline timing in bench.py:
 * slim.conv2d(KSxKS, 10 filters, SAME) + ReLU + slim.max_pool2d(2x2) (base_network.py:103-107).
 * x: fp16 (whitened on the fly with mean_inv) or fp32 NHWC [B][H][W][Cin]; w HWIO; pooled f32 [B][H/2][W/2][10];
 * amax u8 same shape (argmax position 0..3, 4 = ReLU closed). */
int cpp_conv_forward(const void* x, int32_t x_is_f16, const float* mean_inv, const float* w, const float* bias,
                     int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t KS, float* pooled, uint8_t* amax, void* stream);
/* gradient wrt the layer input (needs Cin == 10): dx f32 [B][H][W][10] */
int cpp_conv_dgrad(const float* d_pooled, const uint8_t* amax, const float* w, int32_t B, int32_t H, int32_t W,
                   int32_t KS, float* dx, void* stream);
/* gradient wrt weights (HWIO) and biases; scratch f32[cpp_conv_wgrad_scratch_floats()] */
int64_t cpp_conv_wgrad_scratch_floats(int32_t H, int32_t W, int32_t Cin, int32_t KS);
int cpp_conv_wgrad(const void* x, int32_t x_is_f16, const float* mean_inv, const float* d_pooled, const uint8_t* amax,
                   int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t KS, float* dw, float* db, float* scratch, void* stream);

/* the same layer on the 5th-gen tensor cores (tcgen05.mma, fp16 x fp16 -> fp32 in TMEM) for `nets` (<= 3) sibling
 * networks that read the same input in ONE pass over x: actor+critic on state_1, the two targets on state_2
 * (ddpg_cartpole.py:270-273), NAF's value/mu/l trunks (naf_cartpole.py:104,150,175).  x fp16 NHWC (the replay
 * layout, replay_memory.py:32); rows (dev, optional) = slab row of every batch image, which fuses the gather of
 * ReplayMemory.batch (replay_memory.py:134,138) into the layer; mean_inv as above (folded into the weights);
 * w/bias/pooled/amax: HOST arrays of `nets` device pointers; scratch (dev, 256-byte aligned) of
 * cpp_conv_tc_scratch_bytes() holds the packed weights.  fp32 weights enter as two fp16 pieces (22 mantissa
 * bits), pixels are exact fp16, accumulation is fp32: results agree with cpp_conv_forward to ~1e-6 relative.
 * conv2/conv3 (base_network.py:111-123): x_is_pieces = 1, x holds the fp32 activation of the layer below as fp16 pieces
 * [B][H][W][hi(Cin/2) | lo(Cin/2)], w has Cin/2 input channels, mean_inv = NULL.  pooled_hl (HOST array of `nets` device
 * pointers or NULL): the pooled output once more as fp16 pieces [B][H/2][W/2][hi(10) | lo(10)] for the next layer. */
int64_t cpp_conv_tc_scratch_bytes(int32_t nets, int32_t H, int32_t W, int32_t Cin, int32_t KS);
int cpp_conv_forward_tc(const void* x_f16, const int32_t* rows, const float* mean_inv, int32_t nets,
                        const float* const* w, const float* const* bias, int32_t B, int32_t H, int32_t W, int32_t Cin,
                        int32_t KS, float* const* pooled, uint8_t* const* amax, void* scratch, void* stream,
                        int32_t x_is_pieces, void* const* pooled_hl);

/* gradient wrt the input of a 10 -> 10 channel layer (conv2/conv3; same contract as cpp_conv_dgrad) on the tensor cores:
 * the pooled gradient is un-pooled through the arg-max side band into fp16 pieces (scaled by a power of two from its
 * max) and convolved with the flipped, transposed filters by the tcgen05 kernel in dense-output mode.
 * scratch: cpp_conv_dgrad_tc_scratch_bytes(), 256-byte aligned. */
int64_t cpp_conv_dgrad_tc_scratch_bytes(int32_t B, int32_t H, int32_t W, int32_t KS);
int cpp_conv_dgrad_tc(const float* d_pooled, const uint8_t* amax, const float* w, int32_t B, int32_t H, int32_t W, int32_t KS,
                      float* dx, void* scratch, void* stream);
/* weight and bias gradients of the same layer for `nets` (<= 3) sibling networks in ONE pass over x on the tensor cores
 * (mma.sync m16n8k16, fp16 x fp16 -> fp32): tf.gradients of the conv1 variables in ddpg_cartpole.py:111,213 /
 * naf_cartpole.py:233.  x fp16 NHWC (exact replay pixels); d_pooled/amax/dw/db: HOST arrays of `nets` device pointers;
 * every fp32 gradient enters as two fp16 pieces, the whitening is folded out through a constant-one channel.
 * x_is_pieces = 1: x holds [hi(Cin/2) | lo(Cin/2)] pieces of an fp32 activation (conv2/conv3), mean_inv must be NULL and
 * dw has Cin/2 input channels.  scratch: cpp_conv_wgrad_mma_scratch_bytes(), 256-byte aligned. */
int64_t cpp_conv_wgrad_mma_scratch_bytes(int32_t nets, int32_t H, int32_t W, int32_t Cin, int32_t KS);
int cpp_conv_wgrad_mma(const void* x_f16, const float* mean_inv, int32_t x_is_pieces, int32_t nets, const float* const* d_pooled,
                       const uint8_t* const* amax, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t KS,
                       float* const* dw, float* const* db, void* scratch, void* stream);

/* ------------------------------------------------------------------ a11-a13: clip / optimiser / target copy
 * util.clip_and_debug_gradients util.py:45-58 (tf.clip_by_global_norm): writes
 * scale = clip*min(1/||g||, 1/clip) and ||g|| to out2 f32[2] (dev); scratch f64[cpp_norm_scratch_doubles()].
 * clip <= 0 disables clipping (scale = 1). */
int64_t cpp_norm_scratch_doubles(void);
int cpp_global_norm_scale(const float* grads, int64_t n, float clip, double* scratch, float* out2, void* stream);
/* util.construct_optimiser util.py:73-76: kind 0 GradientDescent, 1 Momentum, 2 Adam.  g is multiplied by
 * *scale (dev, may be NULL).  slots: Momentum f32[n], Adam f32[2n]; opt_state f32[2] (beta powers, dev). */
int cpp_optimiser_apply(int32_t kind, float* params, const float* grads, const float* scale, int64_t n,
                        float lr, float momentum, float beta1, float beta2, float eps,
                        float* slots, float* opt_state, void* stream);
/* Network._create_variables_copy_op base_network.py:20-33: t <- t - c*(t - s) */
int cpp_soft_update(float* target, const float* source, float coeff, int64_t n, void* stream);

/* ------------------------------------------------------------------ a7-a9: DDPG agent
 * ActorNetwork.train ddpg_cartpole.py:140-145, CriticNetwork.train :230-237, check_loss :239-248,
 * action_given :121-138 (noise stays on the host), update_weights base_network.py:45-49. */
typedef struct {
  cpp_net_spec actor, critic;
  float actor_lr, critic_lr, discount, gradient_clip, target_update_rate;
  int32_t max_batch;
  int32_t world_size, rank;    /* data parallel: grads are summed over ranks before clip+apply */
} cpp_ddpg_config;

typedef struct {
  float* params;          /* [actor | critic] contiguous, fp32 */
  float* target_params;   /* [target_actor | target_critic] */
  float* grads;           /* [actor | critic | loss | nonfinite flag] : the flat all-reduce buffer */
  void* workspace; int64_t workspace_bytes;   /* >= cpp_ddpg_workspace_bytes() */
} cpp_ddpg_buffers;

typedef struct cpp_ddpg cpp_ddpg;
int cpp_ddpg_create(const cpp_ddpg_config* cfg, cpp_ddpg** out);
int cpp_ddpg_destroy(cpp_ddpg* a);
int64_t cpp_ddpg_workspace_bytes(const cpp_ddpg* a);
/* flat-buffer layout (floats): out5 = {n_actor, n_critic, offset_critic, offset_loss, total}; every part starts
 * 16-byte aligned, gaps stay zero; grads[offset_loss] = loss, [offset_loss+1] = non-finite count */
int cpp_ddpg_layout(const cpp_ddpg* a, int64_t* out5);
int cpp_ddpg_bind(cpp_ddpg* a, const cpp_ddpg_buffers* b);
/* whitening statistics: by default computed from the batch handed to each call; a data-parallel or
 * replay-resident caller may pin them (mean_inv f32[2*Cin] dev for s1 and s2) for the next calls */
int cpp_ddpg_set_moments(cpp_ddpg* a, const float* mean_inv_s1, const float* mean_inv_s2);
/* phase A: forward/backward only -> grads[actor part]; phase B: clip + SGD.  actor_train = A;B.
 * Between A and B a data-parallel host all-reduces (sum) the actor part of `grads`. */
int cpp_ddpg_actor_backward(cpp_ddpg* a, const void* s1, int32_t is_f16, int32_t B, int32_t B_global, void* stream);
int cpp_ddpg_actor_apply(cpp_ddpg* a, void* stream);
int cpp_ddpg_actor_train(cpp_ddpg* a, const void* s1, int32_t is_f16, int32_t B, void* stream);
int cpp_ddpg_critic_backward(cpp_ddpg* a, const void* s1, const float* action, const float* reward,
                             const float* mask, const void* s2, int32_t is_f16, int32_t B, int32_t B_global,
                             int32_t reuse_s1_trunk, void* stream);
int cpp_ddpg_critic_apply(cpp_ddpg* a, void* stream);
int cpp_ddpg_critic_train(cpp_ddpg* a, const void* s1, const float* action, const float* reward,
                          const float* mask, const void* s2, int32_t is_f16, int32_t B, void* stream);
/* one whole grad-step, actor.train(state_1); critic.train(batch) (ddpg_cartpole.py:332-334), as ONE backward (both flat
 * gradient parts + loss) and ONE apply: the critic gradient does not depend on the actor update, so the results are those
 * of the two reference calls while every pass over state_1 is shared.  A data-parallel host all-reduces `grads` in between. */
int cpp_ddpg_step_backward(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask,
                           const void* s2, int32_t is_f16, int32_t B, int32_t B_global, void* stream);
int cpp_ddpg_step_apply(cpp_ddpg* a, void* stream);
/* both of the above in one call (single replica): after one eager run per argument set the whole step - four chains on
 * forked streams between the shared conv1 passes - is replayed as ONE CUDA graph launch */
int cpp_ddpg_train_step(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask,
                        const void* s2, int32_t is_f16, int32_t B, void* stream);
/* out: loss f32[1], td f32[B], q f32[B] (dev) */
int cpp_ddpg_check_loss(cpp_ddpg* a, const void* s1, const float* action, const float* reward,
                        const float* mask, const void* s2, int32_t is_f16, int32_t B,
                        float* loss, float* td, float* q, void* stream);
int cpp_ddpg_action_given(cpp_ddpg* a, const void* state, int32_t is_f16, int32_t B, float* out_action, void* stream);
/* 8f row 2, the rollout latency path of ActorNetwork.action_given (ddpg_cartpole.py:121-138), called once per env step with
 * B = 1: an fp32 state is copied to fp16 on the device - exact for what the env produces, fp16(k)/255 stored in a float32 array,
 * bullet_cartpole.py:239-242 - so that it takes the tensor-core trunk, and the whole chain (copy, statistics, conv1-3, FC stack)
 * replays as ONE CUDA graph per (state buffer, output buffer).  out_action_and_flag (dev) holds B*A actions followed by one
 * float that is 1.0 when some element of the state was NOT an fp16 number: the caller must then use cpp_ddpg_action_given,
 * which keeps fp32 states exact (CUDA-core route). */
int cpp_ddpg_action_given_fast(cpp_ddpg* a, const void* state, int32_t is_f16, int32_t B, float* out_action_and_flag, void* stream);
int cpp_ddpg_update_targets(cpp_ddpg* a, float coeff, void* stream);
/* parity instrumentation (tests/test_gpu_step_pinned.py): where one intermediate of the LAST step lives inside the bound
 * workspace, so that the fp64 oracle can be evaluated with the routing (2x2 max-pool winners, ReLU gates) the GPU took.
 * part: 0 actor, 1 critic, 2 target actor, 3 target critic.
 * kind 0: routing bytes of conv layer `index` - u8 [B][PH][PW][10], 0..3 = winner of the window (dy*2+dx), 4 = ReLU closed;
 * kind 1: pooled output of conv layer `index` - f32 [B][PH][PW][10];  kind 2: output of FC layer `index` - f32 [B][ld].
 * out4 = {byte offset inside the workspace, B, elements per batch row, valid leading elements per row}. */
int cpp_ddpg_debug_view(const cpp_ddpg* a, int32_t part, int32_t kind, int32_t index, int32_t B, int64_t* out4);

/* ------------------------------------------------------------------ 8e: data parallel, the gradient all-reduce inside the step
 * The reference is a single replica; SURVEY.md 8e shards the minibatch over one process per GPU with ONE NCCL sum all-reduce of
 * the flat gradient buffer per grad-step.  cpp_nccl_unique_id (rank 0) -> the host broadcasts the 128 bytes (torch.distributed
 * is only the rendezvous) -> every rank calls cpp_*_comm_init on its bound agent.  From then on cpp_ddpg_train_step /
 * cpp_ddpg_step_backward / cpp_naf_backward sum the gradients over the replicas inside the step, after its last gradient kernel
 * and captured into the step's CUDA graph; the loss and critic/NAF gradients are scaled by 1 / (B * world).  world = 1 (or
 * never calling comm_init) is the single-replica behaviour.  NCCL is bound at run time (the libnccl.so.2 the process already
 * loaded, else the system one); CPP_ERR_NCCL when it is missing or a call fails.
 *
 * Transport 1 (default, one NVSwitch box, <= 8 ranks): our own all-reduce over NVLink peer memory.  cpp_*_p2p_prepare allocates
 * this rank's exchange block and returns its 64-byte CUDA IPC handle; the host all-gathers the handles; cpp_*_p2p_connect maps
 * the peers' blocks.  Per step ONE kernel after the last gradient kernel: 128-bit peer stores of this rank's buffer into its slot
 * in every peer's block, release flags, acquire of the peers' flags, local sum of the slots in rank order (csrc/comm.cu).
 * Transport 2: NCCL (cpp_nccl_unique_id + cpp_*_comm_init), one ncclAllReduce at the same place. */
int cpp_ddpg_p2p_prepare(cpp_ddpg* a, int32_t rank, int32_t world_size, void* out_host_handle_64_bytes);
int cpp_ddpg_p2p_connect(cpp_ddpg* a, const void* handles_world_x_64_bytes);
/* the exchange alone: the bound gradient buffer summed in place over the replicas (what the step does after its last gradient
 * kernel); for callers that run actor.train / critic.train as separate calls, tests and scripts/bench_allreduce.py */
int cpp_ddpg_all_reduce_grads(cpp_ddpg* a, void* stream);
int cpp_nccl_unique_id(void* out_host_128_bytes);
int cpp_nccl_version(int32_t* out);
int cpp_ddpg_comm_init(cpp_ddpg* a, int32_t rank, int32_t world_size, const void* unique_id_128_bytes);

/* ------------------------------------------------------------------ a10: NAF agent
 * NafNetwork.train naf_cartpole.py:264-272, debug_values :274-284, action_given :247-262,
 * ValueNetwork.value_given :111-114, target update :373. */
typedef struct {
  cpp_net_spec value, mu, l;
  float discount, gradient_clip, target_update_rate;
  int32_t optimiser;            /* 0 GradientDescent 1 Momentum 2 Adam */
  float lr, momentum, beta1, beta2, eps;
  int32_t max_batch, action_dim;
  int32_t world_size, rank;
  /* --share-input-state-representation (naf_cartpole.py:151-154,176-179): `mu` and `l` are then single-layer networks
   * (pixels = 0, input_dim = width of value's last hidden layer) on top of the value network's representation */
  int32_t share_input_state_representation;
} cpp_naf_config;

typedef struct {
  float* params;          /* [value | naf/output_action | naf/l_values] */
  float* target_params;   /* [target_value] */
  float* grads;           /* [value | mu | l | loss | nonfinite flag] */
  float* slots;           /* optimiser slots: 0 / n / 2n floats */
  float* opt_state;       /* f32[2] */
  void* workspace; int64_t workspace_bytes;
} cpp_naf_buffers;

typedef struct cpp_naf cpp_naf;
int cpp_naf_create(const cpp_naf_config* cfg, cpp_naf** out);
int cpp_naf_destroy(cpp_naf* a);
int64_t cpp_naf_workspace_bytes(const cpp_naf* a);
/* out7 = {n_value, n_mu, n_l, offset_mu, offset_l, offset_loss, total} */
int cpp_naf_layout(const cpp_naf* a, int64_t* out7);
int cpp_naf_bind(cpp_naf* a, const cpp_naf_buffers* b);
int cpp_naf_set_moments(cpp_naf* a, const float* mean_inv_s1, const float* mean_inv_s2);
int cpp_naf_backward(cpp_naf* a, const void* s1, const float* action, const float* reward, const float* mask,
                     const void* s2, int32_t is_f16, int32_t B, int32_t B_global, void* stream);
/* returns CPP_ERR_NUMERICS (after a stream sync) when check=1 and l_values/L/loss were non-finite; check=2 skips the update
 * on the device in that case without synchronising (loss_host must be NULL; the caller reads grads[offset_loss .. +1] later) */
int cpp_naf_apply(cpp_naf* a, int32_t check, float* loss_host, void* stream);
int cpp_naf_train(cpp_naf* a, const void* s1, const float* action, const float* reward, const float* mask,
                  const void* s2, int32_t is_f16, int32_t B, float* loss_host, void* stream);
/* out (dev): l_values [B][A(A+1)/2], loss[1], V[B], A[B], V2[B] */
int cpp_naf_debug_values(cpp_naf* a, const void* s1, const float* action, const float* reward, const float* mask,
                         const void* s2, int32_t is_f16, int32_t B,
                         float* l_values, float* loss, float* V, float* Aout, float* V2, void* stream);
int cpp_naf_action_given(cpp_naf* a, const void* state, int32_t is_f16, int32_t B, float* out_action, void* stream);
/* as cpp_ddpg_action_given_fast (naf_cartpole.py:247-262) */
int cpp_naf_action_given_fast(cpp_naf* a, const void* state, int32_t is_f16, int32_t B, float* out_action_and_flag, void* stream);
int cpp_naf_value_given(cpp_naf* a, const void* state, int32_t is_f16, int32_t B, float* out_value, void* stream);
int cpp_naf_update_targets(cpp_naf* a, float coeff, void* stream);
int cpp_naf_comm_init(cpp_naf* a, int32_t rank, int32_t world_size, const void* unique_id_128_bytes);
int cpp_naf_p2p_prepare(cpp_naf* a, int32_t rank, int32_t world_size, void* out_host_handle_64_bytes);
int cpp_naf_p2p_connect(cpp_naf* a, const void* handles_world_x_64_bytes);
int cpp_naf_all_reduce_grads(cpp_naf* a, void* stream);
/* as cpp_ddpg_debug_view; part: 0 value, 1 naf/output_action, 2 naf/l_values, 3 target value */
int cpp_naf_debug_view(const cpp_naf* a, int32_t part, int32_t kind, int32_t index, int32_t B, int64_t* out4);

/* ------------------------------------------------------------------ a14: LRPG
 * LikelihoodRatioPolicyGradientAgent graph lrpg_cartpole.py:80-130, train :165-182, util.standardise util.py:37-43 */
typedef struct {
  cpp_net_spec model;
  float gradient_clip;
  int32_t optimiser; float lr, momentum, beta1, beta2, eps;
  int32_t max_batch;
} cpp_lrpg_config;
typedef struct {
  float* params; float* grads; float* slots; float* opt_state;
  void* workspace; int64_t workspace_bytes;
} cpp_lrpg_buffers;
typedef struct cpp_lrpg cpp_lrpg;
int cpp_lrpg_create(const cpp_lrpg_config* cfg, cpp_lrpg** out);
int cpp_lrpg_destroy(cpp_lrpg* a);
int64_t cpp_lrpg_workspace_bytes(const cpp_lrpg* a);
int64_t cpp_lrpg_num_params(const cpp_lrpg* a);      /* buffers hold this rounded up to a multiple of 4 floats */
int cpp_lrpg_bind(cpp_lrpg* a, const cpp_lrpg_buffers* b);
/* observations f32[N][input_dim], actions int32[N], advantages f32[N] (dev) */
int cpp_lrpg_train(cpp_lrpg* a, const float* observations, const int32_t* actions, const float* advantages,
                   int32_t N, float* loss_host, void* stream);
int cpp_lrpg_logits(cpp_lrpg* a, const float* observations, int32_t N, float* logits, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif  /* CARTPOLEPP_H_ */
