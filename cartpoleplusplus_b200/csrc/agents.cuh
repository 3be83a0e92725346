// Agent-level step objects behind the C ABI (see include/cartpolepp.h).
#pragma once
#include "net.cuh"
#include "comm.cuh"

namespace cpp {

void set_step_options(int streams, int graphs);   // -1 = environment default, -2 = leave unchanged

// One captured step per argument set: a few sets are kept (double-buffered staging alternates between two input buffer
// sets), least recently used goes.  First call with a new set runs eagerly (validates, configures kernels), the second
// captures the fork/join structure on `cap_stream`, later calls replay it with one cudaGraphLaunch.
struct GraphSlot {
  cudaGraphExec_t exec = nullptr;
  const void* key[8] = {};
  int ikey[5] = {};
  int seen = 0, used = 0;
  int launches = 0;                                     // kernels inside the captured step (for cpp_launch_count)
};
struct GraphCache {
  GraphSlot slots[4];
  int clock = 0;
  void clear();
  ~GraphCache() { clear(); }
};
// body(stream) enqueues the step; returns a cpp_status
template <typename Body>
int run_graphed(GraphCache& gc, const void* const (&key)[8], const int (&ikey)[5], cudaStream_t s, cudaStream_t cap_stream, Body body) {
  GraphSlot* hit = nullptr; GraphSlot* lru = &gc.slots[0];
  for (auto& c : gc.slots) {
    bool eq = c.seen > 0;
    for (int i = 0; i < 8 && eq; ++i) eq = c.key[i] == key[i];
    for (int i = 0; i < 5 && eq; ++i) eq = c.ikey[i] == ikey[i];
    if (eq) { hit = &c; break; }
    if (c.used < lru->used) lru = &c;
  }
  const bool same = hit != nullptr;
  GraphSlot& G = same ? *hit : *lru;
  G.used = ++gc.clock;
  if (same && G.exec != nullptr) { CPP_CHECK_CUDA(cudaGraphLaunch(G.exec, s)); g_launch_count += G.launches; return CPP_OK; }
  if (!same) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    for (int i = 0; i < 8; ++i) G.key[i] = key[i];
    for (int i = 0; i < 5; ++i) G.ikey[i] = ikey[i];
    G.seen = 1;
    return body(s);
  }
  CPP_CHECK_CUDA(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
  const long long launches_before = g_launch_count;
  const int st = body(cap_stream);
  cudaGraph_t gr = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cap_stream, &gr);
  if (st != CPP_OK) { if (gr) cudaGraphDestroy(gr); G.seen = 0; return st; }
  if (ce != cudaSuccess) { G.seen = 0; set_error("graph capture failed: %s", cudaGetErrorString(ce)); return CPP_ERR_CUDA; }
  G.launches = (int)(g_launch_count - launches_before);
  const cudaError_t ie = cudaGraphInstantiate(&G.exec, gr, 0);
  cudaGraphDestroy(gr);
  if (ie != cudaSuccess) { G.exec = nullptr; G.seen = 0; set_error("graph instantiate failed: %s", cudaGetErrorString(ie)); return CPP_ERR_CUDA; }
  CPP_CHECK_CUDA(cudaGraphLaunch(G.exec, s));
  return CPP_OK;
}
bool step_streams_enabled();
bool step_graphs_enabled();

struct DDPG {
  cpp_ddpg_config cfg;
  Net actor, critic;
  int64_t n_a = 0, n_c = 0, off_c = 0, off_loss = 0, total = 0;
  cpp_ddpg_buffers buf{};
  bool bound = false;
  size_t ws_bytes = 0;
  // carved workspace
  char *ws_actor = nullptr, *ws_critic = nullptr, *ws_target = nullptr, *ws_target2 = nullptr;
  void *tc_scr1 = nullptr, *tc_scr2 = nullptr;   // packed conv1 weights of the state_1 / state_2 trunk groups (conv_tc.cu)
  void* wg_scr = nullptr;                          // conv1 weight-gradient partials (conv_wgrad_mma.cu)
  float *mu = nullptr, *dqda = nullptr, *neg = nullptr, *mu2 = nullptr, *q = nullptr, *q2 = nullptr, *td = nullptr, *dq = nullptr;
  float *ones = nullptr, *mi1 = nullptr, *mi2 = nullptr, *scale2 = nullptr;
  double *mom_scratch = nullptr, *mom_scratch2 = nullptr, *norm_scratch = nullptr;
  const float *pinned1 = nullptr, *pinned2 = nullptr, *cur_m1 = nullptr;
  bool ones_ready = false, critic_trunk_valid = false;
  int trunk_B = 0;

  int init(const cpp_ddpg_config& c);
  void carve(void* ws, bool assign);
  int bind(const cpp_ddpg_buffers& b);
  int stats_for(const void* x, int is_f16, int B, float* dst, const float* pinned, const float** out, cudaStream_t s);
  int actor_backward(const void* s1, int is_f16, int B, int B_global, cudaStream_t s);
  int actor_apply(cudaStream_t s);
  int critic_forward_td(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                        int B, int B_global, bool reuse, float* td_out, float* dq_out, float* loss_flag, cudaStream_t s);
  int critic_backward(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                      int B, int B_global, int reuse, cudaStream_t s);
  int critic_apply(cudaStream_t s);
  int apply_both(cudaStream_t s);     // actor_apply; critic_apply in three launches
  // one whole grad-step (ddpg_cartpole.py:332-334) as backward-of-both then apply-of-both: the critic gradient does not
  // depend on the actor update, so actor.train(s1); critic.train(batch) can share every pass over state_1
  int step_backward(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                    int B, int B_global, cudaStream_t s);
  // with_apply: clip + SGD of both networks appended (single-GPU train step in ONE graph launch)
  int step(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
           int B, int B_global, bool with_apply, cudaStream_t s);
  // the enqueue-only body of a step: four independent chains (actor, critic, target actor, target critic) forked onto
  // side streams between the shared conv1 passes; captured into a CUDA graph after one eager run
  int step_body(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                int B, int B_global, bool with_apply, bool multi, cudaStream_t s);
  int ensure_streams();
  ~DDPG();
  // streams / graph state
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaStream_t cap_stream = nullptr;                      // origin stream of the graph capture (the caller's may be the legacy default stream)
  cudaEvent_t ev[10] = {};
  Comm comm;                                              // data parallel: the gradient all-reduce runs inside the step (comm.cu)
  BackwardAux aux[2];                                     // side streams of the actor / critic backward chains (weight gradients)
  bool streams_ready = false;
  void* tcs[4] = {nullptr, nullptr, nullptr, nullptr};   // packed-weight scratch per chain (actor, critic, target actor, target critic)
  void* wgs[2] = {nullptr, nullptr};                      // weight-gradient partials per backward chain
  GraphCache graph[2];                                    // [0] backward only, [1] backward + apply
  int check_loss(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16, int B,
                 float* loss, float* td_out, float* q_out, cudaStream_t s);
  int action_given(const void* state, int is_f16, int B, float* out, cudaStream_t s);
  // the rollout path (B = 1 per env step): an fp32 state is copied to fp16 on the device (exact for the env's k/255 pixels), runs
  // the tensor-core trunk, and the whole chain replays as one CUDA graph.  out [B*A + 1]: actions, then 1.0 if some element of the
  // state was NOT an fp16 number (the caller must then use action_given, which keeps fp32 states exact)
  int action_given_fast(const void* state, int is_f16, int B, float* out, cudaStream_t s);
  int action_body(const void* state, int is_f16, int B, float* out, bool fast, cudaStream_t s);
  __half* act_f16 = nullptr;                              // fp16 copy of a fed fp32 state (action_given_fast)
  GraphCache graph_act;
  int update_targets(float coeff, cudaStream_t s);
};

struct NAF {
  cpp_naf_config cfg;
  Net value, mu, l;
  int A = 2;
  bool share = false;                 // --share-input-state-representation: mu / l are heads on value's last hidden layer
  int rep_dim = 0;
  float* d_rep = nullptr;             // [B][rep_dim] gradient the two heads send into the shared representation
  int64_t n_v = 0, n_m = 0, n_l = 0, off_m = 0, off_l = 0, off_loss = 0, total = 0;
  cpp_naf_buffers buf{};
  bool bound = false;
  size_t ws_bytes = 0;
  char *ws_v = nullptr, *ws_m = nullptr, *ws_l = nullptr, *ws_t = nullptr;
  void *tc_scr1 = nullptr, *tc_scr2 = nullptr, *wg_scr = nullptr;
  float *V = nullptr, *V2 = nullptr, *muo = nullptr, *lv = nullptr, *dV = nullptr, *dmu = nullptr, *dl = nullptr;
  float *mi1 = nullptr, *mi2 = nullptr, *scale2 = nullptr;
  double *mom_scratch = nullptr, *norm_scratch = nullptr;
  const float *pinned1 = nullptr, *pinned2 = nullptr, *cur_m1 = nullptr;

  int init(const cpp_naf_config& c);
  void carve(void* ws, bool assign);
  int bind(const cpp_naf_buffers& b);
  int stats_for(const void* x, int is_f16, int B, float* dst, const float* pinned, const float** out, cudaStream_t s);
  int forward_all(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16, int B,
                  int B_global, bool grads, float* adv_out, float* loss_flag, cudaStream_t s);
  int backward(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16, int B,
               int B_global, cudaStream_t s);
  int apply(int check, float* loss_host, cudaStream_t s);
  int debug_values(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16, int B,
                   float* l_out, float* loss, float* V_out, float* A_out, float* V2_out, cudaStream_t s);
  int action_given(const void* state, int is_f16, int B, float* out, cudaStream_t s);
  int value_given(const void* state, int is_f16, int B, float* out, cudaStream_t s);
  int action_given_fast(const void* state, int is_f16, int B, float* out, cudaStream_t s);     // as DDPG::action_given_fast
  int action_body(const void* state, int is_f16, int B, float* out, bool fast, cudaStream_t s);
  __half* act_f16 = nullptr;
  GraphCache graph_act;
  int update_targets(float coeff, cudaStream_t s);
  // share mode: mu and l heads on the representation the value network left in `wsv`
  const float* shared_rep(char* wsv, int B) const;
  int heads_forward(const float* P, char* wsv, int B, float* mu_out, float* l_out, cudaStream_t s);
  // fused backward: value / mu / l chains and the target value chain on forked streams, one CUDA graph per argument set
  int backward_body(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16, int B,
                    int B_global, bool multi, cudaStream_t s);
  int ensure_streams();
  ~NAF();
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};     // mu chain, l chain, target value chain
  cudaStream_t cap_stream = nullptr;
  cudaEvent_t ev[10] = {};
  Comm comm;                                              // data parallel, as in DDPG
  bool streams_ready = false;
  void* tcs[4] = {nullptr, nullptr, nullptr, nullptr};    // packed-weight scratch per chain (value, mu, l, target value)
  void* wgs[3] = {nullptr, nullptr, nullptr};
  BackwardAux aux[3];                                     // side streams of the value / mu / l backward chains (weight gradients)
  double* mom_scratch2 = nullptr;
  GraphCache graph;
};

struct LRPG {
  cpp_lrpg_config cfg;
  Net model;
  int64_t n = 0;
  cpp_lrpg_buffers buf{};
  bool bound = false;
  size_t ws_bytes = 0;
  char* ws_m = nullptr;
  float *logits = nullptr, *dlogits = nullptr, *scale2 = nullptr;
  double* norm_scratch = nullptr;

  int init(const cpp_lrpg_config& c);
  void carve(void* ws, bool assign);
  int bind(const cpp_lrpg_buffers& b);
  int train(const float* obs, const int32_t* actions, const float* adv, int N, float* loss_host, cudaStream_t s);
  int get_logits(const float* obs, int N, float* out, cudaStream_t s);
};

}  // namespace cpp
