// Agent-level hot-path steps: DDPG (ddpg_cartpole.py:102-119,186-248,331-337), NAF
// (naf_cartpole.py:147-284,367-373), LRPG (lrpg_cartpole.py:80-130,165-182).
// Every step is "backward" (forward + loss + gradients into the flat gradient buffer) followed by
// "apply" (global-norm clip + optimiser); a data-parallel host all-reduces the flat buffer in between.
#include <algorithm>
#include <stdlib.h>
#include <string>
#include <utility>
#include <vector>
#include "agents.cuh"
#include "conv_row_tc.cuh"

namespace cpp {

static inline int64_t pad4(int64_t n) { return round_up(n, 4); }

struct Carver {
  char* base; size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<char*>(b)) {}
  template <typename T> T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += (size_t)round_up((int64_t)(count * sizeof(T)), 256);
    return p;
  }
};

// ============================================================================ DDPG
int DDPG::init(const cpp_ddpg_config& c) {
  cfg = c;
  CPP_REQUIRE(c.max_batch >= 1, "max_batch=%d", c.max_batch);
  CPP_TRY(actor.init(c.actor));
  CPP_TRY(critic.init(c.critic));
  CPP_REQUIRE(critic.concat_at >= 0 && critic.out_width() == 1, "critic must take an action and output one q value");
  CPP_REQUIRE(actor.out_width() == critic.action_dim, "actor output %d != critic action_dim %d", actor.out_width(), critic.action_dim);
  CPP_REQUIRE(actor.pixels == critic.pixels, "actor and critic read the same state");
  n_a = actor.nparams; n_c = critic.nparams;
  off_c = pad4(n_a); off_loss = off_c + pad4(n_c); total = off_loss + 4;
  return CPP_OK;
}

void DDPG::carve(void* ws, bool assign) {
  Carver cv(ws);
  const int B = cfg.max_batch, A = critic.action_dim, C = actor.pixels ? actor.spec.Cin : 1;
  char* wa = cv.take<char>(actor.workspace_bytes(B));
  char* wc = cv.take<char>(critic.workspace_bytes(B));
  const size_t wt_b = actor.workspace_bytes(B) > critic.workspace_bytes(B) ? actor.workspace_bytes(B) : critic.workspace_bytes(B);
  char* wt = cv.take<char>(wt_b);
  char* wt2 = cv.take<char>(critic.workspace_bytes(B));
  void* ts1 = cv.take<char>((size_t)trunk_group_scratch_bytes(2, actor));
  void* ts2 = cv.take<char>((size_t)trunk_group_scratch_bytes(2, actor));
  void* wgs = cv.take<char>((size_t)std::max(conv1_wgrad_group_scratch_bytes(1, actor), conv1_wgrad_group_scratch_bytes(2, actor)));
  void* wgs2 = cv.take<char>((size_t)std::max(conv1_wgrad_group_scratch_bytes(1, actor), conv1_wgrad_group_scratch_bytes(2, actor)));
  void* ts3 = cv.take<char>((size_t)trunk_group_scratch_bytes(2, actor));
  void* ts4 = cv.take<char>((size_t)trunk_group_scratch_bytes(2, actor));
  float* mu_ = cv.take<float>((size_t)B * A); float* dqda_ = cv.take<float>((size_t)B * A); float* neg_ = cv.take<float>((size_t)B * A);
  float* mu2_ = cv.take<float>((size_t)B * A);
  float* q_ = cv.take<float>(B); float* q2_ = cv.take<float>(B); float* td_ = cv.take<float>(B); float* dq_ = cv.take<float>(B);
  float* ones_ = cv.take<float>(B);
  float* mi1_ = cv.take<float>(2 * C); float* mi2_ = cv.take<float>(2 * C);
  double* msc = cv.take<double>(moments_scratch_doubles(C)); double* nsc = cv.take<double>(2 * norm_scratch_doubles());
  double* msc2 = cv.take<double>(moments_scratch_doubles(C));
  float* sc = cv.take<float>(4);
  __half* af = cv.take<__half>(actor.pixels ? (size_t)B * actor.spec.H * actor.spec.W * actor.spec.Cin : 8);
  ws_bytes = cv.off;
  if (assign) mom_scratch2 = msc2;
  if (assign) act_f16 = af;
  if (assign) {
    ws_actor = wa; ws_critic = wc; ws_target = wt; ws_target2 = wt2; tc_scr1 = ts1; tc_scr2 = ts2; wg_scr = wgs; mu = mu_;
    tcs[0] = ts1; tcs[1] = ts3; tcs[2] = ts2; tcs[3] = ts4; this->wgs[0] = wgs; this->wgs[1] = wgs2; dqda = dqda_; neg = neg_; mu2 = mu2_; q = q_; q2 = q2_; td = td_; dq = dq_;
    ones = ones_; mi1 = mi1_; mi2 = mi2_; mom_scratch = msc; norm_scratch = nsc; scale2 = sc;
  }
}

int DDPG::bind(const cpp_ddpg_buffers& b) {
  CPP_REQUIRE(b.params && b.target_params && b.grads && b.workspace, "null buffer");
  CPP_REQUIRE((((uintptr_t)b.params | (uintptr_t)b.target_params | (uintptr_t)b.grads | (uintptr_t)b.workspace) & 255) == 0,
              "buffers must be 256-byte aligned");
  carve(nullptr, false);
  CPP_REQUIRE(b.workspace_bytes >= (int64_t)ws_bytes, "workspace too small: %lld < %lld", (long long)b.workspace_bytes, (long long)ws_bytes);
  buf = b; bound = true;
  carve(b.workspace, true);
  ones_ready = false; pinned1 = pinned2 = nullptr;
  for (auto& gc : graph) gc.clear();
  graph_act.clear();
  return CPP_OK;
}

int DDPG::stats_for(const void* x, int is_f16, int B, float* dst, const float* pinned, const float** out, cudaStream_t s) {
  *out = nullptr;
  if (!actor.pixels) return CPP_OK;
  if (pinned) { *out = pinned; return CPP_OK; }
  CPP_TRY(launch_channel_moments(x, is_f16, (int64_t)B * actor.spec.H * actor.spec.W, actor.spec.Cin, mom_scratch, dst, s));
  *out = dst;
  return CPP_OK;
}

#define CPP_NEED_BOUND() do { if (!bound) { set_error("agent buffers not bound (call *_bind first)"); return CPP_ERR_STATE; } } while (0)
#define CPP_NEED_BATCH(B) CPP_REQUIRE((B) >= 1 && (B) <= cfg.max_batch, "batch %d outside [1, max_batch=%d]", (B), cfg.max_batch)

int DDPG::actor_backward(const void* s1, int is_f16, int B, int B_global, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B); (void)B_global;   // actor gradient is a batch SUM (SURVEY A-8)
  const int A = critic.action_dim;
  const float* m1;
  CPP_TRY(stats_for(s1, is_f16, B, mi1, pinned1, &m1, s));
  cur_m1 = m1;
  {  // actor and critic trunks read the same whitened state_1: conv1 of both in one tensor-core pass
    const Net* g[2] = {&actor, &critic};
    const float* pp[2] = {buf.params, buf.params + off_c};
    char* wss[2] = {ws_actor, ws_critic};
    CPP_TRY(trunk_forward_group(2, g, pp, wss, s1, is_f16, m1, B, tc_scr1, s));
  }
  CPP_TRY(actor.forward_fc(buf.params, nullptr, B, ws_actor, mu, s));
  CPP_TRY(critic.forward_fc(buf.params + off_c, mu, B, ws_critic, nullptr, s));
  if (!ones_ready) { CPP_TRY(launch_fill(ones, 1.f, cfg.max_batch, s)); ones_ready = true; }
  // q_gradients_wrt_actions: tf.gradients(q_value, input_action), ddpg_cartpole.py:220-222
  CPP_TRY(critic.backward(buf.params + off_c, s1, is_f16, m1, B, ws_critic, ones, nullptr, dqda, s));
  CPP_TRY(launch_scale_copy(dqda, -1.f, (int64_t)B * A, neg, s));                       // tf.neg(...), :113
  CPP_TRY(actor.backward(buf.params, s1, is_f16, m1, B, ws_actor, neg, buf.grads, nullptr, s, 1, wg_scr, tc_scr1));
  {
    const Net* g[1] = {&actor}; char* wss[1] = {ws_actor}; float* gr[1] = {buf.grads};
    CPP_TRY(conv1_wgrad_group(1, g, wss, gr, s1, is_f16, m1, B, wg_scr, s, 1));
  }
  critic_trunk_valid = true; trunk_B = B;
  return CPP_OK;
}

int DDPG::actor_apply(cudaStream_t s) {
  CPP_NEED_BOUND();
  CPP_TRY(launch_global_norm_scale(buf.grads, pad4(n_a), cfg.gradient_clip, norm_scratch, scale2, s));
  CPP_TRY(launch_optimiser(0, buf.params, buf.grads, scale2, pad4(n_a), cfg.actor_lr, 0, 0, 0, 0, nullptr, nullptr, nullptr, s));
  return CPP_OK;
}

int DDPG::critic_forward_td(const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                            int is_f16, int B, int B_global, bool reuse, float* td_out, float* dq_out, float* loss_flag, cudaStream_t s) {
  const float *m1, *m2;
  CPP_TRY(stats_for(s2, is_f16, B, mi2, pinned2, &m2, s));
  reuse = reuse && critic_trunk_valid && trunk_B == B && critic.concat_at > 0;
  if (reuse) m1 = cur_m1;
  else CPP_TRY(stats_for(s1, is_f16, B, mi1, pinned1, &m1, s));
  cur_m1 = m1;
  // bellman_rhs: target actor/critic on state_2, ddpg_cartpole.py:198-202
  {
    const Net* g[2] = {&actor, &critic};
    const float* pp[2] = {buf.target_params, buf.target_params + off_c};
    char* wss[2] = {ws_target, ws_target2};
    CPP_TRY(trunk_forward_group(2, g, pp, wss, s2, is_f16, m2, B, tc_scr2, s));
  }
  CPP_TRY(actor.forward_fc(buf.target_params, nullptr, B, ws_target, mu2, s));
  CPP_TRY(critic.forward_fc(buf.target_params + off_c, mu2, B, ws_target2, q2, s));
  if (!reuse) {
    const Net* g[1] = {&critic};
    const float* pp[1] = {buf.params + off_c};
    char* wss[1] = {ws_critic};
    CPP_TRY(trunk_forward_group(1, g, pp, wss, s1, is_f16, m1, B, tc_scr1, s));
  }
  CPP_TRY(critic.forward_fc(buf.params + off_c, action, B, ws_critic, q, s, reuse ? critic.concat_at : 0));
  CPP_TRY(launch_td_mse(q, q2, reward, mask, cfg.discount, B, B_global, td_out, dq_out, loss_flag, s));
  return CPP_OK;
}

int DDPG::critic_backward(const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                          int is_f16, int B, int B_global, int reuse, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  CPP_TRY(critic_forward_td(s1, action, reward, mask, s2, is_f16, B, B_global, reuse != 0, td, dq, buf.grads + off_loss, s));
  CPP_TRY(critic.backward(buf.params + off_c, s1, is_f16, cur_m1, B, ws_critic, dq, buf.grads + off_c, nullptr, s, 1, wg_scr, tc_scr1));
  {
    const Net* g[1] = {&critic}; char* wss[1] = {ws_critic}; float* gr[1] = {buf.grads + off_c};
    CPP_TRY(conv1_wgrad_group(1, g, wss, gr, s1, is_f16, cur_m1, B, wg_scr, s, 1));
  }
  critic_trunk_valid = false;
  return CPP_OK;
}

static int g_use_streams = -1, g_use_graphs = -1;     // -1: environment default (CARTPOLEPP_STREAMS / CARTPOLEPP_GRAPHS, on unless "0")
void set_step_options(int streams, int graphs) { if (streams >= -1) g_use_streams = streams; if (graphs >= -1) g_use_graphs = graphs; }
static bool env_flag(const char* name) { const char* e = getenv(name); return !(e && e[0] == '0'); }
bool step_streams_enabled();
static bool use_streams() { static const bool d = env_flag("CARTPOLEPP_STREAMS"); return g_use_streams < 0 ? d : g_use_streams != 0; }
static bool use_graphs() { static const bool d = env_flag("CARTPOLEPP_GRAPHS"); return g_use_graphs < 0 ? d : g_use_graphs != 0; }

void GraphCache::clear() {
  for (auto& g : slots) { if (g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.seen = 0; }
}
bool step_streams_enabled() { return use_streams(); }
bool step_graphs_enabled() { return use_graphs(); }

DDPG::~DDPG() {
  for (auto& gc : graph) gc.clear();
  if (streams_ready) {
    for (auto& st : side) if (st) cudaStreamDestroy(st);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (auto& a : aux) {
      if (a.stream) cudaStreamDestroy(a.stream);
      for (auto& e : a.ready) if (e) cudaEventDestroy(e);
      if (a.done) cudaEventDestroy(a.done);
    }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
  }
}

int DDPG::ensure_streams() {
  if (streams_ready) return CPP_OK;
  for (auto& st : side) CPP_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CPP_CHECK_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
  for (auto& a : aux) {
    CPP_CHECK_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    for (auto& e : a.ready) CPP_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CPP_CHECK_CUDA(cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming));
  }
  for (auto& e : ev) CPP_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  streams_ready = true;
  return CPP_OK;
}

int DDPG::step_body(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                    int B, int B_global, bool with_apply, bool multi, cudaStream_t s0) {
  const int A = critic.action_dim, ca = critic.concat_at;
  cudaStream_t sc = multi ? side[0] : s0, sta = multi ? side[1] : s0, stc = multi ? side[2] : s0;
  enum { E_FORK = 0, E_MU, E_DQDA, E_MU2, E_Q2, E_CB, E_TA };
  auto record = [&](int e, cudaStream_t st) -> int { if (multi) CPP_CHECK_CUDA(cudaEventRecord(ev[e], st)); return CPP_OK; };
  auto wait = [&](cudaStream_t st, int e) -> int { if (multi) CPP_CHECK_CUDA(cudaStreamWaitEvent(st, ev[e], 0)); return CPP_OK; };
  struct CapGuard { ~CapGuard() { g_cta_cap = kNumSMs; } } cap_guard;
  struct Tr { bool on; void mark(const char* l, cudaStream_t st) { if (on) trace_mark(l, st); } void dump() { if (on) trace_dump(); } } tr;
  tr.on = trace_enabled() && (!use_graphs() || trace_level() >= 2);
  if (tr.on) trace_begin();
  tr.mark("start", s0);
  const float* P = buf.params; const float* T = buf.target_params;

  // ---- shared passes over the pixels: whitening statistics and conv1 of {actor, critic}(s1) on s0, of {targets}(s2) on
  // sta, each on half of the GPU
  enum { E_START = 7 };
  const float *m1 = nullptr, *m2 = nullptr;
  const Net* g2[2] = {&actor, &critic};
  int tc1 = 0, tc2 = 0;
  if (!ones_ready) { CPP_TRY(launch_fill(ones, 1.f, cfg.max_batch, s0)); ones_ready = true; }
  CPP_TRY(record(E_START, s0));
  CPP_TRY(wait(sta, E_START));
  // ---- weight preparation of every conv2 / conv3 pass of the step (fp16 weight pieces, 12 small kernels) on the two
  // streams that idle until conv1 is done (critic and target critic chains): off the critical path of the four chains.
  // (Next to the persistent conv1 CTAs a prep kernel only gets left-over issue slots, ~10 us each: two streams, six each.)
  struct PrepGuard { ~PrepGuard() { g_tc_prepped = 0; } } prep_guard;
  enum { E_PREP = 8, E_PREP2 = 9 };
  if (multi && g_prep_hoist && actor.pixels && actor.tc_route(is_f16)) {
    CPP_TRY(wait(sc, E_START)); CPP_TRY(wait(stc, E_START));
    // row-sweep route (conv_row_tc.cu): the twelve packs are collected and written by ONE kernel on sc; passes the route does not
    // cover launch their own prep kernels as before
    tcr::prep_batch_begin();
    int prc = actor.prep_trunk_tc(P, B, tcs[0], true, sc);
    if (prc == CPP_OK) prc = critic.prep_trunk_tc(P + off_c, B, tcs[1], true, stc);
    if (prc == CPP_OK) prc = actor.prep_trunk_tc(T, B, tcs[2], false, sc);
    if (prc == CPP_OK) prc = critic.prep_trunk_tc(T + off_c, B, tcs[3], false, stc);
    const int frc = tcr::prep_batch_flush(sc);
    CPP_TRY(prc); CPP_TRY(frc);
    CPP_TRY(record(E_PREP, sc)); CPP_TRY(wait(stc, E_PREP)); CPP_TRY(record(E_PREP2, stc));
    g_tc_prepped = 1;
  }
  cudaStream_t sx = g_conv1_split ? sta : s0;                    // stream of the conv1 pass over state_2
  if (multi) g_cta_cap = g_conv1_split ? kNumSMs / 2 : kNumSMs;
  {
    CPP_TRY(stats_for(s1, is_f16, B, mi1, pinned1, &m1, s0));
    cur_m1 = m1;
    const float* pp[2] = {P, P + off_c}; char* wss[2] = {ws_actor, ws_critic};
    CPP_TRY(conv1_forward_group(2, g2, pp, wss, s1, is_f16, m1, B, tcs[0], s0, &tc1));
    tr.mark("s0 conv1 fwd {actor,critic}(s1) done", s0);
  }
  {
    double* keep = mom_scratch;
    if (multi) mom_scratch = mom_scratch2;                       // the two statistics passes run concurrently
    const int st = stats_for(s2, is_f16, B, mi2, pinned2, &m2, sta);
    mom_scratch = keep;
    CPP_TRY(st);
    if (sx != sta) { CPP_TRY(record(E_TA, sta)); CPP_TRY(wait(sx, E_TA)); }
    const float* pt[2] = {T, T + off_c}; char* wst[2] = {ws_target, ws_target2};
    CPP_TRY(record(E_FORK, s0));                                 // conv1(s1) done: the critic chain may start
    CPP_TRY(wait(sc, E_FORK));
    CPP_TRY(conv1_forward_group(2, g2, pt, wst, s2, is_f16, m2, B, tcs[2], sx, &tc2));
    tr.mark("sta conv1 fwd {targets}(s2) done", sx);
  }
  CPP_TRY(record(E_TA, sx));                                     // conv1(s2) done: the target chains may start
  CPP_TRY(wait(stc, E_TA));
  if (sx != sta) CPP_TRY(wait(sta, E_TA));
  if (g_tc_prepped) {
    CPP_TRY(wait(s0, E_PREP)); CPP_TRY(wait(sta, E_PREP)); CPP_TRY(wait(stc, E_PREP));
    CPP_TRY(wait(s0, E_PREP2)); CPP_TRY(wait(sta, E_PREP2)); CPP_TRY(wait(sc, E_PREP2));
  }
  const int cap_fa = g_fwd_actor_sms, cap_fo = (kNumSMs - g_fwd_actor_sms) / 3;
  if (multi) g_cta_cap = cap_fa;
  // ---- actor chain (s0): trunk tail, FC stack -> mu                       ddpg_cartpole.py:90-100
  CPP_TRY(actor.forward_trunk(P, s1, is_f16, m1, B, ws_actor, s0, tc1, tc1 ? tcs[0] : nullptr));
  tr.mark("s0 actor conv2/3 fwd done", s0);
  CPP_TRY(actor.forward_fc(P, nullptr, B, ws_actor, mu, s0));
  tr.mark("s0 actor FC fwd done (mu)", s0);
  CPP_TRY(record(E_MU, s0));
  // ---- critic chain (sc): trunk tail, FC below the action concat, then Q(s1, mu(s1)) and dQ/da      :161-184,220-222
  if (multi) g_cta_cap = cap_fo;
  CPP_TRY(critic.forward_trunk(P + off_c, s1, is_f16, m1, B, ws_critic, sc, tc1, tc1 ? tcs[1] : nullptr));
  tr.mark("sc critic conv2/3 fwd done", sc);
  if (ca > 0) CPP_TRY(critic.forward_fc(P + off_c, nullptr, B, ws_critic, nullptr, sc, 0, ca));
  CPP_TRY(wait(sc, E_MU));
  const bool tail = critic_tail_ok(critic);
  if (tail) {
    CPP_TRY(launch_critic_tail_fwd(critic, P + off_c, mu, B, ws_critic, nullptr, dqda, neg, sc));   // Q(s1, mu), dQ/da and tf.neg in one launch
  } else {
    CPP_TRY(critic.forward_fc(P + off_c, mu, B, ws_critic, nullptr, sc, ca > 0 ? ca : 0));
    CPP_TRY(critic.backward(P + off_c, s1, is_f16, m1, B, ws_critic, ones, nullptr, dqda, sc));
    CPP_TRY(launch_scale_copy(dqda, -1.f, (int64_t)B * A, neg, sc));                     // tf.neg(...), :113
  }
  tr.mark("sc critic FC fwd @mu + dQ/da done", sc);
  CPP_TRY(record(E_DQDA, sc));
  // ---- target actor chain (sta) -> mu2, target critic chain (stc) -> q2                            :198-202
  CPP_TRY(actor.forward_trunk(T, s2, is_f16, m2, B, ws_target, sta, tc2, tc2 ? tcs[2] : nullptr));
  CPP_TRY(actor.forward_fc(T, nullptr, B, ws_target, mu2, sta));
  tr.mark("sta target actor done (mu2)", sta);
  CPP_TRY(record(E_MU2, sta));
  CPP_TRY(critic.forward_trunk(T + off_c, s2, is_f16, m2, B, ws_target2, stc, tc2, tc2 ? tcs[3] : nullptr));
  if (ca > 0) CPP_TRY(critic.forward_fc(T + off_c, nullptr, B, ws_target2, nullptr, stc, 0, ca));
  CPP_TRY(wait(stc, E_MU2));
  CPP_TRY(critic.forward_fc(T + off_c, mu2, B, ws_target2, q2, stc, ca > 0 ? ca : 0));
  tr.mark("stc target critic done (q2)", stc);
  CPP_TRY(record(E_Q2, stc));
  // ---- backward chains: actor on s0, critic on sc
  const int cap_critic = g_bwd_critic_sms, cap_actor = kNumSMs - g_bwd_critic_sms;
  if (multi) g_cta_cap = cap_actor;
  CPP_TRY(wait(s0, E_DQDA));
  aux[0].used = aux[1].used = false;
  aux[0].cta_cap = cap_actor; aux[1].cta_cap = cap_critic;
  CPP_TRY(actor.backward(P, s1, is_f16, m1, B, ws_actor, neg, buf.grads, nullptr, s0, 1, wgs[0], tcs[0], multi ? &aux[0] : nullptr));
  tr.mark("s0 actor backward (FC, conv3, conv2) done", s0);
  CPP_TRY(wait(sc, E_Q2));
  if (multi) g_cta_cap = cap_critic;
  if (tail) CPP_TRY(launch_critic_tail_fwd(critic, P + off_c, action, B, ws_critic, q, nullptr, nullptr, sc));
  else CPP_TRY(critic.forward_fc(P + off_c, action, B, ws_critic, q, sc, ca > 0 ? ca : 0));   // Q(s1, a_batch): only the layers above the concat
  CPP_TRY(launch_td_mse(q, q2, reward, mask, cfg.discount, B, B_global, td, dq, buf.grads + off_loss, sc));
  tr.mark("sc critic FC @a + TD done", sc);
  CPP_TRY(critic.backward(P + off_c, s1, is_f16, m1, B, ws_critic, dq, buf.grads + off_c, nullptr, sc, 1, wgs[1], tcs[1], multi ? &aux[1] : nullptr));
  tr.mark("sc critic backward (FC, conv3, conv2) done", sc);
  CPP_TRY(record(E_CB, sc));
  CPP_TRY(wait(s0, E_CB));
  if (multi) for (auto& a : aux) if (a.used) CPP_CHECK_CUDA(cudaStreamWaitEvent(s0, a.done, 0));     // join the weight-gradient side streams
  g_cta_cap = kNumSMs;
  // ---- conv1 weight gradients of both networks in one pass over state_1; whole GPU
  {
    char* wss[2] = {ws_actor, ws_critic}; float* gr[2] = {buf.grads, buf.grads + off_c};
    CPP_TRY(conv1_wgrad_group(2, g2, wss, gr, s1, is_f16, m1, B, wgs[0], s0, 1));
  }
  tr.mark("s0 conv1 wgrad {actor,critic} done", s0);
  if (comm.active()) {                      // data parallel: the flat gradient buffer summed over the replicas (comm.cu)
    CPP_TRY(comm.all_reduce(buf.grads, total, s0));
    tr.mark("s0 all-reduce done", s0);
  }
  if (with_apply) CPP_TRY(apply_both(s0));
  tr.mark("s0 apply done", s0);
  tr.dump();
  return CPP_OK;
}

int DDPG::step(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
               int B, int B_global, bool with_apply, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  critic_trunk_valid = false;
  const bool multi = use_streams();
  if (multi || use_graphs()) CPP_TRY(ensure_streams());
  if (!use_graphs()) return step_body(s1, action, reward, mask, s2, is_f16, B, B_global, with_apply, multi, s);
  const void* const key[8] = {s1, action, reward, mask, s2, pinned1, pinned2, comm.key()};
  const int ikey[5] = {is_f16, B, B_global, (multi ? 1 : 0) | (g_fc_tc << 1) | (g_wgrad_tc << 5) | (tcr::conv_row_enabled() << 8) | (g_mlp_fast << 11), (conv1_tc_enabled() ? 1 : 0) | (fused_mlp_level() << 1) | (g_prep_hoist << 4) | (g_conv1_split << 5) | (g_critic_tail << 6) | (g_bwd_critic_sms << 8) | (g_fwd_actor_sms << 16) | ((g_wgrad_flush_steps / 16) << 24)};
  const int rc = run_graphed(graph[with_apply ? 1 : 0], key, ikey, s, cap_stream, [&](cudaStream_t st) {
    return step_body(s1, action, reward, mask, s2, is_f16, B, B_global, with_apply, multi, st);
  });
  if (rc == CPP_OK && trace_level() >= 2) trace_dump_graph();
  return rc;
}

int DDPG::step_backward(const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                        int is_f16, int B, int B_global, cudaStream_t s) {
  return step(s1, action, reward, mask, s2, is_f16, B, B_global, false, s);
}

int DDPG::apply_both(cudaStream_t s) {
  CPP_NEED_BOUND();
  critic_trunk_valid = false;
  return launch_clip_sgd_dual(buf.params, buf.grads, pad4(n_a), cfg.actor_lr, buf.params + off_c, buf.grads + off_c, pad4(n_c),
                              cfg.critic_lr, cfg.gradient_clip, norm_scratch, scale2, s);
}

int DDPG::critic_apply(cudaStream_t s) {
  CPP_NEED_BOUND();
  CPP_TRY(launch_global_norm_scale(buf.grads + off_c, pad4(n_c), cfg.gradient_clip, norm_scratch, scale2 + 2, s));
  CPP_TRY(launch_optimiser(0, buf.params + off_c, buf.grads + off_c, scale2 + 2, pad4(n_c), cfg.critic_lr, 0, 0, 0, 0, nullptr, nullptr, nullptr, s));
  critic_trunk_valid = false;
  return CPP_OK;
}

int DDPG::check_loss(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                     int B, float* loss, float* td_out, float* q_out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  CPP_TRY(critic_forward_td(s1, action, reward, mask, s2, is_f16, B, B, false, td_out, nullptr, scale2, s));
  critic_trunk_valid = false;
  CPP_CHECK_CUDA(cudaMemcpyAsync(loss, scale2, sizeof(float), cudaMemcpyDeviceToDevice, s));
  CPP_CHECK_CUDA(cudaMemcpyAsync(q_out, q, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return CPP_OK;
}

int DDPG::action_body(const void* state, int is_f16, int B, float* out, bool fast, cudaStream_t s) {
  if (fast && !is_f16 && actor.pixels) {
    CPP_TRY(launch_f32_to_f16_exact(reinterpret_cast<const float*>(state), (int64_t)B * actor.spec.H * actor.spec.W * actor.spec.Cin,
                                    act_f16, out + (size_t)B * critic.action_dim, s));
    state = act_f16; is_f16 = 1;
  } else if (fast) {
    CPP_CHECK_CUDA(cudaMemsetAsync(out + (size_t)B * critic.action_dim, 0, sizeof(float), s));
  }
  const float* m;
  CPP_TRY(stats_for(state, is_f16, B, mi2, nullptr, &m, s));     // statistics of the fed batch itself (B=1 in rollouts)
  const Net* g[1] = {&actor};
  const float* pp[1] = {buf.params};
  char* wss[1] = {ws_target};
  CPP_TRY(trunk_forward_group(1, g, pp, wss, state, is_f16, m, B, tc_scr2, s));
  CPP_TRY(actor.forward_fc(buf.params, nullptr, B, ws_target, out, s));
  return CPP_OK;
}

int DDPG::action_given(const void* state, int is_f16, int B, float* out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  return action_body(state, is_f16, B, out, false, s);
}

int DDPG::action_given_fast(const void* state, int is_f16, int B, float* out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  if (!use_graphs()) return action_body(state, is_f16, B, out, true, s);
  CPP_TRY(ensure_streams());
  const void* const key[8] = {state, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int ikey[5] = {is_f16, B, 0, 0, (conv1_tc_enabled() ? 1 : 0) | (fused_mlp_level() << 1)};
  return run_graphed(graph_act, key, ikey, s, cap_stream, [&](cudaStream_t st) { return action_body(state, is_f16, B, out, true, st); });
}

int DDPG::update_targets(float coeff, cudaStream_t s) {
  CPP_NEED_BOUND();
  return launch_soft_update(buf.target_params, buf.params, coeff, off_loss, s);
}

// ============================================================================ NAF
int NAF::init(const cpp_naf_config& c) {
  cfg = c;
  CPP_REQUIRE(c.max_batch >= 1, "max_batch=%d", c.max_batch);
  CPP_TRY(value.init(c.value)); CPP_TRY(mu.init(c.mu)); CPP_TRY(l.init(c.l));
  A = c.action_dim;
  CPP_REQUIRE(A >= 1 && A <= 8, "action_dim=%d unsupported", A);
  CPP_REQUIRE(value.out_width() == 1 && mu.out_width() == A && l.out_width() == A * (A + 1) / 2, "NAF head widths do not match action_dim");
  share = c.share_input_state_representation != 0;
  rep_dim = value.in_dim[value.n_fc - 1];
  if (share) {
    CPP_REQUIRE(!mu.pixels && !l.pixels && mu.n_fc == 1 && l.n_fc == 1 && mu.feat == rep_dim && l.feat == rep_dim,
                "shared input state representation: mu / l must be single layers on the %d-wide representation", rep_dim);
    CPP_REQUIRE(value.n_fc == 1 || value.out_ld[value.n_fc - 2] == rep_dim, "representation must be contiguous");
  } else {
    CPP_REQUIRE(value.pixels == mu.pixels && mu.pixels == l.pixels, "NAF nets read the same state");
  }
  CPP_REQUIRE(c.optimiser >= 0 && c.optimiser <= 2, "optimiser kind %d", c.optimiser);
  n_v = value.nparams; n_m = mu.nparams; n_l = l.nparams;
  off_m = pad4(n_v); off_l = off_m + pad4(n_m); off_loss = off_l + pad4(n_l); total = off_loss + 4;
  return CPP_OK;
}

void NAF::carve(void* ws, bool assign) {
  Carver cv(ws);
  const int B = cfg.max_batch, C = value.pixels ? value.spec.Cin : 1, NL = A * (A + 1) / 2;
  char* wv = cv.take<char>(value.workspace_bytes(B)); char* wm = cv.take<char>(mu.workspace_bytes(B));
  char* wl = cv.take<char>(l.workspace_bytes(B));
  char* wt = cv.take<char>(std::max(value.workspace_bytes(B), std::max(mu.workspace_bytes(B), l.workspace_bytes(B))));
  void* ts1 = cv.take<char>((size_t)trunk_group_scratch_bytes(3, value));
  void* ts2 = cv.take<char>((size_t)trunk_group_scratch_bytes(3, value));
  void* wgs = cv.take<char>((size_t)conv1_wgrad_group_scratch_bytes(3, value));
  void* wgs1 = cv.take<char>((size_t)conv1_wgrad_group_scratch_bytes(3, value));
  void* wgs2 = cv.take<char>((size_t)conv1_wgrad_group_scratch_bytes(3, value));
  void* ts3 = cv.take<char>((size_t)trunk_group_scratch_bytes(3, value));
  void* ts4 = cv.take<char>((size_t)trunk_group_scratch_bytes(3, value));
  double* msc2 = cv.take<double>(moments_scratch_doubles(value.pixels ? value.spec.Cin : 1));
  float* V_ = cv.take<float>(B); float* V2_ = cv.take<float>(B); float* mu_ = cv.take<float>((size_t)B * A); float* lv_ = cv.take<float>((size_t)B * NL);
  float* dV_ = cv.take<float>(B); float* dmu_ = cv.take<float>((size_t)B * A); float* dl_ = cv.take<float>((size_t)B * NL);
  float* mi1_ = cv.take<float>(2 * C); float* mi2_ = cv.take<float>(2 * C);
  double* msc = cv.take<double>(moments_scratch_doubles(C)); double* nsc = cv.take<double>(norm_scratch_doubles());
  float* sc = cv.take<float>(4);
  float* drep = cv.take<float>((size_t)B * rep_dim);
  __half* af = cv.take<__half>(value.pixels ? (size_t)B * value.spec.H * value.spec.W * value.spec.Cin : 8);
  ws_bytes = cv.off;
  if (assign) {
    d_rep = drep; act_f16 = af;
    ws_v = wv; ws_m = wm; ws_l = wl; ws_t = wt; tc_scr1 = ts1; tc_scr2 = ts2; wg_scr = wgs; V = V_;
    tcs[0] = ts1; tcs[1] = ts3; tcs[2] = ts4; tcs[3] = ts2; this->wgs[0] = wgs; this->wgs[1] = wgs1; this->wgs[2] = wgs2; mom_scratch2 = msc2; V2 = V2_; muo = mu_; lv = lv_; dV = dV_; dmu = dmu_; dl = dl_;
    mi1 = mi1_; mi2 = mi2_; mom_scratch = msc; norm_scratch = nsc; scale2 = sc;
  }
}

int NAF::bind(const cpp_naf_buffers& b) {
  CPP_REQUIRE(b.params && b.target_params && b.grads && b.workspace, "null buffer");
  CPP_REQUIRE(cfg.optimiser == 0 || b.slots != nullptr, "optimiser needs slots");
  CPP_REQUIRE(cfg.optimiser != 2 || b.opt_state != nullptr, "Adam needs opt_state");
  carve(nullptr, false);
  CPP_REQUIRE(b.workspace_bytes >= (int64_t)ws_bytes, "workspace too small: %lld < %lld", (long long)b.workspace_bytes, (long long)ws_bytes);
  buf = b; bound = true;
  carve(b.workspace, true);
  pinned1 = pinned2 = nullptr;
  graph.clear(); graph_act.clear();
  return CPP_OK;
}

NAF::~NAF() {
  graph.clear();
  if (streams_ready) {
    for (auto& st : side) if (st) cudaStreamDestroy(st);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (auto& a : aux) {
      if (a.stream) cudaStreamDestroy(a.stream);
      for (auto& e : a.ready) if (e) cudaEventDestroy(e);
      if (a.done) cudaEventDestroy(a.done);
    }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
  }
}

int NAF::ensure_streams() {
  if (streams_ready) return CPP_OK;
  for (auto& st : side) CPP_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CPP_CHECK_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
  for (auto& a : aux) {
    CPP_CHECK_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    for (auto& e : a.ready) CPP_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CPP_CHECK_CUDA(cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming));
  }
  for (auto& e : ev) CPP_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  streams_ready = true;
  return CPP_OK;
}

// naf.train(batch) up to the gradients (naf_cartpole.py:147-239): conv1 of {value, mu, l}(s1) in one tensor-core pass on s0
// while the target value network runs on state_2 on its own stream; then the three chains fork (trunk tail + FC), join at
// the L.L^T head, fork again for their backward passes and join for the shared conv1 weight gradient.
int NAF::backward_body(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                       int B, int B_global, bool multi, cudaStream_t s0) {
  cudaStream_t sm = multi ? side[0] : s0, sl = multi ? side[1] : s0, st = multi ? side[2] : s0;
  enum { E_START = 0, E_C1, E_MU, E_L, E_V2, E_HEAD, E_BM, E_BL };
  auto record = [&](int e, cudaStream_t x) -> int { if (multi) CPP_CHECK_CUDA(cudaEventRecord(ev[e], x)); return CPP_OK; };
  auto wait = [&](cudaStream_t x, int e) -> int { if (multi) CPP_CHECK_CUDA(cudaStreamWaitEvent(x, ev[e], 0)); return CPP_OK; };
  struct CapGuard { ~CapGuard() { g_cta_cap = kNumSMs; } } cap_guard;
  const float* P = buf.params;
  const float *m1 = nullptr, *m2 = nullptr;
  const int NL = A * (A + 1) / 2;
  CPP_TRY(record(E_START, s0));
  CPP_TRY(wait(st, E_START));
  // ---- target value network on state_2 (its own stream, a quarter of the GPU)
  if (multi) g_cta_cap = kNumSMs / 4;
  {
    double* keep = mom_scratch;
    if (multi) mom_scratch = mom_scratch2;
    const int rc = stats_for(s2, is_f16, B, mi2, pinned2, &m2, st);
    mom_scratch = keep;
    CPP_TRY(rc);
    const Net* g[1] = {&value}; const float* pp[1] = {buf.target_params}; char* wss[1] = {ws_t};
    CPP_TRY(trunk_forward_group(1, g, pp, wss, s2, is_f16, m2, B, tcs[3], st));
    CPP_TRY(value.forward_fc(buf.target_params, nullptr, B, ws_t, V2, st));
    CPP_TRY(record(E_V2, st));
  }
  // ---- weight packs of every conv2 / conv3 pass of the three online networks (forward + input gradient: twelve) in ONE launch on the
  // mu stream, which idles until conv1 is done (row-sweep route; passes it does not cover keep their own prep kernels)
  struct PrepGuard { ~PrepGuard() { g_tc_prepped = 0; } } prep_guard;
  enum { E_PREP = 8 };
  bool hoisted = false;
  if (multi && !share && g_prep_hoist && value.pixels && value.tc_route(is_f16)) {
    CPP_TRY(wait(sm, E_START));
    tcr::prep_batch_begin();
    int prc = value.prep_trunk_tc(P, B, tcs[0], true, sm);
    if (prc == CPP_OK) prc = mu.prep_trunk_tc(P + off_m, B, tcs[1], true, sm);
    if (prc == CPP_OK) prc = l.prep_trunk_tc(P + off_l, B, tcs[2], true, sm);
    const int frc = tcr::prep_batch_flush(sm);
    CPP_TRY(prc); CPP_TRY(frc);
    CPP_TRY(record(E_PREP, sm));
    hoisted = true;
  }
  // ---- conv1 of the three networks on state_1 in one pass
  if (multi) g_cta_cap = kNumSMs - kNumSMs / 4;
  CPP_TRY(stats_for(s1, is_f16, B, mi1, pinned1, &m1, s0));
  cur_m1 = m1;
  if (share) {
    // one trunk: value network, then V / mu / l as three heads on its representation          naf_cartpole.py:151-154,176-179
    const Net* g1[1] = {&value}; const float* pp1[1] = {P}; char* ws1[1] = {ws_v};
    int tc1 = 0;
    CPP_TRY(conv1_forward_group(1, g1, pp1, ws1, s1, is_f16, m1, B, tcs[0], s0, &tc1));
    CPP_TRY(value.forward_trunk(P, s1, is_f16, m1, B, ws_v, s0, tc1, tc1 ? tcs[0] : nullptr));
    CPP_TRY(value.forward_fc(P, nullptr, B, ws_v, V, s0));
    CPP_TRY(heads_forward(P, ws_v, B, muo, lv, s0));
    CPP_TRY(wait(s0, E_V2));
    CPP_TRY(launch_naf_head(V, muo, lv, action, reward, mask, V2, cfg.discount, B, A, B_global, dV, dmu, dl, nullptr,
                            buf.grads + off_loss, s0));
    const float* rep = shared_rep(ws_v, B);
    CPP_TRY(mu.backward(P + off_m, rep, 0, nullptr, B, ws_m, dmu, buf.grads + off_m, nullptr, s0));
    CPP_TRY(l.backward(P + off_l, rep, 0, nullptr, B, ws_l, dl, buf.grads + off_l, nullptr, s0));
    CPP_TRY(launch_heads_dgrad(reinterpret_cast<const float*>(ws_m + mu.layout(B).dTop), P + off_m + mu.off_fc_w[0], A,
                               reinterpret_cast<const float*>(ws_l + l.layout(B).dTop), P + off_l + l.off_fc_w[0], NL, B, rep_dim,
                               d_rep, s0));
    CPP_TRY(value.backward(P, s1, is_f16, m1, B, ws_v, dV, buf.grads, nullptr, s0, 1, wgs[0], tcs[0], nullptr, d_rep));
    g_cta_cap = kNumSMs;
    float* gr[1] = {buf.grads};
    CPP_TRY(conv1_wgrad_group(1, g1, ws1, gr, s1, is_f16, m1, B, wgs[0], s0, 1));
    return comm.all_reduce(buf.grads, total, s0);
  }
  const Net* g3[3] = {&value, &mu, &l};
  const float* pp3[3] = {P, P + off_m, P + off_l};
  char* ws3[3] = {ws_v, ws_m, ws_l};
  int tc1 = 0;
  CPP_TRY(conv1_forward_group(3, g3, pp3, ws3, s1, is_f16, m1, B, tcs[0], s0, &tc1));
  CPP_TRY(record(E_C1, s0));
  CPP_TRY(wait(sm, E_C1)); CPP_TRY(wait(sl, E_C1));
  if (hoisted) { CPP_TRY(wait(s0, E_PREP)); CPP_TRY(wait(sl, E_PREP)); g_tc_prepped = 1; }
  // ---- three forward chains
  if (multi) g_cta_cap = kNumSMs / 4;
  CPP_TRY(value.forward_trunk(P, s1, is_f16, m1, B, ws_v, s0, tc1, tc1 ? tcs[0] : nullptr));
  CPP_TRY(value.forward_fc(P, nullptr, B, ws_v, V, s0));
  CPP_TRY(mu.forward_trunk(P + off_m, s1, is_f16, m1, B, ws_m, sm, tc1, tc1 ? tcs[1] : nullptr));
  CPP_TRY(mu.forward_fc(P + off_m, nullptr, B, ws_m, muo, sm));
  CPP_TRY(record(E_MU, sm));
  CPP_TRY(l.forward_trunk(P + off_l, s1, is_f16, m1, B, ws_l, sl, tc1, tc1 ? tcs[2] : nullptr));
  CPP_TRY(l.forward_fc(P + off_l, nullptr, B, ws_l, lv, sl));
  CPP_TRY(record(E_L, sl));
  // ---- head: Q = V + A, TD target, loss, gradients wrt V / mu / l                          naf_cartpole.py:186-230
  CPP_TRY(wait(s0, E_MU)); CPP_TRY(wait(s0, E_L)); CPP_TRY(wait(s0, E_V2));
  CPP_TRY(launch_naf_head(V, muo, lv, action, reward, mask, V2, cfg.discount, B, A, B_global, dV, dmu, dl, nullptr,
                          buf.grads + off_loss, s0));
  CPP_TRY(record(E_HEAD, s0));
  CPP_TRY(wait(sm, E_HEAD)); CPP_TRY(wait(sl, E_HEAD));
  // ---- three backward chains (conv1 weight gradients deferred)
  // (every weight / bias gradient on a side stream per chain, as in the DDPG step: the chains of input gradients stay short)
  if (multi) g_cta_cap = kNumSMs / 3;
  for (auto& a : aux) { a.used = false; a.cta_cap = kNumSMs / 3; }
  CPP_TRY(value.backward(P, s1, is_f16, m1, B, ws_v, dV, buf.grads, nullptr, s0, 1, wgs[0], tcs[0], multi ? &aux[0] : nullptr));
  CPP_TRY(mu.backward(P + off_m, s1, is_f16, m1, B, ws_m, dmu, buf.grads + off_m, nullptr, sm, 1, wgs[1], tcs[1], multi ? &aux[1] : nullptr));
  CPP_TRY(record(E_BM, sm));
  CPP_TRY(l.backward(P + off_l, s1, is_f16, m1, B, ws_l, dl, buf.grads + off_l, nullptr, sl, 1, wgs[2], tcs[2], multi ? &aux[2] : nullptr));
  CPP_TRY(record(E_BL, sl));
  CPP_TRY(wait(s0, E_BM)); CPP_TRY(wait(s0, E_BL));
  if (multi) for (auto& a : aux) if (a.used) CPP_CHECK_CUDA(cudaStreamWaitEvent(s0, a.done, 0));     // join the weight-gradient side streams
  g_cta_cap = kNumSMs;
  {
    float* gr[3] = {buf.grads, buf.grads + off_m, buf.grads + off_l};
    CPP_TRY(conv1_wgrad_group(3, g3, ws3, gr, s1, is_f16, m1, B, wgs[0], s0, 1));
  }
  return comm.all_reduce(buf.grads, total, s0);      // data parallel: summed over the replicas (comm.cu); no-op for one replica
}

int NAF::stats_for(const void* x, int is_f16, int B, float* dst, const float* pinned, const float** out, cudaStream_t s) {
  *out = nullptr;
  if (!value.pixels) return CPP_OK;
  if (pinned) { *out = pinned; return CPP_OK; }
  CPP_TRY(launch_channel_moments(x, is_f16, (int64_t)B * value.spec.H * value.spec.W, value.spec.Cin, mom_scratch, dst, s));
  *out = dst;
  return CPP_OK;
}

int NAF::forward_all(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                     int B, int B_global, bool grads, float* adv_out, float* loss_flag, cudaStream_t s) {
  const float *m1, *m2;
  CPP_TRY(stats_for(s1, is_f16, B, mi1, pinned1, &m1, s));
  CPP_TRY(stats_for(s2, is_f16, B, mi2, pinned2, &m2, s));
  cur_m1 = m1;
  if (share) {
    const Net* g[1] = {&value}; const float* pp[1] = {buf.params}; char* wss[1] = {ws_v};
    CPP_TRY(trunk_forward_group(1, g, pp, wss, s1, is_f16, m1, B, tc_scr1, s));
    CPP_TRY(value.forward_fc(buf.params, nullptr, B, ws_v, V, s));
    CPP_TRY(heads_forward(buf.params, ws_v, B, muo, lv, s));
  } else {
    {  // value / mu / l trunks read the same whitened state_1 (naf_cartpole.py:104,150,175)
      const Net* g[3] = {&value, &mu, &l};
      const float* pp[3] = {buf.params, buf.params + off_m, buf.params + off_l};
      char* wss[3] = {ws_v, ws_m, ws_l};
      CPP_TRY(trunk_forward_group(3, g, pp, wss, s1, is_f16, m1, B, tc_scr1, s));
    }
    CPP_TRY(value.forward_fc(buf.params, nullptr, B, ws_v, V, s));
    CPP_TRY(mu.forward_fc(buf.params + off_m, nullptr, B, ws_m, muo, s));
    CPP_TRY(l.forward_fc(buf.params + off_l, nullptr, B, ws_l, lv, s));
  }
  {  // target_value_net on state_2, naf_cartpole.py:225-227
    const Net* g[1] = {&value};
    const float* pp[1] = {buf.target_params};
    char* wss[1] = {ws_t};
    CPP_TRY(trunk_forward_group(1, g, pp, wss, s2, is_f16, m2, B, tc_scr2, s));
  }
  CPP_TRY(value.forward_fc(buf.target_params, nullptr, B, ws_t, V2, s));
  CPP_TRY(launch_naf_head(V, muo, lv, action, reward, mask, V2, cfg.discount, B, A, B_global,
                          grads ? dV : nullptr, dmu, dl, adv_out, loss_flag, s));
  return CPP_OK;
}

int NAF::backward(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                  int B, int B_global, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  const bool multi = step_streams_enabled(), graphs = step_graphs_enabled();
  if (multi || graphs) CPP_TRY(ensure_streams());
  if (!graphs) return backward_body(s1, action, reward, mask, s2, is_f16, B, B_global, multi, s);
  const void* const key[8] = {s1, action, reward, mask, s2, pinned1, pinned2, comm.key()};
  const int ikey[5] = {is_f16, B, B_global, (multi ? 1 : 0) | (g_fc_tc << 1) | (g_wgrad_tc << 5) | (tcr::conv_row_enabled() << 8) | (g_mlp_fast << 11), (conv1_tc_enabled() ? 1 : 0) | (fused_mlp_level() << 1) | (g_prep_hoist << 4) | ((g_wgrad_flush_steps / 16) << 24)};
  return run_graphed(graph, key, ikey, s, cap_stream, [&](cudaStream_t x) {
    return backward_body(s1, action, reward, mask, s2, is_f16, B, B_global, multi, x);
  });
}

int NAF::apply(int check, float* loss_host, cudaStream_t s) {
  CPP_NEED_BOUND();
  CPP_TRY(launch_global_norm_scale(buf.grads, off_loss, cfg.gradient_clip, norm_scratch, scale2, s));
  // the update is skipped on device when the non-finite flag is set (tf.check_numerics aborts the step)
  CPP_TRY(launch_optimiser(cfg.optimiser, buf.params, buf.grads, scale2, off_loss, cfg.lr, cfg.momentum, cfg.beta1, cfg.beta2,
                           cfg.eps, buf.slots, buf.opt_state, check ? buf.grads + off_loss + 1 : nullptr, s));
  if (loss_host != nullptr || check == 1) {    // check == 2: device-side skip only, no host synchronisation
    float lf[2];
    CPP_CHECK_CUDA(cudaMemcpyAsync(lf, buf.grads + off_loss, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CPP_CHECK_CUDA(cudaStreamSynchronize(s));
    if (loss_host) *loss_host = lf[0];
    if (check == 1 && (lf[1] != 0.f || !isfinite(lf[0]))) {
      set_error("check_numerics: non-finite l_values / L / loss (naf_cartpole.py:242-245)");
      return CPP_ERR_NUMERICS;
    }
  }
  return CPP_OK;
}

int NAF::debug_values(const void* s1, const float* action, const float* reward, const float* mask, const void* s2, int is_f16,
                      int B, float* l_out, float* loss, float* V_out, float* A_out, float* V2_out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  CPP_TRY(forward_all(s1, action, reward, mask, s2, is_f16, B, B, false, A_out, scale2, s));
  const int NL = A * (A + 1) / 2;
  CPP_CHECK_CUDA(cudaMemcpyAsync(l_out, lv, (size_t)B * NL * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CPP_CHECK_CUDA(cudaMemcpyAsync(loss, scale2, sizeof(float), cudaMemcpyDeviceToDevice, s));
  CPP_CHECK_CUDA(cudaMemcpyAsync(V_out, V, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CPP_CHECK_CUDA(cudaMemcpyAsync(V2_out, V2, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return CPP_OK;
}

int NAF::action_body(const void* state, int is_f16, int B, float* out, bool fast, cudaStream_t s) {
  if (fast && !is_f16 && value.pixels) {
    CPP_TRY(launch_f32_to_f16_exact(reinterpret_cast<const float*>(state), (int64_t)B * value.spec.H * value.spec.W * value.spec.Cin,
                                    act_f16, out + (size_t)B * A, s));
    state = act_f16; is_f16 = 1;
  } else if (fast) {
    CPP_CHECK_CUDA(cudaMemsetAsync(out + (size_t)B * A, 0, sizeof(float), s));
  }
  const float* m;
  CPP_TRY(stats_for(state, is_f16, B, mi2, nullptr, &m, s));
  if (share) {    // value trunk and hidden layers, then the mu head
    const Net* g[1] = {&value}; const float* pp[1] = {buf.params}; char* wss[1] = {ws_t};
    CPP_TRY(trunk_forward_group(1, g, pp, wss, state, is_f16, m, B, tc_scr2, s));
    if (value.n_fc > 1) CPP_TRY(value.forward_fc(buf.params, nullptr, B, ws_t, nullptr, s, 0, value.n_fc - 1));
    return mu.forward(buf.params + off_m, shared_rep(ws_t, B), 0, nullptr, nullptr, B, ws_m, out, s);
  }
  const Net* g[1] = {&mu};
  const float* pp[1] = {buf.params + off_m};
  char* wss[1] = {ws_t};
  CPP_TRY(trunk_forward_group(1, g, pp, wss, state, is_f16, m, B, tc_scr2, s));
  return mu.forward_fc(buf.params + off_m, nullptr, B, ws_t, out, s);
}

int NAF::action_given(const void* state, int is_f16, int B, float* out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  return action_body(state, is_f16, B, out, false, s);
}

int NAF::action_given_fast(const void* state, int is_f16, int B, float* out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  if (!step_graphs_enabled()) return action_body(state, is_f16, B, out, true, s);
  CPP_TRY(ensure_streams());
  const void* const key[8] = {state, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int ikey[5] = {is_f16, B, 0, 0, (conv1_tc_enabled() ? 1 : 0) | (fused_mlp_level() << 1)};
  return run_graphed(graph_act, key, ikey, s, cap_stream, [&](cudaStream_t st) { return action_body(state, is_f16, B, out, true, st); });
}

int NAF::value_given(const void* state, int is_f16, int B, float* out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(B);
  const float* m;
  CPP_TRY(stats_for(state, is_f16, B, mi2, nullptr, &m, s));
  const Net* g[1] = {&value};
  const float* pp[1] = {buf.params};
  char* wss[1] = {ws_t};
  CPP_TRY(trunk_forward_group(1, g, pp, wss, state, is_f16, m, B, tc_scr2, s));
  return value.forward_fc(buf.params, nullptr, B, ws_t, out, s);
}

const float* NAF::shared_rep(char* wsv, int B) const {
  int ld;
  return value.fc_input(value.layout(B), wsv, value.n_fc - 1, &ld);
}

int NAF::heads_forward(const float* P, char* wsv, int B, float* mu_out, float* l_out, cudaStream_t s) {
  const float* rep = shared_rep(wsv, B);
  CPP_TRY(mu.forward(P + off_m, rep, 0, nullptr, nullptr, B, ws_m, mu_out, s));
  return l.forward(P + off_l, rep, 0, nullptr, nullptr, B, ws_l, l_out, s);
}

int NAF::update_targets(float coeff, cudaStream_t s) {
  CPP_NEED_BOUND();
  return launch_soft_update(buf.target_params, buf.params, coeff, pad4(n_v), s);
}

// ============================================================================ LRPG
int LRPG::init(const cpp_lrpg_config& c) {
  cfg = c;
  CPP_REQUIRE(c.max_batch >= 1, "max_batch=%d", c.max_batch);
  CPP_TRY(model.init(c.model));
  CPP_REQUIRE(!model.pixels, "lrpg_cartpole.py:42 asserts low-dim state");
  n = model.nparams;
  return CPP_OK;
}

void LRPG::carve(void* ws, bool assign) {
  Carver cv(ws);
  const int B = cfg.max_batch, K = model.out_width();
  char* wm = cv.take<char>(model.workspace_bytes(B));
  float* lg = cv.take<float>((size_t)B * K); float* dlg = cv.take<float>((size_t)B * K);
  double* nsc = cv.take<double>(norm_scratch_doubles()); float* sc = cv.take<float>(4);
  ws_bytes = cv.off;
  if (assign) { ws_m = wm; logits = lg; dlogits = dlg; norm_scratch = nsc; scale2 = sc; }
}

int LRPG::bind(const cpp_lrpg_buffers& b) {
  CPP_REQUIRE(b.params && b.grads && b.workspace, "null buffer");
  CPP_REQUIRE(cfg.optimiser == 0 || b.slots != nullptr, "optimiser needs slots");
  CPP_REQUIRE(cfg.optimiser != 2 || b.opt_state != nullptr, "Adam needs opt_state");
  carve(nullptr, false);
  CPP_REQUIRE(b.workspace_bytes >= (int64_t)ws_bytes, "workspace too small");
  buf = b; bound = true;
  carve(b.workspace, true);
  return CPP_OK;
}

int LRPG::train(const float* obs, const int32_t* actions, const float* adv, int N, float* loss_host, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(N);
  CPP_TRY(model.forward(buf.params, obs, 0, nullptr, nullptr, N, ws_m, logits, s));
  CPP_TRY(launch_lrpg_loss(logits, actions, adv, N, model.out_width(), dlogits, scale2 + 2, s));
  CPP_TRY(model.backward(buf.params, obs, 0, nullptr, N, ws_m, dlogits, buf.grads, nullptr, s));
  CPP_TRY(launch_global_norm_scale(buf.grads, pad4(n), cfg.gradient_clip, norm_scratch, scale2, s));
  CPP_TRY(launch_optimiser(cfg.optimiser, buf.params, buf.grads, scale2, pad4(n), cfg.lr, cfg.momentum, cfg.beta1, cfg.beta2,
                           cfg.eps, buf.slots, buf.opt_state, nullptr, s));
  if (loss_host) {
    CPP_CHECK_CUDA(cudaMemcpyAsync(loss_host, scale2 + 2, sizeof(float), cudaMemcpyDeviceToHost, s));
    CPP_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  return CPP_OK;
}

int LRPG::get_logits(const float* obs, int N, float* out, cudaStream_t s) {
  CPP_NEED_BOUND(); CPP_NEED_BATCH(N);
  return model.forward(buf.params, obs, 0, nullptr, nullptr, N, ws_m, out, s);
}

}  // namespace cpp
