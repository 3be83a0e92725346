// extern "C" surface of libcartpolepp.so (include/cartpolepp.h).  Thin: argument checks, object
// lifetime, exception firewall.
#include <stdarg.h>
#include <stdlib.h>
#include <string>
#include <utility>
#include <vector>
#include <new>
#include "agents.cuh"
#include "conv_tc.cuh"
#include "conv_row_tc.cuh"
#include "conv_wgrad_mma.cuh"
#include <string.h>

namespace cpp {
long long g_launch_count = 0;
int g_cta_cap = 148;
int g_tc_prepped = 0;
int g_conv1_split = 1;     // the two conv1 passes of the DDPG step side by side on half of the SMs each (1) or one after the other on all (0)
int g_fwd_actor_sms = 37;  // SM budget of the actor's forward chain (mu is needed first); the other three chains share the rest
int g_bwd_critic_sms = 92; // SM budget of the critic's backward chain in the fused DDPG step (the actor's gets the rest of the 148); round 5
                           // A/B (profiles/r5/ab_r5s.txt): 74: 0.459, 88-96: 0.4495, 104: 0.454, 120: 0.486 ms - the critic's chain starts later (TD target)
int g_critic_tail = 1;     // the pixel critic's [hidden2, action] -> hidden3 -> q head as one kernel per evaluation / backward (mlp.cu)
int g_is_training = 1;
int g_dropout_seed = 1, g_dropout_external = 0;
int g_wgrad_tc = 5;     // bit 0: conv1 on tcgen05 (conv_wgrad_tc.cu); bit 2: conv2 / conv3 on the row-sweep tcgen05 kernel (conv_wgrad_row_tc.cu); bit 1: the older piece-mode kernel (parity green, +40 us per c3 step: profiles/r4/wgrad_tc.md)
int g_fc_tc = [] { const char* e = getenv("CARTPOLEPP_FC_TC"); return e ? (atoi(e) & 15) : 0; }();   // off by default: measured slower than the FFMA kernels at every BASELINE size (profiles/r4/fc_tc.md)
int g_wgrad_flush_steps = 32;     // the tensor-core accumulator truncates: 128-step chains cost 1.3e-5 on the conv1 weight gradient, 32 keep it at 5e-6 (profiles/r3/wgrad_flush.md)
int g_prep_hoist = 1;      // cpp_set_option("prep_hoist", 0): weight prep kernels stay in front of their main kernels (A/B timing)
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static bool g_trace_active = false;
static std::vector<std::pair<std::string, cudaEvent_t>> g_trace_pts;      // eager step: recreated every step
static std::vector<std::pair<std::string, cudaEvent_t>> g_graph_pts;      // captured step: event-record nodes inside the graph
int trace_level() { static const int t = [] { const char* e = getenv("CARTPOLEPP_TRACE"); return e ? atoi(e) : 0; }(); return t; }
bool trace_enabled() { return trace_level() >= 1; }
void trace_begin() { g_trace_active = true; }
void trace_mark(const char* label, cudaStream_t st) {
  if (!g_trace_active) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  if (cs == cudaStreamCaptureStatusActive) {
    // inside a capture a plain record is only a dependency edge; the external flag makes it a timed event-record node
    if (cudaEventRecordWithFlags(e, st, cudaEventRecordExternal) != cudaSuccess) { cudaEventDestroy(e); return; }
    g_graph_pts.emplace_back(label, e);
  } else {
    cudaEventRecord(e, st);
    g_trace_pts.emplace_back(label, e);
  }
}
static void dump_list(std::vector<std::pair<std::string, cudaEvent_t>>& pts, const char* title) {
  if (pts.empty()) return;
  cudaDeviceSynchronize();
  fprintf(stderr, "---- %s (us)\n", title);
  for (auto& p : pts) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pts[0].second, p.second) == cudaSuccess) fprintf(stderr, "%9.1f  %s\n", ms * 1e3f, p.first.c_str());
  }
}
void trace_dump() {
  if (!g_trace_active) return;
  g_trace_active = false;
  if (!g_trace_pts.empty()) {
    dump_list(g_trace_pts, "step timeline");
    for (auto& p : g_trace_pts) cudaEventDestroy(p.second);
    g_trace_pts.clear();
  }
}
void trace_dump_graph() { dump_list(g_graph_pts, "step timeline inside the CUDA graph"); }
}  // namespace cpp

using namespace cpp;

struct cpp_net { Net n; };
struct cpp_ddpg { DDPG a; };
struct cpp_naf { NAF a; };
struct cpp_lrpg { LRPG a; };

#define ST(s) reinterpret_cast<cudaStream_t>(s)
// the reference feeds ONE global IS_TRAINING placeholder per Session.run (base_network.py:11): train ops True, the rest False
struct TrainingMode { int saved; explicit TrainingMode(int v) : saved(g_is_training) { g_is_training = v; } ~TrainingMode() { g_is_training = saved; } };
#define API_BEGIN try {
#define API_END } catch (const std::exception& e) { set_error("exception: %s", e.what()); return CPP_ERR_INVALID; } \
                  catch (...) { set_error("unknown exception"); return CPP_ERR_INVALID; }
#define NEED(p) do { if (!(p)) { set_error("%s: null argument `%s`", __func__, #p); return CPP_ERR_INVALID; } } while (0)

extern "C" {

int cpp_version(void) { return CPP_ABI_VERSION; }
const char* cpp_last_error(void) { return get_error(); }
int64_t cpp_launch_count(void) { return g_launch_count; }

int64_t cpp_piece_overflow_count(int32_t reset) {
  unsigned int n = 0;
  if (tc::piece_overflow_count(reset, &n) != CPP_OK) return -1;
  return (int64_t)n;
}

int cpp_set_option(const char* name, int32_t value) {
  API_BEGIN
  NEED(name);
  if (strcmp(name, "conv1_tc") == 0) { set_conv1_tc_enabled(value); return CPP_OK; }
  if (strcmp(name, "fused_mlp") == 0) { set_fused_mlp(value); return CPP_OK; }
  if (strcmp(name, "streams") == 0) { set_step_options(value, -2); return CPP_OK; }
  if (strcmp(name, "graphs") == 0) { set_step_options(-2, value); return CPP_OK; }
  if (strcmp(name, "prep_hoist") == 0) { g_prep_hoist = value != 0; return CPP_OK; }
  if (strcmp(name, "wgrad_flush_steps") == 0) { g_wgrad_flush_steps = value < 16 ? 16 : (value > 1024 ? 1024 : value); return CPP_OK; }
  if (strcmp(name, "dropout_seed") == 0) { g_dropout_seed = value; return CPP_OK; }
  if (strcmp(name, "dropout_external") == 0) { g_dropout_external = value != 0; return CPP_OK; }
  if (strcmp(name, "is_training") == 0) { g_is_training = value != 0; return CPP_OK; }
  if (strcmp(name, "fc_tc") == 0) { g_fc_tc = value & 15; return CPP_OK; }
  if (strcmp(name, "wgrad_tc") == 0) { g_wgrad_tc = value & 7; return CPP_OK; }
  if (strcmp(name, "conv_row") == 0) { tcr::set_conv_row(value); return CPP_OK; }
  if (strcmp(name, "mlp_fast") == 0) { g_mlp_fast = value & 7; return CPP_OK; }
  if (strcmp(name, "conv1_split") == 0) { g_conv1_split = value != 0; return CPP_OK; }
  if (strcmp(name, "critic_tail") == 0) { g_critic_tail = value != 0; return CPP_OK; }
  if (strcmp(name, "fwd_actor_sms") == 0) { g_fwd_actor_sms = value < 16 ? 16 : (value > 100 ? 100 : value); return CPP_OK; }
  if (strcmp(name, "bwd_critic_sms") == 0) { g_bwd_critic_sms = value < 16 ? 16 : (value > 132 ? 132 : value); return CPP_OK; }
  set_error("unknown option `%s`", name);
  return CPP_ERR_INVALID;
  API_END
}

// ---------------------------------------------------------------- replay / moments
int cpp_replay_gather(const void* slab, const int32_t* s1_idx, const int32_t* s2_idx, const float* action, const float* reward,
                      const float* mask, const int64_t* idxs, int32_t B, int64_t row_elems, int32_t action_dim, void* out_s1,
                      void* out_s2, float* out_action, float* out_reward, float* out_mask, void* stream) {
  API_BEGIN
  if (B == 0) return CPP_OK;
  NEED(slab); NEED(s1_idx); NEED(s2_idx); NEED(action); NEED(reward); NEED(mask); NEED(idxs);
  NEED(out_s1); NEED(out_s2); NEED(out_action); NEED(out_reward); NEED(out_mask);
  CPP_REQUIRE(B > 0 && row_elems > 0 && action_dim > 0, "gather: bad sizes");
  return launch_gather(slab, s1_idx, s2_idx, action, reward, mask, idxs, B, row_elems, action_dim, out_s1, out_s2, out_action,
                       out_reward, out_mask, ST(stream));
  API_END
}
int cpp_slot_stats(const void* slab, const int32_t* slots, int32_t n, int64_t n_pix, int32_t C, double* slot_stats, void* stream) {
  API_BEGIN
  NEED(slab); NEED(slots); NEED(slot_stats);
  return launch_slot_stats(slab, slots, n, n_pix, C, slot_stats, ST(stream));
  API_END
}
int cpp_moments_from_slots(const double* slot_stats, const int32_t* slot_table, const int64_t* idxs, int32_t B, int64_t n_pix,
                           int32_t C, float* mean_inv, void* stream) {
  API_BEGIN
  NEED(slot_stats); NEED(slot_table); NEED(idxs); NEED(mean_inv);
  CPP_REQUIRE(B >= 1 && C >= 1 && C <= 1024, "moments_from_slots: bad sizes");
  return launch_moments_from_slots(slot_stats, slot_table, idxs, B, n_pix, C, mean_inv, ST(stream));
  API_END
}
int64_t cpp_moments_scratch_doubles(int32_t C) { return moments_scratch_doubles(C); }
int cpp_channel_moments(const void* x, int32_t is_f16, int64_t n_pix_total, int32_t C, double* scratch, float* mean_inv, void* stream) {
  API_BEGIN
  NEED(x); NEED(scratch); NEED(mean_inv);
  CPP_REQUIRE(n_pix_total >= 1, "moments: empty batch");
  return launch_channel_moments(x, is_f16, n_pix_total, C, scratch, mean_inv, ST(stream));
  API_END
}

// ---------------------------------------------------------------- networks
int cpp_net_create(const cpp_net_spec* spec, cpp_net** out) {
  API_BEGIN
  NEED(spec); NEED(out);
  cpp_net* n = new (std::nothrow) cpp_net();
  NEED(n);
  const int st = n->n.init(*spec);
  if (st != CPP_OK) { delete n; return st; }
  *out = n;
  return CPP_OK;
  API_END
}
int cpp_net_destroy(cpp_net* net) { delete net; return CPP_OK; }
int64_t cpp_net_num_params(const cpp_net* net) { return net ? net->n.nparams : -1; }
int32_t cpp_net_num_vars(const cpp_net* net) { return net ? (int32_t)net->n.vars.size() : -1; }
int cpp_net_var_info(const cpp_net* net, int32_t i, int64_t* offset, int32_t* ndim, int64_t* shape4) {
  API_BEGIN
  NEED(net); NEED(offset); NEED(ndim); NEED(shape4);
  CPP_REQUIRE(i >= 0 && i < (int)net->n.vars.size(), "var index %d out of range", i);
  const VarInfo& v = net->n.vars[i];
  *offset = v.offset; *ndim = v.ndim;
  for (int k = 0; k < 4; ++k) shape4[k] = v.shape[k];
  return CPP_OK;
  API_END
}
int32_t cpp_net_feature_dim(const cpp_net* net) { return net ? net->n.feat : -1; }
int64_t cpp_net_workspace_bytes(const cpp_net* net, int32_t B) { return net ? (int64_t)net->n.workspace_bytes(B) : -1; }
int cpp_net_forward(const cpp_net* net, const float* params, const void* state, int32_t is_f16, const float* mean_inv,
                    const float* action, int32_t B, void* ws, float* out, void* stream) {
  API_BEGIN
  NEED(net); NEED(params); NEED(state); NEED(ws);
  return net->n.forward(params, state, is_f16, mean_inv, action, B, ws, out, ST(stream));
  API_END
}
int cpp_net_backward(const cpp_net* net, const float* params, const void* state, int32_t is_f16, const float* mean_inv, int32_t B,
                     void* ws, const float* d_out, float* grads, float* d_action, void* stream) {
  API_BEGIN
  NEED(net); NEED(params); NEED(state); NEED(ws); NEED(d_out);
  CPP_REQUIRE(grads != nullptr || d_action != nullptr, "backward: nothing requested");
  return net->n.backward(params, state, is_f16, mean_inv, B, ws, d_out, grads, d_action, ST(stream));
  API_END
}

int cpp_conv_forward(const void* x, int32_t x_is_f16, const float* mean_inv, const float* w, const float* bias, int32_t B,
                     int32_t H, int32_t W, int32_t Cin, int32_t KS, float* pooled, uint8_t* amax, void* stream) {
  API_BEGIN
  NEED(x); NEED(w); NEED(bias); NEED(pooled); NEED(amax);
  ConvLayer L; L.H = H; L.W = W; L.Cin = Cin; L.KS = KS;
  return launch_conv_fwd(L, x, x_is_f16, mean_inv, w, bias, B, pooled, amax, ST(stream));
  API_END
}
int cpp_conv_dgrad(const float* d_pooled, const uint8_t* amax, const float* w, int32_t B, int32_t H, int32_t W, int32_t KS,
                   float* dx, void* stream) {
  API_BEGIN
  NEED(d_pooled); NEED(amax); NEED(w); NEED(dx);
  ConvLayer L; L.H = H; L.W = W; L.Cin = kConvCout; L.KS = KS;
  return launch_conv_dgrad(L, d_pooled, amax, w, B, dx, ST(stream));
  API_END
}
int64_t cpp_conv_wgrad_scratch_floats(int32_t H, int32_t W, int32_t Cin, int32_t KS) {
  ConvLayer L; L.H = H; L.W = W; L.Cin = Cin; L.KS = KS;
  return conv_wgrad_scratch_floats(L);
}
int cpp_conv_wgrad(const void* x, int32_t x_is_f16, const float* mean_inv, const float* d_pooled, const uint8_t* amax, int32_t B,
                   int32_t H, int32_t W, int32_t Cin, int32_t KS, float* dw, float* db, float* scratch, void* stream) {
  API_BEGIN
  NEED(x); NEED(d_pooled); NEED(amax); NEED(dw); NEED(db); NEED(scratch);
  CPP_REQUIRE(!x_is_f16 || mean_inv != nullptr, "fp16 input needs whitening stats");
  ConvLayer L; L.H = H; L.W = W; L.Cin = Cin; L.KS = KS;
  return launch_conv_wgrad(L, x, x_is_f16, mean_inv, d_pooled, amax, B, dw, db, scratch, ST(stream));
  API_END
}

int64_t cpp_conv_tc_scratch_bytes(int32_t nets, int32_t H, int32_t W, int32_t Cin, int32_t KS) {
  return tc::conv_tc_scratch_bytes(nets, H, W, Cin, KS);
}
int cpp_conv_forward_tc(const void* x_f16, const int32_t* rows, const float* mean_inv, int32_t nets, const float* const* w,
                        const float* const* bias, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t KS,
                        float* const* pooled, uint8_t* const* amax, void* scratch, void* stream, int32_t x_is_pieces,
                        void* const* pooled_hl) {
  API_BEGIN
  NEED(x_f16); NEED(w); NEED(bias); NEED(pooled); NEED(amax); NEED(scratch);
  return tc::launch_conv_fwd_tc(x_f16, rows, mean_inv, nets, w, bias, B, H, W, Cin, KS, pooled, amax, scratch, ST(stream),
                                x_is_pieces, reinterpret_cast<__half* const*>(pooled_hl));
  API_END
}

int64_t cpp_conv_dgrad_tc_scratch_bytes(int32_t B, int32_t H, int32_t W, int32_t KS) {
  const int64_t pk = tc::conv_tc_scratch_bytes(1, H, W, tc::kC24, KS);
  if (pk < 0 || B < 1) return -1;
  return 256 + round_up((int64_t)B * H * W * tc::kC24 * 2, 256) + pk;
}
int cpp_conv_dgrad_tc(const float* d_pooled, const uint8_t* amax, const float* w, int32_t B, int32_t H, int32_t W, int32_t KS,
                      float* dx, void* scratch, void* stream) {
  API_BEGIN
  NEED(d_pooled); NEED(amax); NEED(w); NEED(dx); NEED(scratch);
  CPP_REQUIRE(((uintptr_t)scratch & 255) == 0, "dgrad_tc: unaligned scratch");
  char* sc = reinterpret_cast<char*>(scratch);
  float* gsc = reinterpret_cast<float*>(sc);
  __half* dyp = reinterpret_cast<__half*>(sc + 256);
  void* pk = sc + 256 + round_up((int64_t)B * H * W * tc::kC24 * 2, 256);
  if (tc::conv_dgrad_fused_supported(H, W, KS))
    return tc::launch_conv_dgrad_tc_fused(d_pooled, amax, gsc, gsc + 1, 0, w, B, H, W, KS, dx, pk, ST(stream));
  CPP_TRY(tc::launch_unpool_split(d_pooled, amax, B, H, W, gsc, gsc + 1, dyp, ST(stream)));
  return tc::launch_conv_dgrad_tc(dyp, gsc + 1, w, B, H, W, KS, dx, pk, ST(stream));
  API_END
}
int64_t cpp_conv_wgrad_mma_scratch_bytes(int32_t nets, int32_t H, int32_t W, int32_t Cin, int32_t KS) {
  // the same buffer serves every x_is_pieces value the caller may pass for this channel count
  int64_t b = wg::conv_wgrad_mma_scratch_bytes(nets, H, W, Cin, KS);
  if (Cin % 2 == 0) b = std::max(b, wg::conv_wgrad_mma_scratch_bytes(nets, H, W, Cin, KS, 1));
  if (Cin == tc::kC24) b = std::max(b, wg::conv_wgrad_mma_scratch_bytes(nets, H, W, Cin, KS, 2));
  return b;
}
int cpp_conv_wgrad_mma(const void* x_f16, const float* mean_inv, int32_t x_is_pieces, int32_t nets, const float* const* d_pooled,
                       const uint8_t* const* amax, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t KS, float* const* dw,
                       float* const* db, void* scratch, void* stream) {
  API_BEGIN
  NEED(x_f16); NEED(d_pooled); NEED(amax); NEED(dw); NEED(db); NEED(scratch);
  return wg::launch_conv_wgrad_mma(x_f16, mean_inv, x_is_pieces, nets, d_pooled, amax, B, H, W, Cin, KS, dw, db, scratch, ST(stream));
  API_END
}

// ---------------------------------------------------------------- clip / optimiser / target copy
int64_t cpp_norm_scratch_doubles(void) { return norm_scratch_doubles(); }
int cpp_global_norm_scale(const float* grads, int64_t n, float clip, double* scratch, float* out2, void* stream) {
  API_BEGIN
  NEED(grads); NEED(scratch); NEED(out2);
  return launch_global_norm_scale(grads, n, clip, scratch, out2, ST(stream));
  API_END
}
int cpp_optimiser_apply(int32_t kind, float* params, const float* grads, const float* scale, int64_t n, float lr, float momentum,
                        float beta1, float beta2, float eps, float* slots, float* opt_state, void* stream) {
  API_BEGIN
  NEED(params); NEED(grads);
  return launch_optimiser(kind, params, grads, scale, n, lr, momentum, beta1, beta2, eps, slots, opt_state, nullptr, ST(stream));
  API_END
}
int cpp_soft_update(float* target, const float* source, float coeff, int64_t n, void* stream) {
  API_BEGIN
  NEED(target); NEED(source);
  return launch_soft_update(target, source, coeff, n, ST(stream));
  API_END
}

// ---------------------------------------------------------------- data parallel (8e)
int cpp_nccl_unique_id(void* out128) { API_BEGIN NEED(out128); return comm_unique_id(out128); API_END }
int cpp_nccl_version(int32_t* out) { API_BEGIN NEED(out); int v = 0; CPP_TRY(comm_version(&v)); *out = v; return CPP_OK; API_END }
int cpp_ddpg_comm_init(cpp_ddpg* a, int32_t rank, int32_t world, const void* id128) {
  API_BEGIN
  NEED(a);
  for (auto& gc : a->a.graph) gc.clear();
  return a->a.comm.init(rank, world, id128);
  API_END
}
int cpp_naf_comm_init(cpp_naf* a, int32_t rank, int32_t world, const void* id128) {
  API_BEGIN
  NEED(a);
  a->a.graph.clear();
  return a->a.comm.init(rank, world, id128);
  API_END
}
int cpp_ddpg_p2p_prepare(cpp_ddpg* a, int32_t rank, int32_t world, void* handle_out64) {
  API_BEGIN
  NEED(a); NEED(handle_out64);
  for (auto& gc : a->a.graph) gc.clear();
  return a->a.comm.p2p_prepare(rank, world, a->a.total, handle_out64);
  API_END
}
int cpp_ddpg_all_reduce_grads(cpp_ddpg* a, void* stream) {
  API_BEGIN
  NEED(a);
  CPP_REQUIRE(a->a.bound, "agent buffers not bound");
  return a->a.comm.all_reduce(a->a.buf.grads, a->a.total, ST(stream));
  API_END
}
int cpp_naf_all_reduce_grads(cpp_naf* a, void* stream) {
  API_BEGIN
  NEED(a);
  CPP_REQUIRE(a->a.bound, "agent buffers not bound");
  return a->a.comm.all_reduce(a->a.buf.grads, a->a.total, ST(stream));
  API_END
}
int cpp_ddpg_p2p_connect(cpp_ddpg* a, const void* handles) { API_BEGIN NEED(a); NEED(handles); return a->a.comm.p2p_connect(handles); API_END }
int cpp_naf_p2p_prepare(cpp_naf* a, int32_t rank, int32_t world, void* handle_out64) {
  API_BEGIN
  NEED(a); NEED(handle_out64);
  a->a.graph.clear();
  return a->a.comm.p2p_prepare(rank, world, a->a.total, handle_out64);
  API_END
}
int cpp_naf_p2p_connect(cpp_naf* a, const void* handles) { API_BEGIN NEED(a); NEED(handles); return a->a.comm.p2p_connect(handles); API_END }

// ---------------------------------------------------------------- DDPG
int cpp_ddpg_create(const cpp_ddpg_config* cfg, cpp_ddpg** out) {
  API_BEGIN
  NEED(cfg); NEED(out);
  cpp_ddpg* a = new (std::nothrow) cpp_ddpg();
  NEED(a);
  const int st = a->a.init(*cfg);
  if (st != CPP_OK) { delete a; return st; }
  a->a.carve(nullptr, false);
  *out = a;
  return CPP_OK;
  API_END
}
int cpp_ddpg_destroy(cpp_ddpg* a) { delete a; return CPP_OK; }
int64_t cpp_ddpg_workspace_bytes(const cpp_ddpg* a) { return a ? (int64_t)a->a.ws_bytes : -1; }
int cpp_ddpg_layout(const cpp_ddpg* a, int64_t* out5) {
  API_BEGIN
  NEED(a); NEED(out5);
  out5[0] = a->a.n_a; out5[1] = a->a.n_c; out5[2] = a->a.off_c; out5[3] = a->a.off_loss; out5[4] = a->a.total;
  return CPP_OK;
  API_END
}
int cpp_ddpg_bind(cpp_ddpg* a, const cpp_ddpg_buffers* b) { API_BEGIN NEED(a); NEED(b); return a->a.bind(*b); API_END }
int cpp_ddpg_set_moments(cpp_ddpg* a, const float* m1, const float* m2) {
  API_BEGIN NEED(a); a->a.pinned1 = m1; a->a.pinned2 = m2; a->a.critic_trunk_valid = false; return CPP_OK; API_END
}
int cpp_ddpg_actor_backward(cpp_ddpg* a, const void* s1, int32_t is_f16, int32_t B, int32_t B_global, void* stream) {
  API_BEGIN TrainingMode tm_(1); NEED(a); NEED(s1); return a->a.actor_backward(s1, is_f16, B, B_global, ST(stream)); API_END
}
int cpp_ddpg_actor_apply(cpp_ddpg* a, void* stream) { API_BEGIN NEED(a); return a->a.actor_apply(ST(stream)); API_END }
int cpp_ddpg_actor_train(cpp_ddpg* a, const void* s1, int32_t is_f16, int32_t B, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1);
  CPP_TRY(a->a.actor_backward(s1, is_f16, B, B, ST(stream)));
  return a->a.actor_apply(ST(stream));
  API_END
}
int cpp_ddpg_critic_backward(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask,
                             const void* s2, int32_t is_f16, int32_t B, int32_t B_global, int32_t reuse, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2);
  CPP_REQUIRE(B_global >= B, "B_global %d < B %d", B_global, B);
  return a->a.critic_backward(s1, action, reward, mask, s2, is_f16, B, B_global, reuse, ST(stream));
  API_END
}
int cpp_ddpg_critic_apply(cpp_ddpg* a, void* stream) { API_BEGIN NEED(a); return a->a.critic_apply(ST(stream)); API_END }
int cpp_ddpg_critic_train(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                          int32_t is_f16, int32_t B, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2);
  CPP_TRY(a->a.critic_backward(s1, action, reward, mask, s2, is_f16, B, B, 0, ST(stream)));
  return a->a.critic_apply(ST(stream));
  API_END
}
int cpp_ddpg_step_backward(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask,
                           const void* s2, int32_t is_f16, int32_t B, int32_t B_global, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2);
  CPP_REQUIRE(B_global >= B, "B_global %d < B %d", B_global, B);
  return a->a.step_backward(s1, action, reward, mask, s2, is_f16, B, B_global, ST(stream));
  API_END
}
int cpp_ddpg_train_step(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask,
                        const void* s2, int32_t is_f16, int32_t B, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2);
  return a->a.step(s1, action, reward, mask, s2, is_f16, B, B * a->a.comm.world, true, ST(stream));
  API_END
}
int cpp_ddpg_step_apply(cpp_ddpg* a, void* stream) {
  API_BEGIN
  NEED(a);
  return a->a.apply_both(ST(stream));
  API_END
}
int cpp_ddpg_check_loss(cpp_ddpg* a, const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                        int32_t is_f16, int32_t B, float* loss, float* td, float* q, void* stream) {
  API_BEGIN TrainingMode tm_(0);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2); NEED(loss); NEED(td); NEED(q);
  return a->a.check_loss(s1, action, reward, mask, s2, is_f16, B, loss, td, q, ST(stream));
  API_END
}
int cpp_ddpg_action_given(cpp_ddpg* a, const void* state, int32_t is_f16, int32_t B, float* out, void* stream) {
  API_BEGIN TrainingMode tm_(0); NEED(a); NEED(state); NEED(out); return a->a.action_given(state, is_f16, B, out, ST(stream)); API_END
}
int cpp_ddpg_action_given_fast(cpp_ddpg* a, const void* state, int32_t is_f16, int32_t B, float* out, void* stream) {
  API_BEGIN TrainingMode tm_(0); NEED(a); NEED(state); NEED(out); return a->a.action_given_fast(state, is_f16, B, out, ST(stream)); API_END
}
int cpp_ddpg_update_targets(cpp_ddpg* a, float coeff, void* stream) { API_BEGIN NEED(a); return a->a.update_targets(coeff, ST(stream)); API_END }

static int debug_view(const Net& net, const char* ws_part, const void* ws_base, int kind, int index, int B, int64_t* out4) {
  CPP_REQUIRE(ws_part != nullptr && ws_base != nullptr, "debug_view: buffers not bound");
  CPP_REQUIRE(B >= 1, "debug_view: batch %d", B);
  const Net::Layout L = net.layout(B);
  size_t off = 0; int64_t per = 0, valid = 0;
  if (kind == 0 || kind == 1) {
    CPP_REQUIRE(net.pixels && index >= 0 && index < 3, "debug_view: conv layer %d of a %s network", index, net.pixels ? "pixel" : "low-dim");
    off = kind == 0 ? L.amax[index] : L.pooled[index];
    per = valid = (int64_t)net.conv[index].PH() * net.conv[index].PW() * kConvCout;
  } else if (kind == 2) {
    CPP_REQUIRE(index >= 0 && index < net.n_fc, "debug_view: FC layer %d", index);
    off = L.h[index]; per = net.out_ld[index]; valid = net.out_dim[index];
  } else if (kind == 3) {
    CPP_REQUIRE(index >= 0 && index < net.n_fc && net.drop[index], "debug_view: FC layer %d has no dropout", index);
    off = L.mask[index]; per = valid = net.out_dim[index];
  } else {
    set_error("debug_view: kind %d", kind);
    return CPP_ERR_INVALID;
  }
  out4[0] = (int64_t)((ws_part - reinterpret_cast<const char*>(ws_base)) + (ptrdiff_t)off);
  out4[1] = B; out4[2] = per; out4[3] = valid;
  return CPP_OK;
}
int cpp_ddpg_debug_view(const cpp_ddpg* a, int32_t part, int32_t kind, int32_t index, int32_t B, int64_t* out4) {
  API_BEGIN
  NEED(a); NEED(out4);
  CPP_REQUIRE(part >= 0 && part < 4, "debug_view: part %d", part);
  const DDPG& d = a->a;
  const char* ws[4] = {d.ws_actor, d.ws_critic, d.ws_target, d.ws_target2};
  return debug_view((part & 1) ? d.critic : d.actor, ws[part], d.buf.workspace, kind, index, B, out4);
  API_END
}

// ---------------------------------------------------------------- NAF
int cpp_naf_debug_view(const cpp_naf* a, int32_t part, int32_t kind, int32_t index, int32_t B, int64_t* out4) {
  API_BEGIN
  NEED(a); NEED(out4);
  CPP_REQUIRE(part >= 0 && part < 4, "debug_view: part %d", part);
  const NAF& d = a->a;
  const char* ws[4] = {d.ws_v, d.ws_m, d.ws_l, d.ws_t};
  const Net* nets[4] = {&d.value, &d.mu, &d.l, &d.value};
  return debug_view(*nets[part], ws[part], d.buf.workspace, kind, index, B, out4);
  API_END
}
int cpp_naf_create(const cpp_naf_config* cfg, cpp_naf** out) {
  API_BEGIN
  NEED(cfg); NEED(out);
  cpp_naf* a = new (std::nothrow) cpp_naf();
  NEED(a);
  const int st = a->a.init(*cfg);
  if (st != CPP_OK) { delete a; return st; }
  a->a.carve(nullptr, false);
  *out = a;
  return CPP_OK;
  API_END
}
int cpp_naf_destroy(cpp_naf* a) { delete a; return CPP_OK; }
int64_t cpp_naf_workspace_bytes(const cpp_naf* a) { return a ? (int64_t)a->a.ws_bytes : -1; }
int cpp_naf_layout(const cpp_naf* a, int64_t* o) {
  API_BEGIN
  NEED(a); NEED(o);
  o[0] = a->a.n_v; o[1] = a->a.n_m; o[2] = a->a.n_l; o[3] = a->a.off_m; o[4] = a->a.off_l; o[5] = a->a.off_loss; o[6] = a->a.total;
  return CPP_OK;
  API_END
}
int cpp_naf_bind(cpp_naf* a, const cpp_naf_buffers* b) { API_BEGIN NEED(a); NEED(b); return a->a.bind(*b); API_END }
int cpp_naf_set_moments(cpp_naf* a, const float* m1, const float* m2) { API_BEGIN NEED(a); a->a.pinned1 = m1; a->a.pinned2 = m2; return CPP_OK; API_END }
int cpp_naf_backward(cpp_naf* a, const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                     int32_t is_f16, int32_t B, int32_t B_global, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2);
  CPP_REQUIRE(B_global >= B, "B_global %d < B %d", B_global, B);
  return a->a.backward(s1, action, reward, mask, s2, is_f16, B, B_global, ST(stream));
  API_END
}
int cpp_naf_apply(cpp_naf* a, int32_t check, float* loss_host, void* stream) { API_BEGIN NEED(a); return a->a.apply(check, loss_host, ST(stream)); API_END }
int cpp_naf_train(cpp_naf* a, const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                  int32_t is_f16, int32_t B, float* loss_host, void* stream) {
  API_BEGIN TrainingMode tm_(1);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2);
  CPP_TRY(a->a.backward(s1, action, reward, mask, s2, is_f16, B, B, ST(stream)));
  return a->a.apply(1, loss_host, ST(stream));
  API_END
}
int cpp_naf_debug_values(cpp_naf* a, const void* s1, const float* action, const float* reward, const float* mask, const void* s2,
                         int32_t is_f16, int32_t B, float* l_values, float* loss, float* V, float* Aout, float* V2, void* stream) {
  API_BEGIN TrainingMode tm_(0);
  NEED(a); NEED(s1); NEED(action); NEED(reward); NEED(mask); NEED(s2); NEED(l_values); NEED(loss); NEED(V); NEED(Aout); NEED(V2);
  return a->a.debug_values(s1, action, reward, mask, s2, is_f16, B, l_values, loss, V, Aout, V2, ST(stream));
  API_END
}
int cpp_naf_action_given(cpp_naf* a, const void* state, int32_t is_f16, int32_t B, float* out, void* stream) {
  API_BEGIN TrainingMode tm_(0); NEED(a); NEED(state); NEED(out); return a->a.action_given(state, is_f16, B, out, ST(stream)); API_END
}
int cpp_naf_action_given_fast(cpp_naf* a, const void* state, int32_t is_f16, int32_t B, float* out, void* stream) {
  API_BEGIN TrainingMode tm_(0); NEED(a); NEED(state); NEED(out); return a->a.action_given_fast(state, is_f16, B, out, ST(stream)); API_END
}
int cpp_naf_value_given(cpp_naf* a, const void* state, int32_t is_f16, int32_t B, float* out, void* stream) {
  API_BEGIN TrainingMode tm_(0); NEED(a); NEED(state); NEED(out); return a->a.value_given(state, is_f16, B, out, ST(stream)); API_END
}
int cpp_naf_update_targets(cpp_naf* a, float coeff, void* stream) { API_BEGIN NEED(a); return a->a.update_targets(coeff, ST(stream)); API_END }

// ---------------------------------------------------------------- LRPG
int cpp_lrpg_create(const cpp_lrpg_config* cfg, cpp_lrpg** out) {
  API_BEGIN
  NEED(cfg); NEED(out);
  cpp_lrpg* a = new (std::nothrow) cpp_lrpg();
  NEED(a);
  const int st = a->a.init(*cfg);
  if (st != CPP_OK) { delete a; return st; }
  a->a.carve(nullptr, false);
  *out = a;
  return CPP_OK;
  API_END
}
int cpp_lrpg_destroy(cpp_lrpg* a) { delete a; return CPP_OK; }
int64_t cpp_lrpg_workspace_bytes(const cpp_lrpg* a) { return a ? (int64_t)a->a.ws_bytes : -1; }
int64_t cpp_lrpg_num_params(const cpp_lrpg* a) { return a ? a->a.n : -1; }
int cpp_lrpg_bind(cpp_lrpg* a, const cpp_lrpg_buffers* b) { API_BEGIN NEED(a); NEED(b); return a->a.bind(*b); API_END }
int cpp_lrpg_train(cpp_lrpg* a, const float* obs, const int32_t* actions, const float* adv, int32_t N, float* loss_host, void* stream) {
  API_BEGIN NEED(a); NEED(obs); NEED(actions); NEED(adv); return a->a.train(obs, actions, adv, N, loss_host, ST(stream)); API_END
}
int cpp_lrpg_logits(cpp_lrpg* a, const float* obs, int32_t N, float* logits, void* stream) {
  API_BEGIN NEED(a); NEED(obs); NEED(logits); return a->a.get_logits(obs, N, logits, ST(stream)); API_END
}

}  // extern "C"
