// Net: parameter layout, activation workspace, forward and backward of one reference network.
#include <stdlib.h>
#include <algorithm>
#include "net.cuh"
#include "conv_tc.cuh"
#include "conv_wgrad_mma.cuh"

namespace cpp {

int Net::init(const cpp_net_spec& s) {
  spec = s;
  pixels = s.pixels != 0;
  bn = pixels && s.batch_norm != 0;
  n_fc = s.n_fc;
  concat_at = s.concat_at;
  action_dim = s.action_dim;
  CPP_REQUIRE(n_fc >= 1 && n_fc <= CPP_MAX_FC, "n_fc=%d out of range", n_fc);
  CPP_REQUIRE(concat_at < n_fc, "concat_at=%d out of range", concat_at);
  CPP_REQUIRE(concat_at < 0 || action_dim >= 1, "action concat needs action_dim >= 1");
  vars.clear();
  any_drop = false;
  int64_t off = 0;
  auto add_var = [&](int nd, int64_t a, int64_t b, int64_t c, int64_t d) {
    VarInfo v; v.offset = off; v.ndim = nd; v.shape[0] = a; v.shape[1] = b; v.shape[2] = c; v.shape[3] = d;
    int64_t n = 1; for (int i = 0; i < nd; ++i) n *= v.shape[i];
    vars.push_back(v); off += n;
  };
  if (pixels) {
    CPP_REQUIRE(s.H >= 8 && s.W >= 8 && s.Cin >= 1, "pixel state %dx%dx%d too small (three 2x2 pools)", s.H, s.W, s.Cin);
    CPP_REQUIRE(!(concat_at == 0), "action concat in front of the first FC layer is a low-dim-only layout");
    int h = s.H, w = s.W, cin = s.Cin;
    const int ks[3] = {5, 5, 3};                       // base_network.py:103,111,119
    for (int i = 0; i < 3; ++i) {
      conv[i].H = h; conv[i].W = w; conv[i].Cin = cin; conv[i].KS = ks[i];
      off_conv_w[i] = off; add_var(4, ks[i], ks[i], cin, kConvCout);
      off_conv_b[i] = off; add_var(1, kConvCout, 1, 1, 1);                 // biases, or BatchNorm/beta (no bias with a normalizer_fn)
      if (bn) {                                                            // slim.batch_norm(center=True, scale=False): TF creation order
        off_bn_mean[i] = off; add_var(1, kConvCout, 1, 1, 1);              // BatchNorm/moving_mean
        off_bn_var[i] = off; add_var(1, kConvCout, 1, 1, 1);               // BatchNorm/moving_variance
      }
      h /= 2; w /= 2; cin = kConvCout;
    }
    feat = h * w * kConvCout;
  } else {
    CPP_REQUIRE(s.input_dim >= 1, "input_dim=%d", s.input_dim);
    feat = s.input_dim;
  }
  int d = feat;
  for (int i = 0; i < n_fc; ++i) {
    CPP_REQUIRE(s.fc_out[i] >= 1, "fc_out[%d]=%d", i, s.fc_out[i]);
    CPP_REQUIRE(s.fc_act[i] >= 0 && s.fc_act[i] <= 2, "fc_act[%d]=%d", i, s.fc_act[i]);
    CPP_REQUIRE(i == n_fc - 1 || s.fc_act[i] == 1, "hidden FC layers are ReLU in the reference (layer %d)", i);
    in_dim[i] = d + (concat_at == i ? action_dim : 0);
    out_dim[i] = s.fc_out[i];
    act[i] = s.fc_act[i];
    drop[i] = s.fc_dropout[i] != 0;
    CPP_REQUIRE(!drop[i] || (i < n_fc - 1 && act[i] == 1), "dropout follows hidden ReLU layers only (layer %d)", i);
    any_drop = any_drop || drop[i];
    off_fc_w[i] = off; add_var(2, in_dim[i], out_dim[i], 1, 1);
    off_fc_b[i] = off; add_var(1, out_dim[i], 1, 1, 1);
    d = out_dim[i];
  }
  for (int i = 0; i < n_fc; ++i) out_ld[i] = (i + 1 < n_fc) ? in_dim[i + 1] : out_dim[i];
  nparams = off;
  return CPP_OK;
}

Net::Layout Net::layout(int B) const {
  Layout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (size_t)round_up((int64_t)bytes, 256); return o; };
  if (B < 1) B = 1;
  if (pixels) {
    for (int i = 0; i < 3; ++i) {
      const size_t n = (size_t)B * conv[i].PH() * conv[i].PW() * kConvCout;
      L.pooled[i] = take(n * sizeof(float));
      L.amax[i] = take(n);
      if (i < 2) L.hl[i] = take(n / kConvCout * tc::kC24 * sizeof(__half));   // fp16 pieces of pooled[i], 24-channel layout (conv_tc.cuh)
    }
    // gradients wrt pooled2 / pooled1 (dense, the size of conv3 / conv2 inputs)
    L.dpool[0] = take((size_t)B * conv[2].H * conv[2].W * kConvCout * sizeof(float));
    L.dpool[1] = take((size_t)B * conv[1].H * conv[1].W * kConvCout * sizeof(float));
    int64_t wg = 0;
    for (int i = 0; i < 3; ++i) { const int64_t w = conv_wgrad_scratch_floats(conv[i]); if (w > wg) wg = w; }
    L.wgrad = take((size_t)wg * sizeof(float));
    L.dyp = take((size_t)B * conv[1].H * conv[1].W * tc::kC24 * sizeof(__half));   // un-pooled gradient pieces (dgrad on tensor cores)
    L.gsc = take(8 * sizeof(float));                                                      // {max|g|, 1/scale} per conv layer
    if (bn) {
      for (int i = 0; i < 3; ++i) {
        L.raw[i] = take((size_t)B * conv[i].H * conv[i].W * kConvCout * sizeof(float));
        L.bnscr[i] = take((size_t)bn_scratch_bytes());
      }
      L.dconv = take((size_t)B * conv[0].H * conv[0].W * kConvCout * sizeof(float));
      L.bnjunk = take(64);
    }
  } else {
    L.x0 = take((size_t)B * in_dim[0] * sizeof(float));
  }
  for (int i = 0; i < n_fc; ++i) L.h[i] = take((size_t)B * out_ld[i] * sizeof(float));
  // gradients: one buffer per FC layer input (a weight gradient running on a side stream may still read the gradient of
  // layer i after the main chain has moved on) and the gradient wrt the last pre-activation
  for (int i = 0; i < n_fc; ++i) L.dX[i] = take((size_t)B * in_dim[i] * sizeof(float));
  L.dTop = take((size_t)B * out_dim[n_fc - 1] * sizeof(float));
  if (any_drop) {
    for (int i = 0; i < n_fc; ++i) if (drop[i]) L.mask[i] = take((size_t)B * out_dim[i]);
    L.dropctr = take(sizeof(unsigned long long));
  }
  L.total = off;
  return L;
}

const float* Net::fc_input(const Layout& L, char* ws, int i, int* ld) const {
  if (i == 0) {
    if (pixels) { *ld = feat; return reinterpret_cast<const float*>(ws + L.pooled[2]); }
    *ld = in_dim[0];
    return reinterpret_cast<const float*>(ws + L.x0);
  }
  *ld = out_ld[i - 1];
  return reinterpret_cast<const float*>(ws + L.h[i - 1]);
}

bool Net::tc_route(int is_f16) const {
  return pixels && !bn && is_f16 && conv1_tc_enabled() && tc::conv_tc_supported(1, conv[0].H, conv[0].W, conv[0].Cin, conv[0].KS) &&
         tc::conv_tc_supported(1, conv[1].H, conv[1].W, tc::kC24, conv[1].KS) &&
         tc::conv_tc_supported(1, conv[2].H, conv[2].W, tc::kC24, conv[2].KS);
}

int Net::forward_trunk(const float* params, const void* state, int is_f16, const float* mean_inv, int B, void* ws_,
                       cudaStream_t s, int first_conv, void* tc_scratch) const {
  CPP_REQUIRE(B >= 1, "batch %d", B);
  char* ws = reinterpret_cast<char*>(ws_);
  const Layout L = layout(B);
  if (pixels) {
    CPP_REQUIRE(first_conv > 0 || mean_inv != nullptr, "pixel network needs whitening statistics");
    const void* x = state; int xf16 = is_f16; const float* mi = mean_inv;
    if (first_conv > 0) { x = ws + L.pooled[first_conv - 1]; xf16 = 0; mi = nullptr; }
    for (int i = first_conv; i < 3; ++i) {
      float* pooled = reinterpret_cast<float*>(ws + L.pooled[i]);
      uint8_t* amax = reinterpret_cast<uint8_t*>(ws + L.amax[i]);
      if (i > 0 && first_conv > 0 && tc_scratch != nullptr) {
        // tensor cores: the layer below left its output as fp16 pieces; this layer leaves pieces for the next one
        const float* w[1] = {params + off_conv_w[i]}; const float* b[1] = {params + off_conv_b[i]};
        float* po[1] = {pooled}; uint8_t* am[1] = {amax};
        __half* hl[1] = {i < 2 ? reinterpret_cast<__half*>(ws + L.hl[i]) : nullptr};
        CPP_TRY(tc::launch_conv_fwd_tc(ws + L.hl[i - 1], nullptr, nullptr, 1, w, b, B, conv[i].H, conv[i].W, tc::kC24, conv[i].KS,
                                       po, am, reinterpret_cast<char*>(tc_scratch) + i * trunk_slot_bytes(*this), s, 2, hl,
                                       g_tc_prepped ? tc::kPhaseMain : tc::kPhaseBoth));
      } else if (bn) {
        // raw conv (no bias) -> batch statistics (IS_TRAINING) or moving statistics -> normalise + beta -> ReLU -> pool
        float* raw = reinterpret_cast<float*>(ws + L.raw[i]);
        CPP_TRY(launch_conv_raw(conv[i], x, xf16, mi, params + off_conv_w[i], B, raw, s));
        CPP_TRY(launch_bn_forward(raw, params + off_conv_b[i], params + off_bn_mean[i], params + off_bn_var[i], g_is_training, B,
                                  conv[i].H, conv[i].W, ws + L.bnscr[i], pooled, amax, s));
      } else {
        CPP_TRY(launch_conv_fwd(conv[i], x, xf16, mi, params + off_conv_w[i], params + off_conv_b[i], B, pooled, amax, s));
      }
      x = pooled; xf16 = 0; mi = nullptr;
      trace_mark(i == 1 ? "   . conv2 fwd" : (i == 2 ? "   . conv3 fwd" : "   . conv1 fwd"), s);
    }
  } else {
    CPP_TRY(launch_state_to_f32(state, is_f16, B, feat, reinterpret_cast<float*>(ws + L.x0), in_dim[0], s));
  }
  return CPP_OK;
}

int Net::forward_fc(const float* params, const float* action, int B, void* ws_, float* out, cudaStream_t s, int first_fc,
                    int end_fc) const {
  CPP_REQUIRE(B >= 1, "batch %d", B);
  if (end_fc < 0 || end_fc > n_fc) end_fc = n_fc;
  CPP_REQUIRE(concat_at < 0 || concat_at < first_fc || concat_at >= end_fc || action != nullptr, "this network needs an action input");
  const bool dropping = any_drop && g_is_training;      // slim.dropout is the identity unless IS_TRAINING (base_network.py:69-70)
  if (!dropping && fused_mlp_enabled() && mlp_fits(*this)) return launch_mlp_forward(*this, params, action, B, ws_, out, s, first_fc, end_fc);
  char* ws = reinterpret_cast<char*>(ws_);
  const Layout L = layout(B);
  if (dropping) CPP_TRY(launch_dropout_tick(reinterpret_cast<unsigned long long*>(ws + L.dropctr), s));
  for (int i = first_fc; i < end_fc; ++i) {
    int ld;
    const float* x = fc_input(L, ws, i, &ld);
    if (concat_at == i)   // tf.concat(1, [hidden, action]), ddpg_cartpole.py:170,175
      CPP_TRY(launch_copy_cols(action, action_dim, B, action_dim, const_cast<float*>(x), ld, in_dim[i] - action_dim, s));
    GemmArgs g{};
    g.A = x; g.lda = ld; g.transA = 0;
    g.B = params + off_fc_w[i]; g.ldb = out_dim[i]; g.transB = 0;
    g.C = reinterpret_cast<float*>(ws + L.h[i]); g.ldc = out_ld[i];
    g.M = B; g.N = out_dim[i]; g.K = in_dim[i];
    g.epi = EPI_BIAS_ACT; g.bias = params + off_fc_b[i]; g.act = act[i];
    CPP_TRY(launch_gemm(g, s));
    if (dropping && drop[i])
      CPP_TRY(launch_dropout(reinterpret_cast<float*>(ws + L.h[i]), out_ld[i], B, out_dim[i], reinterpret_cast<uint8_t*>(ws + L.mask[i]),
                             reinterpret_cast<const unsigned long long*>(ws + L.dropctr), i, s));
  }
  if (out != nullptr && end_fc == n_fc) {
    const int n = out_dim[n_fc - 1];
    CPP_CHECK_CUDA(cudaMemcpyAsync(out, ws + L.h[n_fc - 1], (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  return CPP_OK;
}

int Net::forward(const float* params, const void* state, int is_f16, const float* mean_inv, const float* action,
                 int B, void* ws, float* out, cudaStream_t s, int first_fc) const {
  if (first_fc == 0) CPP_TRY(forward_trunk(params, state, is_f16, mean_inv, B, ws, s, 0));
  return forward_fc(params, action, B, ws, out, s, first_fc);
}

static int g_conv1_tc_override = -1;
void set_conv1_tc_enabled(int on) { g_conv1_tc_override = on; }
bool conv1_tc_enabled() {
  static const bool env_on = [] { const char* e = getenv("CARTPOLEPP_CONV1"); return !(e && std::string(e) == "ffma"); }();
  return g_conv1_tc_override < 0 ? env_on : g_conv1_tc_override != 0;
}

int64_t conv1_wgrad_group_scratch_bytes(int n, const Net& net) {
  if (!net.pixels) return 0;
  int64_t b = wg::conv_wgrad_mma_scratch_bytes(n, net.conv[0].H, net.conv[0].W, net.conv[0].Cin, net.conv[0].KS);
  for (int i = 1; i < 3; ++i) b = std::max(b, wg::conv_wgrad_mma_scratch_bytes(1, net.conv[i].H, net.conv[i].W, tc::kC24, net.conv[i].KS, 2));
  return b > 0 ? b : 0;
}

int conv1_wgrad_group(int n, const Net* const* nets, char* const* ws, float* const* grads, const void* state, int is_f16,
                      const float* mean_inv, int B, void* scratch, cudaStream_t s, int gmax_from_dgrad) {
  const Net& n0 = *nets[0];
  if (!n0.pixels) return CPP_OK;
  CPP_REQUIRE(n >= 1 && n <= wg::kMaxNets, "conv1 wgrad group of %d networks", n);
  const ConvLayer& c1 = n0.conv[0];
  const float* gp[wg::kMaxNets]; const uint8_t* am[wg::kMaxNets]; float* dw[wg::kMaxNets]; float* db[wg::kMaxNets];
  const float* gm[wg::kMaxNets];
  for (int i = 0; i < n; ++i) {
    const Net::Layout L = nets[i]->layout(B);
    gm[i] = reinterpret_cast<const float*>(ws[i] + L.gsc) + 6;
    gp[i] = reinterpret_cast<const float*>(ws[i] + L.dpool[1]);
    am[i] = reinterpret_cast<const uint8_t*>(ws[i] + L.amax[0]);
    dw[i] = grads[i] + nets[i]->off_conv_w[0]; db[i] = grads[i] + nets[i]->off_conv_b[0];
  }
  if (scratch != nullptr && n0.tc_route(is_f16) && wg::conv_wgrad_mma_supported(n, c1.H, c1.W, c1.Cin, c1.KS))
    return wg::launch_conv_wgrad_mma(state, mean_inv, 0, n, gp, am, B, c1.H, c1.W, c1.Cin, c1.KS, dw, db, scratch, s,
                                     gmax_from_dgrad ? gm : nullptr);
  for (int i = 0; i < n; ++i) {
    const Net::Layout L = nets[i]->layout(B);
    if (nets[i]->bn)   // batch norm: Net::backward left the dense d(conv1) (and wrote d(beta) itself)
      CPP_TRY(launch_conv_wgrad(c1, state, is_f16, mean_inv, reinterpret_cast<const float*>(ws[i] + L.dconv), nullptr, B, dw[i],
                                reinterpret_cast<float*>(ws[i] + L.bnjunk), reinterpret_cast<float*>(ws[i] + L.wgrad), s));
    else
      CPP_TRY(launch_conv_wgrad(c1, state, is_f16, mean_inv, gp[i], am[i], B, dw[i], db[i], reinterpret_cast<float*>(ws[i] + L.wgrad), s));
  }
  return CPP_OK;
}

int64_t trunk_slot_bytes(const Net& net) {
  if (!net.pixels) return 0;
  int64_t b = 0;
  for (int n = 1; n <= tc::kMaxNets; ++n) b = std::max(b, tc::conv_tc_scratch_bytes(n, net.conv[0].H, net.conv[0].W, net.conv[0].Cin, net.conv[0].KS));
  for (int i = 1; i < 3; ++i) b = std::max(b, tc::conv_tc_scratch_bytes(1, net.conv[i].H, net.conv[i].W, tc::kC24, net.conv[i].KS));
  return b > 0 ? (int64_t)round_up(b, 256) : 0;
}

int64_t trunk_group_scratch_bytes(int, const Net& net) { return kTcSlots * trunk_slot_bytes(net); }

int Net::prep_trunk_tc(const float* params, int B, void* tc_scratch, bool with_dgrad, cudaStream_t s) const {
  char* base = reinterpret_cast<char*>(tc_scratch);
  const int64_t slot = trunk_slot_bytes(*this);
  for (int i = 1; i < 3; ++i) {
    const float* w[1] = {params + off_conv_w[i]}; const float* b[1] = {params + off_conv_b[i]};
    CPP_TRY(tc::launch_conv_fwd_tc(base, nullptr, nullptr, 1, w, b, B, conv[i].H, conv[i].W, tc::kC24, conv[i].KS, nullptr, nullptr,
                                   base + i * slot, s, 2, nullptr, tc::kPhasePrep));
    if (with_dgrad)
      CPP_TRY(tc::launch_conv_dgrad_tc(base, nullptr, params + off_conv_w[i], B, conv[i].H, conv[i].W, conv[i].KS, nullptr,
                                       base + (i == 2 ? 3 : 4) * slot, s, nullptr, tc::kPhasePrep));
  }
  return CPP_OK;
}

int conv1_forward_group(int n, const Net* const* nets, const float* const* params, char* const* ws, const void* state,
                        int is_f16, const float* mean_inv, int B, void* tc_scratch, cudaStream_t s, int* took_tc) {
  CPP_REQUIRE(n >= 1 && n <= tc::kMaxNets, "trunk group of %d networks", n);
  const Net& n0 = *nets[0];
  *took_tc = 0;
  const bool tc_ok = tc_scratch != nullptr && n0.tc_route(is_f16) &&
                     tc::conv_tc_supported(n, n0.conv[0].H, n0.conv[0].W, n0.conv[0].Cin, n0.conv[0].KS);
  if (!tc_ok) return CPP_OK;
  CPP_REQUIRE(mean_inv != nullptr, "pixel network needs whitening statistics");
  const float *w[tc::kMaxNets], *b[tc::kMaxNets];
  float* pooled[tc::kMaxNets]; uint8_t* amax[tc::kMaxNets]; __half* hl[tc::kMaxNets];
  for (int i = 0; i < n; ++i) {
    const Net& ni = *nets[i];
    CPP_REQUIRE(ni.pixels && ni.conv[0].H == n0.conv[0].H && ni.conv[0].W == n0.conv[0].W && ni.conv[0].Cin == n0.conv[0].Cin,
                "sibling networks must read the same state");
    const Net::Layout L = ni.layout(B);
    w[i] = params[i] + ni.off_conv_w[0]; b[i] = params[i] + ni.off_conv_b[0];
    pooled[i] = reinterpret_cast<float*>(ws[i] + L.pooled[0]); amax[i] = reinterpret_cast<uint8_t*>(ws[i] + L.amax[0]);
    hl[i] = reinterpret_cast<__half*>(ws[i] + L.hl[0]);
  }
  CPP_TRY(tc::launch_conv_fwd_tc(state, nullptr, mean_inv, n, w, b, B, n0.conv[0].H, n0.conv[0].W, n0.conv[0].Cin,
                                 n0.conv[0].KS, pooled, amax, tc_scratch, s, 0, hl));
  *took_tc = 1;
  return CPP_OK;
}

int trunk_forward_group(int n, const Net* const* nets, const float* const* params, char* const* ws, const void* state,
                        int is_f16, const float* mean_inv, int B, void* tc_scratch, cudaStream_t s) {
  int took_tc = 0;
  CPP_TRY(conv1_forward_group(n, nets, params, ws, state, is_f16, mean_inv, B, tc_scratch, s, &took_tc));
  for (int i = 0; i < n; ++i)
    CPP_TRY(nets[i]->forward_trunk(params[i], state, is_f16, mean_inv, B, ws[i], s, took_tc, took_tc ? tc_scratch : nullptr));
  return CPP_OK;
}

// dW = x^T . dPre and db = column sums of dPre of one FC layer: one tcgen05 GEMM with an all-ones row appended to x^T when the
// tensor-core route takes the layer (fc_tc.cu), else the FFMA GEMM + the column-sum kernel
static int fc_wgrad(const float* x, int xld, const float* dpre, int dld, int B, int in, int out, float* dw, float* db, cudaStream_t sw) {
  GemmArgs g{};
  g.A = x; g.lda = xld; g.transA = 1;
  g.B = dpre; g.ldb = dld; g.transB = 0;
  g.C = dw; g.ldc = out;
  g.M = in; g.N = out; g.K = B; g.epi = EPI_NONE;
  if (gemm_tc_wanted(g)) { g.colsum = db; return launch_gemm(g, sw); }
  CPP_TRY(launch_gemm(g, sw));
  return launch_colsum(dpre, dld, B, out, db, sw);
}

int Net::backward(const float* params, const void* state, int is_f16, const float* mean_inv, int B, void* ws_,
                  const float* d_out, float* grads, float* d_action, cudaStream_t s, int defer_conv1, void* wg_scratch,
                  void* tc_scratch, BackwardAux* aux, const float* d_rep_extra) const {
  CPP_REQUIRE(B >= 1, "batch %d", B);
  CPP_REQUIRE(d_action == nullptr || concat_at >= 0, "d_action requested from a network without action input");
  char* ws = reinterpret_cast<char*>(ws_);
  const Layout L = layout(B);
  const int last = n_fc - 1;
  const bool fork = aux != nullptr && aux->stream != nullptr && grads != nullptr && !bn;   // (batch norm: one dense d(conv) buffer shared by the layers)
  cudaStream_t sw = fork ? aux->stream : s;              // weight / bias gradients
  int ev = 0;
  auto ready = [&]() -> int {                            // the gradient produced last on `s` may now be consumed on `sw`
    if (!fork) return CPP_OK;
    CPP_CHECK_CUDA(cudaEventRecord(aux->ready[ev], s));
    CPP_CHECK_CUDA(cudaStreamWaitEvent(sw, aux->ready[ev], 0));
    ++ev; aux->used = true;
    return CPP_OK;
  };
  CPP_REQUIRE(d_rep_extra == nullptr || (grads != nullptr && concat_at < 0), "d_rep_extra needs a full backward pass of a network without action input");
  const bool dropping = any_drop && g_is_training;
  float* gmax3 = nullptr;                                 // set when the fused FC input-gradient kernel has left max |gp| of conv3
  const bool fused = fused_mlp_level() >= 2 && mlp_fits(*this) && d_rep_extra == nullptr && !dropping;
  const int stop_at_f = (grads == nullptr) ? concat_at : 0;
  if (fused) {
    // one launch for the whole chain of input gradients; every weight / bias gradient afterwards (side stream if given)
    const bool need_first = (stop_at_f == 0 && pixels && grads != nullptr) || (concat_at == stop_at_f && d_action != nullptr);
    // tensor-core route: the kernel also leaves max |d(flattened conv3 output)| for conv3's gradient pieces (no max pass later)
    if (stop_at_f == 0 && need_first && pixels && !bn && tc_scratch != nullptr && tc_route(is_f16) &&
        tc::conv_dgrad_fused_supported(conv[2].H, conv[2].W, conv[2].KS)) {
      gmax3 = reinterpret_cast<float*>(ws + L.gsc) + 4;
      CPP_CHECK_CUDA(cudaMemsetAsync(gmax3, 0, sizeof(float), s));
    }
    CPP_TRY(launch_mlp_dgrad(*this, params, B, ws, d_out, stop_at_f, need_first ? 1 : 0, d_action, s, gmax3));
    if (grads != nullptr) {
      CPP_TRY(ready());
      for (int i = last; i >= 0; --i) {
        int xld;
        const float* x = fc_input(L, ws, i, &xld);
        const float* dpre = i == last ? reinterpret_cast<const float*>(ws + L.dTop) : reinterpret_cast<const float*>(ws + L.dX[i + 1]);
        const int dld = i == last ? out_dim[last] : in_dim[i + 1];
        CPP_TRY(fc_wgrad(x, xld, dpre, dld, B, in_dim[i], out_dim[i], grads + off_fc_w[i], grads + off_fc_b[i], sw));
      }
    }
  }
  // gradient wrt the last pre-activation
  float* dcur = reinterpret_cast<float*>(ws + L.dTop);
  int dld = out_dim[last];
  int i_first = last;
  const bool tail = !fused && grads != nullptr && d_rep_extra == nullptr && critic_tail_ok(*this);
  if (tail) {
    // pixel critic: q and hidden3 backwards in one launch (dTop, dX[last], dX[concat_at]); their weight gradients follow
    CPP_TRY(launch_critic_tail_bwd(*this, params, d_out, B, ws, s));
    CPP_TRY(ready());
    for (int i = last; i >= concat_at; --i) {
      int xld;
      const float* x = fc_input(L, ws, i, &xld);
      const float* dpre = i == last ? reinterpret_cast<const float*>(ws + L.dTop) : reinterpret_cast<const float*>(ws + L.dX[last]);
      const int pld = i == last ? out_dim[last] : in_dim[last];
      CPP_TRY(fc_wgrad(x, xld, dpre, pld, B, in_dim[i], out_dim[i], grads + off_fc_w[i], grads + off_fc_b[i], sw));
    }
    dcur = reinterpret_cast<float*>(ws + L.dX[concat_at]);
    dld = in_dim[concat_at];
    if (d_action != nullptr)
      CPP_TRY(launch_copy_cols(dcur + (dld - action_dim), dld, B, action_dim, d_action, action_dim, 0, s));
    i_first = concat_at - 1;
  } else if (!fused) {
    CPP_TRY(launch_act_grad(d_out, out_dim[last], reinterpret_cast<const float*>(ws + L.h[last]), out_ld[last], act[last], B,
                            out_dim[last], dcur, out_dim[last], s));
  }
  const int stop_at = (grads == nullptr) ? concat_at : 0;   // only d_action wanted: stop once it is known
  if (fused) dcur = reinterpret_cast<float*>(ws + L.dX[0]);
  for (int i = i_first; i >= stop_at && !fused; --i) {
    int xld;
    const float* x = fc_input(L, ws, i, &xld);
    if (grads != nullptr) {
      CPP_TRY(ready());
      CPP_TRY(fc_wgrad(x, xld, dcur, dld, B, in_dim[i], out_dim[i], grads + off_fc_w[i], grads + off_fc_b[i], sw));
    }
    const bool need_dx = (i > stop_at) || (i == 0 && pixels && grads != nullptr) || (concat_at == i && d_action != nullptr);
    if (!need_dx) break;     // (a low-dim network without hidden layers: the representation is the state, nothing to send)
    float* dnext = reinterpret_cast<float*>(ws + L.dX[i]);
    GemmArgs g{};                                          // dX = dPre . W^T, masked by the ReLU of the layer below
    g.A = dcur; g.lda = dld; g.transA = 0;
    g.B = params + off_fc_w[i]; g.ldb = out_dim[i]; g.transB = 1;
    g.C = dnext; g.ldc = in_dim[i];
    g.M = B; g.N = in_dim[i]; g.K = out_dim[i];
    // behind a dropout layer the input is 2 m relu(z): the stored activation is > 0 exactly where m = 1 and the ReLU is open
    const float gate_scale = (i > 0 && dropping && drop[i - 1]) ? 2.f : 1.f;
    if (i > 0) { g.epi = EPI_RELU_MASK; g.aux = x; g.aux_ld = xld; g.mask_cols = out_dim[i - 1]; g.mask_scale = gate_scale; }
    else g.epi = EPI_NONE;
    // tensor-core route: the layer-0 input gradient is conv3's gp - its GEMM also leaves max |gp| (no max pass before conv3's kernels)
    if (i == 0 && pixels && !bn && grads != nullptr && d_rep_extra == nullptr && tc_scratch != nullptr && tc_route(is_f16) &&
        tc::conv_dgrad_fused_supported(conv[2].H, conv[2].W, conv[2].KS) && !gemm_tc_wanted(g)) {
      gmax3 = reinterpret_cast<float*>(ws + L.gsc) + 4;
      CPP_CHECK_CUDA(cudaMemsetAsync(gmax3, 0, sizeof(float), s));
      g.absmax = gmax3;
    }
    CPP_TRY(launch_gemm(g, s));
    if (i == last && d_rep_extra != nullptr)
      CPP_TRY(launch_add_gated(dnext, in_dim[i], d_rep_extra, i > 0 ? x : nullptr, xld, B, in_dim[i], s, gate_scale));
    if (concat_at == i && d_action != nullptr)
      CPP_TRY(launch_copy_cols(dnext + (in_dim[i] - action_dim), in_dim[i], B, action_dim, d_action, action_dim, 0, s));
    dcur = dnext;
    dld = in_dim[i];
  }
  auto finish = [&]() -> int {                            // mark the end of the side stream's work for the caller's join
    if (fork && aux->used) CPP_CHECK_CUDA(cudaEventRecord(aux->done, sw));
    return CPP_OK;
  };
  if (!pixels || grads == nullptr) return finish();
  trace_mark("   . FC input gradients", s);


  // conv trunk: dcur = d(pooled3) as (B, F); base_network.py:103-123 backwards
  const float* gp = dcur;
  for (int i = 2; i >= 0; --i) {
    const void* x; int xf16; const float* mi;
    if (i == 0) { x = state; xf16 = is_f16; mi = mean_inv; }
    else { x = ws + L.pooled[i - 1]; xf16 = 0; mi = nullptr; }
    const uint8_t* amax = reinterpret_cast<const uint8_t*>(ws + L.amax[i]);
    if (bn) {
      // through slim.batch_norm with batch statistics: dense d(conv) for EVERY position, d(beta); then the dense-gradient kernels
      CPP_REQUIRE(g_is_training, "backward through batch norm needs the batch statistics of a training-mode forward");
      float* dconv = reinterpret_cast<float*>(ws + L.dconv);
      CPP_TRY(launch_bn_backward(gp, amax, reinterpret_cast<const float*>(ws + L.raw[i]), B, conv[i].H, conv[i].W, ws + L.bnscr[i], dconv,
                                 grads + off_conv_b[i], s));
      if (i == 0 && defer_conv1) break;                     // d(conv1) stays in ws: picked up by conv1_wgrad_group
      CPP_TRY(launch_conv_wgrad(conv[i], x, xf16, mi, dconv, nullptr, B, grads + off_conv_w[i], reinterpret_cast<float*>(ws + L.bnjunk),
                                reinterpret_cast<float*>(ws + L.wgrad), s));
      if (i > 0) {
        float* dx = reinterpret_cast<float*>(ws + L.dpool[2 - i]);
        CPP_TRY(launch_conv_dgrad_dense(conv[i], dconv, params + off_conv_w[i], B, dx, s));
        gp = dx;
      }
      continue;
    }
    if (i == 0 && defer_conv1) break;                       // gp == ws + L.dpool[1]: picked up by conv1_wgrad_group
    // tensor-core route for conv3/conv2: the un-pool + split pass also yields max|gp|, shared by wgrad and dgrad
    const bool tc_dg = i > 0 && tc_scratch != nullptr && tc_route(is_f16);
    float* gsc = reinterpret_cast<float*>(ws + L.gsc) + 2 * i;
    // i == 1: conv3's input-gradient kernel has already left max|gp| in gsc[0] (its epilogue), no separate max pass - and the
    // weight gradient (side stream) can start before the un-pool / split pass instead of behind it
    const bool early_wgrad = tc_dg && i == 1;
    // fused: the row-sweep kernel builds its input strips from gp + arg-max itself (conv_row_tc.cu) - no un-pool / split launch, no
    // piece tensor; conv3's max|gp| then needs its own small pass (conv2's comes from conv3's input-gradient epilogue)
    const bool fused_dg = tc_dg && tc::conv_dgrad_fused_supported(conv[i].H, conv[i].W, conv[i].KS);
    auto unpool = [&]() -> int {
      if (fused_dg) {
        if (i != 1 && gmax3 == nullptr) {
          CPP_TRY(tc::launch_absmax(gp, (int64_t)B * conv[i].PH() * conv[i].PW() * kConvCout, gsc, s));
          trace_mark("   . conv3 max|g|", s);
        }
        return CPP_OK;
      }
      CPP_TRY(tc::launch_unpool_split(gp, amax, B, conv[i].H, conv[i].W, gsc, gsc + 1, reinterpret_cast<__half*>(ws + L.dyp), s,
                                      i == 1 ? 1 : 0));
      trace_mark(i == 2 ? "   . conv3 un-pool/split" : "   . conv2 un-pool/split", s);
      return CPP_OK;
    };
    if (tc_dg && !early_wgrad) CPP_TRY(unpool());
    CPP_TRY(ready());                                      // gp (and its max) are complete on the main stream
    const int cap_saved = g_cta_cap;
    if (fork && aux->cta_cap > 0) g_cta_cap = aux->cta_cap;
    struct CapRestore { int v; ~CapRestore() { g_cta_cap = v; } } cap_restore{cap_saved};
    if (i > 0 && wg_scratch != nullptr && tc_route(is_f16) &&
        wg::conv_wgrad_mma_supported(1, conv[i].H, conv[i].W, tc::kC24, conv[i].KS, 2)) {
      const float* g1[1] = {gp}; const uint8_t* a1[1] = {amax};
      float* dw[1] = {grads + off_conv_w[i]}; float* db[1] = {grads + off_conv_b[i]};
      const float* gm[1] = {gsc};
      CPP_TRY(wg::launch_conv_wgrad_mma(ws + L.hl[i - 1], nullptr, 2, 1, g1, a1, B, conv[i].H, conv[i].W, tc::kC24, conv[i].KS,
                                        dw, db, wg_scratch, sw, tc_dg ? gm : nullptr));
    } else {
      CPP_TRY(launch_conv_wgrad(conv[i], x, xf16, mi, gp, amax, B, grads + off_conv_w[i], grads + off_conv_b[i],
                                reinterpret_cast<float*>(ws + L.wgrad), sw));
    }
    g_cta_cap = cap_saved;
    trace_mark(i == 2 ? "   . conv3 wgrad (side)" : (i == 1 ? "   . conv2 wgrad (side)" : "   . conv1 wgrad"), sw);
    if (early_wgrad) CPP_TRY(unpool());
    if (i > 0) {
      float* dx = reinterpret_cast<float*>(ws + L.dpool[2 - i]);   // i=2 -> dpool[0] (pooled2 grad), i=1 -> dpool[1]
      if (fused_dg) {
        CPP_TRY(tc::launch_conv_dgrad_tc_fused(gp, amax, gsc, gsc + 1, 1, params + off_conv_w[i], B, conv[i].H, conv[i].W, conv[i].KS, dx,
                                               reinterpret_cast<char*>(tc_scratch) + (i == 2 ? 3 : 4) * trunk_slot_bytes(*this), s,
                                               reinterpret_cast<float*>(ws + L.gsc) + (i == 1 ? 6 : 2),
                                               g_tc_prepped ? tc::kPhaseMain : tc::kPhaseBoth));
      } else if (tc_dg) {
        CPP_TRY(tc::launch_conv_dgrad_tc(reinterpret_cast<__half*>(ws + L.dyp), gsc + 1, params + off_conv_w[i], B, conv[i].H, conv[i].W,
                                         conv[i].KS, dx, reinterpret_cast<char*>(tc_scratch) + (i == 2 ? 3 : 4) * trunk_slot_bytes(*this), s,
                                         // max|dx| for the next consumer: conv1's wgrad (i == 1), conv2's un-pool/split (i == 2)
                                         reinterpret_cast<float*>(ws + L.gsc) + (i == 1 ? 6 : 2),
                                         g_tc_prepped ? tc::kPhaseMain : tc::kPhaseBoth));
      } else {
        CPP_TRY(launch_conv_dgrad(conv[i], gp, amax, params + off_conv_w[i], B, dx, s));
      }
      gp = dx;
      trace_mark(i == 2 ? "   . conv3 dgrad" : "   . conv2 dgrad", s);
    }
  }
  return finish();
}

}  // namespace cpp
