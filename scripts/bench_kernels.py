#!/usr/bin/env python
"""Kernel-level timings (CUDA events on the launching stream, L2 flushed between launches) of the conv kernels at the
BASELINE config 3 layer shapes, each through its C-ABI entry point on the whole GPU.

  python scripts/bench_kernels.py                 all kernels, one JSON line each
  python scripts/bench_kernels.py --only NAME     one kernel (for `ncu -k regex:... python scripts/bench_kernels.py --only NAME`)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartpoleplusplus_b200 import _lib as L   # noqa: E402
from tests import gpu_util as U               # noqa: E402

B, H, W, CIN, NETS = 256, 64, 64, 9, 2


def timeit(fn, flush, n=10):
  for _ in range(3):
    fn()
  ts = []
  for _ in range(n):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
  return float(np.median(ts)), float(np.min(ts))


def kernel_table(only=None, reps=10, rm=None):
  """[{kernel, what, launches_per_step, us_median, us_min, useful_gflop | bytes, ...}] at the c3 layer shapes; rm: a filled
  ReplayMemory for the gather line (bench.py passes its own)"""
  lib = L.lib()
  if os.environ.get("CONV_ROW") is not None:     # A/B of the conv2 / conv3 route: 1 row-sweep kernel (default), 0 parity-plane kernel
    L.check(lib.cpp_set_option(b"conv_row", int(os.environ["CONV_ROW"])))
  if os.environ.get("WGRAD_TC") is not None:     # conv weight-gradient routes: bit 0 conv1 tcgen05, bit 1 conv2/3 tcgen05 (old), bit 2 conv2/3 row-sweep
    L.check(lib.cpp_set_option(b"wgrad_tc", int(os.environ["WGRAD_TC"])))
  dev = "cuda"
  g = torch.Generator(device=dev); g.manual_seed(0)
  rs = np.random.RandomState(0)
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  st = L.stream_ptr

  # ---- conv1 (fp16 pixels, whitening folded), actor + critic
  x = (torch.randint(0, 256, (B, H, W, CIN), device=dev, generator=g).to(torch.float16) / 255).contiguous()
  scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(CIN)), dtype=torch.float64, device=dev)
  mi = torch.zeros(2 * CIN, dtype=torch.float32, device=dev)
  L.check(lib.cpp_channel_moments(L.ptr(x), 1, C.c_int64(B * H * W), CIN, L.ptr(scratch), L.ptr(mi), st()))
  w1 = [torch.randn(5, 5, CIN, 10, device=dev) * 0.05 for _ in range(NETS)]
  b1 = [torch.randn(10, device=dev) * 0.1 for _ in range(NETS)]
  p1 = [torch.zeros(B, 32, 32, 10, device=dev) for _ in range(NETS)]
  a1 = [torch.zeros(B, 32, 32, 10, dtype=torch.uint8, device=dev) for _ in range(NETS)]
  hl1 = [torch.zeros(B, 32, 32, 24, dtype=torch.float16, device=dev) for _ in range(NETS)]
  scr1 = torch.zeros(int(lib.cpp_conv_tc_scratch_bytes(NETS, H, W, CIN, 5)), dtype=torch.uint8, device=dev)
  pw1, pb1, pp1, pa1, ph1 = L.ptr_array(w1), L.ptr_array(b1), L.ptr_array(p1), L.ptr_array(a1), L.ptr_array(hl1)
  g1 = [torch.randn(B, 32, 32, 10, device=dev) * 1e-3 for _ in range(NETS)]
  dw1 = [torch.zeros(5, 5, CIN, 10, device=dev) for _ in range(NETS)]
  db1 = [torch.zeros(10, device=dev) for _ in range(NETS)]
  wscr1 = torch.zeros(int(lib.cpp_conv_wgrad_mma_scratch_bytes(NETS, H, W, CIN, 5)), dtype=torch.uint8, device=dev)
  pg1, pdw1, pdb1 = L.ptr_array(g1), L.ptr_array(dw1), L.ptr_array(db1)

  def conv1_fwd_tc():
    L.check(lib.cpp_conv_forward_tc(L.ptr(x), None, L.ptr(mi), NETS, pw1, pb1, B, H, W, CIN, 5, pp1, pa1, L.ptr(scr1), st(), 0, ph1))

  def conv1_wgrad_mma():
    L.check(lib.cpp_conv_wgrad_mma(L.ptr(x), L.ptr(mi), 0, NETS, pg1, pa1, B, H, W, CIN, 5, pdw1, pdb1, L.ptr(wscr1), st()))

  # ---- conv2 / conv3 (one network; input = 24-channel fp16 pieces of the layer below)
  def layer(Hh, KS):
    act = np.maximum(rs.randn(B, Hh, Hh, 10), 0).astype(np.float32)
    xin = torch.from_numpy(U.to_c24(act)[0]).to(dev)
    w = torch.randn(KS, KS, 10, 10, device=dev) * 0.05; b = torch.randn(10, device=dev) * 0.1
    po = torch.zeros(B, Hh // 2, Hh // 2, 10, device=dev); am = torch.zeros(B, Hh // 2, Hh // 2, 10, dtype=torch.uint8, device=dev)
    hl = torch.zeros(B, Hh // 2, Hh // 2, 24, dtype=torch.float16, device=dev)
    scr = torch.zeros(int(lib.cpp_conv_tc_scratch_bytes(1, Hh, Hh, 24, KS)), dtype=torch.uint8, device=dev)
    gp = torch.randn(B, Hh // 2, Hh // 2, 10, device=dev) * 1e-3
    dx = torch.zeros(B, Hh, Hh, 10, device=dev)
    dscr = torch.zeros(int(lib.cpp_conv_dgrad_tc_scratch_bytes(B, Hh, Hh, KS)), dtype=torch.uint8, device=dev)
    dw = torch.zeros(KS, KS, 10, 10, device=dev); db = torch.zeros(10, device=dev)
    wscr = torch.zeros(int(lib.cpp_conv_wgrad_mma_scratch_bytes(1, Hh, Hh, 24, KS)), dtype=torch.uint8, device=dev)
    pw, pb, pp, pa, ph = L.ptr_array([w]), L.ptr_array([b]), L.ptr_array([po]), L.ptr_array([am]), L.ptr_array([hl])
    pg, pdw, pdb = L.ptr_array([gp]), L.ptr_array([dw]), L.ptr_array([db])

    def fwd():
      L.check(lib.cpp_conv_forward_tc(L.ptr(xin), None, None, 1, pw, pb, B, Hh, Hh, 24, KS, pp, pa, L.ptr(scr), st(), 2, ph))

    def dgrad():
      L.check(lib.cpp_conv_dgrad_tc(L.ptr(gp), L.ptr(am), L.ptr(w), B, Hh, Hh, KS, L.ptr(dx), L.ptr(dscr), st()))

    def wgrad():
      L.check(lib.cpp_conv_wgrad_mma(L.ptr(xin), None, 2, 1, pg, pa, B, Hh, Hh, 24, KS, pdw, pdb, L.ptr(wscr), st()))
    keep = (xin, w, b, po, am, hl, scr, gp, dx, dscr, dw, db, wscr, pw, pb, pp, pa, ph, pg, pdw, pdb)
    return fwd, dgrad, wgrad, keep

  c2 = layer(32, 5)
  c3 = layer(16, 3)
  conv1_fwd_tc()          # fills the arg-max side band the wgrad reads
  c2[0](); c3[0]()

  # ---- FC stacks (a4): the actor's 640 -> 100 -> 100 -> 50 -> 2 and the critic's 640 -> 200 -> 50 (+2) -> 50 -> 1 through
  # cpp_net_forward / cpp_net_backward of a low-dim network fed with the flattened trunk output
  def fc_net(sizes, acts, concat_at=-1):
    spec = L.NetSpec()
    spec.pixels, spec.input_dim, spec.n_fc, spec.concat_at, spec.action_dim = 0, 640, len(sizes), concat_at, 2 if concat_at >= 0 else 0
    for i, (n_, a_) in enumerate(zip(sizes, acts)):
      spec.fc_out[i], spec.fc_act[i] = n_, a_
    h = C.c_void_p()
    L.check(lib.cpp_net_create(C.byref(spec), C.byref(h)))
    n_par = int(lib.cpp_net_num_params(h))
    P_ = (torch.randn(n_par, device=dev) * 0.05).contiguous()
    G_ = torch.zeros(n_par, device=dev)
    ws = torch.zeros(int(lib.cpp_net_workspace_bytes(h, B)), dtype=torch.uint8, device=dev)
    xin = torch.relu(torch.randn(B, 640, device=dev)).contiguous()
    act = torch.randn(B, 2, device=dev) if concat_at >= 0 else None
    out = torch.zeros(B, sizes[-1], device=dev); dout = torch.randn(B, sizes[-1], device=dev)
    dact = torch.zeros(B, 2, device=dev) if concat_at >= 0 else None

    def fwd():
      L.check(lib.cpp_net_forward(h, L.ptr(P_), L.ptr(xin), 0, None, L.ptr(act), B, L.ptr(ws), L.ptr(out), st()))

    def bwd():
      L.check(lib.cpp_net_backward(h, L.ptr(P_), L.ptr(xin), 0, None, B, L.ptr(ws), L.ptr(dout), L.ptr(G_), L.ptr(dact), st()))
    macs = 0
    d = 640
    for i, n_ in enumerate(sizes):
      macs += (d + (2 if concat_at == i else 0)) * n_
      d = n_
    return fwd, bwd, macs, (h, P_, G_, ws, xin, act, out, dout, dact)

  fa = fc_net([100, 100, 50, 2], [1, 1, 1, 2])
  fcr = fc_net([200, 50, 50, 1], [1, 1, 1, 0], concat_at=2)
  fa[0](); fcr[0]()

  mac1 = B * H * W * 10 * 25 * CIN * NETS
  mac2, mac3 = B * 32 * 32 * 10 * 25 * 10, B * 16 * 16 * 10 * 9 * 10
  # launches per DDPG grad-step at c3 (csrc/agents.cu DDPG::step_body): conv1 fwd {actor,critic}(s1) + {targets}(s2); conv2/3
  # fwd of the 4 networks; conv2/3 dgrad + wgrad of the 2 trained ones; one conv1 wgrad; FC fwd of 4 networks (the critic's
  # tail twice), FC bwd of 2
  kernels = [
      ("conv1_fwd_tc", conv1_fwd_tc, 2.0 * mac1, 2, "conv1 5x5 9->10 fwd, 2 sibling nets in one pass, incl. weight prep"),
      ("conv1_wgrad_mma", conv1_wgrad_mma, 2.0 * mac1, 1, "conv1 weight gradient, actor+critic (tcgen05, csrc/conv_wgrad_tc.cu), incl. absmax/reduce/finalize"),
      ("conv2_fwd_tc", c2[0], 2.0 * mac2, 4, "conv2 5x5 10->10 fwd, one network, row-sweep tcgen05 kernel (csrc/conv_row_tc.cu, TMA tensor-map strips), incl. weight prep"),
      ("conv2_dgrad_tc", c2[1], 2.0 * mac2, 2, "conv2 input gradient (row-sweep kernel, un-pool / split fused into its producer warps), incl. max|g| pass + weight prep"),
      ("conv2_wgrad_mma", c2[2], 2.0 * mac2, 2, "conv2 weight gradient, one network (row-sweep tcgen05, csrc/conv_wgrad_row_tc.cu), incl. max|g| pass + finalize"),
      ("conv3_fwd_tc", c3[0], 2.0 * mac3, 4, "conv3 3x3 10->10 fwd (row-sweep kernel), incl. weight prep"),
      ("conv3_dgrad_tc", c3[1], 2.0 * mac3, 2, "conv3 input gradient (row-sweep kernel, fused un-pool), incl. max|g| pass + weight prep"),
      ("conv3_wgrad_mma", c3[2], 2.0 * mac3, 2, "conv3 weight gradient (row-sweep tcgen05), incl. max|g| pass + finalize"),
      ("fc_actor_fwd", fa[0], 2.0 * B * fa[2], 2, "actor FC stack 640-100-100-50-2 forward (actor, target actor)"),
      ("fc_actor_bwd", fa[1], 4.0 * B * fa[2], 1, "actor FC stack backward: input gradients + weight / bias gradients"),
      ("fc_critic_fwd", fcr[0], 2.0 * B * fcr[2], 2, "critic FC stack 640-200-50(+2)-50-1 forward (critic, target critic)"),
      ("fc_critic_bwd", fcr[1], 4.0 * B * fcr[2], 1, "critic FC stack backward"),
  ]
  rows = []
  for name, fn, flops, per_step, what in kernels:
    if only and name not in only.split(","):
      continue
    med, mn = timeit(fn, flush, reps)
    rows.append(dict(kernel=name, what=what, launches_per_step=per_step, us_median=med, us_min=mn, useful_gflop=flops / 1e9,
                     useful_tflops=flops / (med * 1e-6) / 1e12))
  if rm is not None and (not only or only == "replay_gather"):
    idx = torch.randint(0, rm.buffer_size, (B,), device=dev, dtype=torch.int64)
    idx_h = idx.cpu().numpy()
    out = rm.batch_at(idx_h, d_idxs=idx)
    med, mn = timeit(lambda: rm.batch_at(idx_h, d_idxs=idx, out=out), flush, reps)
    nbytes = 2 * 2 * B * rm.row_elems * 2                 # s1 + s2 rows, read + written, fp16
    rows.append(dict(kernel="replay_gather", what="ReplayMemory.batch gather of state_1 / state_2 rows + action/reward/mask",
                     launches_per_step=1, us_median=med, us_min=mn, bytes=nbytes, gb_per_s=nbytes / (med * 1e-6) / 1e9))
  return rows


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--only", default=None)
  ap.add_argument("--reps", type=int, default=10)
  args = ap.parse_args()
  for r in kernel_table(args.only, args.reps):
    print(json.dumps(r))


if __name__ == "__main__":
  main()
