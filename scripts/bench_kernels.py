#!/usr/bin/env python
"""Kernel-level timings (CUDA events, L2 flushed between launches) of the conv layers at the BASELINE shapes:
exact-fp32 CUDA-core kernels vs the tcgen05 path.  usage: python scripts/bench_kernels.py [c3|c4|c5]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartpoleplusplus_b200 import _lib as L   # noqa: E402

CFG = dict(c3=(256, 64, 64, 9, 2), c4=(128, 64, 64, 18, 3), c5=(128, 128, 128, 24, 2))


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


def main():
  name = sys.argv[1] if len(sys.argv) > 1 else "c3"
  B, H, W, Cin, nets = CFG[name]
  lib = L.lib()
  dev = "cuda"
  g = torch.Generator(device=dev); g.manual_seed(0)
  x = (torch.randint(0, 256, (B, H, W, Cin), device=dev, generator=g).to(torch.float16) / 255).contiguous()
  scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(Cin)), dtype=torch.float64, device=dev)
  mi = torch.zeros(2 * Cin, dtype=torch.float32, device=dev)
  L.check(lib.cpp_channel_moments(L.ptr(x), 1, C.c_int64(B * H * W), Cin, L.ptr(scratch), L.ptr(mi), L.stream_ptr()))
  ws = [torch.randn(5, 5, Cin, 10, device=dev) * 0.05 for _ in range(nets)]
  bs = [torch.randn(10, device=dev) * 0.1 for _ in range(nets)]
  pooled = [torch.zeros(B, H // 2, W // 2, 10, device=dev) for _ in range(nets)]
  amax = [torch.zeros(B, H // 2, W // 2, 10, dtype=torch.uint8, device=dev) for _ in range(nets)]
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  scr = torch.zeros(int(lib.cpp_conv_tc_scratch_bytes(nets, H, W, Cin, 5)), dtype=torch.uint8, device=dev)
  wa, ba, pa, aa = L.ptr_array(ws), L.ptr_array(bs), L.ptr_array(pooled), L.ptr_array(amax)

  def ffma():
    for n in range(nets):
      L.check(lib.cpp_conv_forward(L.ptr(x), 1, L.ptr(mi), L.ptr(ws[n]), L.ptr(bs[n]), B, H, W, Cin, 5, L.ptr(pooled[n]), L.ptr(amax[n]), L.stream_ptr()))

  def tc():
    L.check(lib.cpp_conv_forward_tc(L.ptr(x), None, L.ptr(mi), nets, wa, ba, B, H, W, Cin, 5, pa, aa, L.ptr(scr), L.stream_ptr()))

  flops = 2.0 * B * H * W * 10 * 25 * Cin * nets
  out = dict(config=name, B=B, H=H, W=W, Cin=Cin, nets=nets, useful_gflop=flops / 1e9)
  for nm, fn in (("ffma", ffma), ("tcgen05", tc)):
    med, mn = timeit(fn, flush)
    out[nm] = dict(us_median=med, us_min=mn, useful_tflops=flops / (med * 1e-6) / 1e12)
  print(json.dumps(out))


if __name__ == "__main__":
  main()
