"""tcgen05 conv forward (cpp_conv_forward_tc) against an fp64 reference and against the exact-fp32 CUDA-core kernel
(cpp_conv_forward), at the conv1 shapes of the BASELINE configs, with sibling networks fused along N, odd sizes,
near-constant ("sparse scene") images that stress the whitening fold, and the fused replay-row gather."""
import ctypes as C
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import gpu_util as U

pytestmark = pytest.mark.gpu


def _lib():
  from cartpoleplusplus_b200 import _lib as L
  return L, L.lib()


def make_input(rs, B, H, W, Cin, sparse):
  if sparse:   # constant background, <= 3 % coloured pixels: tiny channel variance -> large inv
    k = np.full((B, H, W, Cin), 40, dtype=np.int64)
    m = rs.rand(B, H, W, 1) < 0.03
    k = np.where(m, rs.randint(0, 256, (B, H, W, Cin)), k)
  else:
    k = rs.randint(0, 256, (B, H, W, Cin))
  return (k.astype(np.float16) / np.float16(255))


def run_tc(B, H, W, Cin, KS, nets, seed, sparse=False, gather=False):
  L, lib = _lib()
  rs = np.random.RandomState(seed)
  dev = "cuda"
  x_h = make_input(rs, B, H, W, Cin, sparse)
  rows = None
  if gather:   # the batch is rows[b] of a bigger slab
    n_slab = 3 * B
    slab_h = make_input(rs, n_slab, H, W, Cin, sparse)
    rows_h = rs.randint(0, n_slab, B).astype(np.int32)
    x_h = slab_h[rows_h]
    slab = torch.from_numpy(slab_h).to(dev)
    rows = torch.from_numpy(rows_h).to(dev)
  x = torch.from_numpy(x_h).to(dev)
  scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(Cin)), dtype=torch.float64, device=dev)
  mi = torch.zeros(2 * Cin, dtype=torch.float32, device=dev)
  L.check(lib.cpp_channel_moments(L.ptr(x), 1, C.c_int64(B * H * W), Cin, L.ptr(scratch), L.ptr(mi), L.stream_ptr()))
  x64 = torch.from_numpy(x_h.astype(np.float64))
  mean = x64.mean(dim=(0, 1, 2)); var = ((x64 - mean) ** 2).mean(dim=(0, 1, 2))
  xw = ((x64 - mean) * torch.rsqrt(var + 1e-6)).permute(0, 3, 1, 2).contiguous()
  lim = np.sqrt(6.0 / (KS * KS * (Cin + 10)))
  PH, PW = H // 2, W // 2
  ws, bs, pooled, amax, ref = [], [], [], [], []
  for n in range(nets):
    w_h = rs.uniform(-lim, lim, (KS, KS, Cin, 10)).astype(np.float32)
    b_h = rs.uniform(-0.1, 0.1, 10).astype(np.float32)
    ws.append(torch.from_numpy(w_h).to(dev)); bs.append(torch.from_numpy(b_h).to(dev))
    pooled.append(torch.full((B, PH, PW, 10), -7.0, dtype=torch.float32, device=dev))
    amax.append(torch.full((B, PH, PW, 10), 9, dtype=torch.uint8, device=dev))
    y = F.conv2d(xw, torch.from_numpy(w_h.astype(np.float64)).permute(3, 2, 0, 1).contiguous(),
                 torch.from_numpy(b_h.astype(np.float64)), padding=KS // 2)
    ref.append(y)
  nb = int(lib.cpp_conv_tc_scratch_bytes(nets, H, W, Cin, KS))
  assert nb > 0
  scr = torch.zeros(nb, dtype=torch.uint8, device=dev)
  src = slab if gather else x
  L.check(lib.cpp_conv_forward_tc(L.ptr(src), L.ptr(rows), L.ptr(mi), nets, L.ptr_array(ws), L.ptr_array(bs), B, H, W, Cin, KS,
                                  L.ptr_array(pooled), L.ptr_array(amax), L.ptr(scr), L.stream_ptr(), 0, None))
  torch.cuda.synchronize()
  errs = []
  for n in range(nets):
    y = ref[n]
    ref_pool = F.max_pool2d(F.relu(y), 2).permute(0, 2, 3, 1).numpy()
    got = pooled[n].cpu().numpy()
    errs.append(U.assert_close(got, ref_pool, what="tc conv fwd net %d %dx%dx%d k%d" % (n, H, W, Cin, KS)))
    a = amax[n].cpu().numpy().astype(np.int64)
    yw = y[:, :, :PH * 2, :PW * 2].reshape(B, 10, PH, 2, PW, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, PH, PW, 10, 4).numpy()
    closed = a == 4
    assert a.max() <= 4
    assert np.array_equal(closed, got == 0)
    picked = np.take_along_axis(yw, np.minimum(a, 3)[..., None], axis=-1)[..., 0]
    scale = np.abs(yw).max()
    assert np.all(np.abs(picked - yw.max(-1))[~closed] <= 1e-5 * scale)
    assert np.all(yw.max(-1)[closed] <= 1e-5 * scale)
    # the exact-fp32 CUDA-core kernel on the same input: same values to fp32 rounding
    p2 = torch.zeros((B, PH, PW, 10), dtype=torch.float32, device=dev); a2 = torch.zeros((B, PH, PW, 10), dtype=torch.uint8, device=dev)
    L.check(lib.cpp_conv_forward(L.ptr(x), 1, L.ptr(mi), L.ptr(ws[n]), L.ptr(bs[n]), B, H, W, Cin, KS, L.ptr(p2), L.ptr(a2), L.stream_ptr()))
    U.assert_close(got, p2.cpu().numpy(), what="tc vs fp32 kernel")
  return errs


@pytest.mark.parametrize("B,H,W,Cin,KS,nets", [
    (8, 64, 64, 9, 5, 2),       # c3 conv1, actor+critic
    (4, 64, 64, 18, 5, 3),      # c4 conv1, NAF value/mu/l
    (2, 128, 128, 24, 5, 2),    # c5 conv1
    (3, 64, 64, 9, 5, 1),       # a single network (action_given)
    (2, 50, 50, 6, 5, 2),       # reference default render size, R=2 (6 channels: padded group)
    (2, 32, 48, 8, 5, 2),       # exactly one channel group, W != H
    (3, 33, 31, 9, 5, 2),       # odd sizes: VALID pooling drops the last row/column
    (4, 32, 32, 10, 3, 1),      # 3x3 kernel, 8+2 channels
])
def test_conv_tc_forward(B, H, W, Cin, KS, nets):
  errs = run_tc(B, H, W, Cin, KS, nets, seed=B * 1000 + H + Cin)
  print("tc conv fwd B%d %dx%dx%d k%d nets%d: rel err vs fp64" % (B, H, W, Cin, KS, nets), ["%.2e" % e for e in errs])


def test_conv_tc_sparse_scene():
  errs = run_tc(4, 64, 64, 9, 5, 2, seed=5, sparse=True)
  print("tc conv fwd sparse scene: rel err vs fp64", ["%.2e" % e for e in errs])


def test_conv_tc_fused_gather():
  run_tc(6, 64, 64, 9, 5, 2, seed=11, gather=True)


def test_conv_tc_full_batch_c3():
  errs = run_tc(256, 64, 64, 9, 5, 2, seed=3)
  print("tc conv fwd c3 full batch: rel err vs fp64", ["%.2e" % e for e in errs])


def run_tc_pieces(B, H, W, KS, seed):
  """conv2 / conv3 on the tensor cores: the input is an fp32 activation handed over as fp16 pieces (24-channel layout), the
  output comes back as fp32 + arg-max + its own piece copy for the next layer"""
  L, lib = _lib()
  rs = np.random.RandomState(seed)
  dev = "cuda"
  Cin = 10
  x32 = (np.maximum(rs.randn(B, H, W, Cin), 0) * 10.0 ** rs.uniform(-2, 1)).astype(np.float32)
  x_h, x_exact = U.to_c24(x32)
  x = torch.from_numpy(x_h).to(dev)
  x64 = torch.from_numpy(x_exact).permute(0, 3, 1, 2).contiguous()
  lim = np.sqrt(6.0 / (KS * KS * (Cin + 10)))
  w_h = rs.uniform(-lim, lim, (KS, KS, Cin, 10)).astype(np.float32); b_h = rs.uniform(-0.1, 0.1, 10).astype(np.float32)
  w, b = torch.from_numpy(w_h).to(dev), torch.from_numpy(b_h).to(dev)
  PH, PW = H // 2, W // 2
  pooled = torch.full((B, PH, PW, 10), -7.0, dtype=torch.float32, device=dev)
  amax = torch.full((B, PH, PW, 10), 9, dtype=torch.uint8, device=dev)
  hl = torch.full((B, PH, PW, 24), -7.0, dtype=torch.float16, device=dev)
  nb = int(lib.cpp_conv_tc_scratch_bytes(1, H, W, 24, KS))
  assert nb > 0
  scr = torch.zeros(nb, dtype=torch.uint8, device=dev)
  L.check(lib.cpp_conv_forward_tc(L.ptr(x), None, None, 1, L.ptr_array([w]), L.ptr_array([b]), B, H, W, 24, KS,
                                  L.ptr_array([pooled]), L.ptr_array([amax]), L.ptr(scr), L.stream_ptr(), 2, L.ptr_array([hl])))
  torch.cuda.synchronize()
  y = F.conv2d(x64, torch.from_numpy(w_h.astype(np.float64)).permute(3, 2, 0, 1).contiguous(), torch.from_numpy(b_h.astype(np.float64)),
               padding=KS // 2)
  ref_pool = F.max_pool2d(F.relu(y), 2).permute(0, 2, 3, 1).numpy()
  got = pooled.cpu().numpy()
  e = U.assert_close(got, ref_pool, what="tc conv from pieces %dx%d k%d" % (H, W, KS))
  a = amax.cpu().numpy().astype(np.int64)
  yw = y[:, :, :PH * 2, :PW * 2].reshape(B, 10, PH, 2, PW, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, PH, PW, 10, 4).numpy()
  closed = a == 4
  assert a.max() <= 4 and np.array_equal(closed, got == 0)
  picked = np.take_along_axis(yw, np.minimum(a, 3)[..., None], axis=-1)[..., 0]
  scale = np.abs(yw).max()
  assert np.all(np.abs(picked - yw.max(-1))[~closed] <= 1e-5 * scale)
  hi_o, lo_o, one_o, pad_o = U.from_c24(hl.cpu().numpy())
  assert np.array_equal(hi_o.astype(np.float16), got.astype(np.float16))                 # hi = round-to-nearest fp16 of the fp32 value
  assert np.abs(hi_o + lo_o - got).max() <= 2.0 ** -20 * max(np.abs(got).max(), 1e-30)
  assert np.all(one_o == 1.0) and np.all(pad_o == 0.0)
  # the exact-fp32 CUDA-core kernel on the fp32 activation
  p2 = torch.zeros((B, PH, PW, 10), dtype=torch.float32, device=dev); a2 = torch.zeros((B, PH, PW, 10), dtype=torch.uint8, device=dev)
  x32d = torch.from_numpy(x32).to(dev)
  L.check(lib.cpp_conv_forward(L.ptr(x32d), 0, None, L.ptr(w), L.ptr(b), B, H, W, Cin, KS, L.ptr(p2), L.ptr(a2), L.stream_ptr()))
  U.assert_close(got, p2.cpu().numpy(), what="tc from pieces vs fp32 kernel")
  return e


def _with_route(row, fn):
  """row = 1: the row-sweep kernel (conv_row_tc.cu, ky taps along N; TMA tensor-map strips) where it covers the shape (input gradient: fused un-pool producer), 3: the same with cp.async strips, 5: input gradient from a piece tensor;
  row = 0: the parity-plane kernel (conv_tc.cu) everywhere"""
  L, lib = _lib()
  L.check(lib.cpp_set_option(b"conv_row", row))
  try:
    return fn()
  finally:
    L.check(lib.cpp_set_option(b"conv_row", 1))


@pytest.mark.parametrize("row", [1, 3, 5, 0])
@pytest.mark.parametrize("B,H,W,KS", [(8, 32, 32, 5), (8, 16, 16, 3), (2, 64, 64, 5), (4, 32, 32, 3), (3, 25, 25, 5), (3, 12, 12, 3),
                                      (256, 32, 32, 5), (256, 16, 16, 3), (1, 32, 32, 5), (7, 16, 16, 3), (5, 32, 48, 5), (3, 48, 32, 3),
                                      (2, 124, 64, 5), (2, 16, 124, 5), (130, 2, 2, 3), (9, 6, 4, 5)])
def test_conv_tc_from_pieces(B, H, W, KS, row):
  e = _with_route(row, lambda: run_tc_pieces(B, H, W, KS, seed=B + H + KS))
  print("tc conv from pieces B%d %dx%d k%d row=%d: rel err vs fp64 %.2e" % (B, H, W, KS, row, e))


def run_dgrad_tc(B, H, W, KS, seed):
  """input gradient of a 10 -> 10 layer on the tensor cores vs fp64 autograd with the same routing, and vs the CUDA-core kernel"""
  L, lib = _lib()
  rs = np.random.RandomState(seed)
  dev = "cuda"
  PH, PW = H // 2, W // 2
  lim = np.sqrt(6.0 / (KS * KS * 20))
  w_h = rs.uniform(-lim, lim, (KS, KS, 10, 10)).astype(np.float32)
  a = rs.randint(0, 5, (B, PH, PW, 10))                                 # any routing, incl. closed gates (4)
  gp_h = (rs.randn(B, PH, PW, 10) * 10.0 ** rs.uniform(-7, 0)).astype(np.float32)
  onehot = np.zeros((B, PH, PW, 10, 4))
  np.put_along_axis(onehot, np.minimum(a, 3)[..., None], 1.0, axis=-1)
  onehot[a == 4] = 0.0
  gfull = np.zeros((B, 10, H, W))
  gfull[:, :, :PH * 2, :PW * 2] = (onehot * gp_h.astype(np.float64)[..., None]).reshape(B, PH, PW, 10, 2, 2).transpose(0, 3, 1, 4, 2, 5).reshape(B, 10, PH * 2, PW * 2)
  xin = torch.zeros((B, 10, H, W), dtype=torch.float64, requires_grad=True)
  y = F.conv2d(xin, torch.from_numpy(w_h.astype(np.float64)).permute(3, 2, 0, 1).contiguous(), padding=KS // 2)
  gx, = torch.autograd.grad(y, [xin], grad_outputs=torch.from_numpy(gfull))
  want = gx.permute(0, 2, 3, 1).numpy()
  gp = torch.from_numpy(gp_h).to(dev); am = torch.from_numpy(a.astype(np.uint8)).to(dev); w = torch.from_numpy(w_h).to(dev)
  dx = torch.full((B, H, W, 10), 7.0, dtype=torch.float32, device=dev)
  nb = int(lib.cpp_conv_dgrad_tc_scratch_bytes(B, H, W, KS))
  assert nb > 0
  scr = torch.zeros(nb, dtype=torch.uint8, device=dev)
  L.check(lib.cpp_conv_dgrad_tc(L.ptr(gp), L.ptr(am), L.ptr(w), B, H, W, KS, L.ptr(dx), L.ptr(scr), L.stream_ptr()))
  torch.cuda.synchronize()
  e = U.assert_close(dx.cpu().numpy(), want, what="dgrad_tc %dx%d k%d" % (H, W, KS))
  dx2 = torch.zeros((B, H, W, 10), dtype=torch.float32, device=dev)
  L.check(lib.cpp_conv_dgrad(L.ptr(gp), L.ptr(am), L.ptr(w), B, H, W, KS, L.ptr(dx2), L.stream_ptr()))
  U.assert_close(dx.cpu().numpy(), dx2.cpu().numpy(), what="dgrad_tc vs fp32 kernel")
  return e


@pytest.mark.parametrize("row", [1, 3, 5, 0])
@pytest.mark.parametrize("B,H,W,KS", [(8, 32, 32, 5), (8, 16, 16, 3), (2, 64, 64, 5), (3, 25, 25, 5), (3, 12, 13, 3), (256, 32, 32, 5),
                                      (256, 16, 16, 3), (1, 4, 4, 3), (5, 32, 48, 5), (2, 124, 64, 5), (130, 2, 2, 3)])
def test_conv_dgrad_tc(B, H, W, KS, row):
  e = _with_route(row, lambda: run_dgrad_tc(B, H, W, KS, seed=B + H + KS))
  print("dgrad_tc B%d %dx%d k%d row=%d: rel err vs fp64 %.2e" % (B, H, W, KS, row, e))


def test_piece_overflow_is_counted_not_silent():
  """an activation above 2 x 65504 cannot be carried by the fp16 hi + lo piece copy that feeds the next tensor-core layer: the
  fp32 output stays exact, the copy saturates, and cpp_piece_overflow_count reports it (EngineBase.check_piece_overflow raises)"""
  import ctypes as C
  from cartpoleplusplus_b200 import _lib as L
  lib = L.lib()
  B, H, W, Cin = 2, 16, 16, 9
  dev = "cuda"
  x = torch.ones(B, H, W, Cin, dtype=torch.float16, device=dev)
  mi = torch.cat([torch.zeros(Cin), torch.ones(Cin)]).to(dev)                 # no whitening: mean 0, inv 1
  w = [torch.full((5, 5, Cin, 10), 1000.0, device=dev)]                       # 225 taps x 1000 = 225 000 > 131 008 in the interior
  b = [torch.zeros(10, device=dev)]
  pooled = [torch.zeros(B, H // 2, W // 2, 10, device=dev)]
  amax = [torch.zeros(B, H // 2, W // 2, 10, dtype=torch.uint8, device=dev)]
  hl = [torch.zeros(B, H // 2, W // 2, 24, dtype=torch.float16, device=dev)]
  scr = torch.zeros(int(lib.cpp_conv_tc_scratch_bytes(1, H, W, Cin, 5)), dtype=torch.uint8, device=dev)
  assert int(lib.cpp_piece_overflow_count(1)) >= 0
  L.check(lib.cpp_conv_forward_tc(L.ptr(x), None, L.ptr(mi), 1, L.ptr_array(w), L.ptr_array(b), B, H, W, Cin, 5, L.ptr_array(pooled),
                                  L.ptr_array(amax), L.ptr(scr), L.stream_ptr(), 0, L.ptr_array(hl)))
  torch.cuda.synchronize()
  assert abs(float(pooled[0].max()) - 225000.0) < 1.0                          # the fp32 output is exact
  n = int(lib.cpp_piece_overflow_count(1))
  assert n > 0, "saturated piece copies were not counted"
  assert int(lib.cpp_piece_overflow_count(0)) == 0                             # reset
