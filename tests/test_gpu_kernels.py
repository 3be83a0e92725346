"""Kernel-level GPU parity through the C ABI (cpp_conv_forward / cpp_conv_dgrad / cpp_conv_wgrad /
cpp_channel_moments), at the exact layer shapes of the BASELINE configs.  The backward references are computed in
fp64 with the SAME max-pool/ReLU routing (the amax side band the forward kernel produced), so no gate decision can
differ between the two sides and the 1e-5 tolerance applies at full size."""
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


def whiten64(x):
  mean = x.mean(dim=(0, 1, 2)); var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
  return (x - mean) * torch.rsqrt(var + 1e-6)


def run_layer(B, H, W, Cin, KS, first, seed):
  """first=True: fp16 pixel input + whitening (conv1); else fp32 activations (conv2/conv3)"""
  L, lib = _lib()
  rs = np.random.RandomState(seed)
  dev = "cuda"
  if first:
    x_h = (rs.randint(0, 256, (B, H, W, Cin)).astype(np.float16) / np.float16(255))
    x = torch.from_numpy(x_h).to(dev)
    scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(Cin)), dtype=torch.float64, device=dev)
    mi = torch.zeros(2 * Cin, dtype=torch.float32, device=dev)
    L.check(lib.cpp_channel_moments(L.ptr(x), 1, C.c_int64(B * H * W), Cin, L.ptr(scratch), L.ptr(mi), L.stream_ptr()))
    x64 = whiten64(torch.from_numpy(x_h.astype(np.float64)))
  else:
    x_h = np.maximum(rs.randn(B, H, W, Cin), 0).astype(np.float32)      # post-ReLU activations: ~half zeros
    x = torch.from_numpy(x_h).to(dev)
    mi = None
    x64 = torch.from_numpy(x_h.astype(np.float64))
  lim = np.sqrt(6.0 / (KS * KS * (Cin + 10)))
  w_h = rs.uniform(-lim, lim, (KS, KS, Cin, 10)).astype(np.float32)
  b_h = rs.uniform(-0.1, 0.1, 10).astype(np.float32)
  w, b = torch.from_numpy(w_h).to(dev), torch.from_numpy(b_h).to(dev)
  PH, PW = H // 2, W // 2
  pooled = torch.zeros((B, PH, PW, 10), dtype=torch.float32, device=dev)
  amax = torch.zeros((B, PH, PW, 10), dtype=torch.uint8, device=dev)
  L.check(lib.cpp_conv_forward(L.ptr(x), 1 if first else 0, L.ptr(mi), L.ptr(w), L.ptr(b), B, H, W, Cin, KS,
                               L.ptr(pooled), L.ptr(amax), L.stream_ptr()))
  # ---- forward reference (fp64)
  w64 = torch.from_numpy(w_h.astype(np.float64)).permute(3, 2, 0, 1).contiguous().requires_grad_(True)
  b64 = torch.from_numpy(b_h.astype(np.float64)).requires_grad_(True)
  xin = x64.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  y = F.conv2d(xin, w64, b64, padding=KS // 2)                         # (B,10,H,W) pre-activation
  ref_pool = F.max_pool2d(F.relu(y), 2).permute(0, 2, 3, 1)
  e_fwd = U.assert_close(pooled.cpu().numpy(), ref_pool.detach().numpy(), what="conv fwd %dx%dx%d k%d" % (H, W, Cin, KS))
  # amax points at a window element that attains the max (up to fp32 rounding), 4 <=> pooled == 0
  a = amax.cpu().numpy().astype(np.int64)
  yw = y.detach()[:, :, :PH * 2, :PW * 2].reshape(B, 10, PH, 2, PW, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, PH, PW, 10, 4).numpy()
  closed = a == 4
  assert np.array_equal(closed, pooled.cpu().numpy() == 0)
  picked = np.take_along_axis(yw, np.minimum(a, 3)[..., None], axis=-1)[..., 0]
  scale = np.abs(yw).max()
  assert np.all(np.abs(picked - yw.max(-1))[~closed] <= 1e-5 * scale)
  assert np.all(yw.max(-1)[closed] <= 1e-5 * scale)
  # ---- backward with the GPU's routing
  gp_h = rs.randn(B, PH, PW, 10).astype(np.float32)
  gp = torch.from_numpy(gp_h).to(dev)
  onehot = np.zeros((B, PH, PW, 10, 4))
  np.put_along_axis(onehot, np.minimum(a, 3)[..., None], 1.0, axis=-1)
  onehot[closed] = 0.0
  gfull = np.zeros((B, 10, H, W))
  gw = (onehot * gp_h.astype(np.float64)[..., None]).reshape(B, PH, PW, 10, 2, 2).transpose(0, 3, 1, 4, 2, 5).reshape(B, 10, PH * 2, PW * 2)
  gfull[:, :, :PH * 2, :PW * 2] = gw
  gx, gw64, gb64 = torch.autograd.grad(y, [xin, w64, b64], grad_outputs=torch.from_numpy(gfull))
  scr = torch.zeros(int(lib.cpp_conv_wgrad_scratch_floats(H, W, Cin, KS)), dtype=torch.float32, device=dev)
  dw = torch.zeros((KS, KS, Cin, 10), dtype=torch.float32, device=dev); db = torch.zeros(10, dtype=torch.float32, device=dev)
  L.check(lib.cpp_conv_wgrad(L.ptr(x), 1 if first else 0, L.ptr(mi), L.ptr(gp), L.ptr(amax), B, H, W, Cin, KS,
                             L.ptr(dw), L.ptr(db), L.ptr(scr), L.stream_ptr()))
  e_w = U.assert_close(dw.cpu().numpy(), gw64.permute(2, 3, 1, 0).numpy(), what="conv wgrad")
  e_b = U.assert_close(db.cpu().numpy(), gb64.numpy(), what="conv bias grad")
  e_d = None
  if Cin == 10:
    dx = torch.zeros((B, H, W, 10), dtype=torch.float32, device=dev)
    L.check(lib.cpp_conv_dgrad(L.ptr(gp), L.ptr(amax), L.ptr(w), B, H, W, KS, L.ptr(dx), L.stream_ptr()))
    e_d = U.assert_close(dx.cpu().numpy(), gx.permute(0, 2, 3, 1).numpy(), what="conv dgrad")
  return dict(fwd=e_fwd, wgrad=e_w, bgrad=e_b, dgrad=e_d)


LAYERS = [
    # (B, H, W, Cin, KS, first)                                 config
    (256, 64, 64, 9, 5, True),      # c3 conv1
    (256, 32, 32, 10, 5, False),    # c3/c4 conv2
    (256, 16, 16, 10, 3, False),    # c3/c4 conv3
    (64, 64, 64, 18, 5, True),      # c4 conv1 (per-GPU shard 128; 64 keeps the fp64 reference quick)
    (8, 128, 128, 24, 5, True),     # c5 conv1
    (16, 64, 64, 10, 5, False),     # c5 conv2
    (32, 32, 32, 10, 3, False),     # c5 conv3
    (7, 50, 50, 6, 5, True),        # reference default render 50x50 (R=2): 50 -> 25 -> 12 -> 6
    (7, 25, 25, 10, 5, False),      # odd input: VALID pooling drops the last row/col
    (7, 12, 12, 10, 3, False),
    (3, 22, 18, 15, 5, True),       # ragged, channels not a multiple of anything
    (1, 64, 64, 9, 5, True),        # B=1 action_given path
    (5, 8, 8, 3, 5, True),          # smallest legal image
]


@pytest.mark.parametrize("B,H,W,Cin,KS,first", LAYERS, ids=["%dx%dx%dx%d_k%d" % l[:5] for l in LAYERS])
def test_conv_layer_fwd_dgrad_wgrad(B, H, W, Cin, KS, first):
  print(run_layer(B, H, W, Cin, KS, first, seed=B + H + Cin))


@pytest.mark.parametrize("C,n_pix,f16", [(9, 256 * 64 * 64, True), (24, 8 * 128 * 128, True), (18, 1000, True), (6, 37, False), (3, 1, True)])
def test_channel_moments(C, n_pix, f16):
  L, lib = _lib()
  rs = np.random.RandomState(C)
  if f16:
    xh = (rs.randint(0, 256, (n_pix, C)).astype(np.float16) / np.float16(255))
  else:
    xh = rs.randn(n_pix, C).astype(np.float32)
  x = torch.from_numpy(xh).cuda()
  scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(C)), dtype=torch.float64, device="cuda")
  out = torch.zeros(2 * C, dtype=torch.float32, device="cuda")
  L.check(lib.cpp_channel_moments(L.ptr(x), 1 if f16 else 0, n_pix, C, L.ptr(scratch), L.ptr(out), L.stream_ptr()))
  x64 = xh.astype(np.float64)
  got = out.cpu().numpy().astype(np.float64)
  np.testing.assert_allclose(got[:C], x64.mean(0), rtol=1e-7, atol=1e-9)
  np.testing.assert_allclose(got[C:], 1.0 / np.sqrt(x64.var(0) + 1e-6), rtol=2e-7)
  # constant channel: variance exactly 0 -> inv = 1000, whitened value exactly 0 (sparse-scene stress)
  xc = torch.full((4096, C), 200.0 / 255.0, dtype=torch.float16, device="cuda")
  L.check(lib.cpp_channel_moments(L.ptr(xc), 1, 4096, C, L.ptr(scratch), L.ptr(out), L.stream_ptr()))
  o = out.cpu().numpy()
  assert np.all(o[:C] == np.float32(np.float16(200.0 / 255.0))) and np.allclose(o[C:], 1000.0, rtol=1e-6)
