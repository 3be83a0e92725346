"""The arithmetic the planned --use-batch-norm kernels will do, restated in numpy and checked
against the oracle's autograd graph.  No kernels exist yet; this pins the two identities the design relies on so that the
round that writes them starts from verified formulas:

  forward : slim.batch_norm(center=True, scale=False) is a per-channel INCREASING affine map, so
            maxpool(relu(BN(x))) == relu(inv * maxpool(x) + (beta - mean * inv)) with the SAME arg-max as maxpool(x):
            the fused conv + max-pool epilogue stays, plus per-channel sums of x and x^2 over all pre-pool positions;
  backward: with dyh = the pooled gradient routed to the arg-max position through the ReLU gate (dense, mostly zero),
            dbeta = sum dyh,  dx = inv * (dyh - mean(dyh) - xhat * mean(dyh * xhat)),  xhat = (x - mean) * inv.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import nets_oracle as no


def _oracle_layer(x, beta):
  """relu(BN_train(x)) -> 2x2/2 max-pool, exactly as oracle.conv_trunk composes them (NCHW)"""
  zero = torch.zeros_like(beta)
  return F.max_pool2d(F.relu(no.batch_norm(x, beta, zero, zero + 1)), 2)


def test_pool_commutes_with_the_batch_norm_affine_map():
  rs = np.random.RandomState(0)
  x = rs.randn(5, 10, 12, 14)
  beta = rs.randn(10) * 0.3
  want = _oracle_layer(torch.tensor(x), torch.tensor(beta)).numpy()
  mean, var = x.mean(axis=(0, 2, 3)), x.var(axis=(0, 2, 3))
  inv = 1.0 / np.sqrt(var + no.BN_EPSILON)
  win = x.reshape(5, 10, 6, 2, 7, 2).transpose(0, 1, 2, 4, 3, 5).reshape(5, 10, 6, 7, 4)        # the four window positions last
  pooled_raw, arg_raw = win.max(axis=-1), win.argmax(axis=-1)
  got = np.maximum(inv[None, :, None, None] * pooled_raw + (beta - mean * inv)[None, :, None, None], 0.0)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
  # the arg-max of the raw window is the arg-max of the normalised window (strictly increasing map)
  y = (x - mean[None, :, None, None]) * inv[None, :, None, None] + beta[None, :, None, None]
  wy = y.reshape(5, 10, 6, 2, 7, 2).transpose(0, 1, 2, 4, 3, 5).reshape(5, 10, 6, 7, 4)
  assert np.array_equal(wy.argmax(axis=-1), arg_raw)
  # and the statistics are plain sums over ALL pre-pool positions: sum x, sum x^2 per channel
  n = 5 * 12 * 14
  s1, s2 = x.sum(axis=(0, 2, 3)), (x * x).sum(axis=(0, 2, 3))
  np.testing.assert_allclose(s1 / n, mean, rtol=1e-12)
  np.testing.assert_allclose(s2 / n - (s1 / n) ** 2, var, rtol=1e-9)


def test_batch_norm_backward_closed_form():
  rs = np.random.RandomState(1)
  xn = rs.randn(4, 10, 8, 6)
  bn = rs.randn(10) * 0.3
  gp = rs.randn(4, 10, 4, 3)                                       # gradient arriving at the pooled output
  x = torch.tensor(xn, requires_grad=True); beta = torch.tensor(bn, requires_grad=True)
  out = _oracle_layer(x, beta)
  dx_want, dbeta_want = torch.autograd.grad(out, [x, beta], grad_outputs=torch.tensor(gp))
  # --- the planned kernels' arithmetic
  N = 4 * 8 * 6
  mean, var = xn.mean(axis=(0, 2, 3)), xn.var(axis=(0, 2, 3))
  inv = 1.0 / np.sqrt(var + no.BN_EPSILON)
  xhat = (xn - mean[None, :, None, None]) * inv[None, :, None, None]
  y = xhat + bn[None, :, None, None]
  win = y.reshape(4, 10, 4, 2, 3, 2).transpose(0, 1, 2, 4, 3, 5).reshape(4, 10, 4, 3, 4)
  arg, best = win.argmax(axis=-1), win.max(axis=-1)
  routed = np.zeros_like(win)                                     # un-pool through the arg-max byte and the ReLU gate
  np.put_along_axis(routed, arg[..., None], (gp * (best > 0))[..., None], axis=-1)
  dyh = routed.reshape(4, 10, 4, 3, 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(4, 10, 8, 6)
  dbeta = dyh.sum(axis=(0, 2, 3))
  m1 = dyh.sum(axis=(0, 2, 3)) / N                                 # two per-channel reductions: sum dyh, sum dyh * xhat
  m2 = (dyh * xhat).sum(axis=(0, 2, 3)) / N
  dx = inv[None, :, None, None] * (dyh - m1[None, :, None, None] - xhat * m2[None, :, None, None])
  np.testing.assert_allclose(dbeta, dbeta_want.numpy(), rtol=1e-11, atol=1e-12)
  np.testing.assert_allclose(dx, dx_want.numpy(), rtol=1e-9, atol=1e-12)
  assert (dx != 0).mean() > 0.99                                   # dense: why the weight-gradient staging needs a dense-dY variant
