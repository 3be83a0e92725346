"""CPU restatement of the index algebra of the round-5 row-sweep kernels (csrc/conv_row_tc.cu, csrc/conv_wgrad_row_tc.cu): the same
strips, tap-to-column mapping, TMEM slot ring with shadow slots, segment split and piece layout, written as numpy loops over the
"instructions" the kernels issue, checked against a direct SAME convolution / its gradients.  It runs without a GPU and pins the
conventions the CUDA code relies on (t = KS-1-ky, entry e = pixel e - PAD, slot = running output row mod 12 with windows running on
into 4 shadow slots, dY window rows r-PAD..r+PAD as M blocks, tap kx = strip shifted by kx entries as N blocks); the GPU parity tests
(tests/test_gpu_conv_tc.py, tests/test_gpu_kernels.py) check the kernels themselves."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

KRING, KPHYS = 12, 16          # conv_row_tc.cu: kRing, kPhysRing


def c24_hi(c):
  return c if c < 8 else 8 + c


def c24_lo(c):
  return 8 + c if c < 8 else 10 + c


def to_pieces(x):
  """fp32 [.., 10] -> 24-channel piece layout [hi0..7 | lo0..7 | hi8 hi9 lo8 lo9 1 0 0 0] (tc::kC24), as float64 of the fp16 values"""
  hi = x.astype(np.float16)
  lo = (x - hi.astype(np.float32)).astype(np.float16)
  out = np.zeros(x.shape[:-1] + (24,), dtype=np.float64)
  for c in range(10):
    out[..., c24_hi(c)] = hi[..., c]
    out[..., c24_lo(c)] = lo[..., c]
  out[..., 20] = 1.0
  return out


def weight_channel(ch):      # tc::c24_weight_channel
  return ch if ch < 8 else ch - 8 if ch < 16 else ch - 8 if ch < 18 else ch - 10 if ch < 20 else -1


def segments(n_tiles, HP, n_cta):
  """conv_row_tc.cu SegIter: pooled rows of all tiles, tile-major, one contiguous range per CTA, segments never cross a tile"""
  total = n_tiles * HP
  for cta in range(n_cta):
    g, g1 = total * cta // n_cta, total * (cta + 1) // n_cta
    segs = []
    while g < g1:
      tile = g // HP
      pa = g - tile * HP
      pb = min(HP, pa + (g1 - g))
      segs.append((tile, 2 * pa, 2 * pb))
      g += pb - pa
    yield segs


@pytest.mark.parametrize("B,H,W,KS,n_cta", [(5, 8, 8, 5, 3), (4, 6, 10, 3, 2), (7, 16, 16, 3, 5), (3, 32, 32, 5, 4), (2, 4, 124, 5, 1)])
def test_row_sweep_forward_algebra(B, H, W, KS, n_cta):
  rs = np.random.RandomState(B + H + KS)
  PAD, E = KS // 2, W + KS - 1
  ipt = 128 // E
  n_tiles = -(-B // ipt)
  x32 = np.maximum(rs.randn(B, H, W, 10), 0).astype(np.float32)
  xp = to_pieces(x32)                                           # what the layer below left behind
  w = rs.uniform(-0.3, 0.3, (KS, KS, 10, 10))
  # B operand: column block t <-> ky = KS-1-t; K8 half (kx, g) <-> channels 8g..8g+7 of the piece layout at tap column kx
  def bcol(t, kx, g):
    m = np.zeros((8, 10))
    for e in range(8):
      c = weight_channel(8 * g + e)
      if c >= 0:
        m[e] = w[KS - 1 - t, kx, c]
    return m
  out = np.zeros((B, H, W, 10))
  for segs in segments(n_tiles, H // 2, n_cta):                 # one CTA
    slots = np.zeros((KPHYS, 128, 10))                          # TMEM: ring + shadow slots, lanes, filters
    gseg = 0
    for tile, ya, yb in segs:
      ra, rb = max(0, ya - PAD), min(H - 1, yb - 1 + PAD)
      for r in range(ra, rb + 1):
        # the strip of input row r: entry e of image i = pixel e - PAD (zero halo = TMA out-of-bounds fill)
        strip = np.zeros((128 + KS, 24))
        for i in range(ipt):
          b = tile * ipt + i
          if b < B:
            strip[i * E + PAD:i * E + PAD + W] = xp[b, r]
        t0, t1 = max(0, ya - r + PAD), min(KS, yb - r + PAD)
        slot0 = (gseg + r - PAD + t0 - ya) % KRING              # windows never wrap: they run on into the shadow slots
        assert slot0 + (t1 - t0) <= KPHYS
        for kx in range(KS):
          for g in range(3):
            a = np.stack([strip[m + kx, 8 * g:8 * g + 8] for m in range(128)])          # lane m reads entry m + kx
            for t in range(t0, t1):
              slots[slot0 + (t - t0)] += a @ bcol(t, kx, g)
        # rows that received their last tap are drained in pairs; ring + shadow halves are added, both zeroed
        y_lo = max(ya, r - PAD) if r == rb else r - PAD
        y_hi = yb - 1 if r == rb else r - PAD
        for y in range(max(ya, y_lo), y_hi + 1):
          if y & 1:
            for yy in (y - 1, y):
              s = (gseg + yy - ya) % KRING
              acc = slots[s] + (slots[KRING + s] if s < KPHYS - KRING else 0.0)
              slots[s] = 0.0
              if s < KPHYS - KRING:
                slots[KRING + s] = 0.0
              for i in range(ipt):
                b = tile * ipt + i
                if b < B:
                  out[b, yy] = acc[i * E:i * E + W]
      gseg += yb - ya
    assert np.all(slots == 0.0)
  xe = torch.from_numpy(np.stack([xp[..., c24_hi(c)] + xp[..., c24_lo(c)] for c in range(10)], -1)).permute(0, 3, 1, 2)
  ref = F.conv2d(xe, torch.from_numpy(w).permute(3, 2, 0, 1), padding=PAD).permute(0, 2, 3, 1).numpy()
  assert np.abs(out - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("B,H,W,KS,n_cta", [(5, 8, 8, 5, 3), (4, 6, 10, 3, 2), (7, 16, 16, 3, 4), (3, 12, 20, 5, 5)])
def test_row_sweep_weight_gradient_algebra(B, H, W, KS, n_cta):
  """conv_wgrad_row_tc.cu: G[(t, n), (kx, c)] = sum over strip positions q of dY[y = r - PAD + t][q][n] * X[r][q + kx][c]; halo
  positions carry dY = 0; dW[ky][kx][c][o] = G summed over the hi / lo pieces of X (channels of the 24-layout); db from the
  constant-one channel at the centre tap"""
  rs = np.random.RandomState(B * 3 + H + KS)
  PAD, E = KS // 2, W + KS - 1
  ipt = 128 // E
  n_tiles = -(-B // ipt)
  x32 = np.maximum(rs.randn(B, H, W, 10), 0).astype(np.float32)
  xp = to_pieces(x32)
  dy = rs.randn(B, H, W, 10) * (rs.rand(B, H, W, 10) < 0.25)     # un-pooled gradient: one position of four is non-zero
  G = np.zeros((3, KS * 8, KS, 10))                              # [channel group][column kx * 8 + c8][window row t][n]
  total = n_tiles * H
  for cta in range(n_cta):                                       # Seq: input rows of all tiles, one contiguous range per CTA
    g, g1 = total * cta // n_cta, total * (cta + 1) // n_cta
    while g < g1:
      tile = g // H
      ra = g - tile * H
      rb = min(H, ra + (g1 - g))
      g += rb - ra
      for r in range(ra, rb):
        strip = np.zeros((ipt * E + KS, 24))
        for i in range(ipt):
          b = tile * ipt + i
          if b < B:
            strip[i * E + PAD:i * E + PAD + W] = xp[b, r]
        for t in range(KS):                                      # window row t = dY row r - PAD + t (zero outside the image)
          y = r - PAD + t
          if not (0 <= y < H):
            continue
          dyrow = np.zeros((ipt * E, 10))                        # strip position q = (image, column x); halo positions are zero
          for i in range(ipt):
            b = tile * ipt + i
            if b < B:
              dyrow[i * E:i * E + W] = dy[b, y]
          for gq in range(3):
            for kx in range(KS):                                 # N block kx = the strip shifted by kx entries (SBO = 16 bytes)
              xs = strip[kx:kx + ipt * E, 8 * gq:8 * gq + 8]
              G[gq, kx * 8:kx * 8 + 8, t] += xs.T @ dyrow
  dw = np.zeros((KS, KS, 10, 10))
  for ky in range(KS):
    t = KS - 1 - ky
    for kx in range(KS):
      for c in range(10):
        for ch in (c24_hi(c), c24_lo(c)):
          dw[ky, kx, c] += G[ch >> 3, kx * 8 + (ch & 7), t]
  db = G[20 >> 3, PAD * 8 + (20 & 7), PAD]
  xe = torch.from_numpy(np.stack([xp[..., c24_hi(c)] + xp[..., c24_lo(c)] for c in range(10)], -1)).permute(0, 3, 1, 2).requires_grad_(False)
  wt = torch.zeros(10, 10, KS, KS, dtype=torch.float64, requires_grad=True)
  bt = torch.zeros(10, dtype=torch.float64, requires_grad=True)
  yref = F.conv2d(xe, wt, bt, padding=PAD)
  gw, gb = torch.autograd.grad(yref, [wt, bt], grad_outputs=torch.from_numpy(dy).permute(0, 3, 1, 2))
  assert np.abs(dw - gw.permute(2, 3, 1, 0).numpy()).max() <= 1e-9 * max(1.0, float(gw.abs().max()))
  assert np.abs(db - gb.numpy()).max() <= 1e-9 * max(1.0, float(gb.abs().max()))


def test_segments_cover_every_pooled_row_once():
  for n_tiles, HP, n_cta in [(86, 16, 148), (86, 16, 37), (37, 8, 74), (1, 1, 5), (3, 7, 4)]:
    seen = np.zeros((n_tiles, HP), dtype=np.int64)
    for segs in segments(n_tiles, HP, n_cta):
      for tile, ya, yb in segs:
        assert ya % 2 == 0 and yb % 2 == 0 and 0 <= ya < yb <= 2 * HP
        seen[tile, ya // 2:yb // 2] += 1
    assert np.all(seen == 1)
