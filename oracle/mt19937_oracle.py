"""Oracle for replay index sampling.

The reference draws minibatch indexes with the legacy global numpy RNG:
``np.random.randint(0, size, n)`` (/root/reference/replay_memory.py:123-129).  That
arithmetic lives in numpy (third party, unpinned by the reference; numpy 2.3.5 here, the
legacy stream is frozen by NEP 19).  This file restates the published algorithm:

  * ``init_genrand(seed)``   - MT19937 seeding used by ``np.random.seed(int)``
  * ``genrand_int32``        - the MT19937 tempering/twist
  * legacy ``randint(0, N, n)`` for the default int dtype: ``rng = N-1``; ``mask`` = the
    smallest 2^k-1 >= rng; per output draw one 32-bit word (rng <= 0xFFFFFFFF), keep
    ``word & mask`` when it is <= rng, otherwise redraw (masked rejection).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned by tests/golden/mt19937_kat.json
(values produced by numpy itself, see oracle/make_golden.py) and by SURVEY.md Appendix F.
"""
import numpy as np

N, M = 624, 397
MATRIX_A, UPPER, LOWER = 0x9908B0DF, 0x80000000, 0x7FFFFFFF


class MT19937:
  def __init__(self, seed=None):
    self.mt = [0] * N
    self.pos = N
    if seed is not None:
      self.seed(seed)

  def seed(self, s):
    # init_genrand: mt[i] = 1812433253 * (mt[i-1] ^ (mt[i-1] >> 30)) + i
    s &= 0xFFFFFFFF
    mt = self.mt
    mt[0] = s
    for i in range(1, N):
      mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
    self.pos = N

  # numpy state tuple: ('MT19937', key[624] uint32, pos, has_gauss, cached_gaussian)
  def set_numpy_state(self, state):
    self.mt = [int(v) for v in state[1]]
    self.pos = int(state[2])

  def get_numpy_state(self, has_gauss=0, cached=0.0):
    return ('MT19937', np.array(self.mt, dtype=np.uint32), self.pos, has_gauss, cached)

  def _twist(self):
    mt = self.mt
    for k in range(N):
      y = (mt[k] & UPPER) | (mt[(k + 1) % N] & LOWER)
      mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    self.pos = 0

  def next_u32(self):
    if self.pos >= N:
      self._twist()
    y = self.mt[self.pos]
    self.pos += 1
    y ^= y >> 11
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= y >> 18
    return y & 0xFFFFFFFF

  def randint(self, high, n):
    """legacy np.random.randint(0, high, n) -> int64[n]"""
    rng = high - 1
    out = np.empty(n, dtype=np.int64)
    if rng == 0:
      out[:] = 0      # numpy draws nothing from the stream in this case
      return out
    assert 0 < rng <= 0xFFFFFFFF
    mask = rng
    for sh in (1, 2, 4, 8, 16):
      mask |= mask >> sh
    for i in range(n):
      while True:
        v = self.next_u32() & mask
        if v <= rng:
          break
      out[i] = v
    return out


def numpy_legacy_randint(seed, high, n, calls=1):
  """The dependency itself (numpy RandomState) - what the reference actually executes."""
  rs = np.random.RandomState(seed)
  return [rs.randint(0, high, n) for _ in range(calls)]
