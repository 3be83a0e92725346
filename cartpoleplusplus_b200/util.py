"""Mirror of the reference's util.py for the pieces the hot path and its callers need
(/root/reference/util.py): flags :10-20, StopWatch :22-26, collapsed_successive_ranges :60-71,
construct_optimiser :73-76 (returns the optimiser *description* the CUDA step consumes),
OrnsteinUhlenbeckNoise :134-156 (host side, including the rotated np.clip arguments, Appendix C-6).
SaverUtil / PNG helpers are out of scope."""
import json
import time
import numpy as np

from ._engine import parse_optimiser


def add_opts(parser):
  parser.add_argument('--gradient-clip', type=float, default=5, help="do global clipping to this norm")
  parser.add_argument('--print-gradients', action='store_true', help="whether to verbose print all gradients and l2 norms")
  parser.add_argument('--optimiser', type=str, default="GradientDescent", help="tf.train.XXXOptimizer to use")
  parser.add_argument('--optimiser-args', type=str, default="{\"learning_rate\": 0.001}",
                      help="json serialised args for optimiser constructor")
  parser.add_argument('--use-dropout', action='store_true', help="include a dropout layers after each fully connected layer")


class StopWatch:
  def reset(self):
    self.start = time.time()

  def time(self):
    return time.time() - self.start


def collapsed_successive_ranges(values):
  """reduce an array, e.g. [2,3,4,5,13,14,15], to its successive ranges [2-5, 13-15]"""
  last, start, out = None, None, []
  for value in values:
    if start is None:
      start = value
    elif value != last + 1:
      out.append("%d-%d" % (start, last))
      start = value
    last = value
  out.append("%d-%d" % (start, last))
  return ", ".join(out)


def construct_optimiser(opts):
  """-> (kind, hyper-parameter dict) for libcartpolepp's optimiser kernels"""
  return parse_optimiser(opts.optimiser, json.loads(opts.optimiser_args))


def shape_and_product_of(shape):
  n = 1
  for d in shape:
    if d is not None:
      n *= int(d)
  return "%s #%s" % (tuple(shape), n)


class OrnsteinUhlenbeckNoise(object):
  """generate time correlated noise for action exploration"""

  def __init__(self, dim, theta=0.01, sigma=0.2, max_magnitude=1.5):
    self.dim, self.theta, self.sigma, self.max_magnitude = dim, theta, sigma, max_magnitude
    self.state = np.zeros(self.dim)

  def sample(self):
    self.state += self.theta * -self.state
    self.state += self.sigma * np.random.randn(self.dim)
    # util.py:155 passes np.clip(max, -max, state): the arguments are rotated in the reference, which
    # evaluates to minimum(state, max) for these values; reproduced as is
    self.state = np.clip(self.max_magnitude, -self.max_magnitude, self.state)
    return np.copy(self.state)
