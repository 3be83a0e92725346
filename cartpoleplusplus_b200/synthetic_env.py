"""Stand-in for the reference's BulletCartpole (/root/reference/bullet_cartpole.py), which needs pybullet
and gym (both absent, and the physics is out of scope: SURVEY.md section 2 row 7).  It keeps the reference's
flag names/defaults (:13-38), observation/action space shapes (:117-130) and the gym-style reset/step
protocol so the agent loops and command lines run end to end on synthetic data.  The dynamics are a toy
pole-angle integrator, NOT the reference's physics."""
import numpy as np


def add_opts(parser):
  parser.add_argument('--gui', action='store_true')
  parser.add_argument('--delay', type=float, default=0.0)
  parser.add_argument('--action-force', type=float, default=50.0)
  parser.add_argument('--initial-force', type=float, default=55.0)
  parser.add_argument('--no-random-theta', action='store_true')
  parser.add_argument('--action-repeats', type=int, default=2)
  parser.add_argument('--steps-per-repeat', type=int, default=5)
  parser.add_argument('--num-cameras', type=int, default=1)
  parser.add_argument('--event-log-out', type=str, default=None)
  parser.add_argument('--max-episode-len', type=int, default=200)
  parser.add_argument('--use-raw-pixels', action='store_true')
  parser.add_argument('--render-width', type=int, default=50)
  parser.add_argument('--render-height', type=int, default=50)
  parser.add_argument('--reward-calc', type=str, default='fixed')


class _Box(object):
  def __init__(self, shape):
    self.shape = tuple(shape)


class _Discrete(object):
  def __init__(self, n):
    self.n = n

  def sample(self):
    return np.random.randint(0, self.n)


class SyntheticCartpole(object):
  def __init__(self, opts, discrete_actions):
    self.repeats = opts.action_repeats
    if opts.num_cameras not in [1, 2]:
      raise ValueError("--num-cameras must be 1 or 2")
    self.num_cameras = opts.num_cameras
    self.use_raw_pixels = opts.use_raw_pixels
    self.render_width, self.render_height = opts.render_width, opts.render_height
    self.max_episode_len = opts.max_episode_len
    self.discrete_actions = discrete_actions
    if self.use_raw_pixels:
      state_shape = (self.render_height, self.render_width, 3, self.num_cameras, self.repeats)
    else:
      state_shape = (self.repeats, 2, 7)
    self.observation_space = _Box(state_shape)
    self.action_space = _Discrete(5) if discrete_actions else _Box((1, 2))
    assert opts.reward_calc in ['fixed', 'angle', 'action', 'angle_action']
    self.state = np.empty(state_shape, dtype=np.float32)
    self.rs = np.random.RandomState(0)
    self.steps = 0
    self.done = True
    # --event-log-out (bullet_cartpole.py:90-94,221-222,283-285): every episode is appended in the reference's framed
    # protobuf + PNG format, which ReplayMemory.reset_from_event_log (--event-log-in) reads back
    self.event_log = None
    if getattr(opts, "event_log_out", None):
      from . import event_log
      self.event_log = event_log.EventLog(opts.event_log_out, opts.use_raw_pixels)

  def _render(self, cam):
    H, W = self.render_height, self.render_width
    img = np.full((H, W, 3), 200, dtype=np.float16)
    cx = int(np.clip((self.pos[cam % 2] + 1.0) * 0.5 * (W - 1), 0, W - 1))
    top = int(np.clip((1.0 - abs(self.theta[cam % 2])) * (H // 2), 0, H - 1))
    img[H - 4:, max(0, cx - 3):cx + 4] = (30, 30, 220)
    img[top:H - 4, cx] = (220, 30, 30)
    img /= 255           # fp16(k)/255 in fp16, bullet_cartpole.py:239-242
    return img

  def _fill(self, repeat):
    if self.use_raw_pixels:
      for cam in range(self.num_cameras):
        self.state[:, :, :, cam, repeat] = self._render(cam)
    else:
      self.state[repeat][0] = (self.pos[0], self.pos[1], 0.1, 0, 0, 0, 1)
      self.state[repeat][1] = (self.pos[0], self.pos[1], 0.4, self.theta[0], self.theta[1], 0, 1)

  def reset(self):
    self.steps, self.done = 0, False
    self.pos = np.zeros(2)
    self.vel = np.zeros(2)
    self.theta = self.rs.uniform(-0.05, 0.05, 2)
    self.omega = self.rs.uniform(-0.3, 0.3, 2)
    for r in range(self.repeats):
      self._fill(r)
    if self.event_log:
      self.event_log.reset()
      self.event_log.add_just_state(self.state)
    return np.copy(self.state)

  def step(self, action):
    if self.done:
      raise RuntimeError("step() on a finished episode")
    if self.discrete_actions:
      f = {0: (0, 0), 1: (-1, 0), 2: (1, 0), 3: (0, -1), 4: (0, 1)}[int(action)]
      f = np.array(f, dtype=np.float64)
    else:
      f = np.asarray(action, dtype=np.float64).reshape(-1)[:2]
    for r in range(self.repeats):
      self.vel += 0.02 * f
      self.pos += 0.05 * self.vel
      self.omega += 0.05 * (3.0 * self.theta - 0.8 * f)
      self.theta += 0.05 * self.omega
      self._fill(r)
    self.steps += 1
    if np.any(np.abs(self.theta) > 0.35) or np.any(np.abs(self.pos) > 1.0) or self.steps >= self.max_episode_len:
      self.done = True
    if self.event_log:
      self.event_log.add(self.state, int(action) if self.discrete_actions else np.asarray(action, dtype=np.float32).reshape(1, -1), 1.0)
    return np.copy(self.state), 1.0, self.done, {}
