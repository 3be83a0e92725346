"""Reader / writer of the reference's on-disk event log (/root/reference/event_log.py:21-111, event.proto) - SURVEY.md 8f
row 3: `--event-log-in --dont-do-rollouts` is how the reference drives pure-training runs (exps/run_81.sh:6).

Format (event_log.py:50-58,103-111): a sequence of episodes, each framed as struct '=l' (native int32) byte length
followed by a serialised `cp.Episode` protobuf (event.proto:1-35); the file may be gzip-compressed (".gz").
  Episode{ repeated Event event = 1 }       Event{ repeated float action = 1; repeated State state = 2; optional float reward = 3 }
  State{ repeated float cart_pose = 1; repeated float pole_pose = 2; repeated Render render = 3 }
  Render{ optional int32 height = 1; optional int32 width = 2; optional bytes png_bytes = 3 }
The first event of an episode carries only the reset state; pixel states are one PNG per (repeat, camera).

`protoc` is not available, so the message classes are built from a hand-written descriptor; PNGs go through PIL
(the reference uses matplotlib's imsave/imread: RGBA PNG, float RGB in [0,1] = uint8 / 255)."""
import gzip
import io
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory


def _build_messages():
  fd = descriptor_pb2.FileDescriptorProto()
  fd.name = "cartpolepp_event.proto"
  fd.package = "cp"
  fd.syntax = "proto2"
  F = descriptor_pb2.FieldDescriptorProto

  def msg(name, fields):
    m = fd.message_type.add()
    m.name = name
    for fname, number, ftype, label, type_name in fields:
      f = m.field.add()
      f.name, f.number, f.type, f.label = fname, number, ftype, label
      if type_name:
        f.type_name = type_name

  OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED
  msg("Render", [("height", 1, F.TYPE_INT32, OPT, None), ("width", 2, F.TYPE_INT32, OPT, None),
                 ("png_bytes", 3, F.TYPE_BYTES, OPT, None)])
  msg("State", [("cart_pose", 1, F.TYPE_FLOAT, REP, None), ("pole_pose", 2, F.TYPE_FLOAT, REP, None),
                ("render", 3, F.TYPE_MESSAGE, REP, ".cp.Render")])
  msg("Event", [("action", 1, F.TYPE_FLOAT, REP, None), ("state", 2, F.TYPE_MESSAGE, REP, ".cp.State"),
                ("reward", 3, F.TYPE_FLOAT, OPT, None)])
  msg("Episode", [("event", 1, F.TYPE_MESSAGE, REP, ".cp.Event")])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  get = getattr(message_factory, "GetMessageClass", None)
  if get is None:                                   # older protobuf
    factory = message_factory.MessageFactory(pool)
    get = factory.GetPrototype
  return {n: get(pool.FindMessageTypeByName("cp." + n)) for n in ("Render", "State", "Event", "Episode")}


_M = _build_messages()
Render, State, Event, Episode = _M["Render"], _M["State"], _M["Event"], _M["Episode"]


def rgb_to_png(rgb):
  """event_log.py:9-13 (plt.imsave of a float RGB image): 8-bit RGBA PNG, channel = uint8(x * 255) with alpha 255"""
  from PIL import Image
  a = np.asarray(rgb, dtype=np.float64)
  u8 = (np.clip(a, 0.0, 1.0) * 255).astype(np.uint8)                 # matplotlib's to_rgba(bytes=True) truncates
  rgba = np.concatenate([u8, np.full(u8.shape[:2] + (1,), 255, np.uint8)], axis=2)
  sio = io.BytesIO()
  Image.fromarray(rgba, "RGBA").save(sio, format="PNG")
  return sio.getvalue()


def png_to_rgb(png_bytes):
  """event_log.py:15-19 (plt.imread): float32 RGB in [0,1] = uint8 / 255, alpha dropped"""
  from PIL import Image
  im = Image.open(io.BytesIO(png_bytes)).convert("RGBA")
  return (np.asarray(im, dtype=np.float32) / np.float32(255))[:, :, :3]


def read_state_from_event(event):
  """event_log.py:21-39: (H, W, 3, cameras, repeats) for pixel events, (repeats, 2, 7) for low-dim ones"""
  if len(event.state[0].render) > 0:
    num_repeats, num_cameras = len(event.state), len(event.state[0].render)
    eg = event.state[0].render[0]
    state = np.empty((eg.height, eg.width, 3, num_cameras, num_repeats))
    for r_idx in range(num_repeats):
      for c_idx in range(num_cameras):
        state[:, :, :, c_idx, r_idx] = png_to_rgb(event.state[r_idx].render[c_idx].png_bytes)
  else:
    state = np.empty((len(event.state), 2, 7))
    for i, s in enumerate(event.state):
      state[i][0] = s.cart_pose
      state[i][1] = s.pole_pose
  return state


class EventLog(object):
  """event_log.py:41-92: append-only writer, one framed Episode per reset()"""

  def __init__(self, path, use_raw_pixels):
    self.log_file = open(path, "ab")
    self.episode_entry = None
    self.use_raw_pixels = use_raw_pixels

  def reset(self):
    if self.episode_entry is not None:
      buff = self.episode_entry.SerializeToString()
      if len(buff) > 0:
        self.log_file.write(struct.pack('=l', len(buff)))
        self.log_file.write(buff)
        self.log_file.flush()
    self.episode_entry = Episode()

  def add_state_to_event(self, state, event):
    state = np.asarray(state)
    if self.use_raw_pixels:
      for r_idx in range(state.shape[4]):
        s = event.state.add()
        for c_idx in range(state.shape[3]):
          render = s.render.add()
          render.width, render.height = state.shape[1], state.shape[0]
          render.png_bytes = rgb_to_png(state[:, :, :, c_idx, r_idx])
    else:
      for r in range(state.shape[0]):
        s = event.state.add()
        s.cart_pose.extend([float(v) for v in state[r][0]])
        s.pole_pose.extend([float(v) for v in state[r][1]])

  def add(self, state, action, reward):
    event = self.episode_entry.event.add()
    self.add_state_to_event(state, event)
    if isinstance(action, int):
      event.action.append(action)
    else:
      action = np.asarray(action)
      assert action.shape[0] == 1            # never log batch operations
      event.action.extend([float(v) for v in action[0]])
    event.reward = reward

  def add_just_state(self, state):
    self.add_state_to_event(state, self.episode_entry.event.add())

  def close(self):
    self.reset()
    self.log_file.close()


class EventLogReader(object):
  """event_log.py:95-111"""

  def __init__(self, path):
    self.log_file = gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")

  def entries(self):
    while True:
      buff_len_bytes = self.log_file.read(4)
      if len(buff_len_bytes) == 0:
        return
      buff_len = struct.unpack('=l', buff_len_bytes)[0]
      episode = Episode()
      episode.ParseFromString(self.log_file.read(buff_len))
      yield episode
