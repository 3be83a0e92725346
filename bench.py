#!/usr/bin/env python
"""bench.py - DDPG grad-steps/s on BASELINE config 3 (ddpg_cartpole.py pixel state 64x64, R=3, C=1 -> 9 channels,
conv actor/critic, batch 256 per GPU) through the B200-native hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            native arm (this repo's CUDA path)
  python bench.py --impl reference [...]                         the reference's CPU path (oracle port, fp32 torch)

One "step" = ReplayMemory.batch(256) + actor.train(batch.state_1) + critic.train(batch) (+ the tau target update
once per --batches-per-step=5 steps), ddpg_cartpole.py:331-337.  `value` keeps inputs in HBM (replay resident on the
GPU, host only draws the MT19937 indexes); `e2e` feeds host-resident (pinned) batches through the reference-facing
API with the H2D copies and the loss read-back inside the timed region.
Multi-GPU: data parallel, per-GPU batch fixed at 256 (weak scaling), one NCCL all-reduce per optimiser step;
value = N * K / time in batch-256 grad-steps/s."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (64, 64, 3, 1, 3)      # (H, W, rgb, cameras, repeats), bullet_cartpole.py:121-122
BATCH = 256
N_REPLAY = 4096
BATCHES_PER_STEP = 5
METRIC = "DDPG grad-steps/sec on 64x64 pixel batch=256"
UNIT = "grad-steps/s"
WORKLOAD = "c3: ddpg_cartpole.py --use-raw-pixels 64x64 R=3 C=1 (9ch) conv actor/critic, batch 256 per GPU"


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", type=str, default="native", choices=["native", "reference"])
  ap.add_argument("--skip-cpu-baseline", action="store_true")
  ap.add_argument("--skip-e2e", action="store_true")
  ap.add_argument("--skip-roofline", action="store_true")
  return ap.parse_args()


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    d = json.load(open(p))
    return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                source="measured")
  return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------ synthetic workload
def synthetic_tables(n):
  rs = np.random.RandomState(1234)
  action = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
  reward = np.ones((n, 1), dtype=np.float32)                           # bullet_cartpole.py:210
  mask = (rs.rand(n, 1) >= 1.0 / 50).astype(np.float32)
  return action, reward, mask


def fill_replay(rm):
  """fill the GPU replay memory directly (bypassing the env): slot k holds fp16(u8)/255 pixels"""
  import ctypes as C
  import torch
  from cartpoleplusplus_b200 import _lib
  g = torch.Generator(device=rm.device)
  g.manual_seed(1234)
  slots, row = rm.d_state.shape
  for s0 in range(0, slots, 512):
    k = torch.randint(0, 256, (min(512, slots - s0), row), device=rm.device, generator=g, dtype=torch.int32)
    rm.d_state[s0:s0 + k.shape[0]] = k.to(torch.float16) / torch.tensor(255, dtype=torch.float16, device=rm.device)
  n = rm.buffer_size
  action, reward, mask = synthetic_tables(n)
  rm.state_1_idx[:] = np.arange(n, dtype=np.int32)
  rm.state_2_idx[:] = np.arange(1, n + 1, dtype=np.int32)
  rm.action[:], rm.reward[:], rm.terminal_mask[:] = action, reward, mask
  rm.insert, rm.full = 0, True
  rm.state_free_slots.clear()
  rm.d_state_1_idx.copy_(torch.from_numpy(rm.state_1_idx)); rm.d_state_2_idx.copy_(torch.from_numpy(rm.state_2_idx))
  rm.d_action.copy_(torch.from_numpy(action)); rm.d_reward.copy_(torch.from_numpy(reward[:, 0])); rm.d_mask.copy_(torch.from_numpy(mask[:, 0]))
  all_slots = torch.arange(slots, dtype=torch.int32, device=rm.device)
  _lib.check(rm.lib.cpp_slot_stats(_lib.ptr(rm.d_state), _lib.ptr(all_slots), C.c_int32(slots), C.c_int64(rm.n_pix),
                                   C.c_int32(rm.channels), _lib.ptr(rm.d_slot_stats), _lib.stream_ptr()))
  torch.cuda.synchronize()


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
  Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

  def __init__(self, index):
    self.index, self.proc, self.lines = index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
    self.proc.terminate()
    self.t.join(timeout=2)
    sm, mx, reasons = [], None, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for l in self.lines:
      f = [x.strip() for x in l.split(",")]
      if len(f) < 7:
        continue
      try:
        sm.append(float(f[0])); mx = float(f[1])
      except ValueError:
        continue
      for nme, v in zip(names, f[3:7]):
        if v.lower().startswith("active"):
          reasons.add(nme)
    return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
  import ctypes as C
  import torch
  from cartpoleplusplus_b200 import _lib, base_network, ddpg_cartpole, dp as dpmod
  from cartpoleplusplus_b200.replay_memory import ReplayMemory, Batch

  dp = dpmod.DataParallel()
  assert dp.world_size == args.gpus or dp.world_size == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
  torch.cuda.set_device(dp.local_rank)
  dev = torch.device("cuda", dp.local_rank)
  lib = _lib.lib()

  o = ddpg_cartpole.set_opts(ddpg_cartpole.default_opts(
      ["--use-raw-pixels", "--render-width=64", "--render-height=64", "--action-repeats=3", "--num-cameras=1",
       "--batch-size=%d" % BATCH, "--replay-memory-size=%d" % N_REPLAY]))
  s1p, s2p = base_network.Placeholder(SHAPE, "s1"), base_network.Placeholder(SHAPE, "s2")
  actor = ddpg_cartpole.ActorNetwork("actor", s1p, 2)
  critic = ddpg_cartpole.CriticNetwork("critic", actor)
  tactor = ddpg_cartpole.ActorNetwork("target_actor", s2p, 2)
  tcritic = ddpg_cartpole.CriticNetwork("target_critic", tactor)
  actor.init_ops_for_training(critic)
  critic.init_ops_for_training(tcritic)
  eng = actor._engine
  rng = np.random.RandomState(42)                     # identical replicas: same xavier draw on every rank
  for net in (actor, critic, tactor, tcritic):
    net.flat_params().copy_(torch.from_numpy(net.initial_flat(rng)))
  tactor.set_as_target_network_for(actor, o.target_update_rate)
  tcritic.set_as_target_network_for(critic, o.target_update_rate)
  if dp.enabled:
    eng.set_data_parallel(dp)
    eng.max_batch = 0; eng._ensure(BATCH)              # re-create with world_size

  rm = ReplayMemory(N_REPLAY, SHAPE, 2)
  fill_replay(rm)
  rm.sampler.seed(0)                                   # np.random.seed(0) stream, identical on every rank

  G = dp.world_size
  # Two sets of batch buffers: the replay gather of step i+1 (a1, a2 and the whitening statistics) is enqueued on a second
  # stream right after the training step i was launched and runs next to it; every step still does exactly one gather and
  # one training step, in index-stream order.
  gs = torch.cuda.Stream(device=dev)
  idx_pin = [torch.empty(BATCH * G, dtype=torch.int64, pin_memory=True) for _ in range(2)]
  d_idx = [torch.empty(BATCH * G, dtype=torch.int64, device=dev) for _ in range(2)]
  outs = [Batch(torch.empty((BATCH,) + SHAPE, dtype=torch.float16, device=dev), torch.empty((BATCH, 2), device=dev),
                torch.empty((BATCH, 1), device=dev), torch.empty((BATCH, 1), device=dev),
                torch.empty((BATCH,) + SHAPE, dtype=torch.float16, device=dev)) for _ in range(2)]
  moms = [(torch.empty(18, dtype=torch.float32, device=dev), torch.empty(18, dtype=torch.float32, device=dev)) for _ in range(2)]
  copied = [torch.cuda.Event() for _ in range(2)]      # the index vector left its pinned buffer
  ready = [torch.cuda.Event() for _ in range(2)]       # gather + statistics of this buffer set are complete
  free = [torch.cuda.Event() for _ in range(2)]        # the training step that read this buffer set is complete
  pending = {}

  def issue_gather(i):
    k = i % 2
    idxs = rm.random_indexes(BATCH * G)                                    # a1: host MT19937, bit exact
    if i >= 2:
      copied[k].synchronize()
    idx_pin[k].copy_(torch.from_numpy(idxs))
    with torch.cuda.stream(gs):
      if i >= 2:
        gs.wait_event(free[k])
      d_idx[k].copy_(idx_pin[k], non_blocking=True)
      copied[k].record(gs)
      mine = d_idx[k][dp.rank * BATCH:(dp.rank + 1) * BATCH]
      batch = rm.batch_at(idxs[dp.rank * BATCH:(dp.rank + 1) * BATCH], d_idxs=mine, out=outs[k])   # a2: gather kernel
      rm.batch_moments(d_idx[k], 1, out=moms[k][0]); rm.batch_moments(d_idx[k], 2, out=moms[k][1])  # global-batch whitening statistics
      ready[k].record(gs)
    pending[i] = batch

  def step(i):
    k = i % 2
    if i not in pending:
      issue_gather(i)
    batch = pending.pop(i)
    torch.cuda.current_stream().wait_event(ready[k])
    eng.train_step(batch, moments=moms[k])                                 # a3-a12
    free[k].record()
    issue_gather(i + 1)
    if (i + 1) % BATCHES_PER_STEP == 0:
      eng.update_targets()                                                 # a13

  def timed(fn, steps, warmup, clocks=None, drain=None):
    for i in range(warmup):
      fn(i)
    if clocks:
      clocks.start(); time.sleep(0.25)       # before the barrier: a rank that sleeps after it stalls its peers' first collective
    dp.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.cpp_launch_count()
    e0.record()
    for i in range(steps):
      fn(warmup + i)
    if drain is not None:
      drain()                               # work the steps put on other streams belongs to the timed region
    e1.record()
    torch.cuda.synchronize(); dp.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dp.all_reduce_max(ms)
    return float(ms.item()), lib.cpp_launch_count() - l0, (clocks.stop() if clocks else None)

  W = max(3, args.warmup)
  if os.environ.get("BENCH_HOST_PROFILE"):
    # diagnosis, not a bench value: is the loop host bound?  host clock before / after the final synchronize + cProfile
    import cProfile, pstats
    for i in range(10):
      step(i)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for i in range(n):
      step(10 + i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    sys.stderr.write("enqueue %.1f us/step, until the GPU is done %.1f us/step\n" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
    pr = cProfile.Profile(); pr.enable()
    for i in range(n):
      step(1000 + i)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(25)
    pending.clear()
  clocks = ClockSampler(dp.local_rank) if dp.rank == 0 else None
  ms, launches, clk = timed(step, args.steps, W, clocks, drain=lambda: torch.cuda.current_stream().wait_stream(gs))
  value = G * args.steps / (ms / 1e3)

  # ---- e2e: host (pinned) batches through the reference-facing API, H2D + loss D2H inside the timed region
  e2e = None
  if not args.skip_e2e:
    host = []
    for j in range(4):
      idxs = rm.random_indexes(BATCH)
      b = rm.batch_at(idxs)
      host.append(Batch(*[t.cpu().pin_memory() for t in b]))
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    losses = []

    def e2e_step(i):
      b = host[i % len(host)]
      eng.train_step(b)                      # == actor.train(b.state_1); critic.train(b): H2D of this batch unless prefetched
      eng.prefetch(host[(i + 1) % len(host)])   # the next batch's H2D copy overlaps this step's kernels (copy stream)
      if (i + 1) % BATCHES_PER_STEP == 0:
        eng.update_targets()
      losses.append(eng.last_loss())         # D2H read of the step's loss (syncs, like Session.run returning)

    ems, _, _ = timed(e2e_step, args.steps, W)
    e2e = dict(value=G * args.steps / (ems / 1e3), unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=4,
               ms_per_step=ems / args.steps, h2d_gb_per_s=h2d * args.steps / (ems / 1e3) / 1e9,
               note="host->device bound: the fp16 states of one batch are 37.75 MB, i.e. this rate IS the pinned-memory PCIe copy rate",
               api="DDPGEngine.train_step(Batch of pinned host tensors) + prefetch(next batch) + last_loss(); every batch is "
                   "copied host->device inside the timed region, overlapped with the previous step")

  # ---- roofline of the dominant tensor-core kernel: conv1 forward of actor+critic on state_1 in one tcgen05 pass
  # (5x5, 9 -> 2x10 ch, 64x64, B=256), timed alone with CUDA events on the launching stream, L2 flushed in between
  roof = None
  if dp.rank == 0 and not args.skip_roofline:
    peaks = measured_peaks()
    x = outs[0].state_1
    m1 = moms[0][0]
    ws_ = [actor.get_variable("actor/conv1/weights"), critic.get_variable("critic/conv1/weights")]
    bs_ = [actor.get_variable("actor/conv1/biases"), critic.get_variable("critic/conv1/biases")]
    pooled = [torch.empty((BATCH, 32, 32, 10), dtype=torch.float32, device=dev) for _ in range(2)]
    amax = [torch.empty((BATCH, 32, 32, 10), dtype=torch.uint8, device=dev) for _ in range(2)]
    scr = torch.zeros(int(lib.cpp_conv_tc_scratch_bytes(2, 64, 64, 9, 5)), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    wp, bp, pp, ap = _lib.ptr_array(ws_), _lib.ptr_array(bs_), _lib.ptr_array(pooled), _lib.ptr_array(amax)
    def conv1():
      _lib.check(lib.cpp_conv_forward_tc(_lib.ptr(x), None, _lib.ptr(m1), 2, wp, bp, BATCH, 64, 64, 9, 5, pp, ap, _lib.ptr(scr),
                                         _lib.stream_ptr(), 0, None))
    for _ in range(3):
      conv1()
    ts = []
    for _ in range(10):
      flush.fill_(1)                         # L2 flush between timed launches
      a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); conv1(); b_.record(); torch.cuda.synchronize()
      ts.append(a.elapsed_time(b_))
    kms = float(np.mean(ts))
    flops = 2.0 * 2 * BATCH * 64 * 64 * 10 * 25 * 9         # conv1 MACs/sample = H*W*10*25*Cin, two sibling nets
    ach = flops / (kms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):                       # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
      tj = json.load(open(tp))
      traffic, traffic_src = int(tj["dram_read_bytes"]) + int(tj["dram_write_bytes"]), tj["source"]
    roof = dict(kernel="conv_fwd_tc_kernel<5,1> + weight prep (conv1 5x5 fwd of actor+critic, whitening fold, bias, ReLU, 2x2 maxpool; "
                       "tcgen05.mma kind::f16, fp32 weights as 2 fp16 pieces)", bound="tensor",
                achieved=ach, peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=ach / peaks["bf16_tflops"], traffic=traffic,
                traffic_unit="bytes of DRAM traffic per launch", traffic_source=traffic_src,
                peak_source=peaks["source"] + " bf16 dense (burst)", kernel_ms=kms, algorithmic_flops_per_launch=flops,
                algorithmic_bytes_per_launch=BATCH * 64 * 64 * 9 * 2,
                note="algorithmic (useful) FLOPs: N = 20 real filters of 48 issued columns, K = 225 of 256 issued; the kernel is bound "
                     "by the shared-memory read of the A operand (4 KB per tcgen05.mma), not by the math rate")

  cpu = None
  if dp.rank == 0 and not args.skip_cpu_baseline:
    cpu = cpu_reference(steps=3, warmup=1)

  if dp.rank == 0:
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=G, steps=args.steps, warmup=W, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (conv MMAs: fp16 hi+lo operand pieces, fp32 accumulate; FC / elementwise fp32)", data="synthetic",
                config=dict(workload=WORKLOAD, global_batch=BATCH * G, replay=N_REPLAY, batches_per_step=BATCHES_PER_STEP,
                            parallelism="dp%d" % G,
                            l2="inputs larger than L2: each step gathers 37.7 MB of random rows from a 453 MB fp16 replay slab",
                            pipeline="the replay gather of step i+1 runs on a second stream next to training step i (two batch buffer sets); "
                                     "each timed step = one gather + one training step"),
                clocks=clk, e2e=e2e, gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line))
  dp.close()


# ------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference(steps, warmup, budget_s=120.0):
  """the reference's CPU path on this box's host cores: py3 restatement of ReplayMemory.batch + fp16->fp32 feed cast
  + torch-CPU fp32 restatement of the TF graph (oracle/, kind 'port'; TensorFlow itself cannot run here)"""
  import torch
  from oracle import nets_oracle as no
  from oracle.replay_oracle import ReplayOracle
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  rs = np.random.RandomState(42)
  P = {}
  for d in (no.ddpg_actor("actor", SHAPE, True), no.ddpg_critic("critic", SHAPE, True)):
    P.update(no.init_params(d, rs, torch.float32))
  P.update(no.retarget({k: v for k, v in P.items() if k.startswith("actor/")}, "actor", "target_actor"))
  P.update(no.retarget({k: v for k, v in P.items() if k.startswith("critic/")}, "critic", "target_critic"))
  orc = no.DDPGOracle(SHAPE, True, P)
  n = 1024                                              # bounded replay for the CPU arm (sampling cost is per row)
  rm = ReplayOracle(n, SHAPE, 2, rng=np.random.RandomState(0))
  r2 = np.random.RandomState(1234)
  rm.state[:] = (r2.randint(0, 256, rm.state.shape).astype(np.float16) / np.float16(255))
  rm.state_1_idx[:] = np.arange(n); rm.state_2_idx[:] = np.arange(1, n + 1)
  rm.action[:], rm.reward[:], rm.terminal_mask[:] = synthetic_tables(n)
  rm.insert, rm.full = 0, True

  def one(batch_size, i):
    b = rm.batch(batch_size)
    feed = (b.state_1.astype(np.float32), b.action, b.reward, b.terminal_mask, b.state_2.astype(np.float32))   # feed_dict cast
    orc.actor_train(feed[0])
    orc.critic_train(feed)
    if (i + 1) % BATCHES_PER_STEP == 0:
      orc.update_targets()

  t0 = time.time(); one(BATCH, 0); t1 = time.time() - t0
  bs = BATCH
  if t1 * (steps + warmup) > budget_s:                   # keep the arm bounded: shrink the per-step sample
    bs = max(8, int(BATCH * budget_s / (t1 * (steps + warmup))))
  for i in range(warmup):
    one(bs, i)
  ts = []
  for i in range(steps):
    t0 = time.time(); one(bs, warmup + i); ts.append(time.time() - t0)
  per_step = float(np.median(ts)) * (BATCH / bs)         # batch-256 equivalent
  return dict(value=1.0 / per_step, unit=UNIT, cores=cores, kind="port",
              sample="%d timed grad-steps at batch %d (median, scaled to batch 256) incl. replay gather + fp16->fp32 feed cast; "
                     "torch-CPU fp32 oracle port with %d threads" % (steps, bs, cores),
              ms_per_step=per_step * 1e3)


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  W = max(1, min(args.warmup, 3))
  cpu = cpu_reference(steps=args.steps, warmup=W)
  line = dict(impl="reference", metric=METRIC, value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=W,
              ms_per_step=cpu["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
              config=dict(workload=WORKLOAD, global_batch=BATCH, replay=1024, batches_per_step=BATCHES_PER_STEP, parallelism="cpu"),
              cpu_baseline=dict(kind=cpu["kind"], cores=cpu["cores"], sample=cpu["sample"], value=cpu["value"], unit=UNIT),
              e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
  print(json.dumps(line))


if __name__ == "__main__":
  a = parse_args()
  if a.impl == "reference":
    run_reference(a)
  else:
    run_native(a)
