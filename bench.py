#!/usr/bin/env python
"""bench.py - grad-steps/s of the cartpole++ training inner loop through the B200-native hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            native arm, BASELINE config 3 (the headline metric)
  python bench.py --config c4|c5 [--scaling weak|strong] ...      the other BASELINE configs (NAF B=512 / DDPG 128x128 B=1024)
  python bench.py --impl reference [...]                         the reference's CPU path (oracle port, fp32 torch)

One "step" = ReplayMemory.batch(B) + actor.train(batch.state_1) + critic.train(batch) (DDPG, ddpg_cartpole.py:331-337) or
+ naf.train(batch) (NAF, naf_cartpole.py:367-373), + the tau target update once per --batches-per-step=5 steps.
`value` keeps inputs in HBM (replay resident on the GPU, host only draws the MT19937 indexes); `e2e` feeds host-resident
(pinned) batches through the reference-facing API with the H2D copies and the loss read-back inside the timed region.

Multi-GPU: data parallel, one process per GPU (torchrun), the gradient all-reduce inside the step (csrc/comm.cu).
  --scaling weak   : per-GPU batch fixed at the config's batch, global batch = N x that; value = N * K / time in
                     config-batch grad-steps/s (default for c3: "batch=256 at 1/2/4/8 B200")
  --scaling strong : the config's GLOBAL batch split over the N GPUs (c4: 512 over 4 -> 128 per GPU, c5: 1024 over 8 -> 128
                     per GPU, as BASELINE.json configures them; default for c4 / c5); value = K / time."""
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

N_REPLAY = 4096
BATCHES_PER_STEP = 5
UNIT = "grad-steps/s"
# MAC / sample / grad-step of the minimum necessary graph
CONFIGS = {
    "c3": dict(agent="ddpg", shape=(64, 64, 3, 1, 3), batch=256, scaling="weak", mac_per_sample=78503500,
               metric="DDPG grad-steps/sec on 64x64 pixel batch=256",
               workload="c3: ddpg_cartpole.py --use-raw-pixels 64x64 R=3 C=1 (9ch) conv actor/critic, batch 256"),
    "c4": dict(agent="naf", shape=(64, 64, 3, 2, 3), batch=512, scaling="strong", mac_per_sample=157618950,
               metric="NAF grad-steps/sec on 64x64 pixel (18ch) batch=512",
               workload="c4: naf_cartpole.py --use-raw-pixels 64x64 R=3 C=2 (18ch) V/mu/L.L^T head, batch 512 (4 GPUs: 128 per GPU)"),
    "c5": dict(agent="ddpg", shape=(128, 128, 3, 2, 4), batch=1024, scaling="strong", mac_per_sample=682305100,
               metric="DDPG grad-steps/sec on 128x128 pixel (24ch) batch=1024",
               workload="c5: ddpg_cartpole.py --use-raw-pixels 128x128 R=4 C=2 (24ch) conv actor/critic, batch 1024 (8 GPUs: 128 per GPU)"),
}


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", type=str, default="native", choices=["native", "reference"])
  ap.add_argument("--config", type=str, default="c3", choices=sorted(CONFIGS))
  ap.add_argument("--scaling", type=str, default=None, choices=["weak", "strong"])
  ap.add_argument("--sustained-seconds", type=float, default=1.5, help="extra loop of at least this long after the K timed steps (0: off)")
  ap.add_argument("--skip-cpu-baseline", action="store_true")
  ap.add_argument("--skip-e2e", action="store_true")
  ap.add_argument("--skip-roofline", action="store_true")
  ap.add_argument("--host-allreduce", action="store_true", help="A/B: the round-1 all-reduce issued from the host between two C-ABI calls")
  ap.add_argument("--no-allreduce", action="store_true", help="diagnosis only (replicas diverge): N independent replicas, what N processes cost without any exchange")
  return ap.parse_args()


def plan(args):
  """-> (config dict, per-GPU batch, global batch, scaling)"""
  cfg = CONFIGS[args.config]
  scaling = args.scaling or cfg["scaling"]
  G = args.gpus
  if scaling == "weak":
    per, glob = cfg["batch"], cfg["batch"] * G
  else:
    assert cfg["batch"] % G == 0, "global batch %d does not split over %d GPUs" % (cfg["batch"], G)
    per, glob = cfg["batch"] // G, cfg["batch"]
  return cfg, per, glob, scaling


def config_dict(args):
  """the workload description - IDENTICAL for the native and the reference arm"""
  cfg, per, glob, scaling = plan(args)
  return dict(workload=cfg["workload"], config=args.config, agent=cfg["agent"], state_shape=list(cfg["shape"]),
              global_batch=glob, per_gpu_batch=per, replay=N_REPLAY, batches_per_step=BATCHES_PER_STEP,
              parallelism="dp%d" % args.gpus, scaling=scaling)


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
  chunk = max(1, (64 << 20) // row)
  for s0 in range(0, slots, chunk):
    k = torch.randint(0, 256, (min(chunk, slots - s0), row), device=rm.device, generator=g, dtype=torch.int32)
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
def build_agent(cfg, per_gpu_batch):
  """the drop-in network classes of the config's agent with identically seeded weights on every rank -> (engine, train callable)"""
  import torch
  from cartpoleplusplus_b200 import base_network, ddpg_cartpole, naf_cartpole
  shape = cfg["shape"]
  H, W, _, Cc, R = shape
  flags = ["--use-raw-pixels", "--render-width=%d" % W, "--render-height=%d" % H, "--action-repeats=%d" % R, "--num-cameras=%d" % Cc,
           "--batch-size=%d" % per_gpu_batch, "--replay-memory-size=%d" % N_REPLAY]
  s1p, s2p = base_network.Placeholder(shape, "s1"), base_network.Placeholder(shape, "s2")
  rng = np.random.RandomState(42)                     # identical replicas: same xavier draw on every rank
  if cfg["agent"] == "ddpg":
    o = ddpg_cartpole.set_opts(ddpg_cartpole.default_opts(flags))
    actor = ddpg_cartpole.ActorNetwork("actor", s1p, 2)
    critic = ddpg_cartpole.CriticNetwork("critic", actor)
    tactor = ddpg_cartpole.ActorNetwork("target_actor", s2p, 2)
    tcritic = ddpg_cartpole.CriticNetwork("target_critic", tactor)
    actor.init_ops_for_training(critic)
    critic.init_ops_for_training(tcritic)
    eng = actor._engine
    for net in (actor, critic, tactor, tcritic):
      net.flat_params().copy_(torch.from_numpy(net.initial_flat(rng)))
    tactor.set_as_target_network_for(actor, o.target_update_rate)
    tcritic.set_as_target_network_for(critic, o.target_update_rate)
    nets = dict(actor=actor, critic=critic)
  else:
    o = naf_cartpole.set_opts(naf_cartpole.default_opts(flags))
    value = naf_cartpole.ValueNetwork("value", s1p, o.hidden_layers)
    tvalue = naf_cartpole.ValueNetwork("target_value", s2p, o.hidden_layers)
    naf = naf_cartpole.NafNetwork("naf", s1p, s2p, value, tvalue, 2)
    eng = naf._engine
    for net in (value, naf.mu_net, naf.l_net):
      net.flat_params().copy_(torch.from_numpy(net.initial_flat(rng)))
    tvalue.set_as_target_network_for(value, o.target_update_rate)
    nets = dict(value=value, naf=naf)
  return eng, o, nets


def run_native(args):
  import ctypes as C
  import torch
  from cartpoleplusplus_b200 import _lib, dp as dpmod
  from cartpoleplusplus_b200.replay_memory import ReplayMemory, Batch

  cfg, BATCH, GLOBAL, scaling = plan(args)
  SHAPE = cfg["shape"]
  CIN = int(np.prod(SHAPE[2:]))
  is_ddpg = cfg["agent"] == "ddpg"
  dp = dpmod.DataParallel()
  assert dp.world_size == args.gpus or dp.world_size == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
  torch.cuda.set_device(dp.local_rank)
  dev = torch.device("cuda", dp.local_rank)
  lib = _lib.lib()
  G = dp.world_size

  eng, o, nets = build_agent(cfg, BATCH)
  if dp.enabled and not args.no_allreduce:
    eng.set_data_parallel(dp, lib_comm=not args.host_allreduce)     # rank 0's weights broadcast; the all-reduce moves inside the library

  rm = ReplayMemory(N_REPLAY, SHAPE, 2)
  fill_replay(rm)
  rm.sampler.seed(0)                                   # np.random.seed(0) stream, identical on every rank

  # Two sets of batch buffers: the replay gather of step i+1 (a1, a2 and the whitening statistics) is enqueued on a second
  # stream right after the training step i was launched and runs next to it; every step still does exactly one gather and
  # one training step, in index-stream order.
  gs = torch.cuda.Stream(device=dev)
  idx_pin = [torch.empty(GLOBAL, dtype=torch.int64, pin_memory=True) for _ in range(2)]
  d_idx = [torch.empty(GLOBAL, dtype=torch.int64, device=dev) for _ in range(2)]
  outs = [Batch(torch.empty((BATCH,) + SHAPE, dtype=torch.float16, device=dev), torch.empty((BATCH, 2), device=dev),
                torch.empty((BATCH, 1), device=dev), torch.empty((BATCH, 1), device=dev),
                torch.empty((BATCH,) + SHAPE, dtype=torch.float16, device=dev)) for _ in range(2)]
  moms = [(torch.empty(2 * CIN, dtype=torch.float32, device=dev), torch.empty(2 * CIN, dtype=torch.float32, device=dev)) for _ in range(2)]
  copied = [torch.cuda.Event() for _ in range(2)]      # the index vector left its pinned buffer
  ready = [torch.cuda.Event() for _ in range(2)]       # gather + statistics of this buffer set are complete
  free = [torch.cuda.Event() for _ in range(2)]        # the training step that read this buffer set is complete
  pending = {}

  def issue_gather(i):
    k = i % 2
    idxs = rm.random_indexes(GLOBAL)                                       # a1: host MT19937, bit exact; the GLOBAL index vector
    if i >= 2:
      copied[k].synchronize()
    idx_pin[k].copy_(torch.from_numpy(idxs))
    with torch.cuda.stream(gs):
      if i >= 2:
        gs.wait_event(free[k])
      d_idx[k].copy_(idx_pin[k], non_blocking=True)
      copied[k].record(gs)
      mine = d_idx[k][dp.rank * BATCH:(dp.rank + 1) * BATCH]
      batch = rm.batch_at(idxs[dp.rank * BATCH:(dp.rank + 1) * BATCH], d_idxs=mine, out=outs[k])   # a2: gather kernel, this rank's shard
      rm.batch_moments(d_idx[k], 1, out=moms[k][0]); rm.batch_moments(d_idx[k], 2, out=moms[k][1])  # global-batch whitening statistics
      ready[k].record(gs)
    pending[i] = batch

  def train(batch, moments):
    if is_ddpg:
      eng.train_step(batch, moments=moments)                               # a3-a9, a11-a12
    else:
      eng.train(batch, moments=moments, sync=False)                        # a10-a12 (no host read-back in the loop)

  def step(i):
    k = i % 2
    if i not in pending:
      issue_gather(i)
    batch = pending.pop(i)
    torch.cuda.current_stream().wait_event(ready[k])
    train(batch, moms[k])
    free[k].record()
    issue_gather(i + 1)
    if (i + 1) % BATCHES_PER_STEP == 0:
      eng.update_targets()                                                 # a13

  def timed(fn, steps, warmup, clocks=None, drain=None, first=0):
    for i in range(warmup):
      fn(first + i)
    if clocks:
      clocks.start(); time.sleep(0.25)       # before the barrier: a rank that sleeps after it stalls its peers' first collective
    dp.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.cpp_launch_count()
    e0.record()
    for i in range(steps):
      fn(first + warmup + i)
    if drain is not None:
      drain()                               # work the steps put on other streams belongs to the timed region
    e1.record()
    torch.cuda.synchronize(); dp.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dp.all_reduce_max(ms)
    return float(ms.item()), lib.cpp_launch_count() - l0, (clocks.stop() if clocks else None)

  W = max(3, args.warmup)
  drain = lambda: torch.cuda.current_stream().wait_stream(gs)
  clocks = ClockSampler(dp.local_rank) if dp.rank == 0 else None
  ms, launches, clk = timed(step, args.steps, W, clocks, drain=drain)
  steps_per_s = args.steps / (ms / 1e3)
  value = (G if scaling == "weak" else 1) * steps_per_s

  # ---- sustained: the same loop for >= --sustained-seconds (K steps of ~0.65 ms are a 30 ms burst)
  sustained = None
  if args.sustained_seconds > 0:
    n_s = int(max(args.steps, np.ceil(args.sustained_seconds / (ms / args.steps / 1e3))))
    sclk = ClockSampler(dp.local_rank) if dp.rank == 0 else None
    sms, _, sc = timed(step, n_s, 0, sclk, drain=drain, first=W + args.steps)
    sustained = dict(steps=n_s, seconds=sms / 1e3, ms_per_step=sms / n_s,
                     value=(G if scaling == "weak" else 1) * n_s / (sms / 1e3), unit=UNIT, clocks=sc)

  # ---- data parallel: the replicas must hold identical bits after all these steps
  replicas_identical = None
  if dp.enabled and not args.no_allreduce:
    mine = torch.cat([eng.buffers["params"], eng.buffers["target_params"]])
    ref = mine.clone()
    dp.broadcast(ref, 0)
    same = torch.tensor([1.0 if torch.equal(ref, mine) else 0.0], dtype=torch.float64, device=dev)
    same = -dp.all_reduce_max(-same)
    replicas_identical = bool(same.item() == 1.0)
    assert replicas_identical, "data-parallel replicas diverged"

  # ---- e2e: host (pinned) batches through the reference-facing API, H2D + loss D2H inside the timed region
  e2e = None
  if not args.skip_e2e:
    host, hmoms = [], []
    for j in range(4):
      idxs = rm.random_indexes(GLOBAL)
      d = torch.from_numpy(idxs).to(dev)
      b = rm.batch_at(idxs[dp.rank * BATCH:(dp.rank + 1) * BATCH])
      host.append(Batch(*[t.cpu().pin_memory() for t in b]))
      # data parallel: every rank whitens with the statistics of the GLOBAL batch (the batch's producer knows them; 2 x 2C floats)
      hmoms.append((rm.batch_moments(d, 1), rm.batch_moments(d, 2)) if dp.enabled else None)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    losses = []

    def e2e_step(i):
      j = i % len(host)
      b = host[j]
      train(b, hmoms[j])                     # == actor.train(b.state_1); critic.train(b) / naf.train(b): H2D of this batch unless prefetched
      eng.prefetch(host[(i + 1) % len(host)])   # the next batch's H2D copy overlaps this step's kernels (copy stream)
      if (i + 1) % BATCHES_PER_STEP == 0:
        eng.update_targets()
      losses.append(eng.last_loss())         # D2H read of the step's loss (syncs, like Session.run returning)

    ems, _, _ = timed(e2e_step, args.steps, W)
    e2e = dict(value=(G if scaling == "weak" else 1) * args.steps / (ems / 1e3), unit=UNIT, h2d_bytes_per_step=int(h2d) * G,
               d2h_bytes_per_step=4 * G, ms_per_step=ems / args.steps, h2d_gb_per_s_per_rank=h2d * args.steps / (ems / 1e3) / 1e9,
               note="host->device bound: one rank's fp16 states of one batch are %.2f MB, i.e. this rate IS the pinned-memory PCIe copy "
                    "rate of the box (all GPUs of the pool's VMs report NUMA node 0; nothing to pin to)" % (h2d / 1e6),
               api="%s(Batch of pinned host tensors) + prefetch(next batch) + last_loss(); every batch is copied host->device inside "
                   "the timed region, overlapped with the previous step" % ("DDPGEngine.train_step" if is_ddpg else "NAFEngine.train"))

  # ---- roofline: the whole step against the tensor peak, and (c3) every kernel family timed alone
  roof = None
  if dp.rank == 0 and not args.skip_roofline:
    peaks = measured_peaks()
    step_flops = 2.0 * cfg["mac_per_sample"] * BATCH                      # per GPU and step
    step_tf = step_flops / (ms / args.steps / 1e3) / 1e12
    roof = dict(bound="tensor", unit="TFLOP/s", peak=peaks["bf16_tflops"], peak_source=peaks["source"] + " bf16 dense (burst)",
                peak_sustained=peaks["bf16_tflops_sustained"],
                step_achieved=step_tf, step_frac=step_tf / peaks["bf16_tflops"], step_frac_of_sustained=step_tf / peaks["bf16_tflops_sustained"],
                step_algorithmic_gflop_per_gpu=step_flops / 1e9,
                step_note="algorithmic FLOPs of the minimum necessary graph (SURVEY.md Appendix B) / measured step time; the convs "
                          "issue 2.7x (conv1: N 20 of 48 columns, K 225 of 256) to ~11x (conv2: fp16 hi+lo pieces of activations AND "
                          "weights, 20 of 32 columns per tap row, 96 of 128 lanes) more MMA work than that to stay within 1e-5 of fp32")
    if args.config == "c3":
      from scripts.bench_kernels import kernel_table
      rows = kernel_table(rm=rm)
      tot = sum(r["us_median"] * r["launches_per_step"] for r in rows)
      for r in rows:
        r["share_of_serial_kernel_time"] = r["us_median"] * r["launches_per_step"] / tot
        if "useful_tflops" in r:
          r["frac_of_tensor_peak"] = r["useful_tflops"] / peaks["bf16_tflops"]
        else:
          r["frac_of_hbm_peak"] = r["gb_per_s"] / peaks["hbm_gbs"]
      top = max(rows, key=lambda r: r["share_of_serial_kernel_time"])
      traffic, traffic_src = None, None
      tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
      if os.path.exists(tp):               # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per kernel name
        tj = json.load(open(tp))
        ent = tj.get("kernels", {}).get(top["kernel"])
        if ent:
          traffic, traffic_src = int(ent["dram_read_bytes"]) + int(ent["dram_write_bytes"]), ent["source"]
      roof.update(kernel=top["kernel"] + ": " + top["what"], achieved=top["useful_tflops"], frac=top["useful_tflops"] / peaks["bf16_tflops"],
                  kernel_us=top["us_median"], launches_per_step=top["launches_per_step"],
                  algorithmic_flops_per_launch=top["useful_gflop"] * 1e9, traffic=traffic, traffic_unit="bytes of DRAM traffic per launch",
                  traffic_source=traffic_src, kernels=rows,
                  note="`kernel` is the family with the largest share of the step's serialised kernel time (us x launches per step), "
                       "timed alone on the whole GPU with CUDA events, L2 flushed between launches; `kernels` lists every family")
    else:
      roof.update(kernel="whole step (per-kernel table only for c3)", achieved=step_tf, frac=step_tf / peaks["bf16_tflops"], traffic=None)

  # ---- rollout latency (8f row 2): actor.action_given(state) / naf.action_given(state) at B = 1, host fp32 state in, host
  # action out, as bullet_cartpole's env loop calls it once per step (ddpg_cartpole.py:317, naf_cartpole.py:349)
  latency = None
  if dp.rank == 0 and not args.skip_e2e:
    rs_ = np.random.RandomState(5)
    states = [(rs_.randint(0, 256, SHAPE).astype(np.float16) / np.float16(255)).astype(np.float32) for _ in range(8)]
    net = nets["actor"] if is_ddpg else nets["naf"]
    call = (lambda s_: net.action_given(s_)) if is_ddpg else (lambda s_: net.action_given(s_, add_noise=False))
    for i in range(20):
      call(states[i % 8])
    ts = []
    for i in range(200):
      t0 = time.perf_counter(); call(states[i % 8]); ts.append((time.perf_counter() - t0) * 1e6)
    latency = dict(call="%s.action_given(state fp32 %s), B=1, host in / host out" % ("ActorNetwork" if is_ddpg else "NafNetwork", list(SHAPE)),
                   us_median=float(np.median(ts)), us_p10=float(np.percentile(ts, 10)), us_p90=float(np.percentile(ts, 90)), calls=200,
                   h2d_bytes=int(states[0].nbytes), d2h_bytes=12,
                   path="fp32 -> exact fp16 copy on the device, tensor-core trunk, FC stack; one CUDA graph replay per call")
  cpu = None
  if dp.rank == 0 and not args.skip_cpu_baseline:
    cpu = cpu_reference(args, steps=3, warmup=1)

  if dp.rank == 0:
    line = dict(metric=cfg["metric"], value=value, unit=UNIT, n_gpus=G, steps=args.steps, warmup=W, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling=scaling, vs_baseline=None,
                dtype="f32 (conv MMAs: fp16 hi+lo operand pieces, fp32 accumulate; FC / elementwise fp32)", data="synthetic",
                config=config_dict(args),
                notes=dict(l2="inputs larger than L2: each step gathers %.1f MB of random rows from a %.0f MB fp16 replay slab"
                              % (2 * BATCH * rm.row_elems * 2 / 1e6, rm.d_state.numel() * 2 / 1e6),
                           pipeline="the replay gather of step i+1 runs on a second stream next to training step i (two batch buffer "
                                    "sets); each timed step = one gather + one training step",
                           optimiser_steps_per_s=steps_per_s,
                           all_reduce=None if not dp.enabled else ("NONE (diagnosis: independent replicas)" if args.no_allreduce else
                                                                   "host-issued torch.distributed" if args.host_allreduce else
                                                                   "inside the step's CUDA graph (csrc/comm.cu), transport %s" % eng.transport)),
                clocks=clk, sustained=sustained, replicas_identical=replicas_identical, e2e=e2e, gpu_launches=int(launches),
                action_latency=latency, roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line))
  dp.close()


# ------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference(args, steps, warmup, budget_s=120.0):
  """the reference's CPU path on this box's host cores: py3 restatement of ReplayMemory.batch + fp16->fp32 feed cast
  + torch-CPU fp32 restatement of the TF graph (oracle/, kind 'port'; TensorFlow itself cannot run here)"""
  import torch
  from oracle import nets_oracle as no
  from oracle.replay_oracle import ReplayOracle
  cfg, _, _, _ = plan(args)
  SHAPE, BATCH = cfg["shape"], cfg["batch"]            # the reference is one replica: it trains on the config's whole batch
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  rs = np.random.RandomState(42)
  P = {}
  if cfg["agent"] == "ddpg":
    for d in (no.ddpg_actor("actor", SHAPE, True), no.ddpg_critic("critic", SHAPE, True)):
      P.update(no.init_params(d, rs, torch.float32))
    P.update(no.retarget({k: v for k, v in P.items() if k.startswith("actor/")}, "actor", "target_actor"))
    P.update(no.retarget({k: v for k, v in P.items() if k.startswith("critic/")}, "critic", "target_critic"))
    orc = no.DDPGOracle(SHAPE, True, P)
  else:
    for d in (no.naf_value("value", SHAPE, True), no.naf_mu(SHAPE, True), no.naf_l(SHAPE, True)):
      P.update(no.init_params(d, rs, torch.float32))
    P.update(no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value"))
    orc = no.NAFOracle(SHAPE, True, P)
  n = N_REPLAY                                          # the same replay as the native arm
  rm = ReplayOracle(n, SHAPE, 2, rng=np.random.RandomState(0))
  r2 = np.random.RandomState(1234)
  for s0 in range(0, rm.state.shape[0], 256):           # fp16(u8)/255 pixels, filled in chunks (the slab is ~0.9 GB at c3)
    blk = rm.state[s0:s0 + 256]
    blk[:] = (r2.randint(0, 256, blk.shape).astype(np.float16) / np.float16(255))
  rm.state_1_idx[:] = np.arange(n); rm.state_2_idx[:] = np.arange(1, n + 1)
  rm.action[:], rm.reward[:], rm.terminal_mask[:] = synthetic_tables(n)
  rm.insert, rm.full = 0, True

  def one(batch_size, i):
    b = rm.batch(batch_size)
    feed = (b.state_1.astype(np.float32), b.action, b.reward, b.terminal_mask, b.state_2.astype(np.float32))   # feed_dict cast
    if cfg["agent"] == "ddpg":
      orc.actor_train(feed[0])
      orc.critic_train(feed)
    else:
      orc.train(feed)
    if (i + 1) % BATCHES_PER_STEP == 0:
      orc.update_targets()

  bs = BATCH
  t0 = time.time(); one(min(bs, 64), 0); t1 = (time.time() - t0) * (bs / min(bs, 64))
  if t1 * (steps + warmup) > budget_s:                   # keep the arm bounded: shrink the per-step sample
    bs = max(8, int(BATCH * budget_s / (t1 * (steps + warmup))))
  for i in range(warmup):
    one(bs, i)
  ts = []
  for i in range(steps):
    t0 = time.time(); one(bs, warmup + i); ts.append(time.time() - t0)
  per_step = float(np.median(ts)) * (BATCH / bs)         # config-batch equivalent
  return dict(value=1.0 / per_step, unit=UNIT, cores=cores, kind="port",
              sample="%d timed grad-steps at batch %d (median, scaled to batch %d) incl. replay gather + fp16->fp32 feed cast; "
                     "torch-CPU fp32 oracle port with %d threads" % (steps, bs, BATCH, cores),
              ms_per_step=per_step * 1e3)


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cfg, _, _, scaling = plan(args)
  W = max(3, args.warmup)                                # the same warm-up rule as the native arm
  cpu = cpu_reference(args, steps=args.steps, warmup=W, budget_s=240.0)
  line = dict(impl="reference", metric=cfg["metric"], value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=W,
              ms_per_step=cpu["ms_per_step"], higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="f32", data="synthetic",
              config=config_dict(args),
              cpu_baseline=dict(kind=cpu["kind"], cores=cpu["cores"], sample=cpu["sample"], value=cpu["value"], unit=UNIT),
              e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
  print(json.dumps(line))


if __name__ == "__main__":
  a = parse_args()
  if a.impl == "reference":
    run_reference(a)
  else:
    run_native(a)
