#!/usr/bin/env python
"""Experiment: accumulation-chain length of the conv weight-gradient kernel (cpp_set_option "wgrad_flush_steps") against
(a) the worst gradient error of the pinned-routing whole-step test at c3 and (b) the step time."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartpoleplusplus_b200 import _lib as L          # noqa: E402
from tests.test_gpu_step_pinned import run_ddpg_pinned   # noqa: E402

for steps in [int(v) for v in sys.argv[1:]] or [128, 64, 32, 16]:
  L.check(L.lib().cpp_set_option(b"wgrad_flush_steps", steps))
  for seed in (77, 78):
    report, rep = run_ddpg_pinned((64, 64, 3, 1, 3), 256, seed)
    print(json.dumps(dict(flush_steps=steps, seed=seed, worst=report["worst grad"],
                          conv1={k: "%.2e" % v for k, v in rep.items() if "conv1/weights" in k or "conv2/weights" in k})))
