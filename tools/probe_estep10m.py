#!/usr/bin/env python
"""Kernel times of one E-step on the bench shape (one sequence of 10 M steps, 30 states, 10 tracks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine

ctx = _lib.get_context(0); eng = Engine(ctx)
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, 10_000_000, seed=1)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
eng.upload_batch([obs])
for _ in range(3):
    eng.estep(device_result=True)
torch.cuda.synchronize()
ctx.set_option("timing", 1)
t0 = time.perf_counter()
for _ in range(5):
    eng.estep(device_result=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(json.dumps({"estep_ms": 1e3 * dt, "us": {k: ctx.stat("us_" + k) for k in ("emission", "forward", "backward", "xi", "emission_stats")}}))
