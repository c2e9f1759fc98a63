#!/usr/bin/env python
"""Kernel times of one E-step on the C3-shaped batch, without and with segment ratios."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine

ctx = _lib.get_context(0); eng = Engine(ctx)
m = synth.make_model(N=30, seed=0)
lens = synth.bench_lengths("c3")
seqs = [synth.sample_obs(m, n, seed=100 + i)[0] for i, n in enumerate(lens)]
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
eng.upload_batch(seqs)
rng = np.random.RandomState(3)
ratios = eng.upload_ratios([np.minimum(rng.geometric(1.0 / 60.0, size=n), 100) / 100.0 for n in lens])
for name, r in (("no ratios", None), ("ratios", ratios)):
    for _ in range(3):
        eng.estep(ratios=r, device_result=True)
    torch.cuda.synchronize()
    ctx.set_option("timing", 1)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.estep(ratios=r, device_result=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps({"case": name, "estep_ms": 1e3 * dt,
                      "us": {k: ctx.stat("us_" + k) for k in ("emission", "forward", "backward", "xi", "emission_stats")}}))
    ctx.set_option("timing", 0)
