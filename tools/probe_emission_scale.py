#!/usr/bin/env python
"""Emission kernel time against the number of steps (50 states, one sequence): first call into freshly allocated
memory and repeated calls into the cached block."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine

ctx = _lib.get_context(0); eng = Engine(ctx)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
m = synth.make_model(N=N, seed=0)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
base, _ = synth.sample_obs(m, 10_000_000, seed=1)
for reps in (1, 5, 10, 20):
    obs = np.concatenate([base] * reps) if reps > 1 else base
    T = obs.shape[0]
    d_obs = torch.from_numpy(obs.reshape(-1)).to(eng.device)
    eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
    prec, tdt = eng._prec("f32")
    torch.cuda.empty_cache()
    out = {}
    for which, args in (("elog", (True, False)), ("blin", (False, True))):
        ts = []
        for it in range(3):
            ctx.set_option("timing", 1)
            r = eng.run_emission(prec, tdt, None, *args)
            torch.cuda.synchronize()
            ts.append(ctx.stat("us_emission"))
            ctx.set_option("timing", 0)
            if it == 0:
                del r
                torch.cuda.empty_cache()      # the second call allocates afresh too; the third reuses the cached block
            else:
                del r
        out[which] = ts
    print(json.dumps({"N": N, "T": T, "us": out, "us_per_10M": {k: [round(v / reps) for v in vs] for k, vs in out.items()}}), flush=True)
    del d_obs
