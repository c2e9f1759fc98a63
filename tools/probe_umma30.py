#!/usr/bin/env python
"""<= 32 states, one sequence: forward / backward passes by the tcgen05 kernels (option "umma" = 1: one thread per
chunk, 2: two threads per chunk) against the mma.sync tile kernels (default)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = _lib.get_context(0); eng = Engine(ctx)
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
eng.upload_batch([obs])
prec, tdt = eng._prec("f32")
_, blin, rowmax = eng.run_emission(prec, tdt, None, False, True)
res = {}
for opt in (0, 1, 2):
    ctx.set_option("umma", opt)
    for _ in range(2):
        alpha, lp = ctx.optimistic(lambda: eng.run_forward(prec, tdt, blin, rowmax, None))
    torch.cuda.synchronize()
    ctx.set_option("timing", 1)
    for _ in range(3):
        alpha, lp = ctx.optimistic(lambda: eng.run_forward(prec, tdt, blin, rowmax, None))
    torch.cuda.synchronize()
    r = {"forward_us": ctx.stat("us_forward"), "logprob": float(lp[0].item())}
    for fl, name in ((2, "bwd_map_us"), (3, "bwd_post_map_us")):
        for _ in range(3):
            outs = ctx.optimistic(lambda: eng.run_backward(prec, tdt, fl, blin, alpha, None))
        torch.cuda.synchronize()
        r[name] = ctx.stat("us_backward")
        if fl == 2:
            r["map"] = outs[1].cpu().numpy()
        del outs
    ctx.set_option("timing", 0)
    res[opt] = r
ctx.set_option("umma", 0)
base = res[0]
for opt in (1, 2):
    res[opt]["map_agreement"] = float(np.mean(res[opt].pop("map") == base["map"]))
    res[opt]["logprob_rel_diff"] = abs(res[opt]["logprob"] - base["logprob"]) / abs(base["logprob"])
base.pop("map")
print(json.dumps({"T": T, "tile": res[0], "umma_1_thread": res[1], "umma_2_threads": res[2],
                  "umma_passes": ctx.stat("umma_passes"), "repaired_forward": ctx.stat("repaired_chunks_forward"),
                  "repaired_backward": ctx.stat("repaired_chunks_backward")}), flush=True)
