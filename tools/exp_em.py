"""Per-kernel time of one E-step at the bench shape (10 M x 30 x 10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
ctx = _lib.get_context(0)
eng = Engine(ctx)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
d_obs = torch.from_numpy(obs).to("cuda").reshape(-1)
eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
eng.estep(device_result=True)
ctx.set_option("timing", 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    eng.estep(device_result=True)
b.record(); torch.cuda.synchronize()
print("E-step %.3f ms" % (a.elapsed_time(b) / 3))
for k in ("emission", "forward", "backward", "xi", "emission_stats"):
    print("  %-16s %d us" % (k, ctx.stat("us_" + k)))
