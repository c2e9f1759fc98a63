"""Per-kernel times of the N = 50 path (config 5 shape) at 10 M steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
m = synth.make_model(N=N, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
ctx = _lib.get_context(0)
eng = Engine(ctx)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
d_obs = torch.from_numpy(obs).to("cuda").reshape(-1)
eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
prec, tdt = eng._prec("f32")
def sweep():
    elog, blin, rowmax = eng.run_emission(prec, tdt, None, True, True)
    alpha, lp = eng.run_forward(prec, tdt, blin, rowmax, None)
    eng.run_backward(prec, tdt, _lib.BWD_MAP, blin, alpha, None)
    eng.run_viterbi(prec, elog, None, None, want64=False)
sweep()
ctx.set_option("timing", 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(2):
    sweep()
b.record(); torch.cuda.synchronize()
print("N=%d T=%d sweep %.2f ms" % (N, T, a.elapsed_time(b) / 2))
for k in ("emission", "forward", "backward", "viterbi_dp", "traceback", "rescore"):
    print("  %-12s %d us" % (k, ctx.stat("us_" + k)))
