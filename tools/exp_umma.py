import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine
T = 10_000_000
FL = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
ctx = _lib.get_context(0)
eng = Engine(ctx)
ctx.set_option('fine_len', FL)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
d_obs = torch.from_numpy(obs).to("cuda").reshape(-1)
eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
prec, tdt = eng._prec("f32")
_, blin, rowmax = eng.run_emission(prec, tdt, None, False, True)
res = {}
for um in (0, 1):
    ctx.set_option("umma", um)
    alpha, lp = eng.run_forward(prec, tdt, blin, rowmax, None)
    ctx.set_option("timing", 1)
    for _ in range(3):
        alpha, lp = eng.run_forward(prec, tdt, blin, rowmax, None)
    torch.cuda.synchronize()
    _, ms, sc = eng.run_backward(prec, tdt, _lib.BWD_MAP, blin, alpha, None)
    res[um] = (float(lp[0]), float(sc[0]), ms.clone())
    print("umma", um, "forward us", ctx.stat("us_forward"), "logprob", float(lp[0]), "map_score", float(sc[0]),
          "umma_passes", ctx.stat("umma_passes"), "repairs", ctx.stat("repaired_chunks_forward"))
    ctx.set_option("timing", 0)
print("logprob rel diff", abs(res[0][0] - res[1][0]) / abs(res[0][0]), "map states equal", bool((res[0][2] == res[1][2]).all()))
