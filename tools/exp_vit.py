"""Viterbi DP time against the chunk count (tiles of 64 steps per chunk)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine
T = 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
ctx = _lib.get_context(0)
eng = Engine(ctx)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
d_obs = torch.from_numpy(obs).to("cuda").reshape(-1)
for tiles in [int(a) for a in sys.argv[1:]] or [0]:
    ctx.set_option("chunk_tiles", tiles)
    eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
    prec, tdt = eng._prec("f32")
    elog, _, _ = eng.run_emission(prec, tdt, None, True, False)
    eng.run_viterbi(prec, elog, None, None, want64=False)
    ctx.set_option("timing", 1)
    for _ in range(3):
        st, _, lp = eng.run_viterbi(prec, elog, None, None, want64=False)
    torch.cuda.synchronize()
    print("chunk_tiles", tiles, "chunks", ctx.stat("chunks"), "viterbi_dp us", ctx.stat("us_viterbi_dp"), "lp", float(lp[0]))
    ctx.set_option("timing", 0)
