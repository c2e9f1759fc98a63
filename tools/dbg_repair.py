import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine
os.environ["TEHMM_NO_DEFER"] = "1"
ctx = _lib.get_context(0); eng = Engine(ctx)
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, 2_000_000, seed=1)
obs, spans = synth.add_missing_stretches(obs, n_stretches=1, lo=20000, hi=20000)
print("span", spans, "fine chunk len", 2_000_000 / 18868, file=sys.stderr)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
eng.upload_batch([obs])
os.environ["TEHMM_DEBUG_REPAIR"] = "1"
print(eng.score(), file=sys.stderr)
A = m["A"]; ev = np.sort(np.abs(np.linalg.eigvals(A)))[::-1]; print("eig", ev[:4], file=sys.stderr)
