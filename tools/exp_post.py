"""End-to-end time of MultitrackHmm.score_samples (posteriors (T, N) float64 back on the host)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from tehmm_b200 import synth
from run_configs import make_hmm
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
hmm, _ = make_hmm(m)
for it in range(4):
    t0 = time.perf_counter()
    lp, post = hmm.score_samples(obs)
    dt = time.perf_counter() - t0
    print("score_samples %d x 30: %.1f ms  (%s %s, row sums %.6f..%.6f)" % (T, dt * 1e3, post.dtype, post.shape, post[:1000].sum(1).min(), post[:1000].sum(1).max()))
    del post
