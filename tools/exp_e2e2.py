"""Per-call wall time of MultitrackHmm.decode_batch next to the library's own phase trace."""
import os, sys, time
os.environ["TEHMM_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.emission import IndependentMultinomialEmissionModel
from tehmm_b200.hmm import MultitrackHmm
T = 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
host_obs = torch.from_numpy(obs).pin_memory().numpy()
em = IndependentMultinomialEmissionModel(30, list(m["syms"]), zeroAsMissingData=True)
em.logProbs = m["table"].copy()
hv = MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy())
hm = MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy(), algorithm="map")
for it in range(6):
    t0 = time.perf_counter(); rv = hv.decode_batch([host_obs]); t1 = time.perf_counter()
    rm = hm.decode_batch([host_obs]); t2 = time.perf_counter()
    print("iter %d: viterbi %.2f ms  map %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for it in range(3):
    rv = hv.decode_batch([host_obs]); rm = hm.decode_batch([host_obs])
pr.disable()
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(18)
