"""Where does an EM iteration of config 3 (350 sequences, 3.5 M steps) go?"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import synth
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from run_configs import make_hmm
m = synth.make_model(N=30, seed=0)
lens = synth.bench_lengths("c3")
seqs = [synth.sample_obs(m, n, seed=100 + i)[0] for i, n in enumerate(lens)]
m0 = synth.make_model(N=30, seed=7, zero_frac=0.0)
hmm, em = make_hmm(m0, n_iter=3, thresh=0.0)
hmm.fit(seqs)
hmm, em = make_hmm(m0, n_iter=20, thresh=0.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
hmm.fit(seqs)
pr.disable()
torch.cuda.synchronize(); print("per iteration %.2f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
