#!/usr/bin/env python
"""Where an EM iteration of config 3 (350 sequences, 3.5 M steps, MultitrackHmm.fit) spends its time: cProfile of
ten iterations on one GPU, top functions by cumulative time."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import synth
from tehmm_b200.emission import IndependentMultinomialEmissionModel
from tehmm_b200.hmm import MultitrackHmm

m = synth.make_model(N=30, seed=0)
m0 = synth.make_model(N=30, seed=99)
lens = synth.bench_lengths("c3")
seqs = [synth.sample_obs(m, n, seed=400 + i)[0] for i, n in enumerate(lens)]
if len(sys.argv) > 1:            # every k-th sequence only: the shard of one rank of k
    seqs = seqs[::int(sys.argv[1])]

def make(n_iter):
    em = IndependentMultinomialEmissionModel(30, list(m0["syms"]), zeroAsMissingData=True)
    em.logProbs = m0["table"].copy()
    return MultitrackHmm(em, startprob=m0["pi"].copy(), transmat=m0["A"].copy(), n_iter=n_iter, thresh=0.0)

make(2).fit(seqs)
h = make(50)
torch.cuda.synchronize(); t0 = time.perf_counter()
h.fit(seqs)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("seconds per iteration: %.5f" % (dt / 50))
h = make(50)
pr = cProfile.Profile()
pr.enable(); h.fit(seqs); torch.cuda.synchronize(); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:7000])
