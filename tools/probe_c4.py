#!/usr/bin/env python
"""Config 4 (24 sequences, 12.4 M bins) through the host-buffer decode calls: per-phase times of one call."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import synth
from tehmm_b200.emission import IndependentMultinomialEmissionModel
from tehmm_b200.hmm import MultitrackHmm
from tehmm_b200.engine import get_engine

m = synth.make_model(N=30, seed=0)
lens = synth.bench_lengths("c4")
seqs = [synth.sample_obs(m, n, seed=200 + i)[0] for i, n in enumerate(lens)]
em = IndependentMultinomialEmissionModel(30, list(m["syms"]), zeroAsMissingData=True)
em.logProbs = m["table"].copy()
hv = MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy())
eng = get_engine(0)
ctx = eng.ctx
one = [np.concatenate(seqs)]
for name, arg in (("24 sequences", seqs), ("one sequence of the same length", one)):
    for _ in range(2):
        hv.decode_both_batch(arg)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hv.decode_both_batch(arg)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    os.environ["TEHMM_HOST_TRACE"] = "2"
    hv.decode_both_batch(arg)
    br = {k: float(ctx.lib.tehmm_decode_host_phase_ms(ctx.handle, i)) for i, k in enumerate(("set_batch", "h2d_and_emission", "trellis", "d2h_and_widen"))}
    del os.environ["TEHMM_HOST_TRACE"]
    ctx.set_option("timing", 1)
    hv.decode_both_batch(arg)
    torch.cuda.synchronize()
    us = {k: ctx.stat("us_" + k) for k in ("emission", "forward", "backward", "viterbi_dp", "traceback", "rescore")}
    ctx.set_option("timing", 0)
    rep = {k: ctx.stat(k) for k in ("repair_passes_forward", "repair_passes_backward", "repair_passes_viterbi", "repair_passes_traceback", "warmup")}
    print(json.dumps({"kernel_us": us, "stats": rep}), flush=True)
    t0 = time.perf_counter()
    v = hv.decode_batch(arg); t1 = time.perf_counter()
    print(json.dumps({"what": name, "decode_both_ms": ts, "traced_phases_ms": br, "decode_batch_viterbi_ms": (t1 - t0) * 1e3}), flush=True)
