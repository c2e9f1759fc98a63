import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as orc
from tehmm_b200 import synth
from tehmm_b200.engine import get_engine
orc.build()
eng = get_engine(0)
for (N, syms, lens, opts) in [
    (2, (16, 33, 3, 16, 7, 4, 1, 33, 3), [64, 1, 20000, 4097, 1000], dict(chunk_tiles=1, warmup=8, fine_len=100, tile=0)),
    (32, (4,), [1, 20000], dict(chunk_tiles=0, warmup=0, fine_len=100, tile=1)),
    (30, synth.BENCH_SYMS, [20000], dict(chunk_tiles=0, warmup=0, fine_len=0, tile=1)),
    (30, synth.BENCH_SYMS, [20000], dict(chunk_tiles=4, warmup=64, fine_len=0, tile=1))]:
    m = synth.make_model(N=N, syms=syms, seed=5)
    seqs = [synth.sample_obs(m, n, seed=7 + i)[0] for i, n in enumerate(lens)]
    for k, v in opts.items():
        eng.ctx.set_option(k, v)
    eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
    eng.upload_batch(seqs)
    es = eng.estep(precision="f64")
    K, Nn, S = m["table"].shape
    s0, tr, ob = np.zeros(Nn), np.zeros((Nn, Nn)), np.zeros((K, Nn, S))
    for o in seqs:
        orc.estep_sequence(o, m["table"], 1.0, m["log_start"], m["log_trans"], None, s0, tr, ob)
    rel = np.abs(es["trans"] - tr) / np.maximum(np.abs(tr), 1e-300)
    big = tr > 1e-6
    print("N=%d lens=%s opts=%s: trans max abs err %.3g, max rel err (entries > 1e-6) %.3g, obs max rel %.3g, repairs fwd %d bwd %d" % (
        N, lens, opts, np.abs(es["trans"] - tr).max(), rel[big].max() if big.any() else 0,
        (np.abs(es["obs"] - ob) / np.maximum(ob, 1e-300))[ob > 1e-6].max(),
        eng.ctx.stat("repaired_chunks_forward"), eng.ctx.stat("repaired_chunks_backward")))
