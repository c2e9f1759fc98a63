"""Where does the end-to-end decode time go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.emission import IndependentMultinomialEmissionModel
from tehmm_b200.hmm import MultitrackHmm
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
pinned = torch.from_numpy(obs).pin_memory()
host_obs = pinned.numpy()
em = IndependentMultinomialEmissionModel(30, list(m["syms"]), zeroAsMissingData=True)
em.logProbs = m["table"].copy()
hv = MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy())
hm = MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy(), algorithm="map")
def t(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("decode viterbi (pinned in)   %.2f ms" % t(lambda: hv.decode_batch([host_obs])))
print("decode map     (pinned in)   %.2f ms" % t(lambda: hm.decode_batch([host_obs])))
print("decode viterbi (pageable in) %.2f ms" % t(lambda: hv.decode_batch([obs])))
eng = hv._engine()
print("  upload_batch pinned        %.2f ms" % t(lambda: eng.upload_batch([host_obs])))
print("  upload_batch pageable      %.2f ms" % t(lambda: eng.upload_batch([obs])))
print("  _engine()                  %.2f ms" % t(lambda: hv._engine()))
prec, tdt = eng._prec("f32")
def vit():
    elog, _, _ = eng.run_emission(prec, tdt, None, True, False)
    return eng.run_viterbi(prec, elog, None, None)
print("  emission+viterbi device    %.2f ms" % t(vit))
st, st64, lp = vit()
print("  states64.cpu().numpy()     %.2f ms" % t(lambda: st64.cpu().numpy()))
print("  states(u8).cpu()->int64    %.2f ms" % t(lambda: st.cpu().to(torch.int64).numpy()))
pin8 = torch.empty(T, dtype=torch.uint8).pin_memory()
def viapin():
    pin8.copy_(st, non_blocking=True); torch.cuda.synchronize()
    return pin8.to(torch.int64).numpy()
print("  u8 -> pinned -> int64      %.2f ms" % t(viapin))
print("  torch threads", torch.get_num_threads(), "cpus", os.cpu_count())
