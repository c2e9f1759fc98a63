"""Timing experiment: forward (tile kernels) with and without the alpha store; backward MAP."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
ctx = _lib.get_context(0); eng = Engine(ctx)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
d_obs = torch.from_numpy(obs).cuda().reshape(-1)
eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
prec, tdt = eng._prec("f32")
elog, blin, rowmax = eng.run_emission(prec, tdt, None, True, True)
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("forward with alpha   %.3f ms" % timeit(lambda: eng.run_forward(prec, tdt, blin, rowmax, None, True)))
print("forward score only   %.3f ms" % timeit(lambda: eng.run_forward(prec, tdt, blin, rowmax, None, False)))
alpha, lp = eng.run_forward(prec, tdt, blin, rowmax, None, True)
print("backward MAP         %.3f ms" % timeit(lambda: eng.run_backward(prec, tdt, _lib.BWD_MAP, blin, alpha, None)))
print("backward none        %.3f ms" % timeit(lambda: eng.run_backward(prec, tdt, 0, blin, alpha, None)))
print("backward POST        %.3f ms" % timeit(lambda: eng.run_backward(prec, tdt, _lib.BWD_POSTERIORS, blin, alpha, None)))
print("emission both        %.3f ms" % timeit(lambda: eng.run_emission(prec, tdt, None, True, True)))
print("emission lin only    %.3f ms" % timeit(lambda: eng.run_emission(prec, tdt, None, False, True)))
print("viterbi              %.3f ms" % timeit(lambda: eng.run_viterbi(prec, elog, None, None, want64=False)))
import torch
N, K, S = eng.N, eng.K, eng.S
packed = torch.zeros(2 + N + N * N + K * N * S, dtype=torch.float64, device=eng.device)
def bw_trans():
    return eng.run_backward(prec, tdt, _lib.BWD_TRANS | _lib.BWD_POSTERIORS, blin, alpha, None, start_trans=packed[2:2 + N + N * N])
print("backward TRANS|POST  %.3f ms" % timeit(bw_trans))
print("backward TRANS       %.3f ms" % timeit(lambda: eng.run_backward(prec, tdt, _lib.BWD_TRANS, blin, alpha, None, start_trans=packed[2:2 + N + N * N])))
post, _, _ = bw_trans()
print("emission stats       %.3f ms" % timeit(lambda: eng.run_emission_stats(prec, post, None, packed[2 + N + N * N:], S)))
print("estep total          %.3f ms" % timeit(lambda: eng.estep(device_result=True)))
