"""Phases of tehmm_decode_host (TEHMM_HOST_TRACE=1) for pinned and pageable input."""
import os, sys, time
os.environ["TEHMM_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tehmm_b200 import _lib, synth
from tehmm_b200.engine import get_engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
m = synth.make_model(N=30, seed=0)
obs, _ = synth.sample_obs(m, T, seed=1)
pinned = torch.from_numpy(obs).pin_memory().numpy()
eng = get_engine(0)
eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
for name, a in (("pinned", pinned), ("pageable", obs)):
    for alg in (_lib.DECODE_VITERBI, _lib.DECODE_MAP):
        for it in range(3):
            t0 = time.perf_counter()
            eng.decode_host([a], alg)
            print("%s alg %d: %.2f ms" % (name, alg, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
# warm output buffer (no page faults): how much of d2h+widen is first-touch?
import ctypes
states = np.empty(T, dtype=np.int64); states[:] = 0
lp = np.empty(1); sc = np.empty(1)
off = np.array([0, T], dtype=np.int64)
for it in range(3):
    t0 = time.perf_counter()
    _lib.check(eng.lib.tehmm_decode_host(eng.ctx.handle, (ctypes.c_void_p * 1)(pinned.ctypes.data), 1, 1, 1, _lib.ptr(off), 0, 0, _lib.ptr(states), _lib.ptr(lp), _lib.ptr(sc)))
    print("warm output: %.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read(), open("/sys/kernel/mm/transparent_hugepage/defrag").read(), file=sys.stderr)
