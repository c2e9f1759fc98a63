#!/usr/bin/env python
"""What speculation costs on slow-mixing input: repair passes, repaired chunks, final warm-up and time of the
forward / backward(MAP) / Viterbi stages for (a) the bench model, (b) the bench model with all-missing
stretches, (c) a near-reducible transition matrix (diagonal 0.9999) with the same stretches.

    python tools/probe_hard.py [--T 2000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=2_000_000)
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    import torch
    from tehmm_b200 import _lib, synth
    from tehmm_b200.engine import Engine

    ctx = _lib.get_context(0)
    eng = Engine(ctx)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    cases = [("bench", 0.9, False), ("bench+gaps", 0.9, True), ("sticky0.9999+gaps", 0.9999, True)]
    for name, sticky, gaps in cases:
        m = synth.make_model(N=30, seed=0, sticky=sticky)
        obs, _ = synth.sample_obs(m, args.T, seed=1)
        if gaps:
            obs, spans = synth.add_missing_stretches(obs)
        eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
        eng.upload_batch([obs])
        for rep in range(2):          # the second call sees the adapted warm-up
            stats0 = {k: ctx.stat(k) for k in ("repair_passes_forward", "repair_passes_backward", "repair_passes_viterbi",
                                               "repair_passes_traceback", "repaired_chunks_forward", "repaired_chunks_backward",
                                               "repaired_chunks_viterbi", "repaired_chunks_traceback", "fallbacks")}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = eng.posteriors(renorm_eps=True, want_post=False, want_map=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            lp, st = eng.viterbi()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            d = {k: ctx.stat(k) - v for k, v in stats0.items()}
            print(json.dumps({"case": name, "call": rep, "T": args.T, "fwd_bwd_map_ms": 1e3 * (t1 - t0), "viterbi_ms": 1e3 * (t2 - t1),
                              "warmup_now": ctx.stat("warmup"), "fine_chunks": ctx.stat("fine_chunks"), "chunks": ctx.stat("chunks"),
                              **d}), flush=True)


if __name__ == "__main__":
    main()
