#!/usr/bin/env python
"""Full-size runs of BASELINE.json configs[2..4] on the device path (one GPU), with the
size-independent checks the oracle cannot give at these sizes.

    python tools/run_configs.py [c3] [c4] [c5] [--scale 1.0] > profiles/rNN_configs.jsonl

c3  unsupervised Baum-Welch, 30 states, 10 tracks, 350 variable-length sequences (3.5 M steps),
    50 EM iterations through MultitrackHmm.fit: seconds per iteration; EM must not decrease
    the log-likelihood; the re-estimated rows must be distributions.
c4  hg19-scale decode: 24 sequences (12.4 M bins at 250 bp), Viterbi + MAP through
    MultitrackHmm.decode_batch (host buffers): seconds; decoding the batch must equal decoding a
    sequence alone (chr21), and a 300 k prefix must equal the CPU oracle's path.
c5  one sequence of 250 M steps, 50 states: Viterbi, forward log-likelihood and MAP; the Viterbi
    score must not exceed the log-likelihood, repairs are reported, and the first 2 M steps
    must decode to the same path as a stand-alone 2.2 M prefix (paths coalesce).
Each config prints one JSON line.  Uses the oracle only as the checker (tools/ is not product).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def sample_long(synth, m, T, seed, piece=10_000_000):
    out = np.empty((T, m["K"]), dtype=np.uint8)
    for i, a in enumerate(range(0, T, piece)):
        n = min(piece, T - a)
        out[a:a + n] = synth.sample_obs(m, n, seed=seed + i)[0]
    return out


def make_hmm(m, **kw):
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    em = IndependentMultinomialEmissionModel(m["N"], list(m["syms"]), zeroAsMissingData=True)
    em.logProbs = m["table"].copy()
    return MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy(), **kw), em


def run_c3(scale):
    import torch
    from tehmm_b200 import synth
    m = synth.make_model(N=30, seed=0)
    lens = [max(50, int(n * scale)) for n in synth.bench_lengths("c3")]
    seqs = [synth.sample_obs(m, n, seed=100 + i)[0] for i, n in enumerate(lens)]
    # start EM away from the generating model
    m0 = synth.make_model(N=30, seed=7, zero_frac=0.0)
    n_iter = 50
    hmm, em = make_hmm(m0, n_iter=n_iter, thresh=0.0)
    hmm.fit(seqs[:4])            # warm-up (allocator, kernels)
    hmm, em = make_hmm(m0, n_iter=n_iter, thresh=0.0)
    iter_lp = []
    real_estep = hmm._device_estep

    def estep(mine, stats, params, n_total, slots):
        lps = real_estep(mine, stats, params, n_total, slots)
        iter_lp.append(float(np.sum(lps)))
        return lps
    hmm._device_estep = estep
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hmm.fit(seqs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    total_lp = iter_lp
    A = hmm.transmat_
    tab = np.exp(em.getLogProbs())
    rows_ok = bool(np.allclose(A.sum(axis=1), 1.0, atol=1e-9))
    em_ok = True
    for k, s in enumerate(m["syms"]):
        em_ok &= bool(np.allclose(tab[k, :, 1:s + 1].sum(axis=1), 1.0, atol=1e-6))
    return {"config": "c3", "sequences": len(seqs), "steps": int(sum(lens)), "states": 30, "tracks": 10,
            "em_iterations": n_iter, "seconds_total": dt, "seconds_per_em_iteration": dt / n_iter,
            "cells_per_s_per_iteration": sum(lens) * 30 / (dt / n_iter),
            "logprob_first_last": [total_lp[0], total_lp[-1]] if total_lp else None,
            "logprob_monotone": bool(all(b >= a - 1e-6 * abs(a) for a, b in zip(total_lp, total_lp[1:]))) if total_lp else None,
            "transmat_rows_sum_to_1": rows_ok, "emission_rows_sum_to_1": em_ok,
            "deferred_checks": hmm._engine().ctx.stat("deferred_checks"), "deferred_refused_chunks": hmm._engine().ctx.stat("deferred_bad"),
            "repaired_chunks": {k: hmm._engine().ctx.stat("repaired_chunks_" + k) for k in ("forward", "backward")}}


def run_c4(scale):
    import torch
    import oracle as orc
    from tehmm_b200 import synth
    orc.build()
    m = synth.make_model(N=30, seed=0)
    lens = [max(50, int(n * scale)) for n in synth.bench_lengths("c4")]
    seqs = [sample_long(synth, m, n, seed=200 + 7 * i) for i, n in enumerate(lens)]
    hv, _ = make_hmm(m)
    hm, _ = make_hmm(m, algorithm="map")
    for _ in range(2):
        rv = hv.decode_batch(seqs)
        rm = hm.decode_batch(seqs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rv = hv.decode_batch(seqs)
    t1 = time.perf_counter()
    rm = hm.decode_batch(seqs)
    t2 = time.perf_counter()
    i21 = 20
    alone_v = hv.decode(seqs[i21])
    alone_m = hm.decode(seqs[i21])
    n0 = min(300_000, lens[0])
    ref = orc.sweep_sequence(seqs[0][:n0], m["table"], 1.0, m["log_start"], m["log_trans"])
    pre = hv.decode(seqs[0][:n0])
    return {"config": "c4", "sequences": len(seqs), "steps": int(sum(lens)), "longest": int(max(lens)),
            "states": 30, "tracks": 10, "viterbi_seconds_e2e": t1 - t0, "map_seconds_e2e": t2 - t1,
            "cells_per_s_e2e": sum(lens) * 30 / (t2 - t0),
            "batch_equals_alone_viterbi": bool(np.array_equal(rv[i21][1], alone_v[1])),
            "batch_vs_alone_viterbi_logprob_rel_err": float(abs(rv[i21][0] - alone_v[0]) / abs(alone_v[0])),
            "batch_equals_alone_map": bool(np.array_equal(rm[i21][1], alone_m[1])),
            "prefix_path_agreement_with_cpu_oracle": float(np.mean(pre[1] == ref["vit_states"])),
            "prefix_logprob_rel_err": float(abs(pre[0] - ref["vit_logprob"]) / abs(ref["vit_logprob"])),
            "all_int64": bool(all(s.dtype == np.int64 for _, s in rv))}


def run_c5(scale):
    import torch
    from tehmm_b200 import synth
    from tehmm_b200.engine import get_engine
    N = 50
    m = synth.make_model(N=N, seed=0)
    T = max(3_000_000, int(250_000_000 * scale))
    t0 = time.perf_counter()
    obs = sample_long(synth, m, T, seed=300)
    gen_s = time.perf_counter() - t0
    eng = get_engine(0)
    eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
    dev = eng.device
    # the stand-alone prefix first: the reference answer for the first 2 M steps, and every kernel of the full-size
    # run gets loaded (the first launch of a kernel includes ~40 ms of lazy module loading)
    eng.upload_batch([obs[:2_200_000]])
    _, st_pre = eng.viterbi()
    out = eng.posteriors(renorm_eps=True, want_post=False, want_map=True)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    d_obs = torch.from_numpy(obs.reshape(-1)).to(dev)
    eng.use_device_batch(d_obs, 1, np.array([0, T], dtype=np.int64))
    prec, tdt = eng._prec("f32")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    eng.ctx.set_option("timing", 1)        # kernel times by the library's own events (the stage times below include cudaMalloc)
    ev[0].record()
    elog, _, rm = eng.run_emission(prec, tdt, None, True, False)
    states, _, vlp = eng.run_viterbi(prec, elog, None, None, want64=False, rowmax=rm)
    ev[1].record()
    del elog
    # config 5 holds two 60 GiB lattices at a time on a 178 GiB device: give the cached blocks back, or the
    # allocator carves the 2 GB rowmax out of a cached lattice-sized block and a third block of that size does not fit
    # (expandable segments avoid that too, but mapping 60 GiB in 2 MB granules takes seconds)
    torch.cuda.empty_cache()
    _, blin, rowmax = eng.run_emission(prec, tdt, None, False, True)
    alpha, logprob = eng.run_forward(prec, tdt, blin, rowmax, None)
    ev[2].record()
    from tehmm_b200 import _lib
    _, mstates, mscore = eng.run_backward(prec, tdt, _lib.BWD_MAP | _lib.BWD_RENORM_EPS, blin, alpha, None)
    ev[3].record()
    torch.cuda.synchronize()
    peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    kernel_us = {k: eng.ctx.stat("us_" + k) for k in ("emission", "forward", "backward", "viterbi_dp", "traceback", "rescore")}
    eng.ctx.set_option("timing", 0)
    vit_ms, fwd_ms, bwd_ms = (ev[i].elapsed_time(ev[i + 1]) for i in range(3))
    st_full = states[:2_000_000].cpu().numpy().astype(np.int64)
    ms_full = mstates[:2_000_000].cpu().numpy().astype(np.int64)
    vlp, lp = float(vlp[0].item()), float(logprob[0].item())
    repairs = {k: eng.ctx.stat("repaired_chunks_" + k) for k in ("forward", "backward", "viterbi", "traceback")}
    del alpha, blin, rowmax, states, mstates, d_obs
    torch.cuda.empty_cache()
    return {"config": "c5", "steps": T, "states": N, "tracks": 10, "host_generation_seconds": gen_s,
            "viterbi_ms": vit_ms, "forward_ms": fwd_ms, "backward_map_ms": bwd_ms, "kernel_us": kernel_us,
            "umma_passes": eng.ctx.stat("umma_passes"),
            "cells_per_s_sweep": T * N / ((vit_ms + fwd_ms + bwd_ms) * 1e-3),
            # the stage brackets include cudaFree / cudaMalloc of the 60 GiB lattices (two fit at a time); kernels only:
            "cells_per_s_kernels": T * N / ((2 * kernel_us["emission"] + sum(kernel_us[k] for k in ("forward", "backward", "viterbi_dp", "traceback", "rescore"))) * 1e-6),
            "peak_device_memory_gb": peak_gb, "logprob": lp, "viterbi_logprob": vlp,
            "viterbi_le_logprob": bool(vlp <= lp), "repaired_chunks": repairs,
            "first_2M_equals_standalone_prefix_viterbi": float(np.mean(st_full == st_pre[0][:2_000_000])),
            "first_2M_equals_standalone_prefix_map": float(np.mean(ms_full == out["map_states"][0][:2_000_000]))}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
        args = [a for a in args if a != sys.argv[sys.argv.index("--scale") + 1]]
    which = args or ["c3", "c4", "c5"]
    for name in which:
        t0 = time.perf_counter()
        try:
            line = {"c3": run_c3, "c4": run_c4, "c5": run_c5}[name](scale)
        except Exception as e:   # keep going: one JSON line per config either way
            import traceback
            traceback.print_exc()
            line = {"config": name, "error": repr(e)}
        line["scale"] = scale
        line["wall_seconds_incl_generation"] = time.perf_counter() - t0
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
