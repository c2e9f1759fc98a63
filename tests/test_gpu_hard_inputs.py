"""-m gpu: slow-mixing input, where speculation does not pay.

All-missing stretches of 2 000 - 20 000 steps (uniform emission) under the bench model and under a
near-reducible transition matrix (diagonal 0.9999): the speculated chunk boundaries fail, and with
the plain repair loop the truth travels one chunk per pass (probe: 91 passes).  The engine switches
to the exact resolution of csrc/fallback.cu (transfer operators of the flagged chunks, a float64
chain, one re-run), so the number of passes is bounded, and the results must still be the
reference's."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from parity import ATOL, TOL, assert_map_near_ties_only, assert_near_ties_only, oracle_all

pytestmark = pytest.mark.gpu

PASS_STATS = ("repair_passes_forward", "repair_passes_backward", "repair_passes_viterbi", "repair_passes_traceback")


def _run(eng, prec):
    ctx = eng.ctx
    before = {k: ctx.stat(k) for k in PASS_STATS + ("fallbacks",)}
    out = eng.posteriors(renorm_eps=False, want_post=True, want_map=True, precision=prec)
    lps, states = eng.viterbi(precision=prec)
    delta = {k: ctx.stat(k) - v for k, v in before.items()}
    return out, lps, states, delta


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("sticky", [0.9, 0.9999])
def test_gaps_bounded_passes_and_exact(oracle, prec, sticky):
    from tehmm_b200 import synth
    from tehmm_b200.engine import get_engine
    m = synth.make_model(N=30, seed=0, sticky=sticky)
    T = 300_000
    obs, _ = synth.sample_obs(m, T, seed=3)
    obs, spans = synth.add_missing_stretches(obs, n_stretches=4, lo=2_000, hi=20_000, seed=6)
    eng = get_engine(0)
    for k in ("chunk_tiles", "warmup", "fine_len"):
        eng.ctx.set_option(k, 0)
    eng.ctx.set_option("tile", 1)
    eng.ctx.set_option("fallback_after", 2)
    eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
    eng.upload_batch([obs])
    out, lps, states, d = _run(eng, prec)
    print("sticky %g %s: passes %s" % (sticky, prec, d))
    # bounded: the ordinary passes, the resolution, its re-run, and for the traceback a few rounds of growing reach
    assert d["fallbacks"] > 0
    for k in PASS_STATS[:3]:
        assert d[k] <= 6, (k, d)
    assert d["repair_passes_traceback"] <= 10, d
    ref = oracle_all(oracle, obs, m["table"], 1.0, m["log_start"], m["log_trans"])
    if prec == "f64":
        assert out["logprob"][0] == pytest.approx(ref["logprob"], rel=1e-10)
        # (the reference keeps its lattices in log space: at |log alpha| ~ 5e6 an ulp is 1e-9, so ITS
        #  posteriors carry ~1e-8 relative rounding at this length)
        #  under diagonal 0.9999 that noise is carried for ~1e4 steps: 2e-6 measured, while our two float64
        #  routes -- plain loop and exact resolution -- agree to 1e-8, test_plain_loop_agrees)
        assert_allclose(out["post"][0], ref["post"], rtol=1e-6 if sticky < 0.99 else 1e-5, atol=1e-12)
        assert_array_equal(states[0], ref["vit_states"])
        assert lps[0] == pytest.approx(ref["vit_logprob"], rel=1e-10)
    else:
        assert out["logprob"][0] == pytest.approx(ref["logprob"], rel=TOL[prec])
        # fp32 under slow mixing: rounding noise is amplified by 1 / (1 - |lambda_2|),
        # 19x for the bench model, 10^4 for diagonal 0.9999; posteriors are compared accordingly
        rtol, atol = (5e-4, 1e-5) if sticky < 0.99 else (0.1, 1e-3)
        assert_allclose(out["post"][0], ref["post"], rtol=rtol, atol=atol)
        assert_map_near_ties_only(out["map_states"][0], ref["post"], rel=rtol, label="gaps")
        assert_near_ties_only(states[0], ref["vit_states"], ref["frame"], m["log_start"], m["log_trans"], None, label="gaps")
        assert lps[0] == pytest.approx(ref["vit_logprob"], rel=1e-6)


def test_plain_loop_agrees(oracle):
    """fallback_after = -1 (the plain repair loop, one link per pass) gives the same float64 results"""
    from tehmm_b200 import synth
    from tehmm_b200.engine import get_engine
    m = synth.make_model(N=30, seed=0, sticky=0.9999)
    obs, _ = synth.sample_obs(m, 60_000, seed=4)
    obs, _ = synth.add_missing_stretches(obs, n_stretches=2, lo=2_000, hi=6_000, seed=7)
    eng = get_engine(0)
    eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
    eng.upload_batch([obs])
    res = {}
    try:
        for fb in (-1, 2):
            eng.ctx.set_option("fallback_after", fb)
            res[fb] = _run(eng, "f64")
    finally:
        eng.ctx.set_option("fallback_after", 2)
    assert res[2][3]["fallbacks"] > 0 and res[-1][3]["fallbacks"] == 0
    assert_array_equal(res[2][2][0], res[-1][2][0])
    assert_allclose(res[2][0]["post"][0], res[-1][0]["post"][0], rtol=1e-8, atol=1e-13)
    assert res[2][0]["logprob"][0] == pytest.approx(res[-1][0]["logprob"][0], rel=1e-12)


def test_gaps_on_a_50_state_model(oracle):
    """The same input class on the 33..64-state kernels (tcgen05 forward / backward, two-warps-per-chunk Viterbi DP,
    two-chunks-per-warp traceback): all-missing stretches longer than any warm-up, so the first passes of those
    kernels hand flagged chunks to the repair kernels and to the exact resolution (transfer operators of 64 x 64)."""
    from tehmm_b200 import synth
    from tehmm_b200.engine import get_engine
    m = synth.make_model(N=50, seed=2, sticky=0.9)
    T = 120_000
    obs, _ = synth.sample_obs(m, T, seed=4)
    obs, spans = synth.add_missing_stretches(obs, n_stretches=3, lo=2_000, hi=9_000, seed=7)
    eng = get_engine(0)
    for k in ("chunk_tiles", "warmup", "fine_len"):
        eng.ctx.set_option(k, 0)
    eng.ctx.set_option("tile", 1)
    eng.ctx.set_option("fallback_after", 2)
    eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
    eng.upload_batch([obs])
    before = eng.ctx.stat("umma_passes")
    out, lps, states, d = _run(eng, "f32")
    print("50 states: passes %s" % (d,))
    # (the forward pass runs on tcgen05; its flagged chunks raise the warm-up beyond this batch's 64-step fine chunks,
    #  so the backward pass of THIS batch takes the one-chunk-per-warp kernel: tehmm_backward_umma_ok)
    assert eng.ctx.stat("umma_passes") >= before + 1
    for k in PASS_STATS[:3]:
        assert d[k] <= 6, (k, d)
    assert d["repair_passes_traceback"] <= 10, d
    ref = oracle_all(oracle, obs, m["table"], 1.0, m["log_start"], m["log_trans"])
    assert out["logprob"][0] == pytest.approx(ref["logprob"], rel=TOL["f32"])
    assert_allclose(out["post"][0], ref["post"], rtol=5e-4, atol=1e-5)
    assert_map_near_ties_only(out["map_states"][0], ref["post"], rel=5e-4, label="gaps, 50 states")
    assert_near_ties_only(states[0], ref["vit_states"], ref["frame"], m["log_start"], m["log_trans"], None, label="gaps, 50 states")
    assert lps[0] == pytest.approx(ref["vit_logprob"], rel=1e-6)
