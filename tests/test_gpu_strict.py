"""-m gpu: the L0 drop-ins (tehmm_b200._hmm / _emission, strict float64 CUDA
kernels through the C ABI) against the reference's golden vectors and the oracle.

Bar: bit-exact for everything made of IEEE adds / multiplies / compares
(emission sums, Viterbi paths and scores, histograms, supervised counts);
<= 1e-10 relative for forward / backward / lneta (CUDA libdevice exp/log vs glibc).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import golden, golden_names, ratios_of

pytestmark = pytest.mark.gpu

LL_CASES = golden_names("ll_")
RTOL = 1e-10


def close_with_inf(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert_array_equal(np.isinf(a), np.isinf(b))
    fin = np.isfinite(b)
    assert_allclose(a[fin], b[fin], rtol=rtol, atol=1e-300)


@pytest.mark.parametrize("name", LL_CASES)
def test_golden_low_level(name):
    from tehmm_b200 import _emission, _hmm
    g = golden(name)
    N = int(g["N"])
    obs, table = g["obs"], g["table"]
    T = obs.shape[0]
    r = ratios_of(g)
    frame = np.zeros((T, N))
    _emission.fastAllLogProbs(obs, table, frame, float(g["normalize"]), r)
    assert_array_equal(frame, g["frame"])                       # bit-exact
    fwd = np.zeros((T, N))
    _hmm._forward(T, N, g["log_start"], g["log_trans"], g["frame"], r, fwd)
    close_with_inf(fwd, g["fwd"])
    bwd = np.zeros((T, N))
    _hmm._backward(T, N, g["log_start"], g["log_trans"], g["frame"], r, bwd)
    close_with_inf(bwd, g["bwd"])
    states, lp = _hmm._viterbi(T, N, g["log_start"], g["log_trans"], r, g["frame"])
    assert states.dtype == np.int64
    assert_array_equal(states, g["vit_states"])                 # bit-exact
    assert lp == float(g["vit_logprob"])
    if T > 1:
        lneta = np.zeros((N, N))
        _hmm._log_sum_lneta(T, N, g["fwd"], g["log_trans"], g["bwd"], g["frame"],
                            float(g["logprob"]), r, lneta)
        close_with_inf(lneta, g["lneta"])
    stats = np.zeros_like(g["obs_stats"])
    _emission.fastAccumulateStats(obs, stats, g["post"], r)
    assert_array_equal(stats, g["obs_stats"])                   # bit-exact (same order)


def test_impossible_rows_quirk():
    from tehmm_b200 import _emission
    g = golden("impossible_rows")
    for key in "abcd":
        obs = g["obs_" + key]
        out = np.full((obs.shape[0], 2), 7.0)
        _emission.fastAllLogProbs(obs, g["table"], out, 1.0, None)
        assert_array_equal(out, g["out_" + key])


def test_dpbenchmark_frames():
    from tehmm_b200 import _hmm
    g = golden("dpbench")
    frame = g["frame"]
    T, N = frame.shape
    for tag, r in (("", None), ("seg_", g["ratios"])):
        fwd = np.zeros((T, N))
        bwd = np.zeros((T, N))
        _hmm._forward(T, N, g["log_start"], g["log_trans"], frame, r, fwd)
        _hmm._backward(T, N, g["log_start"], g["log_trans"], frame, r, bwd)
        close_with_inf(fwd, g[tag + "fwd"])
        close_with_inf(bwd, g[tag + "bwd"])
        states, lp = _hmm._viterbi(T, N, g["log_start"], g["log_trans"], r, frame)
        assert_array_equal(states, g[tag + "vit_states"])
        assert lp == float(g[tag + "vit_logprob"])


def test_supervised_counts():
    from tehmm_b200 import _emission
    from tehmm_b200.track import IntegerTrackTable
    g = golden("counts")
    N = int(g["N"])
    syms = g["syms"]
    S = int(syms.max()) + 1
    obs = g["obs"]
    tab = IntegerTrackTable(obs.shape[1], "chrG", 0, obs.shape[0])
    tab.data[:] = obs
    for key, r in (("stats", None), ("stats_ratio", g["ratios"])):
        stats = np.zeros((len(syms), N, S))
        for k, s in enumerate(syms):
            stats[k, :, :s + 1] += 1.0
        for a, b, st in g["intervals"]:
            _emission.fastUpdateCounts(("chrG", int(a), int(b), int(st)), tab, stats, r)
        assert_array_equal(stats, g[key])


def test_wikipedia_known_answers():
    """tests/hmmTest.py:48-135 through the L0 functions."""
    import math
    from tehmm_b200 import _emission, _hmm
    g = golden("wikipedia")
    log_start, log_trans = np.log(g["startprob"]), np.log(g["transmat"])
    for key, scale in (("v1", 1.0), ("v3", 1.0), ("v4", 1e-3)):
        obs = g[key + "_obs"].astype(np.int32)
        T = obs.shape[0]
        frame = np.zeros((T, 2))
        _emission.fastAllLogProbs(obs, g[key + "_table"], frame, 1.0, None)
        assert_array_equal(frame, g[key + "_frame"])
        states, lp = _hmm._viterbi(T, 2, log_start, log_trans, None, frame)
        assert math.isclose(math.exp(lp), 0.01344 * scale, rel_tol=1e-12)
        assert_array_equal(states, [1, 0, 0])


def test_strict_vs_oracle_larger(oracle):
    """T = 3000, N = 30, K = 10, with ratios and uint16 symbols: same bar."""
    from tehmm_b200 import _emission, _hmm, synth
    m = synth.make_model(N=30, seed=11)
    obs, _ = synth.sample_obs(m, 3000, seed=12, dtype=np.uint16)
    rng = np.random.RandomState(5)
    ratios = rng.uniform(0.01, 10.0, size=3000)
    T, N = 3000, 30
    for r in (None, ratios):
        f_ref, f_gpu = np.zeros((T, N)), np.zeros((T, N))
        oracle.fastAllLogProbs(obs, m["table"], f_ref, 1.0, r)
        _emission.fastAllLogProbs(obs, m["table"], f_gpu, 1.0, r)
        assert_array_equal(f_gpu, f_ref)
        a_ref, a_gpu = np.zeros((T, N)), np.zeros((T, N))
        oracle._forward(T, N, m["log_start"], m["log_trans"], f_ref, r, a_ref)
        _hmm._forward(T, N, m["log_start"], m["log_trans"], f_ref, r, a_gpu)
        close_with_inf(a_gpu, a_ref)
        b_ref, b_gpu = np.zeros((T, N)), np.zeros((T, N))
        oracle._backward(T, N, m["log_start"], m["log_trans"], f_ref, r, b_ref)
        _hmm._backward(T, N, m["log_start"], m["log_trans"], f_ref, r, b_gpu)
        close_with_inf(b_gpu, b_ref)
        s_ref, lp_ref = oracle._viterbi(T, N, m["log_start"], m["log_trans"], r, f_ref)
        s_gpu, lp_gpu = _hmm._viterbi(T, N, m["log_start"], m["log_trans"], r, f_ref)
        assert_array_equal(s_gpu, s_ref)
        assert lp_gpu == lp_ref
        lp = oracle.logsumexp(a_ref[-1])
        l_ref, l_gpu = np.zeros((N, N)), np.zeros((N, N))
        oracle._log_sum_lneta(T, N, a_ref, m["log_trans"], b_ref, f_ref, lp, r, l_ref)
        _hmm._log_sum_lneta(T, N, a_ref, m["log_trans"], b_ref, f_ref, lp, r, l_gpu)
        close_with_inf(l_gpu, l_ref)


def test_argument_errors():
    from tehmm_b200 import _emission, _hmm
    with pytest.raises(AssertionError):
        _emission.fastAllLogProbs(np.zeros((3, 2), dtype=np.uint8), np.zeros((3, 2, 4)), np.zeros((3, 2)), 1.0, None)
    with pytest.raises(ValueError):
        _hmm._forward(2, 2, np.zeros(2), np.zeros((2, 2)), np.zeros((2, 2)), None, np.zeros((2, 2), dtype=np.float32))
    assert _emission.canFast(np.zeros((1, 1), dtype=np.uint8))
    assert not _emission.canFast(np.zeros((1, 1), dtype=np.int64))
