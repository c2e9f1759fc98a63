"""-m gpu: BASELINE.json configs[2..4] in SHAPE (scaled so that the CPU oracle finishes in
seconds), through the class API, against the oracle on every sequence:

  C3  350 ragged sequences WITH segment ratios (segmentTracks-style tables: segOffsets +
      effectiveSegmentLength), three Baum-Welch iterations -- hmm.py:545-616 via
      oracle.estep_sequence per sequence -- re-estimated parameters at north_star's tolerance;
  C4  24 sequences in hg19's chromosome proportions, Viterbi + MAP decode of the whole batch;
  C5  ONE sequence of 1.2 M steps, 50 states: decode / score, prefix against the oracle.

The full-size runs of the same configs are tools/run_configs.py (size-independent checks).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from parity import (ATOL, TOL, assert_map_near_ties_only, assert_near_ties_only, oracle_all)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _defaults():
    from tehmm_b200 import engine
    from tehmm_b200._lib import get_context
    ctx = get_context(0)
    for k in ("chunk_tiles", "warmup", "fine_len"):
        ctx.set_option(k, 0)
    ctx.set_option("tile", 1)
    engine.set_precision("f32")
    yield
    engine.set_precision("f32")


def make_hmm(m, seg_len=None, **kw):
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    em = IndependentMultinomialEmissionModel(m["N"], list(m["syms"]), zeroAsMissingData=True, fudge=0.0,
                                             effectiveSegmentLength=seg_len)
    em.logProbs = m["table"].copy()
    return MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy(), **kw), em


def segmented_table(obs, seg_lens):
    """an IntegerTrackTable as bin/segmentTracks.py + TrackTable.segment leave it (track.py:449-513)"""
    from tehmm_b200.track import IntegerTrackTable
    t = IntegerTrackTable(obs.shape[1], "chrG", 0, int(seg_lens.sum()))
    t.segOffsets = np.concatenate([[0], np.cumsum(seg_lens)[:-1]]).astype(np.int64)
    t.data = obs.copy()
    t.shape = (len(t), obs.shape[1])
    return t


def c3_inputs():
    from tehmm_b200 import synth
    m = synth.make_model(N=30, seed=0)
    lens = [max(20, n // 30) for n in synth.bench_lengths("c3")]        # 350 sequences, ~117 k steps
    rng = np.random.RandomState(3)
    tables = []
    for i, n in enumerate(lens):
        obs = synth.sample_obs(m, n, seed=100 + i)[0]
        seg = np.minimum(rng.geometric(1.0 / 60.0, size=n), 100).astype(np.int64)   # maxLen-100 segments
        tables.append(segmented_table(obs, seg))
    m0 = synth.make_model(N=30, seed=7, zero_frac=0.0)                  # start EM away from the truth
    return m0, tables


@pytest.fixture(scope="module")
def c3_oracle_fit(oracle):
    """three EM iterations driven by the oracle's per-sequence E-step (the reference flow)"""
    from test_host_logic import oracle_estep
    m0, tables = c3_inputs()
    hmm, em = make_hmm(m0, seg_len=100, n_iter=3, thresh=0.0, fixStart=False, transMatEpsilons=True)
    hmm._device_estep = lambda obs, stats, params, n_total, slots: oracle_estep(oracle)(
        hmm, obs, stats, params, n_total, slots)
    hmm.fit(tables)
    return dict(transmat=hmm.transmat_.copy(), startprob=hmm.startprob_.copy(),
                table=em.getLogProbs().copy(), last=hmm.getLastLogProb(), iters=hmm.current_iteration)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_c3_shaped_baum_welch_with_segment_ratios(c3_oracle_fit, prec):
    from tehmm_b200 import engine
    engine.set_precision(prec)
    m0, tables = c3_inputs()
    assert len(tables) == 350 and all(t.getSegmentOffsets() is not None for t in tables)
    hmm, em = make_hmm(m0, seg_len=100, n_iter=3, thresh=0.0, fixStart=False, transMatEpsilons=True)
    ratios = em.getSegmentRatios(tables[0])
    assert ratios is not None and ratios.max() > 0.5 and ratios.min() < 0.1
    hmm.fit(tables)
    ref = c3_oracle_fit
    assert hmm.current_iteration == ref["iters"]
    assert hmm.getLastLogProb() == pytest.approx(ref["last"], rel=TOL[prec])
    # float64: the reference's log-space lattices lose ulp(|log alpha|) (see test_estep_matches_oracle)
    rt = TOL[prec] if prec == "f32" else 1e-9
    assert_allclose(hmm.transmat_, ref["transmat"], rtol=rt, atol=ATOL[prec])
    assert_allclose(hmm.startprob_, ref["startprob"], rtol=rt, atol=ATOL[prec])
    assert_allclose(np.exp(em.getLogProbs()), np.exp(ref["table"]), rtol=rt, atol=ATOL[prec])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_c4_shaped_genome_decode(oracle, prec):
    from tehmm_b200 import engine, synth
    engine.set_precision(prec)
    m = synth.make_model(N=30, seed=0)
    lens = [-(-n // 100) for n in synth.bench_lengths("c4")]            # 24 sequences, chr1 = 9 971 bins
    assert len(lens) == 24 and max(lens) == lens[0]
    seqs = [synth.sample_obs(m, n, seed=200 + 7 * i)[0] for i, n in enumerate(lens)]
    hv, _ = make_hmm(m)
    hm, _ = make_hmm(m, algorithm="map")
    rv = hv.decode_batch(seqs)
    rm = hm.decode_batch(seqs)
    ss = hv.score_samples_batch(seqs[20:])                               # chr21, 22, X, Y
    for i, obs in enumerate(seqs):
        ref = oracle_all(oracle, obs, m["table"], 1.0, m["log_start"], m["log_trans"])
        post = oracle.posteriors(ref["fwd"], ref["bwd"], renorm_eps=True)
        lp, st = rv[i]
        sc, ms = rm[i]
        assert st.dtype == np.int64 and st.shape == (lens[i],)
        assert lp == pytest.approx(ref["vit_logprob"], rel=1e-6 if prec == "f32" else 1e-10)
        assert sc == pytest.approx(np.max(post, axis=1).sum(), rel=TOL[prec])
        if prec == "f64":
            assert_array_equal(st, ref["vit_states"])
            assert_array_equal(ms, np.argmax(post, axis=1))
        else:
            assert_near_ties_only(st, ref["vit_states"], ref["frame"], m["log_start"], m["log_trans"],
                                  label="c4[%d]" % i)
            assert_map_near_ties_only(ms, post, label="c4[%d]" % i)
        if i >= 20:
            lp2, p2 = ss[i - 20]
            assert lp2 == pytest.approx(ref["logprob"], rel=TOL[prec])
            assert_allclose(p2, post, rtol=TOL[prec], atol=ATOL[prec])


def test_c5_shaped_single_long_sequence_50_states(oracle):
    from tehmm_b200 import synth
    N = 50
    m = synth.make_model(N=N, seed=0)
    T = 1_200_000
    obs = synth.sample_obs(m, T, seed=300)[0]
    hv, _ = make_hmm(m)
    hm, _ = make_hmm(m, algorithm="map")
    lp, st = hv.decode(obs)
    sc, ms = hm.decode(obs)
    ll = hv.score(obs)
    assert st.shape == (T,) and st.min() >= 0 and st.max() < N
    assert lp <= ll < 0 and 0.5 * T < sc <= T * (1 + 1e-6)
    eng = hv._engine()
    # prefix against the oracle: paths coalesce and the filters forget, so the first 50 k steps of the
    # 1.2 M decode must be those of the oracle's stand-alone 60 k decode
    n0, n1 = 60_000, 50_000
    ref = oracle_all(oracle, obs[:n0], m["table"], 1.0, m["log_start"], m["log_trans"])
    assert_near_ties_only(st[:n1], ref["vit_states"][:n1], ref["frame"][:n1], m["log_start"], m["log_trans"],
                          label="c5 prefix")
    post = oracle.posteriors(ref["fwd"], ref["bwd"], renorm_eps=True)
    assert_map_near_ties_only(ms[:n1], post[:n1], label="c5 prefix")
    assert hv.score(obs[:n0]) == pytest.approx(ref["logprob"], rel=TOL["f32"])
    lp0, st0 = hv.decode(obs[:n0])
    assert lp0 == pytest.approx(ref["vit_logprob"], rel=1e-6)
    assert_near_ties_only(st0, ref["vit_states"], ref["frame"], m["log_start"], m["log_trans"])
    # E-step conservation laws at this shape (hmm.py:545-574; 1/N: _hmm.pyx:179)
    eng.upload_batch([obs])
    stt = eng.estep()
    assert stt["logprob"] == pytest.approx(ll, rel=1e-9)
    assert stt["obs"].sum() == pytest.approx(T * m["K"], rel=1e-6)
    assert stt["trans"].sum() * N == pytest.approx(T - 1, rel=1e-6)
    assert np.all(stt["trans"][m["A"] == 0] == 0)


def test_viterbi_with_segment_ratios_on_the_lean_kernels(oracle):
    """decode of segmented tables (hmm.py:674: the DP sees the segment ratios, basehmm.py:327: the emission does
    not): the fp32 lean DP and the four-chunks-per-warp traceback take the ratios, including the from-state-0
    rule of _hmm.pyx:234-237; paths equal the oracle's except at near-ties, float64 (generic kernels) exactly"""
    import time
    import torch
    from parity import assert_near_ties_only, oracle_frame
    from tehmm_b200 import synth
    from tehmm_b200.engine import get_engine
    m = synth.make_model(N=30, seed=5)
    lens = [150_000, 1, 2, 40_001]
    rng = np.random.RandomState(4)
    obs = [synth.sample_obs(m, n, seed=60 + i)[0] for i, n in enumerate(lens)]
    ratios = [np.minimum(rng.geometric(1.0 / 60.0, size=n), 300) / 100.0 for n in lens]      # some > 1, most < 1
    assert max(r.max() for r in ratios) > 1.5
    eng = get_engine(0)
    eng.upload_model(m["log_start"], m["log_trans"], m["table"], 1.0, m["widths"])
    eng.upload_batch(obs)
    out = {}
    for prec in ("f64", "f32"):
        out[prec] = eng.viterbi(ratios_em=None, ratios_dp=ratios, precision=prec)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.viterbi(ratios_em=None, ratios_dp=ratios, precision="f32")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    eng.viterbi(precision="f32")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("viterbi fp32 with ratios %.2f ms, without %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    for i, (o, r) in enumerate(zip(obs, ratios)):
        frame = oracle_frame(oracle, o, m["table"], 1.0, None)
        want, wlp = oracle._viterbi(o.shape[0], 30, m["log_start"], m["log_trans"], r, frame)
        assert_array_equal(out["f64"][1][i], want)
        assert out["f64"][0][i] == pytest.approx(wlp, rel=1e-10)
        assert_near_ties_only(out["f32"][1][i], want, frame, m["log_start"], m["log_trans"], r, label="ratios seq %d" % i)
        assert out["f32"][0][i] == pytest.approx(wlp, rel=1e-6)
