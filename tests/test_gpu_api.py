"""-m gpu: the class-level drop-in (MultitrackHmm / IndependentMultinomial-
EmissionModel) against golden outputs of the reference's own classes
(tests/golden/make_golden.py) and the known answers of tests/hmmTest.py.
"""
import copy
import math
import pickle

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import golden
from parity import ATOL, TOL, assert_near_ties_only, oracle_frame

pytestmark = pytest.mark.gpu

EMISSIONPROB = [[0.1, 0.4, 0.5], [0.6, 0.3, 0.1]]
STARTPROB = [0.6, 0.4]
TRANSMAT = [[0.7, 0.3], [0.4, 0.6]]


@pytest.fixture(autouse=True)
def _defaults():
    from tehmm_b200 import engine
    from tehmm_b200._lib import get_context
    ctx = get_context(0)
    ctx.set_option("chunk_tiles", 0)
    ctx.set_option("warmup", 0)
    engine.set_precision("f32")
    yield
    engine.set_precision("f32")


def make_hmm(syms, params, **kw):
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    em = IndependentMultinomialEmissionModel(2, syms, params, zeroAsMissingData=False)
    return MultitrackHmm(em, startprob=STARTPROB, transmat=TRANSMAT, **kw), em


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_wikipedia_example(prec):
    """tests/hmmTest.py:48-135."""
    from tehmm_b200 import engine
    from tehmm_b200.track import IntegerTrackTable
    engine.set_precision(prec)
    g = golden("wikipedia")
    hmm, em = make_hmm([3], [EMISSIONPROB])
    obs = np.asarray([[0], [1], [2]])
    assert_array_equal(hmm._compute_log_likelihood(obs), g["v1_frame"])
    # fp32: the log-probability is the fp32 DP's own value (viterbi_lean_kernel), contract 1e-5
    rel = 1e-9 if prec == "f64" else 1e-6
    lp, st = hmm.decode(obs)
    assert math.exp(lp) == pytest.approx(0.01344, rel=rel)
    assert_array_equal(st, [1, 0, 0])
    assert st.dtype == np.int64
    hmm3, _ = make_hmm([3, 1, 1], [EMISSIONPROB, [[1.], [1.]], [[1.], [1.]]])
    lp, st = hmm3.decode(np.asarray([[0, 0, 0], [1, 0, 0], [2, 0, 0]]))
    assert math.exp(lp) == pytest.approx(0.01344, rel=rel)
    assert_array_equal(st, [1, 0, 0])
    hmm4, _ = make_hmm([3, 1, 1, 10], [EMISSIONPROB, [[1.], [1.]], [[1.], [1.]], [[.1] * 10, [.1] * 10]])
    obs4 = np.asarray([[0, 0, 0, 0], [1, 0, 0, 5], [2, 0, 0, 7]])
    lp, st = hmm4.decode(obs4)
    assert math.exp(lp) == pytest.approx(0.01344 * 1e-3, rel=rel)
    assert_array_equal(st, [1, 0, 0])
    table4 = IntegerTrackTable(4, "scaffold_1", 10, 13)
    for row in range(4):
        table4.writeRow(row, [obs4[0][row], obs4[1][row], obs4[2][row]])
    lp, st = hmm4.decode(table4)
    assert math.exp(lp) == pytest.approx(0.01344 * 1e-3, rel=rel)
    assert_array_equal(st, [1, 0, 0])
    # posteriors (hmmTest.py:147-151 values hold for the eps-renormalised score_samples)
    sc, post = hmm.score_samples(obs)
    assert sc == pytest.approx(float(g["v1_score"]), rel=1e-6 if prec == "f32" else 1e-12)
    assert_allclose(post, g["v1_post"], rtol=1e-5 if prec == "f32" else 1e-10)
    assert post.dtype == np.float64
    assert hmm.score(obs) == pytest.approx(float(g["v1_score"]), rel=1e-6)
    assert_array_equal(hmm.predict(obs), [1, 0, 0])
    # decoder precedence (basehmm.py:389-392): the constructor's algorithm wins
    assert_array_equal(hmm.decode(obs, algorithm="map")[1], [1, 0, 0])
    hmap, _ = make_hmm([3], [EMISSIONPROB], algorithm="map")
    sc, st = hmap.decode(obs)
    assert_array_equal(st, np.argmax(g["v1_post"], axis=1))
    assert sc == pytest.approx(np.max(g["v1_post"], axis=1).sum(), rel=1e-5)


def load_fit_case(name):
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    from tehmm_b200.track import IntegerTrackTable
    g = golden(name)
    N, syms = int(g["N"]), [int(s) for s in g["syms"]]
    seg = int(g["seg"]) == 1
    em = IndependentMultinomialEmissionModel(N, syms, zeroAsMissingData=True, fudge=0.0,
                                             effectiveSegmentLength=100 if seg else None)
    em.logProbs = g["init_table"].copy()
    hmm = MultitrackHmm(em, startprob=g["init_pi"].copy(), transmat=g["init_A"].copy(),
                        n_iter=int(g["n_iter"]), thresh=0.0, fixStart=False, transMatEpsilons=True)
    tables = []
    for i in range(int(g["nseq"])):
        o = g["obs_%d" % i]
        if seg:
            lens = g["seglens_%d" % i]
            t = IntegerTrackTable(o.shape[1], "chrG", 0, int(lens.sum()))
            t.segOffsets = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
            t.data = o.copy()
            t.shape = (len(t), o.shape[1])
        else:
            t = IntegerTrackTable(o.shape[1], "chrG", 0, o.shape[0])
            t.data[:] = o
        tables.append(t)
    return g, hmm, em, tables


@pytest.mark.parametrize("name", ["fit_n4_k3", "fit_n30_k10", "fit_n5_k2_seg"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_fit_matches_reference(oracle, name, prec):
    """Baum-Welch over several sequences and iterations == the reference's fit,
    then decode / score_samples with the fitted model."""
    from tehmm_b200 import engine
    engine.set_precision(prec)
    g, hmm, em, tables = load_fit_case(name)
    assert_allclose(hmm._log_transmat, g["init_log_trans"], rtol=1e-15)
    hmm.fit(tables)
    rt, at = TOL[prec], ATOL[prec]       # north_star: re-estimated parameters <= 1e-5 (fp32) / 1e-10 (fp64)
    assert hmm.current_iteration == int(g["fit_iterations"])
    assert hmm.getLastLogProb() == pytest.approx(float(g["fit_last_logprob"]), rel=1e-9 if prec == "f64" else 1e-5)
    assert_allclose(hmm.transmat_, g["fit_transmat"], rtol=rt, atol=at)
    assert_allclose(hmm.startprob_, g["fit_startprob"], rtol=rt, atol=at)
    assert_allclose(np.exp(em.getLogProbs()), np.exp(g["fit_table"]), rtol=rt, atol=at)
    # decode with the REFERENCE's fitted parameters so the comparison is not blurred by fit error
    hmm._log_transmat = g["fit_log_trans"].copy()
    hmm._log_startprob = g["fit_log_start"].copy()
    em.logProbs = g["fit_table"].copy()
    dec = hmm.decode_batch(tables)
    ss = hmm.score_samples_batch(tables)
    for i in range(len(tables)):
        lp, st = dec[i]
        assert lp == pytest.approx(float(g["vit_logprob_%d" % i]), rel=1e-6 if prec == "f32" else 1e-10)
        if prec == "f64":
            assert_array_equal(st, g["vit_states_%d" % i])
        else:
            # decode(): emission WITHOUT ratios, DP with ratios (basehmm.py:327, hmm.py:674)
            frame = oracle_frame(oracle, tables[i].data, g["fit_table"], 1.0, None)
            r_dp = em.getSegmentRatios(tables[i])
            assert_near_ties_only(st, g["vit_states_%d" % i], frame, g["fit_log_start"], g["fit_log_trans"],
                                  r_dp, label="%s[%d]" % (name, i))
        sc, post = ss[i]
        assert sc == pytest.approx(float(g["score_%d" % i]), rel=1e-5 if prec == "f32" else 1e-10)
        assert_allclose(post, g["post_%d" % i], rtol=1e-5 if prec == "f32" else 1e-9, atol=2e-6 if prec == "f32" else 1e-12)


def test_hmmtest_fit_param_subsets():
    """tests/hmmTest.py:195-250: the 7 params subsets, float observation arrays."""
    from tehmm_b200 import engine
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    engine.set_precision("f64")
    g = golden("hmmtest_fit")
    train3 = []
    for i in range(int(g["nseq"])):
        o = g["obs_%d" % i]
        o3 = np.zeros((len(o), 3), dtype=np.float64)
        o3[:, 0] = o
        train3.append(o3)
    for params in ["s", "t", "e", "st", "se", "te", "ste"]:
        em3 = IndependentMultinomialEmissionModel(2, [3, 1, 1], zeroAsMissingData=False)
        hmm3 = MultitrackHmm(em3, params=params, init_params=params.replace("e", ""))
        hmm3.transmat_ = [[0.5, 0.5], [0.5, 0.5]]
        hmm3.startprob_ = [0.5, 0.5]
        hmm3.fit(train3)
        assert_allclose(hmm3.transmat_, g[params + "_transmat"], rtol=1e-9, atol=1e-12)
        assert_allclose(hmm3.startprob_, g[params + "_startprob"], rtol=1e-9, atol=1e-12)
        assert_allclose(np.exp(em3.getLogProbs()), np.exp(g[params + "_table"]), rtol=1e-9, atol=1e-12)
        lp, st = hmm3.decode(train3[0])
        assert lp == pytest.approx(float(g[params + "_vit_logprob"]), rel=1e-10)
        assert_array_equal(st, g[params + "_vit_states"])


def test_model_is_picklable_and_copyable():
    """modelIO.py:26-32 / hmm.py:694: nothing CUDA-related may live on the model."""
    hmm, em = make_hmm([3], [EMISSIONPROB], maxProb=True)
    obs = [np.asarray([[0], [1], [2], [1], [0]]), np.asarray([[2], [2], [0]])]
    hmm.fit(obs)
    assert hmm.bestCopy is not None
    clone = pickle.loads(pickle.dumps(hmm))
    again = copy.deepcopy(hmm)
    for other in (clone, again):
        assert_array_equal(other.transmat_, hmm.transmat_)
        lp, st = other.decode(obs[0])
        lp0, st0 = hmm.decode(obs[0])
        assert lp == lp0
        assert_array_equal(st, st0)


def test_emission_model_api():
    """tests/emissionTest.py:61-105 on the drop-in class."""
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    em = IndependentMultinomialEmissionModel(numStates=2, numSymbolsPerTrack=[2])
    em.initParams([[[0.2, 0.8], [0.5, 0.5]]])
    assert em.singleLogProb(0, [1]) == math.log(0.2)
    assert em.singleLogProb(1, [0]) == 0
    truth = np.array([[math.log(0.2), math.log(0.5)], [math.log(0.2), math.log(0.5)],
                      [math.log(0.8), math.log(0.5)]])
    assert np.array_equal(em.allLogProbs(np.array([[1], [1], [2]])), truth)
    stats = em.initStats()
    assert stats[0].shape == (2, 3)
    em.accumulateStats(np.array([[0], [0], [1]]), stats, np.array([[0.01, 0.02], [0.01, 0.02], [0.3, 0.4]]))
    assert stats[0][0][0] == 0.01 + 0.01
    assert stats[0][1][0] == 0.02 + 0.02
    assert stats[0][0][1] == 0.3
    assert stats[0][1][1] == 0.4


def test_more_than_64_states_takes_the_strict_kernels(oracle):
    """The batched kernels stop at 64 states; wider models run the reference's per-sequence flow on
    the strict float64 kernels (MultitrackHmm._wide_*).  decode / score / score_samples / fit
    against the oracle for a 70-state model."""
    from test_host_logic import oracle_estep
    from tehmm_b200 import synth
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    N, syms = 70, (5, 3, 9)
    m = synth.make_model(N=N, syms=syms, seed=13)
    seqs = [synth.sample_obs(m, T, seed=20 + i)[0] for i, T in enumerate([400, 1, 257])]

    def model(**kw):
        em = IndependentMultinomialEmissionModel(N, list(syms), zeroAsMissingData=True)
        em.logProbs = m["table"].copy()
        return MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy(), **kw), em

    hmm, _ = model()
    assert hmm._wide()
    for obs in seqs:
        ref = oracle.sweep_sequence(obs, m["table"], 1.0, m["log_start"], m["log_trans"])
        lp, st = hmm.decode(obs)
        assert_array_equal(st, ref["vit_states"])
        assert lp == pytest.approx(ref["vit_logprob"], rel=1e-12)
        assert hmm.score(obs) == pytest.approx(ref["logprob"], rel=1e-10)
        lp2, post = hmm.score_samples(obs)
        assert lp2 == pytest.approx(ref["logprob"], rel=1e-10)
        assert_allclose(post.sum(axis=1), 1.0, rtol=1e-12)
        sc, ms = model(algorithm="map")[0].decode(obs)
        assert_array_equal(ms, ref["map_states"])
        assert sc == pytest.approx(ref["map_score"], rel=1e-9)
    # two EM iterations: strict-kernel E-step == oracle E-step
    a, ema = model(n_iter=3, thresh=0.0)
    a.fit(seqs)
    b, emb = model(n_iter=3, thresh=0.0)
    b._device_estep = lambda obs, stats, params, n_total, slots: oracle_estep(oracle)(b, obs, stats, params, n_total, slots)
    b.fit(seqs)
    assert_allclose(a.transmat_, b.transmat_, rtol=1e-9, atol=1e-300)
    assert_allclose(a.startprob_, b.startprob_, rtol=1e-9, atol=1e-300)
    assert_allclose(ema.getLogProbs(), emb.getLogProbs(), rtol=1e-9, atol=1e-300)
    assert a.getLastLogProb() == pytest.approx(b.getLastLogProb(), rel=1e-10)
