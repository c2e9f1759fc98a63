"""CPU: host-side logic of the drop-in classes -- M-step, convergence rule,
log-prob bookkeeping, parameter setters, emission-model bookkeeping -- checked
against golden outputs of the reference's classes.  The device E-step is
replaced here by the CPU oracle (tests may do that; the product never does)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import golden


def oracle_estep(oracle):
    def _estep(self, obs, stats, params, n_total, slots):
        em = self.emissionModel
        out = np.zeros(n_total)
        for slot, o in zip(slots, obs):
            ratios = em.getSegmentRatios(o)
            arr = o.getNumPyArray() if hasattr(o, "getNumPyArray") else np.asarray(o)
            s0 = np.zeros(self.n_components)
            tr = np.zeros((self.n_components, self.n_components))
            ob = np.zeros_like(stats['obs'])
            out[slot] = oracle.estep_sequence(arr, em.getLogProbs(), em.normalizeFac, self._log_startprob,
                                              self._log_transmat, ratios, s0, tr, ob)
            stats['nobs'] += 1
            if 's' in params:
                stats['start'] += s0
            if 't' in params:
                stats['trans'] += tr
            if 'e' in params:
                stats['obs'] += ob
        return out
    return _estep


@pytest.mark.parametrize("name", ["fit_n4_k3", "fit_n30_k10", "fit_n5_k2_seg"])
def test_fit_host_loop_matches_reference(oracle, monkeypatch, name):
    from tehmm_b200.hmm import MultitrackHmm
    from test_gpu_api import load_fit_case
    monkeypatch.setattr(MultitrackHmm, "_device_estep", oracle_estep(oracle))
    g, hmm, em, tables = load_fit_case(name)
    assert_allclose(hmm._log_transmat, g["init_log_trans"], rtol=1e-15)
    assert_allclose(hmm._log_startprob, g["init_log_start"], rtol=1e-15)
    hmm.fit(tables)
    assert hmm.current_iteration == int(g["fit_iterations"])
    assert hmm.getLastLogProb() == pytest.approx(float(g["fit_last_logprob"]), rel=1e-12)
    assert_allclose(hmm.transmat_, g["fit_transmat"], rtol=1e-10, atol=1e-300)
    assert_allclose(hmm.startprob_, g["fit_startprob"], rtol=1e-10, atol=1e-300)
    assert_allclose(em.getLogProbs(), g["fit_table"], rtol=1e-10, atol=1e-300)


def test_hmmtest_fit_subsets_host_loop(oracle, monkeypatch):
    """tests/hmmTest.py:195-250 with the oracle E-step."""
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    monkeypatch.setattr(MultitrackHmm, "_device_estep", oracle_estep(oracle))
    g = golden("hmmtest_fit")
    train3 = []
    for i in range(int(g["nseq"])):
        o = g["obs_%d" % i]
        o3 = np.zeros((len(o), 3), dtype=np.int32)
        o3[:, 0] = o
        train3.append(o3)
    for params in ["s", "t", "e", "st", "se", "te", "ste"]:
        em3 = IndependentMultinomialEmissionModel(2, [3, 1, 1], zeroAsMissingData=False)
        hmm3 = MultitrackHmm(em3, params=params, init_params=params.replace("e", ""))
        hmm3.transmat_ = [[0.5, 0.5], [0.5, 0.5]]
        hmm3.startprob_ = [0.5, 0.5]
        hmm3.fit(train3)
        assert_allclose(hmm3.transmat_, g[params + "_transmat"], rtol=1e-12)
        assert_allclose(hmm3.startprob_, g[params + "_startprob"], rtol=1e-12)
        assert_allclose(em3.getLogProbs(), g[params + "_table"], rtol=1e-12)


def test_setters_and_constants():
    from tehmm_b200 import common
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    assert common.LOGZERO == -1e100 and common.ZEROLOGPROB == -1e200
    assert common.myLog(0.0) == -1e100
    assert common.myLog(0.0, logZeroVal=-1e6) == -1e6
    assert_array_equal(common.myLog(np.array([1.0, 0.0])), [0.0, -1e100])
    em = IndependentMultinomialEmissionModel(3, [2, 4])
    assert em.getLogProbs().shape == (2, 3, 5)
    assert_array_equal(em.getLogProbs()[:, :, 0], 0.0)          # missing symbol: log 1
    assert em.trackTableWidths() == [3, 5]
    h = MultitrackHmm(em, transmat=np.array([[1., 0., 0.], [0., .5, .5], [.2, .3, .5]]))
    assert h._log_transmat[0, 1] == -1e100                      # zeros stay zeros (hmm.py:645)
    h2 = MultitrackHmm(em, transmat=np.array([[1., 0., 0.], [0., .5, .5], [.2, .3, .5]]),
                       transMatEpsilons=True)
    assert h2._log_transmat[0, 1] > -40                         # epsilon added (hmm.py:635-636)
    with pytest.raises(ValueError):
        MultitrackHmm(em, transmat=np.ones((3, 3)))
    with pytest.raises(ValueError):
        MultitrackHmm(em, startprob=[0.5, 0.2, 0.2])
    assert h.algorithm == "viterbi"
    with pytest.raises(ValueError):
        h.algorithm = "bogus"


def test_maximize_rules():
    """emission.py:243-267: -1e6 for learned zeros, orphaned rows kept, fudge floor."""
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    em = IndependentMultinomialEmissionModel(2, [3])
    before = em.getLogProbs().copy()
    stats = np.zeros((1, 2, 4))
    stats[0, 0, 1:] = [3.0, 0.0, 1.0]
    em.maximize(stats)
    assert_allclose(np.exp(em.getLogProbs()[0, 0, 1:]), [0.75, 0.0, 0.25])
    assert em.getLogProbs()[0, 0, 2] == -1e6
    assert_array_equal(em.getLogProbs()[0, 1], before[0, 1])    # no mass: unchanged


def test_track_table_ratios_and_overlap():
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.track import IntegerTrackTable
    t = IntegerTrackTable(2, "c", 100, 200)
    assert t.shape == (100, 2)
    t.setSegments([0, 10, 30, 90])
    assert len(t) == 4
    assert_array_equal(t.getSegmentLengthsAsRatio(10), [1.0, 2.0, 6.0, 1.0])
    em = IndependentMultinomialEmissionModel(2, [2, 2], effectiveSegmentLength=10)
    assert_array_equal(em.getSegmentRatios(t), [1.0, 2.0, 6.0, 1.0])
    assert em.getSegmentRatios(t.getNumPyArray()) is None
    assert t.getOverlapInTableCoords(("c", 110, 195, 3)) == ["c", 1, 4, 3]
    assert t.getOverlapInTableCoords(("d", 110, 195, 3)) is None
    u = IntegerTrackTable(2, "c", 100, 200)
    assert u.getOverlapInTableCoords(("c", 50, 120, 1)) == ["c", 0, 20, 1]


def test_result_pool_recycles_only_unreachable_blocks():
    """engine._ResultPool: a block goes back to the pool only when the array handed out
    and every slice of it are gone; live results are never aliased."""
    import gc
    from tehmm_b200.engine import _ResultPool
    pool = _ResultPool()
    a = pool.empty_int64(1000)
    a[:] = 7
    part = a[10:20]
    b = pool.empty_int64(1000)
    b[:] = 9
    assert a[0] == 7 and not np.shares_memory(a, b)
    del a
    gc.collect()
    assert len(pool._free) == 0          # `part` still references the block
    c = pool.empty_int64(1000)
    c[:] = 1
    assert part[0] == 7
    del part
    gc.collect()
    assert len(pool._free) == 1
    d = pool.empty_int64(900)            # recycled
    assert len(pool._free) == 0
    d[:] = 5
    assert b[0] == 9 and c[0] == 1
    assert d.dtype == np.int64 and d.flags.writeable and d.shape == (900,)


@pytest.mark.timeout(120)
def test_host_thread_pool_back_to_back_jobs():
    """the decode path's thread pool under back-to-back parallel loops (a worker of the previous loop
    coming back for more must not consume an index of the next one: that lost a task and hung the call)"""
    from tehmm_b200 import _lib
    lib = _lib.load()
    for threads in (2, 4, 16):
        assert lib.tehmm_host_pool_selftest(threads, 200_000) == 0


def test_fast_almost_equal_keeps_numpys_criterion():
    """validate() (hmm.py:576-616 calls it twice per EM iteration) uses a cheap restatement of
    numpy.testing.assert_array_almost_equal: same verdicts, same AssertionError."""
    import numpy as np
    import pytest
    from numpy.testing import assert_array_almost_equal
    from tehmm_b200.common import assert_almost_equal_fast
    ones = np.ones(30)
    cases = [(ones + 1e-7, ones), (ones + 1.4e-6, ones), (ones + 1.6e-6, ones), (0.9, 1.0), (np.float64(1.0), 1.0),
             (np.array([np.nan]), np.array([1.0])), (np.ones(3), np.ones(4)), (np.array([np.inf]), np.array([np.inf]))]
    for got, want in cases:
        try:
            assert_array_almost_equal(got, want)
            ok = True
        except AssertionError:
            ok = False
        if ok:
            assert_almost_equal_fast(got, want)
        else:
            with pytest.raises(AssertionError):
                assert_almost_equal_fast(got, want)
