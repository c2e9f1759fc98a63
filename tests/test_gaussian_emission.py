"""CPU: IndependentMultinomialAndGaussianEmissionModel (reference
emission.py:483-615) against golden outputs of the reference's own class
(tests/golden/make_golden.py:gaussian_case): constructor-time makeGaussian, maximize() on
posterior-weighted counts, and a user `STATE TRACK MEAN STDEV` line."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import golden


class _ValueMap(object):
    def __init__(self, values):
        self.values = list(values)

    def getMapBack(self, symbol):           # CategoryMap.getMapBack (track.py:714-724), reserved = 1
        return self.values[int(symbol) - 1]


class _Track(object):
    def __init__(self, number, dist, values):
        self.number, self.dist, self.valMap = number, dist, _ValueMap(values)

    def getNumber(self):
        return self.number

    def getName(self):
        return "t%d" % self.number

    def getDist(self):
        return self.dist

    def getValueMap(self):
        return self.valMap


def _model(g):
    from tehmm_b200.emission import IndependentMultinomialAndGaussianEmissionModel
    N, syms = int(g["N"]), [int(s) for s in g["syms"]]
    tracks = [_Track(k, "gaussian" if k == 1 else "multinomial",
                     g["values_1"] if k == 1 else [str(v) for v in range(syms[k])]) for k in range(len(syms))]
    params = [g["init_params_%d" % k].tolist() for k in range(len(syms))]
    em = IndependentMultinomialAndGaussianEmissionModel(N, syms, tracks, params, zeroAsMissingData=True, fudge=0.0)
    return em, tracks, syms


@pytest.mark.gpu
def test_gaussian_model_statistics_on_the_device():
    """accumulateStats of the subclass runs the same CUDA kernel (fastAccumulateStats)"""
    g = golden("gaussian")
    em, tracks, syms = _model(g)
    stats = em.initStats()
    em.accumulateStats(g["obs"], stats, g["post"])
    assert_allclose(np.array(stats), g["stats"], rtol=1e-12)
    em.maximize(stats, tracks)
    assert_allclose(em.getLogProbs(), g["table1"], rtol=1e-10, atol=1e-300)


def test_gaussian_tracks_match_reference():
    g = golden("gaussian")
    em, tracks, syms = _model(g)
    assert_allclose(em.gaussParams, g["gauss0"], rtol=1e-13)
    assert_allclose(em.getLogProbs(), g["table0"], rtol=1e-12, atol=1e-300)
    # the multinomial tracks are untouched by makeGaussian
    assert_array_equal(em.getLogProbs()[0], g["table0"][0])
    em.maximize(g["stats"].copy(), tracks)
    assert_allclose(em.gaussParams, g["gauss1"], rtol=1e-12)
    assert_allclose(em.getLogProbs(), g["table1"], rtol=1e-11, atol=1e-300)
    # every row of the re-fitted gaussian track is a distribution over its symbols
    assert_allclose(np.exp(em.getLogProbs()[1, :, 1:syms[1] + 1]).sum(axis=1), 1.0, rtol=1e-12)
    logProbs = em.getLogProbs().copy()
    mask = np.zeros(logProbs.shape, dtype=np.int8)
    em.applyUserEmissionLine(tracks[1], 2, ["2", "t1", "7.5", "2.25"], logProbs, mask)
    assert_allclose(logProbs, g["table_user"], rtol=1e-11, atol=1e-300)
    assert_array_equal(mask, g["mask_user"])
    assert_allclose(em.gaussParams, g["gauss_user"], rtol=1e-12)
