"""-m gpu: replicate training on worker threads (bin/teHmmTrain.py:279-306): per-thread contexts
and streams, same models as training the replicates one after the other, best replicate selected
like the reference does."""
import threading

import numpy as np
import pytest
from numpy.testing import assert_array_equal

pytestmark = pytest.mark.gpu


def make_trainer(obs):
    from tehmm_b200 import synth
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    seen = []

    def train_one(seed):
        m0 = synth.make_model(N=6, syms=(3, 4, 2), seed=int(seed), zero_frac=0.0)
        em = IndependentMultinomialEmissionModel(m0["N"], list(m0["syms"]), zeroAsMissingData=True)
        em.logProbs = m0["table"].copy()
        hmm = MultitrackHmm(em, startprob=m0["pi"].copy(), transmat=m0["A"].copy(), n_iter=6, thresh=0.0)
        hmm.fit(obs)
        seen.append(threading.get_ident())
        return hmm
    return train_one, seen


def test_replicates_on_threads_equal_serial():
    from tehmm_b200 import synth
    from tehmm_b200.replicates import train_replicates
    truth = synth.make_model(N=6, syms=(3, 4, 2), seed=1)
    obs = [synth.sample_obs(truth, T, seed=10 + i)[0] for i, T in enumerate([4000, 1500, 9000, 700])]
    seeds = [11, 12, 13, 14, 15, 16]
    train_one, seen = make_trainer(obs)
    serial, best_s = train_replicates(train_one, seeds, num_threads=1)
    train_two, seen2 = make_trainer(obs)
    threaded, best_t = train_replicates(train_two, seeds, num_threads=3)
    assert len(set(seen2)) > 1, "the replicates must have run on several threads"
    for a, b in zip(serial, threaded):
        assert a.getLastLogProb() == b.getLastLogProb()
        assert_array_equal(a.transmat_, b.transmat_)
        assert_array_equal(a.emissionModel.getLogProbs(), b.emissionModel.getLogProbs())
    lps = [m.getLastLogProb() for m in threaded]
    assert best_t == best_s == int(np.argmax(lps))
    assert len(set(lps)) > 1, "different seeds must reach different optima in this test"


def test_contexts_are_per_thread():
    """two threads decoding at the same time through their own contexts give the single-thread answer"""
    from concurrent.futures import ThreadPoolExecutor
    from tehmm_b200 import _lib, synth
    from tehmm_b200.emission import IndependentMultinomialEmissionModel
    from tehmm_b200.hmm import MultitrackHmm
    models = [synth.make_model(N=30, seed=s) for s in (3, 4)]
    obs = [synth.sample_obs(m, 200_000, seed=5 + i)[0] for i, m in enumerate(models)]

    def decode(i):
        m = models[i]
        em = IndependentMultinomialEmissionModel(m["N"], list(m["syms"]), zeroAsMissingData=True)
        em.logProbs = m["table"].copy()
        hmm = MultitrackHmm(em, startprob=m["pi"].copy(), transmat=m["A"].copy())
        out = None
        for _ in range(5):
            out = hmm.decode(obs[i])
        return out, id(_lib.get_context())
    want = [decode(0), decode(1)]
    with ThreadPoolExecutor(max_workers=2) as pool:
        got = list(pool.map(decode, [0, 1]))
    assert got[0][1] != got[1][1] and got[0][1] != want[0][1]
    for g, w in zip(got, want):
        assert g[0][0] == w[0][0]
        assert_array_equal(g[0][1], w[0][1])
