"""CPU: the checkers of tests/parity.py themselves, against the oracle."""
import numpy as np
import pytest

from parity import near_tie_report, oracle_all, viterbi_terms


@pytest.mark.parametrize("with_ratios", [False, True])
def test_viterbi_terms_sum_to_the_reference_score(oracle, with_ratios):
    """sum(viterbi_terms(path)) == _hmm._viterbi's logprob for its own path (_hmm.pyx:214-248,
    incl. the from-state-0 ratio asymmetry), and any other path scores no higher."""
    from tehmm_b200 import synth
    m = synth.make_model(N=7, syms=(4, 3, 2), seed=5, zero_frac=0.0)
    T = 400
    obs = synth.sample_obs(m, T, seed=6)[0]
    r = np.random.RandomState(7).uniform(0.05, 6.0, size=T) if with_ratios else None
    ref = oracle_all(oracle, obs, m["table"], 1.0, m["log_start"], m["log_trans"], r_em=None, r_dp=r)
    terms = viterbi_terms(ref["frame"], m["log_start"], m["log_trans"], r, ref["vit_states"])
    assert terms.sum() == pytest.approx(ref["vit_logprob"], rel=1e-13)
    other = ref["vit_states"].copy()
    other[100:110] = (other[100:110] + 1) % 7
    other[300] = (other[300] + 3) % 7
    runs, steps, worst, where = near_tie_report(other, ref["vit_states"], ref["frame"], m["log_start"],
                                                m["log_trans"], r)
    assert runs == 2 and steps == 11
    total = viterbi_terms(ref["frame"], m["log_start"], m["log_trans"], r, other).sum()
    assert total <= ref["vit_logprob"]
    assert worst > 1e-3                      # a real detour is not a near-tie
    assert near_tie_report(ref["vit_states"], ref["vit_states"], ref["frame"], m["log_start"],
                           m["log_trans"], r)[0] == 0
