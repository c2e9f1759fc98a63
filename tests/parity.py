"""Shared parity helpers for the -m gpu tests (test infrastructure, like oracle/).

Tolerances are north_star's: float32 production path <= 1e-5 relative on
log-likelihood, posteriors, expected counts and re-estimated parameters; float64
verification mode <= 1e-10.  Values below ATOL are rounding of zero and are
compared absolutely.

fp32 Viterbi: the path must equal the reference's except at NEAR-TIES.  That is
checked per divergence, not per path: for every maximal run of steps where our
path differs from the reference's, the float64 score of both paths over that run
(the terms of _hmm.pyx:214-248, including the transition that leaves the run)
must differ by at most NEAR_TIE_EPS log units.
"""
import numpy as np

TOL = {"f32": 1e-5, "f64": 1e-10}
ATOL = {"f32": 2e-6, "f64": 1e-12}
# absolute float64 score difference tolerated over one divergent run of an fp32 Viterbi path
NEAR_TIE_EPS = 1e-3


def viterbi_terms(frame, log_start, log_trans, ratios, states):
    """terms[t] = what _hmm._viterbi adds to the lattice at step t along `states`
    (/root/reference/_hmm.pyx:214-248, including the from-state-0 segment-ratio
    asymmetry of :234-237); sum(terms) is the path's score."""
    states = np.asarray(states, dtype=np.int64)
    T = len(states)
    frame = np.asarray(frame, dtype=np.float64)
    terms = np.empty(T, dtype=np.float64)
    diag = np.diag(log_trans)
    s0 = states[0]
    terms[0] = log_start[s0] + frame[0, s0]
    if ratios is not None and ratios[0] > 1.0:
        terms[0] += diag[s0] * (ratios[0] - 1.0)
    if T == 1:
        return terms
    i, j = states[:-1], states[1:]
    t = np.arange(1, T)
    base = log_trans[i, j] + frame[t, j]
    if ratios is not None:
        r = np.asarray(ratios, dtype=np.float64)[1:]
        from0 = base + diag[j] * r - np.where(j == 0, log_trans[0, j], 0.0)
        other = base + np.where(r > 1.0, diag[j] * (r - 1.0), 0.0)
        base = np.where(i == 0, from0, other)
    terms[1:] = base
    return terms


def near_tie_report(ours, ref, frame, log_start, log_trans, ratios=None):
    """Per maximal divergent run [s, e): |score_ours - score_ref| over t in [s, e]
    (the step at e is where the paths re-join: same state, different from-state).
    Returns (number of runs, divergent steps, max abs score difference, its run)."""
    ours = np.asarray(ours, dtype=np.int64)
    ref = np.asarray(ref, dtype=np.int64)
    T = len(ref)
    assert ours.shape == ref.shape
    diff = ours != ref
    if not diff.any():
        return 0, 0, 0.0, None
    to = viterbi_terms(frame, log_start, log_trans, ratios, ours)
    tr = viterbi_terms(frame, log_start, log_trans, ratios, ref)
    d = np.concatenate([[0.0], np.cumsum(to - tr)])          # d[t] = sum over steps < t
    edge = np.diff(np.concatenate([[0], diff.astype(np.int8), [0]]))
    starts = np.nonzero(edge == 1)[0]
    ends = np.nonzero(edge == -1)[0]                           # exclusive
    hi = np.minimum(ends + 1, T)                               # include the re-joining step
    gaps = np.abs(d[hi] - d[starts])
    k = int(np.argmax(gaps))
    return len(starts), int(diff.sum()), float(gaps[k]), (int(starts[k]), int(ends[k]))


def assert_near_ties_only(ours, ref, frame, log_start, log_trans, ratios=None,
                          eps=NEAR_TIE_EPS, min_agree=0.9, label=""):
    """fp32 Viterbi parity: every divergence from the reference path is a near-tie."""
    runs, steps, worst, where = near_tie_report(ours, ref, frame, log_start, log_trans, ratios)
    T = len(ref)
    if runs:
        print("near-ties %s: %d divergent runs, %d / %d steps, max |score diff| %.3e at %s"
              % (label, runs, steps, T, worst, where))
    assert worst <= eps, "divergent run %s differs by %.3e log units (> %g): not a near-tie" % (
        where, worst, eps)
    assert steps <= (1.0 - min_agree) * T or T < 50, (steps, T)
    return runs, steps, worst


def min_rtol(got, want, atol):
    """smallest rtol for which np.allclose(got, want, rtol, atol) holds (diagnostics)"""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = np.abs(got - want) - atol
    den = np.abs(want)
    m = (err > 0) & (den > 0)
    return float((err[m] / den[m]).max()) if m.any() else 0.0


def oracle_frame(oracle, obs, table, normalize=1.0, ratios=None):
    """framelogprob of the reference (_emission.pyx:20-144) for one sequence"""
    frame = np.zeros((obs.shape[0], table.shape[1]))
    oracle.fastAllLogProbs(obs, table, frame, normalize, ratios)
    return frame


def oracle_all(oracle, obs, m_table, normalize, log_start, log_trans, r_em=None, r_dp=None):
    """every per-sequence output of the reference flow (hmm.py:545-574, basehmm.py:238-359)"""
    T, N = obs.shape[0], log_start.shape[0]
    frame = oracle_frame(oracle, obs, m_table, normalize, r_em)
    fwd, bwd = np.zeros((T, N)), np.zeros((T, N))
    oracle._forward(T, N, log_start, log_trans, frame, r_dp, fwd)
    oracle._backward(T, N, log_start, log_trans, frame, r_dp, bwd)
    lp = oracle.logsumexp(fwd[-1])
    post = oracle.posteriors(fwd, bwd)
    states, vlp = oracle._viterbi(T, N, log_start, log_trans, r_dp, frame)
    return dict(frame=frame, fwd=fwd, bwd=bwd, logprob=lp, post=post, vit_states=states, vit_logprob=vlp)


def assert_map_near_ties_only(ours, ref_post, rel=1e-4, label=""):
    """MAP (posterior arg-max) parity for the fp32 path: wherever our state is not the
    reference's arg-max (basehmm.py:357), its reference posterior must be within `rel`
    of the row maximum -- a tie at float32 resolution of the normalised posterior."""
    ours = np.asarray(ours, dtype=np.int64)
    ref = np.argmax(ref_post, axis=1)
    bad = np.nonzero(ours != ref)[0]
    if bad.size == 0:
        return 0, 0.0
    top = ref_post[bad, ref[bad]]
    mine = ref_post[bad, ours[bad]]
    gap = float(np.max((top - mine) / top))
    print("MAP near-ties %s: %d / %d steps, max relative posterior gap %.3e" % (label, bad.size, len(ref), gap))
    assert gap <= rel, "MAP state differs where the posteriors are %.3e apart (> %g)" % (gap, rel)
    return int(bad.size), gap


def trans_counts_extended(frames, log_start, log_trans):
    """Expected transition counts (what hmm.py:559-568 adds to stats['trans'], including the
    1/N of _hmm.pyx:179) in numpy.longdouble, scaled-probability space: an arbiter with 64
    mantissa bits for the cases where our float64 kernels and the reference's float64 LOG-space
    lattices (absolute rounding ulp(|log alpha|), growing with T) disagree beyond 1e-10."""
    LD = np.longdouble
    N = log_start.shape[0]
    A = np.exp(np.where(log_trans <= -1e30, -np.inf, log_trans).astype(LD))
    pi = np.exp(np.where(log_start <= -1e30, -np.inf, log_start).astype(LD))
    total = np.zeros((N, N), dtype=LD)
    for frame in frames:
        T = frame.shape[0]
        if T < 2:
            continue
        f = frame.astype(LD)
        b = np.exp(f - f.max(axis=1, keepdims=True))
        alpha = np.empty((T, N), dtype=LD)
        a = pi * b[0]
        alpha[0] = a / a.sum()
        c = np.empty(T, dtype=LD)
        for t in range(1, T):
            a = (alpha[t - 1] @ A) * b[t]
            c[t] = a.sum()
            alpha[t] = a / c[t]
        beta = np.ones(N, dtype=LD)
        for t in range(T - 2, -1, -1):
            w = b[t + 1] * beta / c[t + 1]
            total += np.outer(alpha[t], w) * A
            beta = A @ w
    return np.asarray(total / LD(N), dtype=np.float64)
