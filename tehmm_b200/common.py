"""Numerics contract shared with the reference (common.py:24-33, basehmm.py:65-67)."""
import logging

import numpy as np

LOGZERO = -1e100
EPSILON = np.finfo(float).eps
ZEROLOGPROB = -1e200
NEGINF = -np.inf
EPS = np.finfo(float).eps

logger = logging.getLogger("tehmm_b200")


def myLog(x, logZeroVal=LOGZERO, epsilonVal=EPSILON):
    """np.log that maps |x| < eps to a finite sentinel instead of -inf
    (reference common.py:27-33; vectorised with np.where instead of np.vectorize).
    Returns an array for array input and a numpy scalar for scalar input."""
    a = np.asarray(x, dtype=np.float64)
    small = np.abs(a) < epsilonVal
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(small, logZeroVal, np.log(np.where(small, 1.0, a)))
    return out if out.ndim else np.float64(out)


def logsumexp(arr, axis=0):
    """log(sum(exp(arr))) along `axis` with the max pulled out (basehmm.py:70-93)."""
    arr = np.moveaxis(np.asarray(arr, dtype=np.float64), axis, 0)
    vmax = arr.max(axis=0)
    out = np.log(np.sum(np.exp(arr - vmax), axis=0))
    out += vmax
    return out


def normalize(A, axis=None):
    """Adds EPS to every entry IN PLACE, then returns A / A.sum(axis)
    (basehmm.py:113-141; the in-place epsilon is part of the contract)."""
    A += EPS
    Asum = A.sum(axis)
    if axis and A.ndim > 1:
        Asum[Asum == 0] = 1
        shape = list(A.shape)
        shape[axis] = 1
        Asum.shape = shape
    return A / Asum


def assert_almost_equal_fast(actual, desired, decimal=6):
    """numpy.testing.assert_array_almost_equal's criterion (abs(desired - actual) < 1.5 * 10**-decimal, NaNs must
    match) without its message machinery: validate() runs twice per EM iteration (hmm.py:576-616), and the numpy
    helper costs 80 us a call.  Falls back to the numpy helper for the error message."""
    a, d = np.asarray(actual, dtype=np.float64), np.asarray(desired, dtype=np.float64)
    if a.shape == d.shape or a.ndim == 0 or d.ndim == 0:
        with np.errstate(invalid="ignore"):          # inf - inf: left to the numpy helper below
            diff = np.abs(d - a)
        if bool(np.all(diff < 1.5 * 10.0 ** (-decimal))):
            return
    from numpy.testing import assert_array_almost_equal
    assert_array_almost_equal(actual, desired, decimal)
