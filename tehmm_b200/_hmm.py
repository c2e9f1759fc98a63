"""Drop-in for the reference's `_hmm` Cython module (/root/reference/_hmm.pyx).

Same function names, positional argument order, in/out conventions and dtypes.
The arithmetic runs in the strict float64 CUDA kernels of libtehmm_b200.so
(csrc/strict.cu), which keep the reference's operation order.
"""
import numpy as np

from . import _lib


def _check_lattice(a, T, N, name):
    assert isinstance(a, np.ndarray), "%s must be a NumPy array" % name
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'dtype_t' but got '%s'" % a.dtype)
    assert a.ndim == 2 and a.shape[0] >= T and a.shape[1] == N, "%s has shape %s, expected (%d, %d)" % (name, a.shape, T, N)


def _ratios(segRatios, T):
    if segRatios is None:
        return None
    r = _lib.f64(segRatios)
    assert r.ndim == 1 and r.shape[0] >= T
    return r


def _forward(n_observations, n_components, log_startprob, log_transmat, framelogprob,
             segRatios, fwdlattice):
    """_hmm.pyx:120-158.  fwdlattice (T,N) float64 is written in place."""
    T, N = int(n_observations), int(n_components)
    _check_lattice(fwdlattice, T, N, "fwdlattice")
    ls, lt, fr = _lib.f64(log_startprob, (N,)), _lib.f64(log_transmat, (N, N)), _lib.f64(framelogprob)
    _check_lattice(fr, T, N, "framelogprob")
    r = _ratios(segRatios, T)
    out = fwdlattice if fwdlattice.flags.c_contiguous else np.empty((T, N))
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_forward(ctx.handle, T, N, _lib.ptr(ls), _lib.ptr(lt), _lib.ptr(fr),
                                            _lib.ptr(r), _lib.ptr(out)))
    if out is not fwdlattice:
        fwdlattice[:T] = out


def _backward(n_observations, n_components, log_startprob, log_transmat, framelogprob,
              segRatios, bwdlattice):
    """_hmm.pyx:160-198.  bwdlattice (T,N) float64 is written in place; last row log(1/N)."""
    T, N = int(n_observations), int(n_components)
    _check_lattice(bwdlattice, T, N, "bwdlattice")
    ls, lt, fr = _lib.f64(log_startprob, (N,)), _lib.f64(log_transmat, (N, N)), _lib.f64(framelogprob)
    _check_lattice(fr, T, N, "framelogprob")
    r = _ratios(segRatios, T)
    out = bwdlattice if bwdlattice.flags.c_contiguous else np.empty((T, N))
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_backward(ctx.handle, T, N, _lib.ptr(ls), _lib.ptr(lt), _lib.ptr(fr),
                                             _lib.ptr(r), _lib.ptr(out)))
    if out is not bwdlattice:
        bwdlattice[:T] = out


def _viterbi(n_observations, n_components, log_startprob, log_transmat, segRatios, framelogprob):
    """_hmm.pyx:201-259.  Returns (state_sequence int64[T], logprob float).
    Note the argument order: segRatios comes BEFORE framelogprob, as in the reference."""
    T, N = int(n_observations), int(n_components)
    ls, lt, fr = _lib.f64(log_startprob, (N,)), _lib.f64(log_transmat, (N, N)), _lib.f64(framelogprob)
    _check_lattice(fr, T, N, "framelogprob")
    r = _ratios(segRatios, T)
    states = np.empty(T, dtype=np.int64)
    lp = np.zeros(1)
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_viterbi(ctx.handle, T, N, _lib.ptr(ls), _lib.ptr(lt), _lib.ptr(r),
                                            _lib.ptr(fr), _lib.ptr(states), _lib.ptr(lp)))
    return states, float(lp[0])


def _log_sum_lneta(n_observations, n_components, fwdlattice, log_transmat, bwdlattice,
                   framelogprob, logprob, segRatios, logsum_lneta):
    """_hmm.pyx:62-117.  logsum_lneta (N,N) must be zeros on entry; written in place."""
    T, N = int(n_observations), int(n_components)
    assert isinstance(logsum_lneta, np.ndarray) and logsum_lneta.dtype == np.float64
    assert logsum_lneta.shape == (N, N)
    fw, bw, fr = _lib.f64(fwdlattice), _lib.f64(bwdlattice), _lib.f64(framelogprob)
    for a, n in ((fw, "fwdlattice"), (bw, "bwdlattice"), (fr, "framelogprob")):
        _check_lattice(a, T, N, n)
    lt = _lib.f64(log_transmat, (N, N))
    r = _ratios(segRatios, T)
    out = logsum_lneta if logsum_lneta.flags.c_contiguous else np.ascontiguousarray(logsum_lneta)
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_log_sum_lneta(ctx.handle, T, N, _lib.ptr(fw), _lib.ptr(lt), _lib.ptr(bw),
                                                  _lib.ptr(fr), float(logprob), _lib.ptr(r), _lib.ptr(out)))
    if out is not logsum_lneta:
        logsum_lneta[...] = out
