"""Drop-in for the reference's `_emission` Cython module (/root/reference/_emission.pyx).

canFast / fastAllLogProbs / fastAccumulateStats / fastUpdateCounts with the
reference's signatures; uint8, uint16 and int32 observation matrices (the three
dtype clones of the reference) run in the strict float64 CUDA kernels.
"""
import numpy as np

from . import _lib
from .track import is_track_table

_FAST_DTYPES = (np.dtype(np.int32), np.dtype(np.uint16), np.dtype(np.uint8))


def canFast(obs):
    """_emission.pyx:14-18."""
    return is_track_table(obs) or (isinstance(obs, np.ndarray) and obs.dtype in _FAST_DTYPES)


def _obs_array(obs):
    if is_track_table(obs):
        obs = obs.getNumPyArray()
    assert isinstance(obs, np.ndarray)
    assert len(obs.shape) == 2
    assert obs.dtype in _FAST_DTYPES, "unsupported observation dtype %s" % obs.dtype
    return np.ascontiguousarray(obs)


def fastAllLogProbs(obs, logProbs, outProbs, normalize, segRatios):
    """_emission.pyx:20-80: outProbs[i,j] = normalize * sum_k logProbs[k,j,obs[i,k]]
    (* segRatios[i]); rows before the first feasible row are zeroed."""
    obs = _obs_array(obs)
    assert isinstance(logProbs, np.ndarray)
    assert isinstance(outProbs, np.ndarray)
    assert len(logProbs.shape) == 3
    assert logProbs.dtype == np.float64
    assert outProbs.dtype == np.float64
    assert outProbs.shape[0] == obs.shape[0]
    assert logProbs.shape[0] == obs.shape[1]
    T, K = obs.shape
    _, N, S = logProbs.shape
    assert outProbs.shape[1] == N
    tab = np.ascontiguousarray(logProbs)
    r = None if segRatios is None else _lib.f64(segRatios)
    out = outProbs if outProbs.flags.c_contiguous else np.empty((T, N))
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_all_log_probs(ctx.handle, _lib.ptr(obs), obs.dtype.itemsize, T, K,
                                                  _lib.ptr(tab), N, S, _lib.ptr(out), float(normalize),
                                                  _lib.ptr(r)))
    if out is not outProbs:
        outProbs[...] = out


def fastAccumulateStats(obs, obsStats, posteriors, segRatios):
    """_emission.pyx:146-190: obsStats[k, j, obs[i,k]] += posteriors[i,j] (* segRatios[i])."""
    obs = _obs_array(obs)
    assert isinstance(obsStats, np.ndarray) and obsStats.dtype == np.float64 and obsStats.ndim == 3
    T, K = obs.shape
    K2, N, S = obsStats.shape
    assert K2 == K
    post = _lib.f64(posteriors)
    assert post.shape == (T, N)
    r = None if segRatios is None else _lib.f64(segRatios)
    st = obsStats if obsStats.flags.c_contiguous else np.ascontiguousarray(obsStats)
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_accumulate_stats(ctx.handle, _lib.ptr(obs), obs.dtype.itemsize, T, K,
                                                     _lib.ptr(st), N, S, _lib.ptr(post), _lib.ptr(r)))
    if st is not obsStats:
        obsStats[...] = st


def fastUpdateCounts(bedInterval, trackTable, obsStats, segRatios):
    """_emission.pyx:236-332: supervised counts over [start,end) in TABLE coordinates."""
    assert is_track_table(trackTable)
    obs = _obs_array(trackTable)
    assert isinstance(obsStats, np.ndarray) and obsStats.dtype == np.float64 and obsStats.ndim == 3
    T, K = obs.shape
    _, N, S = obsStats.shape
    start, end, state = int(bedInterval[1]), int(bedInterval[2]), int(bedInterval[3])
    r = None if segRatios is None else _lib.f64(segRatios)
    st = obsStats if obsStats.flags.c_contiguous else np.ascontiguousarray(obsStats)
    ctx = _lib.get_context()
    _lib.check(ctx.lib.tehmm_strict_update_counts(ctx.handle, _lib.ptr(obs), obs.dtype.itemsize, T, K, start, end,
                                                  state, _lib.ptr(st), N, S, _lib.ptr(r)))
    if st is not obsStats:
        obsStats[...] = st
