"""Decode output path: teHmmEval's statesToBed
(/root/reference/bin/teHmmEval.py:238-270) on the library's native writers
(tehmm_states_to_bed, tehmm_scores_to_bed).  Same lines, same order, one per observation --
contiguous equal states are not merged (teHmmEval.py:241-243); the posterior / emission
score files carry "%s" of a NumPy float64 in the fourth column, as the reference prints it.
Host only.
"""
import ctypes

import numpy as np

from . import _lib


def _intervals(trackTable, n):
    """(chrom, start, seg lengths or None, mask offsets or None) of the n observations"""
    chrom = str(trackTable.getChrom())
    start = int(trackTable.getStart())
    end = int(trackTable.getEnd())
    segOffsets = trackTable.getSegmentOffsets()
    maskOffsets = trackTable.getMaskRunningOffsets() if hasattr(trackTable, "getMaskRunningOffsets") else None
    if segOffsets is None:
        assert n == end - start
    seg = None
    if segOffsets is not None:
        offs = np.asarray(segOffsets, dtype=np.int64)
        assert len(offs) == n
        seg = np.empty(n, dtype=np.int64)
        seg[:-1] = offs[1:] - offs[:-1]
        seg[-1] = end - (start + offs[-1])          # TrackTable.getSegmentLength (track.py:497-502)
    mask = None if maskOffsets is None else np.ascontiguousarray(maskOffsets, dtype=np.int32)
    return chrom, start, seg, mask


def _write_scores(lib, f, chrom, start, scores, seg, mask):
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    f.flush()
    _lib.check(lib.tehmm_scores_to_bed(f.fileno(), chrom.encode(), start, _lib.ptr(scores), scores.shape[0],
                                       _lib.ptr(seg), _lib.ptr(mask), 0 if mask is None else mask.shape[0]))


def statesToBed(trackTable, states, bedFile, posteriors=None, posteriorsMask=None, posteriorsFile=None,
                emProbs=None, emissionsMask=None, emissionsFile=None, stateNames=None):
    """The reference's signature (teHmmEval.py:238-240) plus `stateNames`.

    trackTable : anything with getChrom / getStart / getEnd / getSegmentOffsets (and optionally
                 getSegmentLength, getMaskRunningOffsets), as the reference's TrackTable
    states     : int sequence (decode() output), or a sequence of names when the caller has
                 already mapped them (MultitrackHmm.viterbi does, hmm.py:233-234)
    bedFile    : open file, or None (teHmmEval.py:263)
    posteriors / posteriorsMask / posteriorsFile : (T, N) posterior distribution, (N,) 0/1 mask of the
                 states of interest, open file: line i carries sum(posteriors[i-1] * mask) -- the
                 reference indexes i-1 (teHmmEval.py:266-267: row 0 shows the LAST observation's value)
    emProbs / emissionsMask / emissionsFile : the same with log(sum(exp(emProbs[i-1]) * mask))
    stateNames : optional list, state index -> name
    """
    n = len(states)
    chrom, start, seg, mask = _intervals(trackTable, n)
    lib = _lib.load()
    if posteriors is not None:
        # row sums in NumPy (pairwise summation over the contiguous axis, exactly np.sum of a row)
        sc = (np.asarray(posteriors) * np.asarray(posteriorsMask)).sum(axis=1)
        assert sc.shape[0] == n
        _write_scores(lib, posteriorsFile, chrom, start, np.roll(sc, 1), seg, mask)
    if emProbs is not None:
        with np.errstate(divide="ignore"):
            sc = np.log((np.exp(np.asarray(emProbs)) * np.asarray(emissionsMask)).sum(axis=1))
        assert sc.shape[0] == n
        _write_scores(lib, emissionsFile, chrom, start, np.roll(sc, 1), seg, mask)
    if bedFile is None:
        return
    names = None if stateNames is None else [str(x) for x in stateNames]
    st = np.asarray(states)
    if st.dtype.kind not in "iu":
        # names in, as returned by MultitrackHmm.viterbi with a stateNameMap: index them
        uniq, inv = np.unique(st.astype(str), return_inverse=True)
        names, st = [str(u) for u in uniq], inv
    st = np.ascontiguousarray(st, dtype=np.int64)
    cnames = None
    if names is not None:
        cnames = (ctypes.c_char_p * len(names))(*[s.encode() for s in names])
    bedFile.flush()
    _lib.check(lib.tehmm_states_to_bed(bedFile.fileno(), chrom.encode(), start, _lib.ptr(st), n, _lib.ptr(seg),
                                       _lib.ptr(mask), 0 if mask is None else mask.shape[0], cnames,
                                       0 if names is None else len(names)))
