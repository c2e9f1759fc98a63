"""Decode output path: teHmmEval's statesToBed
(/root/reference/bin/teHmmEval.py:238-262), bedFile part, on the library's native
writer (tehmm_states_to_bed).  Same lines, same order, one per observation --
contiguous equal states are not merged (teHmmEval.py:241-243).  Host only.
"""
import ctypes

import numpy as np

from . import _lib


def statesToBed(trackTable, states, bedFile, stateNames=None):
    """Write `states` (one per observation of `trackTable`) to the open file `bedFile`.

    trackTable : anything with getChrom / getStart / getEnd / getSegmentOffsets (and optionally
                 getSegmentLength, getMaskRunningOffsets), as the reference's TrackTable
    states     : int sequence (decode() output), or a sequence of names when the caller has
                 already mapped them (MultitrackHmm.viterbi does, hmm.py:233-234)
    stateNames : optional list, state index -> name
    """
    chrom = str(trackTable.getChrom())
    start = int(trackTable.getStart())
    end = int(trackTable.getEnd())
    segOffsets = trackTable.getSegmentOffsets()
    maskOffsets = trackTable.getMaskRunningOffsets() if hasattr(trackTable, "getMaskRunningOffsets") else None
    n = len(states)
    if segOffsets is None:
        assert n == end - start
    names = None if stateNames is None else [str(x) for x in stateNames]
    st = np.asarray(states)
    if st.dtype.kind not in "iu":
        # names in, as returned by MultitrackHmm.viterbi with a stateNameMap: index them
        uniq, inv = np.unique(st.astype(str), return_inverse=True)
        names, st = [str(u) for u in uniq], inv
    st = np.ascontiguousarray(st, dtype=np.int64)
    seg = None
    if segOffsets is not None:
        offs = np.asarray(segOffsets, dtype=np.int64)
        assert len(offs) == n
        seg = np.empty(n, dtype=np.int64)
        seg[:-1] = offs[1:] - offs[:-1]
        seg[-1] = end - (start + offs[-1])          # TrackTable.getSegmentLength (track.py:497-502)
    mask = None if maskOffsets is None else np.ascontiguousarray(maskOffsets, dtype=np.int32)
    cnames = None
    if names is not None:
        cnames = (ctypes.c_char_p * len(names))(*[s.encode() for s in names])
    bedFile.flush()
    lib = _lib.load()
    _lib.check(lib.tehmm_states_to_bed(bedFile.fileno(), chrom.encode(), start, _lib.ptr(st), n, _lib.ptr(seg),
                                       _lib.ptr(mask), 0 if mask is None else mask.shape[0], cnames,
                                       0 if names is None else len(names)))
