"""Minimal host-side track containers for the hot path.

The reference's data model (track.py) is out of scope;
the hot path only needs the observation matrix and, for segmented tables, the
per-observation segment lengths.  Anything that quacks like the reference's
TrackTable (getNumPyArray / getSegmentOffsets / getSegmentLengthsAsRatio) is
accepted, so the reference's own TrackData objects work unchanged.
"""
import numpy as np

INTEGER_ARRAY_TYPE = np.uint8


def is_track_table(obs):
    return hasattr(obs, "getNumPyArray") and hasattr(obs, "getSegmentOffsets")


class TrackTable(object):
    """Interval [start, end) of a chromosome over numTracks tracks (track.py:353-513)."""

    def __init__(self, numTracks, chrom, start, end):
        assert end > start
        self.numTracks = numTracks
        self.chrom = chrom
        self.start = start
        self.end = end
        self.origEnd = end
        self.segOffsets = None
        self.shape = (len(self), numTracks)

    def __len__(self):
        return self.end - self.start if self.segOffsets is None else len(self.segOffsets)

    def getNumTracks(self):
        return self.numTracks

    def getChrom(self):
        return self.chrom

    def getStart(self):
        return self.start

    def getEnd(self):
        return self.end

    def getSegmentOffsets(self):
        return self.segOffsets

    def getSegmentLength(self, i):
        if i == len(self.segOffsets) - 1:
            return self.end - (self.start + self.segOffsets[-1])
        return self.segOffsets[i + 1] - self.segOffsets[i]

    def getSegmentLengthsAsRatio(self, effectiveSegmentLength):
        """segment length / effective length per observation (track.py:504-513)."""
        if self.segOffsets is None:
            return None
        eff = float(effectiveSegmentLength)
        assert eff >= 1
        offs = np.asarray(self.segOffsets, dtype=np.int64)
        ends = np.append(offs[1:], self.end - self.start)
        return (ends - offs).astype(np.float64) / eff

    def getOverlapInTableCoords(self, bedInterval, startHint=None):
        """Overlap with a BED interval, in table coordinates (track.py:390-432)."""
        chrom, start, end = bedInterval[0], bedInterval[1], bedInterval[2]
        if self.chrom != chrom or not (self.start < end and self.end > start):
            return None
        lo, hi = max(self.start, start), min(self.end, end)
        out = [self.chrom, lo, hi] + list(bedInterval[3:])
        if self.segOffsets is None:
            out[1] = lo - self.start
            out[2] = hi - self.start
            return out
        offs = self.start + np.asarray(self.segOffsets, dtype=np.int64)
        first = int(np.searchsorted(offs, lo, side="right") - 1)
        last = int(np.searchsorted(offs, hi, side="left") - 1)
        assert first >= 0 and last >= first
        out[1] = first
        out[2] = last + 1
        return out

    def getNumPyArray(self):
        raise RuntimeError("Not implemented")


class IntegerTrackTable(TrackTable):
    """(end-start) x numTracks integer matrix, C order, uint8 by default (track.py:546-583)."""

    def __init__(self, numTracks, chrom, start, end, dtype=INTEGER_ARRAY_TYPE):
        super(IntegerTrackTable, self).__init__(numTracks, chrom, start, end)
        self.data = np.zeros((end - start, numTracks), dtype=dtype)
        self.iinfo = np.iinfo(dtype)
        self.maskArray = None

    def __getitem__(self, index):
        return self.data[index]

    def writeRow(self, row, rowArray):
        assert row < self.getNumTracks()
        assert len(rowArray) == len(self)
        self.data[:, row] = np.clip(np.asarray(rowArray), self.iinfo.min, self.iinfo.max)

    def getRow(self, row):
        return self.data[:, row]

    def initRow(self, row, val):
        self.data[:, row] = val

    def getNumPyArray(self):
        return self.data

    def setSegments(self, segOffsets):
        """Keep one observation per segment (track.py:594-601 compressSegments)."""
        self.segOffsets = np.asarray(segOffsets, dtype=np.int64)
        self.data = self.data[self.segOffsets]
        self.shape = (len(self), self.numTracks)
