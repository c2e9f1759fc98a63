"""Minimal host-side track containers for the hot path.

The hot path needs the observation matrix and, for segmented tables, the
per-observation segment lengths; IntegerTrackTable.segment
compresses a table to one row per segment on the GPU.  Anything that quacks like the reference's
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

    def segment(self, segIntervals, trackList, interpolate=True):
        """Transform the table to one row per segment interval (track.py:449-495): the offsets of the
        segment intervals that fall into this table, the per-track mode of every segment written to
        its first row (interpolateSegments / setAverages, track.py:515-533,603-620), then
        compressSegments -- the last two on the GPU (tehmm_compress_segments).  segIntervals: sorted
        (chrom, start, end, ...) tuples covering the table; trackList: iterable of tracks with
        getNumber() / getDist(), or None (every track takes the mode).  Masked tables and gaussian
        tracks (mean of mapped-back values) are not covered."""
        import torch
        from . import tracks_device
        assert self.maskArray is None, "segment() on a masked table is not implemented"
        offs = [int(iv[1]) - self.start for iv in segIntervals
                if iv[0] == self.chrom and int(iv[1]) >= self.start and int(iv[1]) < self.origEnd]
        assert len(offs) > 0 and offs[0] == 0, "segment intervals must start where the table starts"
        self.segOffsets = np.asarray(offs, dtype=np.int64)
        K = self.numTracks
        use_mode = np.zeros(K, dtype=np.uint8)
        if interpolate:
            use_mode[:] = 1
            if trackList is not None:
                for track in trackList:
                    if track.getDist() == "gaussian":
                        raise NotImplementedError("gaussian tracks are averaged on the host in the reference (track.py:607-617)")
        dev = torch.device("cuda", tracks_device._ctx().device)
        d = torch.from_numpy(np.ascontiguousarray(self.data)).to(dev)
        d_off = torch.from_numpy(self.segOffsets).to(dev)
        self.data = tracks_device.compress(d, d_off, use_mode).cpu().numpy()
        self.shape = (len(self), self.numTracks)

    def compressSegments(self):
        """one row per segment offset (track.py:594-601)"""
        assert self.segOffsets is not None and len(self.segOffsets) > 0
        self.data = self.data[self.segOffsets]

    def setSegments(self, segOffsets):
        """Keep one observation per segment (track.py:594-601 compressSegments)."""
        self.segOffsets = np.asarray(segOffsets, dtype=np.int64)
        self.data = self.data[self.segOffsets]
        self.shape = (len(self), self.numTracks)
