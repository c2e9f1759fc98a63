"""bin/segmentTracks.py:200-277 with the per-base scan on the GPU.

segmentTracks(trackData, args, stats) keeps the script's signature and output (one BED line per
segment, hexadecimal running count as the name); isNewSegment keeps its own for the callers
that probe single columns.  The scan itself -- isNewSegment for every base of every region, a
Python loop in the reference -- is tehmm_segment_table (csrc/tracks.cu).
"""
import numpy as np

from . import tracks_device


def isNewSegment(trackTable, pi, i, curLen, args, stats):
    """host restatement for single columns (segmentTracks.py:241-277); the bulk scan does not call it"""
    assert i > 0 and i < len(trackTable) and pi >= 0 and pi < i
    if args.fixLen > 0:
        return curLen >= args.fixLen
    if args.maxLen > 0 and curLen >= args.maxLen:
        return True
    col, prev = np.asarray(trackTable[i]), np.asarray(trackTable[pi])
    dif = (col != prev) & (np.asarray(args.ignoreList) == 0)
    difCount = int(dif.sum())
    cutTrackFound = bool((dif & (np.asarray(args.cutList) == 1)).any())
    retVal = cutTrackFound or difCount > args.thresh
    if args.stats is not None and retVal:
        for j in np.nonzero(dif)[0]:
            c, p = stats.get(int(j), (0, 0))
            stats[int(j)] = (c + 1, p + 1. / float(difCount))
    return retVal


def segment_tables(tables, args, device_tables=None):
    """cut positions of every table of a list: [(table, ndarray of segment start rows)].
    tables: objects with getNumPyArray() (host) -- or pass the same data already on the device
    as one concatenated (sum T, K) tensor in device_tables."""
    import torch
    lens = [len(t) for t in tables]
    region_off = np.zeros(len(tables) + 1, dtype=np.int64)
    np.cumsum(lens, out=region_off[1:])
    if device_tables is None:
        host = np.concatenate([np.ascontiguousarray(t.getNumPyArray()) for t in tables], axis=0)
        if host.dtype.itemsize > 4 or host.dtype.kind not in "iu":
            host = host.astype(np.int32)
        view = {1: np.uint8, 2: np.int16, 4: np.int32}[host.dtype.itemsize]
        d = torch.from_numpy(host.view(view)).to(torch.device("cuda", tracks_device._ctx().device))
    else:
        d = device_tables
    K = d.shape[1]
    ignore = np.zeros(K, dtype=np.uint8) if getattr(args, "ignoreList", None) is None else np.asarray(args.ignoreList, dtype=np.uint8)
    cut = np.zeros(K, dtype=np.uint8) if getattr(args, "cutList", None) is None else np.asarray(args.cutList, dtype=np.uint8)
    d_cut, d_off, passes = tracks_device.segment(d, region_off, ignore, cut, int(args.thresh), int(args.maxLen), int(args.fixLen),
                                                 getattr(args, "comp", "first") == "prev")
    off = d_off.cpu().numpy()
    out = []
    for r, t in enumerate(tables):
        a, b = np.searchsorted(off, region_off[r]), np.searchsorted(off, region_off[r + 1])
        out.append((t, off[a:b] - region_off[r]))
    return out, d, d_off, region_off


def segmentTracks(trackData, args, stats):
    """write args.outBed like the script (segmentTracks.py:200-239)"""
    tables = trackData.getTrackTableList()
    per_table, d, d_off, region_off = segment_tables(tables, args)
    count = int(args.co)
    with open(args.outBed, "w") as oFile:
        for table, starts in per_table:
            chrom, start, end = table.getChrom(), table.getStart(), table.getEnd()
            ends = np.append(starts[1:], end - start)
            lines = ["%s\t%d\t%d\t%s\n" % (chrom, start + int(a), start + int(b), hex(count + q)[2:])
                     for q, (a, b) in enumerate(zip(starts, ends))]
            oFile.write("".join(lines))
            count += len(starts)
            if args.stats is not None and args.fixLen <= 0:
                _cut_stats(table, starts, args, stats)
    return count


def _cut_stats(table, starts, args, stats):
    """the statistics pass of isNewSegment (segmentTracks.py:268-274) for the cuts the difference rule made"""
    data = np.asarray(table.getNumPyArray())
    prev_mode = getattr(args, "comp", "first") == "prev"
    ign = np.asarray(args.ignoreList) == 0
    cutl = np.asarray(args.cutList) == 1
    for q in range(1, len(starts)):
        i = int(starts[q])
        pi = i - 1 if prev_mode else int(starts[q - 1])
        if args.maxLen > 0 and i - int(starts[q - 1]) >= args.maxLen:
            continue                                  # cut by length: isNewSegment returned before the statistics
        dif = (data[i] != data[pi]) & ign
        difCount = int(dif.sum())
        if not ((dif & cutl).any() or difCount > args.thresh):
            continue
        for j in np.nonzero(dif)[0]:
            c, p = stats.get(int(j), (0, 0))
            stats[int(j)] = (c + 1, p + 1. / float(difCount))
