"""Track file reading with the reference's names (/root/reference/trackIO.py), the per-base fill
on the GPU.

readBedData's per-interval work -- column choice, delta values, the value map (strings ->
categories, which mutates the map) -- is host work and follows trackIO.py:175-197 line by line;
its per-BASE loop (trackIO.py:198-202, the reason rasterising a genome takes hours in the
reference, README.md:341) is one call to tehmm_rasterize_intervals.  bedtools is not used: the
intersection with the query interval and the sort are done here.
"""
import numpy as np

from . import tracks_device


def bedRead(filePath):
    """list of tab-split BED lines with at least three columns (trackIO.py:389-404)"""
    intervals = []
    with open(filePath, "r") as bf:
        for line in bf:
            if len(line) > 0 and line[0] != "#":
                interval = line.rstrip("\n").split("\t")
                if len(interval) > 2:
                    intervals.append(interval)
    return intervals


def _intersect_sort(rows, chrom, start, end, need_intersect, sort):
    """what `intersectBed -a file -b interval | sortBed` (trackIO.py:138-143) or `sortBed`
    (trackIO.py:147-151) leaves: overlapping parts only, clipped, by (chrom, start)"""
    if need_intersect:
        out = []
        for r in rows:
            if r[0] != chrom:
                continue
            s, e = int(r[1]), int(r[2])
            if s < end and e > start:
                r = list(r)
                r[1], r[2] = str(max(s, start)), str(min(e, end))
                out.append(r)
        out.sort(key=lambda r: int(r[1]))         # stable: file order among equal starts
        return out
    if sort:
        return sorted(rows, key=lambda r: (r[0], int(r[1])))
    return rows


def bed_interval_values(intersections, valCol=None, valMap=None, updateMap=False, useDelta=False):
    """(starts, ends, val, val0) per interval: trackIO.py:175-197, in file order (the value map is
    updated in that order, so category numbers come out as in the reference)"""
    n = len(intersections)
    starts = np.empty(n, dtype=np.int64)
    ends = np.empty(n, dtype=np.int64)
    vals, vals0 = [None] * n, [None] * n
    prevInterval, prevVal = None, 0
    for i, overlap in enumerate(intersections):
        starts[i], ends[i] = int(overlap[1]), int(overlap[2])
        if valCol is not None:
            if valCol == 0:
                val = 1
            elif valCol == 4:
                assert overlap[4] is not None and overlap[4] != ""
                val = overlap[4]
            else:
                assert valCol == 3
                assert overlap[3] is not None and overlap[3] != ""
                val = overlap[3]
        else:
            val = overlap[3]
        val0 = val
        if useDelta is True:
            if prevInterval is not None and int(overlap[1]) == int(prevInterval[2]) and prevInterval[0] == overlap[0]:
                try:
                    val0 = float(val) - float(prevVal)
                except Exception:
                    val0 = int(val != prevVal)
            prevVal = val
            prevInterval = overlap
            val = 0
        if valMap is not None:
            val = valMap.getMap(val, update=updateMap)
            val0 = valMap.getMap(val0, update=updateMap)
        vals[i], vals0[i] = val, val0
    return starts, ends, vals, vals0


def bed_interval_values_native(bedPath, chrom, start, end, valCol, valMap, updateMap, needIntersect, sort):
    """the same four arrays without a Python loop over the intervals: the file is parsed,
    intersected and sorted natively (tehmm_bed_open) and only the DISTINCT value strings go through
    the value map, in order of first appearance -- which is the order the reference's loop feeds
    them to it (trackIO.py:175-197 without useDelta)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    vc = 0 if valCol == 0 else (3 if valCol is None else int(valCol))
    _lib.check(lib.tehmm_bed_open(bedPath.encode(), chrom.encode(), int(start), int(end), 1 if needIntersect else 0,
                                  1 if sort else 0, vc, ctypes.byref(h)))
    try:
        n = int(lib.tehmm_bed_count(h))
        s, e, vi = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int32)
        _lib.check(lib.tehmm_bed_fetch(h, _lib.ptr(s), _lib.ptr(e), _lib.ptr(vi)))
        if vc == 0:
            one = valMap.getMap(1, update=updateMap) if valMap is not None else 1
            v = np.full(n, _as_int(one), dtype=np.int64)
        else:
            nu = int(lib.tehmm_bed_nunique(h))
            lut = np.empty(max(nu, 1), dtype=np.int64)
            for i in range(nu):
                u = lib.tehmm_bed_unique(h, i).decode()
                lut[i] = _as_int(valMap.getMap(u, update=updateMap) if valMap is not None else u)
            v = lut[vi] if n else np.empty(0, dtype=np.int64)
    finally:
        lib.tehmm_bed_close(h)
    return s, e, v, v


def _as_int(v):
    if v is None:
        return 0
    try:
        return int(v)
    except (TypeError, ValueError):
        return int(float(v))


def rasterizeBedData(d_table, k, bedPath, chrom, start, end, **kwargs):
    """readBedData straight into column k of a (end - start, K) DEVICE table: nothing but the
    interval list crosses PCIe.  Keyword arguments as readBedData (no outputBuf)."""
    valCol = int(kwargs["valCol"]) if kwargs.get("valCol") is not None else None
    valMap = kwargs.get("valMap")
    if kwargs.get("useDelta", False) is not True and valCol in (None, 0, 3, 4):
        s, e, vi, vi0 = bed_interval_values_native(bedPath, chrom, start, end, valCol, valMap, kwargs.get("updateValMap", False),
                                                   kwargs.get("needIntersect", True), kwargs.get("sort", False) is True)
    else:       # delta values depend on the previous interval: the reference's loop, interval by interval
        rows = _intersect_sort(bedRead(bedPath), chrom, start, end, kwargs.get("needIntersect", True),
                               kwargs.get("sort", False) is True)
        s, e, v, v0 = bed_interval_values(rows, valCol, valMap, kwargs.get("updateValMap", False), kwargs.get("useDelta", False))
        vi = np.array([_as_int(x) for x in v], dtype=np.int64)
        vi0 = np.array([_as_int(x) for x in v0], dtype=np.int64)
    keep = (np.minimum(e, end) > np.maximum(s, start))      # the reference indexes out of range for the others
    info = np.iinfo({1: np.uint8, 2: np.int16, 4: np.int32}[d_table.element_size()])
    vi, vi0 = np.clip(vi, info.min, info.max), np.clip(vi0, info.min, info.max)
    tracks_device.rasterize(d_table, k, s[keep], e[keep], vi[keep], vi0[keep], start, end)
    return int(np.sum(np.minimum(e, end)[keep] - np.maximum(s, start)[keep]))


def readBedData(bedPath, chrom, start, end, **kwargs):
    """Read a bed file into an array with one entry per base (trackIO.py:67-212): same arguments
    (valCol, valMap, updateValMap, sort, needIntersect, useDelta, outputBuf), same result -- the
    outputBuf when one is given, else a list with the value map's missing value where no interval
    lies.  ignoreBed12=False (bed12ToBed6) is not supported."""
    if kwargs.get("ignoreBed12", True) is not True:
        raise NotImplementedError("ignoreBed12=False needs bedtools' bed12ToBed6")
    valMap = kwargs.get("valMap")
    defVal = valMap.getMissingVal() if valMap is not None else None
    outputBuf = kwargs.get("outputBuf")
    n = end - start
    dtype = outputBuf.dtype if outputBuf is not None else np.int32
    if np.dtype(dtype).kind not in "iu" or np.dtype(dtype).itemsize > 4:
        dtype = np.int32
    d = tracks_device.new_table(n, 1, dtype={1: np.uint8, 2: np.int16, 4: np.int32}[np.dtype(dtype).itemsize])
    sentinel = None
    if outputBuf is not None:
        import torch
        d[:, 0] = torch.from_numpy(np.ascontiguousarray(outputBuf).astype(dtype, copy=False).view(
            {1: np.uint8, 2: np.int16, 4: np.int32}[np.dtype(dtype).itemsize])).to(d.device)
    else:
        sentinel = np.iinfo(np.int32).min
        tracks_device.fill_column(d, 0, sentinel)
    kw = dict(kwargs)
    kw.pop("outputBuf", None)
    rasterizeBedData(d, 0, bedPath, chrom, start, end, **kw)
    host = d[:, 0].cpu().numpy()
    if outputBuf is not None:
        outputBuf[:] = host.view(outputBuf.dtype) if host.dtype.itemsize == outputBuf.dtype.itemsize else host
        return outputBuf
    return [defVal if x == sentinel else int(x) for x in host]
