"""Device-side builders of the HMM's input (csrc/tracks.cu):
interval rasterisation, segmentation, segment compression, runSum.  Thin wrappers over the C ABI
on torch device tensors; the reference-named host APIs (trackIO.readBedData, segmentTracks,
IntegerTrackTable.segment, _track.runSum) are built on these."""
import ctypes

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _ctx(device=None):
    return _lib.get_context(device)


def _bind(ctx):
    torch = _torch()
    torch.cuda.set_device(ctx.device)
    s = torch.cuda.current_stream(ctx.device).cuda_stream
    _lib.check(ctx.lib.tehmm_ctx_set_stream(ctx.handle, ctypes.c_uint64(s)))


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def new_table(T, K, dtype=np.uint8, device=None):
    """(T, K) zero table on the device, the layout of IntegerTrackTable.data (track.py:552-558)"""
    torch = _torch()
    ctx = _ctx(device)
    tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}[np.dtype(dtype).itemsize]
    return torch.zeros((int(T), int(K)), dtype=tdt, device=torch.device("cuda", ctx.device))


def fill_column(d_table, k, value, device=None):
    ctx = _ctx(device)
    _bind(ctx)
    T, K = d_table.shape
    _lib.check(ctx.lib.tehmm_track_fill(ctx.handle, _p(d_table), int(T), int(K), d_table.element_size(), int(k), int(value)))


def rasterize(d_table, k, starts, ends, vals, vals0, region_start, region_end, device=None):
    """paint n intervals (genome coordinates, file order: later wins) into column k"""
    ctx = _ctx(device)
    _bind(ctx)
    T, K = d_table.shape
    assert T == region_end - region_start
    s = np.ascontiguousarray(starts, dtype=np.int64)
    e = np.ascontiguousarray(ends, dtype=np.int64)
    v = np.ascontiguousarray(vals, dtype=np.int32)
    v0 = np.ascontiguousarray(vals0, dtype=np.int32)
    assert s.shape == e.shape == v.shape == v0.shape
    _lib.check(ctx.lib.tehmm_rasterize_intervals(ctx.handle, _lib.ptr(s), _lib.ptr(e), _lib.ptr(v), _lib.ptr(v0), int(s.shape[0]),
                                                 int(region_start), int(region_end), _p(d_table), int(K), d_table.element_size(), int(k)))


def segment(d_table, region_off, ignore=None, cut=None, thresh=1, maxLen=0, fixLen=0, prev_mode=False, device=None):
    """segmentTracks.py:200-277 on a (T, K) device table made of len(region_off)-1 regions.
    Returns (d_cut uint8[T], d_seg_off int64[nseg], passes)."""
    torch = _torch()
    ctx = _ctx(device)
    _bind(ctx)
    T, K = d_table.shape
    ro = np.ascontiguousarray(region_off, dtype=np.int64)
    ig = None if ignore is None else np.ascontiguousarray(ignore, dtype=np.uint8)
    ct = None if cut is None else np.ascontiguousarray(cut, dtype=np.uint8)
    dev = d_table.device
    d_cut = torch.zeros(int(T), dtype=torch.uint8, device=dev)
    d_off = torch.empty(int(T), dtype=torch.int64, device=dev)
    nseg = ctypes.c_int64(0)
    passes = ctypes.c_int(0)
    _lib.check(ctx.lib.tehmm_segment_table(ctx.handle, _p(d_table), int(T), int(K), d_table.element_size(), int(ro.shape[0] - 1),
                                           _lib.ptr(ro), _lib.ptr(ig), _lib.ptr(ct), int(thresh), int(maxLen), int(fixLen),
                                           1 if prev_mode else 0, _p(d_cut), _p(d_off), ctypes.byref(nseg), ctypes.byref(passes)))
    return d_cut, d_off[:nseg.value], passes.value


def compress(d_table, d_seg_off, use_mode=None, device=None):
    """one row per segment: per-track mode (interpolateSegments) of the tracks flagged in use_mode,
    first row of the segment otherwise (compressSegments).  Returns the (nseg, K) device table."""
    torch = _torch()
    ctx = _ctx(device)
    _bind(ctx)
    T, K = d_table.shape
    nseg = int(d_seg_off.shape[0])
    um = None if use_mode is None else np.ascontiguousarray(use_mode, dtype=np.uint8)
    d_out = torch.empty((nseg, int(K)), dtype=d_table.dtype, device=d_table.device)
    _lib.check(ctx.lib.tehmm_compress_segments(ctx.handle, _p(d_table), int(T), int(K), d_table.element_size(), _p(d_seg_off),
                                               nseg, _lib.ptr(um), _p(d_out)))
    return d_out


def run_sum(d_mask, device=None):
    torch = _torch()
    ctx = _ctx(device)
    _bind(ctx)
    n = int(d_mask.shape[0])
    d_out = torch.empty(n, dtype=torch.int32, device=d_mask.device)
    _lib.check(ctx.lib.tehmm_run_sum(ctx.handle, _p(d_mask), _p(d_out), n))
    return d_out


def segment_ratios(d_seg_off, region_off, effective_len):
    """per-segment length / effectiveSegmentLength (track.py:504-513), on the device; region_off
    gives the row where each region ends (the last segment of a region runs to the region's end)"""
    torch = _torch()
    ends = torch.empty_like(d_seg_off)
    ends[:-1] = d_seg_off[1:]
    ends[-1] = int(region_off[-1])
    # (a tensor divisor: torch turns division by a Python scalar into a multiplication by its reciprocal,
    #  which is an ulp off the reference's float(segLen) / effectiveSegmentLength)
    div = torch.full((1,), float(effective_len), dtype=torch.float64, device=d_seg_off.device)
    return torch.div((ends - d_seg_off).to(torch.float64), div)
