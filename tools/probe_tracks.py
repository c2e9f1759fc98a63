#!/usr/bin/env python
"""Throughput of the device-side input builders on one GPU, with the
reference's own loops timed on a bounded sample beside them (trackIO.readBedData from oracle/_ref; the
segmentation scan of bin/segmentTracks.py restated, it is a Python-2 script).

    python tools/probe_tracks.py [--T 10000000] > profiles/rNN_tracks.jsonl
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=10_000_000)
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    args = ap.parse_args()
    import torch
    from tehmm_b200 import trackIO, tracks_device
    from test_gpu_tracks import reference_segment_scan, runny_table, write_bed
    T, K = args.T, args.K
    rng = np.random.RandomState(0)
    tmp = tempfile.mkdtemp()
    names = ["LTR", "SINE", "LINE", "DNA", "Simple_repeat", "Low_complexity"]

    class Map(object):
        def __init__(self):
            self.m = {}

        def getMissingVal(self):
            return 0

        def getMap(self, v, update=False):
            if v not in self.m and update:
                self.m[v] = len(self.m) + 1
            return self.m.get(v, 0)
    # ---- rasterisation: K BED files of T / 100 intervals each
    n_iv = T // 100
    paths = []
    for k in range(K):
        p = os.path.join(tmp, "t%d.bed" % k)
        write_bed(p, rng, n_iv, 0, T, names)
        paths.append(p)
    d = tracks_device.new_table(T, K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vals = [trackIO.bed_interval_values_native(p, "chr1", 0, T, 3, Map(), True, True, False) for p in paths]
    t1 = t2 = time.perf_counter()
    for k, (s, e, v, v0) in enumerate(vals):
        tracks_device.fill_column(d, k, 0)
        tracks_device.rasterize(d, k, s, e, np.asarray(v, dtype=np.int32), np.asarray(v0, dtype=np.int32), 0, T)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    out = {"what": "rasterisation", "bases": T, "tracks": K, "intervals_per_track": n_iv,
           "parse_and_value_map_s": t1 - t0, "device_fill_s": t3 - t2,
           "bases_x_tracks_per_s_fill": T * K / (t3 - t2), "bases_x_tracks_per_s_total": T * K / (t3 - t0)}
    try:
        import ref_loader
        R = ref_loader.load()
        tio = importlib.import_module("teHmm.trackIO")
        n = args.cpu_sample
        p = os.path.join(tmp, "s.bed")
        write_bed(p, rng, n // 100, 0, n, names)
        t0 = time.perf_counter()
        tio.readBedData(p, "chr1", 0, n, valCol=3, valMap=R.track.CategoryMap(reserved=1), updateValMap=True,
                        needIntersect=False, outputBuf=np.zeros(n, dtype=np.uint8))
        dt = time.perf_counter() - t0
        out["reference_readBedData_bases_per_s"] = n / dt
        out["reference_sample"] = "%d bases, one track, no bedtools call" % n
    except Exception as ex:      # pragma: no cover
        out["reference_readBedData_bases_per_s"] = None
        out["reference_error"] = repr(ex)
    print(json.dumps(out), flush=True)

    # ---- segmentation + compression of a run-length table
    data = runny_table(rng, T, K, [0.01, 0.02, 0.005, 0.05, 0.002, 0.03, 0.01, 0.01, 0.02, 0.004][:K])
    d = torch.from_numpy(data).cuda()
    a = types.SimpleNamespace(comp="first", thresh=1, maxLen=100, fixLen=0, ignoreList=[0] * K, cutList=[0] * K)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d_cut, d_off, passes = tracks_device.segment(d, [0, T], a.ignoreList, a.cutList, a.thresh, a.maxLen, a.fixLen)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        d_small = tracks_device.compress(d, d_off, np.ones(K, dtype=np.uint8))
        torch.cuda.synchronize(); t2 = time.perf_counter()
    n = args.cpu_sample
    t0 = time.perf_counter()
    want = reference_segment_scan(data[:n], a)
    dt = time.perf_counter() - t0
    got = d_off.cpu().numpy()
    print(json.dumps({"what": "segmentation (--comp first --thresh 1 --maxLen 100) + compression (per-track mode)", "bases": T,
                      "tracks": K, "segments": int(d_off.shape[0]), "scan_passes": passes,
                      "prefix_equals_reference_scan": bool(np.array_equal(got[:len(want) - 1], want[:-1])),
                      "reference_scan_bases_per_s": n / dt, "reference_sample": "%d bases, restated Python loop" % n}), flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d_cut, d_off, passes = tracks_device.segment(d, [0, T], a.ignoreList, a.cutList, a.thresh, a.maxLen, a.fixLen)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d_small = tracks_device.compress(d, d_off, np.ones(K, dtype=np.uint8))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({"what": "timing of the above", "segment_ms": 1e3 * (t1 - t0), "compress_ms": 1e3 * (t2 - t1),
                      "bases_per_s_segment": T / (t1 - t0), "bases_per_s_compress": T / (t2 - t1)}), flush=True)


if __name__ == "__main__":
    main()
