"""Lines per second of the per-observation BED writer: native (tehmm_states_to_bed) vs the
reference's Python loop (teHmmEval.py:238-262), CPU only."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_output_bed import FakeTable, reference_states_to_bed
from tehmm_b200.output import statesToBed
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
states = np.random.RandomState(0).randint(0, 30, size=n)
tab = FakeTable("chr1", 0, n)
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    with open(os.path.join(d, "a.bed"), "w") as f:
        statesToBed(tab, states, f)
    t1 = time.perf_counter()
    m = min(n, 500_000)
    with open(os.path.join(d, "b.bed"), "w") as f:
        reference_states_to_bed(FakeTable("chr1", 0, m), states[:m], f)
    t2 = time.perf_counter()
    print("native: %d lines in %.3f s = %.2e lines/s   python loop: %d lines in %.3f s = %.2e lines/s   (%.0fx)" % (
        n, t1 - t0, n / (t1 - t0), m, t2 - t1, m / (t2 - t1), (n / (t1 - t0)) / (m / (t2 - t1))))
