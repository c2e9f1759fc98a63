"""CPU: the native per-observation BED writer (tehmm_states_to_bed) against
a line-by-line restatement of the reference's statesToBed (bin/teHmmEval.py:238-262, bedFile
part).  Byte-exact output, including segment lengths, mask offsets and state names."""
import io
import os

import numpy as np
import pytest


class FakeTable(object):
    """the accessors statesToBed uses (track.py:434-447,497-502,650-662)"""

    def __init__(self, chrom, start, end, segOffsets=None, maskOffsets=None):
        self.chrom, self.start, self.end = chrom, start, end
        self.segOffsets, self.maskOffsets = segOffsets, maskOffsets

    def getChrom(self): return self.chrom
    def getStart(self): return self.start
    def getEnd(self): return self.end
    def getSegmentOffsets(self): return self.segOffsets
    def getMaskRunningOffsets(self): return self.maskOffsets

    def getSegmentLength(self, i):
        if i == len(self.segOffsets) - 1:
            return self.end - (self.start + self.segOffsets[-1])
        return self.segOffsets[i + 1] - self.segOffsets[i]


def reference_states_to_bed(trackTable, states, bedFile):
    """teHmmEval.py:238-262, bedFile branch, verbatim logic"""
    chrom, start = trackTable.getChrom(), trackTable.getStart()
    segOffsets = trackTable.getSegmentOffsets()
    maskOffsets = trackTable.getMaskRunningOffsets()
    segDist = 0
    for i in range(len(states)):
        curStart = start + segDist
        intLen = 1
        if segOffsets is not None:
            intLen = trackTable.getSegmentLength(i)
        segDist += intLen
        if maskOffsets is not None:
            curStart += maskOffsets[curStart - trackTable.getStart()]
        curEnd = curStart + intLen
        bedFile.write("%s\t%d\t%d\t%s\n" % (chrom, curStart, curEnd, states[i]))


def run_both(tmp_path, table, states, names=None, prefix=""):
    from tehmm_b200.output import statesToBed
    ref = io.StringIO()
    ref.write(prefix)
    mapped = states if names is None else [names[s] for s in states]
    reference_states_to_bed(table, mapped, ref)
    path = os.path.join(str(tmp_path), "out.bed")
    with open(path, "w") as f:
        f.write(prefix)
        statesToBed(table, states, f, stateNames=names)
        f.write("#after\n")
    got = open(path).read()
    assert got == ref.getvalue() + "#after\n"
    return got


def test_plain_table(tmp_path):
    rng = np.random.RandomState(0)
    states = rng.randint(0, 30, size=1000)
    out = run_both(tmp_path, FakeTable("chr1", 12345, 13345), states, prefix="track name=x\n")
    assert out.splitlines()[1] == "chr1\t12345\t12346\t%d" % states[0]


def test_segments_mask_and_names(tmp_path):
    rng = np.random.RandomState(1)
    n = 5000
    lens = rng.randint(1, 200, size=n)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    end = 777 + int(lens.sum())
    states = rng.randint(0, 4, size=n)
    names = ["LTR", "inside", "outside", "TSD|x"]
    run_both(tmp_path, FakeTable("scaffold_12", 777, end, segOffsets=offs), states, names=names)
    mask = np.cumsum(rng.rand(int(lens.sum())) < 0.01).astype(np.int32)
    run_both(tmp_path, FakeTable("scaffold_12", 777, end, segOffsets=offs, maskOffsets=mask), states, names=names)
    run_both(tmp_path, FakeTable("2L", 0, 3000, maskOffsets=mask[:3000]), rng.randint(0, 4, size=3000))


def test_many_blocks_and_name_input(tmp_path):
    """several 65536-line blocks formatted by several threads stay in order; states given as
    names (what MultitrackHmm.viterbi returns with a stateNameMap) are written as such"""
    from tehmm_b200.output import statesToBed
    rng = np.random.RandomState(2)
    n = 300_001
    states = rng.randint(0, 3, size=n)
    run_both(tmp_path, FakeTable("chrX", 10**9, 10**9 + n), states)
    named = np.array(["a", "bb", "c"])[states]
    path = os.path.join(str(tmp_path), "named.bed")
    with open(path, "w") as f:
        statesToBed(FakeTable("chrX", 5, 5 + n), list(named), f)
    lines = open(path).read().splitlines()
    assert len(lines) == n and lines[-1] == "chrX\t%d\t%d\t%s" % (5 + n - 1, 5 + n, named[-1])


def test_unnamed_state_is_an_error(tmp_path):
    from tehmm_b200.output import statesToBed
    with open(os.path.join(str(tmp_path), "x.bed"), "w") as f:
        with pytest.raises(AssertionError):
            statesToBed(FakeTable("c", 0, 3), [0, 1, 2], f, stateNames=["only", "two"])
