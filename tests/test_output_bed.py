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


def reference_scores_to_bed(trackTable, n, posteriors, posteriorsMask, posteriorsFile, emProbs, emissionsMask,
                            emissionsFile):
    """teHmmEval.py:238-270, the posteriorsFile / emissionsFile branches, verbatim logic (incl. the i-1 index)"""
    chrom, start = trackTable.getChrom(), trackTable.getStart()
    segOffsets = trackTable.getSegmentOffsets()
    maskOffsets = trackTable.getMaskRunningOffsets()
    segDist = 0
    for i in range(n):
        curStart = start + segDist
        intLen = 1
        if segOffsets is not None:
            intLen = trackTable.getSegmentLength(i)
        segDist += intLen
        if maskOffsets is not None:
            curStart += maskOffsets[curStart - trackTable.getStart()]
        curEnd = curStart + intLen
        if posteriors is not None:
            posteriorsFile.write("%s\t%d\t%d\t%s\n" % (chrom, curStart, curEnd,
                                 np.sum(posteriors[i-1] * posteriorsMask)))
        if emProbs is not None:
            emissionsFile.write("%s\t%d\t%d\t%s\n" % (chrom, curStart, curEnd,
                                np.log(np.sum(np.exp(emProbs[i-1]) * emissionsMask))))


@pytest.mark.parametrize("N", [3, 8, 30, 50])
def test_posterior_and_emission_score_files(tmp_path, N):
    from tehmm_b200.output import statesToBed
    rng = np.random.RandomState(N)
    n = 4000
    post = rng.dirichlet(np.ones(N) * 0.3, size=n)
    post[5] = 0.0; post[5, 1] = 1.0                       # exact 1.0 / 0.0 sums
    post[6] = 1e-7 / N                                    # scientific notation
    em = np.log(rng.dirichlet(np.ones(N), size=n)) * rng.choice([1.0, 30.0], size=(n, 1))
    pmask = (rng.rand(N) < 0.5).astype(np.float64)
    pmask[1] = 1.0
    emask = np.zeros(N); emask[[0, N - 1]] = 1.0
    offs = np.concatenate([[0], np.cumsum(rng.randint(1, 50, size=n - 1))])
    table = FakeTable("chr7", 1000, 1000 + int(offs[-1]) + 13, segOffsets=offs)
    rp, re_ = io.StringIO(), io.StringIO()
    reference_scores_to_bed(table, n, post, pmask, rp, em, emask, re_)
    pp, pe, pb = (os.path.join(str(tmp_path), x) for x in ("p.bed", "e.bed", "s.bed"))
    states = rng.randint(0, N, size=n)
    with open(pp, "w") as fp, open(pe, "w") as fe, open(pb, "w") as fb:
        statesToBed(table, states, fb, post, pmask, fp, em, emask, fe)
    assert open(pp).read() == rp.getvalue()
    assert open(pe).read() == re_.getvalue()
    rb = io.StringIO()
    reference_states_to_bed(table, states, rb)
    assert open(pb).read() == rb.getvalue()
    # bedFile may be None (teHmmEval.py:263)
    with open(pp, "w") as fp:
        statesToBed(table, states, None, post, pmask, fp)
    assert open(pp).read() == rp.getvalue()


def test_python_float_formatting_matches(tmp_path):
    """fourth column == "%s" % np.float64(x) for awkward values"""
    from tehmm_b200.output import _write_scores
    from tehmm_b200 import _lib
    rng = np.random.RandomState(1)
    vals = np.concatenate([
        [0.0, -0.0, 1.0, -1.0, 0.1, 0.2 + 0.1, 1e-4, 9.999e-5, 1e-5, 1e15, 1e16, 123456789012345678.0, 1e22, 1e23,
         5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, np.inf, -np.inf, np.nan, 100.0, 1e5, 0.5, 2.5e-7,
         1.0 - 2 ** -53, 1.0 + 2 ** -52, 4.35, 0.3, 1 / 3, 2 / 3, 1e-100, 123.456, 9007199254740993.0],
        rng.rand(3000), np.exp(rng.uniform(-50, 50, size=3000)), -rng.rand(100) * 1e-6])
    path = os.path.join(str(tmp_path), "v.bed")
    with open(path, "w") as f:
        _write_scores(_lib.load(), f, "c", 0, vals, None, None)
    got = [l.split("\t")[3] for l in open(path).read().splitlines()]
    want = ["%s" % v for v in vals]
    assert got == want
