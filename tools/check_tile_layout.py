"""Host-side check of the register layout used by csrc/tile.cu: emulate
mma.sync.m16n8k8 (row.col) fragment semantics per the PTX ISA and verify that
feeding the accumulator fragments back as A-operand fragments with the permuted
transition fragments reproduces X @ A (forward) and W @ A.T (backward)."""
import numpy as np

def mma(acc, a, b):
    """acc[lane][4], a[lane][4], b[lane][2] -> acc += A(16x8) @ B(8x8) per the fragment layouts"""
    A = np.zeros((16, 8)); B = np.zeros((8, 8)); C = np.zeros((16, 8))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        A[g, q] = a[lane][0]; A[g + 8, q] = a[lane][1]; A[g, q + 4] = a[lane][2]; A[g + 8, q + 4] = a[lane][3]
        B[q, g] = b[lane][0]; B[q + 4, g] = b[lane][1]
    C = A @ B
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        acc[lane][0] += C[g, 2 * q]; acc[lane][1] += C[g, 2 * q + 1]
        acc[lane][2] += C[g + 8, 2 * q]; acc[lane][3] += C[g + 8, 2 * q + 1]

def run(backward):
    rs = np.random.RandomState(1)
    T = rs.rand(32, 32)
    X = rs.rand(16, 32)
    # lane registers x[r][i] <-> row g+8r, state 8q+i
    x = np.zeros((32, 2, 8))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for r in range(2):
            for i in range(8):
                x[lane, r, i] = X[g + 8 * r, 8 * q + i]
    frag = np.zeros((32, 4, 4, 2))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for kt in range(4):
            for nt in range(4):
                sn = 8 * (g >> 1) + 2 * nt + (g & 1)
                sk = 8 * q + 2 * kt
                for e in range(2):
                    frag[lane, kt, nt, e] = T[sn, sk + e] if backward else T[sk + e, sn]
    d = np.zeros((32, 2, 8))
    for nt in range(4):
        acc = np.zeros((32, 4))
        for kt in range(4):
            a = [[x[l, 0, 2 * kt], x[l, 1, 2 * kt], x[l, 0, 2 * kt + 1], x[l, 1, 2 * kt + 1]] for l in range(32)]
            bb = [[frag[l, kt, nt, 0], frag[l, kt, nt, 1]] for l in range(32)]
            mma(acc, a, bb)
        for l in range(32):
            d[l, 0, 2 * nt] = acc[l][0]; d[l, 0, 2 * nt + 1] = acc[l][1]
            d[l, 1, 2 * nt] = acc[l][2]; d[l, 1, 2 * nt + 1] = acc[l][3]
    D = np.zeros((16, 32))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for r in range(2):
            for i in range(8):
                D[g + 8 * r, 8 * q + i] = d[lane, r, i]
    ref = X @ T.T if backward else X @ T
    err = np.abs(D - ref).max()
    print("backward" if backward else "forward", "max err", err)
    assert err < 1e-12

run(False); run(True)
