"""Deterministic synthetic models and observation matrices.

Used by bench.py and the parity tests; host-side NumPy only.
  model: sticky Dirichlet transitions with exact zeros (exercises LOGZERO),
         uniform start, Dirichlet(0.5) per-(track,state) emissions, symbol 0 =
         missing (log-prob 0);
  obs:   sampled FROM the model with run-length structure, uint8, a fraction of
         entries overwritten with the missing symbol.
"""
import numpy as np

from .common import myLog

BENCH_SYMS = (4, 8, 16, 32, 64, 250, 2, 2, 2, 2)   # 6 multinomial + 4 binary tracks


def make_model(N=30, syms=BENCH_SYMS, seed=0, sticky=0.9, zero_frac=0.2, uniform_start=True):
    """Returns dict(pi, A, table (K,N,S) float64 log-probs, log_start, log_trans, syms, widths)."""
    rng = np.random.RandomState(seed)
    A = rng.dirichlet(np.ones(N), size=N)
    A = sticky * np.eye(N) + (1.0 - sticky) * A
    if N > 2 and zero_frac > 0:
        mask = rng.rand(N, N) < zero_frac
        np.fill_diagonal(mask, False)
        A[mask] = 0.0
    A /= A.sum(axis=1, keepdims=True)
    pi = np.full(N, 1.0 / N) if uniform_start else rng.dirichlet(np.ones(N))
    K = len(syms)
    S = max(syms) + 1
    table = np.zeros((K, N, S))
    probs = []
    for k, s in enumerate(syms):
        p = rng.dirichlet(0.5 * np.ones(s), size=N)          # (N, s)
        p = np.maximum(p, 1e-12)
        p /= p.sum(axis=1, keepdims=True)
        probs.append(p)
        table[k, :, 1:s + 1] = np.log(p)
    return dict(N=N, K=K, S=S, syms=tuple(syms), widths=[s + 1 for s in syms], pi=pi, A=A,
                table=table, probs=probs, log_start=np.asarray(myLog(pi), dtype=np.float64),
                log_trans=np.asarray(myLog(A), dtype=np.float64))


def sample_obs(model, T, seed=1, missing=0.05, dtype=np.uint8):
    """(T,K) symbols emitted by a state path sampled from the model (vectorised:
    the path is drawn run by run from the geometric dwell times)."""
    rng = np.random.RandomState(seed)
    N, K = model["N"], model["K"]
    A = model["A"]
    stay = np.diag(A)
    leave = A.copy()
    np.fill_diagonal(leave, 0.0)
    rs = leave.sum(axis=1, keepdims=True)
    leave = np.where(rs > 0, leave / np.where(rs > 0, rs, 1.0), 1.0 / N)
    cum = np.cumsum(leave, axis=1)
    states = np.empty(T, dtype=np.int64)
    t = 0
    s = int(rng.randint(N))
    while t < T:
        p = min(max(stay[s], 0.0), 1.0 - 1e-9)
        run = int(rng.geometric(1.0 - p))
        run = min(run, T - t)
        states[t:t + run] = s
        t += run
        s = int(min(np.searchsorted(cum[s], rng.rand()), N - 1))
    obs = np.zeros((T, K), dtype=dtype)
    order = np.argsort(states, kind="stable")
    bounds = np.searchsorted(states[order], np.arange(N + 1))
    for k, ns in enumerate(model["syms"]):
        c = np.cumsum(model["probs"][k], axis=1)               # (N, ns)
        u = rng.rand(T)
        sym = np.empty(T, dtype=np.int64)
        for st in range(N):
            idx = order[bounds[st]:bounds[st + 1]]
            if idx.size:
                sym[idx] = np.searchsorted(c[st], u[idx])
        sym = np.minimum(sym, ns - 1) + 1
        sym[rng.rand(T) < missing] = 0
        obs[:, k] = sym
    return obs, states


def bench_lengths(config, rng=None):
    """sequence lengths of the BASELINE.json configs"""
    if config == "c2":
        return [10_000_000]
    if config == "c3":
        rng = rng or np.random.RandomState(2)
        lens = np.exp(rng.uniform(np.log(1000), np.log(100000), size=350))
        lens = np.maximum(1000, (lens * (3.5e6 / lens.sum())).astype(np.int64))
        return [int(x) for x in lens]
    if config == "c4":   # ceil(chromLen / 250) for hg19 chr1..22, X, Y
        hg19 = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
                146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
                102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566,
                155270560, 59373566]
        return [-(-x // 250) for x in hg19]
    raise ValueError(config)


def add_missing_stretches(obs, n_stretches=8, lo=2_000, hi=20_000, seed=5):
    """Overwrite `n_stretches` runs of lo..hi consecutive rows with the missing symbol in EVERY track
    (assembly gaps, unmappable repeats): there the emission is uniform and a filter forgets its start
    only as fast as the transition matrix mixes.  Returns (obs copy, [(start, end)])."""
    rng = np.random.RandomState(seed)
    out = obs.copy()
    T = obs.shape[0]
    spans = []
    for _ in range(n_stretches):
        n = int(rng.randint(lo, hi + 1))
        n = min(n, max(1, T // (2 * n_stretches)))
        a = int(rng.randint(0, max(1, T - n)))
        out[a:a + n] = 0
        spans.append((a, a + n))
    return out, spans
