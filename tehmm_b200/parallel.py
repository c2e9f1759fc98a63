"""Multi-GPU plumbing: one process per GPU, sequences sharded across ranks,
ONE all-reduce of the packed sufficient statistics per EM iteration
(SURVEY.md section 8e).  Decoding needs no collective: each rank decodes its
own sequences.

torch.distributed is optional: without an initialised process group everything
here is the identity.  NCCL is used for CUDA tensors; with a gloo group (CPU
tests, world_size 2) the tensor takes a round trip through host memory.
"""
import heapq


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    """(world_size, rank)"""
    d = _dist()
    if d is None:
        return 1, 0
    return d.get_world_size(), d.get_rank()


def lpt_partition(lengths, nbins):
    """Longest-processing-time greedy bin packing.  Returns nbins lists of indices
    (each list in increasing index order).  Deterministic: ties go to the lower bin."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    heap = [(0, b) for b in range(nbins)]
    heapq.heapify(heap)
    bins = [[] for _ in range(nbins)]
    for i in order:
        load, b = heapq.heappop(heap)
        bins[b].append(i)
        heapq.heappush(heap, (load + int(lengths[i]), b))
    return [sorted(b) for b in bins]


def shard_indices(lengths):
    """indices of the sequences this rank owns"""
    n, r = world()
    if n == 1:
        return list(range(len(lengths)))
    return lpt_partition(lengths, n)[r]


def shard(obs_list):
    """the sub-list of sequences this rank owns (all of them without a process group)"""
    n, _ = world()
    if n == 1:
        return list(obs_list)
    idx = shard_indices([len(o) for o in obs_list])
    return [obs_list[i] for i in idx]


def all_reduce_stats(packed):
    """SUM all-reduce of the packed statistics tensor, in place; returns it."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return packed
    backend = d.get_backend()
    if packed.is_cuda and backend != "nccl":
        host = packed.cpu()
        d.all_reduce(host)
        packed.copy_(host)
    else:
        d.all_reduce(packed)
    return packed


# ---------------------------------------------------------------------------
# ONE long sequence over the ranks: rank r owns
# the core range [a_r, b_r) of the time axis and works on the window
# [a_r - H, b_r + H): the same speculate / verify / repair idea as the chunks
# inside one GPU, one level up.
#   * forward-like recursions (Viterbi DP, forward filter) start H steps to the
#     left of the core from a flat vector; the vector a rank reaches at a_r - 1
#     ("left probe") must agree with the vector its left neighbour computed for
#     the same time as the last step of ITS core ("right probe"), which is exact
#     by induction from rank 0;
#   * backward-like walks (Viterbi traceback, backward filter) enter H steps to
#     the right; the state a rank holds at b_r - 1 must equal the state its right
#     neighbour reaches at b_r - 1, exact by induction from the last rank.
# The only exchange is ONE all-gather of two short vectors per rank.  If any
# boundary disagrees, every rank redoes its window with a four times longer halo
# (H >= T degenerates to every rank doing the whole sequence, which is exact).
def time_shards(T, n):
    """core ranges [(a_0, b_0), ...]: n contiguous, near-equal pieces of [0, T)"""
    return [(T * i // n, T * (i + 1) // n) for i in range(n)]


def _gather_vectors(vec, n, r):
    """all-gather of one float64 vector per rank -> list of n vectors"""
    import numpy as np
    d = _dist()
    if d is None or n == 1:
        return [vec]
    import torch
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    if d.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(n)]
    d.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def probes_agree(right, left, mode, tol):
    """right: what rank r computed for the last step of its core; left: what rank r+1
    reached for the same time step after its warm-up.
    mode 'diff'  : log-domain vectors equal up to an additive constant (Viterbi delta rows)
    mode 'ratio' : non-negative vectors equal up to a factor (forward filter)
    mode 'equal' : identical (state bands)"""
    import numpy as np
    right, left = np.asarray(right, dtype=np.float64), np.asarray(left, dtype=np.float64)
    if mode == "equal":
        return bool(np.array_equal(right, left))
    if mode == "ratio":
        zr, zl = right <= 0, left <= 0
        if not np.array_equal(zr, zl) or zr.all():
            return bool(np.array_equal(zr, zl) and zr.all())
        right, left = np.log(right[~zr]), np.log(left[~zl])
    else:
        fr, fl = np.isfinite(right), np.isfinite(left)
        if not np.array_equal(fr, fl) or not fr.any():
            return bool(np.array_equal(fr, fl))
        right, left = right[fr], left[fl]
    d = right - left
    return bool(d.max() - d.min() <= tol)


def boundary_vector(res):
    """what a rank contributes to the all-gather: [P, left_state, right_state, right_probe[P], left_probe[P]]"""
    import numpy as np
    lp, rp = res.get("left_probe"), res.get("right_probe")
    P = 0 if rp is None else len(rp)
    vec = np.full(3 + 2 * P, np.nan)
    vec[0] = P
    vec[1] = -1 if res.get("left_state") is None else res["left_state"]
    vec[2] = -1 if res.get("right_state") is None else res["right_state"]
    if P:
        vec[3:3 + P] = rp
        if lp is not None:
            vec[3 + P:] = lp
    return vec


def boundaries_agree(allv, shards, mode, tol):
    """the verdict every rank computes from the gathered vectors (identical on all ranks)"""
    ok = True
    for i in range(len(shards) - 1):
        mine, nxt = allv[i], allv[i + 1]
        if shards[i + 1][0] >= shards[i + 1][1] or shards[i][0] >= shards[i][1]:
            continue                                  # empty core (T < n)
        if mine[2] >= 0 and nxt[1] >= 0 and mine[2] != nxt[1]:
            ok = False
        Pi = int(mine[0])
        if Pi and not probes_agree(mine[3:3 + Pi], nxt[3 + Pi:3 + 2 * Pi], mode, tol):
            ok = False
    return ok


def run_time_sharded(T, work, halo=4096, ranks=None, gather=None):
    """Drive `work(core, window) -> dict(result=..., left_state, right_state, left_probe,
    right_probe, mode, tol)` under the scheme above.

    left_state  : state this rank's walk holds at time a_r - 1 (None for rank 0 / not a walk)
    right_state : state it holds at b_r - 1, the last step of its core
    left_probe  : vector its warm-up reached at a_r - 1 (None for rank 0)
    right_probe : vector it computed at b_r - 1
    Returns (result of the accepted attempt, halo used, attempts)."""
    n, r = ranks if ranks is not None else world()
    gather = gather or _gather_vectors
    shards = time_shards(T, n)
    a, b = shards[r]
    H = max(1, int(halo))
    attempts = 0
    while True:
        attempts += 1
        window = (max(0, a - H), min(T, b + H))
        res = work((a, b), window)
        if n == 1:
            return res["result"], H, attempts
        allv = gather(boundary_vector(res), n, r)
        ok = boundaries_agree(allv, shards, res["mode"], res["tol"])
        if ok or H >= T:
            return res["result"], H, attempts
        H = min(T, 4 * H)


def sum_over_ranks(x):
    """float64 scalar summed over the ranks"""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64)
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.all_reduce(t)
    return float(t.item())


def gather_states(core_states, T):
    """every rank's core states -> the whole path (int64[T]) on every rank; one byte per step on the wire"""
    import numpy as np
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return np.asarray(core_states, dtype=np.int64)
    import torch
    n = d.get_world_size()
    shards = time_shards(T, n)
    width = max(b - a for a, b in shards)
    buf = np.zeros(width, dtype=np.uint8)
    buf[:len(core_states)] = core_states
    t = torch.from_numpy(buf)
    if d.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(n)]
    d.all_gather(out, t)
    parts = [o.cpu().numpy()[:b - a] for o, (a, b) in zip(out, shards)]
    return np.concatenate(parts).astype(np.int64)
