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
