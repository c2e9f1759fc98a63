"""Replicate training (bin/teHmmTrain.py:279-306): `--reps` random restarts of Baum-Welch on the same
tracks, the best one kept.

The reference runs the replicates on a ThreadPool (runParallelShellCommands(useThreads=True)) whose
threads take turns at the GIL-bound trellis.  Here every replicate thread gets its own library
context (contexts are per thread, _lib.get_context), its own CUDA stream, and -- when the process
sees several GPUs -- its own device, dealt round robin: the E-steps of different replicates run
side by side on the device(s), the host M-steps interleave under the GIL.  Results do not depend on
the schedule: a replicate is a pure function of its seed (kernels are deterministic).
"""
from concurrent.futures import ThreadPoolExecutor

from . import _lib
from .common import LOGZERO


def train_replicates(train_one, seeds, num_threads=None, devices=None):
    """train_one(seed) -> trained model (anything with getLastLogProb()), called once per seed on
    a worker thread bound to a device and a stream.  Returns (models, best_index) with the
    reference's selection rule (teHmmTrain.py:294-300): the first replicate whose last
    log-probability is strictly greater than every earlier one, starting from LOGZERO."""
    import torch
    seeds = list(seeds)
    if devices is None:
        devices = list(range(max(1, torch.cuda.device_count())))
    if num_threads is None:
        num_threads = max(1, min(len(seeds), 2 * len(devices)))

    def work(item):
        i, seed = item
        dev = devices[i % len(devices)]
        _lib.set_thread_device(dev)
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                model = train_one(seed)
                torch.cuda.current_stream(dev).synchronize()
            return model
        finally:
            _lib.set_thread_device(None)

    if num_threads == 1 or len(seeds) == 1:
        models = [work(it) for it in enumerate(seeds)]
    else:
        with ThreadPoolExecutor(max_workers=num_threads) as pool:
            models = list(pool.map(work, enumerate(seeds)))
    best = (-1, LOGZERO)
    for i, m in enumerate(models):
        lp = m.getLastLogProb()
        if lp is not None and lp > best[1]:
            best = (i, lp)
    return models, best[0]
