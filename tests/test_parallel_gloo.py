"""CPU: the N>1 path of fit() -- sequence sharding + ONE all-reduce of the packed
statistics per EM iteration -- with a world_size-2 gloo group.  The device
E-step is replaced by the CPU oracle inside the workers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_is_balanced_and_complete():
    from tehmm_b200.parallel import lpt_partition
    rng = np.random.RandomState(0)
    lens = rng.randint(1000, 100000, size=350)
    bins = lpt_partition(lens, 8)
    assert sorted(i for b in bins for i in b) == list(range(350))
    loads = [int(sum(lens[i] for i in b)) for b in bins]
    assert max(loads) - min(loads) <= max(lens)
    assert lpt_partition([5], 4) == [[0], [], [], []]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from tehmm_b200 import parallel
    from tehmm_b200.hmm import MultitrackHmm
    from test_gpu_api import load_fit_case
    from test_host_logic import oracle_estep

    inner = oracle_estep(orc)

    def estep(self, obs, stats, params, n_total, slots):
        # same packing as engine.estep, on a CPU tensor, through the real all-reduce
        local = {'nobs': 0, 'start': np.zeros_like(stats['start']), 'trans': np.zeros_like(stats['trans']),
                 'obs': np.zeros_like(stats['obs'])}
        lps = inner(self, obs, local, params, n_total, slots)
        packed = torch.from_numpy(np.concatenate([[lps.sum(), local['nobs']], local['start'],
                                                  local['trans'].ravel(), local['obs'].ravel(), lps]))
        packed = parallel.all_reduce_stats(packed).numpy()
        N = self.n_components
        K, _, S = stats['obs'].shape
        base = 2 + N + N * N + K * N * S
        stats['nobs'] += int(round(packed[1]))
        stats['start'] += packed[2:2 + N]
        stats['trans'] += packed[2 + N:2 + N + N * N].reshape(N, N)
        stats['obs'] += packed[2 + N + N * N:base].reshape(K, N, S)
        return packed[base:]

    MultitrackHmm._device_estep = estep
    g, hmm, em, tables = load_fit_case("fit_n4_k3")
    assert len(parallel.shard(tables)) < len(tables)
    hmm.fit(tables)
    np.savez(out_path % rank, transmat=hmm.transmat_, startprob=hmm.startprob_, table=em.getLogProbs(),
             last=hmm.getLastLogProb())
    dist.destroy_process_group()


def test_two_rank_fit_equals_reference(tmp_path):
    import torch.multiprocessing as mp
    from conftest import golden
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = golden("fit_n4_k3")
    r0, r1 = np.load(out % 0), np.load(out % 1)
    for key in ("transmat", "startprob", "table"):
        np.testing.assert_array_equal(r0[key], r1[key])          # every rank ends identical
    np.testing.assert_allclose(r0["transmat"], g["fit_transmat"], rtol=1e-10)
    np.testing.assert_allclose(r0["startprob"], g["fit_startprob"], rtol=1e-10)
    np.testing.assert_allclose(r0["table"], g["fit_table"], rtol=1e-10)
    assert float(r0["last"]) == pytest.approx(float(g["fit_last_logprob"]), rel=1e-12)
