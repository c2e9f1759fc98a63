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


def _worker(rank, world, port, out_path, single=False):
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

    def local_estep(self, obs, params, n_total, slots, stats_S):
        # same packing as Engine.estep, on a CPU tensor; the real MultitrackHmm._device_estep around it
        # does the all-reduce (gloo here, NCCL on the GPU box) and the empty-shard handling
        N, K = self.n_components, self.emissionModel.getNumTracks()
        local = {'nobs': 0, 'start': np.zeros(N), 'trans': np.zeros((N, N)), 'obs': np.zeros((K, N, stats_S))}
        lps = inner(self, obs, local, params, n_total, slots)
        return torch.from_numpy(np.concatenate([[lps.sum(), local['nobs']], local['start'],
                                                local['trans'].ravel(), local['obs'].ravel(), lps]))

    MultitrackHmm._local_estep = local_estep
    g, hmm, em, tables = load_fit_case("fit_n4_k3")
    if single:
        tables = tables[:1]          # fewer sequences than ranks: one rank's shard is EMPTY
        assert sorted(len(b) for b in parallel.lpt_partition([len(t) for t in tables], world)) == [0, 1]
    else:
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


def test_two_rank_fit_with_an_empty_shard(oracle, tmp_path):
    """ONE sequence on two ranks (single-chromosome training on a multi-GPU box): the rank without
    work must still reach the all-reduce of every iteration (it used to raise in upload_batch while
    the other rank waited forever) and end with the same parameters as a single-process fit."""
    import torch.multiprocessing as mp
    from tehmm_b200.hmm import MultitrackHmm
    from test_gpu_api import load_fit_case
    from test_host_logic import oracle_estep
    port = 31500 + (os.getpid() % 2000)
    out = str(tmp_path / "single%d.npz")
    mp.spawn(_worker, args=(2, port, out, True), nprocs=2, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    g, hmm, em, tables = load_fit_case("fit_n4_k3")
    hmm._device_estep = lambda obs, stats, params, n_total, slots: oracle_estep(oracle)(
        hmm, obs, stats, params, n_total, slots)
    hmm.fit(tables[:1])
    for r in (r0, r1):
        np.testing.assert_allclose(r["transmat"], hmm.transmat_, rtol=1e-12)
        np.testing.assert_allclose(r["startprob"], hmm.startprob_, rtol=1e-12)
        np.testing.assert_allclose(r["table"], em.getLogProbs(), rtol=1e-12)
        assert float(r["last"]) == pytest.approx(hmm.getLastLogProb(), rel=1e-12)
    for key in ("transmat", "startprob", "table"):
        np.testing.assert_array_equal(r0[key], r1[key])


# ---------------------------------------------------------------------------
# one long sequence, time axis split over the ranks (parallel.run_time_sharded)
def _numpy_viterbi_window(frame, log_start, log_trans, core, window):
    """plain NumPy Viterbi of frame[window] (stand-in for Engine.viterbi_window on CPU):
    max-normalised delta rows, lowest-index ties, path score of the core rows."""
    (a, b), (w0, w1) = core, window
    f = frame[w0:w1]
    n = f.shape[0]
    delta = np.empty_like(f)
    bp = np.zeros(f.shape, dtype=np.int64)
    d = log_start + f[0]
    delta[0] = d - d.max()
    for t in range(1, n):
        cand = delta[t - 1][:, None] + log_trans
        bp[t] = np.argmax(cand, axis=0)
        d = cand.max(axis=0) + f[t]
        delta[t] = d - d.max()
    st = np.empty(n, dtype=np.int64)
    st[-1] = int(np.argmax(delta[-1]))
    for t in range(n - 1, 0, -1):
        st[t - 1] = bp[t][st[t]]
    score = 0.0
    for t in range(a, b):
        j = st[t - w0]
        score += (log_start[j] if t == 0 else log_trans[st[t - 1 - w0], j]) + frame[t, j]
    return {"mode": "diff", "tol": 1e-9, "right_state": int(st[b - 1 - w0]), "right_probe": delta[b - 1 - w0],
            "left_state": int(st[a - 1 - w0]) if a > w0 else None,
            "left_probe": delta[a - 1 - w0] if a > w0 else None,
            "result": (score, st[a - w0:b - w0])}


def _sharded_case():
    sys.path.insert(0, ROOT)
    from tehmm_b200 import synth
    m = synth.make_model(N=6, syms=(4, 3, 5), seed=11)
    obs, _ = synth.sample_obs(m, 5000, seed=12)
    return m, obs


def _sharded_worker(rank, world, port, out_path, halo):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from tehmm_b200 import parallel
    m, obs = _sharded_case()
    T, N = obs.shape[0], m["N"]
    frame = np.zeros((T, N))
    orc.fastAllLogProbs(obs, m["table"], frame, 1.0, None)
    (part, core), H, attempts = parallel.run_time_sharded(
        T, lambda c, w: _numpy_viterbi_window(frame, m["log_start"], m["log_trans"], c, w), halo)
    total = parallel.sum_over_ranks(part)
    path = parallel.gather_states(core, T)
    np.savez(out_path % rank, path=path, total=total, H=H, attempts=attempts)
    dist.destroy_process_group()


@pytest.mark.parametrize("halo", [256, 1])
def test_two_rank_time_sharded_viterbi_equals_oracle(tmp_path, halo, oracle):
    """world_size 2 over gloo: the stitched path and the summed score equal the oracle's
    Viterbi of the whole sequence; a hopeless halo (1 step) is detected at the boundary and
    repaired by retrying with longer ones."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000) + halo % 7
    out = str(tmp_path / "shard%d.npz")
    mp.spawn(_sharded_worker, args=(2, port, out, halo), nprocs=2, join=True)
    m, obs = _sharded_case()
    T, N = obs.shape[0], m["N"]
    frame = np.zeros((T, N))
    oracle.fastAllLogProbs(obs, m["table"], frame, 1.0, None)
    st, lp = oracle._viterbi(T, N, m["log_start"], m["log_trans"], None, frame)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    np.testing.assert_array_equal(r0["path"], r1["path"])
    np.testing.assert_array_equal(r0["path"], st)
    assert float(r0["total"]) == pytest.approx(lp, rel=1e-12)
    assert int(r0["H"]) == int(r1["H"])
    if halo == 1:
        assert int(r0["attempts"]) > 1
    else:
        assert int(r0["attempts"]) == 1


def test_boundary_verdict():
    from tehmm_b200.parallel import boundaries_agree, boundary_vector, probes_agree, time_shards
    assert time_shards(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert probes_agree([0.0, -1.0, -np.inf], [-5.0, -6.0, -np.inf], "diff", 1e-9)
    assert not probes_agree([0.0, -1.0, -np.inf], [-5.0, -6.1, -np.inf], "diff", 1e-9)
    assert not probes_agree([0.0, -1.0, -np.inf], [-5.0, -6.0, -7.0], "diff", 1e-9)
    assert probes_agree([1.0, 0.5, 0.0], [2.0, 1.0, 0.0], "ratio", 1e-9)
    assert not probes_agree([1.0, 0.5, 0.0], [2.0, 1.1, 0.0], "ratio", 1e-9)
    left = {"left_state": None, "right_state": 3, "left_probe": None, "right_probe": np.array([0.0, -2.0])}
    good = {"left_state": 3, "right_state": 1, "left_probe": np.array([-1.0, -3.0]), "right_probe": np.array([0.0, -9.0])}
    bad = dict(good, left_state=2)
    sh = time_shards(100, 2)
    assert boundaries_agree([boundary_vector(left), boundary_vector(good)], sh, "diff", 1e-9)
    assert not boundaries_agree([boundary_vector(left), boundary_vector(bad)], sh, "diff", 1e-9)
