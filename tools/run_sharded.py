#!/usr/bin/env python
"""ONE long sequence decoded with its time axis split over the ranks:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded.py [T] [N_states]
Every rank builds the same synthetic sequence, decode_sharded / score_sharded run over NCCL, and
rank 0 checks the stitched result against its own single-GPU decode of the whole sequence.
Prints one JSON line (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tehmm_b200 import synth
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from run_configs import make_hmm, sample_long

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = synth.make_model(N=N, seed=0)
obs = sample_long(synth, m, T, seed=500)
hv, _ = make_hmm(m)
hm, _ = make_hmm(m, algorithm="map")
out = {}
for name, h in (("viterbi", hv), ("map", hm)):
    h.decode_sharded(obs[:200_000])                     # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lp, path = h.decode_sharded(obs)
    torch.cuda.synchronize(); out[name + "_seconds"] = time.perf_counter() - t0
    out[name] = (lp, path)
t0 = time.perf_counter(); lp_s = hv.score_sharded(obs); out["score_seconds"] = time.perf_counter() - t0
if rank == 0:
    t0 = time.perf_counter(); ref_v = hv.decode(obs); one_v = time.perf_counter() - t0
    ref_m = hm.decode(obs)
    ref_lp = hv.score(obs)
    line = {"what": "time-sharded single sequence", "steps": T, "states": N, "ranks": world,
            "viterbi_seconds": out["viterbi_seconds"], "map_seconds": out["map_seconds"], "score_seconds": out["score_seconds"],
            "single_gpu_viterbi_seconds": one_v,
            "viterbi_path_equal": bool(np.array_equal(out["viterbi"][1], ref_v[1])),
            "viterbi_logprob_rel_err": float(abs(out["viterbi"][0] - ref_v[0]) / abs(ref_v[0])),
            "map_path_equal": bool(np.array_equal(out["map"][1], ref_m[1])),
            "logprob_rel_err": float(abs(lp_s - ref_lp) / abs(ref_lp))}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
