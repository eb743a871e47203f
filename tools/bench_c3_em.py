"""Config C3 at full size: factorial DBN (4 ring-coupled chains x 16 states), EM over 16 384
series x 8 slices sharded over the ranks of one box, one NCCL all-reduce of the expected counts
per iteration.   torchrun --nproc-per-node 8 tools/bench_c3_em.py   (N_SERIES / T / ITERS env)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import nip_b200.api as api
from nip_b200.dist import EmWorker, GpuEmBackend
from nip_b200.synth import FactorialSpec

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, T, ITERS = int(os.environ.get("N_SERIES", 16384)), int(os.environ.get("T", 8)), int(os.environ.get("ITERS", 2))
per = N // world
sp = FactorialSpec(16, 4, seed=1)
fm = sp.flat()
data = sp.sample(per, T, seed=100 + rank)                 # this rank's shard
m = api.Model(fm, device=local, engine=int(os.environ.get("ENGINE", 0)))   # 0: engine 3 by itself, 1: grid team
b = m.batch(sp.obs_vars, data)
m.mstep(np.random.default_rng(1234).random(m.counts_size()) + 0.1)   # same initial parameters on every rank
w = EmWorker(GpuEmBackend(m, b), rank, world)
out = []
for it in range(ITERS):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ll, bad = w.iteration()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out.append({"iteration": it, "seconds": float(tt[0]), "loglik_per_slice": ll / (N * T), "bad_luck": bad,
                "estep_kernels_ms_rank0": m.last_kernel_ms()[0]})
if rank == 0:
    s = out[-1]["seconds"]
    print(json.dumps({"config": "C3: factorial 4x16 ring-coupled, %d series x %d slices, EM" % (N, T), "n_gpus": world,
                      "iterations": out, "em_iterations_per_s": 1.0 / s, "slice_steps_per_s": N * T / s,
                      "forward_rows_GB_per_gpu": min(64, per) * T * 65536 * 8 / 1e9,
                      "forward_rows_note": "engine 3 keeps the rows of the 64 sequences in flight only (a store per data row would be %.1f GB per GPU)" % (per * T * 65536 * 8 / 1e9)}))
if world > 1:
    dist.destroy_process_group()
