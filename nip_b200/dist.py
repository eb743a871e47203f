"""Multi-GPU plumbing: one process per GPU, sequences sharded across ranks.

Sequences are independent given the parameters (the reference loops over them
serially only to save memory, src/nip.c:2182-2207), so
  * smoothing / likelihood need NO collective: every rank handles its own series;
  * EM needs exactly one all-reduce per iteration, over the concatenated
    expected-count tables plus {log-likelihood, status}.  The reference's 1.0
    pseudo-count (src/nip.c:2171-2172) is added on rank 0 only, so that it appears
    once in the reduced counts; every rank then runs the (tiny) M-step
    redundantly and parameters never leave HBM.

`torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) is the
transport; the tensor handed to all_reduce aliases the library's own count
accumulator in HBM.
"""
from __future__ import annotations

import numpy as np


def shard_series(lengths, world: int):
    """Deterministic balanced partition of series over ranks by total length
    (longest-processing-time first).  Returns a list of index arrays."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))          # ties -> lowest rank: deterministic
        parts[r].append(int(i))
        load[r] += lengths[i]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def device_tensor(ptr: int, n: int, device: int):
    """float64 torch tensor aliasing `n` doubles of device memory at `ptr`"""
    import torch

    class _Alias:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Alias(), device=torch.device("cuda", device))


class EmWorker:
    """Distributed EM over one rank's shard.

    `backend` provides
        estep(add_pseudocount) -> (counts_tensor [n+2], status_local)
            counts_tensor[:n] expected counts, [n] log-likelihood, [n+1] bad-luck flag
        mstep()                 M-step from the (reduced) counts_tensor
    which is nip_b200.api on GPUs (GpuEmBackend below) and an oracle-backed stand-in
    in the gloo tests.
    """

    def __init__(self, backend, rank: int, world: int, group=None):
        self.backend, self.rank, self.world, self.group = backend, rank, world, group

    def iteration(self):
        """one E-step + all-reduce + M-step; returns (loglik, bad_luck)"""
        counts = self.backend.estep(add_pseudocount=(self.rank == 0))
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
        tail = counts[-2:].cpu()
        ll, bad = float(tail[0]), float(tail[1]) != 0.0
        if not bad:
            self.backend.mstep()
        return ll, bad


class GpuEmBackend:
    """nip_b200.api Model/Batch behind the EmWorker interface"""

    def __init__(self, model, batch, use_evidence=None):
        self.model, self.batch, self.use_evidence = model, batch, use_evidence
        ptr, n = model.counts_device()
        self.counts = device_tensor(ptr, n, model.device)

    def estep(self, add_pseudocount):
        self.batch.estep(use_evidence=self.use_evidence, add_pseudocount=add_pseudocount, want_counts=False)
        return self.counts

    def mstep(self):
        self.model.mstep(None)
