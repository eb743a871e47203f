"""CPU suite: the N > 1 host logic (series sharding, one all-reduce of the EM
sufficient statistics per iteration, pseudo-count added once) with world_size 2
over gloo.  The oracle stands in for the device E-/M-step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import Case
from nip_b200.dist import EmWorker, shard_series


def test_shard_series_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    lengths = rng.integers(1, 1000, size=257)
    for world in (1, 2, 4, 8):
        parts = shard_series(lengths, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(257))
        loads = [int(lengths[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= int(lengths.max())
    assert [p.tolist() for p in shard_series([5, 5, 5, 5], 2)] == [[0, 2], [1, 3]]


class OracleEmBackend:
    def __init__(self, om, obs_vars, series):
        self.om, self.obs_vars, self.series = om, obs_vars, series
        self.n = len(om.fm.counts_offsets()) and int(om.fm.counts_offsets()[-1])
        self.counts = torch.zeros(self.n + 2, dtype=torch.float64)

    def estep(self, add_pseudocount):
        start = np.ones(self.n) if add_pseudocount else np.zeros(self.n)
        c, ll, st = self.om.estep(self.obs_vars, self.series, counts=start)
        self.counts[:self.n] = torch.from_numpy(c)
        self.counts[self.n], self.counts[self.n + 1] = ll, float(st != 0)
        return self.counts

    def mstep(self):
        self.om.mstep(self.counts[:self.n].numpy())


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.bindings import OracleLib
    c = Case(name)
    om = OracleLib().model(c.fm)
    om.mstep(np.array([float.fromhex(x) for x in c.j["em"]["init"]]))
    mine = shard_series([len(s) for s in c.series], world)[rank]
    w = EmWorker(OracleEmBackend(om, c.obs_vars, [c.series[i] for i in mine]), rank, world)
    lls = []
    for _ in range(2):
        ll, bad = w.iteration()
        lls.append(ll)
        assert not bad
    t, p = om.parameters()
    torch.save({"ll": lls, "tables": t, "prior": p, "counts": w.backend.counts.clone()}, "%s.%d" % (out, rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["hmm5", "coupled2x3"])
def test_em_allreduce_world2(tmp_path, name, oracle_lib):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, name, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0", weights_only=False), torch.load(out + ".1", weights_only=False)
    # both ranks end with identical parameters ...
    assert np.array_equal(r0["tables"], r1["tables"]) and np.array_equal(r0["prior"], r1["prior"])
    assert r0["ll"] == r1["ll"]
    # ... equal to the single-process run (sum order differs: 1e-12, all terms >= 0)
    c = Case(name)
    om = oracle_lib.model(c.fm)
    om.mstep(np.array([float.fromhex(x) for x in c.j["em"]["init"]]))
    for k in range(2):
        counts, ll, st = om.estep(c.obs_vars, c.series)
        assert st == 0
        np.testing.assert_allclose(r0["ll"][k], ll, rtol=1e-12)
        om.mstep(counts)
    np.testing.assert_allclose(r0["counts"][:-2].numpy(), counts, rtol=1e-12)   # pseudo-count exactly once
    t, p = om.parameters()
    np.testing.assert_allclose(r0["tables"], t, rtol=1e-12)
    np.testing.assert_allclose(r0["prior"], p, rtol=1e-12)
