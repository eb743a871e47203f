#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 backend for NIP's join-tree hot path.

Workload (BASELINE.json configs[1], SURVEY §8 "C2"): HMM-style DBN, 64 hidden
states x 32 symbols, 4096 sequences x 1000 slices, forward-backward smoothing
with log-likelihood; synthetic sequences sampled from random-init CPTs.

  metric  slice-steps/s   (one slice-step = one time slice of one sequence,
                           forward AND backward, posterior + ll term)
  step    one smoothing pass over the whole resident batch
  value   device-timed, inputs resident in HBM (nipgpu_infer_device)
  e2e     same pass through the C ABI with HOST buffers: pinned H2D of the
          observations and D2H of posteriors + log-likelihoods inside the timing
  roofline  the two DMMA kernels of the pass against the FP64 tensor rate measured
          live on the device (MEASURED_PEAKS.json has HBM and bf16 only)
  cpu_baseline  the reference itself (oracle/_ref) on the host cores, bounded sample

`--impl reference` times the reference's own CPU implementation on the same
config instead (rank 0 only under torchrun).  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, M, N_SERIES, T = 64, 32, 4096, 1000          # configs[1] of BASELINE.json
F_ALG = 4.0 * S * S                              # flops per slice-step: two S x S contractions
B_ALG = 16.0 * S + 8.0 * S + 8.0                 # bytes per slice-step: alpha w+r, posterior, evidence x2
WORKLOAD = "C2: HMM-64x32, %d sequences x %d slices, forward-backward smoothing + loglik" % (N_SERIES, T)
METRIC = "slice-steps/sec (forward-backward)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------- clocks ---
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------- workload ---
def make_workload(seed_model=1, seed_data=2, n_series=N_SERIES, t=T):
    from nip_b200.synth import HmmSpec
    h = HmmSpec(S, M, seed=seed_model)
    data = h.sample(n_series, t, seed=seed_data)
    return h, data


def cpu_reference(h, data, steps=1, warmup=0, budget_s=12.0):
    """the reference's forward_backward_inference (+ll) on the host cores; returns
    (slice-steps/s, cores, kind, sample description, ms per step)"""
    from oracle import bindings
    cores = os.cpu_count() or 1
    if bindings.have_ref():
        R = bindings.RefLib()
        with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
            f.write(h.net_text())
            path = f.name
        rm = R.parse(path)
        os.unlink(path)
        # calibrate on one short series, then size the sample to ~budget_s of wall time
        ts0 = [rm.timeseries(h.obs_vars, data[0, :50])]
        per_step = rm.time_infer(ts0, h.hidden_query, True, 1) / 50.0
        per_core = max(1, int(budget_s / (per_step * T)))
        n = min(data.shape[0], per_core * cores)
        ts = [rm.timeseries(h.obs_vars, data[i]) for i in range(n)]
        times = []
        for k in range(warmup + steps):
            times.append(rm.time_infer(ts, h.hidden_query, True, cores))
        dt = float(np.mean(times[warmup:]))
        kind = "reference"
        what = ("oracle/_ref (reference sources, gcc -O2), forward_backward_inference+ll, %d of %d series x %d "
                "slices, %d forked workers" % (n, data.shape[0], T, cores))
        return n * T / dt, cores, kind, what, dt * 1e3
    O = bindings.OracleLib()
    om = O.model(h.flat())
    n = 4
    t0 = time.perf_counter()
    for i in range(n):
        om.infer(h.obs_vars, data[i], h.hidden_query)
    dt = time.perf_counter() - t0
    return n * T / dt, 1, "port", "oracle/nip_oracle.c, %d series x %d slices, 1 core" % (n, T), dt * 1e3


# ------------------------------------------------------------------- main ---
def main():
    # Everything that libraries print (NCCL's version banner, warnings) goes to stderr: stdout
    # carries exactly one JSON line.
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--em-iters", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return
        h, data = make_workload()
        val, cores, kind, what, ms = cpu_reference(h, data, steps=max(1, args.steps), warmup=args.warmup)
        print(file=json_out, flush=True, *[json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "slice-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": what},
            "cpu_baseline": {"value": val, "unit": "slice-steps/s", "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": val, "unit": "slice-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})])
        return

    import torch
    import torch.distributed as dist
    import nip_b200.api as api

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.load_library()                      # raises if the CUDA library is missing: no fallback

    h, data = make_workload(seed_data=2 + rank)   # weak scaling: every rank smooths its own 4096 x 1000
    model = api.Model(h.flat(), device=local_rank)
    assert model.engine == api.ENGINE_CHAIN
    batch = model.batch(h.obs_vars, data)
    query = h.hidden_query
    stream = torch.cuda.ExternalStream(model.L.nipgpu_model_stream(model.h), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                     # nvidia-smi needs ~1 s to start streaming: start before warm-up
    for _ in range(W):
        batch.infer_device(query)
    barrier()
    sampler.mark()                      # samples from here on belong to the timed regions
    api.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.infer_device(query)
        kernel_ms.append(model.last_kernel_ms()[0])
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = api.launch_count()

    # ---- end to end through the host-buffer entry point (pinned host memory) ----
    obs_host = torch.from_numpy(np.ascontiguousarray(data.reshape(-1, 1))).pin_memory()
    post_host = torch.empty((batch.rows, S), dtype=torch.float64).pin_memory()
    ll_host = torch.empty(N_SERIES, dtype=torch.float64).pin_memory()
    post_np, ll_np, obs_np = post_host.numpy(), ll_host.numpy(), obs_host.numpy()

    def e2e_step():
        batch.update(obs_np)                                  # H2D, every step
        batch.infer(query, out=post_np, ll_out=ll_np)         # kernels + D2H of posteriors and ll
    e2e_step()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t1) / args.e2e_steps
    clocks = sampler.stop()             # covers the device-timed steps and the end-to-end steps

    # ---- EM: E-step + one all-reduce of the sufficient statistics + M-step per iteration ----
    from nip_b200.dist import EmWorker, GpuEmBackend
    rng = np.random.default_rng(7)
    model.mstep(rng.random(model.counts_size()) + 0.1)     # same random start on every rank
    worker = EmWorker(GpuEmBackend(model, batch), rank, world)
    worker.iteration()
    barrier()
    t2 = time.perf_counter()
    em_ll = 0.0
    for _ in range(args.em_iters):
        em_ll, em_bad = worker.iteration()
    barrier()
    em_s = (time.perf_counter() - t2) / args.em_iters

    times = torch.tensor([dev_ms, e2e_s * 1e3, wall * 1e3, em_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms, em_ms = [float(x) for x in times.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    units = N_SERIES * T * world
    value = units * args.steps / (dev_ms * 1e-3)
    dmma_tf, dfma_tf, copy_gbs = api.probe_peaks(local_rank)
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    k_ms = float(np.mean(kernel_ms))                      # fwd + bwd kernels of one pass (CUDA events)
    achieved_tf = F_ALG * N_SERIES * T / (k_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["chain_pass_dram_bytes"]
    except (OSError, ValueError, KeyError):
        pass
    hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
    out = {
        "metric": METRIC, "value": value, "unit": "slice-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "engine": "chain (DMMA m8n8k4, warp-resident recursion)",
                   "l2_policy": "working set 4.2 GB per pass >> 126 MB L2 (no flush needed)",
                   "per_gpu_units_per_step": N_SERIES * T, "wall_ms_per_step": wall_ms / args.steps},
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": dmma_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / dmma_tf if dmma_tf else None, "traffic": traffic,
                     "kernel": "k_chain_forward<8> + k_chain_backward<8> (one launch each per pass)",
                     "kernel_ms_per_pass": k_ms, "flops_per_slice_step": F_ALG,
                     "peak_source": "FP64 DMMA rate measured live by nipgpu_probe_peaks (MEASURED_PEAKS.json "
                                    "has no FP64 entry); DFMA %.1f TF, copy %.0f GB/s in the same probe" % (dfma_tf, copy_gbs),
                     "hbm": {"achieved": B_ALG * N_SERIES * T / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "bytes_per_slice_step": B_ALG,
                             "frac": B_ALG * N_SERIES * T / (k_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks_file else "fallback"}},
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "slice-steps/s",
                "h2d_bytes_per_step": int(obs_np.nbytes), "d2h_bytes_per_step": int(post_np.nbytes + ll_np.nbytes),
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches), "clocks": clocks,
        "em": {"metric": "EM iterations/s (E-step over the whole set + all-reduce + M-step)",
               "value": 1e3 / em_ms, "unit": "iter/s", "ms_per_iteration": em_ms,
               "slice_steps_per_s": units / (em_ms * 1e-3),
               "workload": "same model and data, %d x %d slices per GPU, %d GPU(s)" % (N_SERIES, T, world),
               "allreduce_doubles": int(model.counts_size() + 2), "loglik_per_slice": em_ll / units},
    }
    if not args.no_cpu_baseline:
        v, cores, kind, what, _ = cpu_reference(h, data)
        out["cpu_baseline"] = {"value": v, "unit": "slice-steps/s", "cores": cores, "kind": kind, "sample": what}
    print(json.dumps(out), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
