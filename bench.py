#!/usr/bin/env python
"""bench.py — benchmark of the B200 backend for NIP's join-tree hot path.

Headline workload (BASELINE.json configs[1], SURVEY §8 "C2"): HMM-style DBN, 64 hidden
states x 32 symbols, 4096 sequences x 1000 slices, forward-backward smoothing with
log-likelihood; synthetic sequences sampled from random-init CPTs.  `--config C1|C3|C4|C5`
makes another configuration of BASELINE.json the headline; without it the other four are
measured device-timed in a `configs` block of the same JSON line (1 GPU only).

  metric   slice-steps/s (one slice-step = one time slice of one sequence, forward AND
           backward, posterior + ll term); C5: records/s of the likelihood loop
  step     one pass of the hot path over the whole resident batch
  value    device-timed, inputs resident in HBM
  e2e      the same pass through the C ABI with HOST buffers: pinned H2D of the observations and
           D2H of the results inside the timing (each rank bound to its GPU's NUMA node)
  parity   before any timing counts: the series the CPU arm ran are compared with the GPU
           results at 1e-9 relative; the run aborts on a mismatch
  roofline dominant kernel(s) against the FP64 tensor rate measured live (MEASURED_PEAKS.json has
           HBM and bf16 only) or against the measured HBM bandwidth, whichever bounds the config
  cpu_baseline  the reference itself (oracle/_ref) on the host cores, bounded sample; also at one
           core and with the shipped compiler flags (no -O)
  em / em_strong  EM iterations/s: per-GPU set fixed (weak) and ONE fixed set sharded over the
           GPUs (strong); one all-reduce of the expected counts per iteration
`--impl reference` times the reference's own CPU implementation on the same config instead
(rank 0 only under torchrun).  One JSON line on stdout.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
RTOL = 1e-9


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


def bind_to_gpu_numa(index):
    """pin this process to the cores of the NUMA node the GPU hangs off, so that the pinned host
    buffers of the end-to-end leg are allocated next to it (first touch).  Returns a description."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]                      # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"numa_node": None, "note": "single NUMA node (sysfs reports -1)"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:                     # noqa: BLE001 - placement is best effort
        return {"numa_node": None, "note": "not bound: %s" % e}


def max_rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    den = np.where(want == 0, 1.0, np.abs(want))
    err = np.abs(got - want) / den
    err = np.where((want == 0) & (got != 0), np.inf, err)    # an exact zero must stay an exact zero
    return float(err.max()) if err.size else 0.0


# ---------------------------------------------------------------- configs ---
class HmmConfig:
    """C2 / C4: HMM-style DBN, smoothing + log-likelihood (chain engine: DMMA contraction)"""

    def __init__(self, name, S, M, n_series, T, cpu_T=None, parity_series=4):
        self.name, self.S, self.M, self.n_series, self.T = name, S, M, n_series, T
        self.cpu_T = cpu_T or T            # slices per series in the CPU sample (C4: 0.16 s per slice-step)
        self.parity_series = parity_series
        self.metric, self.unit = "slice-steps/sec (forward-backward)", "slice-steps/s"
        self.workload = "%s: HMM-%dx%d, %d sequences x %d slices, forward-backward smoothing + loglik" % (name, S, M, n_series, T)
        self.f_alg = 4.0 * S * S           # two S x S contractions per slice-step
        self.b_alg = 16.0 * S + 8.0 * S + 8.0   # alpha write + read, posterior, evidence twice
        self.bound = "tensor"
        self.kernel = ("k_chain_forward_team<8,2> + k_chain_backward_team<8,2> (one launch each per pass, + k_chain_final)" if S <= 64 else
                       "k_dense_gemm (one per slice and direction) + settle kernels")
        self.engine = ("chain (DMMA m8n8k4, recursion resident in the registers of two-warp teams)" if S <= 64 else
                       "dense (per-slice DGEMM on DMMA, 128x128x16 tiles)")

    def build(self, seed_data):
        from nip_b200.synth import HmmSpec
        self.spec = HmmSpec(self.S, self.M, seed=1)
        self.data = self.spec.sample(self.n_series, self.T, seed=seed_data)
        self.obs_vars, self.query = self.spec.obs_vars, self.spec.hidden_query
        self.row = self.S
        return self.spec.flat()

    def units(self):
        return self.n_series * self.T

    def device_step(self, model, batch):
        batch.infer_device(self.query)
        return model.last_kernel_ms()[0]

    def e2e_buffers(self, torch):
        self.obs_host = torch.from_numpy(np.ascontiguousarray(self.data.reshape(-1, 1))).pin_memory()
        self.post_host = torch.empty((self.units(), self.row), dtype=torch.float64).pin_memory()
        self.ll_host = torch.empty(self.n_series, dtype=torch.float64).pin_memory()
        return int(self.obs_host.numpy().nbytes), int(self.post_host.numpy().nbytes + self.ll_host.numpy().nbytes)

    def e2e_step(self, batch):
        batch.update(self.obs_host.numpy())                                   # H2D, every step
        batch.infer(self.query, out=self.post_host.numpy(), ll_out=self.ll_host.numpy())   # kernels + D2H

    # -- the checker: the reference (oracle/_ref) when it is built, else the C restatement --
    def _cpu_model(self):
        from oracle import bindings
        if bindings.have_ref():
            R = bindings.RefLib()
            with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
                f.write(self.spec.net_text())
                path = f.name
            rm = R.parse(path)
            os.unlink(path)
            return "reference", rm
        return "port", bindings.OracleLib().model(self.spec.flat())

    def parity(self, api, model, batch):
        """the first series of the workload through the CPU checker (in memory, not the six-digit
        files) against the GPU: full-batch results when the CPU runs whole series, else the same
        truncated series as a small GPU batch of their own"""
        kind, cm = self._cpu_model()
        n, Tc = self.parity_series, self.cpu_T
        series = [self.data[i, :Tc] for i in range(n)]
        if Tc == self.T:
            post, ll = self.post_host.numpy(), self.ll_host.numpy()
            got = [(post[i * self.T:(i + 1) * self.T], ll[i]) for i in range(n)]
        else:
            sb = model.batch(self.obs_vars, series)
            p, l = sb.infer(self.query)
            got = [(g, l[i]) for i, g in enumerate(sb.split(p))]
            sb.close()
        worst = 0.0
        for i, s in enumerate(series):
            if kind == "reference":
                want, llw = cm.infer(cm.timeseries(self.obs_vars, s), self.query)
            else:
                want, llw = cm.infer(self.obs_vars, s, self.query)
            worst = max(worst, max_rel_err(got[i][0], want), max_rel_err(got[i][1], llw))
        return {"checked_series": n, "slices_each": Tc, "against": kind, "max_rel_err": worst, "tolerance": RTOL,
                "what": "posterior marginals of P1 and per-series log-likelihood, in memory",
                "ok": bool(worst <= RTOL)}

    def cpu_baseline(self, budget_s=12.0, steps=1, warmup=0, variants=True):
        from oracle import bindings
        cores = os.cpu_count() or 1
        kind, cm = self._cpu_model()
        Tc = self.cpu_T
        if kind == "port":
            n = 2
            t0 = time.perf_counter()
            for i in range(n):
                cm.infer(self.obs_vars, self.data[i, :Tc], self.query)
            dt = time.perf_counter() - t0
            what = "oracle/nip_oracle.c (C restatement, gcc -O2), %d series x %d slices, 1 core" % (n, Tc)
            return {"value": n * Tc / dt, "unit": self.unit, "cores": 1, "kind": "port", "sample": what}, dt * 1e3
        ts0 = [cm.timeseries(self.obs_vars, self.data[0, :min(50, Tc)])]
        per_step = cm.time_infer(ts0, self.query, True, 1) / min(50, Tc)
        per_core = max(1, int(budget_s / (per_step * Tc)))
        n = min(self.n_series, per_core * cores)
        ts = [cm.timeseries(self.obs_vars, self.data[i, :Tc]) for i in range(n)]
        times = [cm.time_infer(ts, self.query, True, cores) for _ in range(warmup + steps)]
        dt = float(np.mean(times[warmup:]))
        what = ("oracle/_ref (reference sources, gcc -O2), forward_backward_inference+ll, %d of %d series x %d "
                "slices, %d forked workers" % (n, self.n_series, Tc, cores))
        out = {"value": n * Tc / dt, "unit": self.unit, "cores": cores, "kind": "reference", "sample": what}
        if variants:   # the as-shipped single-threaded figure and the shipped flags (no -O)
            n1 = max(1, min(n, per_core // 4 or 1))
            one = cm.time_infer(ts[:n1], self.query, True, 1)
            out["one_core"] = {"value": n1 * Tc / one, "unit": self.unit, "cores": 1, "flags": "-O2",
                               "sample": "%d series x %d slices" % (n1, Tc)}
            o0 = os.path.join(ROOT, "oracle", "_ref", "libnip_ref_O0.so")
            if os.path.exists(o0):
                R0 = bindings.RefLib(o0)
                with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
                    f.write(self.spec.net_text())
                    path = f.name
                m0 = R0.parse(path)
                os.unlink(path)
                n0 = max(1, n1 // 2)
                ts_0 = [m0.timeseries(self.obs_vars, self.data[i, :Tc]) for i in range(n0)]
                t0 = m0.time_infer(ts_0, self.query, True, 1)
                out["one_core_shipped_flags"] = {"value": n0 * Tc / t0, "unit": self.unit, "cores": 1,
                                                 "flags": "-g, no -O (the reference's Makefile:15)",
                                                 "sample": "%d series x %d slices" % (n0, Tc)}
        return out, dt * 1e3

    def roofline(self, k_ms, peaks):
        tf = self.f_alg * self.units() / (k_ms * 1e-3) / 1e12
        gbs = self.b_alg * self.units() / (k_ms * 1e-3) / 1e9
        return {"bound": "tensor", "achieved": tf, "peak": peaks["dmma_tf"], "unit": "TFLOP/s",
                "frac": tf / peaks["dmma_tf"] if peaks["dmma_tf"] else None, "kernel": self.kernel,
                "kernel_ms_per_pass": k_ms, "flops_per_slice_step": self.f_alg,
                "peak_source": "FP64 DMMA rate measured live by nipgpu_probe_peaks (MEASURED_PEAKS.json has no FP64 "
                               "entry); DFMA %.1f TF, copy %.0f GB/s in the same probe" % (peaks["dfma_tf"], peaks["copy_gbs"]),
                "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "bytes_per_slice_step": self.b_alg,
                        "frac": gbs / peaks["hbm_gbs"], "peak_source": peaks["hbm_source"]}}


class SmallModelConfig:
    """C1 / C5: examples/model.net (4 states, 5 symbols; the parsed tables are the golden fixture
    generated from the reference).  C1: forward-backward smoothing of 365 x 24 slices as demo.sh
    does; C5: the niplikelihood loop over 1 M series x 50."""

    def __init__(self, name, n_series, T, likelihood):
        self.name, self.n_series, self.T, self.likelihood = name, n_series, T, likelihood
        self.bound, self.row = "hbm", 4
        if likelihood:
            self.metric, self.unit = "records/sec (niplikelihood loop)", "records/s"
            self.workload = "%s: examples/model.net, niplikelihood over %d series x %d slices (m1, m2 per record)" % (name, n_series, T)
            self.b_alg = 4.0 + 1.0 + 16.0   # observation, first-row flag, (m1, m2)
            self.kernel = "k_jt_like_gather (+ k_jt_likelihood once per evidence configuration)"
            self.engine = "generic join tree, memoised per evidence configuration"
        else:
            self.metric, self.unit = "slice-steps/sec (forward-backward)", "slice-steps/s"
            self.workload = "%s: examples/model.net, %d series x %d slices, forward-backward smoothing + loglik" % (name, n_series, T)
            self.b_alg = 16.0 * 4 + 8.0 * 4 + 8.0   # 104 B per slice-step
            self.kernel = "k_chain_forward<1> + k_chain_backward<1> (365 series); k_chain_small_forward/backward<4> at 1 M series"
            self.engine = "chain (one 8-state DMMA tile below 2048 series, one thread per sequence above)"

    def build(self, seed_data):
        from cases import Case
        c = Case("model_net")
        self.fm, self.obs_vars, self.query = c.fm, c.obs_vars, [1]
        rng = np.random.default_rng(seed_data)
        # M1 drawn from the model's own stationary-ish emission mix: any valid symbol stream does
        self.data = rng.integers(0, int(c.fm.var_card[c.obs_vars[0]]), size=(self.n_series, self.T, 1), dtype=np.int32)
        self.on = np.zeros(c.fm.n_vars, dtype=np.uint8)
        self.on[c.obs_vars[0]] = 1
        return c.fm

    def units(self):
        return self.n_series * self.T

    def device_step(self, model, batch):
        if self.likelihood:
            batch.likelihood(1 - self.on, self.on)     # the ABI call is host-buffered; kernel time from CUDA events
        else:
            batch.infer_device(self.query)
        return model.last_kernel_ms()[0]

    def e2e_buffers(self, torch):
        self.obs_host = torch.from_numpy(np.ascontiguousarray(self.data.reshape(-1, 1))).pin_memory()
        if self.likelihood:
            return int(self.obs_host.numpy().nbytes), int(self.units() * 16)
        self.post_host = torch.empty((self.units(), self.row), dtype=torch.float64).pin_memory()
        self.ll_host = torch.empty(self.n_series, dtype=torch.float64).pin_memory()
        return int(self.obs_host.numpy().nbytes), int(self.post_host.numpy().nbytes + self.ll_host.numpy().nbytes)

    def e2e_step(self, batch):
        batch.update(self.obs_host.numpy())
        if self.likelihood:
            self.like = batch.likelihood(1 - self.on, self.on)
        else:
            batch.infer(self.query, out=self.post_host.numpy(), ll_out=self.ll_host.numpy())

    def parity(self, api, model, batch):
        from oracle import bindings
        om = bindings.OracleLib().model(self.fm)
        n = 16
        worst = 0.0
        for i in range(n):
            s = self.data[i]
            if self.likelihood:
                want = om.likelihood(self.obs_vars, s, 1 - self.on, self.on)
                worst = max(worst, max_rel_err(self.like[i * self.T:(i + 1) * self.T], want))
            else:
                want, llw = om.infer(self.obs_vars, s, self.query)
                worst = max(worst, max_rel_err(self.post_host.numpy()[i * self.T:(i + 1) * self.T], want),
                            max_rel_err(self.ll_host.numpy()[i], llw))
        return {"checked_series": n, "slices_each": self.T, "against": "port (oracle/nip_oracle.c, bit-identical to the reference in tests/test_oracle.py)",
                "max_rel_err": worst, "tolerance": RTOL, "ok": bool(worst <= RTOL)}

    def cpu_baseline(self, budget_s=6.0, steps=1, warmup=0, variants=False):
        from oracle import bindings
        om = bindings.OracleLib().model(self.fm)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s and n < self.n_series:
            if self.likelihood:
                om.likelihood(self.obs_vars, self.data[n], 1 - self.on, self.on)
            else:
                om.infer(self.obs_vars, self.data[n], self.query)
            n += 1
        dt = time.perf_counter() - t0
        what = "oracle/nip_oracle.c (C restatement of the reference, gcc -O2), %d series x %d slices, 1 core" % (n, self.T)
        return {"value": n * self.T / dt, "unit": self.unit, "cores": 1, "kind": "port", "sample": what}, dt * 1e3

    def roofline(self, k_ms, peaks):
        gbs = self.b_alg * self.units() / (k_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "kernel": self.kernel, "kernel_ms_per_pass": k_ms,
                "bytes_per_unit": self.b_alg, "peak_source": peaks["hbm_source"]}


class FactorialConfig:
    """C3: 4 ring-coupled chains x 16 states (three 16^6-entry cliques, 403 MB of tables), the
    E-step that niptrain's em_learn runs.  One step = the E-step over this GPU's series; the
    time is linear in the number of series (64 sequences are in flight at a time), so a
    bounded per-GPU sample stands for the 16 384-series set.  Engine 3 evaluates the join tree
    factor by factor (the cliques are never materialised); NIPGPU_FACTOR=0 / engine 1 streams
    the materialised tables instead (round 1: 572 slice-steps/s)."""

    def __init__(self, name="C3", n_series=128, T=8):
        self.name, self.n_series, self.T = name, n_series, T
        self.metric, self.unit = "slice-steps/sec (EM E-step)", "slice-steps/s"
        self.workload = ("%s: factorial DBN 4 x 16 states ring-coupled, E-step over %d series x %d slices per GPU "
                         "(sample of the 16384-series set)" % (name, n_series, T))
        self.bound = "fp64"
        self.f_alg = 2 * 1.3e8  # 4 stages x 2 x 16^6 flops per direction, factored evaluation
        self.b_alg = 1.0e6      # 16 |I| bytes of forward rows per slice-step
        self.kernel = "k_fac_contract / k_fac_contract2 (multi-operand contractions, ~12 of 16^6 terms per slice-step)"
        self.engine = "engine 3: join tree factor by factor (clique tables never materialised)"

    def build(self, seed_data):
        from nip_b200.synth import FactorialSpec
        self.spec = FactorialSpec(16, 4, seed=1)
        self.data = self.spec.sample(self.n_series, self.T, seed=seed_data)
        self.obs_vars, self.query = self.spec.obs_vars, [4, 5]
        return self.spec.flat()

    def units(self):
        return self.n_series * self.T

    def device_step(self, model, batch):
        batch.estep(want_counts=False)
        return model.last_kernel_ms()[0]

    def roofline(self, k_ms, peaks):
        tf = self.f_alg * self.units() / (k_ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": tf, "peak": peaks["dfma_tf"], "unit": "TFLOP/s",
                "frac": tf / peaks["dfma_tf"], "kernel": self.kernel, "kernel_ms_per_pass": k_ms,
                "flops_per_slice_step": self.f_alg,
                "note": "FP64 FMA pipe (no FP64 tensor shape fits these 16-wide gathers); the kernels execute "
                        "~9e8 flops per slice-step (every message is one 16^6-term multi-operand contraction) "
                        "and are issue-bound on index arithmetic, not on either roof",
                "peak_source": "DFMA rate measured live by nipgpu_probe_peaks"}

    def parity(self, api, model, batch):
        return {"checked_series": 0, "ok": None,
                "note": "the reference needs ~50 s per slice-step at this shape; parity at full size is "
                        "tests/test_gpu_parity.py::test_c3_full_size_vs_oracle (engines 1 and 3 against the "
                        "oracle), at 3/6/8 states test_factorial_vs_oracle and the factorial4x3.json fixture"}


def make_config(name):
    if name == "C1":
        return SmallModelConfig("C1", 365, 24, likelihood=False)
    if name == "C2":
        return HmmConfig("C2", 64, 32, 4096, 1000)
    if name == "C3":
        return FactorialConfig()
    if name == "C4":
        return HmmConfig("C4", 1024, 64, 4096, 100, cpu_T=6, parity_series=2)
    if name == "C5":
        return SmallModelConfig("C5", 1000000, 50, likelihood=True)
    raise SystemExit("unknown config " + name)


def read_peaks(api, device):
    dmma_tf, dfma_tf, copy_gbs = api.probe_peaks(device)
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    return {"dmma_tf": dmma_tf, "dfma_tf": dfma_tf, "copy_gbs": copy_gbs,
            "hbm_gbs": peaks_file.get("hbm_gbs", 6650.0),
            "hbm_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks_file else "fallback of B200_PROFILING.md"}


def side_config(name, api, device, peaks, steps=3):
    """another BASELINE configuration, device-timed, with its roofline and a parity flag"""
    t0 = time.perf_counter()
    cfg = make_config(name)
    if name == "C1":
        pass
    fm = cfg.build(2)
    model = api.Model(fm, device=device, engine=api.ENGINE_AUTO)
    batch = model.batch(cfg.obs_vars, cfg.data)
    if name == "C3":
        model.mstep(np.random.default_rng(7).random(model.counts_size()) + 0.1)
    for _ in range(2):
        cfg.device_step(model, batch)
    ms = [cfg.device_step(model, batch) for _ in range(steps)]
    k_ms = float(np.mean(ms))
    out = {"workload": cfg.workload, "metric": cfg.metric, "unit": cfg.unit, "engine": cfg.engine,
           "value": cfg.units() / (k_ms * 1e-3), "ms_per_step": k_ms, "timing": "CUDA events around the pass, mean of %d" % steps,
           "roofline": cfg.roofline(k_ms, peaks)}
    if name != "C3":
        import torch
        cfg.e2e_buffers(torch)
        cfg.e2e_step(batch)
        out["parity"] = cfg.parity(api, model, batch)
    else:
        out["parity"] = cfg.parity(api, model, batch)
    if name == "C1":    # the same model at a size that fills the machine
        big = np.random.default_rng(3).integers(0, 5, size=(1000000, 50, 1), dtype=np.int32)
        bb = model.batch(cfg.obs_vars, big)
        for _ in range(2):
            bb.infer_device(cfg.query)
        bms = []
        for _ in range(steps):
            bb.infer_device(cfg.query)
            bms.append(model.last_kernel_ms()[0])
        b_ms = float(np.mean(bms))
        gbs = cfg.b_alg * 5e7 / (b_ms * 1e-3) / 1e9
        out["scaled"] = {"workload": "same model, 1 000 000 series x 50 slices", "value": 5e7 / (b_ms * 1e-3),
                         "ms_per_step": b_ms, "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                                           "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]}}
        bb.close()
    batch.close()
    model.close()
    out["wall_s"] = time.perf_counter() - t0
    return out


# --------------------------------------------------------------------- EM ---
def em_legs(api, torch, dist, cfg, model, batch, rank, world, local_rank, iters, barrier):
    """EM iterations/s on the C2 model: weak (this rank's own 4096 x 1000) and strong (one fixed
    16384 x 1000 set, sharded), plus the multi-GPU parity check of the reduced counts."""
    from nip_b200.dist import EmWorker, GpuEmBackend, shard_series
    out = {}
    rng = np.random.default_rng(7)
    init = rng.random(model.counts_size()) + 0.1     # same random start on every rank

    def timed(worker):
        worker.iteration()
        barrier()
        t = time.perf_counter()
        ll = 0.0
        for _ in range(iters):
            ll, _bad = worker.iteration()
        barrier()
        return (time.perf_counter() - t) / iters, ll

    model.mstep(init)
    em_s, em_ll = timed(EmWorker(GpuEmBackend(model, batch), rank, world))
    out["weak"] = (em_s, em_ll)

    # ---- strong scaling: ONE fixed set drawn on the device from the model, same on every rank ----
    n_strong, T = 16384, cfg.T
    model.set_parameters(*_flat_params(cfg))
    full = model.sample(n_strong, T, seed=11)[:, :, cfg.obs_vars[0]:cfg.obs_vars[0] + 1]
    mine = shard_series(np.full(n_strong, T), world)[rank]
    sb = model.batch(cfg.obs_vars, np.ascontiguousarray(full[mine]))
    model.mstep(init)
    st_s, st_ll = timed(EmWorker(GpuEmBackend(model, sb), rank, world))
    sb.close()
    out["strong"] = (st_s, st_ll, n_strong, T)

    # ---- parity of the multi-GPU path: sharded + all-reduced counts == the one-GPU E-step ----
    n_chk, T_chk = 1024, 64
    chk = full[:n_chk, :T_chk]
    model.mstep(init)
    one = model.batch(cfg.obs_vars, np.ascontiguousarray(chk))
    want, ll_want, _ = one.estep(add_pseudocount=True)
    one.close()
    err = 0.0
    if world > 1:
        part = shard_series(np.full(n_chk, T_chk), world)[rank]
        pb = model.batch(cfg.obs_vars, np.ascontiguousarray(chk[part]))
        be = GpuEmBackend(model, pb)
        counts = be.estep(add_pseudocount=(rank == 0))
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        got = counts.cpu().numpy()
        pb.close()
        err = max(max_rel_err(got[:-2], want), max_rel_err(got[-2], ll_want))
    out["check"] = {"series": n_chk, "slices_each": T_chk, "ranks": world, "max_rel_err": err, "tolerance": 1e-12,
                    "what": "all-reduced expected counts + loglik of the sharded set vs the one-GPU E-step of the whole set",
                    "ok": bool(err <= 1e-12)}
    return out


def _flat_params(cfg):
    fm = cfg.spec.flat()
    return fm.clique_tables, fm.var_prior


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
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs-block", action="store_true")
    ap.add_argument("--no-em", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--em-iters", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = make_config(args.config)

    if args.impl == "reference":
        if rank != 0:
            return
        if args.config == "C3":
            print(json.dumps({"impl": "reference", "unavailable": "C3 needs ~50 s per slice-step on the CPU"}),
                  file=json_out, flush=True)
            return
        cfg.build(2)
        base, ms = cfg.cpu_baseline(steps=max(1, args.steps), warmup=args.warmup, variants=False)
        print(json.dumps({
            "impl": "reference", "metric": cfg.metric, "value": base["value"], "unit": cfg.unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg.workload, "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": cfg.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist
    import nip_b200.api as api

    torch.cuda.set_device(local_rank)
    placement = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.load_library()                      # raises if the CUDA library is missing: no fallback

    fm = cfg.build(2 + rank)                # weak scaling: every rank works on its own set
    model = api.Model(fm, device=local_rank, engine=api.ENGINE_AUTO)
    batch = model.batch(cfg.obs_vars, cfg.data)
    if args.config == "C3":
        model.mstep(np.random.default_rng(7).random(model.counts_size()) + 0.1)
    stream = torch.cuda.ExternalStream(model.L.nipgpu_model_stream(model.h), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                     # nvidia-smi needs ~1 s to start streaming: start before warm-up
    for _ in range(W):
        cfg.device_step(model, batch)
    barrier()
    sampler.mark()                      # samples from here on belong to the timed regions
    api.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        kernel_ms.append(cfg.device_step(model, batch))
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = api.launch_count()
    if args.config == "C5":             # the likelihood ABI is host-buffered: the device-timed figure is the kernels'
        dev_ms = float(np.sum(kernel_ms))

    # ---- end to end through the host-buffer entry point (pinned host memory) ----
    e2e_ms, h2d, d2h, parity = None, 0, 0, None
    if args.config != "C3":
        h2d, d2h = cfg.e2e_buffers(torch)
        cfg.e2e_step(batch)
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.e2e_steps):
            cfg.e2e_step(batch)
        barrier()
        e2e_ms = (time.perf_counter() - t1) / args.e2e_steps * 1e3
    clocks = sampler.stop()             # covers the device-timed steps and the end-to-end steps

    # ---- parity gate: nothing below counts unless the GPU results match the CPU checker ----
    if rank == 0:
        parity = cfg.parity(api, model, batch)
        if parity.get("ok") is False:
            log("PARITY FAILURE:", json.dumps(parity))
            raise SystemExit(3)

    em = None
    if args.config == "C2" and not args.no_em:
        em = em_legs(api, torch, dist, cfg, model, batch, rank, world, local_rank, args.em_iters, barrier)
        if not em["check"]["ok"]:
            log("MULTI-GPU PARITY FAILURE:", json.dumps(em["check"]))
            raise SystemExit(3)

    t = [dev_ms, e2e_ms or 0.0, wall * 1e3]
    if em:
        t += [em["weak"][0] * 1e3, em["strong"][0] * 1e3]
    times = torch.tensor(t, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    times = [float(x) for x in times.cpu()]
    dev_ms, e2e_ms, wall_ms = times[:3]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    units = cfg.units() * world
    value = units * args.steps / (dev_ms * 1e-3)
    peaks = read_peaks(api, local_rank)
    k_ms = float(np.mean(kernel_ms))                      # dominant kernels of one pass (CUDA events on the library's stream)
    roof = cfg.roofline(k_ms, peaks)
    if args.config == "C2":
        # ncu dram__bytes of one pass (fwd + bwd), captured once per round by tools/prof_c2.py under
        # `ncu --set full`; a profiler cannot run inside this timed process
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            roof["traffic"] = tj["chain_pass_dram_bytes"]
            roof["traffic_source"] = tj.get("source", "profiles/traffic.json (ncu --set full, one pass)")
        except (OSError, ValueError, KeyError):
            roof["traffic"] = None
    else:
        roof["traffic"] = None
    out = {
        "metric": cfg.metric, "value": value, "unit": cfg.unit, "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg.workload, "engine": cfg.engine,
                   "l2_policy": "working set per pass >> 126 MB L2 (no flush needed)" if cfg.units() * cfg.b_alg > 5e8 else
                                "inputs smaller than L2: a launch-latency-bound case, reported as such",
                   "per_gpu_units_per_step": cfg.units(), "wall_ms_per_step": wall_ms / args.steps,
                   "host_placement": placement},
        "roofline": roof, "parity": parity,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if e2e_ms:
        out["e2e"] = {"value": units / (e2e_ms * 1e-3), "unit": cfg.unit, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                      "d2h_gbs_per_rank": d2h / (e2e_ms * 1e-3) / 1e9}
    if em:
        em_ms, st_ms = times[3], times[4]
        n_strong, Ts = em["strong"][2], em["strong"][3]
        out["em"] = {"metric": "EM iterations/s (E-step over the whole set + all-reduce + M-step)", "scaling": "weak",
                     "value": 1e3 / em_ms, "unit": "iter/s", "ms_per_iteration": em_ms,
                     "slice_steps_per_s": units / (em_ms * 1e-3),
                     "workload": "same model and data, %d x %d slices per GPU, %d GPU(s)" % (cfg.n_series, cfg.T, world),
                     "allreduce_doubles": int(model.counts_size() + 2), "loglik_per_slice": em["weak"][1] / units}
        out["em_strong"] = {"metric": "EM iterations/s on ONE fixed set sharded over the GPUs", "scaling": "strong",
                            "value": 1e3 / st_ms, "unit": "iter/s", "ms_per_iteration": st_ms,
                            "slice_steps_per_s": n_strong * Ts / (st_ms * 1e-3),
                            "workload": "%d series x %d slices drawn on the device from the C2 model (seed 11), "
                                        "%d per GPU" % (n_strong, Ts, n_strong // world),
                            "loglik_per_slice": em["strong"][1] / (n_strong * Ts),
                            "self_check": "loglik_per_slice after the same number of iterations must agree across N"}
        out["parity"]["multi_gpu"] = em["check"]
    if not args.no_cpu_baseline and args.config != "C3":
        out["cpu_baseline"], _ = cfg.cpu_baseline()
    if world == 1 and args.config == "C2" and not args.no_configs_block:
        block = {}
        for name in ("C1", "C5", "C4", "C3"):
            try:
                block[name] = side_config(name, api, local_rank, peaks)
            except Exception as e:            # noqa: BLE001 - the headline line must survive a side config
                block[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        out["configs"] = block
    print(json.dumps(out), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
