#!/usr/bin/env python
"""bench.py -- consensus decisions/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W              # this build (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...    # CPU port of the reference path (rank 0 only)
  torchrun ... bench.py --gpus N ...                         # N ranks, decisions sharded by agent

A "step" is one pass of the hot path (K1 encode -> K2 finalise -> K3 similarity+vote) over one batch
of 256 decisions per GPU (BASELINE config 2: 8-model pool, 512-token candidates, dim 768).  All inputs
are synthetic (quoracle_b200/synth.py); the encoder is this build's definition with random weights.

  value  : whole-job decisions/s with token ids already resident in HBM (CUDA events, max over ranks);
           nothing but the library's own launches sits between the two events
  e2e    : the same metric through the host C-ABI calls qc_decide_batch_submit/_wait with pinned HOST
           buffers, the host->device copy of the step's token ids and the device->host read of its
           verdicts inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"

Extra keys of the same JSON line (none of them changes the headline leg):
  sustained      3000 more steps of the same workload: the clock / power regime a long job runs in
  extra.configs  short legs of BASELINE configs 3, 4 and two config-5 corners, each with its own roofline fraction
                 and its own parity record
  split_pool, tsplit (N > 1)   the two cross-GPU exchange paths with their bit-equality checks
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

UNIT = "decisions/s"
VOCAB = 32768
THRESHOLDS = (0.85, 0.995, 0.9995)      # 0.85 = the most common schema value; the other two make both verdicts occur
                                        # (cosines between unrelated candidates of this encoder sit near 0.99)
MARGIN = 2e-5                           # f64 margin below which an fp32 path may flip an index (reported, never dropped)
# BASELINE.json configs that fit this harness (the default, c2, is the one the metric is quoted on):
CONFIGS = {
    "c2": dict(pool=8, tokens=512, dim=768, agents=256, name="config 2"),
    "c3": dict(pool=8, tokens=512, dim=768, agents=512, name="config 3 (4096-agent tree over 8 GPUs = 512 agents per GPU)"),
    "c4": dict(pool=16, tokens=2048, dim=1536, agents=128, name="config 4 (1024 agents over 8 GPUs = 128 agents per GPU)"),
}
E2E_DEPTH = 4                    # batches in flight through qc_decide_batch_submit/_wait (<= QC_PIPELINE_SLOTS)
L2_BYTES = 126 * 2**20


def metric_name(c) -> str:
    return f"consensus decisions/sec ({c['pool']}-model, {c['tokens']}-tok)"


def workload(c) -> str:
    return (f"{c['name']}: {c['pool']}-model pool, {c['agents']} concurrent agents per GPU, {c['tokens']}-token "
            f"candidates, embedding dim {c['dim']}, vocab {VOCAB}, thresholds drawn from {list(THRESHOLDS)}")


def rotate_for(c) -> int:
    per_batch = c["agents"] * c["pool"] * c["tokens"] * 4
    return max(4, -(-(168 * 2**20) // per_batch))            # enough distinct batches to exceed the 126 MB L2


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region: NVML every 10 ms when pynvml is usable,
    else `nvidia-smi -lms 200`."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines: list[str] = []
        self.proc = None
        self.thread = None
        self.nvml = None                 # (module, handle) when NVML is usable: 10 ms sampling, no subprocess
        self.samples: list[tuple] = []
        self._stop = threading.Event()

    def _nvml_loop(self):
        nv, h = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nvml = (nv, nv.nvmlDeviceGetHandleByIndex(self.index))
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            sm = [x[0] for x in self.samples]
            reasons = {name for x in self.samples for bit, name in bits.items() if x[2] & bit}
            return {"sm_mhz": statistics.median(sm) if sm else None,
                    "sm_max_mhz": max(x[1] for x in self.samples) if self.samples else None,
                    "samples": len(sm), "reasons": sorted(reasons), "source": "nvml, 10 ms"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU legs: the oracle's C port of the path (encoder + reference rule + N x N medoid), all host threads, and the
# oracle's f64 verdict margins.  These four functions are the ONLY place bench.py touches oracle/ (cpu_baseline,
# the parity records and --impl reference).
# ---------------------------------------------------------------------------------------------------
def cpu_path(weights, token_ids, lens, thr, threads):
    from oracle import cport
    b, n, t = token_ids.shape
    emb = cport.encode(weights.table_bf16, weights.w_bf16, token_ids.reshape(b * n, t), lens.reshape(b * n),
                       threads=threads, f32=True, native=True)
    ok, med, _ = cport.similarity_batch(emb.reshape(b, n, -1), thr.astype(np.float64), want_full=True,
                                        threads=threads, want_cos=False, native=True)
    return ok, med, emb


def cpu_sim_only(emb, thr, threads, want_full=True, repeat=1):
    """Similarity stage alone on the host cores, fed the GPU-sized embeddings (the reference has no encoder to time:
    its embed step is a network call).  want_full=False is the reference rule only (N-1 cosines, consensus_rules.ex
    :277-286); True adds the full N x N matrix and the medoid vote."""
    from oracle import cport
    t0 = time.perf_counter()
    for _ in range(repeat):
        out = cport.similarity_batch(emb, thr.astype(np.float64), want_full=want_full, threads=threads, want_cos=False,
                                     native=True)
    return (time.perf_counter() - t0) / repeat, out


def cpu_margins(emb, thr):
    """f64 oracle verdicts and their margins (|cos - thr| of the row-0 rule, top-2 gap of the vote) for a set of
    embeddings [B, N, D]: which decisions an fp32 path may legitimately flip."""
    from oracle import numeric
    ok, med, _, margins = numeric.decide_batch(emb, np.asarray(thr, np.float64))
    return ok, med, margins


def cpu_threads() -> int:
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1, which would silently turn
    the CPU arm into a single-core run, so the count comes from the affinity mask, not from OpenMP."""
    from oracle import cport
    cport.load(native=True)
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def parity_record(gpu_ok, gpu_med, emb_cpu, thr, cpu_ok=None, cpu_med=None,
                  oracle="numpy f64 reference rule + medoid on the C port's f32 embeddings of the same tokens") -> dict:
    """GPU verdicts against the f64 oracle on `emb_cpu` [S, N, D]: mismatches over ALL S decisions, mismatches among
    the margin-safe ones (must be 0), and how many the margin filter set aside -- nothing is dropped silently."""
    ok, med, margins = cpu_margins(emb_cpu, thr)
    s = int(ok.shape[0])
    safe_t, safe_v = margins["thr"] > MARGIN, margins["vote"] > MARGIN
    rec = {
        "checked": s,
        "ref_ok": int((gpu_ok[:s][safe_t] != ok[safe_t]).sum()), "medoid": int((gpu_med[:s][safe_v] != med[safe_v]).sum()),
        "ref_ok_unfiltered": int((gpu_ok[:s] != ok).sum()), "medoid_unfiltered": int((gpu_med[:s] != med).sum()),
        "margin_rejected": {"thr": int((~safe_t).sum()), "vote": int((~safe_v).sum()), "margin": MARGIN},
        "ref_ok_ones": int((ok == 1).sum()), "ref_ok_zeros": int((ok == 0).sum()),
        "gpu_ref_ok_ones": int((gpu_ok[:s] == 1).sum()), "gpu_ref_ok_zeros": int((gpu_ok[:s] == 0).sum()),
        "medoid_histogram": np.bincount(med, minlength=emb_cpu.shape[1]).tolist(),
        "oracle": oracle,
    }
    if cpu_ok is not None:
        rec["c_port_vs_f64_oracle"] = {"ref_ok": int((cpu_ok != ok).sum()), "medoid": int((cpu_med != med).sum())}
    return rec


def run_reference(args):
    """--impl reference: the CPU port of the reference path on the box's host cores (rank 0 only).  Every step is the
    same 256-decision batch shape the GPU arm times (a bounded prefix of it only when the whole run would not finish
    within a few minutes on this box's cores); all host threads of the affinity mask, no auto-tuning of the count."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from quoracle_b200 import synth
    c = CONFIGS[args.config]
    b, n, t = c["agents"], c["pool"], c["tokens"]
    weights = synth.make_encoder(c["dim"], VOCAB, seed=synth.BASE_SEED + 2)
    batch = synth.make_decisions(b, n, t, VOCAB, seed=synth.BASE_SEED + 2, thresholds=THRESHOLDS)
    threads = cpu_threads()
    steps, warmup = args.steps, args.warmup
    cpu_path(weights, batch.token_ids[:1], batch.lens[:1], batch.thr[:1], threads)           # loads the library
    probe = min(b, max(4, (4 * threads + n - 1) // n))
    t0 = time.perf_counter()
    cpu_path(weights, batch.token_ids[:probe], batch.lens[:probe], batch.thr[:probe], threads)
    per = (time.perf_counter() - t0) / probe
    budget_s = 150.0                                   # keep the whole run within a few minutes
    sample = b if per * b * (steps + warmup) <= budget_s else max(1, int(budget_s / (per * (steps + warmup))))
    for _ in range(warmup):
        cpu_path(weights, batch.token_ids[:sample], batch.lens[:sample], batch.thr[:sample], threads)
    t0 = time.perf_counter()
    for k in range(steps):
        lo = (k * sample) % max(1, b - sample + 1)
        cpu_path(weights, batch.token_ids[lo:lo + sample], batch.lens[lo:lo + sample], batch.thr[lo:lo + sample], threads)
    dt = time.perf_counter() - t0
    value = sample * steps / dt
    desc = (f"{sample} of the {b} decisions of a step ({sample * n} candidates x {t} tokens) of the same workload; "
            f"C port of the path (f32 encoder + f64 reference rule + N x N medoid), OpenMP over candidates, "
            f"{threads} threads (the affinity mask)")
    line = {
        "impl": "reference", "metric": metric_name(c), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 encoder / f64 cosine", "data": "synthetic",
        "config": {"workload": workload(c), "decisions_per_step": sample,
                   "note": "the reference's own Elixir path cannot run here (no BEAM); "
                   "its embed step is a remote HTTP call with no local arithmetic, so the CPU arm times this "
                   "build's encoder definition ported to C plus the restated reference rule"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        tt = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in tt]


class DecideLeg:
    """One workload (a CONFIGS entry) resident on this rank's GPU: rotating input batches larger than L2, the
    device-timed loop, and a separate K1-only timing pass."""

    def __init__(self, env: Env, ctx, c, rotate=None, seed_off=0):
        from quoracle_b200 import synth
        torch = env.torch
        self.env, self.ctx, self.c = env, ctx, c
        self.b, self.n, self.t, self.dim = c["agents"], c["pool"], c["tokens"], c["dim"]
        self.rotate = rotate or rotate_for(c)
        seed = synth.BASE_SEED + 2 + seed_off
        # Decisions shard by agent: rank r owns its own agents' decisions; no data-path collective.
        self.batches = [synth.make_decisions(self.b, self.n, self.t, VOCAB, seed=seed + 1000 * env.rank + i,
                                             thresholds=THRESHOLDS) for i in range(self.rotate)]
        self.d_tok = [torch.from_numpy(x.token_ids).to(env.dev) for x in self.batches]
        self.d_len = [torch.from_numpy(x.lens).to(env.dev) for x in self.batches]
        self.d_thr = [torch.from_numpy(x.thr).to(env.dev) for x in self.batches]
        self.d_ok = torch.empty(self.b, dtype=torch.int32, device=env.dev)
        self.d_med = torch.empty(self.b, dtype=torch.int32, device=env.dev)
        # A dedicated non-default stream: handle 0 (torch's default stream) would mean "context stream" to
        # the C ABI, and the CUDA events below must sit on the stream the kernels are launched on.
        self.tstream = torch.cuda.Stream(device=env.dev)
        self.stream = self.tstream.cuda_stream
        self.flops = 2.0 * self.b * self.n * self.t * self.dim * self.dim      # algorithmic encoder flops per launch

    def step(self, i):
        j = i % self.rotate
        self.ctx.decide_batch_dev(self.d_tok[j].data_ptr(), self.d_len[j].data_ptr(), self.b, self.n, self.t,
                                  self.d_thr[j].data_ptr(), self.d_ok.data_ptr(), self.d_med.data_ptr(), 0, 0, self.stream)

    def timed(self, steps, warmup):
        """-> ms for `steps` steps (this rank), kernels launched in the region.  Per-kernel event timing is OFF here.

        The K steps are recorded once into a CUDA graph (qc_capture_*, chunks of at most 100 steps) and replayed
        between the two events: the same kernels on the same rotating inputs as a host loop of qc_decide_batch_dev
        calls, but with no host work between them -- at N = 8 the max over ranks of a 20-step host loop measured the
        slowest process's scheduling hiccups, not the GPUs.  `self.loop` says which form ran (host loop on capture
        failure)."""
        env, torch = self.env, self.env.torch
        env.barrier()
        for i in range(warmup):
            self.step(i)
        env.barrier()
        self.ctx.check_status()
        self.ctx.profile_enable(False)
        self.ctx.profile_reset()
        chunk = max(d for d in range(1, min(steps, 100) + 1) if steps % d == 0)
        graph = None
        try:
            graph = self.ctx.capture(self.stream, lambda: [self.step(warmup + i) for i in range(chunk)])
            per_chunk = self.ctx.launch_count()
            graph.launch(self.stream)                                # untimed replay: instantiation / upload effects
            torch.cuda.synchronize()
            self.loop = f"cuda graph of {chunk} steps ({per_chunk} kernel nodes) x {steps // chunk} launches"
        except Exception as e:                                       # noqa: BLE001 -- report and fall back, never hide
            graph = None
            self.loop = f"host loop (graph capture unavailable: {e})"
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        if graph is not None:
            ev0.record(self.tstream)
            for _ in range(steps // chunk):
                graph.launch(self.stream)
            ev1.record(self.tstream)
            env.barrier()
            launches = per_chunk * (steps // chunk)
        else:
            self.ctx.profile_reset()
            ev0.record(self.tstream)
            for i in range(steps):
                self.step(warmup + i)
            ev1.record(self.tstream)
            env.barrier()
            launches = self.ctx.launch_count()
        ms = ev0.elapsed_time(ev1)
        self.ctx.check_status()
        if graph is not None:
            graph.close()
        return ms, launches

    def kernel_pass(self, steps, first=0):
        """Same steps again with every launch bracketed by CUDA events on the launching stream (outside any timed
        region): per-kernel mean durations."""
        torch = self.env.torch
        steps = min(steps, 200)
        self.ctx.profile_reset()
        self.ctx.profile_enable(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graph = None
        try:
            # the brackets become event-record nodes of a graph: device-side durations, no host time inside a bracket
            graph = self.ctx.capture(self.stream, lambda: [self.step(first + i) for i in range(steps)])
            ev0.record(self.tstream)
            graph.launch(self.stream)
            ev1.record(self.tstream)
            self.k1_pass = "event-record nodes inside a CUDA graph of the same steps"
        except Exception as e:                                       # noqa: BLE001
            if graph is not None:
                graph.close()
                graph = None
            self.ctx.profile_reset()
            self.ctx.profile_enable(True)
            ev0.record(self.tstream)
            for i in range(steps):
                self.step(first + i)
            ev1.record(self.tstream)
            self.k1_pass = f"host loop with events ({e})"
        torch.cuda.synchronize()
        stats = self.ctx.kernel_stats()
        self.ctx.profile_enable(False)
        if graph is not None:
            graph.close()
        stats["pass_ms_per_step"] = ev0.elapsed_time(ev1) / max(1, steps)    # the event-bracketed pass's own step time
        self.ctx.check_status()
        # max over ranks, like every other multi-GPU time: the step is the slowest rank's, so K1 must be too
        names = [k for k, v in stats.items() if isinstance(v, dict)]
        means = [stats[k]["ms"] / stats[k]["launches"] if stats[k]["launches"] else 0.0 for k in names]
        red = self.env.max_over_ranks(*means, stats["pass_ms_per_step"])
        for k, m in zip(names, red[:-1]):
            stats[k]["ms"] = m * stats[k]["launches"]
        stats["pass_ms_per_step"] = red[-1]
        return stats

    def roofline(self, stats, step_ms, peaks, long_run, traffic=None, traffic_source=None):
        k1 = stats["encode"]
        k1_ms = k1["ms"] / max(1, k1["launches"])
        pass_ms = stats.get("pass_ms_per_step")
        # The event-bracketed K1 carries a few microseconds of launch latency that the pipelined step (graph replay,
        # programmatic dependent launch) hides, so on a fast box it can exceed the WHOLE timed step.  K1 cannot take
        # longer than the step that contains it: the duration used for `achieved` is the smaller of the two.
        k1_used = min(k1_ms, step_ms) if step_ms else k1_ms
        achieved = self.flops / (k1_used * 1e-3) / 1e12 if k1_used > 0 else 0.0
        peak = peaks["bf16_sustained"] if long_run else peaks["bf16_burst"]
        return {"bound": "tensor", "kernel": "qc_encode_kernel (K1)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_source": traffic_source,
                "peak_kind": ("sustained" if long_run else "burst") + ", " + peaks["source"],
                "k1_ms_per_launch": k1_ms, "k1_ms_used": k1_used,
                "k1_ms_source": "CUDA events around each K1 launch" if k1_used == k1_ms else
                                "bounded by ms_per_step: the event-bracketed K1 is longer than the whole pipelined step "
                                "(events expose launch latency the graph / PDL step hides)",
                "k1_share_of_step": k1_used / step_ms if step_ms else None,
                "k1_share_note": "k1_ms_used / ms_per_step of the timed region (a CUDA-graph replay without events); "
                                 "inside the event-bracketed pass itself, whose event nodes break the programmatic "
                                 "launch overlap, the share is k1_share_in_event_pass",
                "k1_share_in_event_pass": k1_ms / pass_ms if pass_ms else None,
                "k1_pass_ms_per_step": pass_ms,
                "k1_timing": "CUDA events around every K1 launch on the launching stream, in a separate pass of the "
                             "same steps right after the timed region: " + getattr(self, "k1_pass", "host loop"),
                "flops_per_launch": self.flops}


def short_config_leg(env: Env, name: str, ctx, steps: int, peaks, check: int):
    """extra.configs entry: a few steps of another BASELINE config on this rank's share, with its own roofline
    fraction and (rank 0) a parity record of `check` decisions against the oracle."""
    from quoracle_b200 import synth
    c = CONFIGS[name]
    leg = DecideLeg(env, ctx, c, rotate=max(4, min(rotate_for(c), 10)), seed_off=100)
    ms, launches = leg.timed(steps, 3)
    stats = leg.kernel_pass(steps, first=3)
    (ms,) = env.max_over_ranks(ms)
    step_ms = ms / steps
    rec = {"workload": workload(c), "steps": steps, "warmup": 3, "ms_per_step": step_ms,
           "value": env.world * leg.b * steps / (ms * 1e-3), "unit": UNIT, "gpu_launches": int(launches),
           "inputs": f"{leg.rotate} rotating batches of {leg.b * leg.n * leg.t * 4 / 1e6:.1f} MB", "timed_loop": leg.loop,
           "encoder": ctx.encoder_info(),
           "roofline": leg.roofline(stats, step_ms, peaks, long_run=False),
           "kernels_ms_per_launch": {k: (v["ms"] / v["launches"] if v["launches"] else None) for k, v in stats.items() if isinstance(v, dict)}}
    if env.rank == 0 and check > 0:
        x = leg.batches[0]
        r0 = ctx.decide_batch(x.token_ids[:check], x.lens[:check], x.thr[:check], want_cos=False)
        w = synth.make_encoder(c["dim"], VOCAB, seed=synth.BASE_SEED + 2)
        cok, cmed, emb = cpu_path(w, x.token_ids[:check], x.lens[:check], x.thr[:check], cpu_threads())
        rec["index_mismatch_vs_gpu"] = parity_record(r0.ref_ok, r0.medoid, emb.reshape(check, leg.n, -1), x.thr[:check],
                                                     cok, cmed)
    return rec


def similarity_corner(env: Env, ctx, n: int, d: int, b: int, peaks):
    """extra.configs.c5_*: K3 alone on embeddings resident in HBM (config 5 corner), L2 flushed between iterations,
    HBM roofline; parity of a 256-decision sample against the f64 oracle."""
    from quoracle_b200 import synth
    torch = env.torch
    base = synth.make_embeddings(min(b, 2048), n, d, seed=n * d)
    emb = torch.from_numpy(base).to(env.dev)
    if b > emb.shape[0]:
        emb = emb.repeat((b + emb.shape[0] - 1) // emb.shape[0], 1, 1)[:b].contiguous()
    rng = np.random.default_rng(n + d)
    # thresholds on both sides of the row-0 cosine cloud of these inputs, so both verdicts of the rule occur
    thr_h = rng.choice(np.asarray([-0.5, 0.0, 0.3, 0.5, 0.7] if n > 8 else [0.1, 0.3, 0.5, 0.7], np.float32), size=b)
    thr = torch.from_numpy(thr_h).to(env.dev)
    ok = torch.empty(b, dtype=torch.int32, device=env.dev)
    med = torch.empty(b, dtype=torch.int32, device=env.dev)
    cos = torch.empty((b, n, n), dtype=torch.float32, device=env.dev)
    stream = torch.cuda.Stream(device=env.dev)
    flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=env.dev)

    def run():
        ctx.similarity_batch_dev(emb.data_ptr(), b, n, d, thr.data_ptr(), ok.data_ptr(), med.data_ptr(), cos.data_ptr(),
                                 stream.cuda_stream)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        flush.zero_()                                   # evict the inputs from L2 between iterations
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run()
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ctx.check_status()
    ms = float(np.median(times))
    nbytes = 4 * n * d + 4 * n * n + 8
    gbs = b * nbytes / (ms * 1e-3) / 1e9
    s = min(256, b, base.shape[0])
    traffic, traffic_source = None, None
    tp = ROOT / "profiles" / "k1_traffic.json"            # static ncu numbers, labelled as such (never measured in this run)
    if tp.exists():
        tj = json.loads(tp.read_text()).get("k3_corners", {}).get(f"pool{n}_dim{d}_batch{b}")
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_source = "static: " + tj["source"] + " (NOT measured in this run)"
    rec = {"workload": f"config 5 corner: K3 only, pool {n}, dim {d}, batch {b}, embeddings resident in HBM "
                       f"({b * n * d * 4 / 1e6:.0f} MB), 256 MB L2 flush between iterations",
           "ms": ms, "value": b / (ms * 1e-3), "unit": UNIT,
           "roofline": {"bound": "hbm", "kernel": "qc_sim_small_kernel" if n <= 8 else "qc_sim_tc_kernel",
                        "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                        "bytes_per_decision": nbytes, "algorithmic_bytes": b * nbytes, "traffic": traffic,
                        "traffic_source": traffic_source, "peak_kind": peaks["source"]}}
    rec["index_mismatch_vs_gpu"] = parity_record(ok[:s].cpu().numpy(), med[:s].cpu().numpy(), base[:s], thr_h[:s],
                                                 oracle="numpy f64 reference rule + medoid on the same synthetic embeddings")
    return rec


def run_ours(args):
    from quoracle_b200 import exchange_legs, native, synth
    env = Env()
    torch, dist = env.torch, env.dist
    world, rank, local = env.world, env.rank, env.local
    c = CONFIGS[args.config]
    peaks = measured_peaks()

    weights = synth.make_encoder(c["dim"], VOCAB, seed=synth.BASE_SEED + 2)
    ctx = native.Context(local, weights.table_bf16, weights.w_bf16)
    if args.cta_group:
        ctx.set_cta_group(args.cta_group)
    leg = DecideLeg(env, ctx, c)
    b, n, t = leg.b, leg.n, leg.t

    # ---- headline: device-resident inputs, nothing but the library's launches between the events -----------
    sampler = ClockSampler(local).start()
    ms, launches = leg.timed(args.steps, args.warmup)
    main_loop = leg.loop
    # ---- K1 timing: the same steps again, per-kernel events, outside the timed region ---------------------
    stats = leg.kernel_pass(args.steps, first=args.warmup)

    # ---- e2e: host buffers through the C-ABI host entry point --------------------------------------------
    e2e_rotate = min(8, leg.rotate)
    pinned_tok = [native.PinnedArray((b, n, t), np.int32) for _ in range(e2e_rotate)]
    pin_len = [native.PinnedArray((b, n), np.int32) for _ in range(e2e_rotate)]
    pin_thr = [native.PinnedArray((b,), np.float32) for _ in range(e2e_rotate)]
    for k in range(e2e_rotate):
        pinned_tok[k].array[...] = leg.batches[k].token_ids
        pin_len[k].array[...] = leg.batches[k].lens
        pin_thr[k].array[...] = leg.batches[k].thr
    # Pipelined host API (qc_decide_batch_submit / _wait), E2E_DEPTH batches in flight: the H2D copy of one step
    # overlaps the GPU work of another and an encoder launch is always queued while the host turns a finished
    # slot around; every step still copies its own inputs from pinned host memory and reads its own verdicts
    # back, all inside the timed region.
    pin_ok = [native.PinnedArray((b,), np.int32) for _ in range(E2E_DEPTH)]
    pin_med = [native.PinnedArray((b,), np.int32) for _ in range(E2E_DEPTH)]
    outs = [native.DecideResult(pin_ok[k].array, pin_med[k].array, None, None) for k in range(E2E_DEPTH)]
    e2e_steps = args.steps

    def e2e_run(nsteps):
        for i in range(nsteps):
            k = i % E2E_DEPTH
            if i >= E2E_DEPTH:
                ctx.decide_batch_wait(k)
            j = i % e2e_rotate
            ctx.decide_batch_submit(k, pinned_tok[j].array, pin_len[j].array, pin_thr[j].array, outs[k])
        for i in range(max(0, nsteps - E2E_DEPTH), nsteps):
            ctx.decide_batch_wait(i % E2E_DEPTH)

    # warm-up touches every slot twice (first use of a slot allocates its stream, workspaces and pinned staging)
    e2e_run(max(2 * E2E_DEPTH, min(12, args.warmup)))
    env.barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    env.barrier()
    clocks = sampler.stop()
    gpu_ok0, gpu_med0 = None, None
    if rank == 0:
        r0 = ctx.decide_batch(leg.batches[0].token_ids, leg.batches[0].lens, leg.batches[0].thr, want_cos=False)
        gpu_ok0, gpu_med0 = r0.ref_ok.copy(), r0.medoid.copy()

    ms, e2e_s = env.max_over_ranks(ms, e2e_s)
    value = world * b * args.steps / (ms * 1e-3)
    e2e_value = world * b * e2e_steps / e2e_s
    step_ms = ms / args.steps

    # ---- sustained: the regime a long job runs in (clocks under power cap), same workload ------------------
    sustained = None
    if args.sustained_steps > 0:
        s2 = ClockSampler(local).start()
        sms, _ = leg.timed(args.sustained_steps, 3)
        sstats = leg.kernel_pass(min(200, args.sustained_steps), first=3)
        sclk = s2.stop()
        (sms,) = env.max_over_ranks(sms)
        k1s = min(sstats["encode"]["ms"] / max(1, sstats["encode"]["launches"]), sms / args.sustained_steps)
        tfl = leg.flops / (k1s * 1e-3) / 1e12
        sustained = {"steps": args.sustained_steps, "seconds": sms * 1e-3,
                     "value": world * b * args.sustained_steps / (sms * 1e-3), "unit": UNIT,
                     "ms_per_step": sms / args.sustained_steps, "sm_mhz": sclk.get("sm_mhz"),
                     "reasons": sclk.get("reasons"), "k1_tflops": tfl,
                     "frac_of_bf16_sustained": tfl / peaks["bf16_sustained"] if peaks["bf16_sustained"] else None,
                     "frac_of_bf16_burst": tfl / peaks["bf16_burst"] if peaks["bf16_burst"] else None,
                     "k1_note": "K1 mean over an event-bracketed pass of 200 steps right after the sustained run"}

    # ---- other BASELINE configs, short legs ---------------------------------------------------------------
    extra = {}
    if not args.no_extra:
        if args.config == "c2":
            extra["c3"] = short_config_leg(env, "c3", ctx, steps=max(5, min(10, args.steps)), peaks=peaks, check=32)
        w4 = synth.make_encoder(CONFIGS["c4"]["dim"], VOCAB, seed=synth.BASE_SEED + 2)
        ctx4 = native.Context(local, w4.table_bf16, w4.w_bf16)
        extra["c4"] = short_config_leg(env, "c4", ctx4, steps=5, peaks=peaks, check=2)
        if world == 1:
            extra["c5_pool8_dim768"] = similarity_corner(env, ctx, 8, 768, 65536, peaks)
            extra["c5_pool32_dim768"] = similarity_corner(env, ctx, 32, 768, 16384, peaks)
    else:
        ctx4 = None

    # ---- cross-GPU exchange paths (N > 1 only): driver-visible bit-equality + timing ----------------------
    split_rec = tsplit_rec = None
    if world > 1 and not args.no_exchange:
        split_rec = exchange_legs.split_pool_leg(ctx, env.dev, decisions=256, pool=8, tokens=512, iters=10)
        if ctx4 is None:
            w4 = synth.make_encoder(CONFIGS["c4"]["dim"], VOCAB, seed=synth.BASE_SEED + 2)
            ctx4 = native.Context(local, w4.table_bf16, w4.w_bf16)
        tsplit_rec = exchange_legs.tsplit_leg(ctx4, env.dev, decisions=32, pool=16, tokens=2048, iters=5)

    if rank == 0:
        traffic, traffic_source = None, None
        tp = ROOT / "profiles" / "k1_traffic.json"
        if tp.exists() and args.config == "c2":
            tj = json.loads(tp.read_text())
            traffic = tj.get("dram_bytes_per_launch")
            traffic_source = ("static: " + tj.get("source", "profiles/k1_traffic.json") +
                              " (one `ncu --set full` capture of this kernel on this workload; NOT measured in this run)")
        line = {
            "metric": metric_name(c), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 (fp32 accumulate; fp32 similarity)", "data": "synthetic",
            "config": {"workload": workload(c), "decisions_per_gpu_per_step": b,
                       "l2": f"inputs rotate over {leg.rotate} batches per rank ({leg.rotate * b * n * t * 4 / 1e6:.0f} MB "
                             f"> 126 MB L2); encoder table+weights ({(VOCAB * c['dim'] + c['dim'] ** 2) * 2 / 1e6:.0f} MB) are "
                             "persistent state",
                       "parallelism": f"agents sharded over {world} GPU(s), no data-path collective",
                       "timed_loop": main_loop,
                       "encoder": ctx.encoder_info(), "seed": synth.BASE_SEED + 2},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": b * n * t * 4 + b * n * 4 + b * 4,
                    "d2h_bytes_per_step": 2 * b * 4, "steps": e2e_steps,
                    "api": f"qc_decide_batch_submit/_wait (host C ABI, {E2E_DEPTH} batches in flight), pinned host buffers, wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": leg.roofline(stats, step_ms, peaks, long_run=ms > 2000.0, traffic=traffic,
                                     traffic_source=traffic_source),
            "kernels_ms_per_launch": {k: (v["ms"] / v["launches"] if v["launches"] else None) for k, v in stats.items() if isinstance(v, dict)},
        }
        if sustained is not None:
            line["sustained"] = sustained
        if extra:
            line["extra"] = {"configs": extra}
        if split_rec is not None:
            line["split_pool"] = split_rec
            line["tsplit"] = tsplit_rec
        if world == 1 and not args.no_cpu_baseline:
            threads = cpu_threads()
            x = leg.batches[0]
            cpu_path(weights, x.token_ids[:1], x.lens[:1], x.thr[:1], threads)          # loads the library
            probe = min(b, max(4, (4 * threads + n - 1) // n))
            t0 = time.perf_counter()
            cpu_path(weights, x.token_ids[:probe], x.lens[:probe], x.thr[:probe], threads)
            per = (time.perf_counter() - t0) / probe
            sample = max(1, min(b, int(args.cpu_seconds / max(per, 1e-6))))
            t0 = time.perf_counter()
            ok, med, emb = cpu_path(weights, x.token_ids[:sample], x.lens[:sample], x.thr[:sample], threads)
            dt = time.perf_counter() - t0
            e3 = emb.reshape(sample, n, -1)
            sim_dt, _ = cpu_sim_only(e3, x.thr[:sample], 1)
            rule_dt, _ = cpu_sim_only(e3, x.thr[:sample], 1, want_full=False)
            sim_all_dt, _ = cpu_sim_only(e3, x.thr[:sample], threads, repeat=20)
            line["cpu_baseline"] = {
                "value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"first {sample} decisions of batch 0 ({sample * n} candidates x {t} tokens), C port of the "
                          f"whole path, {dt:.1f} s",
                "similarity_only_1thread_decisions_per_s": sample / sim_dt if sim_dt > 0 else None,
                "reference_rule_only_1thread_decisions_per_s": sample / rule_dt if rule_dt > 0 else None,
                f"similarity_only_{threads}threads_decisions_per_s": sample / sim_all_dt if sim_all_dt > 0 else None,
                "index_mismatch_vs_gpu": parity_record(gpu_ok0, gpu_med0, e3, x.thr[:sample], ok, med),
            }
        print(json.dumps(line), flush=True)
    if ctx4 is not None:
        ctx4.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cta-group", type=int, default=0, help="force the K1 variant (1 or 2); 0 = library default")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="size of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.configs legs (c3, c4, c5 corners)")
    ap.add_argument("--no-exchange", action="store_true", help="skip the split_pool / tsplit legs at N > 1")
    ap.add_argument("--sustained-steps", type=int, default=3000, help="extra sustained leg (0 = off)")
    ap.add_argument("--e2e-depth", type=int, default=E2E_DEPTH, choices=[1, 2, 3, 4],
                    help="batches in flight in the e2e leg (host C ABI submit/wait slots)")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2",
                    help="BASELINE config (c2 = the headline; c3/c4 are the 8-GPU configs, per-GPU share)")
    args = ap.parse_args()
    globals()["E2E_DEPTH"] = args.e2e_depth
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
