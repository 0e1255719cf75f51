#!/usr/bin/env python
"""Benchmark of the IR decay / spectral-decay hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

Metric: IR-seconds analysed per second for third-octave RT60 + broadband EDC + CSD waterfall.
Workload: config 5 of BASELINE.json (6 s @ 48 kHz mono IRs, 288,000 samples), sharded by IR:
each GPU analyses `--rows` IRs per step (default 8192 = the 8-GPU shard of the 65,536-IR config),
so N GPUs analyse N * rows IRs per step (weak scaling, no data-path collective; one packed NCCL gather
of the metric tables per step, inside the timed step).  The full 65,536-IR config at N = 1 is a rate metric
over the same rows: it is not run as one batch.

One step = decay (EDC curve + EDT/T20/T30 fits + early-decay time) + rt60bands (third-octave, T30) +
waterfall (18 slices x 1705 bins) over the whole batch.  `value` is timed with the inputs resident in
HBM; `e2e` through the public batch API from pinned host memory with the result tables copied back.
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

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SR = 48000
N_SAMPLES = 288000
METRIC = "ir_seconds_analysed_per_second"
UNIT = "IR-s/s"


# ----------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference algorithm) - used by cpu_baseline and --impl reference
# ----------------------------------------------------------------------------------------------


def _cpu_one_row(row):
    """The reference's three analyses on one channel (oracle/ref_port.py restates analyse/*.py)."""
    import oracle

    d = oracle.decay_metrics(row, SR, oracle.DecayParams(compute_edt=True))
    b = oracle.band_metrics(row, SR, oracle.BandParams(band_mode="third"))
    # the reference computes every STFT frame (waterfall.py:220); time it the way the reference runs
    w = oracle.waterfall(row, SR, oracle.CsdParams(), only_selected_frames=False)
    return d["start"], len(b["metrics"]), w["rel_db"].shape


def _cpu_init():
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"


def cpu_throughput(rows_np, cores, repeats=1, start_method="spawn"):
    """IR-seconds per wall second of the CPU path on `cores` worker processes.

    spawn, not fork: the calling process holds a CUDA context."""
    import multiprocessing as mp

    ctx = mp.get_context(start_method)
    with ctx.Pool(processes=cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_one_row, [rows_np[0]] * cores)  # warm the workers (imports, FFT plans)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_one_row, list(rows_np), chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return rows_np.shape[0] * (N_SAMPLES / SR) / best, best


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout_s=8.0):
        """Block until the first sample has arrived (nvidia-smi takes a second or more to start on an 8-GPU box)."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def mark(self):
        """Samples from here on belong to the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        # samples of the timed region; a region shorter than the 100 ms sampling period falls back to the samples
        # taken under the same load during the warm-up steps (and says so)
        window = self.lines[getattr(self, "first", 0):]
        source = "timed region"
        if not window:
            window, source = self.lines, "warm-up + timed region (timed region shorter than the sampling period)"
        for line in window:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "sampled_during": source, "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# this build
# ----------------------------------------------------------------------------------------------


def measured_peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from audio_analysis_b200 import _native as nat
    from audio_analysis_b200 import decay, distributed, engine, rt60bands, waterfall
    from audio_analysis_b200.pipeline import run_hot_path
    from audio_analysis_b200.synth import synth_batch_torch

    rank, world, local = distributed.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"note: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: keep the process (and the pinned staging buffers it allocates) on the GPU's NUMA node.
    # Not at N = 1, where the CPU baseline below wants every host core.
    numa_cpus = None
    if world > 1 and not os.environ.get("AA_BENCH_NO_NUMA"):
        numa_cpus = distributed.bind_host_to_gpu(dev)
    lib = nat.load()
    rows_n = int(args.rows)
    n = int(args.samples)

    ds = decay.DecayAnalysisSettings(compute_edt=True)
    bs = rt60bands.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
    ws = waterfall.WaterfallAnalysisSettings()

    # synthetic input, generated on the device; rank r owns IR indices [r*rows, (r+1)*rows)
    x_dev = synth_batch_torch(rows_n, n, SR, first_index=rank * rows_n, device=dev)
    rows = engine.make_rows(x_dev)
    torch.cuda.synchronize()

    def step_device(bands_options=None):
        return run_hot_path(rows, SR, ds, bs, ws, want_edc=True, bands_options=bands_options)

    # the one collective (N > 1): the metric tables of every rank, packed into one record per row, one dist.gather per
    # step INSIDE the timed step.  The gatherer is built and first used during warm-up: the first NCCL call on the
    # communicator sets up the peer channels (hundreds of ms, reported as gather_first_call_ms).
    gatherer = None

    def tables_of(o):
        return {"band_rt60": o.band_metrics, "fits": o.decay.fits, "early10": o.decay.early10, "start": o.decay.start}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_ready()
    for _ in range(max(3, args.warmup)):
        out = step_device()
    gather_first_ms = None
    if world > 1:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gatherer = distributed.TableGatherer(tables_of(out), rows_n)
        gatherer.gather(tables_of(out), unpack=False)
        g1.record()
        barrier()
        gather_first_ms = g0.elapsed_time(g1)
        for _ in range(2):
            gatherer.gather(tables_of(out), unpack=False)
    barrier()
    launches0 = lib.aa_kernel_launch_count()
    lib.aa_profile_enable(1)   # CUDA events around every library kernel, on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    gather_events = []
    e0.record()
    for _ in range(args.steps):
        out = step_device()
        if gatherer is not None:
            ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ga.record()
            gatherer.gather(tables_of(out), unpack=False)
            gb.record()
            gather_events.append((ga, gb))
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    band_levels = dict(engine.last_bands_stats)
    gather_ms = (sum(a.elapsed_time(b) for a, b in gather_events) / len(gather_events)) if gather_events else None
    lib.aa_profile_collect()
    stage_names = ["edc", "fft_cols", "fft_rows", "untangle", "fft_rows_untangle", "fft_cols_band", "fft_rows_band",
                   "band_fit", "band_fused", "band_edc", "csd"]
    stage_ms = {k: lib.aa_profile_ms(k.encode()) / args.steps for k in stage_names}
    stage_scopes = {k: int(lib.aa_profile_scopes(k.encode())) // args.steps for k in stage_names}
    # band stage per decimation class (nested scopes around the class's kernels)
    class_ms = {f"D{d}": round(lib.aa_profile_ms(f"band_class_d{d}".encode()) / args.steps, 4) for d in (1, 2, 3, 4, 6, 8, 12, 16, 24)}
    class_ms = {k: v for k, v in class_ms.items() if v > 0.0}
    lib_refine_ms = lib.aa_profile_ms(b"band_refine") / args.steps
    lib.aa_profile_enable(0)
    launches = (lib.aa_kernel_launch_count() - launches0) // max(1, args.steps)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    ir_seconds_per_step = world * rows_n * (n / SR)
    value = ir_seconds_per_step / (ms_step / 1e3)

    # ---- roofline objects: per-kernel device time (summed over the launches of one step, measured above inside the
    # timed region) against each kernel's ALGORITHMIC bytes
    peak, peak_src = measured_peak_gbs()
    n_bins = int(out.csd.shape[2])
    n_bands = int(out.band_metrics.shape[1])
    units = rows_n * n_bands
    # samples synthesised per row over all bands: bands with a low mask are built on every 2nd / 4th sample
    dec = rt60bands.decimation_classes(out.bands, n, SR, bs.transition_width_octaves, settings=bs)
    band_samples = float(sum(n // d for d in dec))          # per row, all bands
    band_units_total = rows_n * band_samples                # replaces units * n below
    # classes whose transform fits one CTA (csrc/bandfused.cu: M = n / 2D in its plan list) never leave shared memory
    fused_lengths = {9000, 12000, 6000, 8000}
    fused_samples = float(sum(n // d for d in dec if d > 1 and n % (2 * d) == 0 and n // (2 * d) in fused_lengths))
    if os.environ.get("AA_NO_FUSED_BANDS"):
        fused_samples = 0.0
    band_units_total -= rows_n * fused_samples              # the three-kernel path handles the rest
    # what the fused kernel must read per row: the bins of the row spectrum inside each of its bands' supports
    # (pass band + raised-cosine skirts, rt60bands.py:127-167; 8 B per bin) and three float64 results per unit
    fused_bytes_row = 0.0
    if fused_samples > 0.0:
        fac = 2.0 ** float(bs.transition_width_octaves)
        for b_def, d in zip(out.bands, dec):
            if not (d > 1 and n % (2 * d) == 0 and n // (2 * d) in fused_lengths):
                continue
            lo = 0.0 if b_def.low_edge_hz is None else float(b_def.low_edge_hz) / fac
            hi = 0.5 * SR if b_def.high_edge_hz is None else min(0.5 * SR, float(b_def.high_edge_hz) * fac)
            fused_bytes_row += max(0.0, hi - lo) * n / SR * 8.0 + 24.0
    stage_bytes = {
        # read x + write edc_db (8 B/sample)
        "edc": (8.0 * rows_n * n, "hbm", "edc_kernel<true>: fused EDC + crossings + fits (aa_decay_batch)"),
        # forward transform of every row: each pass reads and writes the row once
        "fft_cols": (8.0 * rows_n * n, "hbm", "fft_cols_reg: first pass of the forward real FFT"),
        "fft_rows": (8.0 * rows_n * n, "hbm", "fft_rows_reg: second pass of the forward real FFT"),
        "untangle": (8.0 * rows_n * n, "hbm", "untangle_kernel: packed spectrum -> half spectrum"),
        # second pass + untangle in one kernel: reads the scratch matrix, writes the half spectrum
        "fft_rows_untangle": (8.0 * rows_n * n, "hbm / latency",
                              "fft_rows_untangle_reg: second pass of the forward real FFT fused with the untangle step"),
        # per (row, band) unit: the masked spectrum is generated on the fly, the pass writes N/2 complex values
        "fft_cols_band": (4.0 * band_units_total, "fp32 issue (band source + butterflies); writes 4 B per synthesised sample",
                          "fft_cols_reg<band>: first pass of the masked inverse FFT"),
        # per unit: reads N/2 complex, writes N real samples + the energy partials
        "fft_rows_band": (8.0 * band_units_total + 4.0 * band_units_total / 16.0, "hbm",
                          "fft_rows_reg<energy>: second pass of the masked inverse FFT + tile energies"),
        "band_fit": (4.0 * band_units_total / 16.0, "latency (reads the partials and ~10% of the band signal)",
                     "band_fit_kernel: band EDC suffix scan + crossings + fits"),
        # one CTA per unit: masked spectrum -> in-shared-memory inverse -> Schroeder sum -> fits
        "band_fused": (rows_n * fused_bytes_row, "latency of its phase chain / fp32 issue (the band signal never leaves the SM; "
                       "algorithmic bytes = the support bins of the row spectrum it reads + its results)",
                       "band_fused_kernel: deep-decimation classes, synthesis + EDC + fits in one CTA"),
        "band_edc": (4.0 * units * n, "hbm", "edc_kernel<false> on band signals (generic lengths)"),
        "csd": (rows_n * (18 * 4096 * 4.0 + 18 * n_bins * 4.0), "hbm",
                "csd_tma_kernel: TMA-staged frames, FFT, dB, on-chip normalisation"),
    }
    def roof(nbytes, ms):
        gbs = nbytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        return {"achieved": round(gbs, 1), "frac": round(gbs / peak, 4), "ms": round(ms, 4)}
    by_kernel = {}
    for k in stage_names:
        if stage_ms[k] <= 0:
            continue
        nbytes, bound, what = stage_bytes[k]
        by_kernel[k] = {**roof(nbytes, stage_ms[k]), "bound": bound, "kernel": what, "launches_per_step": stage_scopes[k],
                        "algorithmic_bytes_per_step": nbytes, "share_of_step": round(stage_ms[k] / ms_step, 4)}
    dominant = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
    dk = by_kernel[dominant]
    # DRAM bytes the kernel really moved, from the committed ncu --set full capture (per analysed row), scaled to the
    # rows one launch of this run processes; null when the capture is missing or does not cover the stage
    traffic = None
    traffic_note = None
    try:
        prof = Path(__file__).resolve().parent / "profiles"
        tpath = prof / "r02_traffic.json" if (prof / "r02_traffic.json").exists() else prof / "r01_traffic.json"
        tj = json.loads(tpath.read_text())
        per_row = tj["stages"][dominant]["dram_bytes_per_row"]
        traffic = round(per_row * rows_n / max(1, stage_scopes[dominant]), 1)
        traffic_note = (f"profiles/{tpath.name}: {per_row:.0f} B of DRAM traffic per row for this stage "
                        f"({tj['stages'][dominant]['kernels']} kernels, {tj['rows']}-row capture) x rows / launches")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dk["kernel"], "stage": dominant, "achieved": dk["achieved"], "frac": dk["frac"],
                "ms": dk["ms"], "launches_per_step": dk["launches_per_step"],
                "ms_per_launch": round(dk["ms"] / max(1, dk["launches_per_step"]), 4),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "algorithmic_bytes_per_launch": dk["algorithmic_bytes_per_step"] / max(1, dk["launches_per_step"]),
                "traffic": traffic, "traffic_source": traffic_note, "share_of_step": dk["share_of_step"],
                "note": "dominant kernel by device time inside the timed region; the band stage as a whole has 4 B/sample "
                        "of algorithmic traffic (SURVEY 8d) and is bound by fp32 issue, not HBM (SURVEY D7): see "
                        "roofline_by_kernel"}
    ms_bands = sum(stage_ms[k] for k in ("fft_cols", "fft_rows", "untangle", "fft_rows_untangle", "fft_cols_band",
                                          "fft_rows_band", "band_fit", "band_fused", "band_edc"))
    by_kernel["rt60bands_stage"] = {**roof(4.0 * rows_n * n, ms_bands), "bound": "fp32 issue / intermediate traffic",
                                    "kernel": "whole band stage against SURVEY 8d's 4 B/sample (read x once)",
                                    "share_of_step": round(ms_bands / ms_step, 4)}

    # ---- the same step with every band at full rate (the reference's estimator, no decimation) and with the first
    # decimation level kept uncertified (round-1 behaviour): the price of the default's parity guarantee, side by side
    def timed_variant(opt, steps=3):
        step_device(opt)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step_device(opt)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
        return {"value": round(ir_seconds_per_step / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms, 3)}

    variants = None
    if not args.no_variants:
        variants = {"full_rate_bands": timed_variant(engine.bands_options(max_decimation=1)),
                    "uncertified_first_level": timed_variant(engine.bands_options(refine=False))}

    # ---- end to end through the public batch API: pinned host input -> chunked H2D overlapped with the kernels ->
    # tables and CSD slices D2H (run_hot_path_host)
    from audio_analysis_b200.pipeline import run_hot_path_host
    e2e = None
    try:
        x_host = torch.empty((rows_n, n), dtype=torch.float32, pin_memory=True)
        x_host.copy_(x_dev)
        torch.cuda.synchronize()
        chunk = int(os.environ.get("AA_BENCH_E2E_CHUNK", "512"))

        def step_e2e(prev=None):
            return run_hot_path_host(x_host, SR, ds, bs, ws, chunk_rows=chunk, want_edc=True, reuse=prev)

        res = step_e2e()
        res = step_e2e(res)     # warm: pinned result buffers exist and are reused from here on
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps_e2e = max(3, min(args.steps, 10))
        a.record()
        for _ in range(steps_e2e):
            res = step_e2e(res)
        b.record()
        barrier()
        t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_e2e = float(t2.item()) / steps_e2e
        e2e = {"value": round(ir_seconds_per_step / (ms_e2e / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms_e2e, 3),
               "h2d_bytes_per_step": int(res.h2d_bytes), "d2h_bytes_per_step": int(res.d2h_bytes),
               "chunk_rows": chunk,
               "note": "pipeline.run_hot_path_host: chunked H2D / kernels / D2H on three streams; EDC curves stay in "
                       "HBM (plot data on demand); metric tables and CSD slices are copied back"}
        # the same batch as 16-bit PCM (the reference's WAV sample format, decoded on the device): half the PCIe bytes
        try:
            x_pcm = torch.empty((rows_n, n), dtype=torch.int16, pin_memory=True)
            x_pcm.copy_((x_dev * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
            torch.cuda.synchronize()
            res = run_hot_path_host(x_pcm, SR, ds, bs, ws, chunk_rows=chunk, want_edc=True, reuse=res)
            barrier()
            a.record()
            for _ in range(steps_e2e):
                res = run_hot_path_host(x_pcm, SR, ds, bs, ws, chunk_rows=chunk, want_edc=True, reuse=res)
            b.record()
            barrier()
            t3 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            ms_pcm = float(t3.item()) / steps_e2e
            e2e["pcm16_input"] = {"value": round(ir_seconds_per_step / (ms_pcm / 1e3), 1), "unit": UNIT,
                                  "ms_per_step": round(ms_pcm, 3), "h2d_bytes_per_step": int(res.h2d_bytes),
                                  "d2h_bytes_per_step": int(res.d2h_bytes)}
            del x_pcm
        except Exception as exc:
            e2e["pcm16_input"] = {"value": None, "error": str(exc)[:200]}
        del x_host, res
        # headline e2e = the reference's wire format (16-bit PCM taps, include/analysis/recorder.hpp:55-90); the
        # float32 line stays beside it
        pcm = e2e.get("pcm16_input") or {}
        if pcm.get("value"):
            f32 = {k: e2e[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
            e2e.update({"value": pcm["value"], "ms_per_step": pcm["ms_per_step"],
                        "h2d_bytes_per_step": pcm["h2d_bytes_per_step"], "d2h_bytes_per_step": pcm["d2h_bytes_per_step"],
                        "input": "int16 PCM in pinned host memory, decoded on the device (sample / 32768, io.py:46-64)",
                        "float32_input": f32, "steps": steps_e2e})
            del e2e["pcm16_input"]
        else:
            e2e["input"] = "float32 in pinned host memory"
    except Exception as exc:  # pinned allocation can fail on small hosts
        e2e = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    # ---- the other GPU configurations of BASELINE.json, a few steps each (parity for these shapes is in tests/)
    other = {}
    if not args.no_other_configs:
        del x_dev, rows, out
        torch.cuda.empty_cache()
        try:
            other["config3_bundle"] = bench_config3(world, rank, dev, barrier)
        except Exception as exc:
            other["config3_bundle"] = {"error": str(exc)[:300]}
        try:
            other["config4_deconvolve_chain"] = bench_config4(world, rank, dev, barrier)
        except Exception as exc:
            other["config4_deconvolve_chain"] = {"error": str(exc)[:300]}
        x_dev = synth_batch_torch(min(rows_n, 256), n, SR, first_index=rank * rows_n, device=dev)   # for the CPU sample

    # ---- CPU baseline: the oracle port on the host cores, bounded sample (rank 0, single GPU run only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        sample = max(8, 4 * min(cores, 64))   # ~25 s of CPU work: four IRs per core
        sample_rows = x_dev[:sample].cpu().numpy()
        v, wall = cpu_throughput(sample_rows, cores=min(cores, sample))
        cpu = {"value": round(v, 2), "unit": UNIT, "cores": min(cores, sample), "kind": "port",
               "sample": f"{sample} of the same IRs ({sample * n / SR:.0f} IR-s), one process per core, "
                         f"{wall:.2f} s wall; decay + third-octave rt60bands + waterfall of oracle/ref_port.py"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 carries and regression sums)",
            "data": "synthetic exponentially-decaying gaussian noise IRs, 0.25-0.5 s leading silence, -90 dB floor",
            "config": {"workload": f"config5 shard: {rows_n} x {n / SR:.0f} s @48 kHz mono IRs per GPU "
                                   f"(third-octave RT60 + broadband EDC/EDT/T20/T30 + CSD 18x{n_bins})",
                       "rows_per_gpu": rows_n, "samples_per_row": n, "bands": 26,
                       "band_decimation": {
                           "max": int(lib.aa_bands_get_decimation()), "ms_per_class": class_ms,
                           "bands_per_factor": {str(d): dec.count(d) for d in sorted(set(dec))},
                           "levels": band_levels.get("limits"), "units_per_level": band_levels.get("units"),
                           "redo_fraction": round(sum(band_levels.get("units", [0])[1:]) / max(1, units), 5),
                           "refine_ms": round(lib_refine_ms, 4),
                           "note": "every (row, band) fit of the first level is certified or redone at D <= 4, then at "
                                   "full rate (DESIGN.md 4.3); tests: 0 of 213k metrics outside 0.5 % / 1 ms"},
                       "e2e_returns": "metric tables + CSD slices (EDC curves stay in HBM, fetched on demand)",
                       "gather": ("one packed dist.gather of the metric tables per step, inside the timed step"
                                  if world > 1 else None),
                       "l2": "inputs (rows*samples*4 B per GPU) are larger than L2; no flush needed"},
            "host_binding": (f"{len(numa_cpus)} CPUs local to the GPU" if numa_cpus else None),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "roofline_by_kernel": by_kernel, "cpu_baseline": cpu,
            "variants": variants, "other_configs": other,
            "gather_ms": (round(gather_ms, 4) if gather_ms is not None else None),
            "gather_first_call_ms": (round(gather_first_ms, 2) if gather_first_ms is not None else None),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def _max_ms(ms, world, dev):
    import torch
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_config3(world, rank, dev, barrier, n_taps=256, seconds=4.0, steps=3):
    """BASELINE.json config 3: a bundle of 256 synthetic stereo 4 s taps in the recorder's wire format (16-bit PCM WAV,
    include/analysis/recorder.hpp:55-90) through `run_bundle_report` with decay + third-octave rt60bands: file reads
    (page cache), pinned staging, H2D, device decode, kernels, result records and the Markdown reports - wall clock of
    the call the user makes.  Under torchrun the taps are sharded over the ranks."""
    import json as _json
    import shutil
    import tempfile

    import numpy as np
    import torch
    import torch.distributed as dist

    from audio_analysis_b200 import engine, wire
    from audio_analysis_b200.bundle import BundleRunSettings, run_bundle_report
    from audio_analysis_b200.decay import DecayAnalysisSettings
    from audio_analysis_b200.pipeline import run_hot_path
    from audio_analysis_b200.report import ReportSettings
    from audio_analysis_b200.rt60bands import Rt60BandsAnalysisSettings
    from audio_analysis_b200.synth import synth_batch_torch

    frames = int(seconds * SR)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    root = Path(os.environ.get("AA_BENCH_BUNDLE_DIR") or os.path.join(base or tempfile.gettempdir(), "aa_bench_bundle"))
    names = [f"tap_{i:03d}" for i in range(n_taps)]
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
        (root / "taps").mkdir(parents=True)
        for i0 in range(0, n_taps, 64):
            x = synth_batch_torch(2 * min(64, n_taps - i0), frames, SR, first_index=100000 + 2 * i0, device=dev)
            pcm = (x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16).cpu().numpy()
            for k in range(pcm.shape[0] // 2):
                wire.write_pcm16_wav(root / "taps" / f"{names[i0 + k]}.wav", np.stack([pcm[2 * k], pcm[2 * k + 1]], axis=1), SR)
        (root / "meta.json").write_text(_json.dumps({"sample_rate_hz": SR, "length_samples": frames, "taps": names}))
    barrier()
    rs = ReportSettings(rt60_bands_settings=Rt60BandsAnalysisSettings(band_mode="third"), run_spectrogram=False,
                        run_waterfall=False, run_modal_cloud=False)
    settings = BundleRunSettings(report_settings=rs)
    run_bundle_report(root, settings)                      # warm: plans, pinned pools, page cache
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run_bundle_report(root, settings)
    torch.cuda.synchronize()
    wall_ms = _max_ms((time.perf_counter() - t0) * 1e3 / steps, world, dev)
    # the kernels alone on the same taps (already decoded, resident)
    lo, hi = (rank * n_taps) // world, ((rank + 1) * n_taps) // world
    st = wire.stage_pcm16([root / "taps" / f"{t}.wav" for t in names[lo:hi]])
    rows = engine.pcm16_to_rows(st.pcm.to(dev), st.frames, st.channels, False)
    ds = DecayAnalysisSettings()
    bs = Rt60BandsAnalysisSettings(band_mode="third")
    run_hot_path(rows, SR, ds, bs, None, want_edc=True, do_csd=False)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        run_hot_path(rows, SR, ds, bs, None, want_edc=True, do_csd=False)
    b.record()
    barrier()
    dev_ms = _max_ms(a.elapsed_time(b) / steps, world, dev)
    if world > 1:
        dist.barrier()
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    ir_s = n_taps * 2 * seconds
    return {"workload": f"bundle of {n_taps} stereo {seconds:.0f} s @48 kHz PCM16 taps ({n_taps * 2} channels x {frames} samples): "
                        "decay (T20/T30, EDC curves) + third-octave rt60bands per channel, reports written",
            "value": round(ir_s / (wall_ms / 1e3), 1), "unit": UNIT, "ms_per_bundle": round(wall_ms, 2),
            "timed": "run_bundle_report wall clock (WAV payloads -> pinned int16 -> device decode -> kernels -> "
                     "result records -> Markdown; EDC curves are fetched only when matplotlib can draw them), "
                     "max over ranks",
            "resident_kernels": {"value": round(ir_s / (dev_ms / 1e3), 1), "unit": UNIT, "ms": round(dev_ms, 3)},
            "file_seconds_per_s": round(n_taps * seconds / (wall_ms / 1e3), 1), "steps": steps}


def bench_config4(world, rank, dev, barrier, n_captures=64, steps=3):
    """BASELINE.json config 4: 64 synthetic captures of the 10 s log sweep of test_tones/sweep.wav's shape (13 s at
    48 kHz = 624,000 samples: 1 s silence + sweep + 2 s tail) -> deconvolution at n_fft = 2^20 (deconvolve.py:124-193)
    -> third-octave rt60bands + CSD waterfall on the 624,000-sample impulse responses; captures resident in HBM,
    sharded over the ranks."""
    import numpy as np
    import torch

    from audio_analysis_b200 import deconvolve as dv, engine
    from audio_analysis_b200.decay import DecayAnalysisSettings
    from audio_analysis_b200.pipeline import run_hot_path
    from audio_analysis_b200.rt60bands import Rt60BandsAnalysisSettings
    from audio_analysis_b200.synth import synth_batch_torch
    from audio_analysis_b200.waterfall import WaterfallAnalysisSettings

    n = 624000
    t = np.arange(10 * SR) / SR
    k = np.log(20000.0 / 20.0) / 10.0
    sweep = np.zeros(n, dtype=np.float32)
    sweep[SR:SR + t.size] = (0.5 * np.sin(2.0 * np.pi * 20.0 * (np.exp(k * t) - 1.0) / k)).astype(np.float32)
    lo, hi = (rank * n_captures) // world, ((rank + 1) * n_captures) // world
    mine = max(1, hi - lo)
    # captures = sweep convolved with synthetic rooms (circular convolution at 2^20 on the device; the 2 s of silence
    # after the sweep hold the tails)
    rooms = synth_batch_torch(mine, 96000, SR, first_index=200000 + lo, device=dev)
    nfft = 1 << 20
    S = torch.fft.rfft(torch.from_numpy(sweep).to(dev), n=nfft)
    cap = torch.fft.irfft(torch.fft.rfft(rooms, n=nfft) * S, n=nfft)[:, :n].contiguous()
    cap = (0.9 * cap / cap.abs().amax(dim=1, keepdim=True)).contiguous()
    rows = engine.make_rows(cap)
    ds, bs, ws = DecayAnalysisSettings(), Rt60BandsAnalysisSettings(band_mode="third"), WaterfallAnalysisSettings()
    sweep_t = torch.from_numpy(sweep)

    def step():
        irs, _ = dv.deconvolve_batch(rows, sweep_t, dv.DeconvolveSettings(), channels_per_file=1)
        return run_hot_path(irs, SR, ds, bs, ws, want_edc=False)

    step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = step()
    b.record()
    barrier()
    ms = _max_ms(a.elapsed_time(b) / steps, world, dev)
    ir_s = n_captures * n / SR
    return {"workload": f"{n_captures} captures x 13 s (624,000 samples) of a 10 s log sweep -> deconvolve (n_fft 2^20) -> "
                        "third-octave rt60bands (Bluestein-free 13-smooth length) + CSD 18 x 1705, resident in HBM",
            "value": round(ir_s / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms, 3), "scaling": "strong",
            "captures_per_gpu": mine, "steps": steps,
            "band_levels": dict(engine.last_bands_stats)}


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU algorithm (oracle port; /root/reference is not on the GPU box)
# ----------------------------------------------------------------------------------------------


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from audio_analysis_b200.synth import synth_batch_numpy

    cores = host_cores()
    sample = max(8, 4 * min(cores, 64))   # four IRs per core and step
    rows = synth_batch_numpy(sample, int(args.samples), SR)
    import multiprocessing as mp

    workers = min(cores, sample)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(processes=workers, initializer=_cpu_init) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_one_row, list(rows[:workers]), chunksize=1)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_one_row, list(rows), chunksize=1)
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = args.steps * sample * (int(args.samples) / SR) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": round(1e3 * total / args.steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (numpy)",
        "data": "synthetic exponentially-decaying gaussian noise IRs, 0.25-0.5 s leading silence, -90 dB floor",
        "config": {"workload": f"config5 rows: {sample} x {int(args.samples) / SR:.0f} s @48 kHz mono IRs per step "
                               "(third-octave RT60 + broadband EDC/EDT/T20/T30 + CSD), bounded sample of the GPU arm's "
                               "workload", "samples_per_row": int(args.samples), "bands": 26},
        "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{sample} IRs per step, one process per core; oracle/ref_port.py restates "
                                   "analyse/decay.py, rt60bands.py, waterfall.py (pinned bit-exact to the reference by "
                                   "tests/golden)"},
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rows", type=int, default=int(os.environ.get("AA_BENCH_ROWS", "8192")),
                    help="IRs per GPU per step (default 8192: config 5 sharded over 8 GPUs)")
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-variants", dest="no_variants", action="store_true",
                    help="skip the full-rate / uncertified comparison steps (used for ncu launch lists)")
    ap.add_argument("--no-other-configs", dest="no_other_configs", action="store_true",
                    help="skip the config-3 (bundle) and config-4 (deconvolve chain) lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
