"""CPU tests (`-m "not gpu"`): the C-ABI library loads and exports what include/*.h declares, the host
mirror reproduces the reference's band tables / slice selection / text formats, the CLI surface, and
the multi-rank gather over gloo.  No compute call reaches the GPU here."""

from __future__ import annotations

import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from conftest import load_golden

ROOT = Path(__file__).resolve().parent.parent
SR = 48000


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as entry
    entry.build()
    return True


def test_library_exports_every_declared_symbol(built):
    from audio_analysis_b200 import _native
    header = (ROOT / "include" / "audio_analysis_b200.h").read_text()
    declared = set(re.findall(r"\b(aa_[a-z0-9_]+)\s*\(", header))
    declared -= {"aa_status"}
    lib = _native.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert set(_native.EXPORTED_SYMBOLS) == declared, (set(_native.EXPORTED_SYMBOLS) ^ declared)
    assert lib.aa_abi_version() == 1
    assert lib.aa_kernel_launch_count() == 0


def test_struct_layouts_match_the_header(built):
    """ctypes mirrors of aa_decay_params / aa_band / aa_csd_params: sizes implied by the C declarations."""
    import ctypes as C
    from audio_analysis_b200 import _native as n
    assert C.sizeof(n.DecayParams) == 4 * 4 + 3 * 8 + 2 * 4 + 2 * 3 * 8
    assert C.sizeof(n.Band) == 2 * 4 + 2 * 8
    assert C.sizeof(n.CsdParams) == 6 * 4 + 2 * 8
    assert n.DecayParams.fit_hi_db.offset == 48 and n.DecayParams.fit_lo_db.offset == 72


def test_no_cpu_fallback_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from audio_analysis_b200 import decay
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        decay.analyse_decay_for_channel(np.zeros(256, np.float32), SR, "m", decay.DecayAnalysisSettings())


def test_fft_index_model_on_cpu(built):
    """The FFT passes' planner, butterflies and accessors, emulated serially against direct DFT sums."""
    exe = ROOT / "build" / "fft_model_check"
    assert exe.exists()
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "ALL OK" in out.stdout


@pytest.mark.parametrize("mode", ["three", "octave", "third"])
def test_band_tables_equal_the_oracle(mode):
    from audio_analysis_b200 import rt60bands as rb
    for kw in ({}, {"f_min_hz": 100.0, "f_max_hz": 5000.0}, {"low_upper_hz": 300.0, "mid_width_octaves": 1.0}):
        got = rb._build_band_definitions(rb.Rt60BandsAnalysisSettings(band_mode=mode, **kw), SR)
        want = oracle.band_table(oracle.BandParams(band_mode=mode, **kw), SR)
        assert [(b.name, b.centre_hz, b.kind, b.low_edge_hz, b.high_edge_hz) for b in got] == \
               [(w["name"], w["centre"], w["kind"], w["low"], w["high"]) for w in want]
    with pytest.raises(ValueError, match="Unknown band_mode"):
        rb._build_band_definitions(rb.Rt60BandsAnalysisSettings(band_mode="quarter"), SR)


def test_slice_selection_equals_the_oracle():
    from audio_analysis_b200 import waterfall as wf
    variants = [dict(), dict(slice_mode="uniform_frames", num_slices=7), dict(slice_mode="uniform_time", slice_spacing_seconds=0.08),
                dict(start_time_seconds=0.2, end_time_seconds=1.0, num_slices=30), dict(num_slices=1)]
    for kw in variants:
        s = wf.WaterfallAnalysisSettings(**kw)
        p = oracle.CsdParams(**kw)
        table, counts = wf.slice_table(s, SR, 300)
        for frames in (1, 2, 5, 17, 18, 19, 100, 299):
            ft = wf.frame_times_seconds(frames, s.hop_length, SR)
            want = oracle.pick_slices(ft, p)
            assert np.array_equal(wf._select_slice_frame_indices(ft, s), want)
            assert np.array_equal(table[frames, : counts[frames]], want)
    lo, freq = wf.frequency_selection(wf.WaterfallAnalysisSettings(), SR)
    assert lo == 2 and freq.size == 1705 and freq[0] == np.float32(23.4375)
    assert wf.analysed_span(1000, 100, wf.WaterfallAnalysisSettings(ignore_leading_seconds=0.001), SR) == (148, 852)


def test_text_summaries_follow_the_reference_format():
    """plots/example/verb_report.md:24-54,82-85 pins the formats."""
    from audio_analysis_b200 import decay, rt60bands as rb, waterfall as wf
    fit = decay.LinearDecayFit("T20", (-5.0, -25.0), 0.4713, 2.9514, -8.05884, -1.2, 0.999619, 7.4452)
    r = decay.ChannelDecayAnalysis("left", SR, 9287, np.zeros(0, np.float32), np.zeros(0, np.float32), 1.0557, {"T20": fit})
    text = decay.summarise_decay_results_text([r])
    assert text == ("[left] analysis_start_sample_index=9287\n  early_0_to_-10_time=1.0557s\n  EDT: NA\n"
                    "  T20: range=[-5.0,-25.0]dB time=[0.4713,2.9514]s slope=-8.058840dB/s r2=0.999619 rt60=7.4452s\n"
                    "  T30: NA\n")
    empty = decay.ChannelDecayAnalysis("mono", SR, 0, np.zeros(0, np.float32), np.zeros(0, np.float32), None, {})
    assert decay.summarise_decay_results_text([empty]) == "[mono] analysis_start_sample_index=0\n  early_0_to_-10_time=NA\n  fits=NA\n"
    bands = rb._build_band_definitions(rb.Rt60BandsAnalysisSettings(), SR)
    res = rb.Rt60BandsChannelResult("left", SR, bands, {"Low": rb.Rt60BandMetrics(86.0191, None, None),
                                                          "Mid": rb.Rt60BandMetrics(9.2129, 1.5, None),
                                                          "High": rb.Rt60BandMetrics(None, None, None)})
    assert rb.summarise_rt60_bands_results_text([res], False, False) == "[left]\nBand  T30_RT60(s)\nLow  86.019\nMid  9.213\nHigh  NA\n"
    assert "Band  T30_RT60(s)  T20_RT60(s)\nLow  86.019  NA\nMid  9.213  1.500" in rb.summarise_rt60_bands_results_text([res], True, False)
    w = wf.ChannelWaterfallResult("left", SR, 9287, 1142557, np.zeros(18, np.float32), np.zeros(1705, np.float32),
                                  np.zeros((18, 1705), np.float32))
    assert wf.summarise_waterfall_results_text([w]) == "[left] start_sample=9287  dur=23.803s  slices=18  f_bins=1705"


def test_cli_surface_matches_the_reference_arguments():
    from audio_analysis_b200 import cli
    a = cli.parse_arguments(["decay", "--input", "x.wav"])
    assert a.compute_edt is True and a.trim_to_peak is True and a.no_show is False and a.output_basename is None
    assert cli.parse_arguments(["decay", "--input", "x.wav", "--no-compute_edt", "--no-trim_to_peak"]).compute_edt is False
    a = cli.parse_arguments(["rt60bands", "--input", "x.wav", "--band_mode", "third", "--mono", "--include_t20"])
    assert a.band_mode == "third" and a.legend_values is None and a.trim_to_peak is True and a.include_t20
    assert a.transition_width_octaves == pytest.approx(1.0 / 6.0) and a.f_min_hz == 31.5 and a.f_max_hz == 16000.0
    a = cli.parse_arguments(["waterfall", "--input", "x.wav", "--no_hann_window", "--slice_mode", "uniform_time"])
    assert a.n_fft == 4096 and a.hop_length == 512 and a.num_slices == 18 and a.no_hann_window and a.style == "3d"
    a = cli.parse_arguments(["bundle", "--input", "run"])
    assert a.bundle_root == "run" and a.reports_subdir == "reports"
    a = cli.parse_arguments(["spectrogram", "--input", "x.wav", "--duration", "1.5", "--no_hann_window"])
    assert a.n_fft == 4096 and a.hop_length == 512 and a.floor_db == -120.0 and a.dynamic_range_db == 90.0
    assert a.analysis_duration_seconds == 1.5 and a.no_hann_window and a.trim_to_peak is True
    a = cli.parse_arguments(["modalcloud", "--input", "x.wav", "--metric", "edt", "--no-show_median_curve"])
    assert a.n_fft == 8192 and a.hop_length == 512 and a.metric == "edt" and a.log_bins_per_octave == 24
    assert a.min_fit_points == 10 and a.min_peak_db_above_floor == 20.0 and a.show_median_curve is False
    a = cli.parse_arguments(["deconvolve", "--recorded_wav_file_path", "r.wav", "--sweep_wav_file_path", "s.wav"])
    assert a.output_ir_wav_file_path is None and a.regularization_relative == 1e-10 and a.normalise_peak is True
    assert a.target_peak == 0.95 and a.remove_dc is True and a.output_length_mode == "recorded"
    a = cli.parse_arguments(["deconvolve", "--recorded_wav_file_path", "r.wav", "--sweep_wav_file_path", "s.wav",
                             "--no-normalise_peak", "--no-remove_dc", "--output_length_mode", "full_fft"])
    assert a.normalise_peak is False and a.remove_dc is False and a.output_length_mode == "full_fft"
    with pytest.raises(SystemExit):
        cli.parse_arguments(["zplane", "--input", "x.wav"])  # outside the accelerated path
    out = subprocess.run([sys.executable, "-m", "analyse.cli", "--help"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "rt60bands" in out.stdout and "waterfall" in out.stdout


def test_analyse_package_aliases_the_reference_module_names():
    import analyse.decay
    import analyse.rt60bands
    import analyse.waterfall
    import analyse.spectrogram
    import analyse.modalcloud
    import analyse.bundle
    import analyse.io
    import audio_analysis_b200.decay
    assert analyse.decay is audio_analysis_b200.decay
    for name in ("compute_schroeder_edc_db", "fit_decay_slope_over_db_range", "analyse_decay_for_channel",
                 "analyse_decay_from_wav_file", "plot_decay_from_wav_file", "summarise_decay_results_text"):
        assert hasattr(analyse.decay, name)
    for name in ("analyse_rt60_bands_for_channel", "analyse_rt60_bands_from_wav_file", "plot_rt60_bands_from_wav_file",
                 "summarise_rt60_bands_results_text"):
        assert hasattr(analyse.rt60bands, name)
    for name in ("analyse_waterfall_for_channel", "analyse_waterfall_from_wav_file", "plot_waterfall_from_wav_file",
                 "summarise_waterfall_results_text"):
        assert hasattr(analyse.waterfall, name)
    for name in ("SpectrogramAnalysisSettings", "ChannelSpectrogramResult", "analyse_spectrogram_for_channel",
                 "analyse_spectrogram_from_wav_file", "plot_spectrogram_from_wav_file",
                 "summarise_spectrogram_results_text"):
        assert hasattr(analyse.spectrogram, name)
    for name in ("ModalCloudAnalysisSettings", "ModalPoint", "ChannelModalCloudResult", "analyse_modal_cloud_for_channel",
                 "analyse_modal_cloud_from_wav_file", "plot_modal_cloud_from_wav_file",
                 "summarise_modal_cloud_results_text"):
        assert hasattr(analyse.modalcloud, name)
    assert hasattr(analyse.bundle, "run_bundle_report") and hasattr(analyse, "load_wav_file")
    import analyse.deconvolve
    import audio_analysis_b200.deconvolve
    assert analyse.deconvolve is audio_analysis_b200.deconvolve
    for name in ("DeconvolveSettings", "DeconvolvedImpulseResponse", "deconvolve_impulse_response",
                 "deconvolve_from_wav_files", "default_output_ir_path"):
        assert hasattr(analyse.deconvolve, name)
    assert str(analyse.deconvolve.default_output_ir_path("a/b/capture.wav")).endswith("a/b/capture_ir.wav")


def test_butterworth_design_matches_scipy():
    """butterworth.py designs the sections of the opt-in filterbank without scipy.signal; same response, same
    sosfiltfilt result, same sosfilt_zi and padlen as SciPy's own design."""
    from scipy import signal
    from audio_analysis_b200 import butterworth as bw
    rng = np.random.default_rng(0)
    x = rng.standard_normal(6000) * np.exp(-np.arange(6000) / 1500.0)
    for kind, lo, hi, order in (("bandpass", 890.9, 1122.5, 4), ("bandpass", 35.0, 44.1, 4), ("lowpass", None, 250.0, 4),
                                ("highpass", 4000.0, None, 4), ("lowpass", None, 250.0, 3), ("bandpass", 11313.0, 14254.0, 2)):
        sos = bw.butter_sos(order, kind, lo, hi, 48000)
        wn = [lo, hi] if kind == "bandpass" else (hi if kind == "lowpass" else lo)
        ref = signal.butter(order, wn, btype=kind, fs=48000, output="sos")
        assert sos.shape == ref.shape
        _, h = signal.sosfreqz(sos, 1024, fs=48000)
        _, hr = signal.sosfreqz(ref, 1024, fs=48000)
        assert np.max(np.abs(h - hr)) <= 1e-9
        y, yr = signal.sosfiltfilt(sos, x), signal.sosfiltfilt(ref, x)
        assert np.max(np.abs(y - yr)) <= 1e-9 * np.max(np.abs(yr))
        assert np.allclose(bw.sosfilt_zi(sos), signal.sosfilt_zi(sos), rtol=0, atol=1e-12)
        ntaps = 2 * ref.shape[0] + 1 - min(int((ref[:, 2] == 0).sum()), int((ref[:, 5] == 0).sum()))
        assert bw.filtfilt_padlen(sos) == 3 * ntaps
    sos, zi, edges = bw.band_sections([("lowpass", None, 250.0), ("bandpass", 707.0, 1414.0), ("highpass", 4000.0, None)], 4, 48000)
    assert sos.shape == (3, 4, 6) and zi.shape == (3, 4, 2) and list(edges) == [15, 27, 15]
    assert np.array_equal(sos[0, 2:], np.tile([1.0, 0, 0, 1.0, 0, 0], (2, 1)))


def test_bundle_wire_format_staging(tmp_path):
    """wire.py: the recorder's PCM16 WAV layout (recorder.hpp:55-90) is parsed and its payloads land, untouched, in one
    staging buffer; anything else is left to the scipy loader."""
    from scipy.io import wavfile
    from audio_analysis_b200 import wire
    rng = np.random.default_rng(3)
    taps = [rng.integers(-32768, 32767, size=(1000, 2), dtype=np.int16) for _ in range(3)]
    paths = []
    for i, t in enumerate(taps):
        p = tmp_path / f"tap{i}.wav"
        wire.write_pcm16_wav(p, t, 48000)
        paths.append(p)
        rate, back = wavfile.read(str(p))          # scipy reads what the writer wrote
        assert rate == 48000 and np.array_equal(back, t) and p.stat().st_size == 44 + t.nbytes
    info = wire.probe_pcm16(paths[0])
    assert (info.sample_rate_hz, info.channels, info.frames, info.data_offset) == (48000, 2, 1000, 44)
    st = wire.stage_pcm16(paths, pin=False)
    assert st is not None and (st.sample_rate_hz, st.channels, st.frames) == (48000, 2, 1000)
    assert np.array_equal(st.pcm.numpy().reshape(3, 1000, 2), np.stack(taps))
    # a float32 WAV, a file of another length, a missing file: no staging
    fl = tmp_path / "float.wav"
    wavfile.write(str(fl), 48000, np.zeros((1000, 2), np.float32))
    assert wire.probe_pcm16(fl) is None and wire.stage_pcm16([paths[0], fl], pin=False) is None
    short = tmp_path / "short.wav"
    wire.write_pcm16_wav(short, taps[0][:500], 48000)
    assert wire.stage_pcm16([paths[0], short], pin=False) is None
    assert wire.probe_pcm16(tmp_path / "nope.wav") is None


def test_waterfall_smoothing_tables_reproduce_numpy_interp():
    """waterfall.smoothing_tables: the host half of the device log-frequency smoothing.  Doing the kernel's arithmetic
    with those tables in numpy gives the oracle's (= the reference's, waterfall.py:140-185) smoothed curve bit for bit."""
    import oracle
    from audio_analysis_b200 import waterfall as wf
    s = wf.WaterfallAnalysisSettings(smoothing_log_bins=9, f_min_hz=30.0, f_max_hz=18000.0)
    bin_lo, freq = wf.frequency_selection(s, 48000)
    k0, f_in, grid, grid_seg, bin_seg = wf.smoothing_tables(freq, s)
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal(freq.size) * 10.0 - 40.0).astype(np.float32)

    def interp(x, seg, xp, fp):
        out = np.empty(x.size)
        for i in range(x.size):
            j = int(seg[i])
            if j == -1:
                out[i] = fp[0]
            elif j == -2 or j >= xp.size - 1:
                out[i] = fp[-1]
            else:
                out[i] = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]) * (x[i] - xp[j]) + fp[j]
        return out

    cur = raw[k0:k0 + f_in.size].astype(np.float64)
    on_grid = interp(grid, grid_seg, f_in, cur).astype(np.float32).astype(np.float64)
    w = 9
    back, fwd = w // 2, (w - 1) // 2
    smooth = np.array([sum(on_grid[m] * (1.0 / w) for m in range(max(0, g - back), min(grid.size - 1, g + fwd) + 1))
                       for g in range(grid.size)]).astype(np.float32).astype(np.float64)
    got = raw.copy()
    got[k0:k0 + f_in.size] = interp(f_in, bin_seg, grid, smooth).astype(np.float32)
    want = oracle.ref_port._smooth_logf(freq, raw, 30.0, 18000.0, 9, s.log_bins_per_octave)
    assert np.allclose(got, want, rtol=0, atol=2e-6) and np.array_equal(got[:k0], raw[:k0])


def test_wav_io_rules(tmp_path):
    from scipy.io import wavfile
    from audio_analysis_b200 import io
    pcm = load_golden("smooth_48000")["pcm"][:1000]
    path = tmp_path / "s.wav"
    wavfile.write(str(path), SR, np.stack([pcm, pcm // 2], axis=1))
    a = io.load_wav_file(path, expected_channel_mode="mono_or_stereo", allow_mono_and_upmix_to_stereo=False)
    assert a.samples.dtype == np.float32 and a.samples.shape == (1000, 2)
    assert np.array_equal(a.samples[:, 0], pcm.astype(np.float32) / 32768.0)
    names = [n for n, _ in io.get_analysis_channels(a)]
    assert names == ["left", "right"]
    (name, mono), = io.get_analysis_channels(a, use_mono_downmix_for_stereo=True)
    assert name == "mono" and np.array_equal(mono, 0.5 * (a.samples[:, 0] + a.samples[:, 1]))
    wavfile.write(str(tmp_path / "r.wav"), 44100, pcm)
    with pytest.raises(ValueError, match="Expected sample rate 48000 Hz"):
        io.load_wav_file(tmp_path / "r.wav", expected_channel_mode="mono_or_stereo")
    assert io.convert_wav_samples_to_float32(np.array([2.0, -3.0], np.float64)).tolist() == [1.0, -1.0]
    with pytest.raises(ValueError, match="Unsupported integer PCM dtype"):
        io.convert_wav_samples_to_float32(np.zeros(4, np.int8))


def test_shard_ranges_cover_everything():
    from audio_analysis_b200.distributed import shard_range
    for n, w in ((65536, 8), (10, 4), (7, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from audio_analysis_b200 import distributed as d
    d.init_from_env("gloo")
    n_items = 11
    b, e = d.shard_range(n_items, rank, world)
    idx = torch.arange(b, e, dtype=torch.float64)
    tables = {"band_rt60": idx[:, None, None].expand(-1, 26, 3).contiguous() + 0.5,
              "start": torch.arange(b, e, dtype=torch.int32)}
    got = d.gather_tables(tables, n_items)
    if rank == 0:
        ok = (got["start"].tolist() == list(range(n_items))
              and torch.equal(got["band_rt60"][:, 0, 0], torch.arange(n_items, dtype=torch.float64) + 0.5))
        q.put(ok)
    else:
        q.put(got is None)
    dist.barrier()
    dist.destroy_process_group()


def test_host_binding_helpers_without_a_device():
    """distributed.bind_host_to_gpu: cpulist parsing, and no change of affinity when the topology cannot be read."""
    import os
    from audio_analysis_b200 import distributed
    assert distributed._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert distributed._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    import torch
    if not torch.cuda.is_available():
        assert distributed.bind_host_to_gpu("cuda:0") is None   # no device: nothing is read, nothing is changed
    assert os.sched_getaffinity(0) == before


def test_metric_tables_gather_over_gloo_world_size_2():
    """The multi-GPU path's only collective, on CPU with gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(results) and all(p.exitcode == 0 for p in procs)


def _bundle_shard_worker(rank, world, port, root, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from audio_analysis_b200 import bundle, distributed as d
    d.init_from_env("gloo")
    seen = []
    # the analysis itself needs a GPU; here only the sharding / index logic of run_bundle_report is under test
    bundle.run_reports_batch = lambda wavs, bases, settings=None: seen.extend(p.stem for p in wavs) or []
    index = bundle.run_bundle_report(root)
    q.put((rank, seen, index.exists() if rank == 0 else None))
    dist.barrier()
    dist.destroy_process_group()


def test_bundle_taps_are_sharded_over_ranks_gloo_world_size_2(tmp_path):
    """run_bundle_report under torch.distributed: contiguous blocks of taps per rank, every tap exactly once, the
    index written by rank 0 after a barrier (the multi-GPU form of analyse/bundle.py:35-74)."""
    import json
    import torch.multiprocessing as mp
    root = tmp_path / "run"
    (root / "taps").mkdir(parents=True)
    taps = [f"tap_{i:02d}" for i in range(7)]
    (root / "meta.json").write_text(json.dumps({"sample_rate_hz": 48000, "length_samples": 10, "taps": taps}))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bundle_shard_worker, args=(r, 2, port, str(root), q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert results[0][1] == taps[:4] and results[1][1] == taps[4:] and results[0][2] is True
    text = (root / "reports" / "bundle_report.md").read_text()
    assert all(f"- [{t}](reports/{t}/{t}_report.md)" in text for t in taps)


def _gatherer_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from audio_analysis_b200 import distributed as d
    d.init_from_env("gloo")
    per = 5
    base = rank * per
    tables = {"band_rt60": (torch.arange(per * 26 * 3, dtype=torch.float64).view(per, 26, 3) + 1000.0 * rank),
              "fits": torch.full((per, 3, 8), float(rank), dtype=torch.float64),
              "early10": torch.arange(base, base + per, dtype=torch.float64) * 0.5,
              "start": torch.arange(base, base + per, dtype=torch.int32)}
    tables["band_rt60"][0, 0, 0] = float("nan")
    g = d.TableGatherer(tables, per)
    ok = g.record % 16 == 0 and g.record >= 26 * 3 * 8 + 3 * 8 * 8 + 8 + 4
    for step in range(3):           # the same buffers every step
        tables["early10"] += 1.0
        got = g.gather(tables)
        if rank == 0:
            ok = ok and got["start"].tolist() == list(range(world * per))
            ok = ok and torch.equal(got["early10"], torch.arange(world * per, dtype=torch.float64) * 0.5 + (step + 1))
            ok = ok and bool(torch.isnan(got["band_rt60"][0, 0, 0])) and bool(torch.isnan(got["band_rt60"][per, 0, 0]))
            ok = ok and float(got["band_rt60"][per, 1, 2]) == 1000.0 + 5.0 and float(got["fits"][per + 1, 2, 7]) == 1.0
        else:
            ok = ok and got is None
    q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_packed_table_gatherer_reuses_its_buffers_gloo_world_size_2():
    """distributed.TableGatherer: one packed record per row, one collective per step, buffers allocated once."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gatherer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(results) and all(p.exitcode == 0 for p in procs)


def test_synthetic_generator_is_deterministic_and_has_leading_silence():
    from audio_analysis_b200.synth import ir_parameters, synth_ir_numpy
    a, b = synth_ir_numpy(3, 48000), synth_ir_numpy(3, 48000)
    assert np.array_equal(a, b) and a.dtype == np.float32 and np.max(np.abs(a)) == pytest.approx(0.95, rel=1e-6)
    rt, onset = ir_parameters(3, 48000, SR)
    assert 0.05 <= rt <= 2.0 and onset >= 1
    assert np.max(np.abs(a[: onset - 1])) < 1e-3  # only the -90 dB floor before the onset
