/*
 * audio_analysis_b200 - C ABI of the B200-native IR decay / spectral-decay path.
 *
 * This is the drop-in boundary for the hot path of kianmcevoy/audio_analysis
 * (`analyse.decay`, `analyse.rt60bands`, `analyse.waterfall`, and through them
 * `analyse.bundle`).  The reference has no FFI of its own: the
 * functions below are what a ctypes binding inside the reference's
 * `*_for_channel` functions would call (INTEGRATION.md shows the stubs).  Each
 * entry point cites the reference lines it replaces, relative to the reference
 * tree.
 *
 * Conventions
 *  - plain C types only; every pointer named d_* is a DEVICE pointer owned by
 *    the caller (a torch tensor in the shipped Python host layer), h_* is HOST.
 *  - the library keeps no global state except the FFT plan objects the caller
 *    creates and destroys; workspaces are caller-provided.
 *  - every call takes the CUDA stream to enqueue on (a cudaStream_t passed as
 *    void*); calls are asynchronous with respect to the host and re-entrant per
 *    stream.
 *  - return value: AA_OK (0) or a negative aa_status; no exceptions cross the
 *    ABI.  aa_last_error() returns a thread-local message for the last failure.
 *  - batched layout: rows of float32 samples, `row_stride` elements apart, row
 *    base addresses 16-byte aligned (row_stride % 4 == 0).  A row is one
 *    analysed channel (the reference's 1-D `samples`).
 *  - "None" in the reference's result records is NaN in the output tables.
 */
#ifndef AUDIO_ANALYSIS_B200_H
#define AUDIO_ANALYSIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AA_ABI_VERSION 1

typedef enum aa_status {
    AA_OK = 0,
    AA_ERR_INVALID_ARGUMENT = -1,
    AA_ERR_CUDA = -2,
    AA_ERR_TOO_LONG = -3,       /* row longer than AA_MAX_ROW_SAMPLES */
    AA_ERR_WORKSPACE = -4,      /* workspace too small */
    AA_ERR_UNSUPPORTED = -5,
    AA_ERR_NO_DEVICE = -6
} aa_status;

#define AA_MAX_ROW_SAMPLES (1 << 24) /* float32 arange is exact up to 2^24 (decay.py:169) */
#define AA_MAX_FITS 3
#define AA_FIT_COLS 8  /* t_start, t_end, slope_db_per_s, intercept_db, r2, rt60_s, n_points, valid */
#define AA_MAX_BANDS 64
#define AA_MAX_SLICES 64
#define AA_SOS_MAX_SECTIONS 8

/* per-row status written by the batched kernels */
#define AA_ROW_OK 0
#define AA_ROW_TOO_SHORT 1 /* the reference raises ValueError (decay.py:146, waterfall.py:375) */

/* ---- decay: DecayAnalysisSettings (analyse/decay.py:44-72) -------------- */
typedef struct aa_decay_params {
    int32_t sample_rate_hz;
    int32_t trim_to_peak;            /* decay.py:135 */
    int32_t ignore_leading_samples;  /* int(round(ignore_leading_seconds*sr)), decay.py:141 */
    int32_t smoothing_window;        /* edc_smoothing_window_samples; <=1 = off, decay.py:161 */
    double edc_floor_db;             /* decay.py:167 */
    double edc_epsilon;              /* decay.py:154 */
    double fit_lower_limit_db;       /* decay.py:219 */
    int32_t n_fits;                  /* 0..AA_MAX_FITS line fits, decay.py:202 */
    int32_t want_early10;            /* early_decay_10db_time_seconds, decay.py:280-286 */
    double fit_hi_db[AA_MAX_FITS];   /* range_db[0] */
    double fit_lo_db[AA_MAX_FITS];   /* range_db[1] (before the lower-limit clamp) */
} aa_decay_params;

/* ---- rt60bands: BandDefinition (analyse/rt60bands.py:83-89) ------------- */
#define AA_BAND_LOWPASS 0
#define AA_BAND_BANDPASS 1
#define AA_BAND_HIGHPASS 2
typedef struct aa_band {
    int32_t kind;
    int32_t reserved;
    double low_edge_hz;  /* unused for lowpass */
    double high_edge_hz; /* unused for highpass */
} aa_band;

/* ---- waterfall: WaterfallAnalysisSettings (analyse/waterfall.py:43-80) -- */
#define AA_DBREF_GLOBAL_MAX 0
#define AA_DBREF_SLICE_MAX 1
#define AA_DBREF_NONE 2 /* raw magnitude dB, no normalisation (spectrogram-style) */
typedef struct aa_csd_params {
    int32_t n_fft;          /* power of two, 64..8192 */
    int32_t hop_length;
    int32_t use_hann_window;
    int32_t bin_lo;         /* first kept rfft bin (frequency crop, waterfall.py:304) */
    int32_t n_bins;         /* number of kept bins F */
    int32_t db_reference;   /* AA_DBREF_* (waterfall.py:328-336) */
    double floor_db;        /* magnitude floor before dB (waterfall.py:218) */
    double dynamic_range_db;/* clip range, already max(10, .) (waterfall.py:338) */
} aa_csd_params;

/* ---- library ------------------------------------------------------------ */
int aa_abi_version(void);
const char* aa_last_error(void);
/* number of visible CUDA devices, or a negative aa_status */
int aa_device_count(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
long long aa_kernel_launch_count(void);

/*
 * Optional per-stage device timing (used by bench.py for the roofline objects).  While enabled, the library
 * brackets its kernels with CUDA events on the caller's stream under these names:
 *   "edc" (fused EDC + fits kernel), "fft_cols_band" / "fft_cols" (first FFT pass, band-masked source / other),
 *   "fft_rows" (second FFT pass), "untangle", "band_fit", "band_edc", "csd".
 * aa_profile_enable(1) starts a fresh measurement, aa_profile_collect() waits for the recorded events and sums
 * them per name, aa_profile_ms / aa_profile_scopes read the sums.  Not for concurrent use from several threads.
 */
void aa_profile_enable(int on);
int aa_profile_collect(void);
double aa_profile_ms(const char* name);
long long aa_profile_scopes(const char* name);

/*
 * aa_peak_index_batch - first index of max|x| per row.
 * Replaces np.argmax(np.abs(x)) at analyse/decay.py:136, rt60bands.py:335,
 * waterfall.py:359.  d_peak_out[b] = index (0 for an empty row).
 */
int aa_peak_index_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t n_rows,
                        int32_t* d_peak_out, void* stream);

/*
 * aa_decay_batch - fused Schroeder EDC + crossings + line fits for a batch of rows.
 * Replaces compute_schroeder_edc_db (analyse/decay.py:115-170),
 * _interpolated_crossing_time_seconds (:173-199), fit_decay_slope_over_db_range
 * (:202-260) and the body of analyse_decay_for_channel (:268-329).
 *
 *  d_x, row_stride, d_lengths[n_rows] : input rows (d_lengths may be NULL: all rows `uniform_length`)
 *  d_start_in   : NULL, or per-row analysis start index to use INSTEAD of the
 *                 peak search (rt60bands.py:342/393 applies one shared start to
 *                 every band signal); ignore_leading_samples is then not added.
 *  d_edc_db_out : NULL (metrics only) or rows of `edc_stride` floats; entry j of
 *                 row b is the curve value at ABSOLUTE sample j, valid for
 *                 j >= d_start_out[b] (the reference's edc_db[i] is entry start+i).
 *  d_start_out[n_rows]                : analysis_start_sample_index
 *  d_fits_out[n_rows][AA_MAX_FITS][AA_FIT_COLS] (double), row f = p->fit_*[f]
 *  d_early10_out[n_rows] (double, NaN = None)
 *  d_status_out[n_rows]               : AA_ROW_*
 */
int aa_decay_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t uniform_length,
                   int32_t n_rows, const aa_decay_params* p, const int32_t* d_start_in, float* d_edc_db_out,
                   int64_t edc_stride, int32_t* d_start_out, double* d_fits_out, double* d_early10_out,
                   int32_t* d_status_out, void* stream);

/*
 * aa_curve_fits_batch - crossings + line fits on caller-supplied dB curves (monotone or not).
 * Replaces _interpolated_crossing_time_seconds and fit_decay_slope_over_db_range
 * (analyse/decay.py:173-260) when the curve comes from the caller or was smoothed; the time axis is
 * time_seconds[i] = float32(i)/sample_rate (decay.py:169).
 *  d_curve rows  : float32 dB values, `stride` apart; row b holds the curve in
 *                  [d_offsets[b], length_b) (d_offsets NULL = 0; length from d_lengths or uniform_length)
 *  p             : n_fits / fit_* / fit_lower_limit_db / sample_rate_hz / want_early10 are used
 *  outputs       : as aa_decay_batch
 */
int aa_curve_fits_batch(const float* d_curve, int64_t stride, const int32_t* d_lengths, int32_t uniform_length,
                        const int32_t* d_offsets, int32_t n_rows, const aa_decay_params* p, double* d_fits_out,
                        double* d_early10_out, void* stream);

/*
 * aa_box_smooth_db_batch - np.convolve(edc_db, ones(w)/w, "same") then max(., floor_db) -> float32
 * (analyse/decay.py:161-167).  Same row addressing as aa_curve_fits_batch; d_in != d_out.
 */
int aa_box_smooth_db_batch(const float* d_in, float* d_out, int64_t stride, const int32_t* d_lengths,
                           int32_t uniform_length, const int32_t* d_offsets, int32_t n_rows, int32_t window,
                           double floor_db, void* stream);

/*
 * Band filterbank plan for rows of exactly `n` samples: the length-n circular
 * real FFT pair behind _apply_fft_mask (analyse/rt60bands.py:170-175).
 */
typedef struct aa_bands_plan aa_bands_plan;
int aa_bands_plan_create(int32_t n, aa_bands_plan** out);
void aa_bands_plan_destroy(aa_bands_plan* plan);
/* 0 = direct mixed radix, 1 = Bluestein on n/2, 2 = Bluestein on n (odd n) */
int aa_bands_plan_kind(const aa_bands_plan* plan);
/*
 * Band decimation with certified refinement (rows of 288,000 / 192,000 samples, metrics only).
 * The samples y[start + D j] of a band signal are the length n / D inverse transform of the band's spectrum folded
 * into [0, n / 2D].  A band whose mask support lies inside one Nyquist zone [m n/2D, (m+1) n/2D] of the decimated rate
 * folds without overlap, so those samples ARE samples of the reference's band signal (rt60bands.py:170-175); its
 * Schroeder integral (decay.py:150-158) is taken on that grid with trapezoid end terms.  That sum is an estimate of the
 * reference's full-rate sum, so every (row, band) unit is CERTIFIED before its fits are kept: the reference's level
 * crossings (decay.py:173-199) lie in a window around the decimated ones (modelled quadrature error x cert_safety),
 * and moving a fit's end points across those windows must change its slope by less than cert_tol (estimated from the
 * distance of the curve from the fitted line at the end points; window <= cert_window x span, >= cert_min_span
 * decimated points per 30 dB fit).  The line fit itself (decay.py:224-249) weights the first and last decimated
 * sample by the part of their cell inside the reference's sample range.  Units that fail are redone on the next level: D <= max_decimation (default 16, or the environment
 * variable AA_BANDS_DECIMATION) -> D <= 8 -> D <= 4 -> full rate, whose results are the reference's estimator itself.
 * A call with a fit shorter than 28 dB or starting above -2.5 dB (T20, EDT) starts at D <= 4.
 * D in {2, 3, 4, 6, 8, 12, 16, 24}; max_decimation 1 = everything at full rate.
 * aa_bands_classify reports the first-level factor of each band (host only, no device work).
 * With refinement the call synchronises `stream` with the host once per level and chunk of rows (it reads the
 * lengths of the unit lists); one plan must not be used by two host threads at the same time.
 */
typedef struct aa_bands_options {
    int32_t max_decimation;  /* 0 = process default (aa_bands_set_decimation) */
    int32_t refine;          /* 0 = default (on), 1 = on, -1 = off: keep every first-level result (round-1 behaviour) */
    double cert_safety;      /* 0 = default 9 */
    double cert_window;      /* 0 = default 0.015 */
    double cert_tol;         /* 0 = default 3e-3 */
    int32_t cert_min_span;   /* 0 = default 240 */
    int32_t reserved;
} aa_bands_options;
typedef struct aa_bands_stats {
    int32_t n_levels;
    int32_t level_limit[4];  /* decimation limit of each level (last = 1 when refinement is on) */
    int64_t level_units[4];  /* (row, band) units processed on each level */
    int32_t n_chunks;        /* chunks of rows the workspace forced */
    int32_t rows_per_chunk;
} aa_bands_stats;
int aa_bands_get_decimation(void);
void aa_bands_set_decimation(int max_decimation);
int aa_bands_classify(const aa_bands_plan* plan, const aa_band* bands, int32_t n_bands, double transition_width_octaves,
                      int32_t sample_rate_hz, const aa_decay_params* fits /* may be NULL */, int32_t* out_decimation);
/* bytes of workspace needed to process `rows_per_chunk` rows x `n_bands` bands at once */
size_t aa_bands_workspace_bytes(const aa_bands_plan* plan, int32_t rows_per_chunk, int32_t n_bands);

/*
 * aa_rt60bands_batch - band-limited RT60 table for a batch of equal-length rows.
 * Replaces the band loop of analyse_rt60_bands_for_channel
 * (analyse/rt60bands.py:362-406): raised-cosine masks (:116-167), circular FFT
 * filtering (:170-175), the shared start index (:333-342, :393) and
 * _compute_band_metrics_from_samples (:272-321).
 *
 *  d_start_idx[n_rows] : shared analysis start per row (peak + ignore), e.g.
 *                        d_start_out of aa_decay_batch or aa_peak_index_batch.
 *  bands[n_bands]      : HOST array, from the band table (rt60bands.py:183-264)
 *  p                   : decay policy; p->n_fits fits per band, in p->fit_* order
 *  d_metrics_out[n_rows][n_bands][AA_MAX_FITS] (double): rt60 seconds, NaN = None
 *  d_band_signal_out   : NULL, or [n_rows][n_bands][n] float band signals (debug/parity only)
 *  d_workspace / workspace_bytes : scratch; rows are processed in chunks that fit
 */
int aa_rt60bands_batch(const aa_bands_plan* plan, const float* d_x, int64_t row_stride, int32_t n_rows,
                       const int32_t* d_start_idx, const aa_band* bands, int32_t n_bands,
                       double transition_width_octaves, const aa_decay_params* p, double* d_metrics_out,
                       float* d_band_signal_out, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * aa_rt60bands_sos_batch - OPT-IN band filterbank: Butterworth second-order sections, zero-phase forward-backward
 * filtering as a chunked parallel linear-recurrence scan with float64 states and carries (csrc/sos.cu), then the same
 * band EDC + fits as aa_rt60bands_batch.  This is NOT the reference's band filter (a circular FFT-domain mask,
 * analyse/rt60bands.py:116-175); it implements scipy.signal.sosfiltfilt (odd extension of 3 * ntaps samples, steady-
 * state initial conditions) and is validated against it.  The host designs the sections.
 *  h_sos[n_bands][n_sections][6] : b0 b1 b2 1 a1 a2 per section (HOST, float64); n_sections <= AA_SOS_MAX_SECTIONS
 *  h_zi[n_bands][n_sections][2]  : scipy.signal.sosfilt_zi of each band (HOST)
 *  h_edge[n_bands]               : odd extension at each end = sosfiltfilt's default padlen of the band (HOST)
 *  other arguments as aa_rt60bands_batch; aa_sos_workspace_bytes sizes the workspace.
 */
size_t aa_sos_workspace_bytes(int32_t n, int32_t n_sections, int32_t rows_per_chunk, int32_t n_bands);
int aa_rt60bands_sos_batch(const float* d_x, int64_t row_stride, int32_t n, int32_t n_rows, const int32_t* d_start_idx,
                           const double* h_sos, const double* h_zi, const int32_t* h_edge, int32_t n_bands,
                           int32_t n_sections, const aa_decay_params* p, double* d_metrics_out, float* d_band_signal_out,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * aa_csd_batch - waterfall / cumulative-spectral-decay slices for a batch of rows.
 * Replaces _compute_stft_magnitude_db (analyse/waterfall.py:188-230) for the
 * selected frames only, and _build_rel_db_slices (:289-341) without the optional
 * log-frequency smoothing.
 *
 *  d_start_idx[n_rows]                 : analysis start (trim/ignore already applied)
 *  d_frame_idx[n_rows][max_slices]     : STFT frame index of each slice (waterfall.py:233-286)
 *  d_n_slices[n_rows]                  : slices used in each row (<= max_slices)
 *  d_out[n_rows][max_slices][n_bins]   : relative dB in [-dyn, 0] (or raw dB for AA_DBREF_NONE)
 * Frame s of row b covers samples start + frame*hop .. + n_fft (must lie inside the row).
 */
int aa_csd_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t uniform_length,
                 int32_t n_rows, const int32_t* d_start_idx, const int32_t* d_frame_idx,
                 const int32_t* d_n_slices, int32_t max_slices, const aa_csd_params* p, float* d_out,
                 void* stream);

/*
 * aa_csd_plan_batch - which samples and frames of every row the waterfall analyses: the trim / ignore / duration rules
 * of analyse_waterfall_for_channel (analyse/waterfall.py:358-372), the frame count (:208) and the row's line of the
 * slice-selection table (the host tabulates _select_slice_frame_indices, :233-286, per frame count; the selection
 * depends on nothing else).  Feeds aa_csd_batch without a host round trip.
 *  d_peak_idx[n_rows]  : NULL (trim_to_peak off: start at 0) or the peak index per row
 *  analysis_duration_samples : -1 = whole remainder
 *  d_slice_table[max_frames + 1][table_width], d_slice_counts[max_frames + 1] : selected frames / their number
 *  outputs: d_start_out[n_rows], d_span_out[n_rows] (analysed samples), d_frame_idx_out[n_rows][table_width],
 *           d_n_slices_out[n_rows]
 */
int aa_csd_plan_batch(const int32_t* d_lengths, int32_t uniform_length, int32_t n_rows, const int32_t* d_peak_idx,
                      int32_t ignore_leading_samples, int32_t analysis_duration_samples, int32_t n_fft, int32_t hop_length,
                      const int32_t* d_slice_table, const int32_t* d_slice_counts, int32_t max_frames, int32_t table_width,
                      int32_t* d_start_out, int32_t* d_span_out, int32_t* d_frame_idx_out, int32_t* d_n_slices_out,
                      void* stream);

/*
 * aa_csd_smooth_batch - the optional log-frequency smoothing of the waterfall slices, then reference level and clip.
 * Replaces _smooth_mag_db_log_frequency / _moving_average_1d (analyse/waterfall.py:132-185) and the tail of
 * _build_rel_db_slices (:312-339) when smoothing_log_bins > 1.
 *  d_slices[n_rows][max_slices][n_bins] : RAW dB slices from aa_csd_batch(AA_DBREF_NONE) in, relative dB out
 *  bins [first_bin_inside, + n_bins_inside) lie inside [f_min, f_max] (waterfall.py:162); d_freq_inside holds their
 *  frequencies, d_grid_hz the uniform log2(f) grid (:171-176) - both float64, built by the host with the
 *  reference's expressions - and d_grid_segment[g] / d_bin_segment[k] the interval of the other axis each point
 *  falls in (np.interp's search; -1 = left of the first point, -2 = right of the last).
 *  db_reference: AA_DBREF_GLOBAL_MAX or AA_DBREF_SLICE_MAX (:328-336); the clip is [-dynamic_range_db, 0] (:338-339).
 */
int aa_csd_smooth_batch(float* d_slices, int32_t n_rows, const int32_t* d_n_slices, int32_t max_slices, int32_t n_bins,
                        int32_t first_bin_inside, int32_t n_bins_inside, const double* d_freq_inside,
                        const double* d_grid_hz, int32_t n_grid, const int32_t* d_grid_segment,
                        const int32_t* d_bin_segment, int32_t smoothing_log_bins, int32_t db_reference,
                        double dynamic_range_db, void* stream);

/*
 * aa_stft_db_batch - dense STFT magnitude in dB: every frame of the analysed segment of every row
 * (replaces _compute_stft_magnitude_db of analyse/spectrogram.py:107-160).
 *  d_start_idx[n_rows] : first analysed sample (peak + ignore), d_span[n_rows]: analysed samples from there
 *  frames of row r    : span >= n_fft ? 1 + (span - n_fft) / hop_length : 0, capped at frames_max
 *  d_out[n_rows][frames_max][n_fft/2 + 1]: 20 log10 max(|rfft(frame * window)|, 10^(floor_db/20)), float32,
 *                        frame-major (the reference stores (F, T): transpose on the host); entries of frames a
 *                        row does not have are left untouched
 *  n_fft in {1024, 2048, 4096, 8192}; other powers of two go through aa_csd_batch with AA_DBREF_NONE.
 */
int aa_stft_db_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t uniform_length,
                     int32_t n_rows, const int32_t* d_start_idx, const int32_t* d_span, int32_t n_fft, int32_t hop_length,
                     int32_t use_hann_window, double floor_db, int32_t frames_max, float* d_out, void* stream);

/*
 * aa_modal_cloud_batch - per log-frequency bin decay fits over the dense STFT
 * (replaces _aggregate_to_log_bins + the per-bin loop of analyse/modalcloud.py:176-207, :339-371).
 *  d_mag_db[n_rows][frames_max][n_freq] : aa_stft_db_batch output (frame-major dB)
 *  d_n_frames[n_rows]                   : frames each row has
 *  d_bin_ranges[n_bins][2]              : rfft-bin range [k0, k1) of every log bin (k0 == k1: empty -> no point)
 *  range_hi_db / range_lo_db            : the fit range of the chosen metric, e.g. (-5, -35)
 *  d_curve_ws[n_rows * n_bins * frames_max] float32 scratch (the per-bin decay curves in dB)
 *  d_out[n_rows][n_bins][2] (double)    : rt60 seconds, r^2; NaN = the bin yields no point
 */
int aa_modal_cloud_batch(const float* d_mag_db, int32_t n_rows, int32_t frames_max, int32_t n_freq,
                         const int32_t* d_n_frames, const int32_t* d_bin_ranges, int32_t n_bins, int32_t hop_length,
                         int32_t sample_rate_hz, double range_hi_db, double range_lo_db, double fit_lower_limit_db,
                         int32_t min_fit_points, double floor_db, double min_peak_db_above_floor, float* d_curve_ws,
                         double* d_out, void* stream);

/*
 * aa_pcm16_to_float_batch - int16 interleaved PCM -> planar float32 rows.
 * Replaces _convert_integer_pcm_to_float32 + clip + channel split/downmix
 * (analyse/io.py:46-64, :98-113, :66-95) for the bundle wire format written by
 * include/analysis/recorder.hpp:55-90.
 *  d_pcm[n_files][n_frames * n_channels] int16, rows `pcm_stride` elements apart
 *  mode 0: one output row per channel (row = file*n_channels + ch)
 *  mode 1: one row per file = 0.5*(L+R) (float32 arithmetic, io.py:90)
 */
int aa_pcm16_to_float_batch(const int16_t* d_pcm, int64_t pcm_stride, int32_t n_files, int32_t n_frames,
                            int32_t n_channels, int32_t mode, float* d_out, int64_t out_stride, void* stream);

/*
 * Deconvolution (analyse/deconvolve.py:124-193), SURVEY.md 8(f) N1.
 * H = Y conj(X) / (|X|^2 + eps) at n_fft = next_pow2(max(n_rec, n_sweep)).
 */
typedef struct aa_deconv_plan aa_deconv_plan;
int aa_deconv_plan_create(int32_t n_fft, aa_deconv_plan** out);
void aa_deconv_plan_destroy(aa_deconv_plan* plan);
size_t aa_deconv_workspace_bytes(const aa_deconv_plan* plan, int32_t n_rows);
/*
 *  d_sweep[n_sweep] : the excitation (mono, float32)
 *  d_rec rows       : recorded channels, n_rec samples each (clipped to [-1,1] as deconvolve.py:137)
 *  d_ir_out rows    : first n_out samples of the inverse transform (float32), DC removed if
 *                     remove_dc; peak normalisation is done over groups of `norm_group` consecutive
 *                     rows (the channels of one file, deconvolve.py:190) when target_peak > 0.
 */
int aa_deconvolve_batch(const aa_deconv_plan* plan, const float* d_sweep, int32_t n_sweep, const float* d_rec,
                        int64_t rec_stride, int32_t n_rec, int32_t n_rows, double regularization_relative,
                        int32_t remove_dc, double target_peak, int32_t norm_group, int32_t n_out,
                        float* d_ir_out, int64_t ir_stride, void* d_workspace, size_t workspace_bytes,
                        void* stream);

/*
 * aa_rt60bands_batch_ex - aa_rt60bands_batch with per-call options (NULL = defaults) and, optionally, the number of
 * units each refinement level processed (HOST struct, filled when the call returns).
 */
int aa_rt60bands_batch_ex(const aa_bands_plan* plan, const float* d_x, int64_t row_stride, int32_t n_rows,
                          const int32_t* d_start_idx, const aa_band* bands, int32_t n_bands,
                          double transition_width_octaves, const aa_decay_params* p, double* d_metrics_out,
                          float* d_band_signal_out, void* d_workspace, size_t workspace_bytes, void* stream,
                          const aa_bands_options* options, aa_bands_stats* stats_out);

#ifdef __cplusplus
}
#endif
#endif /* AUDIO_ANALYSIS_B200_H */
