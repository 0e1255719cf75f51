// Internal launchers shared between translation units.
#pragma once

#include "aa_common.cuh"

namespace aa {

int launch_edc(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int max_len, int n_rows,
               const aa_decay_params* p, const int* d_start_in, int start_div, int min_samples, float* d_edc_out,
               long long edc_stride, int* d_start_out, double* d_fits_out, double* d_rt60_out, double* d_early10_out,
               int* d_status_out, cudaStream_t stream);

int launch_peak(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int n_rows, int* d_out,
                cudaStream_t stream);

int launch_csd(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int n_rows,
               const int* d_start_idx, const int* d_frame_idx, const int* d_n_slices, int max_slices,
               const aa_csd_params* p, float* d_out, cudaStream_t stream);

int launch_stft_db(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int n_rows,
                   const int* d_start_idx, const int* d_span, int n_fft, int hop, int use_hann, double floor_db,
                   int frames_max, float* d_out, cudaStream_t stream);

int launch_pcm16(const int16_t* d_pcm, long long pcm_stride, int n_files, int n_frames, int n_channels, int mode,
                 float* d_out, long long out_stride, cudaStream_t stream);

}  // namespace aa

namespace aa {

int launch_curve_fits(const float* d_curve, long long stride, const int* d_lengths, int uniform_len,
                      const int* d_offsets, int n_rows, const aa_decay_params* p, double* d_fits_out,
                      double* d_rt60_out, double* d_early10_out, cudaStream_t stream);

int launch_box_smooth(const float* d_in, float* d_out, long long stride, const int* d_lengths, int uniform_len,
                      const int* d_offsets, int n_rows, int window, double floor_db, cudaStream_t stream);

}  // namespace aa

namespace aa {

// band EDC / fits from the energy partials written by the register-pass rows kernel (bandfit.cu)
int launch_band_fit(const float* d_y, long long y_stride, const float* d_part, long long part_stride, int n_groups,
                    int n_tiles, int tile_len, int n, long long n_units, const aa_decay_params* p,
                    const int* d_start_idx, int start_div, int min_samples, double* d_rt60_out, cudaStream_t stream,
                    int decim = 1, int full_n = 0, int aligned = 0);

}  // namespace aa

namespace aa {

// modal cloud (modal.cu): log-bin aggregation of the dense STFT + per-bin decay fits
int launch_modal_cloud(const float* d_mag_db, int n_rows, int frames_max, int n_freq, const int* d_n_frames,
                       const int* d_bin_ranges, int n_bins, int hop, int sr, double hi_db, double lo_db,
                       double fit_lower_limit_db, int min_points, double floor_db, double min_peak_above_floor,
                       float* d_curve_ws, double* d_out, cudaStream_t stream);

}  // namespace aa
