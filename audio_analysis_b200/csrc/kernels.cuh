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

int launch_csd_plan(const int* d_lengths, int uniform_len, int n_rows, const int* d_peak, int ignore, int keep, int n_fft,
                    int hop, const int* d_table, const int* d_counts, int max_frames, int width, int* d_start_out,
                    int* d_span_out, int* d_frames_out, int* d_nslices_out, cudaStream_t stream);

int launch_csd_smooth(float* d_data, int n_rows, const int* d_n_slices, int max_slices, int n_bins, int k0, int n_in,
                      const double* d_f_in, const double* d_grid, int n_grid, const int* d_grid_seg,
                      const int* d_bin_seg, int window, int slice_max, double dynamic_range_db, cudaStream_t stream);

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
//
// BandFitCert: addressing of a class's results and, for a decimated level, the certification of its fits.
//   out_band[j]  : position of class band j in the caller's table (always used)
//   on           : decide per unit whether the decimated fits may be kept (band_fit_kernel: "certification");
//                  a unit that is not certified is appended to list next_list[j] (which starts at lists + next_off[j]
//                  and whose length is counts[next_list[j]]) as row * next_nbc[j] + next_j[j]
//   m[j]         : half-width of the crossing window in decimated samples of a straight decay (safety factor x the
//                  host's quadrature-error model of band j at this decimation, bands.cu: band_quadrature_error)
//   w            : a crossing window may span at most w x (decimated samples between the fit's end points)
//   min_span[j]  : decimated samples a certified 30 dB fit of band j needs between its end points (the steps of a low
//                  band's curve make the fit sensitive to which end samples the grid includes);
//                  fit f needs min_span[j] x fit_scale[f] (shorter fits need more)
struct BandFitCert {
    int on;
    float w;
    float tol;   // largest modelled relative slope change a certified fit may suffer from its crossing windows
    float fit_scale[AA_MAX_FITS];
    int min_span[AA_MAX_BANDS];
    float m[AA_MAX_BANDS];
    short out_band[AA_MAX_BANDS];
    short next_list[AA_MAX_BANDS], next_j[AA_MAX_BANDS], next_nbc[AA_MAX_BANDS];
    long long next_off[AA_MAX_BANDS];
    int* counts;
    int* lists;
};

// unit = row * nbc + j (dense) or d_unit_map[unit] = row * nbc + j; cert == NULL: nbc == out_nb, identity band map
int launch_band_fit(const float* d_y, long long y_stride, const float* d_part, long long part_stride, int n_groups,
                    int n_tiles, int tile_len, int n, long long n_units, const aa_decay_params* p,
                    const int* d_start_idx, int nbc, int out_nb, const int* d_unit_map, const BandFitCert* cert,
                    int min_samples, double* d_rt60_out, cudaStream_t stream, int decim = 1, int full_n = 0,
                    int aligned = 0);

// bandfused.cu: the same per-unit analysis with the band signal synthesised into shared memory by the CTA that
// fits it (classes whose transform of M complex points fits one CTA).  d_X: time-advanced half spectra of the rows,
// d_bands: the class's bands, twN: exp(-2 pi i k / 2M), scale: 1 / (M D)
bool band_fused_supported(int M);
int launch_band_fused(int M, const float2* d_X, long long x_stride, const struct BandDev* d_bands, double freq_scale,
                      struct TwoLevelTwiddle twN, float scale, long long n_units, const aa_decay_params* p,
                      const int* d_start_idx, int nbc, int out_nb, const int* d_unit_map, const BandFitCert* cert,
                      int min_samples, double* d_rt60_out, cudaStream_t stream, int decim, int full_n);

}  // namespace aa

namespace aa {

// modal cloud (modal.cu): log-bin aggregation of the dense STFT + per-bin decay fits
int launch_modal_cloud(const float* d_mag_db, int n_rows, int frames_max, int n_freq, const int* d_n_frames,
                       const int* d_bin_ranges, int n_bins, int hop, int sr, double hi_db, double lo_db,
                       double fit_lower_limit_db, int min_points, double floor_db, double min_peak_above_floor,
                       float* d_curve_ws, double* d_out, cudaStream_t stream);

}  // namespace aa
