// extern "C" surface of libaudio_analysis_b200 (see include/audio_analysis_b200.h).
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <stdio.h>
#include <string>
#include <vector>

#include "fft.cuh"
#include "kernels.cuh"

struct aa_bands_plan;
struct aa_deconv_plan;

namespace aa {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-stage timing (off by default) ----------------------------------------------------------
struct ProfRec {
    std::string name;
    cudaEvent_t e0, e1;
};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_open, g_prof_done;
static std::map<std::string, std::pair<double, long long>> g_prof_sum;  // name -> (ms, scopes)

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void prof_begin(const char* name, cudaStream_t stream) {
    ProfRec r;
    r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, stream);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_open.push_back(r);
}

void prof_end(cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (g_prof_open.empty()) return;
    ProfRec r = g_prof_open.back();
    g_prof_open.pop_back();
    cudaEventRecord(r.e1, stream);
    g_prof_done.push_back(r);
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int bands_plan_create(int n, aa_bands_plan** out);
void bands_plan_destroy(aa_bands_plan* p);
size_t bands_workspace_bytes(const aa_bands_plan* p, int rows, int nb);
int bands_plan_kind(const aa_bands_plan* p);
int band_decimation_limit();
void set_band_decimation_limit(int max_d);
int bands_classify(const aa_bands_plan* plan, const aa_band* bands, int nb, double tw_oct, int sample_rate_hz,
                   const aa_decay_params* fits, int* out_decimation);
int rt60bands_batch(const aa_bands_plan* plan, const float* d_x, long long row_stride, int n_rows,
                    const int* d_start_idx, const aa_band* bands, int nb, double tw_oct, const aa_decay_params* p,
                    double* d_metrics_out, float* d_band_signal_out, void* d_ws, size_t ws_bytes, cudaStream_t stream,
                    const aa_bands_options* opts, aa_bands_stats* stats);
size_t sos_workspace_bytes(int n, int n_sec_max, int rows, int nb);
int rt60bands_sos_batch(const float* d_x, long long row_stride, int n, int n_rows, const int* d_start_idx,
                        const double* sos, const double* zi, const int* edges, int nb, int n_sec, const aa_decay_params* p,
                        double* d_metrics_out, float* d_band_signal_out, void* d_ws, size_t ws_bytes, cudaStream_t stream);
int deconv_plan_create(int n_fft, aa_deconv_plan** out);
void deconv_plan_destroy(aa_deconv_plan* p);
size_t deconv_workspace_bytes(const aa_deconv_plan* p, int rows);
int deconvolve_batch(const aa_deconv_plan* plan, const float* d_sweep, int n_sweep, const float* d_rec,
                     long long rec_stride, int n_rec, int n_rows, double reg_rel, int remove_dc, double target_peak,
                     int norm_group, int n_out, float* d_ir_out, long long ir_stride, void* d_ws, size_t ws_bytes,
                     cudaStream_t stream);

}  // namespace aa

extern "C" {

int aa_abi_version(void) { return AA_ABI_VERSION; }

const char* aa_last_error(void) { return aa::g_err; }

long long aa_kernel_launch_count(void) { return aa::g_launches.load(std::memory_order_relaxed); }

void aa_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(aa::g_prof_mu);
    aa::g_prof_on.store(on != 0);
    for (auto& r : aa::g_prof_done) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    aa::g_prof_done.clear();
    aa::g_prof_sum.clear();
}

int aa_profile_collect(void) {
    std::lock_guard<std::mutex> lock(aa::g_prof_mu);
    for (auto& r : aa::g_prof_done) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto& acc = aa::g_prof_sum[r.name];
            acc.first += (double)ms;
            acc.second += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    aa::g_prof_done.clear();
    return (int)aa::g_prof_sum.size();
}

double aa_profile_ms(const char* name) {
    std::lock_guard<std::mutex> lock(aa::g_prof_mu);
    auto it = aa::g_prof_sum.find(name ? name : "");
    return it == aa::g_prof_sum.end() ? 0.0 : it->second.first;
}

long long aa_profile_scopes(const char* name) {
    std::lock_guard<std::mutex> lock(aa::g_prof_mu);
    auto it = aa::g_prof_sum.find(name ? name : "");
    return it == aa::g_prof_sum.end() ? 0 : it->second.second;
}

int aa_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        aa::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return AA_ERR_NO_DEVICE;
    }
    return n;
}

int aa_peak_index_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t n_rows,
                        int32_t* d_peak_out, void* stream) {
    AA_REQUIRE(d_lengths != nullptr, AA_ERR_INVALID_ARGUMENT, "d_lengths is required");
    return aa::launch_peak(d_x, row_stride, d_lengths, 0, n_rows, d_peak_out, static_cast<cudaStream_t>(stream));
}

int aa_decay_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t uniform_length,
                   int32_t n_rows, const aa_decay_params* p, const int32_t* d_start_in, float* d_edc_db_out,
                   int64_t edc_stride, int32_t* d_start_out, double* d_fits_out, double* d_early10_out,
                   int32_t* d_status_out, void* stream) {
    AA_REQUIRE(uniform_length >= 0 && uniform_length <= row_stride, AA_ERR_INVALID_ARGUMENT,
               "uniform_length (the maximum row length) must be in [0, row_stride]");
    aa::ProfScope prof(d_edc_db_out ? "edc" : "edc_metrics", static_cast<cudaStream_t>(stream));
    return aa::launch_edc(d_x, row_stride, d_lengths, uniform_length, uniform_length, n_rows, p, d_start_in, 1, 4,
                          d_edc_db_out, edc_stride, d_start_out, d_fits_out, nullptr, d_early10_out, d_status_out,
                          static_cast<cudaStream_t>(stream));
}

int aa_curve_fits_batch(const float* d_curve, int64_t stride, const int32_t* d_lengths, int32_t uniform_length,
                        const int32_t* d_offsets, int32_t n_rows, const aa_decay_params* p, double* d_fits_out,
                        double* d_early10_out, void* stream) {
    return aa::launch_curve_fits(d_curve, stride, d_lengths, uniform_length, d_offsets, n_rows, p, d_fits_out,
                                 nullptr, d_early10_out, static_cast<cudaStream_t>(stream));
}

int aa_box_smooth_db_batch(const float* d_in, float* d_out, int64_t stride, const int32_t* d_lengths,
                           int32_t uniform_length, const int32_t* d_offsets, int32_t n_rows, int32_t window,
                           double floor_db, void* stream) {
    return aa::launch_box_smooth(d_in, d_out, stride, d_lengths, uniform_length, d_offsets, n_rows, window, floor_db,
                                 static_cast<cudaStream_t>(stream));
}

int aa_bands_plan_create(int32_t n, aa_bands_plan** out) { return aa::bands_plan_create(n, out); }
void aa_bands_plan_destroy(aa_bands_plan* plan) { aa::bands_plan_destroy(plan); }
int aa_bands_plan_kind(const aa_bands_plan* plan) { return aa::bands_plan_kind(plan); }
int aa_bands_get_decimation(void) { return aa::band_decimation_limit(); }
void aa_bands_set_decimation(int max_decimation) { aa::set_band_decimation_limit(max_decimation); }
int aa_bands_classify(const aa_bands_plan* plan, const aa_band* bands, int32_t n_bands, double transition_width_octaves,
                      int32_t sample_rate_hz, const aa_decay_params* fits, int32_t* out_decimation) {
    return aa::bands_classify(plan, bands, n_bands, transition_width_octaves, sample_rate_hz, fits, out_decimation);
}
size_t aa_bands_workspace_bytes(const aa_bands_plan* plan, int32_t rows_per_chunk, int32_t n_bands) {
    return aa::bands_workspace_bytes(plan, rows_per_chunk, n_bands);
}

int aa_rt60bands_batch(const aa_bands_plan* plan, const float* d_x, int64_t row_stride, int32_t n_rows,
                       const int32_t* d_start_idx, const aa_band* bands, int32_t n_bands,
                       double transition_width_octaves, const aa_decay_params* p, double* d_metrics_out,
                       float* d_band_signal_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    return aa::rt60bands_batch(plan, d_x, row_stride, n_rows, d_start_idx, bands, n_bands, transition_width_octaves, p,
                               d_metrics_out, d_band_signal_out, d_workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

int aa_rt60bands_batch_ex(const aa_bands_plan* plan, const float* d_x, int64_t row_stride, int32_t n_rows,
                          const int32_t* d_start_idx, const aa_band* bands, int32_t n_bands,
                          double transition_width_octaves, const aa_decay_params* p, double* d_metrics_out,
                          float* d_band_signal_out, void* d_workspace, size_t workspace_bytes, void* stream,
                          const aa_bands_options* options, aa_bands_stats* stats_out) {
    return aa::rt60bands_batch(plan, d_x, row_stride, n_rows, d_start_idx, bands, n_bands, transition_width_octaves, p,
                               d_metrics_out, d_band_signal_out, d_workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream), options, stats_out);
}

size_t aa_sos_workspace_bytes(int32_t n, int32_t n_sections, int32_t rows_per_chunk, int32_t n_bands) {
    return aa::sos_workspace_bytes(n, n_sections, rows_per_chunk, n_bands);
}
int aa_rt60bands_sos_batch(const float* d_x, int64_t row_stride, int32_t n, int32_t n_rows, const int32_t* d_start_idx,
                           const double* h_sos, const double* h_zi, const int32_t* h_edge, int32_t n_bands,
                           int32_t n_sections, const aa_decay_params* p, double* d_metrics_out, float* d_band_signal_out,
                           void* d_workspace, size_t workspace_bytes, void* stream) {
    return aa::rt60bands_sos_batch(d_x, row_stride, n, n_rows, d_start_idx, h_sos, h_zi, h_edge, n_bands, n_sections, p,
                                   d_metrics_out, d_band_signal_out, d_workspace, workspace_bytes,
                                   static_cast<cudaStream_t>(stream));
}

int aa_csd_plan_batch(const int32_t* d_lengths, int32_t uniform_length, int32_t n_rows, const int32_t* d_peak_idx,
                      int32_t ignore_leading_samples, int32_t analysis_duration_samples, int32_t n_fft, int32_t hop_length,
                      const int32_t* d_slice_table, const int32_t* d_slice_counts, int32_t max_frames, int32_t table_width,
                      int32_t* d_start_out, int32_t* d_span_out, int32_t* d_frame_idx_out, int32_t* d_n_slices_out,
                      void* stream) {
    return aa::launch_csd_plan(d_lengths, uniform_length, n_rows, d_peak_idx, ignore_leading_samples,
                               analysis_duration_samples, n_fft, hop_length, d_slice_table, d_slice_counts, max_frames,
                               table_width, d_start_out, d_span_out, d_frame_idx_out, d_n_slices_out,
                               static_cast<cudaStream_t>(stream));
}

int aa_csd_smooth_batch(float* d_slices, int32_t n_rows, const int32_t* d_n_slices, int32_t max_slices, int32_t n_bins,
                        int32_t first_bin_inside, int32_t n_bins_inside, const double* d_freq_inside,
                        const double* d_grid_hz, int32_t n_grid, const int32_t* d_grid_segment,
                        const int32_t* d_bin_segment, int32_t smoothing_log_bins, int32_t db_reference,
                        double dynamic_range_db, void* stream) {
    return aa::launch_csd_smooth(d_slices, n_rows, d_n_slices, max_slices, n_bins, first_bin_inside, n_bins_inside,
                                 d_freq_inside, d_grid_hz, n_grid, d_grid_segment, d_bin_segment, smoothing_log_bins,
                                 db_reference == AA_DBREF_SLICE_MAX, dynamic_range_db, static_cast<cudaStream_t>(stream));
}

int aa_stft_db_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t uniform_length,
                     int32_t n_rows, const int32_t* d_start_idx, const int32_t* d_span, int32_t n_fft, int32_t hop_length,
                     int32_t use_hann_window, double floor_db, int32_t frames_max, float* d_out, void* stream) {
    aa::ProfScope prof("stft", static_cast<cudaStream_t>(stream));
    return aa::launch_stft_db(d_x, row_stride, d_lengths, uniform_length, n_rows, d_start_idx, d_span, n_fft, hop_length,
                              use_hann_window, floor_db, frames_max, d_out, static_cast<cudaStream_t>(stream));
}

int aa_modal_cloud_batch(const float* d_mag_db, int32_t n_rows, int32_t frames_max, int32_t n_freq,
                         const int32_t* d_n_frames, const int32_t* d_bin_ranges, int32_t n_bins, int32_t hop_length,
                         int32_t sample_rate_hz, double range_hi_db, double range_lo_db, double fit_lower_limit_db,
                         int32_t min_fit_points, double floor_db, double min_peak_db_above_floor, float* d_curve_ws,
                         double* d_out, void* stream) {
    return aa::launch_modal_cloud(d_mag_db, n_rows, frames_max, n_freq, d_n_frames, d_bin_ranges, n_bins, hop_length,
                                  sample_rate_hz, range_hi_db, range_lo_db, fit_lower_limit_db, min_fit_points, floor_db,
                                  min_peak_db_above_floor, d_curve_ws, d_out, static_cast<cudaStream_t>(stream));
}

int aa_csd_batch(const float* d_x, int64_t row_stride, const int32_t* d_lengths, int32_t uniform_length,
                 int32_t n_rows, const int32_t* d_start_idx, const int32_t* d_frame_idx, const int32_t* d_n_slices,
                 int32_t max_slices, const aa_csd_params* p, float* d_out, void* stream) {
    aa::ProfScope prof("csd", static_cast<cudaStream_t>(stream));
    return aa::launch_csd(d_x, row_stride, d_lengths, uniform_length, n_rows, d_start_idx, d_frame_idx, d_n_slices,
                          max_slices, p, d_out, static_cast<cudaStream_t>(stream));
}

int aa_pcm16_to_float_batch(const int16_t* d_pcm, int64_t pcm_stride, int32_t n_files, int32_t n_frames,
                            int32_t n_channels, int32_t mode, float* d_out, int64_t out_stride, void* stream) {
    return aa::launch_pcm16(d_pcm, pcm_stride, n_files, n_frames, n_channels, mode, d_out, out_stride,
                            static_cast<cudaStream_t>(stream));
}

int aa_deconv_plan_create(int32_t n_fft, aa_deconv_plan** out) { return aa::deconv_plan_create(n_fft, out); }
void aa_deconv_plan_destroy(aa_deconv_plan* plan) { aa::deconv_plan_destroy(plan); }
size_t aa_deconv_workspace_bytes(const aa_deconv_plan* plan, int32_t n_rows) {
    return aa::deconv_workspace_bytes(plan, n_rows);
}

int aa_deconvolve_batch(const aa_deconv_plan* plan, const float* d_sweep, int32_t n_sweep, const float* d_rec,
                        int64_t rec_stride, int32_t n_rec, int32_t n_rows, double regularization_relative,
                        int32_t remove_dc, double target_peak, int32_t norm_group, int32_t n_out, float* d_ir_out,
                        int64_t ir_stride, void* d_workspace, size_t workspace_bytes, void* stream) {
    return aa::deconvolve_batch(plan, d_sweep, n_sweep, d_rec, rec_stride, n_rec, n_rows, regularization_relative,
                                remove_dc, target_peak, norm_group, n_out, d_ir_out, ir_stride, d_workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
