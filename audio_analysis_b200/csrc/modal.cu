// Modal cloud (analyse/modalcloud.py:121-391): per log-frequency bin, a decay curve from the
// dense STFT and a line fit over it.
//
//   modal_aggregate_kernel  one CTA per (frame, row): the frame's dB spectrum (aa_stft_db_batch output, frame-major)
//                           is converted back to linear magnitude in shared memory and averaged over the rfft bins of
//                           every log bin (modalcloud.py:176-207), giving curve[row][bin][frame] in dB;
//   modal_fit_kernel        one CTA per (bin, row): peak normalisation and pruning (:347-353), first crossings of
//                           the range ends with linear interpolation (:215-236), the float32 time mask and
//                           min_fit_points (:259-261), closed-form least squares, RT60 = -60 / slope (:266-281).
// Times are the reference's float32 frame times, float32(frame * hop) / sr (:158).
#include <math.h>

#include "aa_common.cuh"
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int kAggThreads = 256;
constexpr int kFitThreads = 128;
constexpr int kFitWarps = kFitThreads / 32;

__global__ void __launch_bounds__(kAggThreads) modal_aggregate_kernel(const float* __restrict__ mag_db, int frames_max,
                                                                       int n_freq, const int* __restrict__ n_frames,
                                                                       const int* __restrict__ ranges, int n_bins,
                                                                       float* __restrict__ curve, int curve_pitch) {
    extern __shared__ float lin[];  // [n_freq] linear magnitudes of this frame
    const int t = blockIdx.x, row = blockIdx.y;
    if (t >= n_frames[row]) return;
    const float* __restrict__ src = mag_db + ((long long)row * frames_max + t) * n_freq;
    for (int k = threadIdx.x; k < n_freq; k += kAggThreads)
        lin[k] = exp2f(__ldg(src + k) * 0.16609640474436813f);   // 10^(dB / 20), modalcloud.py:190
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += kAggThreads) {
        const int k0 = ranges[2 * b], k1 = ranges[2 * b + 1];
        float v = nanf("");
        if (k1 > k0) {   // bins without any rfft line stay NaN (:199-200)
            float s = 0.f;
            for (int k = k0; k < k1; ++k) s += lin[k];
            const float m = fmaxf(s / (float)(k1 - k0), 1e-30f);
            v = 6.020599913279624f * log2f(m);   // 20 log10
        }
        curve[((long long)row * n_bins + b) * curve_pitch + t] = v;
    }
}

struct ModalFitArgs {
    const float* curve;
    int curve_pitch;
    const int* n_frames;
    int n_bins;
    int hop, sr;
    double hi_db, lo_eff_db;     // range start target, effective range end target (max(low, fit_lower_limit))
    int min_points;
    float floor_db, min_peak_above_floor;
    double* out;                 // [row][bin][2]: rt60 seconds, r^2; NaN = no point
};

__device__ __forceinline__ float frame_time(int i, int hop, float srf) { return __fdiv_rn((float)(i * hop), srf); }

__global__ void __launch_bounds__(kFitThreads) modal_fit_kernel(ModalFitArgs a) {
    const int b = blockIdx.x, row = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = a.n_frames[row];
    const float* __restrict__ c = a.curve + ((long long)row * a.n_bins + b) * a.curve_pitch;
    double* __restrict__ out = a.out + ((long long)row * a.n_bins + b) * 2;
    const float srf = (float)a.sr;
    __shared__ float s_max[kFitWarps];
    __shared__ int s_bad[kFitWarps];
    __shared__ int s_i0[kFitWarps], s_i1[kFitWarps];
    __shared__ double s_acc[kFitWarps][3];
    __shared__ int s_cnt[kFitWarps], s_first[kFitWarps];
    __shared__ float s_peak;
    __shared__ int s_cross[2];
    __shared__ double s_tc[2];
    __shared__ int s_state;   // 0 = go on, 1 = no point

    // peak over time, and "all finite" (:343-347)
    float mx = -INFINITY;
    int bad = 0;
    for (int i = threadIdx.x; i < T; i += kFitThreads) {
        const float v = c[i];
        if (!isfinite(v)) bad = 1;
        mx = fmaxf(mx, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, d));
        bad |= __shfl_xor_sync(kFullMask, bad, d);
    }
    if (lane == 0) { s_max[warp] = mx; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_max[0];
        int bd = s_bad[0];
        for (int w = 1; w < kFitWarps; ++w) { m = fmaxf(m, s_max[w]); bd |= s_bad[w]; }
        s_peak = m;
        s_state = (T < 1 || bd || (m - a.floor_db) < a.min_peak_above_floor) ? 1 : 0;   // :349-351
    }
    __syncthreads();
    if (s_state) {
        if (threadIdx.x == 0) { out[0] = nan(""); out[1] = nan(""); }
        return;
    }
    const float peak = s_peak;
    // first index with rel <= target for both range ends (:219-224)
    const float t_hi = (float)a.hi_db, t_lo = (float)a.lo_eff_db;
    int i0 = 0x7fffffff, i1 = 0x7fffffff;
    for (int i = threadIdx.x; i < T; i += kFitThreads) {
        const float rel = __fsub_rn(c[i], peak);
        if (rel <= t_hi && i < i0) i0 = i;
        if (rel <= t_lo && i < i1) i1 = i;
    }
    for (int d = 16; d > 0; d >>= 1) {
        i0 = min(i0, __shfl_xor_sync(kFullMask, i0, d));
        i1 = min(i1, __shfl_xor_sync(kFullMask, i1, d));
    }
    if (lane == 0) { s_i0[warp] = i0; s_i1[warp] = i1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        int best = 0x7fffffff;
        for (int w = 0; w < kFitWarps; ++w) best = min(best, threadIdx.x == 0 ? s_i0[w] : s_i1[w]);
        const double target = threadIdx.x == 0 ? a.hi_db : a.lo_eff_db;
        double tc = 0.0;
        if (best != 0x7fffffff) {
            if (best == 0) {
                tc = (double)frame_time(0, a.hop, srf);
            } else {   // :228-236
                const double tA = (double)frame_time(best - 1, a.hop, srf), tB = (double)frame_time(best, a.hop, srf);
                const double y0 = (double)__fsub_rn(c[best - 1], peak), y1 = (double)__fsub_rn(c[best], peak);
                if (y1 == y0) tc = tB;
                else {
                    double fr = (target - y0) / (y1 - y0);
                    fr = fmin(fmax(fr, 0.0), 1.0);
                    tc = tA + fr * (tB - tA);
                }
            }
        }
        s_cross[threadIdx.x] = (best != 0x7fffffff);
        s_tc[threadIdx.x] = tc;
    }
    __syncthreads();
    const bool ok = s_cross[0] && s_cross[1] && s_tc[1] > s_tc[0];   // :256-257
    // mask = (time >= t_start) & (time <= t_end) on float32 times (:259): count, first index, regression sums
    const float ts = (float)s_tc[0], te = (float)s_tc[1];
    const double ymid = 0.5 * (a.hi_db + a.lo_eff_db);
    int cnt = 0, first = 0x7fffffff;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    if (ok) {
        for (int i = threadIdx.x; i < T; i += kFitThreads) {
            const float ti = frame_time(i, a.hop, srf);
            if (ti >= ts && ti <= te) {
                ++cnt;
                first = min(first, i);
                const double yc = (double)__fsub_rn(c[i], peak) - ymid;
                acc0 += yc;
                acc1 = fma((double)i, yc, acc1);
                acc2 = fma(yc, yc, acc2);
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(kFullMask, cnt, d);
        first = min(first, __shfl_xor_sync(kFullMask, first, d));
    }
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    acc2 = warp_sum(acc2);
    if (lane == 0) {
        s_cnt[warp] = cnt;
        s_first[warp] = first;
        s_acc[warp][0] = acc0;
        s_acc[warp][1] = acc1;
        s_acc[warp][2] = acc2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0, f0 = 0x7fffffff;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < kFitWarps; ++w) {
            n += s_cnt[w];
            f0 = min(f0, s_first[w]);
            s0 += s_acc[w][0];
            s1 += s_acc[w][1];
            s2 += s_acc[w][2];
        }
        double rt60 = nan(""), r2 = nan("");
        if (ok && n >= a.min_points && n >= 2) {   // :260-261
            // least squares of y on the frame index (uniform grid: the masked frames are consecutive), then
            // slope per second = slope per frame * sr / hop; r^2 does not depend on the time unit (:266-281)
            const double nn = (double)n;
            s1 -= (double)f0 * s0;                       // u = i - first
            const double ubar = 0.5 * (nn - 1.0);
            const double suu = nn * (nn * nn - 1.0) / 12.0;
            const double suy = s1 - ubar * s0;
            const double slope_u = suy / suu;
            const double slope = slope_u * (double)a.sr / (double)a.hop;
            if (slope < 0.0) {
                const double ss_tot = s2 - s0 * s0 / nn;
                const double ss_res = ss_tot - slope_u * suy;
                r2 = (ss_tot > 0.0) ? 1.0 - ss_res / ss_tot : 0.0;
                rt60 = -60.0 / slope;
            }
        }
        out[0] = rt60;
        out[1] = isnan(rt60) ? nan("") : r2;
    }
}

}  // namespace

int launch_modal_cloud(const float* d_mag_db, int n_rows, int frames_max, int n_freq, const int* d_n_frames,
                       const int* d_bin_ranges, int n_bins, int hop, int sr, double hi_db, double lo_db,
                       double fit_lower_limit_db, int min_points, double floor_db, double min_peak_above_floor,
                       float* d_curve_ws, double* d_out, cudaStream_t stream) {
    AA_REQUIRE(d_mag_db && d_n_frames && d_bin_ranges && d_curve_ws && d_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(lo_db <= hi_db, AA_ERR_INVALID_ARGUMENT, "range_db should be (higher_db, lower_db), e.g. (-5, -35).");
    AA_REQUIRE(hop > 0 && sr > 0 && n_bins >= 1 && n_freq >= 1, AA_ERR_INVALID_ARGUMENT, "bad sizes");
    AA_REQUIRE(n_rows <= 65535 && (long long)frames_max * hop < (1LL << 24), AA_ERR_INVALID_ARGUMENT,
               "too many rows per call or frames (float32 frame times are exact below 2^24 samples)");
    if (n_rows <= 0 || frames_max <= 0) return AA_OK;
    {
        ProfScope prof("modal_aggregate", stream);
        const size_t smem = sizeof(float) * (size_t)n_freq;
        AA_CUDA_TRY(cudaFuncSetAttribute(modal_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        AA_REQUIRE(smem <= 64 * 1024, AA_ERR_UNSUPPORTED, "n_fft too large for the aggregation kernel");
        modal_aggregate_kernel<<<dim3((unsigned)frames_max, (unsigned)n_rows), kAggThreads, smem, stream>>>(
            d_mag_db, frames_max, n_freq, d_n_frames, d_bin_ranges, n_bins, d_curve_ws, frames_max);
        int rc = launch_check("modal_aggregate_kernel");
        if (rc != AA_OK) return rc;
    }
    ModalFitArgs a{};
    a.curve = d_curve_ws;
    a.curve_pitch = frames_max;
    a.n_frames = d_n_frames;
    a.n_bins = n_bins;
    a.hop = hop;
    a.sr = sr;
    a.hi_db = hi_db;
    a.lo_eff_db = lo_db > fit_lower_limit_db ? lo_db : fit_lower_limit_db;   // modalcloud.py:252
    a.min_points = min_points;
    a.floor_db = (float)floor_db;
    a.min_peak_above_floor = (float)min_peak_above_floor;
    a.out = d_out;
    ProfScope prof("modal_fit", stream);
    modal_fit_kernel<<<dim3((unsigned)n_bins, (unsigned)n_rows), kFitThreads, 0, stream>>>(a);
    return launch_check("modal_fit_kernel");
}

}  // namespace aa
