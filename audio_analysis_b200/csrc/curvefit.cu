// Generic level-crossing + line-fit kernel on an arbitrary dB curve, and the optional
// dB-domain box smoothing of the EDC.
//
//  - aa_curve_fits_batch replaces _interpolated_crossing_time_seconds + fit_decay_slope_over_db_range
//    (analyse/decay.py:173-260) when the caller supplies the curve itself, or when the curve is not
//    monotone (smoothed EDC, decay.py:161-164), so the "first sample <= target" search is a true
//    minimum-index reduction instead of the tile search of the fused kernel.
//  - aa_box_smooth_db_batch replaces np.convolve(edc_db, ones(w)/w, "same") followed by the floor
//    (analyse/decay.py:161-167).
#include <math.h>

#include "aa_common.cuh"
#include "fit_math.cuh"
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int kCfThreads = 512;
constexpr int kCfWarps = kCfThreads / 32;
constexpr int kMaxTargets = 2 + 2 * AA_MAX_FITS;

struct CurveArgs {
    const float* curve;
    long long stride;
    const int* lengths;
    int uniform_len;
    const int* offsets;  // per-row index of the first curve sample inside the row (NULL = 0)
    int sr;
    int n_fits;
    int want_early10;
    double target[kMaxTargets];
    double fit_hi[AA_MAX_FITS], fit_lo[AA_MAX_FITS];
    double* fits_out;
    double* rt60_out;  // optional [row][AA_MAX_FITS]
    double* early10_out;
};

__global__ void __launch_bounds__(kCfThreads) curve_fits_kernel(CurveArgs a) {
    const int row = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int off = a.offsets ? a.offsets[row] : 0;
    const int full = a.lengths ? a.lengths[row] : a.uniform_len;
    const int n = full - off;
    const float* __restrict__ c = a.curve + (long long)row * a.stride + off;
    const float srf = (float)a.sr;
    const int nt = 2 + 2 * a.n_fits;

    __shared__ int s_min[kCfWarps][kMaxTargets];
    __shared__ int s_idx[kMaxTargets];
    __shared__ int s_ok[kMaxTargets];
    __shared__ double s_t[kMaxTargets];
    __shared__ int s_valid[AA_MAX_FITS], s_first[AA_MAX_FITS], s_last[AA_MAX_FITS];
    __shared__ double s_acc[kCfWarps][AA_MAX_FITS * 3];

    // pass 1: first index with curve <= target, for every target (decay.py:181-185)
    int best[kMaxTargets];
    float tf[kMaxTargets];
#pragma unroll
    for (int t = 0; t < kMaxTargets; ++t) {
        best[t] = 0x7fffffff;
        tf[t] = (t < nt) ? (float)a.target[t] : -INFINITY;
    }
    for (int i = threadIdx.x; i < n; i += kCfThreads) {
        const float v = __ldg(c + i);
#pragma unroll
        for (int t = 0; t < kMaxTargets; ++t)
            if (v <= tf[t] && i < best[t]) best[t] = i;
    }
#pragma unroll
    for (int t = 0; t < kMaxTargets; ++t) {
        int b = best[t];
        for (int d = 16; d > 0; d >>= 1) b = min(b, __shfl_xor_sync(kFullMask, b, d));
        if (lane == 0) s_min[warp][t] = b;
    }
    __syncthreads();
    if (threadIdx.x < nt) {
        const int t = threadIdx.x;
        int b = 0x7fffffff;
        for (int w = 0; w < kCfWarps; ++w) b = min(b, s_min[w][t]);
        const int ok = (b != 0x7fffffff);
        double tc = 0.0;
        if (ok) {
            const float y1 = c[b];
            const float y0 = (b > 0) ? c[b - 1] : y1;
            tc = crossing_time_rel(b, y0, y1, a.target[t], srf);
        }
        s_idx[t] = ok ? b : 0;
        s_ok[t] = ok;
        s_t[t] = tc;
    }
    __syncthreads();
    if (threadIdx.x < a.n_fits) {
        const int f = threadIdx.x;
        const int ts_i = 2 + 2 * f, te_i = 3 + 2 * f;
        int first, last;
        s_valid[f] = fit_index_range(s_ok[ts_i], s_ok[te_i], s_t[ts_i], s_t[te_i], s_idx[ts_i], s_idx[te_i], srf,
                                     &first, &last);
        s_first[f] = first;
        s_last[f] = last;
    }
    __syncthreads();

    // pass 2: regression sums over each fit's index range
    double acc[AA_MAX_FITS * 3];
#pragma unroll
    for (int q = 0; q < AA_MAX_FITS * 3; ++q) acc[q] = 0.0;
    for (int f = 0; f < a.n_fits; ++f) {
        if (!s_valid[f]) continue;
        const double ymid = 0.5 * (a.fit_hi[f] + a.fit_lo[f]);
        const int first = s_first[f], last = s_last[f];
        for (int i = first + threadIdx.x; i <= last; i += kCfThreads) {
            const double u = (double)(i - first);
            const double yc = (double)__ldg(c + i) - ymid;
            acc[f * 3 + 0] += yc;
            acc[f * 3 + 1] = fma(u, yc, acc[f * 3 + 1]);
            acc[f * 3 + 2] = fma(yc, yc, acc[f * 3 + 2]);
        }
    }
#pragma unroll
    for (int q = 0; q < AA_MAX_FITS * 3; ++q) {
        double s = warp_sum(acc[q]);
        if (lane == 0) s_acc[warp][q] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.early10_out) {
            double e10 = nan("");
            if (a.want_early10 && s_ok[0] && s_ok[1] && s_t[1] >= s_t[0]) e10 = s_t[1] - s_t[0];
            a.early10_out[row] = e10;
        }
        for (int f = 0; f < AA_MAX_FITS; ++f) {
            double res[AA_FIT_COLS];
            for (int q = 0; q < 6; ++q) res[q] = nan("");
            res[6] = 0.0;
            res[7] = 0.0;
            if (f < a.n_fits && s_valid[f]) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int w = 0; w < kCfWarps; ++w) {
                    s0 += s_acc[w][f * 3 + 0];
                    s1 += s_acc[w][f * 3 + 1];
                    s2 += s_acc[w][f * 3 + 2];
                }
                fit_finalize(s0, s1, s2, s_first[f], s_last[f], 0.5 * (a.fit_hi[f] + a.fit_lo[f]), a.sr,
                             s_t[2 + 2 * f], s_t[3 + 2 * f], res);
            }
            if (a.fits_out) {
                double* o = a.fits_out + ((long long)row * AA_MAX_FITS + f) * AA_FIT_COLS;
                for (int q = 0; q < AA_FIT_COLS; ++q) o[q] = res[q];
            }
            if (a.rt60_out) a.rt60_out[(long long)row * AA_MAX_FITS + f] = res[5];
        }
    }
}

// out[i] = max(floor, (1/w) * sum_{m = i - w/2}^{i + (w-1)/2} in[m]), zero outside the curve
__global__ void box_smooth_kernel(const float* __restrict__ in, float* __restrict__ out, long long stride,
                                  const int* __restrict__ lengths, int uniform_len, const int* __restrict__ offsets,
                                  int w, float floor_db, int row0) {
    const int row = row0 + blockIdx.y;
    const int off = offsets ? offsets[row] : 0;
    const int n = (lengths ? lengths[row] : uniform_len) - off;
    const float* __restrict__ c = in + (long long)row * stride + off;
    float* __restrict__ o = out + (long long)row * stride + off;
    const int back = w / 2, fwd = (w - 1) / 2;
    const double inv = 1.0 / (double)w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int lo = max(0, i - back), hi = min(n - 1, i + fwd);
        double s = 0.0;
        for (int m = lo; m <= hi; ++m) s += (double)__ldg(c + m) * inv;
        o[i] = fmaxf((float)s, floor_db);
    }
}

}  // namespace

int launch_curve_fits(const float* d_curve, long long stride, const int* d_lengths, int uniform_len,
                      const int* d_offsets, int n_rows, const aa_decay_params* p, double* d_fits_out,
                      double* d_rt60_out, double* d_early10_out, cudaStream_t stream) {
    AA_REQUIRE(d_curve && p && (d_fits_out || d_rt60_out), AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(p->n_fits >= 0 && p->n_fits <= AA_MAX_FITS, AA_ERR_INVALID_ARGUMENT, "n_fits out of range");
    AA_REQUIRE(p->sample_rate_hz > 0, AA_ERR_INVALID_ARGUMENT, "sample rate must be positive");
    if (n_rows <= 0) return AA_OK;
    CurveArgs a{};
    a.curve = d_curve;
    a.stride = stride;
    a.lengths = d_lengths;
    a.uniform_len = uniform_len;
    a.offsets = d_offsets;
    a.sr = p->sample_rate_hz;
    a.n_fits = p->n_fits;
    a.want_early10 = p->want_early10;
    a.target[0] = 0.0;
    a.target[1] = -10.0;
    for (int f = 0; f < AA_MAX_FITS; ++f) {
        double hi = f < p->n_fits ? p->fit_hi_db[f] : 0.0;
        double lo = f < p->n_fits ? p->fit_lo_db[f] : 0.0;
        if (f < p->n_fits) AA_REQUIRE(lo <= hi, AA_ERR_INVALID_ARGUMENT, "fit range must be (higher_db, lower_db)");
        a.fit_hi[f] = hi;
        a.fit_lo[f] = lo;
        a.target[2 + 2 * f] = hi;
        a.target[3 + 2 * f] = lo > p->fit_lower_limit_db ? lo : p->fit_lower_limit_db;
    }
    a.fits_out = d_fits_out;
    a.rt60_out = d_rt60_out;
    a.early10_out = d_early10_out;
    curve_fits_kernel<<<n_rows, kCfThreads, 0, stream>>>(a);
    return launch_check("curve_fits_kernel");
}

int launch_box_smooth(const float* d_in, float* d_out, long long stride, const int* d_lengths, int uniform_len,
                      const int* d_offsets, int n_rows, int window, double floor_db, cudaStream_t stream) {
    AA_REQUIRE(d_in && d_out && d_in != d_out, AA_ERR_INVALID_ARGUMENT, "in/out must be distinct buffers");
    AA_REQUIRE(window >= 1, AA_ERR_INVALID_ARGUMENT, "window must be >= 1");
    if (n_rows <= 0) return AA_OK;
    int gx = (uniform_len + 255) / 256;
    if (gx > 2048) gx = 2048;
    if (gx < 1) gx = 1;
    // gridDim.y holds at most 65,535 rows: band units (rows x bands) go past that
    for (int r0 = 0; r0 < n_rows; r0 += 65535) {
        const int cnt = (n_rows - r0 < 65535) ? (n_rows - r0) : 65535;
        box_smooth_kernel<<<dim3((unsigned)gx, (unsigned)cnt), 256, 0, stream>>>(d_in, d_out, stride, d_lengths, uniform_len,
                                                                               d_offsets, window, (float)floor_db, r0);
        int rc = launch_check("box_smooth_kernel");
        if (rc != AA_OK) return rc;
    }
    return AA_OK;
}

}  // namespace aa
