// Optional log-frequency smoothing of the waterfall slices, then the reference level and the clip
// (analyse/waterfall.py:140-185 _smooth_mag_db_log_frequency / _moving_average_1d, :312-339 of _build_rel_db_slices).
//
// Per slice: the raw dB curve over the kept bins is interpolated onto a uniform log2(f) grid (np.interp), box-averaged
// there (np.convolve(.., "same") on the float32-rounded grid values, zero padded), interpolated back onto the bins
// inside [f_min, f_max] and rounded to float32; bins outside keep their raw value.  The grid, and for every grid point
// / bin the interval of the other axis it falls in, depend only on the settings: the host builds those tables once per
// call (waterfall.py of this package, with the reference's float64 expressions) and the kernel does the arithmetic in
// float64 like numpy.  One CTA per row walks its <= 64 slices, keeps the smoothed table in the output buffer, finds
// the reference level (global or per-slice maximum) and rewrites the values relative to it, clipped to [-dyn, 0].
#include "aa_common.cuh"
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int kSmoothThreads = 256;

struct SmoothArgs {
    float* data;             // [rows][max_slices][n_bins] raw dB in, relative dB out
    const int* n_slices;
    int max_slices, n_bins;
    int k0, n_in;            // bins [k0, k0 + n_in) lie inside [f_min, f_max]
    const double* f_in;      // [n_in] their frequencies
    const double* grid;      // [n_grid] uniform in log2(f)
    int n_grid;
    const int* grid_seg;     // [n_grid] j with f_in[j] <= grid[g] < f_in[j+1]; -1: left of the bins, -2: right
    const int* bin_seg;      // [n_in]   j with grid[j] <= f_in[k] < grid[j+1]; -1 / -2 likewise
    int window;
    int slice_max;           // 1: per-slice reference level, 0: global
    float dyn;
};

// np.interp for one abscissa whose interval is known
__device__ __forceinline__ double interp_at(double x, int seg, const double* __restrict__ xp, const double* fp, int n) {
    if (seg == -1) return fp[0];
    if (seg == -2 || seg >= n - 1) return fp[n - 1];
    const double slope = (fp[seg + 1] - fp[seg]) / (xp[seg + 1] - xp[seg]);
    return slope * (x - xp[seg]) + fp[seg];
}

__global__ void __launch_bounds__(kSmoothThreads) csd_smooth_kernel(SmoothArgs a) {
    extern __shared__ double sm[];
    double* cur = sm;                      // [n_in]  raw dB of the bins inside the range (float64 copy)
    double* on_grid = cur + a.n_in;        // [n_grid] interpolated, rounded to float32 like the reference
    double* smooth = on_grid + a.n_grid;   // [n_grid] box average, rounded to float32
    __shared__ float red[kSmoothThreads / 32];
    __shared__ float level;
    const int row = blockIdx.x;
    const int ns = min(a.n_slices[row], a.max_slices);
    float* __restrict__ base = a.data + (long long)row * a.max_slices * a.n_bins;
    const int back = a.window / 2, fwd = (a.window - 1) / 2;
    const double inv = 1.0 / (double)a.window;
    float row_max = -INFINITY;
    for (int s = 0; s < ns; ++s) {
        float* __restrict__ v = base + (long long)s * a.n_bins;
        for (int k = threadIdx.x; k < a.n_in; k += kSmoothThreads) cur[k] = (double)v[a.k0 + k];
        __syncthreads();
        for (int g = threadIdx.x; g < a.n_grid; g += kSmoothThreads)
            on_grid[g] = (double)(float)interp_at(a.grid[g], a.grid_seg[g], a.f_in, cur, a.n_in);
        __syncthreads();
        for (int g = threadIdx.x; g < a.n_grid; g += kSmoothThreads) {
            const int lo = max(0, g - back), hi = min(a.n_grid - 1, g + fwd);
            double acc = 0.0;
            for (int m = lo; m <= hi; ++m) acc += on_grid[m] * inv;
            smooth[g] = (double)(float)acc;
        }
        __syncthreads();
        float vmax = -INFINITY;
        for (int k = threadIdx.x; k < a.n_bins; k += kSmoothThreads) {
            float o = v[k];
            const int q = k - a.k0;
            if (q >= 0 && q < a.n_in) {
                o = (float)interp_at(a.f_in[q], a.bin_seg[q], a.grid, smooth, a.n_grid);
                v[k] = o;
            }
            vmax = fmaxf(vmax, o);
        }
        if (a.slice_max) {   // reference level of this slice (waterfall.py:331)
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = vmax;
            __syncthreads();
            if (threadIdx.x == 0) {
                float m = red[0];
                for (int w = 1; w < kSmoothThreads / 32; ++w) m = fmaxf(m, red[w]);
                level = m;
            }
            __syncthreads();
            const float lv = level;
            for (int k = threadIdx.x; k < a.n_bins; k += kSmoothThreads) v[k] = fminf(fmaxf(v[k] - lv, -a.dyn), 0.0f);
        } else {
            row_max = fmaxf(row_max, vmax);
        }
        __syncthreads();
    }
    if (!a.slice_max) {   // global reference level over the row's slices (waterfall.py:335)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) row_max = fmaxf(row_max, __shfl_xor_sync(0xffffffffu, row_max, d));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = row_max;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = red[0];
            for (int w = 1; w < kSmoothThreads / 32; ++w) m = fmaxf(m, red[w]);
            level = m;
        }
        __syncthreads();
        const float lv = level;
        const long long total = (long long)ns * a.n_bins;
        for (long long i = threadIdx.x; i < total; i += kSmoothThreads) base[i] = fminf(fmaxf(base[i] - lv, -a.dyn), 0.0f);
    }
}

}  // namespace

namespace {

// Which samples and frames of a row the waterfall analyses (analyse/waterfall.py:358-372, :208, :233-286): pure integer
// bookkeeping per row, done here so that no host round trip and no framework kernel sits between the decay kernel
// (which finds the peak) and the CSD kernel.  The slice selection itself depends only on the number of frames: the
// host tabulates it once per settings (the reference's numpy expressions), the kernel copies the row's line.
__global__ void csd_plan_kernel(const int* __restrict__ lengths, int uniform_len, int n_rows, const int* __restrict__ peak,
                                int ignore, int keep, int n_fft, int hop, const int* __restrict__ table,
                                const int* __restrict__ counts, int max_frames, int width, int* __restrict__ start_out,
                                int* __restrict__ span_out, int* __restrict__ frames_out, int* __restrict__ nslices_out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int len = lengths ? lengths[row] : uniform_len;
    int start = peak ? min(max(peak[row], 0), len) : 0;          // waterfall.py:358-361
    int span = len - start;
    const int skip = min(max(ignore, 0), span);                  // :363-367
    start += skip;
    span -= skip;
    if (keep >= 0) span = min(span, keep);                       // :369-372
    int nf = (span >= n_fft) ? 1 + (span - n_fft) / hop : 0;     // :208
    nf = min(nf, max_frames);
    start_out[row] = start;
    span_out[row] = span;
    nslices_out[row] = counts[nf];
    for (int s = 0; s < width; ++s) frames_out[(long long)row * width + s] = table[(long long)nf * width + s];
}

}  // namespace

int launch_csd_plan(const int* d_lengths, int uniform_len, int n_rows, const int* d_peak, int ignore, int keep, int n_fft,
                    int hop, const int* d_table, const int* d_counts, int max_frames, int width, int* d_start_out,
                    int* d_span_out, int* d_frames_out, int* d_nslices_out, cudaStream_t stream) {
    AA_REQUIRE(d_table && d_counts && d_start_out && d_span_out && d_frames_out && d_nslices_out, AA_ERR_INVALID_ARGUMENT,
               "null argument");
    AA_REQUIRE(n_fft > 0 && hop > 0 && width >= 1 && width <= AA_MAX_SLICES && max_frames >= 0, AA_ERR_INVALID_ARGUMENT,
               "bad plan sizes");
    if (n_rows <= 0) return AA_OK;
    csd_plan_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, stream>>>(d_lengths, uniform_len, n_rows, d_peak, ignore, keep,
                                                                         n_fft, hop, d_table, d_counts, max_frames, width,
                                                                         d_start_out, d_span_out, d_frames_out, d_nslices_out);
    return launch_check("csd_plan_kernel");
}

int launch_csd_smooth(float* d_data, int n_rows, const int* d_n_slices, int max_slices, int n_bins, int k0, int n_in,
                      const double* d_f_in, const double* d_grid, int n_grid, const int* d_grid_seg,
                      const int* d_bin_seg, int window, int slice_max, double dynamic_range_db, cudaStream_t stream) {
    AA_REQUIRE(d_data && d_n_slices && d_f_in && d_grid && d_grid_seg && d_bin_seg, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(max_slices >= 1 && n_bins >= 1 && k0 >= 0 && n_in >= 1 && k0 + n_in <= n_bins && n_grid >= 2 && window >= 2,
               AA_ERR_INVALID_ARGUMENT, "bad smoothing tables");
    if (n_rows <= 0) return AA_OK;
    const size_t smem = sizeof(double) * ((size_t)n_in + 2 * (size_t)n_grid);
    AA_REQUIRE(smem <= 200 * 1024, AA_ERR_UNSUPPORTED, "smoothing tables do not fit shared memory");
    AA_CUDA_TRY(cudaFuncSetAttribute(csd_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SmoothArgs a{};
    a.data = d_data;
    a.n_slices = d_n_slices;
    a.max_slices = max_slices;
    a.n_bins = n_bins;
    a.k0 = k0;
    a.n_in = n_in;
    a.f_in = d_f_in;
    a.grid = d_grid;
    a.n_grid = n_grid;
    a.grid_seg = d_grid_seg;
    a.bin_seg = d_bin_seg;
    a.window = window;
    a.slice_max = slice_max ? 1 : 0;
    a.dyn = (float)dynamic_range_db;
    ProfScope prof("csd_smooth", stream);
    csd_smooth_kernel<<<(unsigned)n_rows, kSmoothThreads, smem, stream>>>(a);
    return launch_check("csd_smooth_kernel");
}

}  // namespace aa
