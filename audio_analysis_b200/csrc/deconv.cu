// Regularised FFT deconvolution of sweep captures into impulse responses
// (analyse/deconvolve.py:124-193), SURVEY.md 8(f) N1: the stage in front of the
// decay path in config 4.  H = Y conj(X) / (|X|^2 + eps) at n_fft = 2^p, built on
// the same FFT engine as the band filterbank; the spectral division, Hermitian
// packing and conjugation are fused into the first pass of the inverse transform.
#include <math.h>

#include "fft.cuh"
#include "kernels.cuh"

struct aa_deconv_plan {
    int n_fft = 0;
    int M = 0;
    aa::C2CPlan* c2c = nullptr;
    aa::TwoLevelTable twN;
};

namespace aa {

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

__global__ void untangle_rows_kernel(const float2* __restrict__ Z, float2* __restrict__ X, int M, int x_stride,
                                     TwoLevelTwiddle twN) {
    const long long row = blockIdx.y;
    const float2* __restrict__ z = Z + row * (long long)M;
    float2* __restrict__ o = X + row * (long long)x_stride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= M; k += gridDim.x * blockDim.x) {
        const int km = (k == M) ? 0 : k;
        const int mk = (k == 0) ? 0 : M - k;
        o[k] = untangle_half_spectrum(__ldg(z + km), __ldg(z + mk), twiddle2(twN, k));
    }
}

__device__ __forceinline__ float block_reduce_max(float v) {
    __shared__ float red[32];
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = red[0];
    for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmaxf(m, red[w]);
    return m;
}

__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double red[32];
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) s += red[w];
    return s;
}

// eps = rel * max(1e-30, max |X|^2)   (deconvolve.py:154-156); single CTA
__global__ void eps_kernel(const float2* __restrict__ X, int n, float rel, float* __restrict__ eps_out) {
    float m = 0.f;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        float2 v = X[k];
        m = fmaxf(m, v.x * v.x + v.y * v.y);
    }
    m = block_reduce_max(m);
    if (threadIdx.x == 0) *eps_out = rel * fmaxf(1e-30f, m);
}

// per row: subtract the mean of the first n_out samples (deconvolve.py:180-181), record max |h|
__global__ void dc_peak_kernel(float* __restrict__ h, long long stride, int n_out, int remove_dc,
                               float* __restrict__ peak_out) {
    float* __restrict__ r = h + (long long)blockIdx.x * stride;
    float mean = 0.f;
    if (remove_dc) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n_out; i += blockDim.x) s += (double)r[i];
        s = block_reduce_sum(s);
        mean = (float)(s / (double)n_out);
    }
    float pk = 0.f;
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
        float v = r[i] - mean;
        r[i] = v;
        pk = fmaxf(pk, fabsf(v));
    }
    pk = block_reduce_max(pk);
    if (threadIdx.x == 0) peak_out[blockIdx.x] = pk;
}

// peak-normalise groups of rows (the channels of one file) and copy to the output pitch (deconvolve.py:99-105,190)
__global__ void normalise_kernel(const float* __restrict__ h, long long stride, int n_out, const float* __restrict__ peaks,
                                 int group, float target, float* __restrict__ out, long long out_stride) {
    const int row = blockIdx.y;
    float scale = 1.0f;
    if (target > 0.f) {
        const int g0 = (row / group) * group;
        float pk = 0.f;
        for (int c = 0; c < group && g0 + c < (int)gridDim.y; ++c) pk = fmaxf(pk, peaks[g0 + c]);
        if (pk > 0.f) scale = (float)((double)target / (double)pk);
    }
    const float* __restrict__ r = h + (long long)row * stride;
    float* __restrict__ o = out + (long long)row * out_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) o[i] = r[i] * scale;
}

struct Layout {
    size_t off_Xs, off_eps, off_peaks, off_Z, off_Y, off_h, off_scr, total;
};

Layout make_layout(const aa_deconv_plan* p, long long rows) {
    Layout l{};
    size_t o = 0;
    const size_t M = (size_t)p->M;
    l.off_Xs = o;
    o += align_up(sizeof(float2) * (M + 1));
    l.off_eps = o;
    o += align_up(sizeof(float));
    l.off_peaks = o;
    o += align_up(sizeof(float) * (size_t)rows);
    l.off_Z = o;
    o += align_up(sizeof(float2) * M * (size_t)rows);
    l.off_Y = o;
    o += align_up(sizeof(float2) * (M + 1) * (size_t)rows);
    l.off_h = o;
    o += align_up(sizeof(float) * 2 * M * (size_t)rows);
    l.off_scr = o;
    o += align_up(sizeof(float2) * c2c_scratch_elems(p->c2c, rows));
    l.total = o;
    return l;
}

}  // namespace

int deconv_plan_create(int n_fft, aa_deconv_plan** out) {
    AA_REQUIRE(out != nullptr, AA_ERR_INVALID_ARGUMENT, "null plan pointer");
    AA_REQUIRE(n_fft >= 16 && (n_fft & (n_fft - 1)) == 0, AA_ERR_INVALID_ARGUMENT, "n_fft must be a power of two >= 16");
    AA_REQUIRE(n_fft <= (1 << 24), AA_ERR_TOO_LONG, "n_fft too large");
    aa_deconv_plan* p = new aa_deconv_plan();
    p->n_fft = n_fft;
    p->M = n_fft / 2;
    int rc = c2c_plan_create(p->M, &p->c2c);
    if (rc == AA_OK) rc = build_two_level(n_fft, &p->twN);
    if (rc != AA_OK) {
        if (p->c2c) c2c_plan_destroy(p->c2c);
        free_two_level(&p->twN);
        delete p;
        return rc;
    }
    *out = p;
    return AA_OK;
}

void deconv_plan_destroy(aa_deconv_plan* p) {
    if (!p) return;
    if (p->c2c) c2c_plan_destroy(p->c2c);
    free_two_level(&p->twN);
    delete p;
}

size_t deconv_workspace_bytes(const aa_deconv_plan* p, int rows) {
    if (!p || rows < 1) return 0;
    return make_layout(p, rows).total;
}

int deconvolve_batch(const aa_deconv_plan* plan, const float* d_sweep, int n_sweep, const float* d_rec,
                     long long rec_stride, int n_rec, int n_rows, double reg_rel, int remove_dc, double target_peak,
                     int norm_group, int n_out, float* d_ir_out, long long ir_stride, void* d_ws, size_t ws_bytes,
                     cudaStream_t stream) {
    AA_REQUIRE(plan && d_sweep && d_rec && d_ir_out && d_ws, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(n_sweep >= 8 && n_rec >= 8, AA_ERR_INVALID_ARGUMENT,
               "Recorded and sweep must both contain at least a few samples.");  // deconvolve.py:139
    AA_REQUIRE(n_sweep <= plan->n_fft && n_rec <= plan->n_fft, AA_ERR_INVALID_ARGUMENT, "n_fft smaller than the inputs");
    AA_REQUIRE(n_out >= 1 && n_out <= plan->n_fft && ir_stride >= n_out, AA_ERR_INVALID_ARGUMENT, "bad output length");
    AA_REQUIRE(norm_group >= 1, AA_ERR_INVALID_ARGUMENT, "norm_group must be >= 1");
    AA_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, AA_ERR_INVALID_ARGUMENT, "workspace not aligned");
    if (n_rows <= 0) return AA_OK;
    const Layout lay = make_layout(plan, n_rows);
    AA_REQUIRE(lay.total <= ws_bytes, AA_ERR_WORKSPACE, "workspace too small (see aa_deconv_workspace_bytes)");
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    float2* d_Xs = reinterpret_cast<float2*>(ws + lay.off_Xs);
    float* d_eps = reinterpret_cast<float*>(ws + lay.off_eps);
    float* d_peaks = reinterpret_cast<float*>(ws + lay.off_peaks);
    float2* d_Z = reinterpret_cast<float2*>(ws + lay.off_Z);
    float2* d_Y = reinterpret_cast<float2*>(ws + lay.off_Y);
    float* d_h = reinterpret_cast<float*>(ws + lay.off_h);
    float2* d_scr = reinterpret_cast<float2*>(ws + lay.off_scr);
    const int M = plan->M;
    const int gx = (M + 1 + 255) / 256 > 1024 ? 1024 : (M + 1 + 255) / 256;
    int rc;
    // sweep spectrum X and eps
    {
        SrcDesc s{};
        s.mode = SRC_REAL_PACKED;
        s.base = d_sweep;
        s.unit_stride = 0;
        s.n_valid = n_sweep;
        rc = c2c_exec(plan->c2c, s, make_dst_complex(d_Z, M, M, 0, 1.0f), 1, d_scr, stream);
        if (rc != AA_OK) return rc;
        untangle_rows_kernel<<<dim3((unsigned)gx, 1), 256, 0, stream>>>(d_Z, d_Xs, M, M + 1, plan->twN.dev);
        if ((rc = launch_check("untangle_rows_kernel")) != AA_OK) return rc;
        eps_kernel<<<1, 1024, 0, stream>>>(d_Xs, M + 1, (float)reg_rel, d_eps);
        if ((rc = launch_check("eps_kernel")) != AA_OK) return rc;
    }
    // recorded spectra Y
    {
        SrcDesc s{};
        s.mode = SRC_REAL_PACKED;
        s.base = d_rec;
        s.unit_stride = rec_stride;
        s.n_valid = n_rec;
        s.clip = 1;
        rc = c2c_exec(plan->c2c, s, make_dst_complex(d_Z, M, M, 0, 1.0f), n_rows, d_scr, stream);
        if (rc != AA_OK) return rc;
        untangle_rows_kernel<<<dim3((unsigned)gx, (unsigned)n_rows), 256, 0, stream>>>(d_Z, d_Y, M, M + 1, plan->twN.dev);
        if ((rc = launch_check("untangle_rows_kernel")) != AA_OK) return rc;
    }
    // h = irfft(Y conj(X) / (|X|^2 + eps))
    {
        SrcDesc s{};
        s.mode = SRC_DECONV_PACKED;
        s.base = d_Y;
        s.unit_stride = M + 1;
        s.n_valid = M;
        s.conj = 1;
        s.M = M;
        s.N = plan->n_fft;
        s.twN = plan->twN.dev;
        s.aux = d_Xs;
        s.eps_ptr = d_eps;
        rc = c2c_exec(plan->c2c, s, make_dst_complex(reinterpret_cast<float2*>(d_h), M, M, 1, 1.0f / (float)M), n_rows,
                      d_scr, stream);
        if (rc != AA_OK) return rc;
    }
    dc_peak_kernel<<<n_rows, 1024, 0, stream>>>(d_h, 2LL * M, n_out, remove_dc, d_peaks);
    if ((rc = launch_check("dc_peak_kernel")) != AA_OK) return rc;
    const int gn = (n_out + 255) / 256 > 1024 ? 1024 : (n_out + 255) / 256;
    normalise_kernel<<<dim3((unsigned)gn, (unsigned)n_rows), 256, 0, stream>>>(d_h, 2LL * M, n_out, d_peaks, norm_group,
                                                                              (float)target_peak, d_ir_out, ir_stride);
    return launch_check("normalise_kernel");
}

}  // namespace aa
