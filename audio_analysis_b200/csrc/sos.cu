// Butterworth second-order-section band filterbank with zero-phase forward-backward filtering, as a chunked
// parallel linear-recurrence scan (opt-in mode of rt60bands; BASELINE.json north_star stage 1).
//
// NOT the reference's filter: kianmcevoy/audio_analysis filters bands with a circular FFT-domain raised-cosine mask
// (analyse/rt60bands.py:116-175) and that is what the default, parity-graded path implements.  This
// mode is pinned to scipy.signal.sosfiltfilt instead (odd extension of 3 * ntaps samples at both ends, steady-state
// initial conditions scaled by the first sample, forward pass, reversed pass, trim), tests/test_gpu_parity.py.
//
// One biquad in the transposed direct form II scipy's sosfilt uses:
//     y[n] = b0 x[n] + z0[n-1],  z0[n] = b1 x[n] - a1 y[n] + z1[n-1],  z1[n] = b2 x[n] - a2 y[n]
// is the linear recurrence s[n] = A s[n-1] + B x[n] on the state s = (z0, z1) with the 2 x 2 transition matrix
// A = [[-a1, 1], [-a2, 0]] and B = (b1 - a1 b0, b2 - a2 b0).  A CTA owns one (row, band) unit and walks its extended
// signal in chunks of kThreads x kSeg samples; inside a chunk every thread owns kSeg consecutive samples and, per
// section: (1) runs the recurrence over its segment from the zero state, which yields the segment's forced response
// v; (2) the segments' incoming states follow from an affine scan with the CONSTANT matrix M = A^kSeg - Kogge-Stone
// over the warp with the precomputed powers M^1, M^2, .. M^16, then one serial step per warp with M^32 and the
// chunk's carried-in state; (3) the thread re-runs its segment from its true incoming state and overwrites the
// samples with the section's output, which is the next section's input.  All state arithmetic - segments, scan,
// carries between chunks - is float64, so poles at radius 1 - 6e-4 (the 39 Hz third-octave band) do not drift; samples
// are float32 in memory like everywhere else on the path.
#include <math.h>
#include <string.h>

#include <vector>

#include "aa_common.cuh"
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int kSosThreads = 256;
constexpr int kSeg = 16;                       // samples per thread and chunk
constexpr int kChunk = kSosThreads * kSeg;     // samples per CTA step
constexpr int kMaxSec = AA_SOS_MAX_SECTIONS;
constexpr int kPitch = kSeg + 1;               // shared-memory pitch of a segment: conflict-free stride

struct SosBandDev {
    int n_sec;
    int edge;                                  // odd extension at each end (scipy sosfiltfilt's padlen of this band)
    double b0[kMaxSec], b1[kMaxSec], b2[kMaxSec], a1[kMaxSec], a2[kMaxSec];
    double zi0[kMaxSec], zi1[kMaxSec];         // steady-state initial state for a unit step (scipy sosfilt_zi)
    double m_step[kMaxSec][5][4];              // M^(2^k), k = 0..4, M = A^kSeg (row major 2 x 2)
    double m_warp[kMaxSec][4];                 // M^32
    double m_lane[kMaxSec][32][4];             // M^(t+1), t = 0..31
};

struct SosArgs {
    const float* x;            // [rows][row_stride]
    long long row_stride;
    int n;                     // samples per row
    const SosBandDev* bands;
    int n_bands;
    float* fwd;                // [units][fwd_stride] forward-pass output over the extended signal
    long long fwd_stride;
    float* y;                  // [units][y_stride] band signals
    long long y_stride;
};

__device__ __forceinline__ void mat_vec_acc(const double* __restrict__ m, double v0, double v1, double& o0, double& o1) {
    o0 = fma(m[0], v0, fma(m[1], v1, o0));
    o1 = fma(m[2], v0, fma(m[3], v1, o1));
}

// scipy.signal._arraytools.odd_ext: 2 x[0] - x[edge..1], x, 2 x[n-1] - x[n-2 .. n-edge-1]
__device__ __forceinline__ float ext_sample(const float* __restrict__ x, int n, int edge, int i) {
    if (i < edge) return 2.0f * __ldg(x) - __ldg(x + (edge - i));
    const int j = i - edge;
    if (j < n) return __ldg(x + j);
    return 2.0f * __ldg(x + (n - 1)) - __ldg(x + (n - 2 - (j - n)));
}

// kBackward = false: input = odd extension of the row, output -> fwd.  true: input = fwd reversed, output (reversed
// back, trimmed by `edge` at both ends) -> y.
template <bool kBackward>
__global__ void __launch_bounds__(kSosThreads) sos_pass_kernel(SosArgs a) {
    __shared__ float stage[kSosThreads * kPitch];
    __shared__ double warp_tot[kSosThreads / 32][2];
    __shared__ double warp_in[kSosThreads / 32 + 1][2];
    __shared__ double carry[kMaxSec][2];        // state entering the current chunk, per section
    __shared__ float first_sample;
    const int unit = blockIdx.x;
    const int row = unit / a.n_bands;
    const SosBandDev& bd = a.bands[unit - row * a.n_bands];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int edge = bd.edge;
    const int ne = a.n + 2 * edge;
    const float* __restrict__ x = a.x + (long long)row * a.row_stride;
    float* __restrict__ fwd = a.fwd + (long long)unit * a.fwd_stride;
    float* __restrict__ y = a.y + (long long)unit * a.y_stride;
    auto input = [&](int i) -> float {
        if (kBackward) return fwd[ne - 1 - i];
        return ext_sample(x, a.n, edge, i);
    };
    // initial conditions: zi * first input sample (scipy sosfiltfilt)
    if (threadIdx.x == 0) first_sample = input(0);
    __syncthreads();
    if (threadIdx.x < bd.n_sec) {
        carry[threadIdx.x][0] = bd.zi0[threadIdx.x] * (double)first_sample;
        carry[threadIdx.x][1] = bd.zi1[threadIdx.x] * (double)first_sample;
    }
    __syncthreads();
    for (int c0 = 0; c0 < ne; c0 += kChunk) {
        // coalesced load of the chunk, then every thread takes its segment
        for (int e = threadIdx.x; e < kChunk; e += kSosThreads) {
            const int i = c0 + e;
            stage[(e / kSeg) * kPitch + (e % kSeg)] = (i < ne) ? input(i) : 0.0f;
        }
        __syncthreads();
        double u[kSeg];
#pragma unroll
        for (int s = 0; s < kSeg; ++s) u[s] = (double)stage[threadIdx.x * kPitch + s];
        for (int k = 0; k < bd.n_sec; ++k) {
            const double b0 = bd.b0[k], b1 = bd.b1[k], b2 = bd.b2[k], a1 = bd.a1[k], a2 = bd.a2[k];
            // (1) forced response of the segment: final state from the zero state
            double z0 = 0.0, z1 = 0.0;
#pragma unroll
            for (int s = 0; s < kSeg; ++s) {
                const double yy = fma(b0, u[s], z0);
                z0 = fma(b1, u[s], fma(-a1, yy, z1));
                z1 = fma(b2, u[s], -a2 * yy);
            }
            // (2) inclusive affine scan over the warp: P_t = sum_{q <= t} M^(t-q) v_q
            double p0 = z0, p1 = z1;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const double q0 = __shfl_up_sync(0xffffffffu, p0, 1 << d);
                const double q1 = __shfl_up_sync(0xffffffffu, p1, 1 << d);
                if (lane >= (1 << d)) mat_vec_acc(bd.m_step[k][d], q0, q1, p0, p1);
            }
            if (lane == 31) {
                warp_tot[warp][0] = p0;
                warp_tot[warp][1] = p1;
            }
            __syncthreads();
            if (threadIdx.x == 0) {   // state entering every warp; the last entry is the state leaving the chunk
                double w0 = carry[k][0], w1 = carry[k][1];
                for (int w = 0; w < kSosThreads / 32; ++w) {
                    warp_in[w][0] = w0;
                    warp_in[w][1] = w1;
                    double n0 = warp_tot[w][0], n1 = warp_tot[w][1];
                    mat_vec_acc(bd.m_warp[k], w0, w1, n0, n1);
                    w0 = n0;
                    w1 = n1;
                }
                carry[k][0] = w0;
                carry[k][1] = w1;
            }
            __syncthreads();
            // state after this thread's segment = P_t + M^(t+1) W_in; the state entering it is the lane below's
            {
                const double w0 = warp_in[warp][0], w1 = warp_in[warp][1];
                mat_vec_acc(bd.m_lane[k][lane], w0, w1, p0, p1);
                double s0 = __shfl_up_sync(0xffffffffu, p0, 1);
                double s1 = __shfl_up_sync(0xffffffffu, p1, 1);
                if (lane == 0) {
                    s0 = w0;
                    s1 = w1;
                }
                z0 = s0;
                z1 = s1;
            }
            // (3) the section's output over the segment, from the true incoming state
#pragma unroll
            for (int s = 0; s < kSeg; ++s) {
                const double yy = fma(b0, u[s], z0);
                z0 = fma(b1, u[s], fma(-a1, yy, z1));
                z1 = fma(b2, u[s], -a2 * yy);
                u[s] = yy;
            }
        }
#pragma unroll
        for (int s = 0; s < kSeg; ++s) stage[threadIdx.x * kPitch + s] = (float)u[s];
        __syncthreads();
        for (int e = threadIdx.x; e < kChunk; e += kSosThreads) {
            const int i = c0 + e;
            if (i >= ne) break;
            const float v = stage[(e / kSeg) * kPitch + (e % kSeg)];
            if (kBackward) {
                const int pos = ne - 1 - i - edge;   // back to forward time, extension trimmed
                if (pos >= 0 && pos < a.n) y[pos] = v;
            } else {
                fwd[i] = v;
            }
        }
        __syncthreads();
    }
}

void mat_mul(const double* a, const double* b, double* o) {
    double r[4] = {a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3]};
    memcpy(o, r, sizeof(r));
}

}  // namespace

size_t sos_workspace_bytes(int n, int n_sec_max, int rows, int nb) {
    const size_t units = (size_t)rows * (size_t)nb;
    const size_t ne = (size_t)n + 2 * (size_t)(3 * (2 * n_sec_max + 1));
    const size_t fwd = ((ne + 3) / 4) * 4, ys = ((size_t)n + 3) / 4 * 4;
    return 4096 + sizeof(SosBandDev) * AA_MAX_BANDS + sizeof(float) * units * (fwd + ys);
}

// sos: HOST [nb][n_sec][6] (b0 b1 b2 1 a1 a2), zi: HOST [nb][n_sec][2], edges: HOST [nb] extension per band
int rt60bands_sos_batch(const float* d_x, long long row_stride, int n, int n_rows, const int* d_start_idx,
                        const double* sos, const double* zi, const int* edges, int nb, int n_sec, const aa_decay_params* p,
                        double* d_metrics_out, float* d_band_signal_out, void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    AA_REQUIRE(d_x && d_start_idx && sos && zi && edges && p && d_metrics_out && d_ws, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(nb >= 1 && nb <= AA_MAX_BANDS && n_sec >= 1 && n_sec <= kMaxSec, AA_ERR_INVALID_ARGUMENT,
               "band / section count out of range");
    AA_REQUIRE(row_stride >= n && row_stride % 4 == 0, AA_ERR_INVALID_ARGUMENT, "row_stride must be >= n and a multiple of 4");
    AA_REQUIRE(p->smoothing_window <= 1, AA_ERR_UNSUPPORTED, "EDC smoothing is not available in the Butterworth mode");
    if (n_rows <= 0) return AA_OK;
    std::vector<SosBandDev> hb((size_t)nb);
    int edge = 0;
    for (int b = 0; b < nb; ++b) {
        SosBandDev& d = hb[(size_t)b];
        memset(&d, 0, sizeof(d));
        d.n_sec = n_sec;
        const double* sb = sos + (size_t)b * n_sec * 6;
        d.edge = edges[b];
        AA_REQUIRE(d.edge >= 0 && d.edge <= 3 * (2 * kMaxSec + 1), AA_ERR_INVALID_ARGUMENT, "extension length out of range");
        if (d.edge > edge) edge = d.edge;
        for (int k = 0; k < n_sec; ++k) {
            const double* s = sb + 6 * k;
            AA_REQUIRE(s[3] == 1.0, AA_ERR_INVALID_ARGUMENT, "sections must be normalised (a0 = 1)");
            d.b0[k] = s[0];
            d.b1[k] = s[1];
            d.b2[k] = s[2];
            d.a1[k] = s[4];
            d.a2[k] = s[5];
            d.zi0[k] = zi[((size_t)b * n_sec + k) * 2];
            d.zi1[k] = zi[((size_t)b * n_sec + k) * 2 + 1];
            const double A[4] = {-s[4], 1.0, -s[5], 0.0};
            double M[4] = {1.0, 0.0, 0.0, 1.0};
            for (int i = 0; i < kSeg; ++i) mat_mul(A, M, M);          // M = A^kSeg
            double pw[4];
            memcpy(pw, M, sizeof(pw));
            for (int q = 0; q < 5; ++q) {                             // M^(2^q)
                memcpy(d.m_step[k][q], pw, sizeof(pw));
                mat_mul(pw, pw, pw);
            }
            memcpy(d.m_warp[k], pw, sizeof(pw));                      // M^32
            double acc[4];
            memcpy(acc, M, sizeof(acc));
            for (int t = 0; t < 32; ++t) {                            // M^(t+1)
                memcpy(d.m_lane[k][t], acc, sizeof(acc));
                mat_mul(M, acc, acc);
            }
        }
    }
    AA_REQUIRE(n > edge + 1, AA_ERR_INVALID_ARGUMENT, "The length of the input vector x must be greater than padlen.");
    const size_t ne = (size_t)n + 2 * (size_t)edge;
    const size_t fwd_stride = ((ne + 3) / 4) * 4, y_stride = ((size_t)n + 3) / 4 * 4;
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    SosBandDev* d_bands = reinterpret_cast<SosBandDev*>(ws);
    const size_t off = ((sizeof(SosBandDev) * (size_t)nb + 4095) / 4096) * 4096;
    AA_REQUIRE(ws_bytes > off, AA_ERR_WORKSPACE, "workspace too small");
    const size_t per_row = sizeof(float) * (size_t)nb * (fwd_stride + y_stride);
    long long rc_rows = (long long)((ws_bytes - off) / per_row);
    AA_REQUIRE(rc_rows >= 1, AA_ERR_WORKSPACE, "workspace too small for one row (see aa_sos_workspace_bytes)");
    if (rc_rows > n_rows) rc_rows = n_rows;
    AA_CUDA_TRY(cudaMemcpyAsync(d_bands, hb.data(), sizeof(SosBandDev) * (size_t)nb, cudaMemcpyHostToDevice, stream));
    AA_CUDA_TRY(cudaStreamSynchronize(stream));   // hb is a stack-lifetime vector
    aa_decay_params pk = *p;
    pk.smoothing_window = 0;
    for (long long r0 = 0; r0 < n_rows; r0 += rc_rows) {
        const long long rows = (n_rows - r0 < rc_rows) ? (n_rows - r0) : rc_rows;
        const long long units = rows * nb;
        SosArgs a{};
        a.x = d_x + r0 * row_stride;
        a.row_stride = row_stride;
        a.n = n;
        a.bands = d_bands;
        a.n_bands = nb;
        a.fwd = reinterpret_cast<float*>(ws + off);
        a.fwd_stride = (long long)fwd_stride;
        a.y = a.fwd + (size_t)units * fwd_stride;
        a.y_stride = (long long)y_stride;
        {
            ProfScope prof("sos_filtfilt", stream);
            sos_pass_kernel<false><<<(unsigned)units, kSosThreads, 0, stream>>>(a);
            int rc = launch_check("sos_pass_kernel<forward>");
            if (rc != AA_OK) return rc;
            sos_pass_kernel<true><<<(unsigned)units, kSosThreads, 0, stream>>>(a);
            rc = launch_check("sos_pass_kernel<backward>");
            if (rc != AA_OK) return rc;
        }
        if (d_band_signal_out)
            AA_CUDA_TRY(cudaMemcpy2DAsync(d_band_signal_out + r0 * nb * (long long)n, sizeof(float) * (size_t)n, a.y,
                                          sizeof(float) * y_stride, sizeof(float) * (size_t)n, (size_t)units,
                                          cudaMemcpyDeviceToDevice, stream));
        // band EDC + fits with the row's shared start index (rt60bands.py:342,393), metrics only
        ProfScope prof("band_edc", stream);
        int rc = launch_edc(a.y, (long long)y_stride, nullptr, n, n, (int)units, &pk, d_start_idx + r0, nb, 8, nullptr, 0,
                            nullptr, nullptr, d_metrics_out + r0 * nb * AA_MAX_FITS, nullptr, nullptr, stream);
        if (rc != AA_OK) return rc;
    }
    return AA_OK;
}

}  // namespace aa
