// Fused Schroeder EDC + level crossings + least-squares decay fits (sm_100a).
//
// Replaces analyse/decay.py:115-329 for a batch of rows.  One thread-block
// CLUSTER analyses one row:
//
//   phase 1  every warp streams its span of 128-sample tiles once from HBM:
//            |x| arg-max (decay.py:136) and a float32 tile energy sum per tile;
//            a per-CTA float64 suffix scan of the tile sums gives the chunk
//            carries, and the CTAs exchange (peak, energy) through distributed
//            shared memory.
//   resolve  E0 = E[start]; for every dB target (0, -10, fit range ends) the
//            tile whose boundary values straddle the target is found from the
//            carries alone and the first crossing sample inside it is resolved
//            exactly (decay.py:173-199); fit index ranges follow (decay.py:227).
//   phase 2  the span is read again (L2 hit: a cluster keeps only one row in
//            flight), the reverse prefix sum is rebuilt per tile with warp
//            shuffles, converted to dB, stored with 128-bit streaming stores,
//            and the regression sums of every fit are accumulated in float64
//            in the same pass.  In metrics-only mode (band signals of
//            rt60bands.py:272-321) only tiles inside a fit range are touched.
//
// Numerics: squares and in-tile sums are float32 (error relative to the tile
// sum, which never exceeds the running tail energy), carries across tiles are
// float64, the dB curve is float32 like the reference's stored curve.  Phase 1
// and phase 2 share ONE scan routine so boundary values are bit-identical and
// the curve is monotone across tile boundaries.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "aa_common.cuh"
#include "edc_body.cuh"
#include "fit_math.cuh"

namespace cg = cooperative_groups;

namespace aa {

namespace {

using namespace edc;
using namespace edc::base;

template <bool kWriteDb>
__global__ void __launch_bounds__(kThreads, 1024 / kThreads) edc_kernel(EdcArgs a) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ CtaShared sh;
    const int T = (int)cg::this_cluster().num_blocks();
    const int row = blockIdx.x / T;
    edc_row<kWriteDb, false>(a, row, a.x + (long long)row * a.row_stride, dyn_smem, sh);
}

// ---------------------------------------------------------------------------------------------
// peak index only (standalone rt60bands / waterfall callers)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_kernel(const float* __restrict__ xs, long long row_stride,
                                                   const int* __restrict__ lengths, int uniform_len,
                                                   int* __restrict__ out) {
    const int row = blockIdx.x;
    const int len = lengths ? lengths[row] : uniform_len;
    const float* __restrict__ x = xs + (long long)row * row_stride;
    float amax = -1.0f;
    int aidx = 0x7fffffff;
    const int nv = len >> 2;
    for (int q = threadIdx.x; q < nv; q += blockDim.x) {
        float4 v = ldg_f4(x + 4 * q);
        const int i0 = 4 * q;
        float m;
        m = fabsf(v.x); if (m > amax) { amax = m; aidx = i0; }
        m = fabsf(v.y); if (m > amax) { amax = m; aidx = i0 + 1; }
        m = fabsf(v.z); if (m > amax) { amax = m; aidx = i0 + 2; }
        m = fabsf(v.w); if (m > amax) { amax = m; aidx = i0 + 3; }
    }
    for (int i = 4 * nv + threadIdx.x; i < len; i += blockDim.x) {
        float m = fabsf(__ldg(x + i));
        if (m > amax) { amax = m; aidx = i; }
    }
    __shared__ float sm[8];
    __shared__ int si[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        float om = __shfl_xor_sync(kFullMask, amax, d);
        int oi = __shfl_xor_sync(kFullMask, aidx, d);
        if (om > amax || (om == amax && oi < aidx)) { amax = om; aidx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = amax; si[threadIdx.x >> 5] = aidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sm[w] > amax || (sm[w] == amax && si[w] < aidx)) { amax = sm[w]; aidx = si[w]; }
        out[row] = (aidx == 0x7fffffff) ? 0 : aidx;
    }
}

int pick_cluster(int max_len) {
    if (const char* env = getenv("AA_EDC_CLUSTER")) {  // tuning override: 1, 2, 4 or 8 CTAs per row
        int t = atoi(env);
        if (t == 1 || t == 2 || t == 4 || t == 8) return t;
    }
    if (max_len <= 16384) return 1;
    if (max_len <= 65536) return 2;
    if (max_len <= 131072) return 4;
    return kMaxCluster;
}

// Many rows: one CTA per row keeps every SM streaming (no cluster barriers on the critical path);
// few rows: split each row over a cluster so the whole chip works on them and the row stays in L2.
int pick_cluster_for_batch(int max_len, int n_rows) {
    if (getenv("AA_EDC_CLUSTER")) return pick_cluster(max_len);
    const int t = pick_cluster(max_len);
    if (n_rows >= 2 * 148) {
        const size_t dyn1 = (size_t)((max_len + kTile - 1) / kTile + kWarps) * (sizeof(double) + sizeof(float));
        if (dyn1 <= 100 * 1024) return 1;
    }
    return t;
}

}  // namespace

// edc_staged.cu: the same kernel body with the row kept in the cluster's shared memory between the phases
int launch_edc_staged(const edc::EdcArgs& a, int n_rows, int max_len, bool write_db, cudaStream_t stream, bool* launched);

// Shared launcher: decay (full rows of fits) and band metrics (rt60 table) both come here.
int launch_edc(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int max_len, int n_rows,
               const aa_decay_params* p, const int* d_start_in, int start_div, int min_samples, float* d_edc_out,
               long long edc_stride, int* d_start_out, double* d_fits_out, double* d_rt60_out, double* d_early10_out,
               int* d_status_out, cudaStream_t stream) {
    AA_REQUIRE(p != nullptr && d_x != nullptr, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(n_rows >= 0, AA_ERR_INVALID_ARGUMENT, "negative row count");
    if (n_rows == 0) return AA_OK;
    AA_REQUIRE(max_len >= 0 && max_len <= AA_MAX_ROW_SAMPLES, AA_ERR_TOO_LONG, "row longer than AA_MAX_ROW_SAMPLES");
    AA_REQUIRE(row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0, AA_ERR_INVALID_ARGUMENT,
               "rows must be 16-byte aligned (row_stride % 4 == 0)");
    AA_REQUIRE(p->n_fits >= 0 && p->n_fits <= AA_MAX_FITS, AA_ERR_INVALID_ARGUMENT, "n_fits out of range");
    AA_REQUIRE(p->sample_rate_hz > 0, AA_ERR_INVALID_ARGUMENT, "sample rate must be positive");
    AA_REQUIRE(p->smoothing_window <= 1, AA_ERR_UNSUPPORTED,
               "EDC smoothing is handled by the host layer, not by the fused kernel");
    if (d_edc_out)
        AA_REQUIRE(edc_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_edc_out) & 15) == 0,
                   AA_ERR_INVALID_ARGUMENT, "edc rows must be 16-byte aligned");

    EdcArgs a{};
    {
        int rc = fill_edc_args(p, &a);
        if (rc != AA_OK) return rc;
    }
    a.x = d_x;
    a.row_stride = row_stride;
    a.lengths = d_lengths;
    a.uniform_len = uniform_len;
    a.max_len = max_len;
    a.start_in = d_start_in;
    a.start_div = start_div > 0 ? start_div : 1;
    a.edc_out = d_edc_out;
    a.edc_stride = edc_stride;
    a.start_out = d_start_out;
    a.fits_out = d_fits_out;
    a.rt60_out = d_rt60_out;
    a.early10_out = d_early10_out;
    a.status_out = d_status_out;
    a.ignore = d_start_in ? 0 : (p->ignore_leading_samples > 0 ? p->ignore_leading_samples : 0);
    a.min_samples = min_samples;

    {   // large batches of rows that fit a cluster's shared memory: read the row from HBM once
        bool launched = false;
        int rc = launch_edc_staged(a, n_rows, max_len, d_edc_out != nullptr, stream, &launched);
        if (rc != AA_OK || launched) return rc;
    }
    const int T = pick_cluster_for_batch(max_len, n_rows);
    const int nt_max = (max_len + kTile - 1) / kTile;
    int tpw = (nt_max + T * kWarps - 1) / (T * kWarps);
    if (tpw < 1) tpw = 1;
    a.tiles_per_warp = tpw;
    const size_t dyn = (size_t)tpw * kWarps * (sizeof(double) + sizeof(float));
    AA_REQUIRE(dyn <= 200 * 1024, AA_ERR_TOO_LONG, "row too long for the EDC kernel's carry table");

    auto kern = d_edc_out ? edc_kernel<true> : edc_kernel<false>;
    AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)n_rows * T));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)T;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return AA_OK;
}

int launch_peak(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int n_rows, int* d_out,
                cudaStream_t stream) {
    AA_REQUIRE(d_x && d_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0, AA_ERR_INVALID_ARGUMENT,
               "rows must be 16-byte aligned (row_stride % 4 == 0)");
    if (n_rows <= 0) return AA_OK;
    peak_kernel<<<n_rows, 256, 0, stream>>>(d_x, row_stride, d_lengths, uniform_len, d_out);
    return launch_check("peak_kernel");
}

}  // namespace aa
