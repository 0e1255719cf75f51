// Band EDC + crossings + decay fits from per-tile energy partials (metrics only).
//
// The register-pass rows kernel (fft.cu: fft_rows_reg<.., true>) stores, next to a band signal y, the energy of
// every group of 8 rows of its output matrix.  A band signal therefore arrives with its energy already summed over
// "tiles" of TL = 2 M1 consecutive samples (M2 tiles, M1 / 8 partials each), and the Schroeder integral
// (analyse/decay.py:150-158) is a suffix scan over those M2 numbers.  Only the few tiles that hold the start
// sample, a level crossing (decay.py:173-199) or fit points (decay.py:224-249) are read again at sample
// resolution - about a tenth of the signal - instead of streaming it twice as the general kernel (edc.cu) does.
//
// One CTA of 2 warps per (row, band) unit (kFitThreads).  Semantics are those of edc.cu / edc_body.cuh: float32 squares and
// in-tile sums, float64 carries, float32 dB curve and times, first sample <= target, fit mask on float32 times,
// closed-form least squares in float64 (fit_math.cuh).
#include "bandfit_body.cuh"

namespace aa {

namespace {

using namespace bandfit;

#ifndef AA_FIT_THREADS
#define AA_FIT_THREADS 64
#endif
constexpr int kFitThreads = AA_FIT_THREADS;
constexpr int kFitWarps = kFitThreads / 32;

template <int NS>
// sixteen resident CTAs of two warps per SM: the kernel is a
// chain of short phases separated by barriers, and more, smaller CTAs in flight hide them - per 8192 rows x 26 bands
// 8.05 ms (256 threads, 3 CTAs) -> 6.83 (256 x 5) -> 4.29 ms (128 x 10); with the certification phases of round 2:
// 8.84 ms (256 x 4), 6.97 (128 x 10), 6.62 ms (64 x 16), 8.05 (32 x 32)
#ifndef AA_FIT_MIN_BLOCKS
#define AA_FIT_MIN_BLOCKS 16
#endif
__global__ void __launch_bounds__(kFitThreads, AA_FIT_MIN_BLOCKS) band_fit_kernel(BandFitArgs a) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ FitShared<kFitWarps> sh;
    const int unit = blockIdx.x;
    const int um = a.unit_map ? __ldg(a.unit_map + unit) : unit;
    const int row = um / a.nbc, cj = um - row * a.nbc;
    band_fit_unit<NS, kFitThreads, false>(a, row, cj, a.y + (long long)unit * a.y_stride,
                                          a.part + (long long)unit * a.part_stride, dyn_smem, sh);
}

}  // namespace

// fills the kernel argument block shared by band_fit_kernel and band_fused_kernel (bandfused.cu)
int fill_band_fit_args(bandfit::BandFitArgs* out, int n, const aa_decay_params* p, const int* d_start_idx, int nbc, int out_nb,
                       const int* d_unit_map, const BandFitCert* cert, int min_samples, double* d_rt60_out, int decim,
                       int full_n, int aligned) {
    bandfit::BandFitArgs& a = *out;
    AA_REQUIRE(p && d_start_idx && d_rt60_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(p->n_fits >= 0 && p->n_fits <= AA_MAX_FITS, AA_ERR_INVALID_ARGUMENT, "n_fits out of range");
    int rc = edc::fill_edc_args(p, &a.e);
    if (rc != AA_OK) return rc;
    for (int tg = 0; tg < bandfit::kMaxTargets; ++tg) a.gain10[tg] = (tg < 2 + 2 * p->n_fits) ? exp10(-0.1 * a.e.target[tg]) : 0.0;
    a.n = n;
    AA_REQUIRE(nbc >= 1 && nbc <= AA_MAX_BANDS && out_nb >= 1, AA_ERR_INVALID_ARGUMENT, "band counts out of range");
    a.e.start_in = d_start_idx;
    a.e.start_div = nbc;
    a.e.rt60_out = d_rt60_out;
    a.unit_map = d_unit_map;
    a.nbc = nbc;
    a.out_nb = out_nb;
    if (cert) {
        a.cert = *cert;
    } else {
        AA_REQUIRE(nbc == out_nb, AA_ERR_INVALID_ARGUMENT, "a class needs its band map");
        for (int j = 0; j < nbc; ++j) a.cert.out_band[j] = (short)j;
    }
    AA_REQUIRE(!a.cert.on || (a.cert.counts && a.cert.lists), AA_ERR_INVALID_ARGUMENT, "certification without unit lists");
    a.e.min_samples = min_samples;
    a.decim = decim > 1 ? decim : 1;
    a.inv_decim = 1.0 / (double)a.decim;
    a.full_n = full_n;
    a.aligned = (aligned || decim > 1) ? 1 : 0;
    a.kappa = decim > 1 ? (float)(decim - 1) / (float)(2 * decim) : 0.0f;
    a.min_fit_points = decim > 1 ? (8 + decim - 1) / decim : 8;
    if (decim > 1) {
        AA_REQUIRE(p->sample_rate_hz % decim == 0, AA_ERR_INVALID_ARGUMENT, "sample rate not divisible by the decimation");
        a.e.sr = p->sample_rate_hz / decim;
    }
    return AA_OK;
}

int launch_band_fit(const float* d_y, long long y_stride, const float* d_part, long long part_stride, int n_groups,
                    int n_tiles, int tile_len, int n, long long n_units, const aa_decay_params* p,
                    const int* d_start_idx, int nbc, int out_nb, const int* d_unit_map, const BandFitCert* cert,
                    int min_samples, double* d_rt60_out, cudaStream_t stream, int decim, int full_n, int aligned) {
    AA_REQUIRE(d_y && d_part && p && d_start_idx && d_rt60_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(tile_len % 4 == 0 && tile_len >= 4 && tile_len <= kSub * kMaxSub, AA_ERR_UNSUPPORTED,
               "tile length must be a multiple of 4 and at most 1024");
    AA_REQUIRE((long long)n_tiles * tile_len >= n && y_stride % 4 == 0, AA_ERR_INVALID_ARGUMENT, "tiles do not cover the row");
    AA_REQUIRE(n_units < 2147483647LL, AA_ERR_INVALID_ARGUMENT, "too many band units");
    if (n_units <= 0) return AA_OK;
    BandFitArgs a{};
    int rc = fill_band_fit_args(&a, n, p, d_start_idx, nbc, out_nb, d_unit_map, cert, min_samples, d_rt60_out, decim, full_n,
                                aligned);
    if (rc != AA_OK) return rc;
    a.y = d_y;
    a.y_stride = y_stride;
    a.part = d_part;
    a.part_stride = part_stride;
    a.n_groups = n_groups;
    a.n_tiles = n_tiles;
    a.tile_len = tile_len;
    const size_t smem = (size_t)n_tiles * (sizeof(double) + sizeof(float));
    AA_REQUIRE(smem <= 48 * 1024, AA_ERR_UNSUPPORTED, "too many tiles");
    const unsigned grid = (unsigned)n_units;
    switch ((tile_len + kSub - 1) / kSub) {
        case 1: band_fit_kernel<1><<<grid, kFitThreads, smem, stream>>>(a); break;
        case 2: band_fit_kernel<2><<<grid, kFitThreads, smem, stream>>>(a); break;
        case 3: band_fit_kernel<3><<<grid, kFitThreads, smem, stream>>>(a); break;
        case 4: band_fit_kernel<4><<<grid, kFitThreads, smem, stream>>>(a); break;
        case 5: band_fit_kernel<5><<<grid, kFitThreads, smem, stream>>>(a); break;
        case 6: band_fit_kernel<6><<<grid, kFitThreads, smem, stream>>>(a); break;
        case 7: band_fit_kernel<7><<<grid, kFitThreads, smem, stream>>>(a); break;
        default: band_fit_kernel<8><<<grid, kFitThreads, smem, stream>>>(a); break;
    }
    return launch_check("band_fit_kernel");
}

}  // namespace aa
