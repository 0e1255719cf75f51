// One kernel per (row, band) unit of a deep decimation class: masked spectrum -> band signal -> Schroeder integral ->
// crossings -> decay fits -> certification, all inside one CTA's shared memory (sm_100a: 227 KB per SM).
//
// Replaces, for the classes whose transform fits (M = n / 2D <= 12,000 complex points), the three kernels
// fft_cols_reg<band> -> fft_rows_reg<energy> -> band_fit_kernel and their traffic: the first pass's scratch matrix
// (written and read once), the band signal (written once, ~10 % read back) and the tile partials never leave the
// SM; HBM sees the band's bins of the row spectrum and three float64 results per unit.
// Reference semantics: analyse/rt60bands.py:170-175 (masked inverse), :272-321 (metrics from the band signal),
// analyse/decay.py:150-249 - on the decimated grid of bands.cu, with the certification of bandfit_body.cuh deciding
// per unit whether the result stands or the unit is redone on a finer level.
//
// Transform: fft_fused.cuh (three in-place stages, butterflies in registers).  Fits: bandfit_body.cuh, the body
// band_fit_kernel runs, reading the signal from shared memory in tiles of 128 samples.
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "bandfit_body.cuh"
#include "fft.cuh"
#include "fft_fused.cuh"

namespace aa {

namespace {

using namespace bandfit;

struct FusedArgs {
    const float2* X;       // half spectra of the rows (time-advanced to the analysis start: bands.cu)
    long long x_stride;
    const BandDev* bands;  // the class's bands (zone placement, mask tables)
    double freq_scale;
    TwoLevelTwiddle twN;   // exp(-2 pi i k / n_c), n_c = 2 M
    const float2* tw;      // stage twiddles (FusedPlan::tw_offset)
    float scale;
};

constexpr int fused_tiles(int M) { return (2 * M + kSub - 1) / kSub; }
constexpr size_t fused_smem(int M) { return sizeof(float2) * (size_t)M + (size_t)fused_tiles(M) * (sizeof(double) + sizeof(float)); }

template <class P, int kMinBlocks>
__global__ void __launch_bounds__(P::kThreads, kMinBlocks) band_fused_kernel(FusedArgs f, BandFitArgs a) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ FitShared<P::kThreads / 32> sh;
    float2* data = reinterpret_cast<float2*>(dyn_smem);
    const int unit = blockIdx.x;
    const int um = a.unit_map ? __ldg(a.unit_map + unit) : unit;
    const int row = um / a.nbc, cj = um - row * a.nbc;
    const int t = (int)threadIdx.x;
    RegSrcBand s;
    s.X = f.X + (size_t)row * f.x_stride;
    s.b = f.bands[cj];
    s.freq_scale = f.freq_scale;
    s.M = P::M;
    s.twN = f.twN;
    const FusedSupport sup = fused_support<P>(s);
    FusedPrefetch pre;
    fused_scatter_load<P>(t, s, sup, &pre);   // in flight while the buffer is cleared
    fused_clear<P>(t, data);
    __syncthreads();
    fused_scatter<P>(t, s, sup, pre, data);
    __syncthreads();
    fused_stage1<P>(t, data);
    __syncthreads();
    fused_stage<P, 2>(t, data, f.tw, 1, f.scale);
    __syncthreads();
    if constexpr (P::kStages >= 3) {
        fused_stage<P, 3>(t, data, f.tw, 1, f.scale);
        __syncthreads();
    }
    if constexpr (P::kStages >= 4) {
        fused_stage<P, 4>(t, data, f.tw, 1, f.scale);
        __syncthreads();
    }
    band_fit_unit<1, P::kThreads, true>(a, row, cj, reinterpret_cast<const float*>(data), nullptr,
                                        dyn_smem + sizeof(float2) * (size_t)P::M, sh);
}

struct FusedTables {
    int M = 0, device = -1, total = 0;
    float2* d = nullptr;   // per stage j >= 2: exp(-2 pi i k' / L_j), k' < L_{j-1}
};
std::mutex g_tables_mu;
FusedTables g_tables[16];

template <class P>
int fused_tables(const float2** tw) {
    int device = 0;
    AA_CUDA_TRY(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(g_tables_mu);
    FusedTables* slot = nullptr;
    for (FusedTables& ft : g_tables) {
        if (ft.d && ft.M == P::M && ft.device == device && ft.total == P::kTwTotal) {
            *tw = ft.d;
            return AA_OK;
        }
        if (!ft.d && !slot) slot = &ft;
    }
    AA_REQUIRE(slot != nullptr, AA_ERR_UNSUPPORTED, "too many fused band plans");
    std::vector<float2> h((size_t)P::kTwTotal);
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 2; j <= P::kStages; ++j)
        for (int k = 0; k < P::len(j - 1); ++k)
            h[(size_t)(P::tw_offset(j) + k)] = make_float2((float)cos(two_pi * k / P::len(j)), (float)-sin(two_pi * k / P::len(j)));
    float2* d = nullptr;
    AA_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(float2) * h.size()));
    // synchronous copy from pageable memory: done once per length and device, before the first launch reads it
    if (cudaMemcpy(d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        cudaGetLastError();
        set_error("upload of the fused band tables failed");
        return AA_ERR_CUDA;
    }
    slot->M = P::M;
    slot->device = device;
    slot->total = P::kTwTotal;
    slot->d = d;
    *tw = d;
    return AA_OK;
}

template <class P, int kMinBlocks>
int launch_fused(FusedArgs f, const BandFitArgs& a, long long n_units, cudaStream_t stream) {
    int rc = fused_tables<P>(&f.tw);
    if (rc != AA_OK) return rc;
    auto kern = band_fused_kernel<P, kMinBlocks>;
    size_t smem = fused_smem(P::M);
    if (const char* env = getenv("AA_FUSED_PAD_SMEM")) smem += (size_t)atoi(env);   // tuning: fewer resident CTAs per SM
    AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)n_units, P::kThreads, smem, stream>>>(f, a);
    return launch_check("band_fused_kernel");
}

}  // namespace

bool band_fused_supported(int M) {
    if (getenv("AA_NO_FUSED_BANDS") != nullptr) return false;   // A/B switch: the three-kernel path
#define X(m, th, mb, ...) \
    if (M == (m)) return true;
    AA_FUSED_LIST(X)
#undef X
    return false;
}

int launch_band_fused(int M, const float2* d_X, long long x_stride, const BandDev* d_bands, double freq_scale,
                      TwoLevelTwiddle twN, float scale, long long n_units, const aa_decay_params* p,
                      const int* d_start_idx, int nbc, int out_nb, const int* d_unit_map, const BandFitCert* cert,
                      int min_samples, double* d_rt60_out, cudaStream_t stream, int decim, int full_n) {
    AA_REQUIRE(d_X && d_bands, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(n_units < 2147483647LL, AA_ERR_INVALID_ARGUMENT, "too many band units");
    if (n_units <= 0) return AA_OK;
    BandFitArgs a{};
    int rc = fill_band_fit_args(&a, 2 * M, p, d_start_idx, nbc, out_nb, d_unit_map, cert, min_samples, d_rt60_out, decim,
                                full_n, 1);
    if (rc != AA_OK) return rc;
    a.n_tiles = fused_tiles(M);
    a.tile_len = kSub;
    FusedArgs f{};
    f.X = d_X;
    f.x_stride = x_stride;
    f.bands = d_bands;
    f.freq_scale = freq_scale;
    f.twN = twN;
    f.scale = scale;
#define X(m, th, mb, ...)                                                         \
    static_assert(FusedPlan<th, __VA_ARGS__>::M == (m), "radices do not multiply to M"); \
    if (M == (m)) return launch_fused<FusedPlan<th, __VA_ARGS__>, mb>(f, a, n_units, stream);
    AA_FUSED_LIST(X)
#undef X
    set_error("no fused band plan for M = %d", M);
    return AA_ERR_UNSUPPORTED;
}

}  // namespace aa
