// Band-limited RT60 (analyse/rt60bands.py:324-413) for a batch of equal-length rows.
//
// Per chunk of rows:
//   1. one forward length-N real FFT per row (the reference recomputes it for every band,
//      rt60bands.py:172) -> half spectrum X kept in the workspace;
//   2. one inverse per (row, band): the raised-cosine mask (rt60bands.py:116-167), the
//      Hermitian packing and the conjugation are generated inside the first FFT pass while
//      it loads its columns, so the masked spectrum never exists in memory;
//   3. the fused EDC kernel in metrics-only mode on every band signal with the row's shared
//      start index (rt60bands.py:342,393), writing only the RT60 table.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "fft.cuh"
#include "fft_plan_host.hpp"
#include "kernels.cuh"

struct aa_bands_plan {
    int n = 0;
    int kind = 0;   // 0 direct, 1 Bluestein on n/2, 2 odd n (complex length n; Bluestein unless n is smooth)
    int M = 0;      // complex transform length: n/2 (even) or n (odd)
    int x_stride = 0;   // float2 per row of the kept half spectrum
    int y_stride = 0;   // floats per band signal row (n rounded up to 4)
    aa::C2CPlan* c2c = nullptr;
    aa::TwoLevelTable twN;
    // decimated inverse transforms (length n / D) for bands whose mask support ends below n / (2 D): register-pass
    // plans with tile energies only (otherwise the band stays at full rate)
    struct Decim {
        int D = 0, n = 0, M = 0, y_stride = 0;
        aa::C2CPlan* c2c = nullptr;
        aa::TwoLevelTable twN;
    };
    static constexpr int kDecims = 8;
    Decim dec[kDecims];   // D = kDecimList[i]
    // pinned (device-visible) host words the refinement levels report their unit-list lengths through
    int* h_counts = nullptr;
};

namespace aa {

// largest decimation the band stage may use (1 = off); AA_BANDS_DECIMATION overrides the default
constexpr int kDefaultBandDecimation = 16;
constexpr int kDecimList[aa_bands_plan::kDecims] = {2, 3, 4, 6, 8, 12, 16, 24};
constexpr int kMaxBandDecimation = 24;
// largest mask value a bin may have and still be left out when a band's support pokes past its Nyquist zone
constexpr float kZoneDropMask = 0.02f;
// process default of aa_bands_options::max_decimation (per-call options override it)
static std::atomic<int> g_band_decimation{-1};
static int clamp_band_decimation(int v) { return v < 1 ? 1 : (v > kMaxBandDecimation ? kMaxBandDecimation : v); }
int band_decimation_limit() {
    int v = g_band_decimation.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* env = getenv("AA_BANDS_DECIMATION");
        v = clamp_band_decimation(env ? atoi(env) : kDefaultBandDecimation);
        g_band_decimation.store(v, std::memory_order_relaxed);
    }
    return v;
}
void set_band_decimation_limit(int max_d) { g_band_decimation.store(clamp_band_decimation(max_d), std::memory_order_relaxed); }

// Certification defaults (aa_bands_options; bandfit.cu: band_fit_kernel).  Calibrated on a numpy model of the decimated
// estimator over 160 k (row, band) units and confirmed on the device over 4096 rows x 26 bands per length
// (tools/decimation_tail.py, profiles/r02_decimation_tail.txt): no metric further than 1.5e-3 from the full-rate path
// (tolerance 5e-3), 8 % of the units redone at D <= 4 and 0.1 % at full rate.
constexpr double kCertSafety = 9.0;     // crossing window half-width = safety x modelled quadrature error
constexpr double kCertWindow = 0.015;   // a crossing window may cover at most this fraction of the fit span
constexpr double kCertTol = 3.0e-3;     // modelled relative slope change allowed from both crossing windows together
constexpr int kCertMinSpan = 240;       // decimated samples between the end points of a certified 30 dB fit, at least
constexpr int kRefineLimit = 4;         // decimation limit of the last decimated level (and the first one of short fits)
constexpr int kRefineLimitMid = 8;      // ... of the level between max_decimation and kRefineLimit

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

__global__ void untangle_kernel(const float2* __restrict__ Z, float2* __restrict__ X, int M, int x_stride,
                                TwoLevelTwiddle twN, const int* __restrict__ advance) {
    // advance != NULL: the half spectrum is stored time-advanced by advance[row] samples,
    // X'[k] = X[k] exp(+2 pi i k s / N), i.e. the row circularly rotated so that its analysis start sits at
    // sample 0 - every band signal synthesised from it starts at the analysis start (decimation classes).
    // A thread makes the PAIR X[k], X[M - k] from one load of Z[k] and Z[M - k] (the packed spectrum is read once):
    // w(M - k) = -conj(w(k)), and ((M - k) s) mod N = (M s mod N) - (k s mod N).
    const long long row = blockIdx.y;
    const float2* __restrict__ z = Z + row * (long long)M;
    float2* __restrict__ o = X + row * (long long)x_stride;
    const int s = advance ? advance[row] : 0;
    const long long N = 2LL * M;
    const long long half_turn = (s & 1) ? M : 0;   // (M s) mod N
    // (k s) mod N advances by a constant per grid stride: one 64-bit reduction per thread, adds afterwards
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long r = ((long long)(blockIdx.x * blockDim.x + threadIdx.x) * (long long)s) % N;
    const long long dr = (stride * (long long)s) % N;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; 2 * k <= M; k += (int)stride) {
        const int mk = M - k;                       // k = 0 pairs bin 0 with bin M, both from Z[0]
        const float2 zk = __ldg(z + k), zm = __ldg(z + (k == 0 ? 0 : mk));
        const float2 w = twiddle2(twN, k);
        float2 a = untangle_half_spectrum(zk, zm, w);
        float2 b = untangle_half_spectrum(zm, zk, make_float2(-w.x, w.y));
        if (s != 0) {
            long long rm = half_turn - r;
            if (rm < 0) rm += N;
            a = cmulc(a, twiddle2(twN, r));
            b = cmulc(b, twiddle2(twN, rm));
        }
        o[k] = a;
        o[mk] = b;
        r += dr;
        if (r >= N) r -= N;
    }
}

// band_mask_value() for every bin of every band's support, once per call: the masks are the same for all rows
// (rt60bands.py:365-387 rebuilds them per channel and band), so the FFT passes read a table instead of
// evaluating two float32 cosines per bin.
__global__ void band_mask_kernel(const BandDev* __restrict__ bands, double freq_scale) {
    BandDev b = bands[blockIdx.y];
    float* __restrict__ out = const_cast<float*>(b.mask);
    b.mask = nullptr;
    for (int k = b.k_first + blockIdx.x * blockDim.x + threadIdx.x; k <= b.k_last; k += gridDim.x * blockDim.x)
        out[k - b.k_first] = band_mask_value(b, (float)((double)k * freq_scale));
}

// The band table travels to the device as a kernel ARGUMENT, not through the copy engine: a host-to-device copy
// on the compute stream would queue behind the bulk input transfer of the next chunk (pipeline.run_hot_path_host)
// and stall the kernels that follow it.
constexpr int kBandsPerStore = 32;
struct BandTableArg {
    BandDev b[kBandsPerStore];
};
__global__ void band_table_kernel(BandTableArg tab, BandDev* __restrict__ out, int n) {
    if ((int)threadIdx.x < n) out[threadIdx.x] = tab.b[threadIdx.x];
}

struct BandMapArg {
    int map[AA_MAX_BANDS];   // class band -> caller band
};

// unit-list lengths -> pinned host memory (a kernel store, not the copy engine: see band_table_kernel)
__global__ void publish_counts_kernel(const int* __restrict__ counts, int* __restrict__ host_counts, int n) {
    if ((int)threadIdx.x < n) host_counts[threadIdx.x] = counts[threadIdx.x];
}

// Host mirror of _make_lowpass_mask / _make_highpass_mask / _make_bandpass_mask edge handling
// (analyse/rt60bands.py:127-167); python floats are doubles, numpy rounds them to float32 when they
// meet the float32 frequency array.
int make_band_dev(const aa_band& b, double tw_oct, double nyq, double freq_scale, int n_bins, BandDev* o) {
    memset(o, 0, sizeof(*o));
    o->kind = b.kind;
    const double fac = pow(2.0, tw_oct);
    auto clipd = [](double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); };
    double lo_edge = b.low_edge_hz, hi_edge = b.high_edge_hz;
    if (b.kind == AA_BAND_BANDPASS) {
        lo_edge = clipd(lo_edge, 1.0, nyq);
        hi_edge = clipd(hi_edge, 1.0, nyq);
        if (hi_edge <= lo_edge) o->degenerate = 1;
    } else if (b.kind != AA_BAND_LOWPASS && b.kind != AA_BAND_HIGHPASS) {
        set_error("unknown band kind %d", b.kind);
        return AA_ERR_INVALID_ARGUMENT;
    }
    double hp_stop = 0.0, lp_stop = nyq;
    if (b.kind != AA_BAND_LOWPASS) {
        double pass = clipd(lo_edge, 1.0, nyq);
        double stop = fmax(1.0, pass / fac);
        if (pass <= stop) stop = fmax(1.0, pass - 1.0);
        o->hp_pass = (float)pass;
        o->hp_stop = (float)stop;
        o->hp_span = (float)(pass - stop);
        o->hp_step = (pass <= stop) ? 1 : 0;
        hp_stop = stop;
    }
    if (b.kind != AA_BAND_HIGHPASS) {
        double pass = clipd(hi_edge, 1.0, nyq);
        double stop = fmin(nyq, pass * fac);
        if (stop <= pass) stop = fmin(nyq, pass + 1.0);
        o->lp_pass = (float)pass;
        o->lp_stop = (float)stop;
        o->lp_span = (float)(stop - pass);
        o->lp_step = (stop <= pass) ? 1 : 0;
        lp_stop = stop;
    }
    // conservative superset of the bins where the mask can be non-zero
    int kf = 0, kl = n_bins - 1;
    if (b.kind != AA_BAND_LOWPASS) kf = (int)floor(hp_stop / freq_scale) - 1;
    if (b.kind != AA_BAND_HIGHPASS) kl = (int)ceil(lp_stop / freq_scale) + 1;
    o->k_first = kf < 0 ? 0 : kf;
    o->k_last = kl > n_bins - 1 ? n_bins - 1 : kl;
    return AA_OK;
}

// band table + mask tables -> device (same stream as the transforms that read them)
int store_band_table(const BandDev* hb, int nb, BandDev* d_bands, cudaStream_t stream) {
    for (int i0 = 0; i0 < nb; i0 += kBandsPerStore) {
        BandTableArg arg;
        const int cnt = (nb - i0 < kBandsPerStore) ? (nb - i0) : kBandsPerStore;
        memset(&arg, 0, sizeof(arg));
        for (int i = 0; i < cnt; ++i) arg.b[i] = hb[i0 + i];
        band_table_kernel<<<1, kBandsPerStore, 0, stream>>>(arg, d_bands + i0, cnt);
        int rc = launch_check("band_table_kernel");
        if (rc != AA_OK) return rc;
    }
    return AA_OK;
}

int upload_bands(const aa_band* bands, int nb, double tw_oct, double nyq, double freq_scale, int N, BandDev* d_bands,
                 float* d_mask, cudaStream_t stream, std::vector<BandDev>* host_out = nullptr) {
    std::vector<BandDev> hb((size_t)nb);
    size_t mask_off = 0;
    int widest = 0;
    for (int i = 0; i < nb; ++i) {
        int rc = make_band_dev(bands[i], tw_oct, nyq, freq_scale, N / 2 + 1, &hb[(size_t)i]);
        if (rc != AA_OK) return rc;
        const int support = hb[(size_t)i].k_last - hb[(size_t)i].k_first + 1;
        hb[(size_t)i].mask = d_mask + mask_off;
        if (support > 0) mask_off += (size_t)support;
        if (support > widest) widest = support;
    }
    {
        int rc = store_band_table(hb.data(), nb, d_bands, stream);
        if (rc != AA_OK) return rc;
    }
    if (host_out) *host_out = hb;
    if (widest > 0) {
        int bx = (widest + 255) / 256;
        if (bx > 64) bx = 64;
        band_mask_kernel<<<dim3((unsigned)bx, (unsigned)nb), 256, 0, stream>>>(d_bands, freq_scale);
        int rc = launch_check("band_mask_kernel");
        if (rc != AA_OK) return rc;
    }
    return AA_OK;
}

struct Layout {
    size_t off_bands, off_mask, off_X, off_y, off_scratch, off_part, off_counts, off_lists, off_db0, off_db1,
        off_unit_start, total;
};
constexpr int kMaxLevels = 4;
constexpr int kMaxLists = (kMaxLevels - 1) * (aa_bands_plan::kDecims + 1);

// Largest per-row footprint over the decimation classes of a call (elements per row): the buffers are shared by
// the classes, which run one after the other, so decimated calls fit several times more rows per chunk.
struct ClassFootprint {
    size_t y_floats = 0, scratch_c = 0, part_floats = 0;
};

Layout make_layout(const aa_bands_plan* p, long long rows, int nb, bool smoothing = false,
                   const ClassFootprint* fp = nullptr) {
    Layout l{};
    const long long units = rows * nb;
    size_t o = 0;
    l.off_bands = o;  // the call's band table, then one re-ordered copy per decimation class
    o += align_up(sizeof(BandDev) * AA_MAX_BANDS * 2);
    l.off_counts = o;  // lengths of the refinement levels' unit lists
    o += align_up(sizeof(int) * kMaxLists);
    l.off_mask = o;  // mask tables: at most one value per rfft bin per band (offsets up to here do not depend on rows)
    o += align_up(sizeof(float) * (size_t)nb * (size_t)(p->n / 2 + 1));
    l.off_lists = o;  // unit lists of the refinement levels: one entry per (row, band) unit and level at most
    o += align_up(sizeof(int) * (size_t)units * (kMaxLevels - 1));
    l.off_X = o;
    o += align_up(sizeof(float2) * (size_t)rows * (size_t)p->x_stride);
    l.off_y = o;  // also the forward transform's packed output Z (rows * M float2 <= units * y_stride floats)
    size_t ybytes = sizeof(float) * (fp ? (size_t)rows * fp->y_floats : (size_t)units * (size_t)p->y_stride);
    size_t zbytes = sizeof(float2) * (size_t)rows * (size_t)p->M;
    o += align_up(ybytes > zbytes ? ybytes : zbytes);
    l.off_scratch = o;
    {   // the forward transform of the rows needs rows * M, the masked inverses units * M (or the class maximum)
        size_t inv = fp ? (size_t)rows * fp->scratch_c : c2c_scratch_elems(p->c2c, units);
        size_t fwd = c2c_scratch_elems(p->c2c, rows);
        o += align_up(sizeof(float2) * (inv > fwd ? inv : fwd));
    }
    {  // per-tile energy partials of the band signals (register-pass lengths; bandfit.cu)
        int tile_c = 0, n_tiles = 0, n_groups = 0;
        l.off_part = o;
        if (fp)
            o += align_up(sizeof(float) * (size_t)rows * fp->part_floats);
        else if (c2c_energy_layout(p->c2c, &tile_c, &n_tiles, &n_groups))
            o += align_up(sizeof(float) * (size_t)units * (size_t)n_tiles * (size_t)n_groups);
    }
    if (smoothing) {  // dB curves of the band signals, before and after the box filter (decay.py:161-164)
        l.off_db0 = o;
        o += align_up(sizeof(float) * (size_t)units * (size_t)p->y_stride);
        l.off_db1 = o;
        o += align_up(sizeof(float) * (size_t)units * (size_t)p->y_stride);
        l.off_unit_start = o;
        o += align_up(sizeof(int) * (size_t)units);
    }
    l.total = o;
    return l;
}

// ---- decimation classes (register-pass lengths, metrics only).  The samples y[start + D j] of a band signal are
// the length n / D inverse transform of the band's spectrum folded into [0, n / 2D] (every bin k lands on k mod n/D).
// A band whose support [k_first, k_last] lies inside ONE Nyquist zone [m n/2D, (m+1) n/2D] of the decimated rate folds
// without overlap (zone 0 is plain low-pass decimation; odd zones arrive mirrored and conjugated), so those samples
// ARE samples of the reference's band signal, and the Schroeder integral is taken on that grid with the trapezoid
// end correction (bandfit.cu).  A support that pokes past the zone edge only with mask values <= kZoneDropMask
// (the far end of a raised-cosine skirt) is cut at the edge.  Measured effect on the band RT60s: <= 8e-4 relative
// (tests/test_gpu_parity.py, tools/decimation_error.py), against a tolerance of 5e-3.
struct BandClass {
    int D = 1;
    const aa_bands_plan::Decim* sub = nullptr;
    std::vector<BandDev> bands;
    BandMapArg map{};
};

// zone placement of band b for a decimated transform of complex length Mc; on success *out is the band as the
// decimated transform sees it (support cut to the zone, z_off / z_neg set)
bool place_in_zone(const BandDev& b, int Mc, double freq_scale, BandDev* out) {
    if (b.k_last < b.k_first) return false;
    const int m = ((b.k_first + b.k_last) / 2) / Mc;
    const int lo = (m == 0) ? 0 : m * Mc + 1, hi = (m + 1) * Mc - 2;   // bins 0 and Mc of the decimated transform stay real
    BandDev o = b;
    if (b.k_first < lo) {
        if (band_mask_value(b, (float)((double)(lo - 1) * freq_scale)) > kZoneDropMask) return false;
        o.k_first = lo;
        if (o.mask) o.mask += lo - b.k_first;
    }
    if (b.k_last > hi) {
        if (band_mask_value(b, (float)((double)(hi + 1) * freq_scale)) > kZoneDropMask) return false;
        o.k_last = hi;
    }
    if (o.k_last < o.k_first) return false;
    o.z_neg = m & 1;
    o.z_off = (m & 1) ? (m + 1) * Mc : m * Mc;
    *out = o;
    return true;
}

// Decimation limits of the levels a metrics-only call runs (first = coarsest, applied to every unit; a unit whose
// fits are not certified there is redone on the next level; the last level is the full-rate path when refinement is on).
// Long late fits (T30) start at max_decimation (default 16); a call with a fit shorter than 28 dB or starting above
// -2.5 dB (T20, EDT: few decimated points per fit) starts at 4.
int fit_decimation_limit(const aa_decay_params* p) {
    int lim = 16;
    if (!p) return lim;
    for (int f = 0; f < p->n_fits && f < AA_MAX_FITS; ++f) {
        const double hi = p->fit_hi_db[f], lo = fmax(p->fit_lo_db[f], p->fit_lower_limit_db);
        if (hi > -2.5 || hi - lo < 28.0) lim = kRefineLimit;
    }
    return lim;
}

void level_limits(const aa_decay_params* fits, int max_decimation, bool refine, std::vector<int>* limits) {
    const int fit_lim = fit_decimation_limit(fits);
    const int top = (fit_lim >= 16) ? max_decimation : std::min(max_decimation, fit_lim);
    limits->clear();
    limits->push_back(top);
    if (!refine || top <= 1) return;
    if (top > kRefineLimitMid) limits->push_back(kRefineLimitMid);
    if (top > kRefineLimit) limits->push_back(kRefineLimit);
    limits->push_back(1);
}

void classify_bands(const aa_bands_plan* plan, const std::vector<BandDev>& hb, int sample_rate_hz, double freq_scale,
                    bool metrics_only, int max_d, std::vector<BandClass>* classes) {
    int tile_c = 0, n_tiles = 0, n_groups = 0;
    const bool can = (plan->n % 2) == 0 && metrics_only && c2c_energy_layout(plan->c2c, &tile_c, &n_tiles, &n_groups);
    BandClass by_d[aa_bands_plan::kDecims + 1];   // [0] full rate, [1 + c] plan->dec[c]
    for (size_t i = 0; i < hb.size(); ++i) {
        BandClass* target = &by_d[0];
        BandDev placed = hb[i];
        if (can) {
            for (int c = 0; c < aa_bands_plan::kDecims; ++c) {   // increasing D: the last fit wins
                const aa_bands_plan::Decim& d = plan->dec[c];
                BandDev z;
                if (d.c2c && d.D <= max_d && sample_rate_hz % d.D == 0 && place_in_zone(hb[i], d.M, freq_scale, &z)) {
                    target = &by_d[1 + c];
                    target->D = d.D;
                    target->sub = &d;
                    placed = z;
                }
            }
        }
        target->map.map[target->bands.size()] = (int)i;
        target->bands.push_back(placed);
    }
    classes->clear();
    for (const BandClass& c : by_d)
        if (!c.bands.empty()) classes->push_back(c);
}

// Modelled error of the decimated Schroeder sum of a band, in decimated samples of a straight decay: the 2 f part of
// y^2 (frequency theta = 4 pi f / sr per sample) has the running sum 1 / (1 - e^{-i theta}) at full rate and
// D / (1 - e^{-i D theta}) on the grid; after the trapezoid end term the two differ by (i / 2) (cot(theta / 2) -
// D cot(D theta / 2)) times the component, i.e. by that many samples' worth of local energy.  0.7 / D x the largest
// value over the pass band matches the measured rms error (numpy model) within 30 % for every band / D.
double band_quadrature_error(const aa_band& b, int D, double sr) {
    if (D <= 1) return 0.0;
    const double nyq = 0.5 * sr;
    double lo = (b.kind == AA_BAND_LOWPASS) ? 1.0 : b.low_edge_hz, hi = (b.kind == AA_BAND_HIGHPASS) ? nyq : b.high_edge_hz;
    lo = fmin(fmax(lo, 1.0), nyq - 1.0);
    hi = fmin(fmax(hi, lo), nyq - 1.0);
    double worst = 0.0;
    for (int i = 0; i < 64; ++i) {
        const double f = lo + (hi - lo) * (double)i / 63.0;
        const double th = 4.0 * 3.14159265358979323846 * f / sr;
        const double v = 0.5 * fabs(1.0 / tan(0.5 * th) - (double)D / tan(0.5 * (double)D * th));
        if (v > worst || v != v) worst = (v != v) ? 1e9 : v;
    }
    return fmin(0.7 * worst / (double)D, 1e6);
}

// One level of a call: the decimation classes at its limit, and for every caller band where it sits.
struct Level {
    int limit = 1;
    std::vector<BandClass> classes;
    int class_of[AA_MAX_BANDS], j_of[AA_MAX_BANDS], d_of[AA_MAX_BANDS];
    int list_of_class[aa_bands_plan::kDecims + 1];   // unit list of a class (levels >= 1), -1 = none
};

}  // namespace

int bands_plan_create(int n, aa_bands_plan** out) {
    AA_REQUIRE(out != nullptr, AA_ERR_INVALID_ARGUMENT, "null plan pointer");
    AA_REQUIRE(n >= 8, AA_ERR_INVALID_ARGUMENT, "Not enough samples for rt60bands analysis.");  // rt60bands.py:346
    AA_REQUIRE(n <= AA_MAX_ROW_SAMPLES, AA_ERR_TOO_LONG, "row longer than AA_MAX_ROW_SAMPLES");
    aa_bands_plan* p = new aa_bands_plan();
    p->n = n;
    const bool even = (n % 2) == 0;
    p->M = even ? n / 2 : n;
    p->x_stride = even ? n / 2 + 1 : n;
    p->y_stride = (n + 3) & ~3;
    int rc = c2c_plan_create(p->M, &p->c2c);
    if (rc == AA_OK) rc = build_two_level(n, &p->twN);
    if (rc != AA_OK) {
        if (p->c2c) c2c_plan_destroy(p->c2c);
        free_two_level(&p->twN);
        delete p;
        return rc;
    }
    p->kind = even ? (p->c2c->bluestein ? 1 : 0) : 2;
    if (cudaMallocHost(reinterpret_cast<void**>(&p->h_counts), sizeof(int) * kMaxLists) != cudaSuccess) {
        cudaGetLastError();
        p->h_counts = nullptr;   // refinement is refused without it (rt60bands_batch)
    }
    {
        int tc, nt, ng;
        if (p->kind == 0 && band_decimation_parent(n) && c2c_energy_layout(p->c2c, &tc, &nt, &ng)) {
            for (int i = 0; i < aa_bands_plan::kDecims; ++i) {
                const int D = kDecimList[i];
                if (n % (2 * D) != 0) continue;
                aa_bands_plan::Decim& d = p->dec[i];
                C2CPlan* sub = nullptr;
                if (c2c_plan_create(n / D / 2, &sub) != AA_OK) continue;
                if (!c2c_energy_layout(sub, &tc, &nt, &ng) || build_two_level(n / D, &d.twN) != AA_OK) {
                    c2c_plan_destroy(sub);
                    continue;
                }
                d.D = D;
                d.n = n / D;
                d.M = n / D / 2;
                d.y_stride = (d.n + 3) & ~3;
                d.c2c = sub;
            }
        }
    }
    *out = p;
    return AA_OK;
}

int bands_plan_kind(const aa_bands_plan* p) { return p ? p->kind : AA_ERR_INVALID_ARGUMENT; }

void bands_plan_destroy(aa_bands_plan* p) {
    if (!p) return;
    if (p->c2c) c2c_plan_destroy(p->c2c);
    if (p->h_counts) cudaFreeHost(p->h_counts);
    free_two_level(&p->twN);
    for (auto& d : p->dec) {
        if (d.c2c) c2c_plan_destroy(d.c2c);
        free_two_level(&d.twN);
    }
    delete p;
}

// decimation factor the band stage uses for each band of a metrics-only call (host only; no device work)
int bands_classify(const aa_bands_plan* plan, const aa_band* bands, int nb, double tw_oct, int sample_rate_hz,
                   const aa_decay_params* fits, int* out_decimation) {
    AA_REQUIRE(plan && bands && out_decimation, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(nb >= 1 && nb <= AA_MAX_BANDS, AA_ERR_INVALID_ARGUMENT, "band count out of range");
    AA_REQUIRE(sample_rate_hz > 0, AA_ERR_INVALID_ARGUMENT, "sample rate must be positive");
    const double sr = (double)sample_rate_hz;
    const double freq_scale = 1.0 / ((double)plan->n * (1.0 / sr));
    std::vector<BandDev> hb((size_t)nb);
    for (int i = 0; i < nb; ++i) {
        int rc = make_band_dev(bands[i], tw_oct, 0.5 * sr, freq_scale, plan->n / 2 + 1, &hb[(size_t)i]);
        if (rc != AA_OK) return rc;
    }
    std::vector<BandClass> classes;
    std::vector<int> limits;
    level_limits(fits, band_decimation_limit(), false, &limits);
    classify_bands(plan, hb, sample_rate_hz, freq_scale, true, limits[0], &classes);
    for (const BandClass& c : classes)
        for (size_t j = 0; j < c.bands.size(); ++j) out_decimation[c.map.map[j]] = c.D;
    return AA_OK;
}

size_t bands_workspace_bytes(const aa_bands_plan* p, int rows, int nb) {
    if (!p || rows < 1 || nb < 1) return 0;
    return make_layout(p, rows, nb, true).total;  // sized for the (larger) smoothing variant
}

int rt60bands_batch(const aa_bands_plan* plan, const float* d_x, long long row_stride, int n_rows,
                    const int* d_start_idx, const aa_band* bands, int nb, double tw_oct, const aa_decay_params* p,
                    double* d_metrics_out, float* d_band_signal_out, void* d_ws, size_t ws_bytes,
                    cudaStream_t stream, const aa_bands_options* opts, aa_bands_stats* stats) {
    if (stats) memset(stats, 0, sizeof(*stats));
    AA_REQUIRE(plan && d_x && d_start_idx && bands && p && d_metrics_out && d_ws, AA_ERR_INVALID_ARGUMENT,
               "null argument");
    AA_REQUIRE(nb >= 1 && nb <= AA_MAX_BANDS, AA_ERR_INVALID_ARGUMENT, "band count out of range");
    AA_REQUIRE(row_stride >= plan->n && row_stride % 4 == 0, AA_ERR_INVALID_ARGUMENT,
               "row_stride must be >= n and a multiple of 4");
    AA_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
               AA_ERR_INVALID_ARGUMENT, "input rows / workspace not aligned");
    AA_REQUIRE(p->sample_rate_hz > 0, AA_ERR_INVALID_ARGUMENT, "sample rate must be positive");
    if (n_rows <= 0) return AA_OK;
    const int N = plan->n, M = plan->M;
    const bool even = (N % 2) == 0;
    const double sr = (double)p->sample_rate_hz;
    const double freq_scale = 1.0 / ((double)N * (1.0 / sr));  // numpy.fft.rfftfreq(n, d=1/sr) (rt60bands.py:350)
    const double nyq = 0.5 * sr;

    const bool smoothing = p->smoothing_window > 1;
    aa_decay_params pk = *p;  // kernel-level parameters: smoothing is applied here, not in the fused kernel
    pk.smoothing_window = 0;
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    const Layout fixed = make_layout(plan, 1, nb, smoothing);   // band table and mask tables sit at row-independent offsets
    AA_REQUIRE(fixed.total <= ws_bytes, AA_ERR_WORKSPACE, "workspace too small for one row (see aa_bands_workspace_bytes)");
    BandDev* d_bands = reinterpret_cast<BandDev*>(ws + fixed.off_bands);
    std::vector<BandDev> hb;
    {
        int rc = upload_bands(bands, nb, tw_oct, nyq, freq_scale, N, d_bands, reinterpret_cast<float*>(ws + fixed.off_mask),
                              stream, &hb);
        if (rc != AA_OK) return rc;
    }
    // levels of this call: decimation classes per limit (classify_bands above), coarsest first
    const int opt_max_d = (opts && opts->max_decimation > 0) ? clamp_band_decimation(opts->max_decimation) : band_decimation_limit();
    const bool refine = !(opts && opts->refine < 0);
    const double cert_k = (opts && opts->cert_safety > 0.0) ? opts->cert_safety : kCertSafety;
    const double cert_w = (opts && opts->cert_window > 0.0) ? opts->cert_window : kCertWindow;
    const int cert_min_span = (opts && opts->cert_min_span > 0) ? opts->cert_min_span : kCertMinSpan;
    const double cert_tol = (opts && opts->cert_tol > 0.0) ? opts->cert_tol : kCertTol;
    const bool metrics_only = !smoothing && d_band_signal_out == nullptr;
    std::vector<int> limits;
    level_limits(p, opt_max_d, refine, &limits);
    std::vector<Level> levels;
    int n_lists = 0;
    for (size_t li = 0; li < limits.size(); ++li) {
        Level lv;
        lv.limit = limits[li];
        classify_bands(plan, hb, p->sample_rate_hz, freq_scale, metrics_only, lv.limit, &lv.classes);
        for (size_t c = 0; c < lv.classes.size(); ++c) {
            lv.list_of_class[c] = (li > 0) ? n_lists++ : -1;
            for (size_t j = 0; j < lv.classes[c].bands.size(); ++j) {
                const int b = lv.classes[c].map.map[j];
                lv.class_of[b] = (int)c;
                lv.j_of[b] = (int)j;
                lv.d_of[b] = lv.classes[c].D;
            }
        }
        levels.push_back(lv);
        if (li == 0 && lv.classes.size() == 1 && lv.classes[0].D == 1) break;   // nothing is decimated: one level
    }
    AA_REQUIRE(n_lists <= kMaxLists, AA_ERR_UNSUPPORTED, "too many refinement lists");
    const std::vector<BandClass>& classes = levels[0].classes;
    const bool one_class = classes.size() == 1 && classes[0].D == 1;
    AA_REQUIRE(one_class || levels.size() == 1 || plan->h_counts != nullptr, AA_ERR_CUDA,
               "pinned host memory for the refinement counters is missing");
    if (stats) {
        stats->n_levels = (int)levels.size();
        for (size_t li = 0; li < levels.size() && li < 4; ++li) stats->level_limit[li] = levels[li].limit;
    }
    // rows per chunk that fit the workspace; the classes share the signal / scratch / partial buffers
    ClassFootprint fp;
    if (!one_class) {
        for (size_t li = 0; li < levels.size(); ++li)
            for (const BandClass& bc : levels[li].classes) {
                const C2CPlan* c2c = bc.sub ? bc.sub->c2c : plan->c2c;
                int tile_c = 0, n_tiles = 0, n_groups = 0;
                c2c_energy_layout(c2c, &tile_c, &n_tiles, &n_groups);
                if (bc.D > 1 && band_fused_supported(bc.sub ? bc.sub->M : M)) continue;   // stays inside its CTAs
                const size_t nbc = (li == 0) ? bc.bands.size() : 1;   // later levels run unit lists in batches that fit
                fp.y_floats = std::max(fp.y_floats, nbc * (size_t)(bc.sub ? bc.sub->y_stride : plan->y_stride));
                fp.scratch_c = std::max(fp.scratch_c, nbc * (size_t)(bc.sub ? bc.sub->M : M));
                fp.part_floats = std::max(fp.part_floats, nbc * (size_t)n_tiles * (size_t)n_groups);
            }
    }
    const ClassFootprint* fpp = one_class ? nullptr : &fp;
    long long rc_rows = 0;
    {
        long long lo = 0, hi = n_rows;
        while (lo < hi) {
            long long mid = (lo + hi + 1) / 2;
            if (make_layout(plan, mid, nb, smoothing, fpp).total <= ws_bytes) lo = mid; else hi = mid - 1;
        }
        rc_rows = lo;
    }
    AA_REQUIRE(rc_rows >= 1, AA_ERR_WORKSPACE, "workspace too small for one row (see aa_bands_workspace_bytes)");
    {   // the register-pass kernels index a chunk's transforms with 32 bits: units x M of every class stays below 2^31
        long long widest = (long long)nb * M;   // one class at full rate
        if (!one_class) {
            widest = M;                          // the forward transform of the rows
            for (const BandClass& bc : classes)
                if (!(bc.D > 1 && band_fused_supported(bc.sub ? bc.sub->M : M)))
                    widest = std::max(widest, (long long)bc.bands.size() * (long long)(bc.sub ? bc.sub->M : M));
        }
        const long long cap = (2147483647LL - 1) / widest;
        if (rc_rows > cap) rc_rows = cap;
        AA_REQUIRE(rc_rows >= 1, AA_ERR_TOO_LONG, "row too long for one chunk");
    }
    const Layout lay = make_layout(plan, rc_rows, nb, smoothing, fpp);
    if (stats) {
        stats->rows_per_chunk = (int)rc_rows;
        stats->n_chunks = (int)((n_rows + rc_rows - 1) / rc_rows);
    }
    float2* d_X = reinterpret_cast<float2*>(ws + lay.off_X);
    float* d_y = reinterpret_cast<float*>(ws + lay.off_y);
    float2* d_Z = reinterpret_cast<float2*>(ws + lay.off_y);
    float2* d_scr = reinterpret_cast<float2*>(ws + lay.off_scratch);
    BandDev* d_class_bands = d_bands + AA_MAX_BANDS;
    int* d_counts = reinterpret_cast<int*>(ws + lay.off_counts);
    int* d_lists = reinterpret_cast<int*>(ws + lay.off_lists);

    for (long long r0 = 0; r0 < n_rows; r0 += rc_rows) {
        const long long rows = (n_rows - r0 < rc_rows) ? (n_rows - r0) : rc_rows;
        const long long units = rows * nb;
        const float* xin = d_x + r0 * row_stride;
        int rc;
        // ---- 1. forward real FFT of every row
        if (even) {
            SrcDesc s{};
            s.mode = SRC_REAL_PACKED;
            s.base = xin;
            s.unit_stride = row_stride;
            s.n_valid = N;
            if (c2c_can_untangle(plan->c2c)) {
                // pass 2 untangles in shared memory and writes the (time-advanced) half spectrum directly
                DstDesc d = make_dst_complex(d_Z, M, M, 0, 1.0f);
                d.untangle_X = d_X;
                d.untangle_stride = plan->x_stride;
                d.untangle_twN = plan->twN.dev;
                d.untangle_advance = one_class ? nullptr : d_start_idx + r0;
                rc = c2c_exec(plan->c2c, s, d, rows, d_scr, stream);
                if (rc != AA_OK) return rc;
            } else {
                rc = c2c_exec(plan->c2c, s, make_dst_complex(d_Z, M, M, 0, 1.0f), rows, d_scr, stream);
                if (rc != AA_OK) return rc;
                dim3 grid((unsigned)((M / 2 + 1 + 255) / 256 > 36 ? 36 : (M / 2 + 1 + 255) / 256), (unsigned)rows);  // ~8 pairs per thread: amortises the index set-up
                ProfScope prof("untangle", stream);
                untangle_kernel<<<grid, 256, 0, stream>>>(d_Z, d_X, M, plan->x_stride, plan->twN.dev,
                                                          one_class ? nullptr : d_start_idx + r0);
                rc = launch_check("untangle_kernel");
                if (rc != AA_OK) return rc;
            }
        } else {
            SrcDesc s{};
            s.mode = SRC_REAL_AS_COMPLEX;
            s.base = xin;
            s.unit_stride = row_stride;
            s.n_valid = N;
            rc = c2c_exec(plan->c2c, s, make_dst_complex(d_X, plan->x_stride, N, 0, 1.0f), rows, d_scr, stream);
            if (rc != AA_OK) return rc;
        }
        if (!one_class) {
            // ---- 2 + 3 per level and decimation class: masked inverse on the class's grid, tile energies, band_fit.
            // Level 0 runs every unit; a unit whose fits band_fit does not certify is appended to the list of the class
            // that owns its band on the next finer level, and the host reads the list lengths between levels.
            // a list holds at most one entry per (row, band of its class): lists of a level share rows x nb entries
            long long list_off[kMaxLists + 1];
            list_off[0] = 0;
            for (size_t lj = 1; lj < levels.size(); ++lj)
                for (size_t cj = 0; cj < levels[lj].classes.size(); ++cj) {
                    const int id = levels[lj].list_of_class[cj];
                    list_off[id + 1] = list_off[id] + rows * (long long)levels[lj].classes[cj].bands.size();
                }
            if (levels.size() > 1) AA_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int) * kMaxLists, stream));
            for (size_t li = 0; li < levels.size(); ++li) {
                const Level& lv = levels[li];
                if (li > 0) {   // list lengths written by the previous levels
                    publish_counts_kernel<<<1, kMaxLists, 0, stream>>>(d_counts, plan->h_counts, n_lists);
                    rc = launch_check("publish_counts_kernel");
                    if (rc != AA_OK) return rc;
                    AA_CUDA_TRY(cudaStreamSynchronize(stream));
                }
                for (size_t ci = 0; ci < lv.classes.size(); ++ci) {
                    const BandClass& bc = lv.classes[ci];
                    const int nbc = (int)bc.bands.size();
                    const int list_id = lv.list_of_class[ci];
                    const long long todo = (li == 0) ? rows * nbc : (long long)plan->h_counts[list_id];
                    if (todo <= 0) continue;
                    if (stats && li < 4) stats->level_units[li] += todo;
                    const int D = bc.D;
                    const C2CPlan* c2c = bc.sub ? bc.sub->c2c : plan->c2c;
                    const int n_c = bc.sub ? bc.sub->n : N;
                    const int M_c = bc.sub ? bc.sub->M : M;
                    const int ystride_c = bc.sub ? bc.sub->y_stride : plan->y_stride;
                    // nested stage timing per class (aa_profile_ms("band_class_d<D>")): the kernels inside keep their own scopes
                    static char class_names[kMaxBandDecimation + 1][24];
                    if (!class_names[1][0])
                        for (int d = 0; d <= kMaxBandDecimation; ++d) snprintf(class_names[d], sizeof(class_names[d]), "band_class_d%d", d);
                    ProfScope class_prof(li == 0 ? class_names[D <= kMaxBandDecimation ? D : 0] : "band_refine", stream);
                    rc = store_band_table(bc.bands.data(), nbc, d_class_bands, stream);
                    if (rc != AA_OK) return rc;
                    int tile_c = 0, n_tiles = 0, n_groups = 0;
                    AA_REQUIRE(c2c_energy_layout(c2c, &tile_c, &n_tiles, &n_groups), AA_ERR_UNSUPPORTED,
                               "decimated plan without tile energies");
                    // where the class's results go, and where its uncertified units go
                    BandFitCert cert{};
                    cert.on = (D > 1 && li + 1 < levels.size()) ? 1 : 0;
                    cert.w = (float)cert_w;
                    cert.tol = (float)cert_tol;
                    for (int f = 0; f < AA_MAX_FITS; ++f) {   // shorter fits need more points for the same slope error
                        const double range = (f < pk.n_fits) ? pk.fit_hi_db[f] - fmax(pk.fit_lo_db[f], pk.fit_lower_limit_db) : 30.0;
                        cert.fit_scale[f] = (float)fmax(1.0, 30.0 / fmax(range, 1.0));
                    }
                    cert.counts = d_counts;
                    cert.lists = d_lists;
                    for (int j = 0; j < nbc; ++j) {
                        const int b = bc.map.map[j];
                        cert.out_band[j] = (short)b;
                        cert.m[j] = (float)(cert_k * band_quadrature_error(bands[b], D, sr));
                        cert.min_span[j] = cert_min_span;
                        if (cert.on) {
                            size_t ln = li + 1;   // next level on which the band is sampled more finely
                            while (ln + 1 < levels.size() && levels[ln].d_of[b] >= D) ++ln;
                            const Level& nx = levels[ln];
                            cert.next_list[j] = (short)nx.list_of_class[nx.class_of[b]];
                            cert.next_off[j] = list_off[nx.list_of_class[nx.class_of[b]]];
                            cert.next_j[j] = (short)nx.j_of[b];
                            cert.next_nbc[j] = (short)nx.classes[(size_t)nx.class_of[b]].bands.size();
                        }
                    }
                    // unit lists run in batches that fit the buffers sized for level 0
                    long long batch = todo;
                    if (li > 0) {
                        const long long by_y = (long long)((size_t)rows * fp.y_floats / (size_t)ystride_c);
                        const long long by_s = (long long)((size_t)rows * fp.scratch_c / (size_t)M_c);
                        const long long by_p = (long long)((size_t)rows * fp.part_floats / ((size_t)n_tiles * (size_t)n_groups));
                        batch = std::max(1LL, std::min(by_y, std::min(by_s, by_p)));
                    }
                    // classes whose transform fits one CTA's shared memory: spectrum -> band signal -> fits in ONE
                    // kernel (bandfused.cu); nothing of the unit goes through the signal / scratch / partial buffers
                    if (D > 1 && band_fused_supported(M_c)) {
                        ProfScope prof("band_fused", stream);
                        rc = launch_band_fused(M_c, d_X, plan->x_stride, d_class_bands, freq_scale, bc.sub->twN.dev,
                                               1.0f / ((float)M_c * (float)D), todo, &pk, d_start_idx + r0, nbc, nb,
                                               (li == 0) ? nullptr : d_lists + list_off[list_id], &cert, 8,
                                               d_metrics_out + r0 * nb * AA_MAX_FITS, stream, D, N);
                        if (rc != AA_OK) return rc;
                        continue;
                    }
                    for (long long u0 = 0; u0 < todo; u0 += batch) {
                        const long long units_c = std::min(batch, todo - u0);
                        const int* d_map = (li == 0) ? nullptr : d_lists + list_off[list_id] + u0;
                        SrcDesc s{};
                        s.mode = SRC_BAND_PACKED;
                        s.base = d_X;
                        s.unit_stride = plan->x_stride;
                        s.n_valid = M_c;
                        s.conj = 1;
                        s.n_bands = nbc;
                        s.unit_map = d_map;
                        s.bands = d_class_bands;
                        s.bands_have_tables = 1;
                        s.freq_scale = freq_scale;
                        s.N = n_c;
                        s.M = M_c;
                        s.twN = bc.sub ? bc.sub->twN.dev : plan->twN.dev;
                        // the spectra of this path are stored time-advanced (untangle_kernel): y[j] = band signal at start + D j
                        float* d_part = reinterpret_cast<float*>(ws + lay.off_part);
                        DstDesc d = make_dst_complex(reinterpret_cast<float2*>(d_y), ystride_c / 2, M_c, 1,
                                                     1.0f / ((float)M_c * (float)D));
                        d.energy = d_part;
                        d.energy_stride = (long long)n_tiles * n_groups;
                        rc = c2c_exec(c2c, s, d, units_c, d_scr, stream);
                        if (rc != AA_OK) return rc;
                        ProfScope prof("band_fit", stream);
                        rc = launch_band_fit(d_y, ystride_c, d_part, (long long)n_tiles * n_groups, n_groups, n_tiles,
                                             2 * tile_c, n_c, units_c, &pk, d_start_idx + r0, nbc, nb, d_map, &cert, 8,
                                             d_metrics_out + r0 * nb * AA_MAX_FITS, stream, D, N, 1);
                        if (rc != AA_OK) return rc;
                    }
                }
            }
            continue;
        }
        // ---- 2. masked inverse per (row, band)
        SrcDesc s{};
        s.mode = even ? SRC_BAND_PACKED : SRC_BAND_FULL;
        s.base = d_X;
        s.unit_stride = plan->x_stride;
        s.n_valid = M;
        s.conj = 1;
        s.n_bands = nb;
        s.bands = d_bands;
        s.bands_have_tables = 1;
        s.freq_scale = freq_scale;
        s.N = N;
        s.M = even ? M : 0;
        s.twN = plan->twN.dev;
        DstDesc d{};
        if (even) {
            d = make_dst_complex(reinterpret_cast<float2*>(d_y), plan->y_stride / 2, M, 1, 1.0f / (float)M);
        } else {
            d.mode = DST_REAL_PART;
            d.base = d_y;
            d.unit_stride = plan->y_stride;
            d.n_valid = N;
            d.conj1 = 1;
            d.scale1 = 1.0f / (float)N;
            d.scale2 = 1.0f;
        }
        int tile_c = 0, n_tiles = 0, n_groups = 0;
        const bool tiled = even && !smoothing && c2c_energy_layout(plan->c2c, &tile_c, &n_tiles, &n_groups);
        float* d_part = reinterpret_cast<float*>(ws + lay.off_part);
        if (tiled) {
            d.energy = d_part;
            d.energy_stride = (long long)n_tiles * n_groups;
        }
        rc = c2c_exec(plan->c2c, s, d, units, d_scr, stream);
        if (rc != AA_OK) return rc;
        if (d_band_signal_out) {
            AA_CUDA_TRY(cudaMemcpy2DAsync(d_band_signal_out + r0 * nb * (long long)N, sizeof(float) * (size_t)N, d_y,
                                          sizeof(float) * (size_t)plan->y_stride, sizeof(float) * (size_t)N,
                                          (size_t)units, cudaMemcpyDeviceToDevice, stream));
        }
        // ---- 3. band EDC + fits, metrics only
        if (tiled) {
            // the rows pass left the energy of every tile of 2 M1 samples next to the signal: suffix-scan those and
            // revisit only the tiles with the start sample, a crossing or fit points
            ProfScope prof("band_fit", stream);
            rc = launch_band_fit(d_y, plan->y_stride, d_part, (long long)n_tiles * n_groups, n_groups, n_tiles,
                                 2 * tile_c, N, units, &pk, d_start_idx + r0, nb, nb, nullptr, nullptr, 8,
                                 d_metrics_out + r0 * nb * AA_MAX_FITS, stream);
            if (rc != AA_OK) return rc;
        } else if (!smoothing) {
            ProfScope prof("band_edc", stream);
            rc = launch_edc(d_y, plan->y_stride, nullptr, N, N, (int)units, &pk, d_start_idx + r0, nb, 8, nullptr, 0,
                            nullptr, nullptr, d_metrics_out + r0 * nb * AA_MAX_FITS, nullptr, nullptr, stream);
            if (rc != AA_OK) return rc;
        } else {
            // slow path: raw dB curve -> box filter -> floor -> generic (non-monotone) crossings and fits
            float* d_db0 = reinterpret_cast<float*>(ws + lay.off_db0);
            float* d_db1 = reinterpret_cast<float*>(ws + lay.off_db1);
            int* d_ustart = reinterpret_cast<int*>(ws + lay.off_unit_start);
            aa_decay_params praw = pk;
            praw.n_fits = 0;
            praw.want_early10 = 0;
            praw.edc_floor_db = -INFINITY;
            rc = launch_edc(d_y, plan->y_stride, nullptr, N, N, (int)units, &praw, d_start_idx + r0, nb, 8, d_db0,
                            plan->y_stride, d_ustart, nullptr, nullptr, nullptr, nullptr, stream);
            if (rc != AA_OK) return rc;
            rc = launch_box_smooth(d_db0, d_db1, plan->y_stride, nullptr, N, d_ustart, (int)units, p->smoothing_window,
                                   p->edc_floor_db, stream);
            if (rc != AA_OK) return rc;
            rc = launch_curve_fits(d_db1, plan->y_stride, nullptr, N, d_ustart, (int)units, &pk, nullptr,
                                   d_metrics_out + r0 * nb * AA_MAX_FITS, nullptr, stream);
            if (rc != AA_OK) return rc;
        }
    }
    return AA_OK;
}

}  // namespace aa
