// Element accessors for the FFT pass kernels: where a logical complex element
// comes from (SrcDesc) and where a transform output goes (DstDesc).
//
// Fusing these into the first / last pass removes whole round trips through HBM:
// the band-masked, real-packed spectrum of rt60bands.py:170-175 is never
// materialised, it is generated while pass 1 loads its columns.
// Host/device so that tests/csrc/fft_model_check.cu can run the same code on CPU.
#pragma once

#include <math.h>

#include "../../include/audio_analysis_b200.h"
#include "fft_core.cuh"

namespace aa {

// Raised-cosine band mask, float32 arithmetic in the reference's operation order
// (analyse/rt60bands.py:116-167).  All edges are precomputed on the host in double
// (python floats in the reference) and rounded to float32 exactly where numpy's
// weak-scalar promotion rounds them.
struct BandDev {
    int kind;        // AA_BAND_*
    int degenerate;  // bandpass with high <= low: all-zero mask (rt60bands.py:162-163)
    // highpass part: 0 at f <= hp_stop, ramp up to 1 at f >= hp_pass
    float hp_stop, hp_pass, hp_span;  // hp_span = float(pass - stop)
    int hp_step;                      // pass <= stop: plain step at hp_pass (rt60bands.py:120-121)
    // lowpass part: 1 at f <= lp_pass, ramp down to 0 at f >= lp_stop
    float lp_pass, lp_stop, lp_span;
    int lp_step;
    int k_first, k_last;  // inclusive rfft-bin range outside of which the mask is exactly 0
    const float* mask;    // optional table of band_mask_value() for k in [k_first, k_last] (same float32 arithmetic)
    // Nyquist zone of a decimated synthesis (bands.cu; register passes only): bin kp of the length n / D transform is
    // the row's bin z_off + kp (even zone) or the conjugate of bin z_off - kp (odd zone: the spectrum folds back).
    // 0, 0 = baseband.
    int z_off, z_neg;
};

AA_HD float ramp01(float f, float x0, float span, int step, float x1) {
    if (step) return (f >= x1) ? 1.0f : 0.0f;
    float t = (f - x0) / span;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
#ifdef __CUDA_ARCH__
    float c = cosf(__fmul_rn(3.14159274101257324f, t));
    return __fsub_rn(0.5f, __fmul_rn(0.5f, c));
#else
    volatile float a = 3.14159274101257324f * t;
    volatile float c = cosf(a);
    volatile float h = 0.5f * c;
    return 0.5f - h;
#endif
}

AA_HD float band_mask_value(const BandDev& b, float f) {
    if (b.degenerate) return 0.0f;
    float hp = 1.0f, lp = 1.0f;
    if (b.kind != AA_BAND_LOWPASS) {  // highpass part (rt60bands.py:140-150)
        hp = ramp01(f, b.hp_stop, b.hp_span, b.hp_step, b.hp_pass);
        if (f <= b.hp_stop) hp = 0.0f;
        if (f >= b.hp_pass) hp = 1.0f;
    }
    if (b.kind != AA_BAND_HIGHPASS) {  // lowpass part (rt60bands.py:127-137)
        lp = 1.0f - ramp01(f, b.lp_pass, b.lp_span, b.lp_step, b.lp_stop);
        if (f <= b.lp_pass) lp = 1.0f;
        if (f >= b.lp_stop) lp = 0.0f;
    }
    return hp * lp;
}

enum SrcMode {
    SRC_COMPLEX = 0,          // float2 array
    SRC_REAL_AS_COMPLEX = 1,  // float array, imag = 0                      (odd-length forward)
    SRC_REAL_PACKED = 2,      // float array viewed as z[m] = (x[2m], x[2m+1]), zero padded, optional clip
    SRC_BAND_PACKED = 3,      // Z[k] built from the masked half spectrum    (even-length inverse)
    SRC_BAND_FULL = 4,        // Hermitian extension of the masked spectrum  (odd-length inverse)
    SRC_DECONV_PACKED = 5     // Z[k] built from H = Y conj(X) / (|X|^2 + eps)
};

struct SrcDesc {
    int mode;
    const void* base;
    long long unit_stride;  // elements of the base type between units
    int n_valid;            // logical elements (complex for 0/3/4/5, real samples for 1/2) that exist
    int conj;               // conjugate the produced value
    int clip;               // SRC_REAL_PACKED: clip samples to [-1, 1] (deconvolve.py:137)
    // band modes: unit = row * n_bands + band, or unit_map[unit] = row * n_bands + band when a unit list is given
    // (bands.cu: the full-rate redo of the decimated units whose fits are not certified)
    int n_bands;
    const int* unit_map;
    const BandDev* bands;
    int bands_have_tables;  // every bands[i].mask points at a table (required by the register-pass kernels)
    double freq_scale;      // rfftfreq: f_k = float(k * freq_scale)
    int N, M;               // real length and N/2
    TwoLevelTwiddle twN;    // exp(-2 pi i k / N)
    // band modes, register-pass kernels only: per-row circular time advance of the band signal, y'[j] = y[j + shift]
    // (Y_k * exp(+2 pi i k shift / shift_N)); the decimated band path aligns its grid with the analysis start
    const int* shift;       // [rows] or NULL
    int shift_N;            // length of the transform the spectrum came from
    TwoLevelTwiddle twShift;  // exp(-2 pi i m / shift_N)
    // deconvolution: base = Y spectra (float2, per unit), aux = sweep spectrum X (shared)
    const float2* aux;
    const float* eps_ptr;   // device scalar: regularisation (deconvolve.py:156)
    // Bluestein wrapper: value(idx) = idx < chirp_len ? inner(idx) * chirp[idx] : 0
    const float2* chirp;
    int chirp_len;
};

AA_HD float2 load_c(const float2* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// L2-only load (ld.global.cg): for buffers rewritten by other SMs during the same kernel
AA_HD float2 load_c_cg(const float2* p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}
AA_HD float load_r(const float* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// masked half-spectrum value Y[k], k in [0, N/2]
AA_HD float2 band_Y(const SrcDesc& s, const float2* X, const BandDev& b, int k) {
    if (k < b.k_first || k > b.k_last) return make_float2(0.f, 0.f);
    const float m = b.mask ? load_r(b.mask + (k - b.k_first)) : band_mask_value(b, (float)((double)k * s.freq_scale));
    if (m == 0.0f) return make_float2(0.f, 0.f);
    return cscale(load_c(X + k), m);
}

// Packed spectrum for a length-N real inverse through an N/2-point complex inverse:
//   Z[k] = Xe[k] + i Xo[k],  Xe = (Y[k] + conj Y[M-k]) / 2,  Xo = (Y[k] - conj Y[M-k]) / 2 * exp(+2 pi i k / N)
AA_HD float2 pack_half_spectrum(float2 yk, float2 ymk, float2 w_minus /* exp(-2 pi i k/N) */) {
    float2 c = cconj(ymk);
    float2 xe = cscale(cadd(yk, c), 0.5f);
    float2 xo = cmulc(cscale(csub(yk, c), 0.5f), w_minus);  // * conj(w_minus) = * exp(+2 pi i k/N)
    return cadd(xe, cmul_pi(xo));
}

AA_HD float2 src_inner(const SrcDesc& s, long long unit, int idx) {
    switch (s.mode) {
        case SRC_COMPLEX: {
            if (idx >= s.n_valid) return make_float2(0.f, 0.f);
            return load_c((const float2*)s.base + unit * s.unit_stride + idx);
        }
        case SRC_REAL_AS_COMPLEX: {
            if (idx >= s.n_valid) return make_float2(0.f, 0.f);
            return make_float2(load_r((const float*)s.base + unit * s.unit_stride + idx), 0.f);
        }
        case SRC_REAL_PACKED: {
            const float* p = (const float*)s.base + unit * s.unit_stride;
            int i0 = 2 * idx;
            float a = (i0 < s.n_valid) ? load_r(p + i0) : 0.f;
            float b = (i0 + 1 < s.n_valid) ? load_r(p + i0 + 1) : 0.f;
            if (s.clip) {
                a = fminf(fmaxf(a, -1.f), 1.f);
                b = fminf(fmaxf(b, -1.f), 1.f);
            }
            return make_float2(a, b);
        }
        case SRC_BAND_PACKED: {
            const long long um = s.unit_map ? (long long)s.unit_map[unit] : unit;
            long long row = um / s.n_bands;
            const BandDev& b = s.bands[um - row * s.n_bands];
            const float2* X = (const float2*)s.base + row * s.unit_stride;
            float2 yk = band_Y(s, X, b, idx);
            float2 ym = band_Y(s, X, b, s.M - idx);
            if (yk.x == 0.f && yk.y == 0.f && ym.x == 0.f && ym.y == 0.f) return make_float2(0.f, 0.f);
            return pack_half_spectrum(yk, ym, twiddle2(s.twN, idx));
        }
        case SRC_BAND_FULL: {
            const long long um = s.unit_map ? (long long)s.unit_map[unit] : unit;
            long long row = um / s.n_bands;
            const BandDev& b = s.bands[um - row * s.n_bands];
            const float2* X = (const float2*)s.base + row * s.unit_stride;
            int half = s.N / 2;
            if (idx <= half) return band_Y(s, X, b, idx);
            return cconj(band_Y(s, X, b, s.N - idx));
        }
        case SRC_DECONV_PACKED: {
            const float2* Y = (const float2*)s.base + unit * s.unit_stride;
            const float2* X = s.aux;
            int k = idx, mk = s.M - idx;
            const float eps = *s.eps_ptr;
            float2 xk = load_c(X + k), xm = load_c(X + mk);
            float2 hk = cscale(cmulc(load_c(Y + k), xk), 1.0f / (xk.x * xk.x + xk.y * xk.y + eps));
            float2 hm = cscale(cmulc(load_c(Y + mk), xm), 1.0f / (xm.x * xm.x + xm.y * xm.y + eps));
            return pack_half_spectrum(hk, hm, twiddle2(s.twN, idx));
        }
        default: return make_float2(0.f, 0.f);
    }
}

AA_HD float2 src_get(const SrcDesc& s, long long unit, int idx) {
    float2 v;
    if (s.chirp) {
        if (idx >= s.chirp_len) return make_float2(0.f, 0.f);
        v = src_inner(s, unit, idx);
        if (s.conj) v = cconj(v);
        return cmul(v, load_c(s.chirp + idx));
    }
    v = src_inner(s, unit, idx);
    return s.conj ? cconj(v) : v;
}

enum DstMode {
    DST_COMPLEX = 0,    // float2 array
    DST_REAL_PART = 1,  // float array, real part only
    DST_MUL_TABLE = 2   // float2 array, value * table[idx]   (Bluestein: multiply by FFT of the chirp)
};

struct DstDesc {
    int mode;
    void* base;
    long long unit_stride;
    int n_valid;          // outputs with idx >= n_valid are dropped
    // order of operations on a raw transform output v:
    int conj1;            // 1. conjugate        (inverse transform through a forward kernel)
    float scale1;         // 2. scale
    const float2* chirp;  // 3. multiply by chirp[idx]   (Bluestein post-multiply)
    int conj2;            // 4. conjugate        (caller-level inverse on top of Bluestein)
    float scale2;         // 5. scale
    const float2* table;  // DST_MUL_TABLE
    // optional (register-pass lengths only, see c2c_energy_layout): per-tile energy partials of the stored values,
    // energy[unit * energy_stride + row_group * M2 + k2] = sum over the group's rows k1 of |out[k2 * M1 + k1]|^2
    float* energy;
    long long energy_stride;
    // optional (register-pass lengths with an even M1, see c2c_can_untangle): the transform is the packed transform of
    // a real signal; pass 2 untangles it in shared memory and stores the half spectrum X[unit][0..M] here instead of
    // the packed output (base is then unused).  untangle_advance[unit] = s stores X[k] exp(+2 pi i k s / N).
    float2* untangle_X;
    long long untangle_stride;
    TwoLevelTwiddle untangle_twN;   // exp(-2 pi i m / N), N = 2 M
    const int* untangle_advance;
};

AA_HD void dst_put(const DstDesc& d, long long unit, int idx, float2 v) {
    if (idx >= d.n_valid) return;
    if (d.conj1) v = cconj(v);
    v = cscale(v, d.scale1);
    if (d.chirp) v = cmul(v, load_c(d.chirp + idx));
    if (d.conj2) v = cconj(v);
    v = cscale(v, d.scale2);
    switch (d.mode) {
        case DST_COMPLEX: ((float2*)d.base)[unit * d.unit_stride + idx] = v; break;
        case DST_REAL_PART: ((float*)d.base)[unit * d.unit_stride + idx] = v.x; break;
        case DST_MUL_TABLE: ((float2*)d.base)[unit * d.unit_stride + idx] = cmul(v, load_c(d.table + idx)); break;
        default: break;
    }
}

// rfft untangle: Z = FFT_M(z), z[m] = x[2m] + i x[2m+1]  ->  X[k], k in [0, M]
AA_HD float2 untangle_half_spectrum(float2 zk, float2 zmk, float2 w /* exp(-2 pi i k/N) */) {
    float2 c = cconj(zmk);
    float2 xe = cscale(cadd(zk, c), 0.5f);
    float2 xo = cmul_mi(cscale(csub(zk, c), 0.5f));
    return cadd(xe, cmul(w, xo));
}

}  // namespace aa
