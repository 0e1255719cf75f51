// Whole-transform-in-shared-memory synthesis of a decimated band signal (bandfused.cu).
//
// The deep decimation classes of the rt60bands filterbank (rt60bands.py:170-175 on a grid of every D-th sample,
// bands.cu) need complex transforms of M = 4,000 ... 12,000 points per (row, band) unit: 32 ... 96 KB, which one CTA
// keeps in shared memory from the masked spectrum to the decay fits.  The transform is a decimation-in-time chain
// of three (or four) in-place stages with register-resident butterflies (fft_radix.cuh); for M = R1 R2 R3:
//
//   input index  n = d3 + R3 (d2 + R2 d1)          output index  k = k1 + R1 (k2 + R2 k3)
//   source:  the masked, packed, conjugated half spectrum (RegSrcBand) is scattered to address (d3 R2 + d2) R1 + d1 -
//            the digit reversal costs nothing because the source is addressed freely - and only where it is not
//            zero (the band's support and its mirror image);
//   stage 1: for every (d3, d2) an R1-point DFT over d1 on R1 consecutive addresses, in place (k1 replaces d1);
//   stage 2: for every (d3, k1) an R2-point DFT over d2 on addresses d3 L2 + d2 R1 + k1, inputs times
//            w_L2^{d2 k1}, L2 = R1 R2, outputs in place (k2 takes the place of d2);
//   stage 3: for every k' = k1 + R1 k2 an R3-point DFT over d3 on addresses d3 L2 + k', inputs times w_M^{d3 k'},
//            outputs in place: X[k' + L2 k3] at address k3 L2 + k', i.e. natural order.
//
// A butterfly reads and writes the same addresses, so a stage needs no second buffer and no barrier inside it; the
// stages are separated by one barrier each.  Stage twiddles are powers of ONE table value per butterfly
// (w^{k'} from a table of L_{j-1} entries), built by four interleaved chains so that the longest product has
// R / 4 + 2 factors.
//
// Host/device: tests/csrc/fft_model_check.cu runs the stages thread by thread on the CPU.
#pragma once

#include "fft_reg.cuh"

namespace aa {

namespace fused_detail {
template <int... Rs>
struct Radices {
    static constexpr int kN = (int)sizeof...(Rs);
    AA_HD static constexpr int get(int j) {   // radix of stage j = 1 .. kN
        constexpr int r[] = {Rs...};
        return r[j - 1];
    }
    AA_HD static constexpr int len(int j) {   // L_j = R_1 ... R_j, L_0 = 1
        constexpr int r[] = {Rs...};
        int l = 1;
        for (int q = 0; q < j; ++q) l *= r[q];
        return l;
    }
};
}  // namespace fused_detail

// FusedPlan<threads, R1, R2, ..., Rs>: s = 3 or 4 stages
template <int kThreads_, int... Rs>
struct FusedPlan {
    using Rad = fused_detail::Radices<Rs...>;
    static constexpr int kStages = Rad::kN;
    static constexpr int M = Rad::len(kStages);
    static constexpr int R1 = Rad::get(1);
    static constexpr int kThreads = kThreads_;
    AA_HD static constexpr int radix(int j) { return Rad::get(j); }
    AA_HD static constexpr int len(int j) { return Rad::len(j); }
    // stage j >= 2 reads tw[tw_offset(j) + k'] = exp(-2 pi i k' / L_j), k' < L_{j-1}
    AA_HD static constexpr int tw_offset(int j) {
        int o = 0;
        for (int q = 2; q < j; ++q) o += Rad::len(q - 1);
        return o;
    }
    static constexpr int kTwTotal = tw_offset(kStages + 1);
    static_assert(kStages >= 2 && kStages <= 4, "two to four stages");
};

// (M, threads, resident CTAs per SM, R1, R2, ...) per supported length.  R1 is odd so that stage 1's accesses (R1
// consecutive values per thread) do not collide on banks.
// Tuning of the round (bench classes D = 16 / 12, ms per 8192 rows): 9000 with 192 / 224 / 256 / 320 threads 8.19 / 8.30 /
// 7.89 / 8.31; four-stage plans 9 x 10 x 10 x 10 and 15 x 15 x 8 x 5 with 384 / 512 threads 8.6-9.4; 12000 with
// 256 / 320 / 384 / 448 / 512 threads 3.85 / 3.73 / 3.58 / 3.55 / 3.50.
#ifndef AA_FUSED_LIST
#define AA_FUSED_LIST(X)                                                                  \
    X(9000, 256, 3, 25, 18, 20)  /* N = 288000 decimated by 16;           73.7 KB */      \
    X(12000, 512, 2, 25, 24, 20) /* 288000 / 12, 192000 / 8;              97.7 KB */      \
    X(6000, 256, 4, 15, 20, 20)  /* 192000 / 16;                          49.1 KB */      \
    X(8000, 256, 3, 25, 16, 20)  /* 192000 / 12;                          65.5 KB */
#endif

// v[r] *= w^r, r = 1 .. R-1, with four interleaved power chains
template <int R>
AA_HD void mul_powers(float2* v, float2 w1) {
    if constexpr (R > 1) v[1] = cmul(v[1], w1);
    if constexpr (R > 2) {
        const float2 w2 = cmul(w1, w1);
        v[2] = cmul(v[2], w2);
        if constexpr (R > 3) {
            const float2 w3 = cmul(w2, w1);
            v[3] = cmul(v[3], w3);
            if constexpr (R > 4) {
                const float2 w4 = cmul(w2, w2);
                float2 p[4] = {w4, w1, w2, w3};   // p[r & 3] = w^(last multiple handled in that residue class)
                v[4] = cmul(v[4], w4);
                static_for<5, R>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    p[r & 3] = cmul(p[r & 3], w4);
                    v[r] = cmul(v[r], p[r & 3]);
                });
            }
        }
    }
}

// ---- stage 1, source side.  The packed spectrum Z of a band is zero outside the band's support and its mirror image
// (Z[idx] needs Y[idx] and Y[M - idx]), typically a tenth or two of the M points: the CTA clears its buffer, walks the
// SUPPORT once - consecutive threads read consecutive bins of the row spectrum and of the mask table - and scatters
// the pair Z[kp], Z[M - kp], both built from the same loads, to their stage-1 addresses.

template <class P>
AA_HD void fused_clear(int t, float2* data) {
    static_assert(P::M % 2 == 0, "the buffer is cleared in 16-byte words");
    float4* p = reinterpret_cast<float4*>(data);
    constexpr int kWords = P::M / 2, kFull = kWords / P::kThreads;
#pragma unroll
    for (int j = 0; j < kFull; ++j) p[t + j * P::kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t + kFull * P::kThreads < kWords) p[t + kFull * P::kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// address of input n = d_s + R_s (d_{s-1} + R_{s-1} ( ... + R_2 d_1)) in the layout the stages transform in place:
// sum over j of d_j L_{j-1}
template <class P, int J = P::kStages>
AA_HD int fused_input_address(int n) {
    if constexpr (J == 1) {
        return n;
    } else {
        const int q = n / P::radix(J), d = n - q * P::radix(J);
        return d * P::len(J - 1) + fused_input_address<P, J - 1>(q);
    }
}

// the first kFusedPre rounds of a thread's support bins are loaded BEFORE the buffer is cleared (fused_scatter_load), so
// that the latency of the spectrum / mask / twiddle loads hides behind the clear and its barrier
constexpr int kFusedPre = 4;
struct FusedPrefetch {
    float2 y[kFusedPre];   // masked (and, for odd zones, conjugated) spectrum value Y[kp]
    float2 w[kFusedPre];   // exp(-2 pi i kp / N)
};

struct FusedSupport {
    int A, B;        // support in the coordinates of the decimated transform (bin kp <-> row bin z_off +- kp)
    bool straddle;   // the support meets its own mirror image [M - B, M - A] (it crosses the middle of the zone)
};

template <class P>
AA_HD FusedSupport fused_support(const RegSrcBand& s) {
    FusedSupport f;
    f.A = s.b.z_neg ? s.b.z_off - s.b.k_last : s.b.k_first - s.b.z_off;
    f.B = s.b.z_neg ? s.b.z_off - s.b.k_first : s.b.k_last - s.b.z_off;
    if (s.b.k_last < s.b.k_first) f.B = f.A - 1;   // empty band
    f.straddle = (P::M - f.B <= f.B) && (P::M - f.A >= f.A);
    return f;
}

template <class P>
AA_HD float2 fused_band_value(const RegSrcBand& s, int kp) {
    float2 y = s.masked(s.zone_bin(kp));
    if (s.b.z_neg) y.y = -y.y;
    return y;
}

// Y[M - kp] = 0: with h = Y[kp] / 2 and t = h conj(w), pack_half_spectrum gives
//   Z[kp] = h + i t  and  Z[M - kp] = conj(h) + i conj(t)   (w(M - kp) = -conj(w)); both are stored conjugated
template <class P>
AA_HD void fused_store_pair(float2* data, int kp, float2 y, float2 w) {
    const float2 h = cscale(y, 0.5f);
    const float2 tt = cmulc(h, w);
    const int mk = P::M - kp;
    data[fused_input_address<P>(kp)] = make_float2(h.x - tt.y, -(h.y + tt.x));
    if (mk < P::M) data[fused_input_address<P>(mk)] = make_float2(h.x + tt.y, h.y - tt.x);
}

template <class P>
AA_HD void fused_scatter_load(int t, const RegSrcBand& s, const FusedSupport& f, FusedPrefetch* pre) {
    static_for<0, kFusedPre>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        const int kp = f.A + t + j * P::kThreads;
        if (kp <= f.B) {
            pre->y[j] = fused_band_value<P>(s, kp);
            pre->w[j] = twiddle2(s.twN, kp);
        }
    });
}

template <class P>
AA_HD void fused_scatter(int t, const RegSrcBand& s, const FusedSupport& f, const FusedPrefetch& pre, float2* data) {
    if (f.straddle) {
        // both Y[kp] and Y[M - kp] exist for some bins: the general element, one point per iteration (the mirrored
        // point of a bin whose partner is outside the support is still built here)
        for (int kp = f.A + t; kp <= f.B; kp += P::kThreads) {
            const int mk = P::M - kp;
            const float2 w = twiddle2(s.twN, kp);
            if (mk >= f.A && mk <= f.B) data[fused_input_address<P>(kp)] = s.element(kp, true, true, w);
            else fused_store_pair<P>(data, kp, fused_band_value<P>(s, kp), w);
        }
        return;
    }
    static_for<0, kFusedPre>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        const int kp = f.A + t + j * P::kThreads;
        if (kp <= f.B) fused_store_pair<P>(data, kp, pre.y[j], pre.w[j]);
    });
    for (int kp = f.A + t + kFusedPre * P::kThreads; kp <= f.B; kp += P::kThreads)
        fused_store_pair<P>(data, kp, fused_band_value<P>(s, kp), twiddle2(s.twN, kp));
}

// stage 1: butterfly b transforms the R1 consecutive addresses b R1 .. b R1 + R1 - 1 in place (nothing to do where
// the band left them all zero)
template <class P>
AA_HD void fused_stage1(int t, float2* data) {
    for (int b = t; b < P::M / P::R1; b += P::kThreads) {
        float2* base = data + b * P::R1;
        float2 v[P::R1];
        bool any = false;
        static_for<0, P::R1>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            v[r] = base[r];
            any |= (v[r].x != 0.0f) | (v[r].y != 0.0f);
        });
        if (!any) continue;
        dft_r<P::R1>(v);
        static_for<0, P::R1>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            base[r] = v[r];
        });
    }
}

// stage J >= 2: butterfly i = k' + L_{J-1} b works on addresses b L_J + rho L_{J-1} + k', inputs times w_LJ^{rho k'}
// (powers of tw[tw_offset(J) + k']), outputs in place.  The last stage conjugates (conj != 0) and scales its outputs:
// with the conjugated source this is the inverse transform (fft_access.cuh: DstDesc::conj1 / scale1).
template <class P, int J>
AA_HD void fused_stage(int t, float2* data, const float2* tw, int conj, float scale) {
    constexpr int R = P::radix(J), Lp = P::len(J - 1), Lj = P::len(J);
    constexpr bool last = (J == P::kStages);
    for (int i = t; i < P::M / R; i += P::kThreads) {
        const int b = i / Lp, kq = i - b * Lp;
        float2* base = data + b * Lj + kq;
        float2 v[R];
        static_for<0, R>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            v[r] = base[r * Lp];
        });
        if (kq != 0) mul_powers<R>(v, load_c(tw + P::tw_offset(J) + kq));
        dft_r<R>(v);
        static_for<0, R>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            float2 o = v[r];
            if (last) {
                if (conj) o.y = -o.y;
                o = cscale(o, scale);
            }
            base[r * Lp] = o;
        });
    }
}

}  // namespace aa
