// Whole-transform-in-shared-memory synthesis of a decimated band signal (bandfused.cu).
//
// The deep decimation classes of the rt60bands filterbank (rt60bands.py:170-175 on a grid of every D-th sample,
// bands.cu) need complex transforms of M = 4,000 ... 12,000 points per (row, band) unit: 32 ... 96 KB, which one CTA
// keeps in shared memory from the masked spectrum to the decay fits.  The transform is a decimation-in-time chain
// of three in-place stages with register-resident butterflies (fft_radix.cuh), M = R1 R2 R3:
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

template <int R1_, int R2_, int R3_, int kThreads_>
struct FusedPlan {
    static constexpr int R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int L2 = R1_ * R2_, M = R1_ * R2_ * R3_;
    static constexpr int kThreads = kThreads_;
    static constexpr int kTw2 = R1_;        // entries of the stage-2 table  w_L2^{k1}
    static constexpr int kTw3 = R1_ * R2_;  // entries of the stage-3 table  w_M^{k'}
};

// (R1, R2, R3, threads, resident CTAs per SM) per supported length; R1 is odd so that stage 1's stores of R1
// consecutive values per thread do not collide on banks
#define AA_FUSED_LIST(X)                                                       \
    X(25, 18, 20, 256, 3) /* M =  9000: N = 288000 decimated by 16; 73.7 KB */ \
    X(25, 24, 20, 256, 2) /* M = 12000: 288000 / 12, 192000 / 8;    97.7 KB */ \
    X(15, 20, 20, 256, 4) /* M =  6000: 192000 / 16;                49.1 KB */ \
    X(25, 16, 20, 256, 3) /* M =  8000: 192000 / 12;                65.5 KB */

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
    for (int i = t; i < P::M / 2; i += P::kThreads) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// address of input idx = d3 + R3 (d2 + R2 d1) in the layout stage 1 transforms in place: (d3 R2 + d2) R1 + d1
template <class P>
AA_HD int fused_input_address(int idx) {
    constexpr int S = P::R2 * P::R3;
    const int d1 = idx / S, i = idx - d1 * S;
    const int d2 = i / P::R3, d3 = i - d2 * P::R3;
    return (d3 * P::R2 + d2) * P::R1 + d1;
}

template <class P>
AA_HD void fused_scatter(int t, const RegSrcBand& s, float2* data) {
    if (s.b.k_last < s.b.k_first) return;   // empty band: all zero
    // support in the coordinates of the decimated transform (bin kp <-> row bin z_off +- kp)
    const int A = s.b.z_neg ? s.b.z_off - s.b.k_last : s.b.k_first - s.b.z_off;
    const int B = s.b.z_neg ? s.b.z_off - s.b.k_first : s.b.k_last - s.b.z_off;
    for (int kp = A + t; kp <= B; kp += P::kThreads) {
        const int mk = P::M - kp;
        const bool in_m = mk >= A && mk <= B;   // only when the support straddles the middle of the zone
        const float2 w = twiddle2(s.twN, kp);   // exp(-2 pi i kp / N), N = 2 M
        data[fused_input_address<P>(kp)] = s.element(kp, true, in_m, w);
        // the mirrored point; when mk is in the support itself, its own iteration builds it
        if (!in_m && mk < P::M) data[fused_input_address<P>(mk)] = s.element(mk, false, true, make_float2(-w.x, w.y));
    }
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

// stage 2: butterflies i = k1 + R1 d3; tw2[k1] = exp(-2 pi i k1 / L2)
template <class P>
AA_HD void fused_stage2(int t, float2* data, const float2* tw2) {
    for (int i = t; i < P::M / P::R2; i += P::kThreads) {
        const int d3 = i / P::R1, k1 = i - d3 * P::R1;
        float2* base = data + d3 * P::L2 + k1;
        float2 v[P::R2];
        static_for<0, P::R2>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            v[r] = base[r * P::R1];
        });
        if (k1 != 0) mul_powers<P::R2>(v, load_c(tw2 + k1));
        dft_r<P::R2>(v);
        static_for<0, P::R2>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            base[r * P::R1] = v[r];
        });
    }
}

// stage 3: butterflies i = k'; tw3[k'] = exp(-2 pi i k' / M).  The outputs are conjugated (conj != 0) and scaled:
// with the conjugated source this is the inverse transform (fft_access.cuh: DstDesc::conj1 / scale1).
template <class P>
AA_HD void fused_stage3(int t, float2* data, const float2* tw3, int conj, float scale) {
    for (int i = t; i < P::L2; i += P::kThreads) {
        float2* base = data + i;
        float2 v[P::R3];
        static_for<0, P::R3>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            v[r] = base[r * P::L2];
        });
        if (i != 0) mul_powers<P::R3>(v, load_c(tw3 + i));
        dft_r<P::R3>(v);
        static_for<0, P::R3>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            float2 o = v[r];
            if (conj) o.y = -o.y;
            base[r * P::L2] = cscale(o, scale);
        });
    }
}

}  // namespace aa
