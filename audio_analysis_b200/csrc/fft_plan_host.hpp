// Host-side FFT planning: factorisation, two-pass split, twiddle / chirp tables.
// No CUDA runtime calls here - fft.cu uploads these tables; the CPU model check
// (tests/csrc/fft_model_check.cu) uses them directly.
#pragma once

#include <math.h>
#include <string.h>

#include <vector>

#include "fft_core.cuh"

namespace aa {

constexpr int kMaxSubLen = 4096;      // longest in-CTA transform
constexpr int kSinglePassMax = 2048;  // M <= this: one pass
constexpr int kTwoLevelShift = 10;

// (M1 = RA1 * RB1, M2 = RA2 * RB2) splits served by the two-stage register-resident kernels of fft.cu
// (fft_reg.cuh): X(RA1, RB1, G1, RA2, RB2, G2).  host_direct_plan() prefers these splits.
#ifndef AA_G_D3
#define AA_G_D3 16   /* columns per CTA of the 120 x 400 split: 16 (7.22 ms for the class) against 8 (7.34) and 10 (7.66) */
#endif
#ifndef AA_G_D6
#define AA_G_D6 8   /* 120 x 200 split: 8 (3.77 ms) against 10 (3.96) and 20 (3.94) */
#endif
#ifndef AA_G_D12
#define AA_G_D12 20   /* 120 x 100 split: 20 (4.45 ms) against 4 (4.94) and 10 (4.51) */
#endif
#ifndef AA_D16_G1
#define AA_D16_G1 15
#endif
#define AA_REG_LIST(X)                                        \
    X(18, 20, 8, 20, 20, 8) /* N = 288000: M = 360 x 400 (400 x 360 with rows in groups of 16 measured 10 % slower) */ \
    X(12, 20, 8, 20, 20, 8) /* N = 192000: M = 240 x 400 */ \
    X(18, 20, 8, 10, 20, 8) /* N = 144000 (288000 decimated by 2): M = 360 x 200 */ \
    X(10, 12, 10, 15, 20, 8) /* N =  72000 (288000 decimated by 4): M = 120 x 300 */ \
    X(10, 12, AA_G_D3, 20, 20, 8) /* N =  96000 (288000 / 3, 192000 / 2): M = 120 x 400 - short columns: the issue-bound pass does less, the memory-bound rows pass more */ \
    X(10, 12, AA_G_D6, 10, 20, 8) /* N =  48000 (192000 decimated by 4): M = 120 x 200 */ \
    X(10, 12, 10, 10, 15, 8) /* N =  36000 (288000 decimated by 8): M = 120 x 150 */ \
    X(5, 8, AA_D16_G1, 15, 15, 8) /* N =  18000 (288000 decimated by 16): M = 40 x 225 (120 x 75 left 40 of 128 threads busy in the rows pass) */ \
    X(10, 12, AA_G_D12, 10, 10, 8) /* N =  24000 (192000 decimated by 8): M = 120 x 100 */ \
    X(5, 8, 15, 10, 15, 8) /* N =  12000 (192000 decimated by 16): M = 40 x 150 */ \
    X(8, 10, 8, 20, 20, 8) /* N =  64000 (192000 decimated by 3): M = 80 x 400 */ \
    X(8, 20, 4, 10, 10, 8) /* N =  32000 (192000 decimated by 6): M = 160 x 100 */ \
    X(8, 10, 4, 10, 10, 8) /* N =  16000 (192000 decimated by 12): M = 80 x 100 */ \
    X(5, 8, 4, 10, 10, 8) /* N =   8000 (192000 decimated by 24): M = 40 x 100 */

// Row lengths whose band stage may synthesise low bands on a decimated grid (bands.cu): the batch shapes the
// approximation of the Schroeder sums was validated on (configs 3 and 5).  The sub-lengths in AA_REG_LIST are
// register-pass lengths in their own right but stay at full rate when they are the row length: short files with
// the onset at sample 0 wrap the filters' pre-ringing into the tail, where a sub-sampled sum is ill-conditioned.
inline bool band_decimation_parent(int n) { return n == 288000 || n == 192000; }

inline bool reg_split_for(int M, int* M1, int* M2) {
#define X(ra1, rb1, g1, ra2, rb2, g2)                   \
    if (M == (ra1) * (rb1) * (ra2) * (rb2)) {           \
        *M1 = (ra1) * (rb1);                            \
        *M2 = (ra2) * (rb2);                            \
        return true;                                    \
    }
    AA_REG_LIST(X)
#undef X
    return false;
}

struct HostSubPlan {
    int L = 1;
    int nst = 0;
    int radix[kMaxStages] = {0};
    std::vector<float2> tw;  // exp(-2 pi i m / L)
};

struct HostTwoLevel {
    int shift = kTwoLevelShift;
    std::vector<float2> hi, lo;
};

struct HostC2C {
    int M = 0;
    bool bluestein = false;
    int M1 = 1, M2 = 1;
    HostSubPlan s1, s2;
    HostTwoLevel twM;
    // Bluestein
    int P = 0;
    std::vector<float2> chirp;  // exp(-i pi n^2 / M)
    std::vector<float2> hw;     // wrapped conj chirp, length P (its FFT is the table H)
};

inline bool is_smooth_host(long long n) {
    if (n < 1) return false;
    for (int p : {2, 3, 5, 7, 11, 13})
        while (n % p == 0) n /= p;
    return n == 1;
}

inline bool factor_radices(int L, int* radix, int* nst) {
    int n = 0;
    int rem = L;
    int twos = 0;
    while (rem % 2 == 0) {
        rem /= 2;
        ++twos;
    }
    // as many radix-8 stages as possible, the rest as radix 4 / 2 (2^4 -> 4,4)
    while (twos >= 3 && twos != 4) {
        radix[n++] = 8;
        twos -= 3;
    }
    while (twos >= 2) {
        radix[n++] = 4;
        twos -= 2;
    }
    if (twos == 1) radix[n++] = 2;
    for (int p : {3, 5, 7, 11, 13}) {
        while (rem % p == 0) {
            if (n >= kMaxStages) return false;
            radix[n++] = p;
            rem /= p;
        }
    }
    if (rem != 1) return false;
    *nst = n;
    return true;
}

inline bool host_subplan(int L, HostSubPlan* sp) {
    sp->L = L;
    memset(sp->radix, 0, sizeof(sp->radix));
    if (!factor_radices(L, sp->radix, &sp->nst)) return false;
    sp->tw.resize((size_t)L);
    for (int m = 0; m < L; ++m) {
        double a = -2.0 * M_PI * (double)m / (double)L;
        sp->tw[(size_t)m] = make_float2((float)cos(a), (float)sin(a));
    }
    return true;
}

inline void host_two_level(long long len, HostTwoLevel* t) {
    const int R = 1 << t->shift;
    long long nhi = (len + R - 1) / R + 1;
    t->hi.resize((size_t)nhi);
    t->lo.resize((size_t)R);
    for (long long a = 0; a < nhi; ++a) {
        double ang = -2.0 * M_PI * (double)((a * R) % len) / (double)len;
        t->hi[(size_t)a] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (int b = 0; b < R; ++b) {
        double ang = -2.0 * M_PI * (double)(b % len) / (double)len;
        t->lo[(size_t)b] = make_float2((float)cos(ang), (float)sin(ang));
    }
}

// true if a direct (mixed radix, <= 2 pass) plan exists for M
inline bool direct_plan_possible(int M) {
    if (!is_smooth_host(M)) return false;
    if (M <= kSinglePassMax) return true;
    for (int d = 1; (long long)d * d <= M; ++d)
        if (M % d == 0 && M / d <= kMaxSubLen) return true;
    return false;
}

inline bool host_direct_plan(int M, HostC2C* p) {
    p->M = M;
    p->bluestein = false;
    int best1 = -1, reg2 = 0;
    if (M <= kSinglePassMax) {
        best1 = 1;
    } else if (reg_split_for(M, &best1, &reg2)) {
    } else {
        int best_cost = 1 << 30;
        for (int d = 1; (long long)d * d <= M; ++d) {
            if (M % d) continue;
            int e = M / d;  // d <= e; rows (pass 2, contiguous) get the longer side
            if (e > kMaxSubLen) continue;
            if (e < best_cost) {
                best_cost = e;
                best1 = d;
            }
        }
        if (best1 < 0) return false;
    }
    p->M1 = best1;
    p->M2 = M / best1;
    if (p->M1 > 1) {
        if (!host_subplan(p->M1, &p->s1)) return false;
        host_two_level(M, &p->twM);
    }
    return host_subplan(p->M2, &p->s2);
}

inline long long bluestein_length(int M) {
    long long P = 1;
    while (P < 2LL * M - 1) P <<= 1;
    return P;
}

// Bluestein tables: X[k] = w[k] * sum_n (x[n] w[n]) conj(w)[k-n],  w[n] = exp(-i pi n^2 / M)
inline void host_bluestein_tables(int M, HostC2C* p) {
    p->M = M;
    p->bluestein = true;
    p->P = (int)bluestein_length(M);
    p->chirp.resize((size_t)M);
    p->hw.assign((size_t)p->P, make_float2(0.f, 0.f));
    const unsigned long long twoM = 2ULL * (unsigned long long)M;
    for (int n = 0; n < M; ++n) {
        unsigned long long q = ((unsigned long long)n * (unsigned long long)n) % twoM;
        double ang = -M_PI * (double)q / (double)M;
        p->chirp[(size_t)n] = make_float2((float)cos(ang), (float)sin(ang));
        float2 c = make_float2((float)cos(ang), (float)(-sin(ang)));
        p->hw[(size_t)n] = c;
        if (n > 0) p->hw[(size_t)(p->P - n)] = c;
    }
}

}  // namespace aa
