// Stockham autosort FFT building blocks, usable from host and device code.
//
// The same functions run inside the shared-memory FFT kernels (fft.cu, csd.cu)
// and inside the CPU index-math check (tests/csrc/fft_model_check.cu), so the
// butterfly / twiddle / pass index arithmetic is verified without a GPU.
#pragma once

#include <cuda_runtime.h>

#define AA_HD __host__ __device__ __forceinline__

namespace aa {

constexpr int kMaxRadix = 13;
constexpr int kMaxStages = 24;

struct SubPlanDev {
    int L;                   // transform length (product of radix[])
    int nst;                 // number of stages
    int radix[kMaxStages];   // each in {2,3,4,5,7,8,11,13}
    const float2* tw;        // exp(-2*pi*i*m/L), m in [0, L)
    // per stage, filled by fill_stage_constants(): Ns = product of earlier radices, nb = L / radix,
    // tmul = L / (Ns * radix), and multiply-high reciprocals for the two divisions in the butterfly loop
    int Ns[kMaxStages];
    int nb[kMaxStages];
    int tmul[kMaxStages];
    unsigned nb_rcp[kMaxStages];
    unsigned ns_rcp[kMaxStages];
    unsigned L_rcp;
};

// q = n / d through one multiply-high: rcp = ceil(2^32 / d), exact while n * d < 2^32; d == 1 -> rcp = 0
AA_HD unsigned fast_rcp(unsigned d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + d - 1) / d); }
AA_HD unsigned fast_div(unsigned n, unsigned rcp) {
#ifdef __CUDA_ARCH__
    return rcp ? __umulhi(n, rcp) : n;
#else
    return rcp ? (unsigned)(((unsigned long long)n * rcp) >> 32) : n;
#endif
}

inline void fill_stage_constants(SubPlanDev* sp) {
    int Ns = 1;
    for (int s = 0; s < sp->nst; ++s) {
        const int R = sp->radix[s];
        sp->Ns[s] = Ns;
        sp->nb[s] = sp->L / R;
        sp->tmul[s] = sp->L / (Ns * R);
        sp->nb_rcp[s] = fast_rcp((unsigned)(sp->L / R));
        sp->ns_rcp[s] = fast_rcp((unsigned)Ns);
        Ns *= R;
    }
    sp->L_rcp = fast_rcp((unsigned)sp->L);
}

AA_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
AA_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
AA_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AA_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
AA_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
AA_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
AA_HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
AA_HD float2 cmul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)

// Forward DFTs of small length (sign -), in place.
AA_HD void dft2(float2* v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

AA_HD void dft3(float2* v) {
    const float s = 0.8660254037844386f;  // sin(pi/3)
    float2 t = cadd(v[1], v[2]);
    float2 d = csub(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
    float2 r = make_float2(s * d.y, -s * d.x);  // -i*s*d
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, r);
    v[2] = csub(m, r);
}

AA_HD void dft4(float2* v) {
    float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    float2 c = cadd(v[1], v[3]), d = cmul_mi(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
}

AA_HD void dft5(float2* v) {
    const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f;  // cos(2pi/5), cos(4pi/5)
    const float s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;    // sin(2pi/5), sin(4pi/5)
    float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
    float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
    float2 x0 = v[0];
    v[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
    float2 p1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
    float2 p2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
    // q = -i * (s1*b1 + s2*b2),  r = -i * (s2*b1 - s1*b2)
    float2 q = make_float2(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));
    float2 r = make_float2(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));
    v[1] = cadd(p1, q);
    v[4] = csub(p1, q);
    v[2] = cadd(p2, r);
    v[3] = csub(p2, r);
}

AA_HD void dft6(float2* v) {
    // 6 = 2 x 3 (Good-Thomas free form): two 3-point DFTs on even / odd inputs, twiddle, 2-point combine
    float2 e[3] = {v[0], v[2], v[4]};
    float2 o[3] = {v[1], v[3], v[5]};
    dft3(e);
    dft3(o);
    const float h = 0.8660254037844386f;
    float2 o1 = make_float2(0.5f * o[1].x + h * o[1].y, 0.5f * o[1].y - h * o[1].x);    // * exp(-i pi/3)
    float2 o2 = make_float2(-0.5f * o[2].x + h * o[2].y, -0.5f * o[2].y - h * o[2].x);  // * exp(-2i pi/3)
    v[0] = cadd(e[0], o[0]);
    v[3] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[4] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[5] = csub(e[2], o2);
}

AA_HD void dft8(float2* v) {
    const float h = 0.7071067811865476f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    dft4(e);
    dft4(o);
    float2 o1 = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));    // * exp(-i pi/4)
    float2 o2 = cmul_mi(o[2]);                                              // * exp(-i pi/2)
    float2 o3 = make_float2(h * (-o[3].x + o[3].y), h * (-o[3].y - o[3].x));  // * exp(-3i pi/4)
    v[0] = cadd(e[0], o[0]);
    v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);
    v[7] = csub(e[3], o3);
}

// Generic prime radix (7, 11, 13): O(R^2) with roots taken from the length-L table (R divides L).
AA_HD void dft_generic(float2* v, int R, const float2* tw, int L) {
    float2 out[kMaxRadix];
    const int step = L / R;
    for (int q = 0; q < R; ++q) {
        float2 acc = v[0];
        int idx = 0;
        for (int r = 1; r < R; ++r) {
            idx += q;
            if (idx >= R) idx -= R;
            acc = cadd(acc, cmul(v[r], tw[idx * step]));
        }
        out[q] = acc;
    }
    for (int q = 0; q < R; ++q) v[q] = out[q];
}

// One Stockham butterfly: butterfly index j in [0, L/R), Ns = product of the radices already done.
// Reads src[j + r*L/R], writes dst[(j/Ns)*Ns*R + (j%Ns) + r*Ns].  src/dst are the two ping-pong rows.
template <int R>
AA_HD void stockham_butterfly_fixed(int j, int L, int Ns, const float2* src, float2* dst, const float2* tw) {
    float2 v[R];
    const int T = L / R;
    const int k = j % Ns;
    const int tstep = k * (L / (Ns * R));  // twiddle exp(-2 pi i r k / (Ns R)) = tw[r * tstep]
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float2 x = src[j + r * T];
        v[r] = (r == 0 || tstep == 0) ? x : cmul(x, tw[r * tstep]);
    }
    if (R == 2) dft2(v);
    else if (R == 3) dft3(v);
    else if (R == 4) dft4(v);
    else if (R == 5) dft5(v);
    else if (R == 8) dft8(v);
    else dft_generic(v, R, tw, L);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[base + r * Ns] = v[r];
}

// Shared-memory index padding: one float2 of padding every 2^kPad elements (0 = none).
template <int kPad>
AA_HD int pad_index(int i) { return kPad ? i + (i >> kPad) : i; }

// Offset of stage-s twiddles in the compact table: stages are stored back to back, stage with Ns earlier points
// starts at Ns - 1 (sum over earlier stages of Ns' * (R' - 1) telescopes), r-major: [(r - 1) * Ns + k] so that
// consecutive k (consecutive threads) read consecutive words - no bank conflicts for any radix.
AA_HD int compact_twiddle_offset(int Ns) { return Ns - 1; }

// Butterfly with compact stage twiddles: twc = table + compact_twiddle_offset(Ns); full = exp(-2 pi i m / L) table
// (only the generic prime radices read it, for their roots of unity).
template <int R, int kPad>
AA_HD void stockham_butterfly_compact(int j, int k, int T, int L, int Ns, const float2* src, float2* dst,
                                      const float2* twc, const float2* full) {
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float2 x = src[pad_index<kPad>(j + r * T)];
        v[r] = (r == 0 || Ns == 1) ? x : cmul(x, twc[(r - 1) * Ns + k]);
    }
    if (R == 2) dft2(v);
    else if (R == 3) dft3(v);
    else if (R == 4) dft4(v);
    else if (R == 5) dft5(v);
    else if (R == 6) dft6(v);
    else if (R == 8) dft8(v);
    else dft_generic(v, R, full, L);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[pad_index<kPad>(base + r * Ns)] = v[r];
}

// Same butterfly with k = j % Ns and the twiddle step supplied by the caller (no division here).
template <int R>
AA_HD void stockham_butterfly_pre(int j, int k, int tstep, int T, int L, int Ns, const float2* src, float2* dst,
                                  const float2* tw) {
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float2 x = src[j + r * T];
        v[r] = (r == 0 || tstep == 0) ? x : cmul(x, tw[r * tstep]);
    }
    if (R == 2) dft2(v);
    else if (R == 3) dft3(v);
    else if (R == 4) dft4(v);
    else if (R == 5) dft5(v);
    else if (R == 8) dft8(v);
    else dft_generic(v, R, tw, L);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[base + r * Ns] = v[r];
}

AA_HD void stockham_butterfly(int R, int j, int L, int Ns, const float2* src, float2* dst, const float2* tw) {
    switch (R) {
        case 2: stockham_butterfly_fixed<2>(j, L, Ns, src, dst, tw); break;
        case 3: stockham_butterfly_fixed<3>(j, L, Ns, src, dst, tw); break;
        case 4: stockham_butterfly_fixed<4>(j, L, Ns, src, dst, tw); break;
        case 5: stockham_butterfly_fixed<5>(j, L, Ns, src, dst, tw); break;
        case 7: stockham_butterfly_fixed<7>(j, L, Ns, src, dst, tw); break;
        case 8: stockham_butterfly_fixed<8>(j, L, Ns, src, dst, tw); break;
        case 11: stockham_butterfly_fixed<11>(j, L, Ns, src, dst, tw); break;
        case 13: stockham_butterfly_fixed<13>(j, L, Ns, src, dst, tw); break;
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------
// Two-pass ("four-step") decomposition of a length-M transform, M = M1 * M2:
//   input  index n = M2*n1 + n2,  output index k = k1 + M1*k2
//   X[k1 + M1 k2] = sum_{n2} W_{M2}^{n2 k2} [ W_M^{n2 k1} sum_{n1} x[M2 n1 + n2] W_{M1}^{n1 k1} ]
// pass 1: for every column n2, an M1-point FFT over n1, times W_M^{n2 k1}, stored at [k1*M2 + n2]
// pass 2: for every row k1, an M2-point FFT over n2, stored at [k1 + M1*k2]
// ---------------------------------------------------------------------------------------------
struct TwoLevelTwiddle {
    const float2* hi;  // exp(-2 pi i (a*R) / M), a in [0, ceil(M/R))
    const float2* lo;  // exp(-2 pi i b / M),     b in [0, R)
    int shift;         // R = 1 << shift
};

AA_HD float2 twiddle2(const TwoLevelTwiddle& t, long long m) {
    const int a = (int)(m >> t.shift);
    const int b = (int)(m & ((1 << t.shift) - 1));
    return cmul(t.hi[a], t.lo[b]);
}

AA_HD long long pass1_src_index(int M2, int col, int i) { return (long long)i * M2 + col; }
AA_HD long long pass1_dst_index(int M2, int col, int i) { return (long long)i * M2 + col; }
AA_HD long long pass2_src_index(int M2, int row, int i) { return (long long)row * M2 + i; }
AA_HD long long pass2_dst_index(int M1, int row, int i) { return (long long)i * M1 + row; }

}  // namespace aa
