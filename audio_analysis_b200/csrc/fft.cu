// Batched complex FFT engine for sm_100a: shared-memory Stockham passes, a
// two-pass (four-step) split for transforms that do not fit one CTA, and a
// Bluestein chirp-z fallback for lengths with prime factors above 13.
//
// Used for the exact length-N circular filterbank of analyse/rt60bands.py:170-175
// (numpy pocketfft rfft/irfft on arbitrary N) and for the power-of-two
// deconvolution transform (analyse/deconvolve.py:147-171).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fft.cuh"
#include "fft_plan_host.hpp"
#include "fft_reg.cuh"

namespace aa {

namespace {

constexpr int kPassThreads = 256;
constexpr size_t kSmemSoft = 56 * 1024;  // per CTA: aim for 4 CTAs (32 warps) per SM
constexpr size_t kSmemHard = 200 * 1024;

int upload(const std::vector<float2>& h, float2** d) {
    AA_CUDA_TRY(cudaMalloc(d, sizeof(float2) * h.size()));
    AA_CUDA_TRY(cudaMemcpy(*d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    return AA_OK;
}

int upload_subplan(const HostSubPlan& h, SubPlan* sp) {
    memset(&sp->dev, 0, sizeof(sp->dev));
    sp->dev.L = h.L;
    sp->dev.nst = h.nst;
    memcpy(sp->dev.radix, h.radix, sizeof(h.radix));
    fill_stage_constants(&sp->dev);
    int rc = upload(h.tw, &sp->d_tw);
    if (rc != AA_OK) return rc;
    sp->dev.tw = sp->d_tw;
    return AA_OK;
}

int upload_two_level(const HostTwoLevel& h, TwoLevelTable* out) {
    int rc = upload(h.hi, &out->d_hi);
    if (rc != AA_OK) return rc;
    rc = upload(h.lo, &out->d_lo);
    if (rc != AA_OK) return rc;
    out->dev.hi = out->d_hi;
    out->dev.lo = out->d_lo;
    out->dev.shift = h.shift;
    return AA_OK;
}

void choose_group(int L, int* G, int* Lp, size_t* smem) {
    int lp = L | 1;  // odd row pitch: column-major accesses hit distinct banks
    size_t per = 2 * sizeof(float2) * (size_t)lp;
    size_t twb = sizeof(float2) * (size_t)L;  // stage twiddles live in shared memory too
    int g = (int)((kSmemSoft - (twb < kSmemSoft ? twb : 0)) / per);
    if (g < 4) g = (int)((kSmemHard - twb) / per);
    if (g >= 16) g = 16;
    else if (g >= 8) g = 8;
    else if (g >= 4) g = 4;
    if (g < 1) g = 1;
    *G = g;
    *Lp = lp;
    *smem = per * (size_t)g + twb;
}

// ------------------------------------------------------------------------------------
// pass kernels
// ------------------------------------------------------------------------------------

struct PassArgs {
    SubPlanDev sp;
    int G, Lp;
    unsigned G_rcp;
    int M1, M2;
    int groups;  // CTAs per unit
    TwoLevelTwiddle twM;
};

template <int R>
__device__ __forceinline__ void stage_loop(const SubPlanDev& sp, int s, int G, int Lp, const float2* __restrict__ a,
                                           float2* __restrict__ b, const float2* __restrict__ tw) {
    const int L = sp.L, Ns = sp.Ns[s], nb = sp.nb[s];
    const unsigned nb_rcp = sp.nb_rcp[s], ns_rcp = sp.ns_rcp[s];
    const int total = G * nb;
    const float2* twc = tw + compact_twiddle_offset(Ns);
    for (int e = threadIdx.x; e < total; e += kPassThreads) {
        const int g = (int)fast_div((unsigned)e, nb_rcp);
        const int j = e - g * nb;
        const int k = j - (int)fast_div((unsigned)j, ns_rcp) * Ns;
        stockham_butterfly_compact<R, 0>(j, k, nb, L, Ns, a + g * Lp, b + g * Lp, twc, sp.tw);
    }
}

__device__ __forceinline__ void run_stages(const SubPlanDev& sp, int G, int Lp, float2*& a, float2*& b,
                                           const float2* tw) {
    for (int s = 0; s < sp.nst; ++s) {
        switch (sp.radix[s]) {  // one branch per stage, not per butterfly
            case 2: stage_loop<2>(sp, s, G, Lp, a, b, tw); break;
            case 3: stage_loop<3>(sp, s, G, Lp, a, b, tw); break;
            case 4: stage_loop<4>(sp, s, G, Lp, a, b, tw); break;
            case 5: stage_loop<5>(sp, s, G, Lp, a, b, tw); break;
            case 6: stage_loop<6>(sp, s, G, Lp, a, b, tw); break;
            case 7: stage_loop<7>(sp, s, G, Lp, a, b, tw); break;
            case 8: stage_loop<8>(sp, s, G, Lp, a, b, tw); break;
            case 11: stage_loop<11>(sp, s, G, Lp, a, b, tw); break;
            case 13: stage_loop<13>(sp, s, G, Lp, a, b, tw); break;
            default: break;
        }
        __syncthreads();
        float2* t = a;
        a = b;
        b = t;
    }
}

// compact per-stage twiddles: stage with Ns earlier points holds w_{Ns R}^{r k} at [Ns - 1 + (r - 1) Ns + k]
__device__ __forceinline__ void load_stage_twiddles(const SubPlanDev& sp, float2* tw_s) {
    for (int s = 0; s < sp.nst; ++s) {
        const int Ns = sp.Ns[s], R = sp.radix[s], tmul = sp.tmul[s];
        const unsigned ns_rcp = sp.ns_rcp[s];
        float2* t = tw_s + compact_twiddle_offset(Ns);
        for (int m = threadIdx.x; m < Ns * (R - 1); m += kPassThreads) {
            const int r1 = (int)fast_div((unsigned)m, ns_rcp);
            const int k = m - r1 * Ns;
            t[m] = __ldg(sp.tw + (r1 + 1) * k * tmul);
        }
    }
}

// Source fast paths: kSrc = SRC_COMPLEX / SRC_REAL_PACKED / SRC_BAND_PACKED are inlined without the mode
// switch and the Bluestein wrapper; kSrc = -1 is the generic accessor.
template <int kSrc>
__device__ __forceinline__ float2 fetch(const SrcDesc& s, long long unit, int idx) {
    if (kSrc == SRC_COMPLEX) {
        float2 v = (idx < s.n_valid) ? __ldg((const float2*)s.base + unit * s.unit_stride + idx) : make_float2(0.f, 0.f);
        return s.conj ? cconj(v) : v;
    } else if (kSrc == SRC_REAL_PACKED) {
        const float* p = (const float*)s.base + unit * s.unit_stride;
        const int i0 = 2 * idx;
        float2 v;
        if (i0 + 1 < s.n_valid) v = __ldg((const float2*)(p + i0));
        else v = make_float2(i0 < s.n_valid ? __ldg(p + i0) : 0.f, 0.f);
        if (s.clip) {
            v.x = fminf(fmaxf(v.x, -1.f), 1.f);
            v.y = fminf(fmaxf(v.y, -1.f), 1.f);
        }
        return v;
    } else if (kSrc == SRC_BAND_PACKED) {
        const long long um = s.unit_map ? (long long)__ldg(s.unit_map + unit) : unit;
        const long long row = um / s.n_bands;
        const BandDev& b = s.bands[um - row * s.n_bands];
        const int mk = s.M - idx;
        const bool in_k = (idx >= b.k_first && idx <= b.k_last), in_m = (mk >= b.k_first && mk <= b.k_last);
        if (!in_k && !in_m) return make_float2(0.f, 0.f);
        const float2* X = (const float2*)s.base + row * s.unit_stride;
        float2 yk = band_Y(s, X, b, idx);
        float2 ym = band_Y(s, X, b, mk);
        float2 v = pack_half_spectrum(yk, ym, twiddle2(s.twN, idx));
        return s.conj ? cconj(v) : v;
    } else {
        return src_get(s, unit, idx);
    }
}

// pass 1: columns.  CTA = G adjacent columns n2 of one unit; M1-point FFT over n1; twiddle; store [k1*M2 + n2].
template <int kSrc>
__global__ void __launch_bounds__(kPassThreads, 4) fft_pass1_kernel(PassArgs pa, SrcDesc src, float2* __restrict__ scratch,
                                                                 long long scratch_stride) {
    extern __shared__ float2 smem[];
    const long long unit = blockIdx.x / pa.groups;
    const int cg = blockIdx.x - (int)(unit * pa.groups);
    const int col0 = cg * pa.G;
    const int L = pa.sp.L, G = pa.G, Lp = pa.Lp;
    float2* a = smem;
    float2* b = smem + (size_t)G * Lp;
    float2* tw_s = b + (size_t)G * Lp;
    load_stage_twiddles(pa.sp, tw_s);
    const int total = G * L;
#pragma unroll 4
    for (int e = threadIdx.x; e < total; e += kPassThreads) {
        const int i = (int)fast_div((unsigned)e, pa.G_rcp);
        const int g = e - i * G;
        const int col = col0 + g;
        float2 v = make_float2(0.f, 0.f);
        if (col < pa.M2) v = fetch<kSrc>(src, unit, i * pa.M2 + col);
        a[g * Lp + i] = v;
    }
    __syncthreads();
    run_stages(pa.sp, G, Lp, a, b, tw_s);
    float2* __restrict__ out = scratch + unit * scratch_stride;
#pragma unroll 4
    for (int e = threadIdx.x; e < total; e += kPassThreads) {
        const int i = (int)fast_div((unsigned)e, pa.G_rcp);
        const int g = e - i * G;
        const int col = col0 + g;
        if (col < pa.M2) {
            float2 v = cmul(a[g * Lp + i], twiddle2(pa.twM, (long long)col * i));
            out[(long long)i * pa.M2 + col] = v;
        }
    }
}

// pass 2: rows.  CTA = G adjacent rows k1 of one unit; M2-point FFT over n2; store [k1 + M1*k2].
// kDst = DST_COMPLEX inlines the plain conj/scale store; -1 is the generic sink.
template <int kSrc, int kDst>
__global__ void __launch_bounds__(kPassThreads, 4) fft_pass2_kernel(PassArgs pa, SrcDesc src, DstDesc dst) {
    extern __shared__ float2 smem[];
    const long long unit = blockIdx.x / pa.groups;
    const int rg = blockIdx.x - (int)(unit * pa.groups);
    const int row0 = rg * pa.G;
    const int L = pa.sp.L, G = pa.G, Lp = pa.Lp;
    float2* a = smem;
    float2* b = smem + (size_t)G * Lp;
    float2* tw_s = b + (size_t)G * Lp;
    load_stage_twiddles(pa.sp, tw_s);
    const int total = G * L;
#pragma unroll 4
    for (int e = threadIdx.x; e < total; e += kPassThreads) {
        const int g = (int)fast_div((unsigned)e, pa.sp.L_rcp);
        const int i = e - g * L;
        const int row = row0 + g;
        float2 v = make_float2(0.f, 0.f);
        if (row < pa.M1) v = fetch<kSrc>(src, unit, row * pa.M2 + i);
        a[g * Lp + i] = v;
    }
    __syncthreads();
    run_stages(pa.sp, G, Lp, a, b, tw_s);
#pragma unroll 4
    for (int e = threadIdx.x; e < total; e += kPassThreads) {
        const int i = (int)fast_div((unsigned)e, pa.G_rcp);
        const int g = e - i * G;
        const int row = row0 + g;
        if (row < pa.M1) {
            const int idx = i * pa.M1 + row;
            float2 v = a[g * Lp + i];
            if (kDst == DST_COMPLEX) {
                if (idx < dst.n_valid) {
                    if (dst.conj1) v = cconj(v);
                    ((float2*)dst.base)[unit * dst.unit_stride + idx] = cscale(v, dst.scale1);
                }
            } else {
                dst_put(dst, unit, idx, v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// two-stage register-resident passes for the hot lengths (fft_reg.cuh); everything else uses the kernels above
// ------------------------------------------------------------------------------------

// transform pitch in shared memory: conflict-free for 64-bit accesses with g as the fastest thread index
constexpr int reg_pitch(int G, int min_elems) {
    int tp = min_elems;
    if (G == 8) {
        while (tp % 8 != 6) ++tp;
    } else if (G == 4) {
        while (tp % 16 != 4) ++tp;
    } else {
        tp |= 1;
    }
    return tp;
}

template <int RA, int RB, int G>
using RegCfg = RegPass<RA, RB, G, reg_pitch(G, RB * (RA | 1))>;

// resident CTAs per SM the band variant is compiled for: a cap of 64 registers per thread (AA_COLS_REG_CAP; 56 and no cap measured 5 % and 2 % slower) (the band source's
// batched loads otherwise push it to 70-80 and the occupancy to a third)
#ifndef AA_COLS_REG_CAP
#define AA_COLS_REG_CAP 64
#endif
constexpr int cols_min_blocks(int threads, int src) {
    return (src == SRC_BAND_PACKED && AA_COLS_REG_CAP > 0) ? 65536 / (AA_COLS_REG_CAP * threads) : 1;
}

template <class P, int kSrc, int M2>
__global__ void __launch_bounds__(P::kThreads, cols_min_blocks(P::kThreads, kSrc)) fft_cols_reg(SrcDesc src, float2* __restrict__ scratch, TwoLevelTwiddle twM,
                                                            const float2* __restrict__ tw_full) {
    static_assert(M2 % P::G == 0, "columns must split evenly over CTAs");
    constexpr int groups = M2 / P::G;
    extern __shared__ float2 smem[];
    float2* buf = smem;
    float2* tws = smem + P::kBufElems;
    float2* colq = tws + P::L;
    const int unit = blockIdx.x / groups;
    const int col0 = (blockIdx.x - unit * groups) * P::G;
    if (kSrc == SRC_BAND_PACKED) {
        const int um = src.unit_map ? __ldg(src.unit_map + unit) : unit;
        const int row = um / src.n_bands;
        RegSrcBand s;
        s.X = (const float2*)src.base + (size_t)row * src.unit_stride;
        s.b = src.bands[um - row * src.n_bands];
        s.freq_scale = src.freq_scale;
        s.M = src.M;
        s.twN = src.twN;
        // no per-row time advance here: the spectra arrive already advanced (bands.cu: untangle_kernel), and
        // c2c_exec sends a SrcDesc with a shift to the generic pass
        reg_cols_stage_a<P>((int)threadIdx.x, s, M2, col0, buf);
    } else if (kSrc == SRC_COMPLEX) {
        RegSrcComplex s{(const float2*)src.base + (size_t)unit * src.unit_stride};
        reg_cols_stage_a<P>((int)threadIdx.x, s, M2, col0, buf);
    } else {  // SRC_REAL_PACKED with even length: z[m] = (x[2m], x[2m+1]) is a plain float2 view of the row
        RegSrcComplex s{(const float2*)((const float*)src.base + (size_t)unit * src.unit_stride)};
        reg_cols_stage_a<P>((int)threadIdx.x, s, M2, col0, buf);
    }
    // stage B's tables are filled after stage A's loads have been issued: their latency hides behind the butterflies
    reg_fill_twiddles<P>((int)threadIdx.x, P::kThreads, tws, tw_full);
    reg_cols_fill_colq<P>((int)threadIdx.x, colq, col0, twM);
    __syncthreads();
    reg_cols_stage_b<P>((int)threadIdx.x, buf, tws, colq, scratch + (size_t)unit * ((size_t)P::L * M2), M2, col0, twM);
}

// (a register cap on this kernel - 56 or 48 per thread through a min-blocks bound - measured 2-6 % slower)
template <class P, int M1, bool kEnergy>
__global__ void __launch_bounds__(P::kThreads) fft_rows_reg(const float2* __restrict__ scratch, float2* __restrict__ out,
                                                            long long out_stride, int n_valid, int conj, float scale,
                                                            const float2* __restrict__ tw_full, float* __restrict__ energy,
                                                            long long energy_stride) {
    static_assert(M1 % P::G == 0, "rows must split evenly over CTAs");
    static_assert(!kEnergy || ((P::G == 8 || P::G == 16) && P::kThreadsB % P::G == 0 && 32 % P::G == 0),
                  "energy partials reduce over aligned groups of 8 or 16 lanes");
    constexpr int groups = M1 / P::G;
    extern __shared__ float2 smem[];
    float2* buf = smem;
    float2* tws = smem + P::kBufElems;
    const int unit = blockIdx.x / groups;
    const int rg = blockIdx.x - unit * groups;
    const int row0 = rg * P::G;
    reg_rows_stage_a<P>((int)threadIdx.x, scratch + (size_t)unit * ((size_t)M1 * P::L), row0, buf);
    reg_fill_twiddles<P>((int)threadIdx.x, P::kThreads, tws, tw_full);  // after stage A's loads are in flight
    __syncthreads();
    if (kEnergy) {
        float e[P::RB] = {};
        reg_rows_stage_b<P>((int)threadIdx.x, buf, tws, out + (size_t)unit * out_stride, M1, row0, n_valid, conj, scale, e);
        // sum over the G rows of the group: g is the low 3 (4) bits of the lane index
        float* __restrict__ dst = energy + (size_t)unit * energy_stride + (size_t)rg * P::L;
        const int jb = (int)threadIdx.x / P::G, g = (int)threadIdx.x - jb * P::G;
#pragma unroll
        for (int q = 0; q < P::RB; ++q) {
            float v = e[q];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (P::G == 16) v += __shfl_xor_sync(0xffffffffu, v, 8);
            if (g == 0 && threadIdx.x < P::kThreadsB) dst[jb + q * P::RA] = v;
        }
    } else {
        reg_rows_stage_b<P>((int)threadIdx.x, buf, tws, out + (size_t)unit * out_stride, M1, row0, n_valid, conj, scale);
    }
}

#ifndef AA_ROWS_PLAIN_G
#define AA_ROWS_PLAIN_G 4   /* 96-thread CTAs, 12 per SM: 5.53 -> 5.11 ms per 8192 forward transforms against groups of 8 */
#endif

#ifndef AA_UNTANGLE_G
#define AA_UNTANGLE_G 4     /* rows per group of the fused pass 2 + untangle kernel */
#endif

// forward real transform: pass 2 + untangle in one kernel (fft_reg.cuh).  Threads [0, kThreads) own group A,
// [kThreads, 2 kThreads) the mirrored group B; CTA 0 of a transform handles the self-paired rows 0 and M1 / 2.
template <class P, int M1>
__global__ void __launch_bounds__(2 * P::kThreads) fft_rows_untangle_reg(const float2* __restrict__ scratch, float2* __restrict__ X,
                                                                         long long x_stride, const float2* __restrict__ tw_full,
                                                                         TwoLevelTwiddle twN, const int* __restrict__ advance) {
    static_assert(M1 % 2 == 0 && M1 / 2 >= P::G, "mirrored row groups need an even M1");
    constexpr int half_groups = (M1 / 2 - 1 + P::G - 1) / P::G;
    constexpr int groups = half_groups + 1;
    extern __shared__ float2 smem[];
    float2* bufA = smem;
    float2* bufB = smem + P::kBufElems;
    float2* tws = bufB + P::kBufElems;
    float2* w2 = tws + P::L;
    float2* a2 = w2 + P::L;
    const int unit = blockIdx.x / groups;
    const int c = blockIdx.x - unit * groups;
    const int half = (int)threadIdx.x / P::kThreads;
    const int t = (int)threadIdx.x - half * P::kThreads;
    const float2* scratch_unit = scratch + (size_t)unit * ((size_t)M1 * P::L);
    UntangleOut u;
    u.X = X + (size_t)unit * x_stride;
    u.M1 = M1;
    u.M2 = P::L;
    u.twN = twN;
    u.advance = advance ? advance[unit] : 0;
    u.w2 = w2;
    u.a2 = a2;
    static_assert((2 * P::kThreads) % P::G == 0, "a thread keeps its row slot across the untangle loop");
    const int rA0 = 1 + (c - 1) * P::G, rB0 = M1 - rA0 - P::G + 1;
    if (c == 0) {
        if (half == 0) reg_rows_stage_a_rows<P>(t, scratch_unit, RowsSelfPaired{M1 / 2}, bufA);
    } else if (half == 0) {
        reg_rows_stage_a_rows<P>(t, scratch_unit, RowsFrom{rA0}, bufA);
    } else {
        reg_rows_stage_a_rows<P>(t, scratch_unit, RowsFrom{rB0}, bufB);
    }
    reg_fill_twiddles<P>((int)threadIdx.x, 2 * P::kThreads, tws, tw_full);
    reg_untangle_fill_tables<P>((int)threadIdx.x, 2 * P::kThreads, w2, a2, u);
    __syncthreads();
    const bool active = t < P::kThreadsB && (c != 0 || half == 0);
    float2 v[P::RB];
    if (active) reg_rows_stage_b_regs<P>(t, half ? bufB : bufA, tws, v);
    __syncthreads();
    if (active) reg_rows_store_local<P>(t, v, half ? bufB : bufA);
    __syncthreads();
    if (c == 0) reg_untangle_special<P>((int)threadIdx.x, 2 * P::kThreads, bufA, u);
    else reg_untangle_general<P>((int)threadIdx.x, 2 * P::kThreads, bufA, bufB, u, rA0);
}

bool launch_rows_untangle_reg(const C2CPlan* p, const float2* scratch, const DstDesc& dst, long long units, cudaStream_t stream,
                              bool dry_run) {
#define X(ra1, rb1, g1, ra2, rb2, g2)                                                                         \
    if (p->M1 == ra1 * rb1 && p->M2 == ra2 * rb2) {                                                           \
        constexpr int m1 = ra1 * rb1;                                                                         \
        constexpr int gp = (m1 % AA_UNTANGLE_G == 0) ? AA_UNTANGLE_G : g2;                                    \
        if constexpr (m1 % 2 == 0 && m1 / 2 >= gp) {                                                          \
            using PP = RegCfg<ra2, rb2, gp>;                                                                  \
            if (dry_run) return true;                                                                         \
            constexpr int groups = (m1 / 2 - 1 + gp - 1) / gp + 1;                                            \
            constexpr size_t smem = sizeof(float2) * (size_t)(2 * PP::kBufElems + 3 * PP::L);                     \
            auto k = fft_rows_untangle_reg<PP, m1>;                                                           \
            static bool once = false;                                                                         \
            if (!once) {                                                                                      \
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
                once = true;                                                                                  \
            }                                                                                                 \
            k<<<(unsigned)(units * groups), 2 * PP::kThreads, smem, stream>>>(scratch, dst.untangle_X, dst.untangle_stride, \
                                                                              p->s2.dev.tw, dst.untangle_twN,  \
                                                                              dst.untangle_advance);          \
            return true;                                                                                      \
        }                                                                                                     \
        return false;                                                                                         \
    }
    AA_REG_LIST(X)
#undef X
    return false;
}

template <int kSrc>
bool launch_cols_reg(const C2CPlan* p, const SrcDesc& src, float2* scratch, long long units, cudaStream_t stream) {
#define X(ra1, rb1, g1, ra2, rb2, g2)                                                                         \
    if (p->M1 == ra1 * rb1 && p->M2 == ra2 * rb2) {                                                           \
        using P = RegCfg<ra1, rb1, g1>;                                                                       \
        auto k = fft_cols_reg<P, kSrc, ra2 * rb2>;                                                            \
        static bool once = false;                                                                             \
        if (!once) {                                                                                          \
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::kSmemBytes);         \
            once = true;                                                                                      \
        }                                                                                                     \
        k<<<(unsigned)(units * ((ra2 * rb2) / g1)), P::kThreads, P::kSmemBytes, stream>>>(src, scratch, p->twM.dev, \
                                                                                         p->s1.dev.tw);       \
        return true;                                                                                          \
    }
    AA_REG_LIST(X)
#undef X
    return false;
}

bool launch_rows_reg(const C2CPlan* p, const float2* scratch, const DstDesc& dst, long long units, cudaStream_t stream) {
#define X(ra1, rb1, g1, ra2, rb2, g2)                                                                         \
    if (p->M1 == ra1 * rb1 && p->M2 == ra2 * rb2) {                                                           \
        using P = RegCfg<ra2, rb2, g2>;                                                                       \
        /* plain variant (forward transforms, band signals on request): AA_ROWS_PLAIN_G rows per CTA */      \
        constexpr int gp = ((ra1 * rb1) % AA_ROWS_PLAIN_G == 0) ? AA_ROWS_PLAIN_G : g2;                       \
        using PP = RegCfg<ra2, rb2, gp>;                                                                      \
        static bool once = false;                                                                             \
        if (!once) {                                                                                          \
            cudaFuncSetAttribute(fft_rows_reg<PP, ra1 * rb1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)PP::kSmemBytes);                                                        \
            cudaFuncSetAttribute(fft_rows_reg<P, ra1 * rb1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)P::kSmemBytes);                                                         \
            once = true;                                                                                      \
        }                                                                                                     \
        if (dst.energy)                                                                                       \
            fft_rows_reg<P, ra1 * rb1, true><<<(unsigned)(units * ((ra1 * rb1) / g2)), P::kThreads, P::kSmemBytes, stream>>>( \
                scratch, (float2*)dst.base, dst.unit_stride, dst.n_valid, dst.conj1, dst.scale1, p->s2.dev.tw, \
                dst.energy, dst.energy_stride);                                                               \
        else                                                                                                  \
            fft_rows_reg<PP, ra1 * rb1, false><<<(unsigned)(units * ((ra1 * rb1) / gp)), PP::kThreads, PP::kSmemBytes, stream>>>( \
                scratch, (float2*)dst.base, dst.unit_stride, dst.n_valid, dst.conj1, dst.scale1, p->s2.dev.tw, \
                nullptr, 0);                                                                                  \
        return true;                                                                                          \
    }
    AA_REG_LIST(X)
#undef X
    return false;
}

template <typename K>
int set_smem(K kernel) {
    AA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemHard));
    return AA_OK;
}

int configure_kernels_once() {
    static bool done = false;
    if (done) return AA_OK;
    int rc;
    if ((rc = set_smem(fft_pass1_kernel<SRC_COMPLEX>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass1_kernel<SRC_REAL_PACKED>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass1_kernel<SRC_BAND_PACKED>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass1_kernel<-1>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass2_kernel<SRC_COMPLEX, DST_COMPLEX>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass2_kernel<SRC_COMPLEX, -1>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass2_kernel<-1, DST_COMPLEX>)) != AA_OK) return rc;
    if ((rc = set_smem(fft_pass2_kernel<-1, -1>)) != AA_OK) return rc;
    done = true;
    return AA_OK;
}

bool plain_complex_dst(const DstDesc& d) {
    return d.mode == DST_COMPLEX && d.chirp == nullptr && d.conj2 == 0 && d.scale2 == 1.0f;
}

int exec_direct(const C2CPlan* p, const SrcDesc& src, const DstDesc& dst, long long units, float2* scratch,
                cudaStream_t stream) {
    int rc = configure_kernels_once();
    if (rc != AA_OK) return rc;
    if (units <= 0) return AA_OK;
    PassArgs pa{};
    pa.M1 = p->M1;
    pa.M2 = p->M2;
    pa.twM = p->twM.dev;
    SrcDesc src2 = src;
    const bool fast_src = (src.chirp == nullptr);
    if (p->M1 > 1) {
        AA_REQUIRE(scratch != nullptr, AA_ERR_WORKSPACE, "fft scratch missing");
        pa.sp = p->s1.dev;
        pa.G = p->G1;
        pa.G_rcp = fast_rcp((unsigned)p->G1);
        pa.Lp = p->Lp1;
        pa.groups = (p->M2 + p->G1 - 1) / p->G1;
        long long grid = units * pa.groups;
        AA_REQUIRE(grid < 2147483647LL, AA_ERR_INVALID_ARGUMENT, "fft grid too large");
        const unsigned g = (unsigned)grid;
        const bool use_spec = getenv("AA_FFT_GENERIC") == nullptr && units * (long long)p->M < 2147483647LL;
        ProfScope prof(src.mode == SRC_BAND_PACKED || src.mode == SRC_BAND_FULL ? "fft_cols_band" : "fft_cols", stream);
        if (use_spec && fast_src && src.mode == SRC_BAND_PACKED && src.conj && src.bands_have_tables &&
            src.shift == nullptr && launch_cols_reg<SRC_BAND_PACKED>(p, src, scratch, units, stream)) {
        } else if (use_spec && fast_src && src.mode == SRC_REAL_PACKED && !src.clip && src.n_valid == 2 * p->M &&
                   (src.unit_stride % 2) == 0 && launch_cols_reg<SRC_REAL_PACKED>(p, src, scratch, units, stream)) {
        } else if (use_spec && fast_src && src.mode == SRC_COMPLEX && !src.conj && src.n_valid >= p->M &&
                   launch_cols_reg<SRC_COMPLEX>(p, src, scratch, units, stream)) {
        } else if (fast_src && src.mode == SRC_COMPLEX)
            fft_pass1_kernel<SRC_COMPLEX><<<g, kPassThreads, p->smem1, stream>>>(pa, src, scratch, (long long)p->M);
        else if (fast_src && src.mode == SRC_REAL_PACKED)
            fft_pass1_kernel<SRC_REAL_PACKED><<<g, kPassThreads, p->smem1, stream>>>(pa, src, scratch, (long long)p->M);
        else if (fast_src && src.mode == SRC_BAND_PACKED)
            fft_pass1_kernel<SRC_BAND_PACKED><<<g, kPassThreads, p->smem1, stream>>>(pa, src, scratch, (long long)p->M);
        else
            fft_pass1_kernel<-1><<<g, kPassThreads, p->smem1, stream>>>(pa, src, scratch, (long long)p->M);
        rc = launch_check("fft_pass1_kernel");
        if (rc != AA_OK) return rc;
        src2 = make_src_complex(scratch, p->M, p->M, 0);
    }
    pa.sp = p->s2.dev;
    pa.G = p->G2;
    pa.G_rcp = fast_rcp((unsigned)p->G2);
    pa.Lp = p->Lp2;
    pa.groups = (p->M1 + p->G2 - 1) / p->G2;
    long long grid = units * pa.groups;
    AA_REQUIRE(grid < 2147483647LL, AA_ERR_INVALID_ARGUMENT, "fft grid too large");
    const unsigned g = (unsigned)grid;
    const bool src_plain = (src2.mode == SRC_COMPLEX && src2.chirp == nullptr);
    const bool dst_plain = plain_complex_dst(dst);
    if (dst.untangle_X != nullptr) {   // forward real transform: pass 2 and the untangle step in one kernel
        ProfScope prof("fft_rows_untangle", stream);
        AA_REQUIRE(p->M1 > 1 && src_plain && dst.conj1 == 0 && dst.scale1 == 1.0f && getenv("AA_FFT_GENERIC") == nullptr &&
                       launch_rows_untangle_reg(p, scratch, dst, units, stream, false),
                   AA_ERR_UNSUPPORTED, "no fused untangle pass for this plan (see c2c_can_untangle)");
        return launch_check("fft_rows_untangle_reg");
    }
    ProfScope prof(dst.energy ? "fft_rows_band" : "fft_rows", stream);
    const bool spec2 = getenv("AA_FFT_GENERIC") == nullptr && p->M1 > 1 && src_plain && dst_plain &&
                       units * (long long)p->M < 2147483647LL &&
                       launch_rows_reg(p, scratch, dst, units, stream);
    if (!spec2 && dst.energy != nullptr) {
        set_error("energy partials need the register-pass rows kernel");
        return AA_ERR_UNSUPPORTED;
    }
    if (spec2) {
    } else if (src_plain && dst_plain)
        fft_pass2_kernel<SRC_COMPLEX, DST_COMPLEX><<<g, kPassThreads, p->smem2, stream>>>(pa, src2, dst);
    else if (src_plain)
        fft_pass2_kernel<SRC_COMPLEX, -1><<<g, kPassThreads, p->smem2, stream>>>(pa, src2, dst);
    else if (dst_plain)
        fft_pass2_kernel<-1, DST_COMPLEX><<<g, kPassThreads, p->smem2, stream>>>(pa, src2, dst);
    else
        fft_pass2_kernel<-1, -1><<<g, kPassThreads, p->smem2, stream>>>(pa, src2, dst);
    return launch_check("fft_pass2_kernel");
}

int build_direct(int M, C2CPlan* p) {
    HostC2C h;
    if (!host_direct_plan(M, &h)) {
        set_error("no direct FFT plan for length %d", M);
        return AA_ERR_UNSUPPORTED;
    }
    p->M = M;
    p->bluestein = false;
    p->M1 = h.M1;
    p->M2 = h.M2;
    int rc;
    if (p->M1 > 1) {
        if ((rc = upload_subplan(h.s1, &p->s1)) != AA_OK) return rc;
        choose_group(p->M1, &p->G1, &p->Lp1, &p->smem1);
        if ((rc = upload_two_level(h.twM, &p->twM)) != AA_OK) return rc;
    }
    if ((rc = upload_subplan(h.s2, &p->s2)) != AA_OK) return rc;
    choose_group(p->M2, &p->G2, &p->Lp2, &p->smem2);
    return AA_OK;
}

}  // namespace

bool is_smooth(long long n) { return is_smooth_host(n); }

bool c2c_can_untangle(const C2CPlan* p) {
    if (!p || p->bluestein || p->M1 <= 1 || getenv("AA_FFT_GENERIC") != nullptr || getenv("AA_NO_FUSED_UNTANGLE") != nullptr)
        return false;
    if ((long long)p->M2 * 2LL * (long long)p->M >= 2147483647LL) return false;   // 32-bit phase arithmetic
    return launch_rows_untangle_reg(p, nullptr, DstDesc{}, 0, nullptr, true);
}

// Tiles of the energy partials a DstDesc may request from c2c_exec: rows are transformed in groups of 8, so a
// transform's outputs fall into M2 tiles of M1 consecutive complex values and M1 / G2 partial sums per tile.
bool c2c_energy_layout(const C2CPlan* p, int* tile_complex, int* n_tiles, int* n_groups) {
    if (p->bluestein || p->M1 <= 1 || getenv("AA_FFT_GENERIC") != nullptr) return false;
#define X(ra1, rb1, g1, ra2, rb2, g2)                                              \
    if (p->M1 == (ra1) * (rb1) && p->M2 == (ra2) * (rb2) && ((g2) == 8 || (g2) == 16)) { \
        *tile_complex = p->M1;                                                     \
        *n_tiles = p->M2;                                                          \
        *n_groups = p->M1 / (g2);                                                  \
        return true;                                                               \
    }
    AA_REG_LIST(X)
#undef X
    return false;
}

int build_two_level(long long len, TwoLevelTable* out) {
    HostTwoLevel h;
    host_two_level(len, &h);
    return upload_two_level(h, out);
}

void free_two_level(TwoLevelTable* t) {
    if (t->d_hi) cudaFree(t->d_hi);
    if (t->d_lo) cudaFree(t->d_lo);
    t->d_hi = t->d_lo = nullptr;
}

void c2c_plan_destroy(C2CPlan* p) {
    if (!p) return;
    if (p->inner) c2c_plan_destroy(p->inner);
    if (p->s1.d_tw) cudaFree(p->s1.d_tw);
    if (p->s2.d_tw) cudaFree(p->s2.d_tw);
    free_two_level(&p->twM);
    if (p->d_chirp) cudaFree(p->d_chirp);
    if (p->d_H) cudaFree(p->d_H);
    delete p;
}

int c2c_plan_create(int M, C2CPlan** out) {
    AA_REQUIRE(M >= 1, AA_ERR_INVALID_ARGUMENT, "fft length must be positive");
    C2CPlan* p = new C2CPlan();
    int rc = AA_OK;
    if (direct_plan_possible(M)) {
        rc = build_direct(M, p);
        if (rc != AA_OK) {
            c2c_plan_destroy(p);
            return rc;
        }
        *out = p;
        return AA_OK;
    }
    if (bluestein_length(M) > (long long)kMaxSubLen * kMaxSubLen) {
        set_error("length %d too long for the Bluestein path", M);
        c2c_plan_destroy(p);
        return AA_ERR_TOO_LONG;
    }
    HostC2C h;
    host_bluestein_tables(M, &h);
    p->M = M;
    p->bluestein = true;
    const size_t P = (size_t)h.P;
    rc = c2c_plan_create(h.P, &p->inner);
    if (rc != AA_OK) {
        c2c_plan_destroy(p);
        return rc;
    }
    float2* d_hw = nullptr;
    float2* d_scr = nullptr;
    auto fail = [&](int code) {
        if (d_hw) cudaFree(d_hw);
        if (d_scr) cudaFree(d_scr);
        c2c_plan_destroy(p);
        return code;
    };
    if ((rc = upload(h.chirp, &p->d_chirp)) != AA_OK) return fail(rc);
    if ((rc = upload(h.hw, &d_hw)) != AA_OK) return fail(rc);
    if (cudaMalloc(&p->d_H, sizeof(float2) * P) != cudaSuccess) {
        set_error("cudaMalloc failed building Bluestein tables for length %d", M);
        return fail(AA_ERR_CUDA);
    }
    size_t scr = c2c_scratch_elems(p->inner, 1);
    if (scr && cudaMalloc(&d_scr, sizeof(float2) * scr) != cudaSuccess) {
        set_error("cudaMalloc failed building Bluestein tables for length %d", M);
        return fail(AA_ERR_CUDA);
    }
    rc = c2c_exec(p->inner, make_src_complex(d_hw, (long long)P, (int)P, 0),
                  make_dst_complex(p->d_H, (long long)P, (int)P, 0, 1.0f), 1, d_scr, 0);
    if (rc != AA_OK) return fail(rc);
    if (cudaStreamSynchronize(0) != cudaSuccess) {
        set_error("Bluestein table transform failed");
        return fail(AA_ERR_CUDA);
    }
    cudaFree(d_hw);
    if (d_scr) cudaFree(d_scr);
    *out = p;
    return AA_OK;
}

size_t c2c_scratch_elems(const C2CPlan* p, long long units) {
    if (!p->bluestein) return p->M1 > 1 ? (size_t)units * (size_t)p->M : 0;
    size_t P = (size_t)p->inner->M;
    return (size_t)units * P + c2c_scratch_elems(p->inner, units);
}

int c2c_exec(const C2CPlan* p, SrcDesc src, DstDesc dst, long long units, float2* scratch, cudaStream_t stream) {
    if (!p->bluestein) return exec_direct(p, src, dst, units, scratch, stream);
    const int P = p->inner->M;
    float2* mid = scratch;
    float2* inner_scratch = scratch + (size_t)units * (size_t)P;
    // A = FFT_P(x * w, zero padded); mid = A * H
    SrcDesc s1 = src;
    s1.chirp = p->d_chirp;
    s1.chirp_len = p->M;
    DstDesc d1{};
    d1.mode = DST_MUL_TABLE;
    d1.base = mid;
    d1.unit_stride = P;
    d1.n_valid = P;
    d1.scale1 = 1.0f;
    d1.scale2 = 1.0f;
    d1.table = p->d_H;
    int rc = c2c_exec(p->inner, s1, d1, units, inner_scratch, stream);
    if (rc != AA_OK) return rc;
    // X = w * IFFT_P(mid), first M outputs; then the caller's conj / scale
    SrcDesc s2 = make_src_complex(mid, P, P, 1);
    DstDesc d2 = dst;
    d2.conj1 = 1;
    d2.scale1 = 1.0f / (float)P;
    d2.chirp = p->d_chirp;
    d2.conj2 = dst.conj1;
    d2.scale2 = dst.scale1 * dst.scale2;
    d2.n_valid = dst.n_valid < p->M ? dst.n_valid : p->M;
    return c2c_exec(p->inner, s2, d2, units, inner_scratch, stream);
}

}  // namespace aa
