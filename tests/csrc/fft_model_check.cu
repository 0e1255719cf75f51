// CPU check of the FFT engine's index arithmetic.
//
// Runs the SAME host/device functions the CUDA kernels use (fft_core.cuh butterflies,
// fft_access.cuh accessors, fft_plan_host.hpp planner) in a serial emulation of
// fft_pass1_kernel / fft_pass2_kernel / the Bluestein sequence, and compares with
// direct DFT sums in double precision.  No GPU is needed: nvcc compiles this as a
// host program.  Exit code 0 = all checks passed.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <complex>
#include <vector>

#include "../../audio_analysis_b200/csrc/fft_access.cuh"
#include "../../audio_analysis_b200/csrc/fft_plan_host.hpp"
#include "../../audio_analysis_b200/csrc/fft_fused.cuh"
#include "../../audio_analysis_b200/csrc/fft_reg.cuh"

using namespace aa;
typedef std::complex<double> cd;

static unsigned long long rng_state = 0x9E3779B97F4A7C15ULL;
static double urand() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

static TwoLevelTwiddle dev_view(const HostTwoLevel& t) {
    TwoLevelTwiddle d;
    d.hi = t.hi.data();
    d.lo = t.lo.data();
    d.shift = t.shift;
    return d;
}

// compact per-stage twiddles exactly as load_stage_twiddles / spec_twiddles build them on the device
static std::vector<float2> compact_table(int L, const int* radix, int nst, const float2* tw) {
    std::vector<float2> t((size_t)L + 1, make_float2(0, 0));
    int Ns = 1;
    for (int s = 0; s < nst; ++s) {
        const int R = radix[s], tmul = L / (Ns * R);
        for (int m = 0; m < Ns * (R - 1); ++m) {
            const int r1 = m / Ns, k = m % Ns;
            t[(size_t)compact_twiddle_offset(Ns) + m] = tw[(r1 + 1) * k * tmul];
        }
        Ns *= R;
    }
    return t;
}

template <int R, int kPad>
static void emu_stage(const SubPlanDev& d, int s, const float2* a, float2* b, const float2* compact) {
    // mirrors stage_loop<R> in fft.cu: multiply-high division, compact twiddles
    for (int e = 0; e < d.nb[s]; ++e) {
        const int g = (int)fast_div((unsigned)e, d.nb_rcp[s]);
        const int j = e - g * d.nb[s];
        const int k = j - (int)fast_div((unsigned)j, d.ns_rcp[s]) * d.Ns[s];
        if (g != 0) { printf("fast_div broke at e=%d\n", e); exit(2); }
        stockham_butterfly_compact<R, kPad>(j, k, d.nb[s], d.L, d.Ns[s], a, b, compact + compact_twiddle_offset(d.Ns[s]),
                                            d.tw);
    }
}

template <int kPad>
static void emu_chain(SubPlanDev& d, std::vector<float2>& a, std::vector<float2>& b) {
    fill_stage_constants(&d);
    std::vector<float2> compact = compact_table(d.L, d.radix, d.nst, d.tw);
    for (int s = 0; s < d.nst; ++s) {
        switch (d.radix[s]) {
            case 2: emu_stage<2, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 3: emu_stage<3, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 4: emu_stage<4, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 5: emu_stage<5, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 6: emu_stage<6, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 7: emu_stage<7, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 8: emu_stage<8, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 11: emu_stage<11, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            case 13: emu_stage<13, kPad>(d, s, a.data(), b.data(), compact.data()); break;
            default: printf("bad radix\n"); exit(2);
        }
        a.swap(b);
    }
}

static void emu_stages(const HostSubPlan& sp, std::vector<float2>& a, std::vector<float2>& b) {
    SubPlanDev d{};
    d.L = sp.L;
    d.nst = sp.nst;
    for (int s = 0; s < sp.nst; ++s) d.radix[s] = sp.radix[s];
    d.tw = sp.tw.data();
    emu_chain<0>(d, a, b);
    static bool div_checked = false;
    if (!div_checked) {  // exhaustive check of the multiply-high division for every index the kernels can form
        div_checked = true;
        for (unsigned dd = 1; dd <= 4096; dd += (dd < 64 ? 1 : 37)) {
            unsigned r = fast_rcp(dd);
            for (unsigned n = 0; n < 70000; n += 13)
                if (fast_div(n, r) != n / dd) { printf("fast_div(%u,%u) wrong\n", n, dd); exit(2); }
        }
    }
}

// ---- serial emulation of fft_cols_reg / fft_rows_reg (fft.cu): every thread of every CTA, stage A then stage B ----
static constexpr int emu_reg_pitch(int G, int min_elems) {  // mirrors reg_pitch() in fft.cu
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

template <int RA1, int RB1, int G1, int RA2, int RB2, int G2>
static void emu_reg_unit(const HostC2C& p, const SrcDesc& src, const DstDesc& dst, long long u, int* used) {
    using P1 = RegPass<RA1, RB1, G1, emu_reg_pitch(G1, RB1 * (RA1 | 1))>;
    using P2 = RegPass<RA2, RB2, G2, emu_reg_pitch(G2, RB2 * (RA2 | 1))>;
    const int M1 = P1::L, M2 = P2::L;
    std::vector<float2> scratch((size_t)M1 * M2), buf, tws, colq((size_t)P1::kColQElems);
    TwoLevelTwiddle twM = dev_view(p.twM);
    buf.assign((size_t)P1::kBufElems, make_float2(0, 0));
    tws.assign((size_t)P1::L, make_float2(0, 0));
    for (int t = 0; t < P1::kThreads; ++t) reg_fill_twiddles<P1>(t, P1::kThreads, tws.data(), p.s1.tw.data());
    for (int cta = 0; cta < M2 / G1; ++cta) {
        const int col0 = cta * G1;
        for (int t = 0; t < P1::kThreads; ++t) {
            if (src.mode == SRC_BAND_PACKED) {
                const long long row = u / src.n_bands;
                RegSrcBand s;
                s.X = (const float2*)src.base + row * src.unit_stride;
                s.b = src.bands[u - row * src.n_bands];
                s.freq_scale = src.freq_scale;
                s.M = src.M;
                s.twN = src.twN;
                s.shift = src.shift ? src.shift[row] : 0;
                s.shift_N = src.shift_N;
                s.shift_rcp = src.shift_N > 0 ? 1.0 / (double)src.shift_N : 0.0;
                s.twShift = src.twShift;
                reg_cols_stage_a<P1>(t, s, M2, col0, buf.data());
            } else if (src.mode == SRC_REAL_PACKED) {
                RegSrcComplex s{(const float2*)((const float*)src.base + u * src.unit_stride)};
                reg_cols_stage_a<P1>(t, s, M2, col0, buf.data());
            } else {
                RegSrcComplex s{(const float2*)src.base + u * src.unit_stride};
                reg_cols_stage_a<P1>(t, s, M2, col0, buf.data());
            }
        }
        for (int t = 0; t < P1::kThreads; ++t) reg_cols_fill_colq<P1>(t, colq.data(), col0, twM);
        for (int t = 0; t < P1::kThreads; ++t)
            reg_cols_stage_b<P1>(t, buf.data(), tws.data(), colq.data(), scratch.data(), M2, col0, twM);
    }
    buf.assign((size_t)P2::kBufElems, make_float2(0, 0));
    tws.assign((size_t)P2::L, make_float2(0, 0));
    for (int t = 0; t < P2::kThreads; ++t) reg_fill_twiddles<P2>(t, P2::kThreads, tws.data(), p.s2.tw.data());
    for (int cta = 0; cta < M1 / G2; ++cta) {
        const int row0 = cta * G2;
        for (int t = 0; t < P2::kThreads; ++t) reg_rows_stage_a<P2>(t, scratch.data(), row0, buf.data());
        for (int t = 0; t < P2::kThreads; ++t) {
            float e[P2::RB];
            reg_rows_stage_b<P2>(t, buf.data(), tws.data(), (float2*)dst.base + u * dst.unit_stride, M1, row0, dst.n_valid,
                                 dst.conj1, dst.scale1, dst.energy ? e : nullptr);
            if (dst.energy && t < P2::kThreadsB) {  // the kernel reduces over the 8 rows of the group with shuffles
                const int jb = t / P2::G;
                for (int q = 0; q < P2::RB; ++q)
                    dst.energy[u * dst.energy_stride + (long long)cta * P2::L + jb + q * P2::RA] += e[q];
            }
        }
    }
    *used = 1;
}

// mirrors the dispatch in exec_direct(): register passes for the listed splits and plain sources / sinks
static bool emu_reg_direct(const HostC2C& p, const SrcDesc& src, const DstDesc& dst, long long units) {
    const bool src_ok = src.chirp == nullptr &&
                        ((src.mode == SRC_BAND_PACKED && src.conj && src.bands_have_tables) ||
                         (src.mode == SRC_REAL_PACKED && !src.clip && src.n_valid == 2 * p.M && src.unit_stride % 2 == 0) ||
                         (src.mode == SRC_COMPLEX && !src.conj && src.n_valid >= p.M));
    const bool dst_ok = dst.mode == DST_COMPLEX && dst.chirp == nullptr && dst.conj2 == 0 && dst.scale2 == 1.0f;
    if (!src_ok || !dst_ok) return false;
    int used = 0;
#define X(ra1, rb1, g1, ra2, rb2, g2)                                                         \
    if (!used && p.M1 == (ra1) * (rb1) && p.M2 == (ra2) * (rb2))                              \
        for (long long u = 0; u < units; ++u) emu_reg_unit<ra1, rb1, g1, ra2, rb2, g2>(p, src, dst, u, &used);
    AA_REG_LIST(X)
#undef X
    return used != 0;
}

static void emu_direct(const HostC2C& p, const SrcDesc& src, const DstDesc& dst, long long units) {
    if (p.M1 > 1 && emu_reg_direct(p, src, dst, units)) return;
    std::vector<float2> scratch;
    for (long long u = 0; u < units; ++u) {
        SrcDesc src2 = src;
        long long u2 = u;
        if (p.M1 > 1) {
            scratch.assign((size_t)p.M, make_float2(0, 0));
            TwoLevelTwiddle tw = dev_view(p.twM);
            std::vector<float2> a((size_t)p.M1), b((size_t)p.M1);
            for (int col = 0; col < p.M2; ++col) {
                for (int i = 0; i < p.M1; ++i) a[(size_t)i] = src_get(src, u, (int)pass1_src_index(p.M2, col, i));
                emu_stages(p.s1, a, b);
                for (int i = 0; i < p.M1; ++i)
                    scratch[(size_t)pass1_dst_index(p.M2, col, i)] = cmul(a[(size_t)i], twiddle2(tw, (long long)col * i));
            }
            src2 = SrcDesc{};
            src2.mode = SRC_COMPLEX;
            src2.base = scratch.data();
            src2.unit_stride = 0;
            src2.n_valid = p.M;
            u2 = 0;
        }
        std::vector<float2> a((size_t)p.M2), b((size_t)p.M2);
        for (int row = 0; row < p.M1; ++row) {
            for (int i = 0; i < p.M2; ++i) a[(size_t)i] = src_get(src2, u2, (int)pass2_src_index(p.M2, row, i));
            emu_stages(p.s2, a, b);
            for (int i = 0; i < p.M2; ++i) dst_put(dst, u, (int)pass2_dst_index(p.M1, row, i), a[(size_t)i]);
        }
    }
}

struct EmuPlan {
    HostC2C top;
    HostC2C inner;  // Bluestein: length P
    std::vector<float2> H;
};

static SrcDesc src_complex(const float2* base, long long stride, int n_valid, int conj) {
    SrcDesc s{};
    s.mode = SRC_COMPLEX;
    s.base = base;
    s.unit_stride = stride;
    s.n_valid = n_valid;
    s.conj = conj;
    return s;
}
static DstDesc dst_complex(float2* base, long long stride, int n_valid, int conj, float scale) {
    DstDesc d{};
    d.mode = DST_COMPLEX;
    d.base = base;
    d.unit_stride = stride;
    d.n_valid = n_valid;
    d.conj1 = conj;
    d.scale1 = scale;
    d.scale2 = 1.0f;
    return d;
}

static bool emu_plan(int M, EmuPlan* ep) {
    if (direct_plan_possible(M)) return host_direct_plan(M, &ep->top);
    host_bluestein_tables(M, &ep->top);
    if (!host_direct_plan(ep->top.P, &ep->inner)) return false;
    ep->H.assign((size_t)ep->top.P, make_float2(0, 0));
    emu_direct(ep->inner, src_complex(ep->top.hw.data(), ep->top.P, ep->top.P, 0),
               dst_complex(ep->H.data(), ep->top.P, ep->top.P, 0, 1.0f), 1);
    return true;
}

// mirrors c2c_exec in fft.cu
static void emu_exec(const EmuPlan& ep, SrcDesc src, DstDesc dst, long long units) {
    if (!ep.top.bluestein) {
        emu_direct(ep.top, src, dst, units);
        return;
    }
    const int P = ep.top.P;
    std::vector<float2> mid((size_t)units * P);
    SrcDesc s1 = src;
    s1.chirp = ep.top.chirp.data();
    s1.chirp_len = ep.top.M;
    DstDesc d1{};
    d1.mode = DST_MUL_TABLE;
    d1.base = mid.data();
    d1.unit_stride = P;
    d1.n_valid = P;
    d1.scale1 = 1.0f;
    d1.scale2 = 1.0f;
    d1.table = ep.H.data();
    emu_direct(ep.inner, s1, d1, units);
    SrcDesc s2 = src_complex(mid.data(), P, P, 1);
    DstDesc d2 = dst;
    d2.conj1 = 1;
    d2.scale1 = 1.0f / (float)P;
    d2.chirp = ep.top.chirp.data();
    d2.conj2 = dst.conj1;
    d2.scale2 = dst.scale1 * dst.scale2;
    d2.n_valid = dst.n_valid < ep.top.M ? dst.n_valid : ep.top.M;
    emu_direct(ep.inner, s2, d2, units);
}

static int failures = 0;
static void report(const char* what, int n, double err, double tol) {
    bool ok = err <= tol;
    printf("%-34s n=%-8d rel_err=%.3e  %s\n", what, n, err, ok ? "ok" : "FAIL");
    if (!ok) ++failures;
}

// ---- 1. complex transforms ---------------------------------------------------------------
static void check_c2c(int M) {
    EmuPlan ep;
    if (!emu_plan(M, &ep)) {
        printf("plan failed for %d\n", M);
        ++failures;
        return;
    }
    const int units = 2;
    std::vector<float2> x((size_t)units * M), X((size_t)units * M);
    double norm = 0;
    for (auto& v : x) {
        v = make_float2((float)urand(), (float)urand());
        norm += (double)v.x * v.x + (double)v.y * v.y;
    }
    norm = sqrt(norm / units);
    emu_exec(ep, src_complex(x.data(), M, M, 0), dst_complex(X.data(), M, M, 0, 1.0f), units);
    // spot check bins by direct summation
    int nchk = M <= 512 ? M : 48;
    double worst = 0;
    for (int c = 0; c < nchk; ++c) {
        int k = M <= 512 ? c : (int)((urand() * 0.5 + 0.5) * (M - 1));
        for (int u = 0; u < units; ++u) {
            cd acc = 0;
            for (int n = 0; n < M; ++n) {
                double ang = -2.0 * M_PI * (double)(((long long)n * k) % M) / M;
                acc += cd(x[(size_t)u * M + n].x, x[(size_t)u * M + n].y) * cd(cos(ang), sin(ang));
            }
            double e = std::abs(acc - cd(X[(size_t)u * M + k].x, X[(size_t)u * M + k].y)) / norm;
            if (e > worst) worst = e;
        }
    }
    char name[64];
    snprintf(name, sizeof(name), "c2c %s M1=%d M2=%d", ep.top.bluestein ? "bluestein" : "direct",
             ep.top.bluestein ? ep.inner.M1 : ep.top.M1, ep.top.bluestein ? ep.inner.M2 : ep.top.M2);
    report(name, M, worst, ep.top.bluestein ? 3e-5 : 5e-6);
    // inverse through the conj trick must return the input
    std::vector<float2> back((size_t)units * M);
    emu_exec(ep, src_complex(X.data(), M, M, 1), dst_complex(back.data(), M, M, 1, 1.0f / M), units);
    double e2 = 0;
    for (size_t i = 0; i < back.size(); ++i) {
        double d = hypot((double)back[i].x - x[i].x, (double)back[i].y - x[i].y);
        if (d > e2) e2 = d;
    }
    report("  round trip", M, e2, ep.top.bluestein ? 1e-4 : 2e-5);
}

// ---- 2. real forward + band-masked real inverse (the rt60bands filter) -------------------
static BandDev make_bandpass(double low, double high, double tw_oct, double nyq, double freq_scale, int nbins) {
    BandDev b{};
    b.kind = AA_BAND_BANDPASS;
    double fac = pow(2.0, tw_oct);
    double hp_pass = low, hp_stop = fmax(1.0, low / fac);
    double lp_pass = high, lp_stop = fmin(nyq, high * fac);
    b.hp_stop = (float)hp_stop;
    b.hp_pass = (float)hp_pass;
    b.hp_span = (float)(hp_pass - hp_stop);
    b.lp_pass = (float)lp_pass;
    b.lp_stop = (float)lp_stop;
    b.lp_span = (float)(lp_stop - lp_pass);
    b.k_first = nbins;
    b.k_last = -1;
    BandDev open = b;
    open.k_first = 0;
    open.k_last = nbins - 1;
    for (int k = 0; k < nbins; ++k) {
        float f = (float)((double)k * freq_scale);
        if (band_mask_value(open, f) != 0.0f) {
            if (k < b.k_first) b.k_first = k;
            b.k_last = k;
        }
    }
    return b;
}

// serial emulation of fft_rows_untangle_reg (fft.cu): forward real transform, pass 1 as above, pass 2 fused with the
// untangle step.  x: N real samples; X: M + 1 half-spectrum bins, stored time-advanced by `advance` samples.
template <int RA1, int RB1, int G1, int RA2, int RB2, int GP>
static void emu_fused_forward(const HostC2C& p, const float* x, int N, float2* X, const HostTwoLevel& twN_h, int advance) {
    using P1 = RegPass<RA1, RB1, G1, emu_reg_pitch(G1, RB1 * (RA1 | 1))>;
    using P2 = RegPass<RA2, RB2, GP, emu_reg_pitch(GP, RB2 * (RA2 | 1))>;
    const int M1 = P1::L, M2 = P2::L;
    std::vector<float2> scratch((size_t)M1 * M2), buf((size_t)P1::kBufElems), tws((size_t)P1::L), colq((size_t)P1::kColQElems);
    TwoLevelTwiddle twM = dev_view(p.twM);
    for (int t = 0; t < P1::kThreads; ++t) reg_fill_twiddles<P1>(t, P1::kThreads, tws.data(), p.s1.tw.data());
    for (int cta = 0; cta < M2 / G1; ++cta) {
        const int col0 = cta * G1;
        RegSrcComplex s{(const float2*)x};
        for (int t = 0; t < P1::kThreads; ++t) reg_cols_stage_a<P1>(t, s, M2, col0, buf.data());
        for (int t = 0; t < P1::kThreads; ++t) reg_cols_fill_colq<P1>(t, colq.data(), col0, twM);
        for (int t = 0; t < P1::kThreads; ++t)
            reg_cols_stage_b<P1>(t, buf.data(), tws.data(), colq.data(), scratch.data(), M2, col0, twM);
    }
    std::vector<float2> bufA((size_t)P2::kBufElems), bufB((size_t)P2::kBufElems), tws2((size_t)P2::L);
    const int nthreads = 2 * P2::kThreads;
    for (int t = 0; t < nthreads; ++t) reg_fill_twiddles<P2>(t, nthreads, tws2.data(), p.s2.tw.data());
    UntangleOut u;
    u.X = X;
    u.M1 = M1;
    u.M2 = M2;
    u.twN = dev_view(twN_h);
    u.advance = advance;
    std::vector<float2> w2((size_t)M2), a2((size_t)M2);
    u.w2 = w2.data();
    u.a2 = a2.data();
    for (int t = 0; t < nthreads; ++t) reg_untangle_fill_tables<P2>(t, nthreads, w2.data(), a2.data(), u);
    const int half_groups = (M1 / 2 - 1 + GP - 1) / GP, groups = half_groups + 1;
    for (int c = 0; c < groups; ++c) {
        const int rA0 = 1 + (c - 1) * GP, rB0 = M1 - rA0 - GP + 1;
        for (int tt = 0; tt < nthreads; ++tt) {
            const int half = tt / P2::kThreads, t = tt - half * P2::kThreads;
            if (c == 0) {
                if (half == 0) reg_rows_stage_a_rows<P2>(t, scratch.data(), RowsSelfPaired{M1 / 2}, bufA.data());
            } else if (half == 0) {
                reg_rows_stage_a_rows<P2>(t, scratch.data(), RowsFrom{rA0}, bufA.data());
            } else {
                reg_rows_stage_a_rows<P2>(t, scratch.data(), RowsFrom{rB0}, bufB.data());
            }
        }
        std::vector<float2> regs((size_t)nthreads * P2::RB);
        for (int tt = 0; tt < nthreads; ++tt) {
            const int half = tt / P2::kThreads, t = tt - half * P2::kThreads;
            if (t < P2::kThreadsB && (c != 0 || half == 0))
                reg_rows_stage_b_regs<P2>(t, half ? bufB.data() : bufA.data(), tws2.data(), &regs[(size_t)tt * P2::RB]);
        }
        for (int tt = 0; tt < nthreads; ++tt) {
            const int half = tt / P2::kThreads, t = tt - half * P2::kThreads;
            if (t < P2::kThreadsB && (c != 0 || half == 0))
                reg_rows_store_local<P2>(t, &regs[(size_t)tt * P2::RB], half ? bufB.data() : bufA.data());
        }
        for (int tt = 0; tt < nthreads; ++tt) {
            if (c == 0) reg_untangle_special<P2>(tt, nthreads, bufA.data(), u);
            else reg_untangle_general<P2>(tt, nthreads, bufA.data(), bufB.data(), u, rA0);
        }
    }
}

static void check_fused_untangle(int N, int advance) {
    const int M = N / 2;
    std::vector<float> x((size_t)N);
    double norm = 0;
    for (int n = 0; n < N; ++n) {
        x[(size_t)n] = (float)(urand() * exp(-6.0 * n / N));
        norm += (double)x[(size_t)n] * x[(size_t)n];
    }
    norm = sqrt(norm);
    HostC2C h;
    if (!host_direct_plan(M, &h)) {
        ++failures;
        return;
    }
    HostTwoLevel twN;
    host_two_level(N, &twN);
    std::vector<float2> X((size_t)M + 1, make_float2(1e30f, 1e30f));
    bool used = false;
#define X_(ra1, rb1, g1, ra2, rb2, g2)                                                                   \
    if (!used && h.M1 == (ra1) * (rb1) && h.M2 == (ra2) * (rb2)) {                                       \
        constexpr int m1 = (ra1) * (rb1);                                                                \
        constexpr int gp = (m1 % 4 == 0) ? 4 : (g2);                                                     \
        if constexpr (m1 % 2 == 0 && m1 / 2 >= gp) {                                                     \
            emu_fused_forward<ra1, rb1, g1, ra2, rb2, gp>(h, x.data(), N, X.data(), twN, advance);       \
            used = true;                                                                                 \
        }                                                                                                \
    }
    AA_REG_LIST(X_)
#undef X_
    if (!used) {
        printf("fused untangle: no register-pass plan for N = %d\n", N);
        ++failures;
        return;
    }
    // every bin must have been written, and a sample of them must equal the direct sums (time-advanced)
    double worst = 0;
    for (int k = 0; k <= M; ++k)
        if (X[(size_t)k].x > 1e29f) {
            printf("fused untangle N=%d: bin %d not written\n", N, k);
            ++failures;
            return;
        }
    for (int c = 0; c < 40; ++c) {
        int k = (c == 0) ? 0 : (c == 1) ? M : (c == 2) ? M / 2 : (c == 3) ? h.M1 / 2 : (c == 4) ? h.M1 : (c == 5) ? M - h.M1 / 2
                : (c == 6) ? 1 : (c == 7) ? M - 1 : (int)(urand() * 0.5 * M + 0.5 * M) % (M + 1);
        if (k < 0) k = -k;
        cd acc = 0;
        for (int n = 0; n < N; ++n) {
            double ang = -2.0 * M_PI * (double)(((long long)(n - advance + (long long)N * 4) % N * k) % N) / N;
            acc += (double)x[(size_t)n] * cd(cos(ang), sin(ang));
        }
        double e = std::abs(acc - cd(X[(size_t)k].x, X[(size_t)k].y)) / norm;
        if (e > worst) worst = e;
    }
    char name[64];
    snprintf(name, sizeof(name), "rfft fused untangle, advance %d", advance);
    report(name, N, worst, 5e-6);
}

static void check_real_band(int N, double lo0 = 890.9, double hi0 = 1122.5) {
    const double sr = 48000.0;
    const double freq_scale = 1.0 / ((double)N * (1.0 / sr));
    const int half = N / 2;
    std::vector<float> x((size_t)N);
    double norm = 0;
    for (int n = 0; n < N; ++n) {
        x[(size_t)n] = (float)(urand() * exp(-6.0 * n / N));
        norm += (double)x[(size_t)n] * x[(size_t)n];
    }
    norm = sqrt(norm);
    std::vector<float2> X;  // half spectrum, k in [0, N/2]
    long long X_stride;
    HostTwoLevel twN;
    host_two_level(N, &twN);
    EmuPlan ep;
    const bool even = (N % 2) == 0;
    const int M = even ? N / 2 : N;
    if (!emu_plan(M, &ep)) {
        ++failures;
        return;
    }
    if (even) {
        std::vector<float2> Z((size_t)M);
        SrcDesc s{};
        s.mode = SRC_REAL_PACKED;
        s.base = x.data();
        s.unit_stride = N;
        s.n_valid = N;
        emu_exec(ep, s, dst_complex(Z.data(), M, M, 0, 1.0f), 1);
        X.resize((size_t)M + 1);
        TwoLevelTwiddle t = dev_view(twN);
        for (int k = 0; k <= M; ++k)
            X[(size_t)k] = untangle_half_spectrum(Z[(size_t)(k % M)], Z[(size_t)((M - k) % M)], twiddle2(t, k));
        X_stride = M + 1;
    } else {
        X.resize((size_t)N);
        SrcDesc s{};
        s.mode = SRC_REAL_AS_COMPLEX;
        s.base = x.data();
        s.unit_stride = N;
        s.n_valid = N;
        emu_exec(ep, s, dst_complex(X.data(), N, N, 0, 1.0f), 1);
        X_stride = N;
    }
    // forward check on a few bins
    double worst = 0;
    for (int c = 0; c < 24; ++c) {
        int k = (c == 0) ? 0 : (c == 1 ? half : (int)((urand() * 0.5 + 0.5) * half));
        cd acc = 0;
        for (int n = 0; n < N; ++n) {
            double ang = -2.0 * M_PI * (double)(((long long)n * k) % N) / N;
            acc += (double)x[(size_t)n] * cd(cos(ang), sin(ang));
        }
        double e = std::abs(acc - cd(X[(size_t)k].x, X[(size_t)k].y)) / norm;
        if (e > worst) worst = e;
    }
    char name[64];
    snprintf(name, sizeof(name), "rfft %s%s", even ? "packed" : "odd", ep.top.bluestein ? "+bluestein" : "");
    report(name, N, worst, ep.top.bluestein ? 3e-5 : 5e-6);

    // two bands in one batch (unit = row*n_bands + band)
    std::vector<BandDev> bands;
    bands.push_back(make_bandpass(lo0, hi0, 1.0 / 6.0, 24000.0, freq_scale, half + 1));
    bands.push_back(make_bandpass(55.7, 70.2, 1.0 / 6.0, 24000.0, freq_scale, half + 1));
    const int nb = (int)bands.size();
    std::vector<float> y((size_t)nb * N, 0.f);
    SrcDesc s{};
    s.mode = even ? SRC_BAND_PACKED : SRC_BAND_FULL;
    s.base = X.data();
    s.unit_stride = X_stride;
    s.n_valid = M;
    s.conj = 1;
    s.n_bands = nb;
    s.bands = bands.data();
    s.bands_have_tables = 1;  // the host accessor evaluates the mask when no table is attached
    s.freq_scale = freq_scale;
    s.N = N;
    s.M = even ? M : 0;
    s.twN = dev_view(twN);
    DstDesc d{};
    if (even) {
        d = dst_complex((float2*)y.data(), M, M, 1, 1.0f / (float)M);
    } else {
        d.mode = DST_REAL_PART;
        d.base = y.data();
        d.unit_stride = N;
        d.n_valid = N;
        d.conj1 = 1;
        d.scale1 = 1.0f / (float)N;
        d.scale2 = 1.0f;
    }
    std::vector<float> energy;
    int e_m1 = 0, e_m2 = 0, e_g = 0;
    if (even && !ep.top.bluestein && reg_split_for(M, &e_m1, &e_m2)) {  // tile energies next to the band signal
#define X(ra1, rb1, g1, ra2, rb2, g2) \
    if (e_m1 == (ra1) * (rb1) && e_m2 == (ra2) * (rb2)) e_g = (g2);
        AA_REG_LIST(X)
#undef X
        energy.assign((size_t)nb * (size_t)(e_m1 / e_g) * (size_t)e_m2, 0.f);
        d.energy = energy.data();
        d.energy_stride = (long long)(e_m1 / e_g) * e_m2;
    }
    emu_exec(ep, s, d, nb);
    if (!energy.empty()) {
        double worst_e = 0;
        for (int bi = 0; bi < nb; ++bi)
            for (int k2 = 0; k2 < e_m2; k2 += 7) {
                double got = 0, want = 0;
                for (int g = 0; g < e_m1 / e_g; ++g) got += energy[(size_t)bi * d.energy_stride + (size_t)g * e_m2 + k2];
                for (int i = 0; i < 2 * e_m1; ++i) {
                    const double v = y[(size_t)bi * N + (size_t)k2 * 2 * e_m1 + i];
                    want += v * v;
                }
                if (want > 0) worst_e = fmax(worst_e, fabs(got - want) / want);
            }
        report("  tile energy partials", N, worst_e, 2e-5);
    }
    // direct evaluation of irfft(X * mask) at a few samples, from the double-precision spectrum of x
    for (int bi = 0; bi < nb; ++bi) {
        BandDev open = bands[(size_t)bi];
        std::vector<cd> Y;
        std::vector<int> ks;
        double ynorm = 0;
        for (int k = open.k_first; k <= open.k_last; ++k) {
            float f = (float)((double)k * freq_scale);
            double m = band_mask_value(open, f);
            if (m == 0.0) continue;
            cd acc = 0;
            for (int n = 0; n < N; ++n) {
                double ang = -2.0 * M_PI * (double)(((long long)n * k) % N) / N;
                acc += (double)x[(size_t)n] * cd(cos(ang), sin(ang));
            }
            Y.push_back(acc * m);
            ks.push_back(k);
            ynorm += std::norm(acc * m);
        }
        ynorm = sqrt(2.0 * ynorm / N);  // rms-ish scale of the band signal times sqrt(N)
        double w2 = 0;
        for (int c = 0; c < 40; ++c) {
            int n = (int)((urand() * 0.5 + 0.5) * (N - 1));
            double acc = 0;
            for (size_t q = 0; q < ks.size(); ++q) {
                int k = ks[q];
                double ang = 2.0 * M_PI * (double)(((long long)n * k) % N) / N;
                double wgt = (k == 0 || (even && k == half)) ? 1.0 : 2.0;
                acc += wgt * (Y[q] * cd(cos(ang), sin(ang))).real();
            }
            acc /= N;
            double e = fabs(acc - (double)y[(size_t)bi * N + n]) / (ynorm / sqrt((double)N) + 1e-30);
            if (e > w2) w2 = e;
        }
        snprintf(name, sizeof(name), "  band %d inverse (%d bins)", bi, (int)ks.size());
        report(name, N, w2, ep.top.bluestein ? 2e-3 : 2e-4);
    }
}

// serial emulation of band_fused_kernel's transform (bandfused.cu, fft_fused.cuh): every thread of the CTA, stage by stage
template <class P>
static void emu_fused_unit(const RegSrcBand& src, int conj, float scale, float2* data) {
    std::vector<float2> tw((size_t)P::kTwTotal);
    for (int j = 2; j <= P::kStages; ++j)
        for (int k = 0; k < P::len(j - 1); ++k)
            tw[(size_t)(P::tw_offset(j) + k)] = make_float2((float)cos(2.0 * M_PI * k / P::len(j)), (float)-sin(2.0 * M_PI * k / P::len(j)));
    const FusedSupport sup = fused_support<P>(src);
    std::vector<FusedPrefetch> pre((size_t)P::kThreads);
    for (int t = 0; t < P::kThreads; ++t) fused_scatter_load<P>(t, src, sup, &pre[(size_t)t]);
    for (int t = 0; t < P::kThreads; ++t) fused_clear<P>(t, data);
    for (int t = 0; t < P::kThreads; ++t) fused_scatter<P>(t, src, sup, pre[(size_t)t], data);
    for (int t = 0; t < P::kThreads; ++t) fused_stage1<P>(t, data);
    for (int t = 0; t < P::kThreads; ++t) fused_stage<P, 2>(t, data, tw.data(), conj, scale);
    if constexpr (P::kStages >= 3)
        for (int t = 0; t < P::kThreads; ++t) fused_stage<P, 3>(t, data, tw.data(), conj, scale);
    if constexpr (P::kStages >= 4)
        for (int t = 0; t < P::kThreads; ++t) fused_stage<P, 4>(t, data, tw.data(), conj, scale);
}

static bool emu_fused(int M, const RegSrcBand& src, int conj, float scale, float2* data) {
#define X(m, th, mb, ...)                                                   \
    if (M == (m)) {                                                         \
        emu_fused_unit<FusedPlan<th, __VA_ARGS__>>(src, conj, scale, data); \
        return true;                                                        \
    }
    AA_FUSED_LIST(X)
#undef X
    return false;
}

// ---- 2b. decimated band synthesis (bands.cu): length N / D inverse of the time-advanced masked spectrum must
// return y[(shift + D j) mod N], every D-th sample of the full-rate band signal from the analysis start
static void check_decimated_band(int N, int D, double lo0, double hi0, int shift) {
    const double sr = 48000.0;
    const double freq_scale = 1.0 / ((double)N * (1.0 / sr));
    const int half = N / 2, Nd = N / D, Md = Nd / 2;
    BandDev band = make_bandpass(lo0, hi0, 1.0 / 6.0, 24000.0, freq_scale, half + 1);
    // wide bands get a sparse input (a few dozen impulses) so that the direct sums below stay cheap
    const bool sparse = band.k_last - band.k_first > 2000;
    std::vector<float> x((size_t)N, 0.f);
    std::vector<int> nz;
    for (int n = 0; n < N; ++n) {
        if (sparse && (n % 4801) != 7) continue;
        x[(size_t)n] = (float)(urand() * exp(-6.0 * n / N));
        nz.push_back(n);
    }
    {   // Nyquist zone of the decimated rate that holds the band (zone 0 = plain low-pass decimation)
        const int m = ((band.k_first + band.k_last) / 2) / Md;
        const int lo = (m == 0) ? 0 : m * Md + 1, hi = (m + 1) * Md - 2;
        if (band.k_first < lo || band.k_last > hi) {
            printf("decimated band check: band does not fit a zone of D=%d\n", D);
            ++failures;
            return;
        }
        band.z_neg = m & 1;
        band.z_off = (m & 1) ? (m + 1) * Md : m * Md;
    }
    // double-precision masked spectrum on the band's support
    std::vector<cd> Y;
    std::vector<int> ks;
    double ynorm = 0;
    std::vector<float2> X((size_t)half + 1, make_float2(0, 0));
    for (int k = band.k_first; k <= band.k_last; ++k) {
        cd acc = 0;
        for (int n : nz) {
            double ang = -2.0 * M_PI * (double)(((long long)n * k) % N) / N;
            acc += (double)x[(size_t)n] * cd(cos(ang), sin(ang));
        }
        X[(size_t)k] = make_float2((float)acc.real(), (float)acc.imag());
        const double m = band_mask_value(band, (float)((double)k * freq_scale));
        if (m == 0.0) continue;
        Y.push_back(acc * m);
        ks.push_back(k);
        ynorm += std::norm(acc * m);
    }
    ynorm = sqrt(2.0 * ynorm / N);
    EmuPlan ep;
    if (!emu_plan(Md, &ep) || ep.top.bluestein) {
        printf("no direct plan for %d\n", Md);
        ++failures;
        return;
    }
    HostTwoLevel twNd, twNfull;
    host_two_level(Nd, &twNd);
    host_two_level(N, &twNfull);
    std::vector<float> y((size_t)Nd, 0.f);
    SrcDesc s{};
    s.mode = SRC_BAND_PACKED;
    s.base = X.data();
    s.unit_stride = half + 1;
    s.n_valid = Md;
    s.conj = 1;
    s.n_bands = 1;
    s.bands = &band;
    s.bands_have_tables = 1;
    s.freq_scale = freq_scale;
    s.N = Nd;
    s.M = Md;
    s.twN = dev_view(twNd);
    s.shift = &shift;
    s.shift_N = N;
    s.twShift = dev_view(twNfull);
    DstDesc d = dst_complex((float2*)y.data(), Md, Md, 1, 1.0f / ((float)Md * (float)D));
    emu_exec(ep, s, d, 1);
    // the same unit through the in-shared-memory transform of band_fused_kernel, where a plan for this length exists
    std::vector<float> yf((size_t)Nd, 0.f);
    RegSrcBand rs;
    rs.X = X.data();
    rs.b = band;
    rs.freq_scale = freq_scale;
    rs.M = Md;
    rs.twN = dev_view(twNd);
    rs.shift = shift;
    rs.shift_N = N;
    rs.shift_rcp = 1.0 / (double)N;
    rs.twShift = dev_view(twNfull);
    const bool fused = emu_fused(Md, rs, 1, 1.0f / ((float)Md * (float)D), (float2*)yf.data());
    double worst = 0, worst_f = 0;
    for (int c = 0; c < 60; ++c) {
        const int j = (int)((urand() * 0.5 + 0.5) * (Nd - 1));
        const long long n = ((long long)shift + (long long)D * j) % N;
        double acc = 0;
        for (size_t q = 0; q < ks.size(); ++q) {
            double ang = 2.0 * M_PI * (double)((n * ks[q]) % N) / N;
            acc += 2.0 * (Y[q] * cd(cos(ang), sin(ang))).real();
        }
        acc /= N;
        worst = fmax(worst, fabs(acc - (double)y[(size_t)j]) / (ynorm / sqrt((double)N) + 1e-30));
        worst_f = fmax(worst_f, fabs(acc - (double)yf[(size_t)j]) / (ynorm / sqrt((double)N) + 1e-30));
    }
    char name[64];
    snprintf(name, sizeof(name), "decimated band D=%d zone=%d shift=%d (%d bins)", D, band.z_neg ? band.z_off / Md - 1 : band.z_off / Md, shift, (int)ks.size());
    report(name, N, worst, 2e-4);
    if (fused) {
        // and sample by sample against the two-pass result (same source arithmetic, different butterfly order)
        double diff = 0, ymax = 0;
        for (int j = 0; j < Nd; ++j) {
            diff = fmax(diff, fabs((double)yf[(size_t)j] - (double)y[(size_t)j]));
            ymax = fmax(ymax, fabs((double)y[(size_t)j]));
        }
        snprintf(name, sizeof(name), "  fused in-CTA transform M=%d vs direct sums", Md);
        report(name, N, worst_f, 2e-4);
        snprintf(name, sizeof(name), "  fused vs two-pass, all %d samples (of peak)", Nd);
        report(name, N, diff / (ymax + 1e-30), 2e-5);
    }
}

// the compile-time chains of fft.cu's AA_SPEC_LIST and csd.cu (radix 6, padded shared-memory layout)
template <int kPad>
static void check_chain(int L, std::vector<int> radices) {
    HostSubPlan hp;
    host_subplan(L, &hp);  // only its twiddle table is used
    SubPlanDev d{};
    d.L = L;
    d.nst = (int)radices.size();
    for (int s = 0; s < d.nst; ++s) d.radix[s] = radices[(size_t)s];
    d.tw = hp.tw.data();
    const size_t padded = (size_t)pad_index<kPad>(L - 1) + 1;
    std::vector<float2> x((size_t)L), a(padded, make_float2(0, 0)), b(padded, make_float2(0, 0));
    double norm = 0;
    for (int i = 0; i < L; ++i) {
        x[(size_t)i] = make_float2((float)urand(), (float)urand());
        a[(size_t)pad_index<kPad>(i)] = x[(size_t)i];
        norm += (double)x[(size_t)i].x * x[(size_t)i].x + (double)x[(size_t)i].y * x[(size_t)i].y;
    }
    norm = sqrt(norm);
    emu_chain<kPad>(d, a, b);
    double worst = 0;
    for (int k = 0; k < L; k += (L > 512 ? 17 : 1)) {
        cd acc = 0;
        for (int n = 0; n < L; ++n) {
            double ang = -2.0 * M_PI * (double)(((long long)n * k) % L) / L;
            acc += cd(x[(size_t)n].x, x[(size_t)n].y) * cd(cos(ang), sin(ang));
        }
        float2 got = a[(size_t)pad_index<kPad>(k)];
        worst = fmax(worst, std::abs(acc - cd(got.x, got.y)) / norm);
    }
    char name[64];
    snprintf(name, sizeof(name), "chain pad%d first radix %d (%d stages)", kPad, radices[0], d.nst);
    report(name, L, worst, 5e-6);
}

int main() {
    check_chain<0>(384, {8, 8, 6});
    check_chain<0>(375, {3, 5, 5, 5});
    check_chain<0>(320, {5, 8, 8});
    check_chain<0>(300, {3, 5, 5, 4});
    check_chain<4>(2048, {4, 8, 8, 8});
    check_chain<4>(512, {8, 8, 8});
    check_chain<4>(4096, {8, 8, 8, 8});
    check_chain<0>(36, {6, 6});
    const int sizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 11, 13, 16, 30, 64, 77, 143, 360, 1001, 2048,
                         2304, 4096, 24000, 30000, 31200, 65536, 17, 397, 4099, 25011, 45001, 144000, 96000, 72000, 36000, 48000, 24000, 18000, 9000, 12000, 6000, 16000, 8000, 4000};
    for (int M : sizes) check_c2c(M);
    const int real_sizes[] = {48000, 60000, 62400, 50022, 45001, 4410, 64, 1155};
    for (int N : real_sizes) check_real_band(N);
    check_real_band(288000, 111.4, 140.3);  // the register-pass splits (narrow bands keep the direct sums cheap)
    check_real_band(192000, 7127.2, 7300.0);
    check_fused_untangle(288000, 0);
    check_fused_untangle(288000, 17041);
    check_fused_untangle(192000, 9999);
    check_fused_untangle(192000, 1);
    check_decimated_band(288000, 4, 111.4, 140.3, 17041);
    check_decimated_band(288000, 2, 5610.0, 5700.0, 23456);
    check_decimated_band(192000, 4, 222.7, 280.6, 9999);
    check_decimated_band(192000, 2, 3500.0, 3600.0, 1);
    check_decimated_band(288000, 3, 10900.0, 11000.0, 777);    // zone 1 of 8 kHz (mirrored)
    check_decimated_band(288000, 4, 14400.0, 14500.0, 31);     // zone 2 of 6 kHz
    check_decimated_band(192000, 3, 6100.0, 6200.0, 5);        // zone 0, N / 3 = 64000
    check_decimated_band(288000, 6, 5610.0, 5700.0, 4242);     // zone 1 of 4 kHz
    check_decimated_band(288000, 12, 2820.0, 2900.0, 99);      // zone 1 of 2 kHz
    check_decimated_band(288000, 24, 222.7, 280.6, 1234);
    check_decimated_band(192000, 6, 2400.0, 2500.0, 8);
    check_decimated_band(192000, 12, 1150.0, 1200.0, 70000);
    check_decimated_band(192000, 24, 111.4, 140.3, 3);
    check_decimated_band(288000, 8, 111.4, 140.3, 20011);
    check_decimated_band(288000, 16, 157.5, 198.4, 13);
    check_decimated_band(288000, 16, 707.0, 891.0, 4321);      // support across the middle of the zone: the fused kernel's paired path
    check_decimated_band(192000, 8, 2227.0, 2300.0, 4321);
    check_decimated_band(192000, 16, 55.7, 70.2, 16001);
    printf("%s (%d failures)\n", failures ? "FAILED" : "ALL OK", failures);
    return failures ? 1 : 0;
}
