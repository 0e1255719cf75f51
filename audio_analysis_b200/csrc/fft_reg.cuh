// Two-stage register-resident FFT passes (the hot lengths of the rt60bands filterbank).
//
// A length L = RA * RB transform is done by RB butterflies of radix RA (stage A) followed by RA
// butterflies of radix RB (stage B), each butterfly living in one thread's registers
// (fft_radix.cuh).  Stage A reads its inputs straight from global memory, stage B writes its
// outputs straight to global memory: the data crosses shared memory exactly once per pass.
// A CTA works on G transforms side by side; the transform index g is the fastest thread index in
// both stages, so that global accesses of adjacent transforms (adjacent columns in pass 1,
// adjacent rows in pass 2) coalesce into whole sectors.
//
// Stockham index algebra (fft_core.cuh) for two stages:
//   stage A (Ns = 1):  butterfly ja in [0, RB) reads x[ja + r RB], writes d[ja RA + r]
//   stage B (Ns = RA): butterfly jb in [0, RA) reads d[jb + r RA] * w_L^{r jb}, writes X[jb + q RA]
// so stage B's input r of butterfly jb is stage A's output jb of butterfly r.
//
// The stage bodies are host/device functions of the thread index so that
// tests/csrc/fft_model_check.cu can run every thread of a CTA serially on the CPU.
#pragma once

#include "fft_access.cuh"
#include "fft_radix.cuh"

#ifndef AA_BAND_SRC_SINGLE_RA
#define AA_BAND_SRC_SINGLE_RA 5
#endif

namespace aa {

template <int RA_, int RB_, int G_, int TP_>
struct RegPass {
    static constexpr int RA = RA_, RB = RB_, L = RA_ * RB_, G = G_, TP = TP_;
    static constexpr int PA = RA_ | 1;  // odd pitch between stage-A butterflies: no bank conflicts
    static constexpr int kThreadsA = G_ * RB_, kThreadsB = G_ * RA_;
    static constexpr int kThreadsMax = kThreadsA > kThreadsB ? kThreadsA : kThreadsB;
    static constexpr int kThreads = (kThreadsMax + 31) / 32 * 32;
    static constexpr int kBufElems = G_ * TP_;
    static constexpr int kColQElems = G_ * RB_;  // pass 1 only: w_M^{RA col q} per (column, output block q)
    static constexpr size_t kSmemBytes = sizeof(float2) * (size_t)(kBufElems + L + kColQElems);
    static_assert(TP_ >= RB_ * PA, "transform pitch too small");
    AA_HD static int slot(int g, int ja, int ra) { return g * TP + ja * PA + ra; }
};

// inter-stage twiddle table in shared memory: tws[r * RA + jb] = w_L^{r jb}; full = exp(-2 pi i m / L)
template <class P>
AA_HD void reg_fill_twiddles(int t, int nthreads, float2* tws, const float2* full) {
    for (int m = t; m < P::L; m += nthreads) {
        const int r = m / P::RA, jb = m - r * P::RA;
        tws[m] = load_c(full + r * jb);
    }
}

// ---- element sources ---------------------------------------------------------------------------
// get(idx, &v) returns false when the element is structurally zero (v is then zero too).

struct RegSrcComplex {
    const float2* p;
    AA_HD bool get(int idx, float2* v) const {
        *v = load_c(p + idx);
        return true;
    }
    template <int RA>
    AA_HD bool get_all(int base, int stride, float2* v) const {
        static_for<0, RA>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            v[r] = load_c(p + base + r * stride);
        });
        return true;
    }
};

// Packed, band-masked, conjugated half spectrum of one (row, band) unit (rt60bands.py:170-175):
// the same arithmetic as fetch<SRC_BAND_PACKED> of the generic passes.
struct RegSrcBand {
    const float2* X;  // half spectrum of the row, k in [0, M]
    BandDev b;
    double freq_scale;
    int M;
    TwoLevelTwiddle twN;
    // optional circular time advance by `shift` samples of the length-shift_N signal the spectrum belongs to
    int shift = 0;
    int shift_N = 0;
    double shift_rcp = 0.0;     // 1 / shift_N
    TwoLevelTwiddle twShift{};  // exp(-2 pi i m / shift_N)
    AA_HD float2 advance(float2 v, int k) const {
#ifdef __CUDA_ARCH__
        return v;  // device callers store the spectrum already advanced (bands.cu: untangle_kernel); fft.cu rejects a shift
#else
        if (shift == 0) return v;
        // (k * shift) mod shift_N without a 64-bit division: the float64 quotient is exact to +-1
        const long long prod = (long long)k * (long long)shift;
        long long q = (long long)((double)prod * shift_rcp);
        long long r = prod - q * (long long)shift_N;
        if (r < 0) r += shift_N;
        if (r >= shift_N) r -= shift_N;
        return cmulc(v, twiddle2(twShift, r));   // * exp(+2 pi i k shift / N)
#endif
    }
    AA_HD float2 masked(int k) const {  // k inside [k_first, k_last]
#ifdef __CUDA_ARCH__
        // device: the per-call mask table is mandatory here (bands.cu builds it); evaluating the raised cosines
        // inline would put 2 x RA copies of cosf into the kernel and overflow the instruction cache
        const float m = load_r(b.mask + (k - b.k_first));
#else
        const float m = b.mask ? b.mask[k - b.k_first] : band_mask_value(b, (float)((double)k * freq_scale));
#endif
        return advance(cscale(load_c(X + k), m), k);
    }
    // decimated transforms read the row's spectrum through the band's Nyquist zone (BandDev::z_off / z_neg)
    AA_HD int zone_bin(int kp) const { return b.z_neg ? b.z_off - kp : b.z_off + kp; }
    AA_HD bool in_support(int k) const {
        // k_last < k_first (empty band) is rejected by the second term
        return (unsigned)(k - b.k_first) <= (unsigned)(b.k_last - b.k_first) && b.k_last >= b.k_first;
    }
    // Element idx of the packed, masked, conjugated spectrum, WITHOUT branches: out-of-support bins load the band's
    // first bin with a zero mask, so that the loads of a whole butterfly can be issued together.  (The per-element
    // branchy form serialised them: 49 % of the warp stall samples of this kernel were long-scoreboard waits.)
    AA_HD float2 element(int idx, bool in_k, bool in_m, float2 w /* exp(-2 pi i idx / N) */) const {
        const int mk = M - idx;
        const int ka = in_k ? zone_bin(idx) : b.k_first, kb = in_m ? zone_bin(mk) : b.k_first;
#ifdef __CUDA_ARCH__
        float ma = load_r(b.mask + (ka - b.k_first)), mb = load_r(b.mask + (kb - b.k_first));
#else
        float ma = b.mask ? b.mask[ka - b.k_first] : band_mask_value(b, (float)((double)ka * freq_scale));
        float mb = b.mask ? b.mask[kb - b.k_first] : band_mask_value(b, (float)((double)kb * freq_scale));
#endif
        const float2 xa = load_c(X + ka), xb = load_c(X + kb);
        ma = in_k ? ma : 0.0f;
        mb = in_m ? mb : 0.0f;
        float2 ya = advance(cscale(xa, ma), ka), yb = advance(cscale(xb, mb), kb);
        if (b.z_neg) {
            ya.y = -ya.y;
            yb.y = -yb.y;
        }
        return cconj(pack_half_spectrum(ya, yb, w));
    }
    AA_HD bool get(int idx, float2* v) const {
        const bool in_k = in_support(zone_bin(idx)), in_m = in_support(zone_bin(M - idx));
        if (!in_k && !in_m) {
            *v = make_float2(0.f, 0.f);
            return false;
        }
        *v = element(idx, in_k, in_m, twiddle2(twN, idx));
        return true;
    }
    // the RA inputs base + r * stride of one stage-A butterfly; false = all structurally zero
    template <int RA>
    AA_HD bool get_all(int base, int stride, float2* v) const {
        unsigned bits_k = 0, bits_m = 0;
        static_for<0, RA>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            const int idx = base + r * stride;
            bits_k |= in_support(zone_bin(idx)) ? (1u << r) : 0u;
            bits_m |= in_support(zone_bin(M - idx)) ? (1u << r) : 0u;
        });
        if ((bits_k | bits_m) == 0u) {
            static_for<0, RA>([&](auto r_) { v[decltype(r_)::value] = make_float2(0.f, 0.f); });
            return false;
        }
        // the butterfly's inputs sit N / (2 RA) apart (stride = RB M2, N = 2 RA RB M2), so their packing twiddles are
        // w(base) times compile-time roots of unity: one table lookup per butterfly instead of RA.  Inputs are built
        // in groups of four: a group with no bin in the band's support costs one test (narrow bands touch one or two
        // groups), inside a group everything is branch-free so that its loads are in flight together.
        const float2 w0 = twiddle2(twN, base);
        const unsigned bits = bits_k | bits_m;
        // group length: 4 for the long butterflies (wide bands: most inputs are in the support, batching the loads
        // pays), 1 for radix <= AA_BAND_SRC_SINGLE_RA (the deep-decimation classes hold narrow bands: one or two
        // inputs of a butterfly are in the support, and a group of four would build three zeros around each)
        constexpr int kGroup = (RA <= AA_BAND_SRC_SINGLE_RA) ? 1 : 4;
        static_for<0, (RA + kGroup - 1) / kGroup>([&](auto c_) {
            constexpr int lo = kGroup * decltype(c_)::value, hi = (lo + kGroup < RA) ? lo + kGroup : RA;
            if ((bits >> lo) & ((1u << (hi - lo)) - 1u)) {
                static_for<lo, hi>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    v[r] = element(base + r * stride, (bits_k >> r) & 1u, (bits_m >> r) & 1u, mul_root<r, 2 * RA>(w0));
                });
            } else {
                static_for<lo, hi>([&](auto r_) { v[decltype(r_)::value] = make_float2(0.f, 0.f); });
            }
        });
        return true;
    }
};

// ---- pass 1: columns.  CTA = G adjacent columns, transform over n1 (stride M2), output [k1 * M2 + col] -------

template <class P, class Src>
AA_HD void reg_cols_stage_a(int t, const Src& src, int M2, int col0, float2* buf) {
    if (t >= P::kThreadsA) return;
    const int ja = t / P::G, g = t - ja * P::G;
    float2 v[P::RA];
    const bool any = src.template get_all<P::RA>(ja * M2 + col0 + g, P::RB * M2, v);
    if (any) dft_r<P::RA>(v);
    static_for<0, P::RA>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        buf[P::slot(g, ja, r)] = v[r];
    });
}

// four-step twiddle w_M^{col k1}, k1 = jb + q RA, factored as w_M^{col jb} * w_M^{RA col q}: the second factor is
// shared by every butterfly jb of the column and lives in shared memory, colq[q * G + g]
template <class P>
AA_HD void reg_cols_fill_colq(int t, float2* colq, int col0, const TwoLevelTwiddle& twM) {
    if (t >= P::kColQElems) return;
    const int q = t / P::G, g = t - q * P::G;
    colq[t] = twiddle2(twM, (long long)(P::RA * (col0 + g)) * q);
}

template <class P>
AA_HD void reg_cols_stage_b(int t, const float2* buf, const float2* tws, const float2* colq, float2* scratch_unit,
                            int M2, int col0, const TwoLevelTwiddle& twM) {
    if (t >= P::kThreadsB) return;
    const int jb = t / P::G, g = t - jb * P::G;
    float2 v[P::RB];
    static_for<0, P::RB>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        const float2 x = buf[P::slot(g, r, jb)];
        v[r] = (r == 0) ? x : cmul(x, tws[r * P::RA + jb]);
    });
    dft_r<P::RB>(v);
    const int col = col0 + g;
    const float2 c = twiddle2(twM, (long long)(col * jb));
    static_for<0, P::RB>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int k1 = jb + q * P::RA;
        scratch_unit[k1 * M2 + col] = cmul(v[q], cmul(c, colq[q * P::G + g]));
    });
}

// ---- pass 2: rows.  CTA = G adjacent rows k1, transform over n2 (contiguous), output [k2 * M1 + row] ----------

template <class P, bool kCg = false>
AA_HD void reg_rows_stage_a(int t, const float2* scratch_unit, int row0, float2* buf) {
    if (t >= P::kThreadsA) return;
    const int ja = t / P::G, g = t - ja * P::G;
    const float2* in = scratch_unit + (row0 + g) * P::L + ja;
    float2 v[P::RA];
    static_for<0, P::RA>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        v[r] = kCg ? load_c_cg(in + r * P::RB) : load_c(in + r * P::RB);
    });
    dft_r<P::RA>(v);
    static_for<0, P::RA>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        buf[P::slot(g, ja, r)] = v[r];
    });
}

// energy != nullptr: also returns |output|^2 of the thread's RB outputs (the band signal's two real samples per
// complex output), for the per-tile energy partials that feed the band EDC (bands.cu: band_fit_kernel)
template <class P>
AA_HD void reg_rows_stage_b(int t, const float2* buf, const float2* tws, float2* out_unit, int M1, int row0,
                            int n_valid, int conj, float scale, float* energy = nullptr) {
    if (t >= P::kThreadsB) return;
    const int jb = t / P::G, g = t - jb * P::G;
    float2 v[P::RB];
    static_for<0, P::RB>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        const float2 x = buf[P::slot(g, r, jb)];
        v[r] = (r == 0) ? x : cmul(x, tws[r * P::RA + jb]);
    });
    dft_r<P::RB>(v);
    const int row = row0 + g;
    static_for<0, P::RB>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        const int idx = (jb + q * P::RA) * M1 + row;
        float2 o = v[q];
        if (conj) o.y = -o.y;
        o = cscale(o, scale);
        if (energy) energy[q] = (idx < n_valid) ? o.x * o.x + o.y * o.y : 0.0f;
        if (idx < n_valid) out_unit[idx] = o;
    });
}

// ---- pass 2 of the forward REAL transform fused with the untangle step ---------------------------------------------
// The packed transform Z = FFT_M(x[2m] + i x[2m+1]) becomes the half spectrum through pairs:
//   X[k] = untangle(Z[k], Z[M - k], w^k),  X[M - k] = untangle(Z[M - k], Z[k], -conj w^k),  w = exp(-2 pi i / N).
// With k = k2 M1 + k1 the partner M - k is (M2 - 1 - k2) M1 + (M1 - k1) for k1 > 0: row k1 of the rows pass pairs with
// row M1 - k1.  A CTA therefore transforms a group A of G rows k1 = rA0 .. rA0 + G - 1 (1 <= k1 < M1 / 2) AND the
// mirrored group B = rows M1 - k1, keeps both groups' outputs in shared memory and writes X directly - the packed
// spectrum never goes to memory and the separate untangle pass (a read and a write of the whole spectrum) disappears.
// The self-paired rows 0 and M1 / 2 (M1 even) are handled by one extra CTA per transform.
// Optionally the spectrum is stored time-advanced by s samples (X[k] e^{+2 pi i k s / N}, bands.cu).

struct RowsFrom {      // slot g -> row r0 + g
    int r0;
    AA_HD int operator()(int g) const { return r0 + g; }
};
struct RowsSelfPaired {   // slot 0 -> row 0, slot 1 -> row M1 / 2 (the other slots repeat row 0 and are ignored)
    int half;
    AA_HD int operator()(int g) const { return g == 1 ? half : 0; }
};

// stage A with an arbitrary row per slot g (row_of(g)); same shared-memory layout as reg_rows_stage_a
template <class P, class RowOf>
AA_HD void reg_rows_stage_a_rows(int t, const float2* scratch_unit, RowOf row_of, float2* buf) {
    if (t >= P::kThreadsA) return;
    const int ja = t / P::G, g = t - ja * P::G;
    const float2* in = scratch_unit + (long long)row_of(g) * P::L + ja;
    float2 v[P::RA];
    static_for<0, P::RA>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        v[r] = load_c(in + r * P::RB);
    });
    dft_r<P::RA>(v);
    static_for<0, P::RA>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        buf[P::slot(g, ja, r)] = v[r];
    });
}

// stage B into registers: v[q] = output k2 = jb + q RA of the row in slot g (t = jb G + g)
template <class P>
AA_HD void reg_rows_stage_b_regs(int t, const float2* buf, const float2* tws, float2* v) {
    const int jb = t / P::G, g = t - jb * P::G;
    static_for<0, P::RB>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        const float2 x = buf[P::slot(g, r, jb)];
        v[r] = (r == 0) ? x : cmul(x, tws[r * P::RA + jb]);
    });
    dft_r<P::RB>(v);
}

// ... and from registers into the group's local output table out[k2 * G + g] (may alias the stage buffer once every
// thread of the group has finished reg_rows_stage_b_regs)
template <class P>
AA_HD void reg_rows_store_local(int t, const float2* v, float2* out) {
    const int jb = t / P::G, g = t - jb * P::G;
    static_for<0, P::RB>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        out[(jb + q * P::RA) * P::G + g] = v[q];
    });
}

struct UntangleOut {
    float2* X;              // half spectrum of this transform, M + 1 entries
    int M1, M2;             // M = M1 M2, N = 2 M
    TwoLevelTwiddle twN;    // exp(-2 pi i m / N)
    int advance;            // s: store X[k] exp(+2 pi i k s / N); 0 = plain
    // per-CTA tables (shared memory), filled by reg_untangle_fill_tables: the twiddles of bin k = k2 M1 + k1 factor
    // into a k2 part (these tables, read with consecutive k2: coalesced) and a k1 part (one value per thread);
    // looking w^k up per element scatters 32 sectors per warp request over the twiddle tables
    const float2* w2;       // [M2] exp(-2 pi i (k2 M1) / N)
    const float2* a2;       // [M2] exp(-2 pi i ((k2 M1 s) mod N) / N)
};

template <class P>
AA_HD void reg_untangle_fill_tables(int t, int nthreads, float2* w2, float2* a2, const UntangleOut& u) {
    const unsigned N = 2u * (unsigned)u.M1 * (unsigned)u.M2;
    const unsigned s = (unsigned)u.advance;
    const unsigned a_step = (unsigned)(((unsigned long long)u.M1 * s) % N);   // (M1 s) mod N
    for (int k2 = t; k2 < P::L; k2 += nthreads) {
        w2[k2] = twiddle2(u.twN, (long long)k2 * u.M1);
        a2[k2] = s ? twiddle2(u.twN, (long long)(((unsigned long long)k2 * a_step) % N)) : make_float2(1.f, 0.f);
    }
}

// X[k], X[mk] from the pair (zk, zm); w = exp(-2 pi i k / N), adv = exp(-2 pi i ((k s) mod N) / N)
AA_HD void untangle_store_pair(const UntangleOut& u, int k, int mk, float2 zk, float2 zm, float2 w, float2 adv) {
    float2 a = untangle_half_spectrum(zk, zm, w);
    float2 b = untangle_half_spectrum(zm, zk, make_float2(-w.x, w.y));
    if (u.advance != 0) {
        // exp(+2 pi i k s / N) = conj(adv);  exp(+2 pi i (M - k) s / N) = (-1)^s adv
        a = cmulc(a, adv);
        b = cmul(b, adv);
        if (u.advance & 1) b = make_float2(-b.x, -b.y);
    }
    u.X[k] = a;
    u.X[mk] = b;
}

// general groups: element (k2, g) of group A (row k1 = rA0 + g) with its partner in group B, whose slot G - 1 - g
// holds row M1 - k1.  nthreads must be a multiple of G: a thread keeps its slot g, hence its row, in every iteration.
template <class P>
AA_HD void reg_untangle_general(int t, int nthreads, const float2* outA, const float2* outB, const UntangleOut& u, int rA0) {
    const unsigned N = 2u * (unsigned)u.M1 * (unsigned)u.M2;
    const unsigned s = (unsigned)u.advance;
    const int g = t % P::G;
    const int k1 = rA0 + g;
    if (2 * k1 >= u.M1) return;                            // beyond the last mirrored pair
    const int dk2 = nthreads / P::G;
    const float2 w1 = twiddle2(u.twN, k1);
    const float2 b1 = s ? twiddle2(u.twN, (long long)(((unsigned long long)k1 * s) % N)) : make_float2(1.f, 0.f);
    for (int k2 = t / P::G; k2 < P::L; k2 += dk2) {
        const int k = k2 * u.M1 + k1;
        const int mk = (u.M2 - 1 - k2) * u.M1 + (u.M1 - k1);
        untangle_store_pair(u, k, mk, outA[k2 * P::G + g], outB[(u.M2 - 1 - k2) * P::G + (P::G - 1 - g)],
                            cmul(u.w2[k2], w1), cmul(u.a2[k2], b1));
    }
}

// the self-paired rows: slot 0 holds row 0, slot 1 row M1 / 2
template <class P>
AA_HD void reg_untangle_special(int t, int nthreads, const float2* out, const UntangleOut& u) {
    const unsigned N = 2u * (unsigned)u.M1 * (unsigned)u.M2;
    const unsigned s = (unsigned)u.advance;
    const int M = u.M1 * u.M2;
    const int kh = u.M1 / 2;
    const float2 w1h = twiddle2(u.twN, kh);
    const float2 b1h = s ? twiddle2(u.twN, (long long)(((unsigned long long)kh * s) % N)) : make_float2(1.f, 0.f);
    for (int e = t; e < 2 * P::L; e += nthreads) {
        const int which = e / P::L, k2 = e - which * P::L;
        if (which == 0) {   // row 0: k = k2 M1 pairs with (M2 - k2) M1; k2 = 0 pairs bin 0 with bin M (both from Z[0])
            const int p2 = (k2 == 0) ? 0 : u.M2 - k2;
            if (k2 != 0 && 2 * k2 > u.M2) continue;
            const int k = k2 * u.M1;
            const int mk = (k2 == 0) ? M : p2 * u.M1;
            untangle_store_pair(u, k, mk, out[k2 * P::G + 0], out[p2 * P::G + 0], u.w2[k2], u.a2[k2]);
        } else {            // row M1 / 2: k2 pairs with M2 - 1 - k2
            const int p2 = u.M2 - 1 - k2;
            if (k2 > p2) continue;
            const int k = k2 * u.M1 + kh, mk = p2 * u.M1 + kh;
            untangle_store_pair(u, k, mk, out[k2 * P::G + 1], out[p2 * P::G + 1], cmul(u.w2[k2], w1h), cmul(u.a2[k2], b1h));
        }
    }
}

}  // namespace aa
