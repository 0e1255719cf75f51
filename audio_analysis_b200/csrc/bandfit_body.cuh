// Band EDC + crossings + decay fits of ONE (row, band) unit from per-tile energies: the body shared by
// band_fit_kernel (bandfit.cu: band signal and tile partials in global memory, written by the rows pass of the
// masked inverse) and band_fused_kernel (bandfused.cu: band signal synthesised into shared memory by the same CTA).
// Semantics: analyse/decay.py:150-158 (Schroeder integral), :173-199 (crossings), :224-249 (fits), as edc_body.cuh.
#pragma once

#include "edc_body.cuh"
#include "kernels.cuh"

namespace aa {
namespace bandfit {

using edc::EdcArgs;
using edc::kMaxTargets;
using edc::TileScan;
using edc::tile_scan;
using edc::ScanMasks;
using edc::make_scan_masks;
using edc::to_db;

constexpr int kSub = 128;     // samples per warp step: one float4 per lane
constexpr int kMaxSub = 8;    // tile length <= 1024 samples

template <bool kSmem>
__device__ __forceinline__ float4 fit_ld4(const float* p) {
    if constexpr (kSmem) return *reinterpret_cast<const float4*>(p);
    else return ldg_f4(p);
}
template <bool kSmem>
__device__ __forceinline__ float fit_ld1(const float* p) {
    if constexpr (kSmem) return *p;
    else return __ldg(p);
}

struct BandFitArgs {
    const float* y;
    long long y_stride;
    const float* part;        // [unit][n_groups][n_tiles]
    long long part_stride;
    int n_groups, n_tiles, tile_len, n;
    // decimated band signals (bands.cu): y holds every decim-th sample of the band signal, starting AT the analysis
    // start (the spectrum was time-advanced), so the curve starts at index 0, ends after ceil((full_n - start) /
    // decim) samples (the rest is the wrapped-around head of the circular signal), and the tail sum of squares is
    // corrected by the end-point term of the trapezoid rule: E_j = decim * (S_j - kappa y_j^2), kappa = (decim-1)/(2 decim)
    int decim, full_n, aligned;
    float kappa;
    double inv_decim;         // 1 / decim
    int min_fit_points;
    // addressing: dense (unit = row * nbc + j) or through a unit list (unit_map[unit] = row * nbc + j); the result of
    // class band j of a row goes to rt60_out[(row * out_nb + out_band[j]) * AA_MAX_FITS + f]
    const int* unit_map;
    int nbc, out_nb;
    // certification of a decimated unit (launch_band_fit's BandFitCert): see certify_unit below
    BandFitCert cert;
    double gain10[kMaxTargets];   // 10^(-target / 10) per dB target (certification: one far-end cell in dB at that level)
    EdcArgs e;                // start_in / rt60_out / targets / fit ranges / eps / floor / sr / min_samples
};

// append (row, band) to the unit list of the level that will redo it: entries are row * nbc' + j' in the addressing
// of the class that owns the band on that level
__device__ __forceinline__ void flag_unit(const BandFitCert& c, int row, int cj) {
    const int l = c.next_list[cj];
    const int pos = atomicAdd(c.counts + l, 1);
    c.lists[c.next_off[cj] + pos] = row * (int)c.next_nbc[cj] + (int)c.next_j[cj];
}

// energy of sample c (0..3) of the lane's four from the in-lane suffix sums
__device__ __forceinline__ float lane_sample_energy(const TileScan& sc, int c) {
    return (c == 3) ? sc.s[3] : (c == 2) ? __fsub_rn(sc.s[2], sc.s[3]) : (c == 1) ? __fsub_rn(sc.s[1], sc.s[2])
                                                                                   : __fsub_rn(sc.s[0], sc.s[1]);
}

template <int kW>
struct FitShared {
    int cert_idx[AA_MAX_FITS * 4];
    float fit_wa[AA_MAX_FITS], fit_wb[AA_MAX_FITS];   // weights of the first / last sample of a fit (decimated grids)
    double fit_slope_u[AA_MAX_FITS], fit_b0[AA_MAX_FITS];
    int fit_n_full[AA_MAX_FITS];
    double warp_tot[kW];
    int cross_ok[kMaxTargets], cross_idx[kMaxTargets];
    double cross_t[kMaxTargets];
    int fit_valid[AA_MAX_FITS], fit_first[AA_MAX_FITS], fit_last[AA_MAX_FITS];
    double fit_ts[AA_MAX_FITS], fit_te[AA_MAX_FITS];
    double warp_acc[kW][AA_MAX_FITS * 3];
};

}  // namespace bandfit

// bandfit.cu: fills the argument block shared by band_fit_kernel and band_fused_kernel (everything but the signal /
// partial pointers and the tile layout)
int fill_band_fit_args(bandfit::BandFitArgs* out, int n, const aa_decay_params* p, const int* d_start_idx, int nbc, int out_nb,
                       const int* d_unit_map, const BandFitCert* cert, int min_samples, double* d_rt60_out, int decim,
                       int full_n, int aligned);

namespace bandfit {

// One tile at sample resolution, visited by a warp from its last 128-sample step to its first:
// f(j, scan, after) gets the step's suffix scan and the float64 energy that follows the step
// (later steps of the tile + later tiles).
template <int NS, bool kSmem, class F>
__device__ __forceinline__ void visit_tile_reverse(const float* __restrict__ y, int k, int tile_len, int n, double carry_k,
                                                   int lane, const ScanMasks& masks, F&& f) {
    float4 v[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int off = j * kSub + 4 * lane;
        const int i0 = k * tile_len + off;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (off < tile_len) {
            if (i0 + 3 < n) v[j] = fit_ld4<kSmem>(y + i0);
            else {
                if (i0 < n) v[j].x = fit_ld1<kSmem>(y + i0);
                if (i0 + 1 < n) v[j].y = fit_ld1<kSmem>(y + i0 + 1);
                if (i0 + 2 < n) v[j].z = fit_ld1<kSmem>(y + i0 + 2);
            }
        }
    }
    double after = carry_k;
#pragma unroll
    for (int j = NS - 1; j >= 0; --j) {
        const TileScan sc = tile_scan(v[j], masks);
        f(j, sc, after);
        after += (double)__shfl_sync(kFullMask, __fadd_rn(sc.excl, sc.s[0]), 0);
    }
}

// One unit by a CTA of kT threads.  y: the unit's band signal (global, or shared when kSmem), part: its tile energy
// partials (global; unused when kSmem), dyn: NT * 12 bytes of shared memory for the carries, sh: the CTA's scratch.
template <int NS, int kT, bool kSmem>
__device__ __forceinline__ void band_fit_unit(const BandFitArgs& a, int row, int cj, const float* __restrict__ y,
                                              const float* __restrict__ part, unsigned char* dyn_smem,
                                              FitShared<kT / 32>& sh) {
    constexpr int kW = kT / 32;
    double* __restrict__ rt60_out = a.e.rt60_out + ((long long)row * a.out_nb + a.cert.out_band[cj]) * AA_MAX_FITS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ScanMasks masks = make_scan_masks();
    const int NT = a.n_tiles, TL = a.tile_len;
    double* carry = reinterpret_cast<double*>(dyn_smem);   // [NT] energy after tile k
    float* tsum = reinterpret_cast<float*>(carry + NT);    // [NT] tile energy
    const EdcArgs& p = a.e;
    int n = a.n, start = min(max(p.start_in[row], 0), a.aligned ? a.full_n : a.n);
    const int full_len = a.full_n - start;   // analysed samples at full rate (aligned grids)
    if (a.aligned) {   // the (decimated) grid starts at the analysis start and stops before the signal wraps around
        n = min(a.n, (a.full_n - start + a.decim - 1) / a.decim);
        start = 0;
    }
    const double kappa = (double)a.kappa;
    const int n_an = n - start;
    const float srf = (float)p.sr;

    // rt60bands.py:395-397: fewer than 8 samples -> every metric None (aligned grids test the full-rate count)
    if (a.aligned ? (full_len < p.min_samples) : (n_an < p.min_samples)) {
        if (threadIdx.x < AA_MAX_FITS) rt60_out[threadIdx.x] = nan("");
        return;
    }

    // ---- tile energies and their float64 exclusive suffix scan
    const int per = (NT + kT - 1) / kT;
    const int t0 = threadIdx.x * per;
    double mine = 0.0;
    if constexpr (kSmem) {
        // the band signal sits in shared memory (bandfused.cu): one THREAD per tile of 128 samples sums its squares (no
        // shuffles); thread k starts at 16-byte word k of its tile so that the 32 lanes of a request spread over all banks
        static_assert(!kSmem || NS == 1, "shared-memory band signals use tiles of one warp step");
        for (int k = threadIdx.x; k < NT; k += kT) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const int cnt = min(TL, n - k * TL);   // analysed samples of the tile (<= 0: none)
            const float4* __restrict__ p = reinterpret_cast<const float4*>(y + (long long)k * TL);
            if (cnt == kSub) {
#pragma unroll 8
                for (int j = 0; j < kSub / 4; ++j) {
                    const float4 v = p[(j + k) & (kSub / 4 - 1)];
                    s0 = fmaf(v.x, v.x, s0);
                    s1 = fmaf(v.y, v.y, s1);
                    s2 = fmaf(v.z, v.z, s2);
                    s3 = fmaf(v.w, v.w, s3);
                }
            } else {
                for (int q = 0; 4 * q < cnt; ++q) {
                    const float4 v = p[q];
                    s0 = fmaf(v.x, v.x, s0);
                    if (4 * q + 1 < cnt) s1 = fmaf(v.y, v.y, s1);
                    if (4 * q + 2 < cnt) s2 = fmaf(v.z, v.z, s2);
                    if (4 * q + 3 < cnt) s3 = fmaf(v.w, v.w, s3);
                }
            }
            tsum[k] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
    } else {
        for (int q = 0; q < per; ++q) {
            const int k = t0 + q;
            if (k < NT) {
                float s = 0.f;
                if ((long long)k * TL < n)
                    for (int g = 0; g < a.n_groups; ++g) s += __ldg(part + (long long)g * NT + k);
                tsum[k] = s;
            }
        }
        __syncthreads();
        const int k_end = (n > 0) ? (n - 1) / TL : 0;   // tile that holds the last analysed sample
        if (warp == 0 && n % TL != 0 && k_end < NT) {   // partly analysed tile: its energy up to sample n only
            float tot = 0.f;
            visit_tile_reverse<NS, kSmem>(y, k_end, TL, n, 0.0, lane, masks, [&](int j, const TileScan& sc, double after) {
                tot += __shfl_sync(kFullMask, __fadd_rn(sc.excl, sc.s[0]), 0);
            });
            if (lane == 0) tsum[k_end] = tot;
        }
        __syncthreads();
    }
    for (int q = 0; q < per; ++q) {
        const int k = t0 + q;
        if (k < NT) mine += (double)tsum[k];
    }
    double inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_down_sync(kFullMask, inc, d);
        if (lane + d < 32) inc += o;
    }
    if (lane == 0) sh.warp_tot[warp] = inc;
    __syncthreads();
    {
        double higher = 0.0;
        for (int w = warp + 1; w < kW; ++w) higher += sh.warp_tot[w];
        double after = __shfl_down_sync(kFullMask, inc, 1);
        if (lane == 31) after = 0.0;
        // far end of a decimated grid: the last sample J stands for the r = full_len - decim J samples that are left,
        // and the trapezoid rule gives it half a step: S += (r - (decim + 1) / 2) / decim * y_J^2
        double s_end = 0.0;
        if (a.decim > 1 && n > 0) {
            const float yl = fit_ld1<kSmem>(y + (n - 1));
            const double r = (double)(full_len - a.decim * (n - 1));
            s_end = (r - 0.5 * (double)(a.decim + 1)) * a.inv_decim * (double)__fmul_rn(yl, yl);
        }
        double running = higher + after + s_end;
        for (int q = per - 1; q >= 0; --q) {
            const int k = t0 + q;
            if (k < NT) {
                carry[k] = running;
                running += (double)tsum[k];
            }
        }
    }
    __syncthreads();

    // ---- E0 = E[start]: every warp works it out for itself (one tile at sample resolution, identical arithmetic and
    // therefore identical values), which saves the hand-over through shared memory and its barrier
    const int ks = min(start / TL, NT - 1);
    double e0 = 0.0;
    {
        const int pofs = start - ks * TL;
        const int pj = pofs / kSub, pl = (pofs % kSub) >> 2, pc = pofs & 3;
        visit_tile_reverse<NS, kSmem>(y, ks, TL, n, carry[ks], lane, masks, [&](int j, const TileScan& sc, double after) {
            if (j == pj && lane == pl) {
                const float sj = (pc == 0) ? sc.s[0] : (pc == 1) ? sc.s[1] : (pc == 2) ? sc.s[2] : sc.s[3];
                e0 = after + (double)__fadd_rn(sc.excl, sj) - kappa * (double)lane_sample_energy(sc, pc);
            }
        });
        e0 = __shfl_sync(kFullMask, e0, pl);
    }
    const double inv_e0 = 1.0 / fmax(e0, p.eps);
    const int n_targets = 2 + 2 * p.n_fits;

    // ---- crossings (decay.py:173-199): first sample <= target, by one warp (all lanes return the same values)
    // hint_tile >= 0: the tile where the crossing is expected (the certification targets sit within a few samples of a
    // crossing that is already known): one 32-tile window is tried before the search over all tiles
    auto resolve = [&](float tauf, int& found, int& idx, float& y0, float& y1, int hint_tile = -1) {
        found = 0;
        idx = 0;
        y0 = y1 = 0.f;
        if (tauf >= 0.0f) {
            found = 1;  // the curve starts at exactly 0 dB (decay.py:186-187)
            idx = start;
            return;
        }
        int hit = -1;
        // the tile with dB(first sample) > target >= dB(first sample of the next tile) is unique (monotone curve)
        auto probe = [&](int kb) {
            const int k = kb + lane;
            bool ok = false;
            if (k < NT) {
                const float b_next = to_db(carry[k], p.eps, inv_e0, p.floor_db);
                float b_first = 0.0f;
                if (k > ks) b_first = to_db(carry[k] + (double)tsum[k], p.eps, inv_e0, p.floor_db);
                ok = (b_first > tauf) && (b_next <= tauf);
            }
            const unsigned m = __ballot_sync(kFullMask, ok);
            if (m) hit = kb + (__ffs(m) - 1);
        };
        if (hint_tile >= 0) probe(max(ks, hint_tile - 2));
        for (int kb = ks; kb < NT && hit < 0; kb += 32) probe(kb);
        if (hit < 0) return;
        const int k = hit;
        // steps are visited last to first; the curve is monotone, so the last step that reports a
        // sample <= target holds the first such sample
        visit_tile_reverse<NS, kSmem>(y, k, TL, n, carry[k], lane, masks, [&](int j, const TileScan& sc, double after) {
            int first_c = 4;
            float dbv = 0.f;
            double ev = 0.0;
#pragma unroll
            for (int c = 3; c >= 0; --c) {
                const int off = j * kSub + 4 * lane + c;
                const int i = k * TL + off;
                const double eu = after + (double)__fadd_rn(sc.excl, sc.s[c]);
                const float d = to_db(eu - kappa * (double)lane_sample_energy(sc, c), p.eps, inv_e0, p.floor_db);
                if (off < TL && i >= start && i < n && d <= tauf) {
                    first_c = c;
                    dbv = d;
                    ev = eu;
                }
            }
            const unsigned m = __ballot_sync(kFullMask, first_c < 4);
            if (m) {
                const int fl = __ffs(m) - 1;
                float py0 = 0.f;
                int pidx = 0;
                if (lane == fl) {
                    pidx = k * TL + j * kSub + 4 * lane + first_c;
                    // value one sample earlier: E[i-1] = E[i] + y[i-1]^2 (i > start because b_first > tau)
                    const float yp = fit_ld1<kSmem>(y + pidx - 1);
                    py0 = to_db(ev + (1.0 - kappa) * (double)__fmul_rn(yp, yp), p.eps, inv_e0, p.floor_db);
                }
                found = 1;
                idx = __shfl_sync(kFullMask, pidx, fl);
                y1 = __shfl_sync(kFullMask, dbv, fl);
                y0 = __shfl_sync(kFullMask, py0, fl);
            }
        });
        if (!found && k + 1 < NT && (k + 1) * TL < n) {
            // every sample of the tile is above the target: the crossing is the first sample of tile k+1
            found = 1;
            idx = (k + 1) * TL;
            y1 = to_db(carry[k], p.eps, inv_e0, p.floor_db);
            const float yp = fit_ld1<kSmem>(y + idx - 1);
            y0 = to_db(carry[k] + (1.0 - kappa) * (double)__fmul_rn(yp, yp), p.eps, inv_e0, p.floor_db);
        }
    };
    // warp w resolves target w; targets 0 / 1 (0 dB, -10 dB) only serve early10, which band tables do not carry
    for (int tg = 2 + warp; tg < n_targets; tg += kW) {
        const double tau = p.target[tg];
        int found, idx;
        float y0, y1;
        resolve((float)tau, found, idx, y0, y1);
        if (lane == 0) {
            sh.cross_ok[tg] = found;
            sh.cross_idx[tg] = idx;
            sh.cross_t[tg] = found ? crossing_time_rel(idx - start, y0, y1, tau, srf) : 0.0;
        }
    }
    __syncthreads();
    // ---- fit index ranges, by the last warp (the certification targets after it keep the other warps busy meanwhile)
    if (warp == kW - 1 && lane < p.n_fits) {
        const int f = lane;
        const int ts_i = 2 + 2 * f, te_i = 3 + 2 * f;
        int first = 0, last = -1, valid;
        float wa = 1.0f, wb = 1.0f;
        if (a.decim > 1) {
            // The reference fits the full-rate samples i0..i1 between the crossing times (decay.py:224-228).  On the
            // decimated grid sample j stands for the D full-rate samples around it, so the fit takes every j whose
            // cell overlaps [i0, i1] and weights the first and the last by the covered fraction of their cell: which
            // end samples the grid happens to include then moves the slope ~50 x less (a band's curve is bumpy, and
            // its end points can sit several dB off the fitted line).
            const double D = (double)a.decim, sr_full = (double)p.sr * D;
            const double ts = sh.cross_t[ts_i], te = sh.cross_t[te_i];
            valid = sh.cross_ok[ts_i] && sh.cross_ok[te_i] && te > ts;
            const double i0 = ceil(ts * sr_full - 1e-6), i1 = floor(te * sr_full + 1e-6);
            valid = valid && (i1 - i0 + 1.0 >= 8.0);
            const double lo_r = i0 - 0.5, hi_r = i1 + 0.5;
            int ja = (int)floor((lo_r - 0.5 * D) / D) + 1, jb = (int)ceil((hi_r + 0.5 * D) / D) - 1;
            ja = max(ja, 0);
            jb = min(jb, n - 1);
            valid = valid && jb >= ja;
            if (valid) {
                wa = (float)((fmin(D * ja + 0.5 * D, hi_r) - fmax(D * ja - 0.5 * D, lo_r)) / D);
                wb = (jb > ja) ? (float)((fmin(D * jb + 0.5 * D, hi_r) - fmax(D * jb - 0.5 * D, lo_r)) / D) : wa;
            }
            first = ja;
            last = jb;
            sh.fit_n_full[f] = (int)(i1 - i0 + 1.0);
        } else {
            valid = fit_index_range(sh.cross_ok[ts_i], sh.cross_ok[te_i], sh.cross_t[ts_i], sh.cross_t[te_i],
                                    sh.cross_idx[ts_i] - start, sh.cross_idx[te_i] - start, srf, &first, &last,
                                    a.min_fit_points);
            sh.fit_n_full[f] = last - first + 1;
        }
        sh.fit_valid[f] = valid;
        sh.fit_first[f] = start + first;
        sh.fit_last[f] = start + last;
        sh.fit_wa[f] = wa;
        sh.fit_wb[f] = wb;
        sh.fit_ts[f] = sh.cross_t[ts_i];
        sh.fit_te[f] = sh.cross_t[te_i];
    }
    // ---- certification of a decimated unit, part 1.  The sub-sampled Schroeder sum differs from the reference's by
    // at most delta dB near a fit end point: the 2 f ripple of y^2 folds to another frequency on the decimated grid
    // (the host's per-band error model times a safety factor, scaled by the unit's own decay rate), and the far end of
    // the grid is known to about one cell of energy.  The reference's crossing then lies between the first decimated
    // samples <= target + delta and <= target - delta: resolve those four extra targets per fit.
    // (Runs beside the fit-range computation: it needs the crossings only.)
    if (a.cert.on) {
        const float m_b = a.cert.m[cj];
        const float yl = fit_ld1<kSmem>(y + (n - 1));
        const double end_cell = (double)__fmul_rn(yl, yl) * inv_e0;   // one far-end cell relative to E0
        for (int ct = warp; ct < 4 * p.n_fits; ct += kW) {
            const int f = ct >> 2, e = (ct >> 1) & 1, minus = ct & 1;
            const int ts_i = 2 + 2 * f, te_i = 3 + 2 * f;
            int cidx = -1;
            const int span = sh.cross_idx[te_i] - sh.cross_idx[ts_i];
            if (sh.cross_ok[ts_i] && sh.cross_ok[te_i] && span > 0) {   // (an invalid fit is flagged whatever these find)
                const double tau0 = p.target[e ? te_i : ts_i];
                // (a modelled bound: the quotient in float)
                const double delta = (double)(m_b * (float)(p.target[ts_i] - p.target[te_i]) / (float)span) +
                                     4.342944819 * end_cell * a.gain10[e ? te_i : ts_i];   // 10^(-tau0 / 10), from the host
                const double tau = tau0 + (minus ? -delta : delta);
                int found, idx;
                float y0, y1;
                resolve((float)tau, found, idx, y0, y1, sh.cross_idx[e ? te_i : ts_i] / TL);
                if (found) cidx = idx;
            }
            if (lane == 0) sh.cert_idx[ct] = cidx;
        }
    }
    __syncthreads();

    // ---- regression sums over the fit ranges (decay.py:224-249)
    double acc[AA_MAX_FITS * 3];
#pragma unroll
    for (int q = 0; q < AA_MAX_FITS * 3; ++q) acc[q] = 0.0;
    int ffirst[AA_MAX_FITS], flast[AA_MAX_FITS];
    float ymid_f[AA_MAX_FITS], fwa[AA_MAX_FITS], fwb[AA_MAX_FITS];
    int lo = 0x7fffffff, hi = -1;
#pragma unroll
    for (int f = 0; f < AA_MAX_FITS; ++f) {
        const bool on = f < p.n_fits && sh.fit_valid[f];
        ffirst[f] = on ? sh.fit_first[f] : 0x7fffffff;
        flast[f] = on ? sh.fit_last[f] : -1;
        ymid_f[f] = (float)(0.5 * (p.fit_hi[f] + p.fit_lo[f]));
        fwa[f] = on ? sh.fit_wa[f] : 1.0f;
        fwb[f] = on ? sh.fit_wb[f] : 1.0f;
        if (on) {
            lo = min(lo, ffirst[f]);
            hi = max(hi, flast[f]);
        }
    }
    if (hi >= lo) {
        const int k_lo = lo / TL, k_hi = min(hi / TL, NT - 1);
        const double eps = p.eps;
        const float floor_db = p.floor_db;
        for (int k = k_lo + warp; k <= k_hi; k += kW) {
            visit_tile_reverse<NS, kSmem>(y, k, TL, n, carry[k], lane, masks, [&](int j, const TileScan& sc, double after) {
                const int i0 = k * TL + j * kSub + 4 * lane;
                const int s_lo = k * TL + j * kSub, s_hi = s_lo + kSub - 1;
                bool any = false;
#pragma unroll
                for (int f = 0; f < AA_MAX_FITS; ++f) any |= (ffirst[f] <= s_hi && flast[f] >= s_lo);
                if (!any) return;  // warp-uniform
                // same arithmetic as phase 2 of edc_body.cuh: float64 ratio rounded once, MUFU log2 on the mantissa,
                // float sums per step promoted to the float64 accumulators
                float db[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float r = (float)(fmax(after + (double)__fadd_rn(sc.excl, sc.s[c]) -
                                                 kappa * (double)lane_sample_energy(sc, c), eps) * inv_e0);
                    db[c] = fmaxf(db10_from_ratio_fast(r), floor_db);
                }
                const int in_tile = TL - (j * kSub + 4 * lane);  // samples of this lane that belong to the tile
#pragma unroll
                for (int f = 0; f < AA_MAX_FITS; ++f) {
                    if (ffirst[f] <= s_hi && flast[f] >= s_lo) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int i = i0 + c;
                            const float yc = (c < in_tile && i >= ffirst[f] && i <= flast[f]) ? db[c] - ymid_f[f] : 0.0f;
                            // end weights (1 at full rate: the products below are then exact)
                            const float yw = yc * ((i == ffirst[f]) ? fwa[f] : (i == flast[f]) ? fwb[f] : 1.0f);
                            s0 += yw;
                            s1 = fmaf((float)c, yw, s1);
                            s2 = fmaf(yw, yc, s2);
                        }
                        const double ub = (double)(i0 - ffirst[f]);
                        acc[f * 3 + 0] += (double)s0;
                        acc[f * 3 + 1] += fma(ub, (double)s0, (double)s1);
                        acc[f * 3 + 2] += (double)s2;
                    }
                }
            });
        }
    }
#pragma unroll
    for (int q = 0; q < AA_MAX_FITS * 3; ++q) {
        if (q < 3 * p.n_fits) {   // (uniform) only the fits the call asked for: each sum is a chain of ten shuffles
            const double s = warp_sum(acc[q]);
            if (lane == 0) sh.warp_acc[warp][q] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < AA_MAX_FITS) {
        const int f = threadIdx.x;
        double rt60 = nan("");
        if (f < p.n_fits && sh.fit_valid[f]) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int w = 0; w < kW; ++w) {
                s0 += sh.warp_acc[w][f * 3 + 0];
                s1 += sh.warp_acc[w][f * 3 + 1];
                s2 += sh.warp_acc[w][f * 3 + 2];
            }
            if (a.decim > 1) {
                // weighted least squares over u = 0..m (u = j - first): only the end weights differ from 1
                const double m = (double)(sh.fit_last[f] - sh.fit_first[f]);
                const double wa = (double)sh.fit_wa[f], wb = (m > 0.0) ? (double)sh.fit_wb[f] : 0.0;
                const double inner = fmax(m - 1.0, 0.0);                       // u = 1..m-1
                const double W = wa + wb + inner;
                const double U1 = wb * m + 0.5 * (m - 1.0) * m * (m > 0.0 ? 1.0 : 0.0);
                const double U2 = wb * m * m + (m > 0.0 ? (m - 1.0) * m * (2.0 * m - 1.0) / 6.0 : 0.0);
                const double den = W * U2 - U1 * U1;
                if (den > 0.0) {
                    const double slope_u = (W * s1 - U1 * s0) / den;          // dB per decimated sample
                    sh.fit_slope_u[f] = slope_u;
                    sh.fit_b0[f] = (s0 - slope_u * U1) / W + (double)ymid_f[f];  // line value at u = 0
                    const double slope = slope_u * (double)p.sr;               // dB per second (decay.py:241-249)
                    if (slope < 0.0) rt60 = -60.0 / slope;
                }
                if (rt60 != rt60) sh.fit_valid[f] = 0;
            } else {
                double res[AA_FIT_COLS];
                rt60 = fit_finalize(s0, s1, s2, sh.fit_first[f] - start, sh.fit_last[f] - start, (double)ymid_f[f], p.sr,
                                    sh.fit_ts[f], sh.fit_te[f], res);
            }
        }
        rt60_out[f] = rt60;
    }
    // ---- certification, part 2: with the crossing windows [lo, hi] (decimated samples) and the fitted line, moving an
    // end point of the fit by `win` full-rate samples changes the slope by about 6 (win / n) (r / range), r = distance
    // of the curve from the line at that end.  The unit is certified when both ends together stay under cert.tol;
    // otherwise it goes on the list of the next, finer level (bands.cu), whose last level is the full-rate path.
    // (The finalising threads sit in warp 0: a warp-level barrier orders their shared-memory writes before the decision,
    // and the other warps are done.  The decision is a threshold on a modelled bound: float arithmetic.)
    if (a.cert.on && warp == 0) {
        __syncwarp();
        if (lane == 0) {
            bool ok = true;
            for (int f = 0; f < p.n_fits && ok; ++f) {
                const int ts_i = 2 + 2 * f, te_i = 3 + 2 * f;
                const int span = sh.cross_idx[te_i] - sh.cross_idx[ts_i];
                ok = sh.fit_valid[f] && (float)span >= (float)a.cert.min_span[cj] * a.cert.fit_scale[f];
                if (!ok) break;
                const float inv_range = 1.0f / (float)(p.target[ts_i] - p.target[te_i]);
                const float inv_n_full = 1.0f / (float)sh.fit_n_full[f];
                const float slope_u = (float)sh.fit_slope_u[f], b0 = (float)sh.fit_b0[f];
                float moved = 0.0f;
                for (int e = 0; e < 2 && ok; ++e) {
                    const int lo = sh.cert_idx[4 * f + 2 * e], hi = sh.cert_idx[4 * f + 2 * e + 1];
                    ok = lo >= 0 && hi >= 0 && hi - lo <= max(1, (int)(a.cert.w * (float)span));
                    const float tc = (float)((e ? sh.fit_te[f] : sh.fit_ts[f]) * (double)p.sr - (double)(sh.fit_first[f] - start));
                    const float line = fmaf(slope_u, tc, b0);
                    const float tgt = (float)p.target[e ? te_i : ts_i];
                    const float r = fmaxf(fabsf(tgt - line), 0.5f) + fabsf(slope_u) * (float)(hi - lo + 1);
                    // a start target at 0 dB is the first sample whatever the curve (decay.py:186-187)
                    const float win = (tgt >= 0.0f) ? 0.0f : ((float)(hi - lo) + 0.25f) * (float)a.decim;
                    moved += 6.0f * (win * inv_n_full) * (r * inv_range);
                }
                ok = ok && moved <= a.cert.tol;
            }
            if (!ok) flag_unit(a.cert, row, cj);
        }
    }
}

}  // namespace bandfit
}  // namespace aa
