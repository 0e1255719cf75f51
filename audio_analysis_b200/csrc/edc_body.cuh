// Fused Schroeder EDC + level crossings + decay fits for one row per thread-block cluster: the device code
// shared by edc_kernel (edc.cu) and the fused band kernel (bands_fused.cu).  See edc.cu for the algorithm.
#pragma once

#include <cooperative_groups.h>
#include <math.h>

#include "aa_common.cuh"
#include "fit_math.cuh"

namespace aa {
namespace edc {

namespace cg = cooperative_groups;

#ifndef AA_EDC_THREADS
#define AA_EDC_THREADS 256   /* 4 CTAs per SM: the serial resolve phase of one row hides behind the streaming phases of three others (512 x 2: 3.5 % slower, 384 x 2: 9 % slower) */
#endif
// The kernel body below is compiled once per VARIANT (translation unit): `base` (edc.cu: AA_EDC_THREADS threads, the
// row is read from global memory in both phases) and `staged` (edc_staged.cu: AA_EDC_STAGE = 1, each CTA of the cluster
// keeps its span of the row in shared memory after phase 1, so the row crosses HBM once).
#ifndef AA_EDC_VARIANT
#define AA_EDC_VARIANT base
#endif
#ifndef AA_EDC_STAGE
#define AA_EDC_STAGE 0
#endif
constexpr int kTile = 128;  // samples per warp tile: one float4 per lane
constexpr int kMaxTargets = 2 + 2 * AA_MAX_FITS;
constexpr int kMaxCluster = 8;
constexpr int kUnroll = 4;  // tiles loaded ahead per warp (memory-level parallelism)

struct EdcArgs {
    const float* x;
    long long row_stride;
    const int* lengths;
    int uniform_len;
    int max_len;
    const int* start_in;
    int start_div;  // start_in is indexed by row / start_div (band units share their row's start)
    float* edc_out;
    long long edc_stride;
    int* start_out;
    double* fits_out;     // [row][AA_MAX_FITS][AA_FIT_COLS]   (metrics_mode 0)
    double* rt60_out;     // [row][AA_MAX_FITS]                (metrics_mode 1)
    double* early10_out;
    int* status_out;
    int sr;
    int trim_to_peak;
    int ignore;
    int min_samples;
    int tiles_per_warp;
    float floor_db;
    double eps;
    int n_fits;
    int want_early10;
    double target[kMaxTargets];  // 0: 0 dB, 1: -10 dB, 2+2f: fit hi, 3+2f: effective fit lo
    double fit_hi[AA_MAX_FITS], fit_lo[AA_MAX_FITS];
};

struct Crossing {
    int found;
    int idx;  // absolute sample index of the first sample <= target
    float y0, y1;
};

namespace AA_EDC_VARIANT {
constexpr int kThreads = AA_EDC_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr bool kStage = AA_EDC_STAGE != 0;

struct CtaShared {
    float warp_amax[kWarps];
    int warp_aidx[kWarps];
    double warp_tot[kWarps];
    // published to the cluster
    float cta_amax;
    int cta_aidx;
    double cta_total;
    double e0;
    Crossing cand[kMaxTargets];
    double fit_sums[AA_MAX_FITS][3];
    // resolved, identical in every CTA of the cluster
    int start;
    double cta_carry;
    double inv_e0;
    int fit_valid[AA_MAX_FITS];
    int fit_first[AA_MAX_FITS];  // absolute sample index range of the fit mask
    int fit_last[AA_MAX_FITS];
    double fit_ts[AA_MAX_FITS], fit_te[AA_MAX_FITS];
    double cross_t[kMaxTargets];
    int cross_ok[kMaxTargets];
    int cross_idx[kMaxTargets];
    double warp_acc[kWarps][AA_MAX_FITS * 3];
};
}  // namespace AA_EDC_VARIANT

struct TileScan {
    float s[4];  // suffix sums inside the lane's 4 samples
    float excl;  // sum over higher lanes
};

template <bool kCg>
__device__ __forceinline__ float4 load_tile(const float* __restrict__ x, int tile, int lane, int len) {
    const int i0 = tile * kTile + lane * 4;
    if (i0 + 3 < len) return kCg ? __ldcg(reinterpret_cast<const float4*>(x + i0)) : ldg_f4(x + i0);
    float4 v;
    v.x = (i0 < len) ? (kCg ? __ldcg(x + i0) : __ldg(x + i0)) : 0.f;
    v.y = (i0 + 1 < len) ? (kCg ? __ldcg(x + i0 + 1) : __ldg(x + i0 + 1)) : 0.f;
    v.z = (i0 + 2 < len) ? (kCg ? __ldcg(x + i0 + 2) : __ldg(x + i0 + 2)) : 0.f;
    v.w = 0.f;
    return v;
}

// Per-lane 0/1 factors of the warp suffix scan: step d adds the value of lane + d when that lane exists.
// inc = fma(o, 1, inc) rounds exactly like inc + o, and fma(o, 0, inc) leaves inc alone, so the scan costs one
// shuffle and one FMA per step with no predicates.
struct ScanMasks {
    float m[5];   // d = 1, 2, 4, 8, 16
    float last;   // 0 for lane 31 (nothing above it), 1 otherwise
};

__device__ __forceinline__ ScanMasks make_scan_masks() {
    ScanMasks k;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 5; ++i) k.m[i] = (lane + (1 << i) < 32) ? 1.0f : 0.0f;
    k.last = (lane == 31) ? 0.0f : 1.0f;
    return k;
}

// The single scan routine shared by both phases (bit-identical partial sums).
__device__ __forceinline__ TileScan tile_scan(float4 v, const ScanMasks& k) {
    TileScan t;
    t.s[3] = __fmul_rn(v.w, v.w);
    t.s[2] = __fmaf_rn(v.z, v.z, t.s[3]);
    t.s[1] = __fmaf_rn(v.y, v.y, t.s[2]);
    t.s[0] = __fmaf_rn(v.x, v.x, t.s[1]);
    float inc = t.s[0];
#pragma unroll
    for (int i = 0; i < 5; ++i) inc = __fmaf_rn(__shfl_down_sync(kFullMask, inc, 1 << i), k.m[i], inc);
    t.excl = __fmul_rn(__shfl_down_sync(kFullMask, inc, 1), k.last);
    return t;
}

__device__ __forceinline__ float to_db(double e, double eps, double inv_e0, float floor_db) {
    float r = (float)(fmax(e, eps) * inv_e0);  // decay.py:154,157
    return fmaxf(db10_from_ratio(r), floor_db);  // decay.py:158,167
}

namespace AA_EDC_VARIANT {

// One row analysed by the calling cluster (every thread of every CTA of the cluster calls this).
//   row      index into the per-row arguments (lengths, start_in, outputs)
//   x        the row's samples
//   kCg      read x through L2 only (ld.global.cg): the fused band kernel re-uses one private buffer for
//            successive band signals written by other SMs, so L1 / the read-only path could be stale
//   dyn_smem tiles_per_warp * kWarps * 12 bytes (float64 carries + float32 tile sums)
// Ends with a cluster barrier: when it returns no CTA of the cluster reads x or `sh` any more.
template <bool kWriteDb, bool kCg>
__device__ __forceinline__ void edc_row(const EdcArgs& a, int row, const float* __restrict__ x, unsigned char* dyn_smem,
                                        CtaShared& sh) {
    cg::cluster_group cluster = cg::this_cluster();
    const int T = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const ScanMasks masks = make_scan_masks();

    const int len = min(max(a.lengths ? a.lengths[row] : a.uniform_len, 0), a.max_len);
    const int nt = (len + kTile - 1) / kTile;
    const int tpw = a.tiles_per_warp;
    const int tpc = tpw * kWarps;
    double* carry = reinterpret_cast<double*>(dyn_smem);            // [tpc] exclusive suffix (this CTA only)
    float* tile_sum = reinterpret_cast<float*>(carry + tpc);         // [tpc]
    // staged variant: the CTA's span of the row, one float4 per lane and tile, written in phase 1 and read by
    // everything after it (the tile reload of the peak, E0, the crossings, phase 2)
    [[maybe_unused]] float4* stage = reinterpret_cast<float4*>(dyn_smem + (((size_t)tpc * 12 + 15) / 16) * 16);
    const int cta_tile0 = rank * tpc;
    auto tile_again = [&](int k) -> float4 {
        if constexpr (kStage) return stage[(k - cta_tile0) * 32 + lane];
        else return load_tile<kCg>(x, k, lane, len);
    };
    const int my_tile0 = cta_tile0 + warp * tpw;
    const int my_tile1 = min(nt, my_tile0 + tpw);

    const bool find_peak = (a.start_in == nullptr) && a.trim_to_peak;
    int start_known = -1;
    if (a.start_in) start_known = min(max(a.start_in[row / a.start_div], 0), len);
    else if (!a.trim_to_peak) start_known = min(a.ignore, len);

    // ------------------------------------------------------------------ phase 1
    float amax = -1.0f;   // per lane: largest |x| seen and the first tile that holds it (index resolved below)
    int atile = -1;
    int aidx = 0x7fffffff;
    {
        const int k_skip = (start_known >= 0) ? start_known / kTile : 0;
        for (int kb = my_tile0; kb < my_tile0 + tpw; kb += kUnroll) {
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {  // all loads first: kUnroll x 512 B in flight per warp
                const int k = kb + u;
                v[u] = (k < my_tile1 && k >= k_skip) ? load_tile<kCg>(x, k, lane, len) : make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (kStage) {
                    if (k < my_tile1 && k >= k_skip) stage[(k - cta_tile0) * 32 + lane] = v[u];
                }
            }
            if constexpr (!kWriteDb) {
                // metrics only: tile totals by one 4-way transpose-reduce (6 shuffles for 4 tiles).  The totals
                // may differ from phase 2's scan totals in the last bit; nothing is stored, so curve
                // monotonicity across tile boundaries is not at stake (the crossing search tolerates it).
                float ts[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int k = kb + u;
                    ts[u] = 0.f;
                    if (k < my_tile1 && k >= k_skip) {
                        if (find_peak) {
                            const float m = fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
                            if (m > amax) { amax = m; atile = k; }
                        }
                        ts[u] = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
                    }
                }
                static_assert(kUnroll == 4, "transpose-reduce is written for 4 tiles");
                const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
                float a0 = hi16 ? ts[2] : ts[0], a1 = hi16 ? ts[3] : ts[1];
                a0 += __shfl_xor_sync(kFullMask, hi16 ? ts[0] : ts[2], 16);
                a1 += __shfl_xor_sync(kFullMask, hi16 ? ts[1] : ts[3], 16);
                float c = hi8 ? a1 : a0;
                c += __shfl_xor_sync(kFullMask, hi8 ? a0 : a1, 8);
                c += __shfl_xor_sync(kFullMask, c, 4);
                c += __shfl_xor_sync(kFullMask, c, 2);
                c += __shfl_xor_sync(kFullMask, c, 1);
                const int u_own = lane >> 3;  // lanes 0, 8, 16, 24 own tiles kb + 0..3
                if ((lane & 7) == 0 && kb + u_own < my_tile0 + tpw) tile_sum[kb + u_own - cta_tile0] = c;
            } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int k = kb + u;
                if (k >= my_tile0 + tpw) break;
                float tsum = 0.f;
                if (k < my_tile1 && k >= k_skip) {
                    if (find_peak) {
                        const float m = fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
                        if (m > amax) { amax = m; atile = k; }
                    }
                    TileScan t = tile_scan(v[u], masks);
                    tsum = __fadd_rn(t.excl, t.s[0]);  // lane 0: whole-tile sum, same ops as phase 2
                }
                if (lane == 0) tile_sum[k - cta_tile0] = tsum;
            }
            }
        }
    }
    if (find_peak) {
        if (atile >= 0) {  // first sample of the lane's four that attains its maximum (first index wins ties, decay.py:136)
            const float4 w = tile_again(atile);
            const int i0 = atile * kTile + lane * 4;
            aidx = (fabsf(w.x) == amax) ? i0 : (fabsf(w.y) == amax) ? i0 + 1 : (fabsf(w.z) == amax) ? i0 + 2 : i0 + 3;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            float om = __shfl_xor_sync(kFullMask, amax, d);
            int oi = __shfl_xor_sync(kFullMask, aidx, d);
            if (om > amax || (om == amax && oi < aidx)) { amax = om; aidx = oi; }
        }
        if (lane == 0) { sh.warp_amax[warp] = amax; sh.warp_aidx[warp] = aidx; }
    }
    __syncthreads();

    // float64 exclusive suffix scan of the tile sums of this CTA
    {
        const int per = (tpc + kThreads - 1) / kThreads;
        const int t0 = threadIdx.x * per;
        double mine = 0.0;
        for (int q = 0; q < per; ++q) {
            int k = t0 + q;
            if (k < tpc) mine += (double)tile_sum[k];
        }
        // suffix over threads: inclusive via shuffles, then across warps
        double inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double o = __shfl_down_sync(kFullMask, inc, d);
            if (lane + d < 32) inc += o;
        }
        if (lane == 0) sh.warp_tot[warp] = inc;
        __syncthreads();
        double higher = 0.0;
        for (int w = warp + 1; w < kWarps; ++w) higher += sh.warp_tot[w];
        double after = __shfl_down_sync(kFullMask, inc, 1);
        if (lane == 31) after = 0.0;
        double running = higher + after;  // sum of everything after this thread's tiles
        for (int q = per - 1; q >= 0; --q) {
            int k = t0 + q;
            if (k < tpc) {
                carry[k] = running;
                running += (double)tile_sum[k];
            }
        }
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < kWarps; ++w) tot += sh.warp_tot[w];
            sh.cta_total = tot;
            if (find_peak) {
                float bm = -1.0f;
                int bi = 0x7fffffff;
                for (int w = 0; w < kWarps; ++w) {
                    float om = sh.warp_amax[w];
                    int oi = sh.warp_aidx[w];
                    if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
                }
                sh.cta_amax = bm;
                sh.cta_aidx = bi;
            }
        }
    }
    cluster.sync();  // A: (peak, energy) of every CTA visible

    if (warp == 0) {   // lane r reads CTA r's record: one remote-memory latency instead of a chain of 3 T loads
        double tot = 0.0;
        float bm = -1.0f;
        int bi = 0x7fffffff;
        if (lane < T) {
            const CtaShared* o = cluster.map_shared_rank(&sh, lane);
            tot = o->cta_total;
            if (find_peak) {
                bm = o->cta_amax;
                bi = o->cta_aidx;
            }
        }
        const double cc = warp_sum((lane > rank && lane < T) ? tot : 0.0);
        if (find_peak) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float om = __shfl_xor_sync(kFullMask, bm, d);
                const int oi = __shfl_xor_sync(kFullMask, bi, d);
                if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
            }
        }
        if (lane == 0) {
            int start = start_known;
            if (find_peak) {
                const int peak = (bi == 0x7fffffff) ? 0 : bi;
                start = min(len, peak + a.ignore);  // decay.py:140-144
            }
            sh.start = start;
            sh.cta_carry = cc;
        }
    }
    __syncthreads();
    const int start = sh.start;
    const double cta_carry = sh.cta_carry;
    const int n_an = len - start;  // analysed samples
    const float srf = (float)a.sr;

    if (n_an < a.min_samples) {  // decay.py:146 / rt60bands.py:395: caller turns this into ValueError / None
        if (rank == 0 && threadIdx.x == 0) {
            if (a.start_out) a.start_out[row] = start;
            if (a.status_out) a.status_out[row] = AA_ROW_TOO_SHORT;
            if (a.early10_out) a.early10_out[row] = nan("");
            for (int f = 0; f < AA_MAX_FITS; ++f) {
                if (a.fits_out) {
                    double* o = a.fits_out + ((long long)row * AA_MAX_FITS + f) * AA_FIT_COLS;
                    for (int c = 0; c < AA_FIT_COLS; ++c) o[c] = (c >= 6) ? 0.0 : nan("");
                }
                if (a.rt60_out) a.rt60_out[(long long)row * AA_MAX_FITS + f] = nan("");
            }
        }
        cluster.sync();  // nobody leaves while a peer may still read its shared memory
        return;
    }

    const int start_tile = start / kTile;
    const int owner = start_tile / tpc;

    // ------------------------------------------------------------------ E0 = E[start]
    if (rank == owner && warp == 0) {
        float4 v = tile_again(start_tile);
        TileScan t = tile_scan(v, masks);
        const int p = start - start_tile * kTile;
        const int pl = p >> 2, pj = p & 3;
        float sj = (pj == 0) ? t.s[0] : (pj == 1) ? t.s[1] : (pj == 2) ? t.s[2] : t.s[3];
        float loc = __fadd_rn(t.excl, sj);
        loc = __shfl_sync(kFullMask, loc, pl);
        if (lane == 0) sh.e0 = (carry[start_tile - cta_tile0] + cta_carry) + (double)loc;
    }
    cluster.sync();  // C: E0 visible
    if (threadIdx.x == 0) {
        const CtaShared* o = cluster.map_shared_rank(&sh, owner);
        sh.inv_e0 = 1.0 / fmax(o->e0, a.eps);
    }
    __syncthreads();
    const double inv_e0 = sh.inv_e0;
    const int n_targets = 2 + 2 * a.n_fits;

    // ------------------------------------------------------------------ crossings (decay.py:173-199)
    if (warp < n_targets) {
        const double tau = a.target[warp];
        const float tauf = (float)tau;
        Crossing c;
        c.found = 0; c.idx = 0; c.y0 = 0.f; c.y1 = 0.f;
        if (tauf >= 0.0f) {
            // curve starts at exactly 0 dB: the crossing is the first sample (decay.py:186-187)
            if (rank == owner) { c.found = 1; c.idx = start; }
        } else {
            // tile k holds the crossing iff dB(first valid sample of k) > tau >= dB(first sample of k+1)
            const int k_lo = max(cta_tile0, start_tile);
            const int k_hi = min(nt, cta_tile0 + tpc);
            int hit = -1;
            for (int kb = k_lo; kb < k_hi && hit < 0; kb += 32) {
                const int k = kb + lane;
                bool ok = false;
                if (k < k_hi) {
                    const int kl = k - cta_tile0;
                    float b_next = to_db(carry[kl] + cta_carry, a.eps, inv_e0, a.floor_db);
                    float b_first = 0.0f;
                    if (k > start_tile)
                        b_first = to_db((carry[kl] + cta_carry) + (double)tile_sum[kl], a.eps, inv_e0, a.floor_db);
                    ok = (b_first > tauf) && (b_next <= tauf);
                }
                unsigned m = __ballot_sync(kFullMask, ok);
                if (m) hit = kb + (__ffs(m) - 1);
            }
            if (hit >= 0) {
                const int k = hit;
                const double base = carry[k - cta_tile0] + cta_carry;
                float4 v = tile_again(k);
                TileScan t = tile_scan(v, masks);
                float db[4];
                int first_j = 4;
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    const int idx = k * kTile + lane * 4 + j;
                    db[j] = to_db(base + (double)__fadd_rn(t.excl, t.s[j]), a.eps, inv_e0, a.floor_db);
                    if (idx >= start && idx < len && db[j] <= tauf) first_j = j;
                }
                unsigned m = __ballot_sync(kFullMask, first_j < 4);
                if (m) {
                    const int fl = __ffs(m) - 1;
                    const int fj = __shfl_sync(kFullMask, first_j, fl);
                    const int idx = k * kTile + fl * 4 + fj;
                    // value at idx and at idx-1 (same tile: idx is never the first valid sample of the tile)
                    const int pl = (idx - 1 - k * kTile) >> 2, pj = (idx - 1 - k * kTile) & 3;
                    float y1 = (fj == 0) ? db[0] : (fj == 1) ? db[1] : (fj == 2) ? db[2] : db[3];
                    y1 = __shfl_sync(kFullMask, y1, fl);
                    float y0 = (pj == 0) ? db[0] : (pj == 1) ? db[1] : (pj == 2) ? db[2] : db[3];
                    y0 = __shfl_sync(kFullMask, y0, pl);
                    c.found = 1; c.idx = idx; c.y0 = y0; c.y1 = y1;
                } else if (k + 1 < nt) {
                    // every sample of this tile is above the target: the crossing is the first sample of tile k+1
                    float y0 = __shfl_sync(kFullMask, db[3], 31);
                    c.found = 1; c.idx = (k + 1) * kTile; c.y0 = y0;
                    c.y1 = to_db(base, a.eps, inv_e0, a.floor_db);
                }
            }
        }
        if (lane == 0) sh.cand[warp] = c;
    }
    cluster.sync();  // D: candidates visible

    if (warp < n_targets) {   // warp tg collects target tg: lane r reads CTA r's candidate, the lowest rank that found one wins
        const int tg = warp;
        Crossing c;
        c.found = 0; c.idx = 0; c.y0 = 0.f; c.y1 = 0.f;
        if (lane < T) c = cluster.map_shared_rank(&sh, lane)->cand[tg];
        const unsigned m = __ballot_sync(kFullMask, c.found != 0);
        const int src = m ? (__ffs(m) - 1) : 0;
        c.found = m ? 1 : 0;
        c.idx = __shfl_sync(kFullMask, c.idx, src);
        c.y0 = __shfl_sync(kFullMask, c.y0, src);
        c.y1 = __shfl_sync(kFullMask, c.y1, src);
        if (lane == 0) {
            const double tc = c.found ? crossing_time_rel(c.idx - start, c.y0, c.y1, a.target[tg], srf) : 0.0;
            sh.cross_ok[tg] = c.found;
            sh.cross_t[tg] = tc;
            sh.cross_idx[tg] = c.idx;
        }
    }
    __syncthreads();
    if (threadIdx.x < a.n_fits) {
        const int f = threadIdx.x;
        const int ts_i = 2 + 2 * f, te_i = 3 + 2 * f;
        int first = 0, last = -1;
        const int valid = fit_index_range(sh.cross_ok[ts_i], sh.cross_ok[te_i], sh.cross_t[ts_i], sh.cross_t[te_i],
                                          sh.cross_idx[ts_i] - start, sh.cross_idx[te_i] - start, srf, &first, &last);
        sh.fit_valid[f] = valid;
        sh.fit_first[f] = start + first;
        sh.fit_last[f] = start + last;
        sh.fit_ts[f] = sh.cross_t[ts_i];
        sh.fit_te[f] = sh.cross_t[te_i];
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2
    // Per sample: ratio = (base + in-tile suffix) / E0 in float64 rounded once to float32, then 10 log10 through
    // the exponent / mantissa split and one MUFU log2 (db10_from_ratio_fast), floor.  Per tile and fit: float sums of (y - ymid), j (y - ymid), (y - ymid)^2 over the lane's four
    // samples, promoted to the float64 accumulators once per tile.
    double acc[AA_MAX_FITS * 3];
#pragma unroll
    for (int q = 0; q < AA_MAX_FITS * 3; ++q) acc[q] = 0.0;
    int ffirst[AA_MAX_FITS], flast[AA_MAX_FITS];
    float ymid_f[AA_MAX_FITS];
#pragma unroll
    for (int f = 0; f < AA_MAX_FITS; ++f) {
        const bool on = f < a.n_fits && sh.fit_valid[f];
        ffirst[f] = on ? sh.fit_first[f] : 0x7fffffff;
        flast[f] = on ? sh.fit_last[f] : -1;
        ymid_f[f] = (float)(0.5 * (a.fit_hi[f] + a.fit_lo[f]));
    }
    float* __restrict__ out = kWriteDb ? a.edc_out + (long long)row * a.edc_stride : nullptr;
    const float floor_db = a.floor_db;

    for (int kb = max(my_tile0, start_tile); kb < my_tile1; kb += kUnroll) {
        float4 v[kUnroll];
        bool need[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int k = kb + u;
            const int t_lo = k * kTile, t_hi = t_lo + kTile - 1;
            bool fit_here = false;
#pragma unroll
            for (int f = 0; f < AA_MAX_FITS; ++f) fit_here |= (ffirst[f] <= t_hi && flast[f] >= t_lo);
            need[u] = (k < my_tile1) && (kWriteDb || fit_here);
            v[u] = need[u] ? tile_again(k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!need[u]) continue;  // warp-uniform
            const int k = kb + u;
            const int t_lo = k * kTile, t_hi = t_lo + kTile - 1;
            const TileScan t = tile_scan(v[u], masks);
            const double base = carry[k - cta_tile0] + cta_carry;
            float db[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // the ratio is formed in float64 and rounded once, so the float32 curve inherits the exact
                // monotonicity of the suffix sums (decay.py:151-157)
                const float r = (float)(fmax(base + (double)__fadd_rn(t.excl, t.s[j]), a.eps) * inv_e0);
                db[j] = fmaxf(db10_from_ratio_fast(r), floor_db);   // decay.py:158,167
            }
            const int i0 = t_lo + lane * 4;
            if (kWriteDb) {
                if (i0 >= start && i0 + 3 < len) {
                    st_cs_f4(out + i0, make_float4(db[0], db[1], db[2], db[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + j >= start && i0 + j < len) out[i0 + j] = db[j];
                }
            }
#pragma unroll
            for (int f = 0; f < AA_MAX_FITS; ++f) {
                if (ffirst[f] <= t_hi && flast[f] >= t_lo) {  // warp-uniform
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int idx = i0 + j;
                        const float yc = (idx >= ffirst[f] && idx <= flast[f]) ? db[j] - ymid_f[f] : 0.0f;
                        s0 += yc;
                        s1 = fmaf((float)j, yc, s1);
                        s2 = fmaf(yc, yc, s2);
                    }
                    const double ub = (double)(i0 - ffirst[f]);  // u of the lane's first sample (may be negative)
                    acc[f * 3 + 0] += (double)s0;
                    acc[f * 3 + 1] += fma(ub, (double)s0, (double)s1);
                    acc[f * 3 + 2] += (double)s2;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < AA_MAX_FITS * 3; ++q) {
        double s = warp_sum(acc[q]);
        if (lane == 0) sh.warp_acc[warp][q] = s;
    }
    __syncthreads();
    if (threadIdx.x < AA_MAX_FITS * 3) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sh.warp_acc[w][threadIdx.x];
        sh.fit_sums[threadIdx.x / 3][threadIdx.x % 3] = s;
    }
    cluster.sync();  // E: per-CTA regression sums visible

    if (rank == 0 && warp == 0) {
        // lane r reads CTA r's nine regression sums (independent remote loads), then nine warp reductions
        double part[AA_MAX_FITS * 3];
#pragma unroll
        for (int q = 0; q < AA_MAX_FITS * 3; ++q) part[q] = 0.0;
        if (lane < T) {
            const CtaShared* o = cluster.map_shared_rank(&sh, lane);
#pragma unroll
            for (int q = 0; q < AA_MAX_FITS * 3; ++q) part[q] = o->fit_sums[q / 3][q % 3];
        }
#pragma unroll
        for (int q = 0; q < AA_MAX_FITS * 3; ++q) part[q] = warp_sum(part[q]);
        if (lane == 0) {
            if (a.start_out) a.start_out[row] = start;
            if (a.status_out) a.status_out[row] = AA_ROW_OK;
            if (a.early10_out) {
                double e10 = nan("");
                if (a.want_early10 && sh.cross_ok[0] && sh.cross_ok[1] && sh.cross_t[1] >= sh.cross_t[0])
                    e10 = sh.cross_t[1] - sh.cross_t[0];  // decay.py:283-286
                a.early10_out[row] = e10;
            }
            for (int f = 0; f < AA_MAX_FITS; ++f) {
                double res[AA_FIT_COLS];
                for (int c = 0; c < 6; ++c) res[c] = nan("");
                res[6] = 0.0;
                res[7] = 0.0;
                double rt60 = nan("");
                if (f < a.n_fits && sh.fit_valid[f])
                    rt60 = fit_finalize(part[f * 3 + 0], part[f * 3 + 1], part[f * 3 + 2], sh.fit_first[f] - start,
                                        sh.fit_last[f] - start, (double)ymid_f[f], a.sr, sh.fit_ts[f], sh.fit_te[f], res);
                if (a.fits_out) {
                    double* o = a.fits_out + ((long long)row * AA_MAX_FITS + f) * AA_FIT_COLS;
                    for (int c = 0; c < AA_FIT_COLS; ++c) o[c] = res[c];
                }
                if (a.rt60_out) a.rt60_out[(long long)row * AA_MAX_FITS + f] = rt60;
            }
        }
    }
    cluster.sync();  // F: rank 0 has finished reading remote shared memory
}

}  // namespace AA_EDC_VARIANT

// Settings -> kernel arguments (the per-call pointers are filled in by the caller).
inline int fill_edc_args(const aa_decay_params* p, EdcArgs* out) {
    EdcArgs& a = *out;
    a.sr = p->sample_rate_hz;
    a.trim_to_peak = p->trim_to_peak;
    a.floor_db = (float)p->edc_floor_db;
    a.eps = p->edc_epsilon;
    a.n_fits = p->n_fits;
    a.want_early10 = p->want_early10;
    a.target[0] = 0.0;
    a.target[1] = -10.0;
    for (int f = 0; f < AA_MAX_FITS; ++f) {
        double hi = f < p->n_fits ? p->fit_hi_db[f] : 0.0;
        double lo = f < p->n_fits ? p->fit_lo_db[f] : 0.0;
        if (f < p->n_fits) AA_REQUIRE(lo <= hi, AA_ERR_INVALID_ARGUMENT, "fit range must be (higher_db, lower_db)");
        a.fit_hi[f] = hi;
        a.fit_lo[f] = lo;
        a.target[2 + 2 * f] = hi;
        a.target[3 + 2 * f] = lo > p->fit_lower_limit_db ? lo : p->fit_lower_limit_db;  // decay.py:219
    }
    return AA_OK;
}

}  // namespace edc
}  // namespace aa
