// Crossing-time interpolation, fit index ranges and closed-form least squares shared by the
// fused EDC kernel (edc.cu) and the generic curve kernel (curvefit.cu).
#pragma once

#include "aa_common.cuh"

namespace aa {

// time_seconds[i] = float32(i) / float32(sr)   (analyse/decay.py:169)
__device__ __forceinline__ float rel_time(int i, float sr) { return __fdiv_rn((float)i, sr); }

// analyse/decay.py:186-199: i = index (relative to the curve start) of the first sample <= target
__device__ __forceinline__ double crossing_time_rel(int i, float y0f, float y1f, double target, float srf) {
    if (i == 0) return (double)rel_time(0, srf);
    const double t0 = (double)rel_time(i - 1, srf), t1 = (double)rel_time(i, srf);
    const double y0 = (double)y0f, y1 = (double)y1f;
    if (y1 == y0) return t1;
    double fr = (target - y0) / (y1 - y0);
    fr = fmin(fmax(fr, 0.0), 1.0);
    return t0 + fr * (t1 - t0);
}

// analyse/decay.py:224-228: mask = (time >= start_t) & (time <= end_t) on float32 values, >= 8 points.
// is / ie are the crossing sample indices (relative) of the range start / end targets.
__device__ __forceinline__ int fit_index_range(int ok_s, int ok_e, double ts, double te, int is, int ie, float srf,
                                               int* first, int* last, int min_points = 8) {
    *first = 0;
    *last = -1;
    if (!(ok_s && ok_e && te > ts)) return 0;
    const float tsf = (float)ts, tef = (float)te;
    int f = is;
    if (is > 0 && rel_time(is - 1, srf) >= tsf) f = is - 1;
    int l = ie;
    if (!(rel_time(ie, srf) <= tef)) l = ie - 1;
    *first = f;
    *last = l;
    return (l - f + 1 >= min_points) ? 1 : 0;
}

// Closed-form least squares over u = 0..n-1 (t = (first + u)/sr) from S0 = sum(yc), S1 = sum(u*yc),
// S2 = sum(yc^2) with yc = y - ymid.  Equivalent to np.linalg.lstsq on [t, 1] (analyse/decay.py:235-249).
// res: t_start, t_end, slope, intercept, r2, rt60, n_points, valid.  Returns rt60 or NaN.
__device__ __forceinline__ double fit_finalize(double s0, double s1, double s2, int first, int last, double ymid,
                                               int sr, double ts, double te, double* res) {
    for (int c = 0; c < 6; ++c) res[c] = nan("");
    const double n = (double)(last - first + 1);
    res[6] = n;
    res[7] = 0.0;
    const double ubar = 0.5 * (n - 1.0);
    const double suu = n * (n * n - 1.0) / 12.0;
    const double suy = s1 - ubar * s0;
    const double slope_u = suy / suu;             // dB per sample
    const double slope = slope_u * (double)sr;    // dB per second
    if (!(slope < 0.0)) return nan("");           // decay.py:241
    const double ybar = s0 / n + ymid;
    const double tbar = ((double)first + ubar) / (double)sr;
    const double ss_tot = s2 - s0 * s0 / n;
    const double ss_res = ss_tot - slope_u * suy;
    res[0] = ts;
    res[1] = te;
    res[2] = slope;
    res[3] = ybar - slope * tbar;
    res[4] = (ss_tot > 0.0) ? 1.0 - ss_res / ss_tot : 0.0;
    res[5] = -60.0 / slope;                       // decay.py:249
    res[7] = 1.0;
    return res[5];
}

}  // namespace aa
