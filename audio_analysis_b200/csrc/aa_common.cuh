// Shared helpers for the audio_analysis_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/audio_analysis_b200.h"

namespace aa {

void set_error(const char* fmt, ...);
void count_launch();  // every kernel launch of this library is counted (aa_kernel_launch_count)

// Optional per-stage device timing for bench.py (aa_profile_* in the C ABI): when enabled, launch sites wrap a
// kernel (or a short group of kernels) in a ProfScope, which records CUDA events on the launching stream.
bool prof_enabled();
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);
struct ProfScope {
    cudaStream_t s;
    bool on;
    ProfScope(const char* name, cudaStream_t stream) : s(stream), on(prof_enabled()) {
        if (on) prof_begin(name, s);
    }
    ~ProfScope() {
        if (on) prof_end(s);
    }
};

#define AA_CUDA_TRY(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::aa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return AA_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define AA_REQUIRE(cond, code, msg)                                  \
    do {                                                             \
        if (!(cond)) {                                               \
            ::aa::set_error("%s (%s:%d)", msg, __FILE__, __LINE__);  \
            return code;                                             \
        }                                                            \
    } while (0)

inline int launch_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return AA_ERR_CUDA;
    }
    count_launch();
    return AA_OK;
}

constexpr int kWarp = 32;
constexpr unsigned kFullMask = 0xffffffffu;

#ifdef __CUDACC__

// 10*log10(2): dB per octave of energy ratio.
__device__ __forceinline__ float db10_from_ratio(float r) { return 3.010299956639812f * log2f(r); }

// 10*log10(r) for a positive normal float32 r: log2 r = exponent + log2(mantissa), with the hardware log2
// (absolute error about 2^-22 on [1, 2)) applied to the mantissa only, so the error stays at 1e-6 dB where the
// plain fast intrinsic would lose two ulps of a value near -40.  The two roundings (product, fused sum) are
// monotone, which keeps a monotone input sequence monotone.
__device__ __forceinline__ float db10_from_ratio_fast(float r) {
    const int bits = __float_as_int(r);
    const float e = (float)((bits >> 23) - 127);
    const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    const float k = 3.010299956639812f;
    return __fmaf_rn(e, k, __fmul_rn(k, __log2f(m)));
}

// Inclusive -> exclusive suffix sum over the lanes of a warp (sum of strictly
// HIGHER lanes).  One fixed instruction sequence so that every caller gets
// bit-identical partial sums for the same inputs (the EDC tile carries rely on it).
__device__ __forceinline__ float warp_suffix_exclusive(float v) {
    float inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_down_sync(kFullMask, inc, d);
        if ((threadIdx.x & 31) + d < 32) inc = __fadd_rn(inc, o);
    }
    float ex = __shfl_down_sync(kFullMask, inc, 1);
    return ((threadIdx.x & 31) == 31) ? 0.0f : ex;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
    return v;
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming 128-bit store (output written once, never re-read by this kernel)
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

#endif  // __CUDACC__

}  // namespace aa
