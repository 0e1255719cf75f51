// Compile-time specialised Stockham stage chains (shared by fft.cu and csd.cu).
#pragma once

#include "fft_core.cuh"

namespace aa {

// compact twiddles for a compile-time radix chain, built from the full exp(-2 pi i m / L) table
template <int L, int Ns>
__device__ __forceinline__ void spec_twiddles(float2* tw_s, const float2* __restrict__ tw_g) {}

template <int L, int Ns, int R, int... Rest>
__device__ __forceinline__ void spec_twiddles(float2* tw_s, const float2* __restrict__ tw_g) {
    constexpr int tmul = L / (Ns * R);
    if (Ns > 1) {
        float2* t = tw_s + compact_twiddle_offset(Ns);
        for (int m = threadIdx.x; m < Ns * (R - 1); m += blockDim.x) {
            const int r1 = m / Ns, k = m % Ns;
            t[m] = __ldg(tw_g + (r1 + 1) * k * tmul);
        }
    }
    spec_twiddles<L, Ns * R, Rest...>(tw_s, tw_g);
}

// G transforms of length L side by side (row pitch Lp), ping-pong between a and b; ends with a barrier
template <int kThreads, int kPad, int L, int G, int Lp, int Ns>
__device__ __forceinline__ void spec_stages(float2*& a, float2*& b, const float2* tw_s, const float2* tw_full) {}

template <int kThreads, int kPad, int L, int G, int Lp, int Ns, int R, int... Rest>
__device__ __forceinline__ void spec_stages(float2*& a, float2*& b, const float2* __restrict__ tw_s,
                                            const float2* __restrict__ tw_full) {
    constexpr int nb = L / R, total = G * nb;
    const float2* twc = tw_s + compact_twiddle_offset(Ns);
#pragma unroll 2
    for (int e = threadIdx.x; e < total; e += kThreads) {
        const int g = e / nb;  // constant divisors: multiply-shift
        const int j = e - g * nb;
        const int k = j % Ns;
        stockham_butterfly_compact<R, kPad>(j, k, nb, L, Ns, a + g * Lp, b + g * Lp, twc, tw_full);
    }
    __syncthreads();
    float2* t = a;
    a = b;
    b = t;
    spec_stages<kThreads, kPad, L, G, Lp, Ns * R, Rest...>(a, b, tw_s, tw_full);
}

}  // namespace aa
