// Register-resident DFTs of composite length (9 ... 25 points) for the two-stage FFT passes.
//
// A pass over a length L = RA * RB transform keeps each butterfly's RA (then RB) points in
// registers, so the data crosses shared memory once per pass instead of once per small-radix
// stage.  Composite radices are built from the small DFTs of fft_core.cuh either by the
// prime-factor (Good-Thomas) map, which needs no internal twiddles when the factors are
// coprime, or by one Cooley-Tukey step whose twiddles are compile-time constants.
// Every index below is a template constant: after inlining there is no indexing at run time.
// Host/device so that tests/csrc/fft_model_check.cu runs the same code on the CPU.
#pragma once

#include <type_traits>

#include "fft_core.cuh"

namespace aa {

// ---- compile-time roots of unity --------------------------------------------------------------
namespace cx {

constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double sin_series(double x) {  // |x| <= pi/2
    double term = x, sum = x;
    for (int n = 1; n < 16; ++n) {
        term *= -x * x / (double)((2 * n) * (2 * n + 1));
        sum += term;
    }
    return sum;
}
constexpr double cos_series(double x) {  // |x| <= pi/2
    double term = 1.0, sum = 1.0;
    for (int n = 1; n < 16; ++n) {
        term *= -x * x / (double)((2 * n - 1) * (2 * n));
        sum += term;
    }
    return sum;
}
// cos / sin of 2 pi m / L with exact argument reduction on the integers
constexpr double cos2pi(long long m, long long L) {
    m %= L;
    if (m < 0) m += L;
    if (2 * m > L) m = L - m;              // cos is even around pi
    if (4 * m > L) return -cos_series(kPi * (double)(L - 2 * m) / (double)L);  // cos(x) = -cos(pi - x)
    return cos_series(2.0 * kPi * (double)m / (double)L);
}
constexpr double sin2pi(long long m, long long L) {
    m %= L;
    if (m < 0) m += L;
    if (2 * m > L) return -sin2pi(L - m, L);
    if (4 * m > L) return sin_series(kPi * (double)(L - 2 * m) / (double)L);   // sin(x) = sin(pi - x)
    return sin_series(2.0 * kPi * (double)m / (double)L);
}
__host__ __device__ constexpr int mod_inverse(int a, int n) {  // a^-1 mod n, gcd(a, n) == 1
    for (int x = 1; x < n; ++x)
        if ((a * x) % n == 1) return x;
    return 1;
}

}  // namespace cx

template <int m, int L>
struct Root {  // exp(-2 pi i m / L)
    static constexpr float c = (float)cx::cos2pi(m, L);
    static constexpr float s = (float)(-cx::sin2pi(m, L));
};

template <int m, int L>
AA_HD float2 mul_root(float2 a) {
    constexpr int r = ((m % L) + L) % L;
    if constexpr (r == 0) {
        return a;
    } else if constexpr (4 * r == L) {
        return cmul_mi(a);
    } else if constexpr (2 * r == L) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * r == 3 * L) {
        return cmul_pi(a);
    } else {
        constexpr float c = Root<r, L>::c, s = Root<r, L>::s;
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

template <int I, int N, class F>
AA_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <int R>
AA_HD void dft_r(float2* v);

// Cooley-Tukey step, L = A * B:  X[ka + A kb] = sum_nb w_B^{nb kb} ( w_L^{nb ka} sum_na x[B na + nb] w_A^{na ka} )
template <int A, int B>
AA_HD void dft_ct(float2* v) {
    constexpr int L = A * B;
    float2 u[L];  // u[nb * A + ka]
    static_for<0, B>([&](auto nb_) {
        constexpr int nb = decltype(nb_)::value;
        float2 t[A];
        static_for<0, A>([&](auto na_) {
            constexpr int na = decltype(na_)::value;
            t[na] = v[B * na + nb];
        });
        dft_r<A>(t);
        static_for<0, A>([&](auto ka_) {
            constexpr int ka = decltype(ka_)::value;
            u[nb * A + ka] = mul_root<nb * ka, L>(t[ka]);
        });
    });
    static_for<0, A>([&](auto ka_) {
        constexpr int ka = decltype(ka_)::value;
        float2 s[B];
        static_for<0, B>([&](auto nb_) {
            constexpr int nb = decltype(nb_)::value;
            s[nb] = u[nb * A + ka];
        });
        dft_r<B>(s);
        static_for<0, B>([&](auto kb_) {
            constexpr int kb = decltype(kb_)::value;
            v[ka + A * kb] = s[kb];
        });
    });
}

// Prime-factor step, gcd(A, B) = 1:  n = (B na + A nb) mod L,  k = (B B' ka + A A' kb) mod L,
// B' = B^-1 mod A, A' = A^-1 mod B; no twiddles between the two sets of small DFTs.
template <int A, int B>
AA_HD void dft_pfa(float2* v) {
    constexpr int L = A * B;
    constexpr int Bi = cx::mod_inverse(B % A, A), Ai = cx::mod_inverse(A % B, B);
    float2 u[L];
    static_for<0, B>([&](auto nb_) {
        constexpr int nb = decltype(nb_)::value;
        float2 t[A];
        static_for<0, A>([&](auto na_) {
            constexpr int na = decltype(na_)::value;
            t[na] = v[(B * na + A * nb) % L];
        });
        dft_r<A>(t);
        static_for<0, A>([&](auto ka_) {
            constexpr int ka = decltype(ka_)::value;
            u[nb * A + ka] = t[ka];
        });
    });
    static_for<0, A>([&](auto ka_) {
        constexpr int ka = decltype(ka_)::value;
        float2 s[B];
        static_for<0, B>([&](auto nb_) {
            constexpr int nb = decltype(nb_)::value;
            s[nb] = u[nb * A + ka];
        });
        dft_r<B>(s);
        static_for<0, B>([&](auto kb_) {
            constexpr int kb = decltype(kb_)::value;
            v[(B * Bi * ka + A * Ai * kb) % L] = s[kb];
        });
    });
}

template <int R>
AA_HD void dft_r(float2* v) {
    if constexpr (R == 1) {
    } else if constexpr (R == 2) {
        dft2(v);
    } else if constexpr (R == 3) {
        dft3(v);
    } else if constexpr (R == 4) {
        dft4(v);
    } else if constexpr (R == 5) {
        dft5(v);
    } else if constexpr (R == 6) {
        dft6(v);
    } else if constexpr (R == 8) {
        dft8(v);
    } else if constexpr (R == 9) {
        dft_ct<3, 3>(v);
    } else if constexpr (R == 10) {
        dft_pfa<2, 5>(v);
    } else if constexpr (R == 12) {
        dft_pfa<3, 4>(v);
    } else if constexpr (R == 15) {
        dft_pfa<3, 5>(v);
    } else if constexpr (R == 16) {
        dft_ct<4, 4>(v);
    } else if constexpr (R == 18) {
        dft_pfa<2, 9>(v);
    } else if constexpr (R == 20) {
        dft_pfa<4, 5>(v);
    } else if constexpr (R == 24) {
        dft_pfa<3, 8>(v);
    } else if constexpr (R == 25) {
        dft_ct<5, 5>(v);
    } else {
        static_assert(R == 1, "unsupported register radix");
    }
}

}  // namespace aa
