// Batched complex FFT engine: plans, pass kernels, Bluestein fallback.
#pragma once

#include "aa_common.cuh"
#include "fft_access.cuh"

namespace aa {

struct SubPlan {
    SubPlanDev dev{};
    float2* d_tw = nullptr;
};

struct TwoLevelTable {
    TwoLevelTwiddle dev{};
    float2* d_hi = nullptr;
    float2* d_lo = nullptr;
};

struct C2CPlan {
    int M = 0;          // logical transform length
    bool bluestein = false;
    // direct two-pass plan (M = M1 * M2; M1 == 1 -> single pass)
    int M1 = 1, M2 = 1;
    SubPlan s1, s2;
    int G1 = 1, G2 = 1;    // transforms per CTA in pass 1 / pass 2
    int Lp1 = 1, Lp2 = 1;  // padded row length in shared memory
    size_t smem1 = 0, smem2 = 0;
    TwoLevelTable twM;     // exp(-2 pi i m / M)
    // Bluestein: length-M DFT as a length-P circular convolution (P = 2^p >= 2M-1)
    C2CPlan* inner = nullptr;
    float2* d_chirp = nullptr;  // exp(-i pi n^2 / M), n in [0, M)
    float2* d_H = nullptr;      // FFT_P of the wrapped conjugate chirp
};

// layout of the optional energy partials (DstDesc::energy); false when this plan cannot produce them
bool c2c_energy_layout(const C2CPlan* p, int* tile_complex, int* n_tiles, int* n_groups);
// true when c2c_exec can fuse the real-FFT untangle step into pass 2 (DstDesc::untangle_X)
bool c2c_can_untangle(const C2CPlan* p);
// true if n factors completely into the supported radices
bool is_smooth(long long n);
int build_two_level(long long len, TwoLevelTable* out);
void free_two_level(TwoLevelTable* t);

int c2c_plan_create(int M, C2CPlan** out);
void c2c_plan_destroy(C2CPlan* p);
// float2 elements of scratch needed for `units` simultaneous transforms
size_t c2c_scratch_elems(const C2CPlan* p, long long units);
// forward DFT (sign -) of `units` transforms; inverse = conj on both sides via the descriptors
int c2c_exec(const C2CPlan* p, SrcDesc src, DstDesc dst, long long units, float2* scratch, cudaStream_t stream);

inline SrcDesc make_src_complex(const float2* base, long long unit_stride, int n_valid, int conj) {
    SrcDesc s{};
    s.mode = SRC_COMPLEX;
    s.base = base;
    s.unit_stride = unit_stride;
    s.n_valid = n_valid;
    s.conj = conj;
    return s;
}

inline DstDesc make_dst_complex(float2* base, long long unit_stride, int n_valid, int conj, float scale) {
    DstDesc d{};
    d.mode = DST_COMPLEX;
    d.base = base;
    d.unit_stride = unit_stride;
    d.n_valid = n_valid;
    d.conj1 = conj;
    d.scale1 = scale;
    d.scale2 = 1.0f;
    return d;
}

}  // namespace aa
