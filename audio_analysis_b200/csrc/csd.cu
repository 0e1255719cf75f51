// Waterfall / cumulative-spectral-decay slices (sm_100a).
//
// Replaces analyse/waterfall.py:188-230 (windowed rFFT magnitude in dB) for the
// <= num_slices frames that reach the result, and analyse/waterfall.py:289-341
// (frequency crop, subtraction of the global / per-slice maximum, clip).  The
// reference computes every STFT frame and then keeps <= 18 of them.
//
// One CTA owns one row (channel): for each selected frame it stages n_fft
// samples in shared memory with 128-bit loads (aligned superset, shifted while
// windowing), runs an n_fft/2-point complex Stockham FFT in shared memory,
// untangles to the real spectrum, converts |X| to dB and stores the cropped
// bins.  The running maximum stays on chip; a final in-place pass over the
// CTA's own (L2-resident) output applies the reference level and the clip.
#include <math.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <vector>

#include "aa_common.cuh"
#include "fft_access.cuh"
#include "fft_plan_host.hpp"
#include "fft_spec.cuh"

#include <cooperative_groups.h>

namespace aa {

namespace {

constexpr int kCsdThreads = 256;

struct CsdTables {
    SubPlanDev sp{};        // n_fft/2-point complex plan
    float2* d_tw = nullptr;
    float2* d_twN = nullptr;  // exp(-2 pi i k / n_fft), k in [0, n_fft/2]
    float* d_hann = nullptr;  // np.hanning(n_fft) (waterfall.py:212)
    float2* d_twc = nullptr;  // compact per-stage twiddles of the compile-time radix chain (TMA kernel), or nullptr
};

// radix chains of the TMA kernel's n_fft/2-point transforms (must match csd_tma_dispatch below)
inline bool tma_chain(int M, std::vector<int>* r) {
    switch (M) {
        case 512: *r = {8, 8, 8}; return true;
        case 1024: *r = {4, 4, 8, 8}; return true;
        case 2048: *r = {4, 8, 8, 8}; return true;
        case 4096: *r = {8, 8, 8, 8}; return true;
        default: return false;
    }
}

std::mutex g_mu;
std::map<std::pair<int, int>, CsdTables> g_tables;  // (device, n_fft) -> immutable tables

int get_tables(int n_fft, CsdTables* out) {
    int dev = 0;
    AA_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_mu);
    auto key = std::make_pair(dev, n_fft);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) {
        *out = it->second;
        return AA_OK;
    }
    const int M = n_fft / 2;
    HostSubPlan hp;
    if (!host_subplan(M, &hp)) {
        set_error("n_fft %d unsupported", n_fft);
        return AA_ERR_UNSUPPORTED;
    }
    CsdTables t;
    t.sp.L = hp.L;
    t.sp.nst = hp.nst;
    memcpy(t.sp.radix, hp.radix, sizeof(hp.radix));
    std::vector<float2> twn((size_t)M + 1);
    for (int k = 0; k <= M; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)n_fft;
        twn[(size_t)k] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<float> hann((size_t)n_fft);
    for (int k = 0; k < n_fft; ++k)  // numpy.hanning: 0.5 - 0.5 cos(2 pi k / (n - 1)); n == 1 -> [1]
        hann[(size_t)k] = (n_fft == 1) ? 1.0f : (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)k / (double)(n_fft - 1)));
    AA_CUDA_TRY(cudaMalloc(&t.d_tw, sizeof(float2) * (size_t)M));
    AA_CUDA_TRY(cudaMalloc(&t.d_twN, sizeof(float2) * ((size_t)M + 1)));
    AA_CUDA_TRY(cudaMalloc(&t.d_hann, sizeof(float) * (size_t)n_fft));
    AA_CUDA_TRY(cudaMemcpy(t.d_tw, hp.tw.data(), sizeof(float2) * (size_t)M, cudaMemcpyHostToDevice));
    AA_CUDA_TRY(cudaMemcpy(t.d_twN, twn.data(), sizeof(float2) * ((size_t)M + 1), cudaMemcpyHostToDevice));
    AA_CUDA_TRY(cudaMemcpy(t.d_hann, hann.data(), sizeof(float) * (size_t)n_fft, cudaMemcpyHostToDevice));
    std::vector<int> chain;
    if (tma_chain(M, &chain)) {  // same layout as spec_twiddles() builds in shared memory (fft_spec.cuh)
        std::vector<float2> twc((size_t)M, make_float2(1.f, 0.f));
        int Ns = 1;
        for (int R : chain) {
            const int tmul = M / (Ns * R);
            for (int m = 0; m < Ns * (R - 1); ++m) {
                const int r1 = m / Ns, k = m % Ns;
                twc[(size_t)compact_twiddle_offset(Ns) + m] = hp.tw[(size_t)((r1 + 1) * k * tmul)];
            }
            Ns *= R;
        }
        AA_CUDA_TRY(cudaMalloc(&t.d_twc, sizeof(float2) * (size_t)M));
        AA_CUDA_TRY(cudaMemcpy(t.d_twc, twc.data(), sizeof(float2) * (size_t)M, cudaMemcpyHostToDevice));
    }
    t.sp.tw = t.d_tw;
    g_tables[key] = t;
    *out = t;
    return AA_OK;
}

struct CsdArgs {
    const float* x;
    long long row_stride;
    const int* lengths;
    int uniform_len;
    const int* start_idx;
    const int* frame_idx;
    const int* n_slices;
    int max_slices;
    int n_fft, hop, use_hann;
    int bin_lo, n_bins;
    int db_reference;
    float floor_lin_sq;  // (10^(floor_db/20))^2
    float dyn;
    float* out;
    SubPlanDev sp;
    const float2* twN;
    const float* hann;
    const float2* twc;
};

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    return m;
}

__global__ void __launch_bounds__(kCsdThreads) csd_kernel(CsdArgs a) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int row = blockIdx.x;
    const int M = a.n_fft >> 1;
    float2* bufA = reinterpret_cast<float2*>(dyn_smem);
    float2* bufB = bufA + M;
    float* raw = reinterpret_cast<float*>(bufB + M);  // n_fft + 8 floats (aligned superset of the frame)
    __shared__ float red[kCsdThreads / 32];
    __shared__ float slice_ref[AA_MAX_SLICES];

    const int len = a.lengths ? a.lengths[row] : a.uniform_len;
    const float* __restrict__ x = a.x + (long long)row * a.row_stride;
    const int start = a.start_idx[row];
    const int ns = min(a.n_slices[row], a.max_slices);
    float* __restrict__ out_row = a.out + (long long)row * a.max_slices * a.n_bins;
    const int len4 = (len + 3) & ~3;  // row_stride >= len4: the pitch padding is readable

    float run_max = -INFINITY;
    for (int s = 0; s < ns; ++s) {
        const int f0 = start + a.frame_idx[(long long)row * a.max_slices + s] * a.hop;
        const int a0 = f0 & ~3;
        const int shift = f0 - a0;
        const int nvec = (a.n_fft + shift + 3) >> 2;  // float4s covering [a0, f0 + n_fft)
        for (int q = threadIdx.x; q < nvec; q += blockDim.x) {
            const int g = a0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g >= 0 && g + 4 <= len4) v = ldg_f4(x + g);
            reinterpret_cast<float4*>(raw)[q] = v;
        }
        __syncthreads();
        // window (waterfall.py:223) and pack z[m] = (w x)[2m] + i (w x)[2m+1]
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            float x0 = raw[shift + 2 * m], x1 = raw[shift + 2 * m + 1];
            if (f0 + 2 * m >= len) x0 = 0.f;
            if (f0 + 2 * m + 1 >= len) x1 = 0.f;
            if (a.use_hann) {
                x0 *= __ldg(a.hann + 2 * m);
                x1 *= __ldg(a.hann + 2 * m + 1);
            }
            bufA[m] = make_float2(x0, x1);
        }
        __syncthreads();
        float2* pa = bufA;
        float2* pb = bufB;
        int Ns = 1;
        for (int st = 0; st < a.sp.nst; ++st) {
            const int R = a.sp.radix[st];
            const int nb = M / R;
            for (int j = threadIdx.x; j < nb; j += blockDim.x) stockham_butterfly(R, j, M, Ns, pa, pb, a.sp.tw);
            __syncthreads();
            float2* t = pa;
            pa = pb;
            pb = t;
            Ns *= R;
        }
        // untangle -> |X|^2 -> dB, cropped to [bin_lo, bin_lo + n_bins)   (waterfall.py:224-227, :304-309)
        float smax = -INFINITY;
        float* __restrict__ o = out_row + (long long)s * a.n_bins;
        for (int b = threadIdx.x; b < a.n_bins; b += blockDim.x) {
            const int k = a.bin_lo + b;
            float2 zk = pa[k == M ? 0 : k];
            float2 zm = pa[(M - k) == M ? 0 : (M - k)];
            float2 X = untangle_half_spectrum(zk, zm, __ldg(a.twN + k));
            float p = fmaxf(X.x * X.x + X.y * X.y, a.floor_lin_sq);
            float db = 10.0f * log10f(p);
            o[b] = db;
            smax = fmaxf(smax, db);
        }
        if (a.db_reference == AA_DBREF_SLICE_MAX) {
            float m = block_max(smax, red);
            if (threadIdx.x == 0) slice_ref[s] = m;
        } else {
            run_max = fmaxf(run_max, smax);
        }
        __syncthreads();
    }
    if (a.db_reference == AA_DBREF_NONE) return;
    float gref = 0.f;
    if (a.db_reference == AA_DBREF_GLOBAL_MAX) gref = block_max(run_max, red);  // waterfall.py:335
    __syncthreads();
    // the CTA re-reads its own stores: ordered by the barrier above
    for (int s = 0; s < ns; ++s) {
        const float ref = (a.db_reference == AA_DBREF_SLICE_MAX) ? slice_ref[s] : gref;
        float* __restrict__ o = out_row + (long long)s * a.n_bins;
        for (int b = threadIdx.x; b < a.n_bins; b += blockDim.x) {
            float r = o[b] - ref;
            o[b] = fminf(fmaxf(r, -a.dyn), 0.0f);  // waterfall.py:338-339
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA kernel: a cluster of CTAs per row, each CTA owning every C-th selected frame.
//   * frames are staged by 1-D bulk copies (cp.async.bulk -> UBLKCP) of the 16-byte aligned superset of the
//     frame into a two-deep shared-memory ring, completion on an mbarrier: the next frame lands while the
//     current one is transformed;
//   * compile-time radix chain with a padded shared-memory layout (fft_spec.cuh);
//   * the CTA keeps its slices' dB values in shared memory, the cluster agrees on the reference level through
//     distributed shared memory, and every output value is written to HBM exactly once, already normalised and
//     clipped (waterfall.py:328-339): traffic = frames in + slices out, the algorithmic minimum.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

constexpr int kTmaFramesPerCta = 3;
constexpr int kTmaPad = 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int M>
struct TmaCfg {
    static constexpr int kNfft = 2 * M;
    static constexpr int kRawFloats = kNfft + 8;                 // aligned superset of a frame
    static constexpr int kFftElems = M + (M >> kTmaPad);         // padded complex buffer
    static constexpr size_t kRawBytes = sizeof(float) * (size_t)kRawFloats;   // ONE staging buffer: 72 KB per CTA, three CTAs per SM
    static constexpr size_t kFftBytes = sizeof(float2) * 2 * (size_t)kFftElems;
    static size_t smem_bytes(int n_bins) {
        return kRawBytes + kFftBytes + sizeof(float) * (size_t)kTmaFramesPerCta * (size_t)((n_bins + 3) & ~3);
    }
};

template <int M, int... Rs>
__global__ void __launch_bounds__(kCsdThreads) csd_tma_kernel(CsdArgs a) {
    using Cfg = TmaCfg<M>;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int row = blockIdx.x / C;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float* raw = reinterpret_cast<float*>(dyn_smem);                       // [kRawFloats]
    float2* fa = reinterpret_cast<float2*>(dyn_smem + Cfg::kRawBytes);   // [kFftElems]
    float2* fb = fa + Cfg::kFftElems;
    float* outs = reinterpret_cast<float*>(dyn_smem + Cfg::kRawBytes + Cfg::kFftBytes);
    const int out_pitch = (a.n_bins + 3) & ~3;
    __shared__ unsigned long long bars[1];
    __shared__ float red[kCsdThreads / 32];
    __shared__ float cta_max;
    __shared__ float slice_ref[kTmaFramesPerCta];

    const int len = a.lengths ? a.lengths[row] : a.uniform_len;
    const float* __restrict__ x = a.x + (long long)row * a.row_stride;
    const int start = a.start_idx[row];
    const int ns = min(a.n_slices[row], a.max_slices);
    const int len4 = (len + 3) & ~3;
    const int n_mine = (ns > rank) ? (ns - rank + C - 1) / C : 0;   // slices rank, rank + C, ...

    auto frame_start = [&](int i) { return start + a.frame_idx[(long long)row * a.max_slices + rank + i * C] * a.hop; };
    // thread 0 stages frame i into the staging buffer (bulk copy when the aligned superset lies inside the row); the
    // copy of frame i + 1 is issued as soon as frame i has been windowed into the FFT buffer, so it lands during the
    // transform of frame i (a second staging buffer would cost the third resident CTA per SM)
    auto stage_frame = [&](int i) {
        const int f0 = frame_start(i);
        const int a0 = f0 & ~3;
        const int nvec = (Cfg::kNfft + (f0 - a0) + 3) >> 2;
        if (f0 >= 0 && a0 + 4 * nvec <= len4) {
            mbar_expect_tx(&bars[0], 16u * (unsigned)nvec);
            bulk_g2s(raw, x + a0, 16u * (unsigned)nvec, &bars[0]);
        }
    };
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (n_mine > 0) stage_frame(0);
    }
    __syncthreads();

    float run_max = -INFINITY;
    int bulk_phase = 0;   // completed bulk copies on bars[0]: frames that took the guarded-load path do not use it
    for (int i = 0; i < n_mine; ++i) {
        const int f0 = frame_start(i);
        const int a0 = f0 & ~3;
        const int shift = f0 - a0;
        const int nvec = (Cfg::kNfft + shift + 3) >> 2;
        float* rb = raw;
        if (f0 >= 0 && a0 + 4 * nvec <= len4) {
            mbar_wait(&bars[0], (unsigned)(bulk_phase & 1));
            ++bulk_phase;
        } else {  // frame sticks out of the row (not produced by the host-side frame selection): guarded loads
            for (int q = threadIdx.x; q < nvec; q += kCsdThreads) {
                const int g = a0 + 4 * q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g >= 0 && g + 4 <= len4) v = ldg_f4(x + g);
                reinterpret_cast<float4*>(rb)[q] = v;
            }
            __syncthreads();
        }
        // window (waterfall.py:223) and pack z[m] = (w x)[2m] + i (w x)[2m+1]
        for (int m = threadIdx.x; m < M; m += kCsdThreads) {
            float x0 = rb[shift + 2 * m], x1 = rb[shift + 2 * m + 1];
            if (f0 + 2 * m >= len) x0 = 0.f;
            if (f0 + 2 * m + 1 >= len) x1 = 0.f;
            if (a.use_hann) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(a.hann) + m);
                x0 *= w.x;
                x1 *= w.y;
            }
            fa[pad_index<kTmaPad>(m)] = make_float2(x0, x1);
        }
        __syncthreads();  // staging buffer consumed
        if (threadIdx.x == 0 && i + 1 < n_mine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async overwrite
            stage_frame(i + 1);
        }
        float2* pa = fa;
        float2* pb = fb;
        spec_stages<kCsdThreads, kTmaPad, M, 1, Cfg::kFftElems, 1, Rs...>(pa, pb, a.twc, a.sp.tw);
        // untangle -> |X|^2 -> dB, cropped to [bin_lo, bin_lo + n_bins)   (waterfall.py:224-227, :304-309)
        float smax = -INFINITY;
        float* __restrict__ o = outs + i * out_pitch;
        for (int b = threadIdx.x; b < a.n_bins; b += kCsdThreads) {
            const int k = a.bin_lo + b;
            const float2 zk = pa[pad_index<kTmaPad>(k == M ? 0 : k)];
            const float2 zm = pa[pad_index<kTmaPad>((M - k) == M ? 0 : (M - k))];
            const float2 X = untangle_half_spectrum(zk, zm, __ldg(a.twN + k));
            const float p = fmaxf(X.x * X.x + X.y * X.y, a.floor_lin_sq);
            const float db = 3.010299956639812f * log2f(p);  // 10 log10 |X|^2
            o[b] = db;
            smax = fmaxf(smax, db);
        }
        if (a.db_reference == AA_DBREF_SLICE_MAX) {
            const float m = block_max(smax, red);
            if (threadIdx.x == 0) slice_ref[i] = m;
        } else {
            run_max = fmaxf(run_max, smax);
        }
        __syncthreads();  // fa / fb free for the next frame
    }
    float gref = 0.f;
    if (a.db_reference == AA_DBREF_GLOBAL_MAX) {  // waterfall.py:335: one level for all slices of the row
        const float m = block_max(run_max, red);
        if (threadIdx.x == 0) cta_max = m;
        cluster.sync();
        gref = -INFINITY;
        for (int r = 0; r < C; ++r) gref = fmaxf(gref, *cluster.map_shared_rank(&cta_max, r));
    }
    __syncthreads();
    float* __restrict__ out_row = a.out + (long long)row * a.max_slices * a.n_bins;
    for (int i = 0; i < n_mine; ++i) {
        const float ref = (a.db_reference == AA_DBREF_SLICE_MAX) ? slice_ref[i] : gref;
        const float* __restrict__ src = outs + i * out_pitch;
        float* __restrict__ o = out_row + (long long)(rank + i * C) * a.n_bins;
        for (int b = threadIdx.x; b < a.n_bins; b += kCsdThreads) {
            float v = src[b];
            if (a.db_reference != AA_DBREF_NONE) v = fminf(fmaxf(v - ref, -a.dyn), 0.0f);  // waterfall.py:338-339
            __stcs(o + b, v);
        }
    }
    if (a.db_reference == AA_DBREF_GLOBAL_MAX) cluster.sync();  // peers may still be reading cta_max
}

template <int M, int... Rs>
int launch_csd_tma(const CsdArgs& a, int n_rows, int cluster, cudaStream_t stream) {
    auto kern = csd_tma_kernel<M, Rs...>;
    const size_t smem = TmaCfg<M>::smem_bytes(a.n_bins);
    AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)n_rows * cluster));
    cfg.blockDim = dim3(kCsdThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return AA_OK;
}

// returns AA_ERR_UNSUPPORTED when the TMA kernel does not cover the request (the caller falls back)
int csd_tma_dispatch(const CsdArgs& a, int n_rows, cudaStream_t stream) {
    if (a.twc == nullptr || getenv("AA_CSD_GENERIC") != nullptr) return AA_ERR_UNSUPPORTED;
    const int cluster = (a.max_slices + kTmaFramesPerCta - 1) / kTmaFramesPerCta;
    if (cluster > 8) return AA_ERR_UNSUPPORTED;
    switch (a.n_fft) {
        case 1024: return launch_csd_tma<512, 8, 8, 8>(a, n_rows, cluster, stream);
        case 2048: return launch_csd_tma<1024, 4, 4, 8, 8>(a, n_rows, cluster, stream);
        case 4096: return launch_csd_tma<2048, 4, 8, 8, 8>(a, n_rows, cluster, stream);
        case 8192: return launch_csd_tma<4096, 8, 8, 8, 8>(a, n_rows, cluster, stream);
        default: return AA_ERR_UNSUPPORTED;
    }
}

// ---------------------------------------------------------------------------------------------
// Dense STFT magnitude in dB (analyse/spectrogram.py:107-160, the same loop as waterfall.py:188-230 but every
// frame is kept).  A CTA takes a run of consecutive frames of one row: the frames overlap (hop < n_fft), so the
// whole run is staged by ONE bulk copy and each sample is read from HBM once; every frame is windowed,
// transformed with the compile-time radix chain and written as n_fft/2 + 1 float32 dB values, frame-major.
// ---------------------------------------------------------------------------------------------
struct StftArgs {
    const float* x;
    long long row_stride;
    const int* lengths;
    int uniform_len;
    const int* start_idx;
    const int* span;        // analysed samples per row (after trim / ignore / duration)
    int n_fft, hop, use_hann;
    int frames_per_cta, frames_max;
    float floor_lin_sq;
    float* out;             // [row][frames_max][n_fft/2 + 1]
    SubPlanDev sp;
    const float2* twN;
    const float* hann;
    const float2* twc;
};

template <int M, int... Rs>
__global__ void __launch_bounds__(kCsdThreads) stft_db_kernel(StftArgs a) {
    using Cfg = TmaCfg<M>;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float2* fa = reinterpret_cast<float2*>(dyn_smem);
    float2* fb = fa + Cfg::kFftElems;
    float* raw = reinterpret_cast<float*>(dyn_smem + Cfg::kFftBytes);   // (frames_per_cta - 1) * hop + n_fft + 8
    __shared__ unsigned long long bar;

    const int row = blockIdx.y;
    const int len = a.lengths ? a.lengths[row] : a.uniform_len;
    const int start = a.start_idx[row];
    const int span = min(a.span[row], len - start);
    const int n_frames = (span >= a.n_fft) ? min(1 + (span - a.n_fft) / a.hop, a.frames_max) : 0;
    const int j0 = blockIdx.x * a.frames_per_cta;
    if (j0 >= n_frames) return;
    const int nf = min(a.frames_per_cta, n_frames - j0);
    const float* __restrict__ x = a.x + (long long)row * a.row_stride;
    const int len4 = (len + 3) & ~3;
    const int f0 = start + j0 * a.hop;
    const int a0 = f0 & ~3;
    const int shift = f0 - a0;
    const int nvec = ((nf - 1) * a.hop + a.n_fft + shift + 3) >> 2;
    const bool bulk = f0 >= 0 && a0 + 4 * nvec <= len4;
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar, 16u * (unsigned)nvec);
            bulk_g2s(raw, x + a0, 16u * (unsigned)nvec, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    } else {
        for (int q = threadIdx.x; q < nvec; q += kCsdThreads) {
            const int g = a0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g >= 0 && g + 4 <= len4) v = ldg_f4(x + g);
            reinterpret_cast<float4*>(raw)[q] = v;
        }
        __syncthreads();
    }
    float* __restrict__ out_row = a.out + ((long long)row * a.frames_max + j0) * (M + 1);
    for (int r = 0; r < nf; ++r) {
        const float* rb = raw + shift + r * a.hop;
        for (int m = threadIdx.x; m < M; m += kCsdThreads) {
            float x0 = rb[2 * m], x1 = rb[2 * m + 1];
            if (a.use_hann) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(a.hann) + m);
                x0 *= w.x;
                x1 *= w.y;
            }
            fa[pad_index<kTmaPad>(m)] = make_float2(x0, x1);
        }
        __syncthreads();
        float2* pa = fa;
        float2* pb = fb;
        spec_stages<kCsdThreads, kTmaPad, M, 1, Cfg::kFftElems, 1, Rs...>(pa, pb, a.twc, a.sp.tw);
        float* __restrict__ o = out_row + (long long)r * (M + 1);
        for (int k = threadIdx.x; k <= M; k += kCsdThreads) {
            const float2 zk = pa[pad_index<kTmaPad>(k == M ? 0 : k)];
            const float2 zm = pa[pad_index<kTmaPad>((M - k) == M ? 0 : (M - k))];
            const float2 X = untangle_half_spectrum(zk, zm, __ldg(a.twN + k));
            const float p = fmaxf(X.x * X.x + X.y * X.y, a.floor_lin_sq);
            __stcs(o + k, 3.010299956639812f * log2f(p));   // 20 log10 |X|  (spectrogram.py:154-155)
        }
        __syncthreads();
    }
}

template <int M, int... Rs>
int launch_stft(const StftArgs& a, int n_rows, cudaStream_t stream) {
    auto kern = stft_db_kernel<M, Rs...>;
    const size_t smem = TmaCfg<M>::kFftBytes + sizeof(float) * ((size_t)(a.frames_per_cta - 1) * a.hop + a.n_fft + 8);
    AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((a.frames_max + a.frames_per_cta - 1) / a.frames_per_cta), (unsigned)n_rows);
    kern<<<grid, kCsdThreads, smem, stream>>>(a);
    return launch_check("stft_db_kernel");
}

__global__ void pcm16_kernel(const short* __restrict__ pcm, long long pcm_stride, int n_frames, int n_channels,
                             int mode, float* __restrict__ out, long long out_stride) {
    const int file = blockIdx.y;
    const short* __restrict__ p = pcm + (long long)file * pcm_stride;
    const float k = 1.0f / 32768.0f;  // io.py:59; |pcm/32768| <= 1 so the clip at io.py:109 never acts
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_frames; i += gridDim.x * blockDim.x) {
        if (n_channels == 1) {
            out[(long long)file * out_stride + i] = (float)p[i] * k;
        } else if (mode == 1) {
            float l = (float)p[2 * i] * k, r = (float)p[2 * i + 1] * k;
            out[(long long)file * out_stride + i] = 0.5f * __fadd_rn(l, r);  // io.py:90
        } else {
            for (int c = 0; c < n_channels; ++c)
                out[((long long)file * n_channels + c) * out_stride + i] = (float)p[(long long)i * n_channels + c] * k;
        }
    }
}

// slices a row does not have (n_slices < max_slices) read as zero: the caller's buffer need not be cleared
__global__ void csd_clear_unused_kernel(float* __restrict__ out, const int* __restrict__ n_slices, int max_slices, int n_bins) {
    const int row = blockIdx.x;
    const int ns = min(max(n_slices[row], 0), max_slices);
    float* __restrict__ p = out + ((long long)row * max_slices + ns) * n_bins;
    const long long total = (long long)(max_slices - ns) * n_bins;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) p[i] = 0.0f;
}

}  // namespace

int launch_csd(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int n_rows,
               const int* d_start_idx, const int* d_frame_idx, const int* d_n_slices, int max_slices,
               const aa_csd_params* p, float* d_out, cudaStream_t stream) {
    AA_REQUIRE(d_x && d_start_idx && d_frame_idx && d_n_slices && p && d_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0, AA_ERR_INVALID_ARGUMENT,
               "rows must be 16-byte aligned (row_stride % 4 == 0)");
    AA_REQUIRE(p->n_fft >= 64 && p->n_fft <= 8192 && (p->n_fft & (p->n_fft - 1)) == 0, AA_ERR_UNSUPPORTED,
               "n_fft must be a power of two in [64, 8192]");
    AA_REQUIRE(p->hop_length > 0, AA_ERR_INVALID_ARGUMENT, "hop_length must be positive");
    AA_REQUIRE(max_slices >= 1 && max_slices <= AA_MAX_SLICES, AA_ERR_INVALID_ARGUMENT, "max_slices out of range");
    AA_REQUIRE(p->bin_lo >= 0 && p->n_bins >= 1 && p->bin_lo + p->n_bins <= p->n_fft / 2 + 1, AA_ERR_INVALID_ARGUMENT,
               "bin range outside the rfft");
    if (n_rows <= 0) return AA_OK;
    CsdTables t;
    int rc = get_tables(p->n_fft, &t);
    if (rc != AA_OK) return rc;
    CsdArgs a{};
    a.x = d_x;
    a.row_stride = row_stride;
    a.lengths = d_lengths;
    a.uniform_len = uniform_len;
    a.start_idx = d_start_idx;
    a.frame_idx = d_frame_idx;
    a.n_slices = d_n_slices;
    a.max_slices = max_slices;
    a.n_fft = p->n_fft;
    a.hop = p->hop_length;
    a.use_hann = p->use_hann_window;
    a.bin_lo = p->bin_lo;
    a.n_bins = p->n_bins;
    a.db_reference = p->db_reference;
    const double fl = pow(10.0, p->floor_db / 20.0);  // waterfall.py:218
    a.floor_lin_sq = (float)(fl * fl);
    a.dyn = (float)p->dynamic_range_db;
    a.out = d_out;
    a.sp = t.sp;
    a.twN = t.d_twN;
    a.hann = t.d_hann;
    a.twc = t.d_twc;
    csd_clear_unused_kernel<<<(unsigned)n_rows, 128, 0, stream>>>(d_out, d_n_slices, max_slices, p->n_bins);
    rc = launch_check("csd_clear_unused_kernel");
    if (rc != AA_OK) return rc;
    {
        const int rc_tma = csd_tma_dispatch(a, n_rows, stream);
        if (rc_tma != AA_ERR_UNSUPPORTED) return rc_tma;
    }
    const size_t smem = sizeof(float2) * (size_t)p->n_fft + sizeof(float) * ((size_t)p->n_fft + 8);
    AA_CUDA_TRY(cudaFuncSetAttribute(csd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    csd_kernel<<<n_rows, kCsdThreads, smem, stream>>>(a);
    return launch_check("csd_kernel");
}

int launch_stft_db(const float* d_x, long long row_stride, const int* d_lengths, int uniform_len, int n_rows,
                   const int* d_start_idx, const int* d_span, int n_fft, int hop, int use_hann, double floor_db,
                   int frames_max, float* d_out, cudaStream_t stream) {
    AA_REQUIRE(d_x && d_start_idx && d_span && d_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0, AA_ERR_INVALID_ARGUMENT,
               "rows must be 16-byte aligned (row_stride % 4 == 0)");
    AA_REQUIRE(n_fft == 1024 || n_fft == 2048 || n_fft == 4096 || n_fft == 8192, AA_ERR_UNSUPPORTED,
               "dense STFT kernel: n_fft must be 1024, 2048, 4096 or 8192");
    AA_REQUIRE(hop > 0, AA_ERR_INVALID_ARGUMENT, "hop_length must be positive");
    AA_REQUIRE(n_rows <= 65535, AA_ERR_INVALID_ARGUMENT, "at most 65535 rows per call");
    if (n_rows <= 0 || frames_max <= 0) return AA_OK;
    CsdTables t;
    int rc = get_tables(n_fft, &t);
    if (rc != AA_OK) return rc;
    StftArgs a{};
    a.x = d_x;
    a.row_stride = row_stride;
    a.lengths = d_lengths;
    a.uniform_len = uniform_len;
    a.start_idx = d_start_idx;
    a.span = d_span;
    a.n_fft = n_fft;
    a.hop = hop;
    a.use_hann = use_hann;
    a.frames_max = frames_max;
    const double fl = pow(10.0, floor_db / 20.0);  // spectrogram.py:154
    a.floor_lin_sq = (float)(fl * fl);
    a.out = d_out;
    a.sp = t.sp;
    a.twN = t.d_twN;
    a.hann = t.d_hann;
    a.twc = t.d_twc;
    // as many overlapping frames per CTA as fit ~64 KB of staged samples (at most 16)
    long long fpc = 1 + ((long long)(64 * 1024 / 4) - n_fft - 8) / hop;
    if (fpc < 1) fpc = 1;
    if (fpc > 16) fpc = 16;
    if (fpc > frames_max) fpc = frames_max;
    a.frames_per_cta = (int)fpc;
    switch (n_fft) {
        case 1024: return launch_stft<512, 8, 8, 8>(a, n_rows, stream);
        case 2048: return launch_stft<1024, 4, 4, 8, 8>(a, n_rows, stream);
        case 4096: return launch_stft<2048, 4, 8, 8, 8>(a, n_rows, stream);
        default: return launch_stft<4096, 8, 8, 8, 8>(a, n_rows, stream);
    }
}

int launch_pcm16(const int16_t* d_pcm, long long pcm_stride, int n_files, int n_frames, int n_channels, int mode,
                 float* d_out, long long out_stride, cudaStream_t stream) {
    AA_REQUIRE(d_pcm && d_out, AA_ERR_INVALID_ARGUMENT, "null argument");
    AA_REQUIRE(n_channels == 1 || n_channels == 2, AA_ERR_INVALID_ARGUMENT, "1 or 2 channels (io.py:172-178)");
    if (n_files <= 0 || n_frames <= 0) return AA_OK;
    int bx = (n_frames + 255) / 256;
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)bx, (unsigned)n_files);
    pcm16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const short*>(d_pcm), pcm_stride, n_frames, n_channels, mode,
                                           d_out, out_stride);
    return launch_check("pcm16_kernel");
}

}  // namespace aa
