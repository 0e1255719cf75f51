// Fused Schroeder EDC + crossings + fits with the row held ON CHIP between the two phases (sm_100a).
//
// edc_kernel (edc.cu) streams a row twice: phase 1 finds the peak and the tile energies, phase 2 needs E0 = E[start]
// before it can write the first dB value, so it reads the row again - 12 B per sample of HBM traffic for 8 B of
// algorithmic traffic once more rows are in flight than fit L2.  Here a thread-block cluster owns the row and every
// CTA keeps ITS span of it in shared memory after phase 1 (288,000 samples = 1.15 MB = 8 CTAs x 144 KB, or 16 x 72 KB);
// E0, the crossing tiles and phase 2 are served from shared memory.  HBM sees x once and the curve once.
// The body is edc_body.cuh compiled with AA_EDC_STAGE: identical arithmetic, bit-identical results.
#define AA_EDC_VARIANT staged
#define AA_EDC_STAGE 1
#ifndef AA_EDC_STAGED_THREADS
#define AA_EDC_STAGED_THREADS 512
#endif
#undef AA_EDC_THREADS
#define AA_EDC_THREADS AA_EDC_STAGED_THREADS

#include <cooperative_groups.h>
#include <stdlib.h>

#include "aa_common.cuh"
#include "edc_body.cuh"

namespace cg = cooperative_groups;

namespace aa {

namespace {

using namespace edc;
using namespace edc::staged;

template <bool kWriteDb>
__global__ void __launch_bounds__(kThreads, 1) edc_staged_kernel(EdcArgs a) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ CtaShared sh;
    const int T = (int)cg::this_cluster().num_blocks();
    const int row = blockIdx.x / T;
    edc_row<kWriteDb, false>(a, row, a.x + (long long)row * a.row_stride, dyn_smem, sh);
}

constexpr size_t kSmemLimit = 200 * 1024;

size_t stage_bytes(int tpw) {
    const size_t tpc = (size_t)tpw * kWarps;
    return ((tpc * 12 + 15) / 16) * 16 + tpc * 32 * sizeof(float4);
}

}  // namespace

int launch_edc_staged(const edc::EdcArgs& a_in, int n_rows, int max_len, bool write_db, cudaStream_t stream,
                      bool* launched) {
    *launched = false;
    // Opt-in (AA_EDC_STAGED = 8 or 16: CTAs per row).  Measured on B200 (profiles/r02_edc_staged_ncu.txt):
    // the row does cross HBM once (2.18 MB of DRAM traffic per 288,000-sample row against 3.2 MB for edc_kernel),
    // but with 144 KB of shared memory per CTA only ONE row is resident per 8 SMs, its phases run one after the other
    // (loads, resolve, dB pass) and nothing overlaps them: 19.6 ms per 8192 rows against 6.0 ms for edc_kernel, whose
    // four resident rows per SM sit in different phases.  Kept for the persistent, TMA-pipelined version.
    int T = 8;
    const char* env = getenv("AA_EDC_STAGED");
    if (!env) return AA_OK;
    {
        const int v = atoi(env);
        if (v != 8 && v != 16) return AA_OK;
        T = v;
    }
    const int min_rows = 1;
    if (!write_db || n_rows < min_rows || max_len < 4096) return AA_OK;
    const int nt_max = (max_len + kTile - 1) / kTile;
    int tpw = (nt_max + T * kWarps - 1) / (T * kWarps);
    if (tpw < 1) tpw = 1;
    if (stage_bytes(tpw) > kSmemLimit) {
        if (T == 8) {   // longer rows: 16 CTAs per row (non-portable cluster size)
            T = 16;
            tpw = (nt_max + T * kWarps - 1) / (T * kWarps);
        }
        if (stage_bytes(tpw) > kSmemLimit) return AA_OK;
    }
    EdcArgs a = a_in;
    a.tiles_per_warp = tpw;
    auto kern = edc_staged_kernel<true>;
    static bool configured = false;
    if (!configured) {
        AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)n_rows * T));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = stage_bytes(tpw);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)T;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (T > 8) {   // is a cluster of 16 schedulable on this part at all?
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
            cudaGetLastError();
            return AA_OK;
        }
    }
    AA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    *launched = true;
    return AA_OK;
}

}  // namespace aa
