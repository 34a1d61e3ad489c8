// Diagnostic micro-benchmark: how long does the tensor pipe take per tcgen05.mma of the shapes the decoder issues, for a single
// CTA (cta_group::1, M = 128) and for a CTA pair (cta_group::2, M = 256), with A from shared memory (SS) or TMEM (TS)?
// One thread issues `n` back-to-back MMAs on garbage operands (timing only) and commits; out[0] = cycles until the last issue
// returned, out[1] = cycles until the commit's mbarrier flipped.
#include "nb_internal.h"
#include "nb_tc_ptx.cuh"

namespace nb {
namespace mmabench {

constexpr int SMEM = 64 * 1024 + 64;

// (two instantiations: a kernel that contains cta_group::2 instructions can only be launched as a cluster of 2)
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int variant, int n, int N, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64 * 1024 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr bool pair = PAIR;
    const bool ts = variant & 1;
    uint32_t rank = 0u;
    if constexpr (PAIR) rank = tc::cluster_ctarank();
    for (int i = tid; i < 16 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;     // halves 1.0
    if (warp == 0) { if constexpr (PAIR) tc::tmem_alloc_pair<512>(tmem_slot); else tc::tmem_alloc<512>(tmem_slot); }
    if (tid == 32) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if constexpr (PAIR) tc::cluster_sync_all();
    const uint32_t tmem = *tmem_slot;
    if (tid == 64 && rank == 0) {
        const uint32_t idesc = tc::make_idesc_f16(pair ? 256 : 128, N);
        const int nb = pair ? N / 2 : N;                                   // B rows in this CTA's shared memory
        const uint64_t ad = tc::make_smem_desc(tc::smem_u32(smem), 128 * 16, 128);
        const uint64_t bd = tc::make_smem_desc(tc::smem_u32(smem) + 8192, nb * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < n; ++i) {
            if constexpr (PAIR) { if (ts) tc::mma_ts_pair(tmem, tmem + 256 + (i & 7) * 8, bd, idesc, i > 0); else tc::mma_ss_pair(tmem, ad, bd, idesc, i > 0); }
            else { if (ts) tc::mma_ts(tmem, tmem + 256 + (i & 7) * 8, bd, idesc, i > 0); else tc::mma_ss(tmem, ad, bd, idesc, i > 0); }
        }
        const long long t1 = clock64();
        if constexpr (PAIR) tc::mma_commit_pair(bar, 0b01); else tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    __syncthreads();
    if constexpr (PAIR) tc::cluster_sync_all();
    if (warp == 0) { if constexpr (PAIR) tc::tmem_dealloc_pair<512>(tmem); else tc::tmem_dealloc<512>(tmem); }
}

}  // namespace mmabench
}  // namespace nb

// variant bit 0: A from TMEM (TS) instead of shared memory (SS); bit 1: CTA pair (cta_group::2, M = 256).  out: device i64[2]
extern "C" int nb_debug_mma_rate(int variant, int n_mma, int N, long long* out, void* stream) {
    using namespace nb;
    if (!out || n_mma <= 0 || N < 16 || N > 256 || N % 16) { set_error("nb_debug_mma_rate: bad argument"); return NB_ERR_BAD_ARG; }
    const bool pair = variant & 2;
    cudaError_t e = pair ? cudaFuncSetAttribute(mmabench::mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mmabench::SMEM)
                         : cudaFuncSetAttribute(mmabench::mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mmabench::SMEM);
    if (e == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pair ? 2 : 1);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = mmabench::SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pair ? 1 : 0;
        e = pair ? cudaLaunchKernelEx(&cfg, mmabench::mma_rate_kernel<true>, variant, n_mma, N, out)
                 : cudaLaunchKernelEx(&cfg, mmabench::mma_rate_kernel<false>, variant, n_mma, N, out);
    }
    if (e != cudaSuccess) { set_error("nb_debug_mma_rate: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
    return NB_OK;
}
