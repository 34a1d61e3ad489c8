// Diagnostic: the CTA-PAIR (cta_group::2) building blocks of the tensor-core decoder in isolation, on one 256-row tile
// (128 rows per CTA of a 2-cluster):
//   * tcgen05.alloc / dealloc with cta_group::2 in both CTAs,
//   * one M = 256 MMA issued by the leader: A and the accumulator CTA-local, B split by N halves across the two CTAs' shared
//     memory (rank 0: rows [0, N/2), rank 1: rows [N/2, N)) -- SS form (A from shared memory) and TS form (A from TMEM),
//   * N = 256, N = 144 (72 rows per CTA) and N = 16 (8 rows per CTA, accumulated into the middle of a wider accumulator),
//   * multicast commit to both CTAs' mbarriers, remote (cluster-scope) mbarrier arrives from the peer's producer threads,
//   * the relay of a CTA-local bulk-copy completion to the leader.
// layer 0 (SS): D0[256 x 256] = [A0 | 1 1 0..] (K = 64 + 16) * W0^T          W0 halves: [2][128 x 80] packed K-major
// layer 1 (TS): D1[256 x 144] = relu(D0[:, :128]) (fp16, TMEM) * W1^T        W1 halves: [2][72 x 128] packed K-major
//               then D1[:, 64:80] += relu(D0[:, :128]) * [W1 half r rows 64..71]^T   (one N = 16 MMA at column 64)
#include "nb_internal.h"
#include "nb_tc_ptx.cuh"

namespace nb {
namespace probe2 {

constexpr int K0 = 64, K0P = 80, N0 = 256, K1 = 128, N1 = 144;
constexpr int OFF_A = 0;                                   // 128 x 80 halves, canonical          20480 B
constexpr int OFF_W0 = 20480;                              // (N0/2) x 80 canonical               20480 B
constexpr int OFF_W1 = 40960;                              // (N1/2) x 128 canonical              18432 B
constexpr int OFF_BAR = 59392;
enum { B_W = 0, B_WPEER, B_AREADY, B_D0, B_HREADY, B_D1, NBARS };
constexpr int OFF_TMEM = OFF_BAR + 64;
constexpr int SMEM = OFF_TMEM + 16;

__global__ void __launch_bounds__(192, 1)
tc_probe2_kernel(const __half* __restrict__ a0, const __half* __restrict__ w0h, const __half* __restrict__ w1h,
                 float* __restrict__ d0_out, float* __restrict__ d1_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem + OFF_A);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 4) tc::tmem_alloc_pair<512>(tmem_slot);
    if (tid == 160) {
        tc::mbar_init(&bars[B_W], 1);
        tc::mbar_init(&bars[B_WPEER], 1);
        tc::mbar_init(&bars[B_AREADY], 8);       // 4 warps of each CTA
        tc::mbar_init(&bars[B_D0], 1);
        tc::mbar_init(&bars[B_HREADY], 8);
        tc::mbar_init(&bars[B_D1], 1);
        tc::fence_mbar_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tc::cluster_sync_all();
    const uint32_t tmem = *tmem_slot;

    if (tid == 160) {   // every CTA loads ITS half of the weights into its own shared memory
        tc::mbar_arrive_expect_tx(&bars[B_W], (N0 / 2) * K0P * 2 + (N1 / 2) * K1 * 2);
        tc::bulk_g2s(smem + OFF_W0, w0h + (size_t)rank * (N0 / 2) * K0P, (N0 / 2) * K0P * 2, &bars[B_W]);
        tc::bulk_g2s(smem + OFF_W1, w1h + (size_t)rank * (N1 / 2) * K1, (N1 / 2) * K1 * 2, &bars[B_W]);
        if (!leader) {  // relay: the peer's half has landed -> tell the leader
            tc::mbar_wait(&bars[B_W], 0);
            tc::mbar_arrive_remote(tc::map_to_cta(&bars[B_WPEER], 0));
        }
    }
    if (tid < 128) {    // A0 rows of this CTA: row r -> canonical [k/8][r/8][r%8][8]
        const int r = tid;
        const __half* src = a0 + (size_t)(rank * 128 + r) * K0;
#pragma unroll
        for (int j = 0; j < K0P / 8; ++j) {
            uint4 v;
            if (j < K0 / 8) v = *reinterpret_cast<const uint4*>(src + j * 8);
            else if (j == K0 / 8) v = make_uint4(0x3C003C00u, 0u, 0u, 0u);   // halves (1, 1, 0, 0, 0, 0, 0, 0)
            else v = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sA + ((size_t)(j * 16 + (r >> 3)) * 64 + (r & 7) * 8)) = v;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_remote(tc::map_to_cta(&bars[B_AREADY], 0));   // leader included: its own address maps to itself
    }

    if (tid == 128 && leader) {   // layer 0 (SS), M = 256 over both CTAs
        tc::mbar_wait(&bars[B_W], 0);
        tc::mbar_wait_cluster(&bars[B_WPEER], 0);
        tc::mbar_wait_cluster(&bars[B_AREADY], 0);
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::make_idesc_f16(256, N0);
#pragma unroll
        for (int ks = 0; ks < K0P / 16; ++ks) {
            const uint64_t ad = tc::make_smem_desc(tc::smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t bd = tc::make_smem_desc(tc::smem_u32(smem + OFF_W0) + ks * 2 * ((N0 / 2) * 16), (N0 / 2) * 16, 128);
            tc::mma_ss_pair(tmem + 0, ad, bd, idesc, ks > 0);
        }
        tc::mma_commit_pair(&bars[B_D0], 0b11);
    }
    if (warp < 4) {     // epilogue 0 in BOTH CTAs: D0 -> global, relu(D0[:, :128]) -> fp16 -> TMEM cols [256, 320)
        tc::mbar_wait(&bars[B_D0], 0);
        tc::tc_fence_after();
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const int row = rank * 128 + warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < N0 / 32; ++c) {
            uint32_t v[32];
            tc::tmem_ld32(lane_base + c * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) d0_out[(size_t)row * N0 + c * 32 + i] = __uint_as_float(v[i]);
            if (c < K1 / 32) {
                uint32_t h[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = tc::cvt_relu_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                tc::tmem_st16(lane_base + 256 + c * 16, h);
            }
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_remote(tc::map_to_cta(&bars[B_HREADY], 0));
    }
    if (tid == 128 && leader) {   // layer 1 (TS): A = h in each CTA's TMEM, 8 columns per K = 16 step
        tc::mbar_wait_cluster(&bars[B_HREADY], 0);
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::make_idesc_f16(256, N1), idesc16 = tc::make_idesc_f16(256, 16);
        const uint32_t w1 = tc::smem_u32(smem + OFF_W1);
#pragma unroll
        for (int ks = 0; ks < K1 / 16; ++ks) {
            const uint64_t bd = tc::make_smem_desc(w1 + ks * 2 * ((N1 / 2) * 16), (N1 / 2) * 16, 128);
            tc::mma_ts_pair(tmem + 320, tmem + 256 + ks * 8, bd, idesc, ks > 0);
        }
        // N = 16: 8 rows per CTA = local rows 64..71 of each half, accumulated into columns 64..79 of D1
#pragma unroll
        for (int ks = 0; ks < K1 / 16; ++ks) {
            const uint64_t bd = tc::make_smem_desc(w1 + ks * 2 * ((N1 / 2) * 16) + 64 * 16, (N1 / 2) * 16, 128);
            tc::mma_ts_pair(tmem + 320 + 64, tmem + 256 + ks * 8, bd, idesc16, true);
        }
        tc::mma_commit_pair(&bars[B_D1], 0b11);
    }
    if (warp < 4) {
        tc::mbar_wait(&bars[B_D1], 0);
        tc::tc_fence_after();
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const int row = rank * 128 + warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < N1 / 16; ++c) {
            uint32_t v[16];
            tc::tmem_ld16(lane_base + 320 + c * 16, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) d1_out[(size_t)row * N1 + c * 16 + i] = __uint_as_float(v[i]);
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == 4) tc::tmem_dealloc_pair<512>(tmem);
}

}  // namespace probe2
}  // namespace nb

extern "C" int nb_debug_tc_probe2(const void* a0, const void* w0_halves, const void* w1_halves, float* d0_out, float* d1_out,
                                  void* stream) {
    using namespace nb;
    if (!a0 || !w0_halves || !w1_halves || !d0_out || !d1_out) { set_error("nb_debug_tc_probe2: null pointer"); return NB_ERR_BAD_ARG; }
    cudaError_t e = cudaFuncSetAttribute(probe2::tc_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, probe2::SMEM);
    if (e == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = probe2::SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, probe2::tc_probe2_kernel, (const __half*)a0, (const __half*)w0_halves, (const __half*)w1_halves,
                               d0_out, d1_out);
    }
    if (e != cudaSuccess) { set_error("nb_debug_tc_probe2: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
    return NB_OK;
}
