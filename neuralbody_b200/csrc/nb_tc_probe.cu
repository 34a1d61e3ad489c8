// Diagnostic: a two-layer tcgen05 micro-pipeline on ONE 128-row tile that exercises, in isolation,
// every hardware contract the fused tensor-core render kernel relies on:
//   * K-major no-swizzle shared-memory operand layout + descriptor fields (LBO/SBO),
//   * the instruction descriptor for kind::f16 with fp32 accumulation,
//   * bulk async copy (TMA engine) + mbarrier transaction counts for the weight operand,
//   * tcgen05.alloc / mma (SS and TS forms) / commit / ld / st,
//   * the bias-as-extra-K-step trick (A column of ones, B row = bias split hi+lo),
//   * cvt.rn.relu.f16x2 packing of the next layer's A operand straight into TMEM.
// layer 0: D0[128x128] = [A0 | 1 1 0..] (K=64+16) * W0p^T      (SS: A from smem)
// layer 1: D1[128x64]  = relu(D0) (fp16, TMEM-resident) * W1^T  (TS: A from TMEM)
#include "nb_internal.h"
#include "nb_tc_ptx.cuh"

namespace nb {
namespace probe {

constexpr int K0 = 64, K0P = 80, N0 = 128, K1 = 128, N1 = 64;

__global__ void __launch_bounds__(192, 1)
tc_probe_kernel(const __half* __restrict__ a0, const __half* __restrict__ w0p, const __half* __restrict__ w1p,
                float* __restrict__ d0_out, float* __restrict__ d1_out, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);                          // [K0P/8][16][8][8] halves = 20480 B
    __half* sW0 = reinterpret_cast<__half*>(smem + 20480);                 // N0 x K0P canonical = 20480 B
    __half* sW1 = reinterpret_cast<__half*>(smem + 40960);                 // N1 x K1 canonical  = 16384 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 57344);            // [0] weights, [1] mma0, [2] mma1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 57344 + 64);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 4) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 160) {
        tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::mbar_init(&bars[2], 1);
        tc::fence_mbar_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (tid == 160) {   // weight loader: two bulk copies complete on bars[0]
        tc::mbar_arrive_expect_tx(&bars[0], N0 * K0P * 2 + N1 * K1 * 2);
        tc::bulk_g2s(sW0, w0p, N0 * K0P * 2, &bars[0]);
        tc::bulk_g2s(sW1, w1p, N1 * K1 * 2, &bars[0]);
    }
    if (tid < 128) {    // A0 tile: row r -> canonical [k/8][r/8][r%8][8]
        const int r = tid;
#pragma unroll
        for (int j = 0; j < K0P / 8; ++j) {
            uint4 v;
            if (j < K0 / 8) v = *reinterpret_cast<const uint4*>(a0 + r * K0 + j * 8);
            else if (j == K0 / 8) v = make_uint4(0x3C003C00u, 0u, 0u, 0u);   // halves (1, 1, 0, 0, 0, 0, 0, 0)
            else v = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sA + ((size_t)(j * 16 + (r >> 3)) * 64 + (r & 7) * 8)) = v;
        }
        tc::fence_proxy_async();
    }
    __syncthreads();

    const bool swap = variant & 1;
    auto desc = [&](const void* base, uint32_t byte_off, uint32_t lbo, uint32_t sbo) {
        return tc::make_smem_desc(tc::smem_u32(base) + byte_off, swap ? sbo : lbo, swap ? lbo : sbo);
    };

    if (tid == 128) {   // layer 0 (SS)
        tc::mbar_wait(&bars[0], 0);
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::make_idesc_f16(128, N0);
#pragma unroll
        for (int ks = 0; ks < K0P / 16; ++ks) {
            const uint64_t ad = desc(sA, ks * 2 * (128 * 16), 128 * 16, 128);       // A: LBO = 128 rows * 16 B
            const uint64_t bd = desc(sW0, ks * 2 * (N0 * 16), N0 * 16, 128);        // B: LBO = N rows * 16 B
            tc::mma_ss(tmem + 0, ad, bd, idesc, ks > 0);
        }
        tc::mma_commit(&bars[1]);
    }
    if (warp < 4) {     // epilogue 0: D0 -> global (fp32) and relu -> fp16 -> TMEM cols [256, 320)
        tc::mbar_wait(&bars[1], 0);
        tc::tc_fence_after();
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const int row = warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < N0 / 32; ++c) {
            uint32_t v[32];
            tc::tmem_ld32(lane_base + c * 32, v);
            tc::tmem_ld_wait();
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 32; ++i) d0_out[row * N0 + c * 32 + i] = __uint_as_float(v[i]);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float lo = __uint_as_float(v[2 * i]), hi = __uint_as_float(v[2 * i + 1]);
                h[i] = (variant & 2) ? tc::cvt_relu_f16x2(hi, lo) : tc::cvt_relu_f16x2(lo, hi);
            }
            tc::tmem_st16(lane_base + 256 + c * 16, h);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
    }
    __syncthreads();
    if (tid == 128) {   // layer 1 (TS): A = h in TMEM, 8 columns per K=16 step
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::make_idesc_f16(128, N1);
#pragma unroll
        for (int ks = 0; ks < K1 / 16; ++ks) {
            const uint64_t bd = desc(sW1, ks * 2 * (N1 * 16), N1 * 16, 128);
            tc::mma_ts(tmem + 128, tmem + 256 + ks * 8, bd, idesc, ks > 0);
        }
        tc::mma_commit(&bars[2]);
    }
    if (warp < 4) {
        tc::mbar_wait(&bars[2], 0);
        tc::tc_fence_after();
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const int row = warp * 32 + lane;
#pragma unroll
        for (int c = 0; c < N1 / 32; ++c) {
            uint32_t v[32];
            tc::tmem_ld32(lane_base + 128 + c * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) d1_out[row * N1 + c * 32 + i] = __uint_as_float(v[i]);
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc<512>(tmem);
}

}  // namespace probe
}  // namespace nb

extern "C" int nb_debug_tc_probe(const void* a0, const void* w0_packed, const void* w1_packed, float* d0_out, float* d1_out,
                                 int variant, void* stream) {
    using namespace nb;
    if (!a0 || !w0_packed || !w1_packed || !d0_out || !d1_out) { set_error("nb_debug_tc_probe: null pointer"); return NB_ERR_BAD_ARG; }
    const int smem = 57344 + 128;
    cudaError_t e = cudaFuncSetAttribute(probe::tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) {
        probe::tc_probe_kernel<<<1, 192, smem, (cudaStream_t)stream>>>((const __half*)a0, (const __half*)w0_packed,
                                                                      (const __half*)w1_packed, d0_out, d1_out, variant);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { set_error("nb_debug_tc_probe: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
    return NB_OK;
}
