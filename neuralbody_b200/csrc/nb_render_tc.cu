// Tensor-core fused render kernel (NB_PRECISION_TC_FP16) for sm_100a.
//
// One persistent CTA per SM walks 128-point tiles (= 128/S whole rays) and does, per tile, everything
// Renderer.render's chunk loop does upstream (if_clight_renderer.py:107-120): sampling, world->SMPL->grid,
// 4-level trilinear gather, decoder MLP (latent_xyzc.py:99-121, folded as in nb_layout.h), PE, composite.
//
// Data flow (nothing per-point ever touches global memory between the gather and the 24-byte result):
//   producer warps (8)  : geometry + trilinear gather from the fp16 channels-last volume -> fp16 A tile "F"
//                         in shared memory, in the tcgen05 K-major no-swizzle operand layout
//   loader warp         : streams the decoder weights (fp16, pre-packed per K=16 step, consumption order)
//                         from L2 into a 12-slot shared-memory ring with bulk async copies (TMA engine)
//   MMA warp (1 thread) : tcgen05.mma kind::f16, fp32 accumulators in TMEM.  Layer 0 reads A from smem (F);
//                         layers 1..4 read A straight from TMEM, where the previous layer's epilogue left it
//   epilogue warps (4)  : tcgen05.ld accumulator -> cvt.rn.relu.f16x2 -> tcgen05.st next layer's A operand
//                         (activations never visit shared memory); PE(xyz) / PE(view) tile; sigma / rgb
//                         read-out; warp-scan alpha composite; 24 B/ray written to HBM
// Biases ride in the GEMMs as an extra K=16 step against a constant column of ones (hi+lo fp16 split, so
// they stay fp32-accurate); so do alpha_fc (2 extra N rows of layer 3) and rgb_fc (a 16-wide layer 4).
//
// TMEM (512 columns): [0,256) fp32 accumulator | [256,384) hA | [384,512) hB  (fp16 activations, 2 per column)
#include "nb_device.cuh"
#include "nb_tc_ptx.cuh"

namespace nb {
namespace tcr {

constexpr int TP = 128;                       // points per tile = UMMA M
constexpr int NUM_SLOTS = 12;
constexpr int SLOT_BYTES = 8192;              // one K=16 step of an N=256 layer
constexpr int CHUNK_BYTES = 2048;             // one 8-wide K chunk of a 128-row A tile
constexpr int F_CHUNKS = 46;                  // 44 feature chunks + the "ones" K-step (chunks 44,45)
constexpr int PE_CHUNKS = 12;                 // 96-wide per-point tile of layer 3
constexpr int EPI_WARPS = 4, MMA_WARP = 4, LOAD_WARP = 5, PROD_WARP0 = 6, PROD_WARPS = 8;
constexpr int NT = (PROD_WARP0 + PROD_WARPS) * 32;   // 448
constexpr int PROD_THREADS = PROD_WARPS * 32;

// shared-memory map (bytes)
constexpr int OFF_F = 0;
constexpr int OFF_PE = OFF_F + F_CHUNKS * CHUNK_BYTES;            //  94208
constexpr int OFF_RING = OFF_PE + PE_CHUNKS * CHUNK_BYTES;        // 118784
constexpr int OFF_GEOM = OFF_RING + NUM_SLOTS * SLOT_BYTES;       // 217088  float4[128] (wx,wy,wz,z)   producer-owned
constexpr int OFF_GRID = OFF_GEOM + TP * 16;                      // 219136  float4[128] (gx,gy,gz,-)   producer-owned
constexpr int OFF_RAW = OFF_GRID + TP * 16;                       // 221184  float4[128] (r,g,b,sigma)  epilogue-owned
constexpr int OFF_Z = OFF_RAW + TP * 16;                          // 223232  float[128] z               epilogue-owned
constexpr int OFF_XF = OFF_Z + TP * 4;                            // 223744  FrameXf (producer-owned)
constexpr int OFF_BAR = OFF_XF + 128;                             // 223872
constexpr int NUM_BARS = 2 * NUM_SLOTS + 6;
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;                  // 224112
constexpr int SMEM_BYTES = OFF_TMEM + 16;

enum { BAR_W_FULL = 0, BAR_W_EMPTY = NUM_SLOTS, BAR_F_FULL = 2 * NUM_SLOTS, BAR_F_EMPTY, BAR_ACC_FULL, BAR_H_READY,
       BAR_GEOM_FREE, BAR_SPARE };

constexpr uint32_t TM_ACC = 0, TM_HA = 256, TM_HB = 384;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& v, float w) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        acc[2 * i] = fmaf(f.x, w, acc[2 * i]);
        acc[2 * i + 1] = fmaf(f.y, w, acc[2 * i + 1]);
    }
}

struct TileCoord { int b, r0, nr; };
__device__ __forceinline__ TileCoord tile_coord(const RenderParams& P, int tile) {
    TileCoord t;
    t.b = tile / P.groups_per_frame;
    t.r0 = (tile % P.groups_per_frame) * P.rays_per_group;
    t.nr = min(P.rays_per_group, P.n_rays - t.r0);
    return t;
}

__global__ void __launch_bounds__(NT, 1) render_tc_kernel(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = P.n_samples;

    // ------------------------------------------------------------------ one-time set-up
    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == LOAD_WARP * 32) {
        for (int i = 0; i < NUM_SLOTS; ++i) { tc::mbar_init(&bars[BAR_W_FULL + i], 1); tc::mbar_init(&bars[BAR_W_EMPTY + i], 1); }
        tc::mbar_init(&bars[BAR_F_FULL], PROD_WARPS);
        tc::mbar_init(&bars[BAR_F_EMPTY], 1);
        tc::mbar_init(&bars[BAR_ACC_FULL], 1);
        tc::mbar_init(&bars[BAR_H_READY], EPI_WARPS * 32);
        tc::mbar_init(&bars[BAR_GEOM_FREE], EPI_WARPS * 32);
        tc::fence_mbar_init();
    }
    if (warp >= PROD_WARP0) {   // constant "ones" K-step of the A operand: chunk 44 = (1,1,0,..), chunk 45 = 0
        const int pt = tid - PROD_WARP0 * 32;
        if (pt < TP) {
            unsigned char* f = smem + OFF_F;
            *reinterpret_cast<uint4*>(f + (44 * 16 + (pt >> 3)) * 128 + (pt & 7) * 16) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(f + (45 * 16 + (pt >> 3)) * 128 + (pt & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        tc::fence_proxy_async();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int n_tiles = P.n_groups;

    // ================================================================== PRODUCERS: geometry + gather
    if (warp >= PROD_WARP0) {
        const int pt = tid - PROD_WARP0 * 32;          // 0..255
        const int pw = warp - PROD_WARP0;
        float4* geom = reinterpret_cast<float4*>(smem + OFF_GEOM);
        float4* grid = reinterpret_cast<float4*>(smem + OFF_GRID);
        FrameXf* xf = reinterpret_cast<FrameXf*>(smem + OFF_XF);
        unsigned char* F = smem + OFF_F;
        const unsigned char* volbase = reinterpret_cast<const unsigned char*>(P.volume);
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc_ = tile_coord(P, tile);
            // F and grid are free once layer 0 of the previous tile has been consumed; geom once the
            // epilogue has copied what it needs from the previous tile
            tc::mbar_wait(&bars[BAR_F_EMPTY], (it & 1) ^ 1);
            tc::mbar_wait(&bars[BAR_GEOM_FREE], (it & 1) ^ 1);
            if (pt < 9) xf->R[pt] = __ldg(P.R + tc_.b * 9 + pt);
            if (pt < 3) {
                xf->Th[pt] = __ldg(P.Th + tc_.b * 3 + pt);
                xf->min_dhw[pt] = __ldg(P.bounds + tc_.b * 6 + (2 - pt));
                xf->voxel[pt] = P.voxel_size[pt];
                xf->out_sh[pt] = P.out_sh[pt];
            }
            named_bar_sync(1, PROD_THREADS);
            if (pt < TP) {
                const int ry = pt / S, s = pt % S;
                float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), gr = make_float4(-4.f, -4.f, -4.f, 0.f);
                if (ry < tc_.nr) {
                    const size_t ri = (size_t)tc_.b * P.n_rays + tc_.r0 + ry;
                    const float ox = __ldg(P.ray_o + ri * 3), oy = __ldg(P.ray_o + ri * 3 + 1), oz = __ldg(P.ray_o + ri * 3 + 2);
                    const float dx = __ldg(P.ray_d + ri * 3), dy = __ldg(P.ray_d + ri * 3 + 1), dz = __ldg(P.ray_d + ri * 3 + 2);
                    const float z = z_sample(__ldg(P.near + ri), __ldg(P.far + ri), P.t_vals, s, S,
                                             P.t_rand ? P.t_rand + ri * S : nullptr);
                    gm.x = __fadd_rn(ox, __fmul_rn(dx, z));
                    gm.y = __fadd_rn(oy, __fmul_rn(dy, z));
                    gm.z = __fadd_rn(oz, __fmul_rn(dz, z));
                    gm.w = z;
                    world_to_grid(*xf, gm.x, gm.y, gm.z, gr.x, gr.y, gr.z);
                }
                geom[pt] = gm;
                grid[pt] = gr;
            }
            named_bar_sync(1, PROD_THREADS);
            // gather: a warp-iteration covers 8 points x 4 consecutive 8-channel chunks (lane%8 = point,
            // lane/8 = chunk): every quarter-warp stores 128 contiguous bytes of one chunk (conflict-free),
            // every point reads 64 contiguous bytes per corner.
            const int lp = lane & 7, lj = lane >> 3;
            for (int blk = pw; blk < 16 * 11; blk += PROD_WARPS) {
                const int p = (blk & 15) * 8 + lp;
                const int j = (blk >> 4) * 4 + lj;           // chunk 0..43
                int lvl, c0;
                if (j < 4) { lvl = 0; c0 = j * 8; }
                else if (j < 12) { lvl = 1; c0 = (j - 4) * 8; }
                else if (j < 28) { lvl = 2; c0 = (j - 12) * 8; }
                else { lvl = 3; c0 = (j - 28) * 8; }
                const int C = P.lvl_C[lvl], D = P.lvl_D[lvl], H = P.lvl_H[lvl], W = P.lvl_W[lvl];
                const float4 g = grid[p];
                Corners cn;
                corner_setup(unnormalize(g.x, W), unnormalize(g.y, H), unnormalize(g.z, D), W, H, D, cn);
                const __half* vol = reinterpret_cast<const __half*>(volbase + P.lvl_off[lvl]) + (size_t)tc_.b * P.lvl_bstride[lvl];
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                uint4 v[8];
                float wgt[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
                    const bool ok = corner_valid(cn, dx, dy, dz, W, H, D);
                    wgt[c] = ok ? corner_weight(cn, dx, dy, dz) : 0.f;
                    const size_t vox = ok ? ((size_t)(cn.z0 + dz) * H + (cn.y0 + dy)) * W + (cn.x0 + dx) : 0;
                    v[c] = ok ? ldg_nc_v4(vol + vox * C + c0) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) fma8(acc, v[c], wgt[c]);
                uint4 o;
                o.x = tc::cvt_f16x2(acc[0], acc[1]); o.y = tc::cvt_f16x2(acc[2], acc[3]);
                o.z = tc::cvt_f16x2(acc[4], acc[5]); o.w = tc::cvt_f16x2(acc[6], acc[7]);
                *reinterpret_cast<uint4*>(F + (j * 16 + (p >> 3)) * 128 + (p & 7) * 16) = o;
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars[BAR_F_FULL]);
        }
    }
    // ================================================================== LOADER: weight stream -> ring
    else if (warp == LOAD_WARP) {
        if (lane == 0) {
            uint32_t cnt = 0;
            const unsigned char* seq = reinterpret_cast<const unsigned char*>(P.wf16);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int b = tile / P.groups_per_frame;
                for (int st = 0; st < kStepsPerTile; ++st, ++cnt) {
                    const unsigned char* src;
                    uint32_t bytes;
                    if (st < kStepsL0 + kStepsL1 + kStepsL2) { src = seq + (size_t)st * SLOT_BYTES; bytes = SLOT_BYTES; }
                    else if (st < kStepsL0 + kStepsL1 + kStepsL2 + kStepsL3) {
                        const int s3 = st - (kStepsL0 + kStepsL1 + kStepsL2);
                        bytes = kStepHalves3 * 2;
                        src = (s3 == kStepsL3 - 1) ? reinterpret_cast<const unsigned char*>(P.wframe) + (size_t)b * bytes
                                                   : seq + sL3 * 2 + (size_t)s3 * bytes;
                    } else {
                        const int s4 = st - (kStepsL0 + kStepsL1 + kStepsL2 + kStepsL3);
                        bytes = kStepHalves4 * 2;
                        src = seq + sL4 * 2 + (size_t)s4 * bytes;
                    }
                    const uint32_t slot = cnt % NUM_SLOTS, round = cnt / NUM_SLOTS;
                    tc::mbar_wait(&bars[BAR_W_EMPTY + slot], (round & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&bars[BAR_W_FULL + slot], bytes);
                    tc::bulk_g2s(smem + OFF_RING + slot * SLOT_BYTES, src, bytes, &bars[BAR_W_FULL + slot]);
                }
            }
        }
    }
    // ================================================================== MMA ISSUER
    else if (warp == MMA_WARP) {
        if (lane == 0) {
            uint32_t cnt = 0, hcnt = 0;
            int it = 0;
            const uint32_t f_addr = tc::smem_u32(smem + OFF_F), pe_addr = tc::smem_u32(smem + OFF_PE);
            const uint32_t ring_addr = tc::smem_u32(smem + OFF_RING);
            constexpr uint32_t ID256 = tc::make_idesc_f16(128, 256), ID3 = tc::make_idesc_f16(128, kN3),
                               ID4 = tc::make_idesc_f16(128, kN4);
            auto wait_slot = [&](uint32_t& slot) {
                slot = cnt % NUM_SLOTS;
                tc::mbar_wait(&bars[BAR_W_FULL + slot], (cnt / NUM_SLOTS) & 1);
                tc::tc_fence_after();
            };
            auto release_slot = [&](uint32_t slot) { tc::mma_commit(&bars[BAR_W_EMPTY + slot]); ++cnt; };
            auto a_desc = [&](uint32_t base, int ks) { return tc::make_smem_desc(base + ks * 2 * CHUNK_BYTES, CHUNK_BYTES, 128); };
            auto b_desc = [&](uint32_t slot, int N) { return tc::make_smem_desc(ring_addr + slot * SLOT_BYTES, N * 16, 128); };
            auto wait_h = [&]() { tc::mbar_wait(&bars[BAR_H_READY], hcnt & 1); ++hcnt; tc::tc_fence_after(); };

            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                uint32_t slot;
                // ---- layer 0: A = F (smem), 22 feature steps + ones step
                tc::mbar_wait(&bars[BAR_F_FULL], it & 1);
                if (it > 0) wait_h();                 // previous tile's last epilogue has drained the accumulator
                tc::tc_fence_after();
                for (int ks = 0; ks < kStepsL0; ++ks) {
                    wait_slot(slot);
                    tc::mma_ss(tmem + TM_ACC, a_desc(f_addr, ks), b_desc(slot, 256), ID256, ks > 0);
                    release_slot(slot);
                }
                tc::mma_commit(&bars[BAR_F_EMPTY]);
                tc::mma_commit(&bars[BAR_ACC_FULL]);
                // ---- layers 1, 2: A = h (TMEM), 16 steps + ones step (A from smem chunk pair 22)
                for (int layer = 1; layer <= 2; ++layer) {
                    wait_h();
                    const uint32_t hin = (layer == 1) ? TM_HA : TM_HB;
                    for (int ks = 0; ks < 16; ++ks) {
                        wait_slot(slot);
                        tc::mma_ts(tmem + TM_ACC, tmem + hin + ks * 8, b_desc(slot, 256), ID256, ks > 0);
                        release_slot(slot);
                    }
                    wait_slot(slot);
                    tc::mma_ss(tmem + TM_ACC, a_desc(f_addr, 22), b_desc(slot, 256), ID256, true);
                    release_slot(slot);
                    tc::mma_commit(&bars[BAR_ACC_FULL]);
                }
                // ---- layer 3: N = 144: A = h2 (hA), then the per-point tile (PE | ones) from smem
                wait_h();
                for (int ks = 0; ks < 16; ++ks) {
                    wait_slot(slot);
                    tc::mma_ts(tmem + TM_ACC, tmem + TM_HA + ks * 8, b_desc(slot, kN3), ID3, ks > 0);
                    release_slot(slot);
                }
                for (int ks = 0; ks < kPeK / 16; ++ks) {
                    wait_slot(slot);
                    tc::mma_ss(tmem + TM_ACC, a_desc(pe_addr, ks), b_desc(slot, kN3), ID3, true);
                    release_slot(slot);
                }
                tc::mma_commit(&bars[BAR_ACC_FULL]);
                // ---- layer 4: N = 16: A = relu(w) (hB, K = 128) + ones step
                wait_h();
                for (int ks = 0; ks < 8; ++ks) {
                    wait_slot(slot);
                    tc::mma_ts(tmem + TM_ACC, tmem + TM_HB + ks * 8, b_desc(slot, kN4), ID4, ks > 0);
                    release_slot(slot);
                }
                wait_slot(slot);
                tc::mma_ss(tmem + TM_ACC, a_desc(f_addr, 22), b_desc(slot, kN4), ID4, true);
                release_slot(slot);
                tc::mma_commit(&bars[BAR_ACC_FULL]);
            }
        }
    }
    // ================================================================== EPILOGUE (thread = tile row)
    else {
        const int row = tid;                                  // 0..127
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const float4* geom = reinterpret_cast<const float4*>(smem + OFF_GEOM);
        float4* rawbuf = reinterpret_cast<float4*>(smem + OFF_RAW);
        float* zbuf = reinterpret_cast<float*>(smem + OFF_Z);
        unsigned char* PE = smem + OFF_PE;
        uint32_t acnt = 0;
        int it = 0;
        auto wait_acc = [&]() { tc::mbar_wait(&bars[BAR_ACC_FULL], acnt & 1); ++acnt; tc::tc_fence_after(); };
        // accumulator columns [0, ncols) -> relu -> fp16 pairs -> TMEM h buffer
        auto relu_to_h = [&](uint32_t hout, int ncols) {
            for (int c = 0; c < ncols / 32; ++c) {
                uint32_t v[32];
                tc::tmem_ld32(lane_base + TM_ACC + c * 32, v);
                tc::tmem_ld_wait();
                uint32_t h[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = tc::cvt_relu_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                tc::tmem_st16(lane_base + hout + c * 16, h);
            }
            tc::tmem_st_wait();
        };
        auto h_done = [&]() { tc::tc_fence_before(); tc::mbar_arrive(&bars[BAR_H_READY]); };

        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc_ = tile_coord(P, tile);
            const int ry = row / S;
            const bool valid = ry < tc_.nr;
            const size_t ri = (size_t)tc_.b * P.n_rays + tc_.r0 + (valid ? ry : 0);
            // copy what this thread needs from the producer-owned geometry, then hand it back
            tc::mbar_wait(&bars[BAR_F_FULL], it & 1);
            const float4 gm = geom[row];
            zbuf[row] = gm.w;
            tc::mbar_arrive(&bars[BAR_GEOM_FREE]);

            // ---- layer 0 epilogue -> hA
            wait_acc();
            relu_to_h(TM_HA, 256);
            h_done();
            // ---- per-point tile of layer 3: [PE10(world xyz) 63 | 0 | PE4(viewdir) 27 | 0 | 1 | 1 | 0 | 0]
            {
                auto store8 = [&](int j, const float* f) {
                    uint4 o;
                    o.x = tc::cvt_f16x2(f[0], f[1]); o.y = tc::cvt_f16x2(f[2], f[3]);
                    o.z = tc::cvt_f16x2(f[4], f[5]); o.w = tc::cvt_f16x2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(PE + (j * 16 + (row >> 3)) * 128 + (row & 7) * 16) = o;
                };
                {
                    float pf[64];
                    positional_embed<10>(gm.x, gm.y, gm.z, [&](int j, float v) { pf[j] = v; });
                    pf[63] = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) store8(j, pf + 8 * j);
                }
                {
                    float pf[32];
                    const float dx = __ldg(P.ray_d + ri * 3), dy = __ldg(P.ray_d + ri * 3 + 1), dz = __ldg(P.ray_d + ri * 3 + 2);
                    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                    positional_embed<4>(__fdiv_rn(dx, nrm), __fdiv_rn(dy, nrm), __fdiv_rn(dz, nrm), [&](int j, float v) { pf[j] = v; });
                    pf[27] = 0.f; pf[28] = 1.f; pf[29] = 1.f; pf[30] = 0.f; pf[31] = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) store8(8 + j, pf + 8 * j);
                }
                tc::fence_proxy_async();
            }
            // ---- layer 1 -> hB, layer 2 -> hA
            wait_acc();
            relu_to_h(TM_HB, 256);
            h_done();
            wait_acc();
            relu_to_h(TM_HA, 256);
            h_done();
            // ---- layer 3: colour hidden -> hB (fp16), sigma = acc[128] + acc[129]
            wait_acc();
            float sigma;
            {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + TM_ACC + 128, v);
                tc::tmem_ld_wait();
                sigma = __uint_as_float(v[0]) + __uint_as_float(v[1]);
            }
            relu_to_h(TM_HB, 128);
            h_done();
            // ---- layer 4: rgb logits = hi rows + lo rows
            wait_acc();
            {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + TM_ACC, v);
                tc::tmem_ld_wait();
                rawbuf[row] = make_float4(__uint_as_float(v[0]) + __uint_as_float(v[3]), __uint_as_float(v[1]) + __uint_as_float(v[4]),
                                          __uint_as_float(v[2]) + __uint_as_float(v[5]), sigma);
            }
            h_done();                                  // accumulator drained: next tile's layer 0 may start
            named_bar_sync(2, EPI_WARPS * 32);
            // ---- composite (a10): one warp per ray
            for (int g = warp; g < tc_.nr; g += EPI_WARPS) {
                const size_t rg = (size_t)tc_.b * P.n_rays + tc_.r0 + g;
                const float dx = __ldg(P.ray_d + rg * 3), dy = __ldg(P.ray_d + rg * 3 + 1), dz = __ldg(P.ray_d + rg * 3 + 2);
                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                float* wout = P.weights ? P.weights + rg * S : nullptr;
                RayOut o = composite_ray(rawbuf + g * S, zbuf + g * S, S, nrm, wout, lane);
                if (P.raw) {
                    float4* rdst = reinterpret_cast<float4*>(P.raw) + rg * S;
                    for (int s = lane; s < S; s += 32) rdst[s] = rawbuf[g * S + s];
                }
                if (lane == 0) {
                    const float add = P.white_bkgd ? __fsub_rn(1.f, o.acc) : 0.f;
                    P.rgb_map[rg * 3 + 0] = o.r + add;
                    P.rgb_map[rg * 3 + 1] = o.g + add;
                    P.rgb_map[rg * 3 + 2] = o.b + add;
                    P.depth_map[rg] = o.depth;
                    P.acc_map[rg] = o.acc;
                    P.disp_map[rg] = disparity(o.depth, o.acc);
                }
            }
            named_bar_sync(2, EPI_WARPS * 32);     // rawbuf / zbuf reused by the next tile
        }
    }

    // ------------------------------------------------------------------ teardown
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        tc::tmem_dealloc<512>(tmem);
    }
}

}  // namespace tcr

bool tc_available() { return true; }

int launch_render_tc(const RenderParams& p_in, int volume_dtype, cudaStream_t stream) {
    RenderParams p = p_in;
    const int S = p.n_samples;
    if (volume_dtype != NB_DTYPE_F16) { set_error("NB_PRECISION_TC_FP16 needs an fp16-packed volume (NB_DTYPE_F16)"); return NB_ERR_UNSUPPORTED; }
    if (S > tcr::TP) {
        set_error("NB_PRECISION_TC_FP16 supports n_samples <= 128 (got %d); use NB_PRECISION_FP32", S);
        return NB_ERR_UNSUPPORTED;
    }
    p.rays_per_group = tcr::TP / S;
    p.tiles_per_group = 1;
    p.groups_per_frame = (p.n_rays + p.rays_per_group - 1) / p.rays_per_group;
    p.n_groups = p.groups_per_frame * p.batch;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.n_groups < sms ? p.n_groups : sms;
    if (grid == 0) return NB_OK;
    cudaError_t e = cudaFuncSetAttribute(tcr::render_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcr::SMEM_BYTES);
    if (e == cudaSuccess) {
        tcr::render_tc_kernel<<<grid, tcr::NT, tcr::SMEM_BYTES, stream>>>(p);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { set_error("render_tc launch failed: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
    return NB_OK;
}

}  // namespace nb
