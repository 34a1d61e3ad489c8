// Tensor-core fused render kernel for sm_100a (NB_PRECISION_TC_FP16 = 1 pass, NB_PRECISION_TC_FP16X3 = 3 passes).
//
// One persistent CTA per SM walks 128-point tiles (= floor(128/S) whole rays) and does, per tile, everything
// Renderer.render's chunk loop does upstream (if_clight_renderer.py:107-120): sampling, world->SMPL->grid,
// 4-level trilinear gather, decoder MLP (latent_xyzc.py:99-121, folded as in nb_layout.h), PE, composite.
// Nothing per-point touches global memory between the gather and the 24-byte per-ray result.
//
//   producer warps (16) : geometry, then the trilinear gather from the channels-last volume into the fp16 A
//                         operand of layer 0, written in the tcgen05 K-major no-swizzle layout into a 2-deep
//                         ring of 64-channel K SEGMENTS (layer 0 is K-pipelined against the gather)
//   loader warp         : streams the decoder weights (fp16, pre-packed groups of 4 K=16 steps per bulk copy, in
//                         consumption order) from L2 into a 3 x 32 KB shared-memory ring (TMA engine, UBLKCP),
//                         multicast to both CTAs of a cluster pair
//   MMA warp (1 thread) : tcgen05.mma kind::f16, fp32 accumulator in TMEM.  Layer 0 reads A from shared memory;
//                         layers 1..4 read A straight from TMEM, where the previous epilogue left it
//   epilogue warps (4)  : tcgen05.ld accumulator -> relu -> fp16 (hi [+ lo]) -> tcgen05.st as the next layer's
//                         A operand (activations never visit shared memory); PE tile; sigma / rgb read-out;
//                         warp-scan alpha composite
// Biases ride in the GEMMs as an extra K=16 step against a constant column of ones (hi+lo fp16 split => fp32
// accurate); so do alpha_fc (2 extra N rows of layer 3) and rgb_fc (a 16-wide layer 4).
//
// Precision.  fp16 has an 11-bit significand; on a trained-like decoder one fp16 rounding of ANY density-path
// operand (volume, features, weights, h0, h1, h2) moves depth_map by ~1e-3, the whole parity budget.  The
// 3-pass mode therefore carries every density-path operand as hi + lo fp16 (22 bits) and issues
//   A_hi W_hi + A_lo W_hi + A_hi W_lo      (the A_lo W_lo term is 2^-22 relative and dropped)
// for layers 0-2 and the alpha rows, gathering from an fp32 volume; the colour tail stays single fp16.
//
// TMEM (512 columns): [0,256) fp32 accumulator | [256,384) h_hi | [384,512) h_lo  (fp16 pairs, in place)
#include "nb_tc_common.cuh"

namespace nb {
namespace tcr {

constexpr int TP = 128;                       // points per tile = UMMA M
constexpr int NUM_SLOTS = 3;
constexpr int STEP_BYTES = 8192;              // one K=16 step of an N=256 layer
constexpr int SLOT_BYTES = 4 * STEP_BYTES;    // one group of up to 4 K-steps per ring slot / bulk copy / mbarrier hand-off
constexpr int CHUNK_BYTES = 2048;             // one 8-wide K chunk of a 128-row A tile
constexpr int SEG_CHUNKS = 8;                 // 64 channels per segment
constexpr int NUM_SEGS = 6;                   // 44 feature chunks = 5 x 8 + 4
constexpr int SEG_RING_BYTES = 6 * SEG_CHUNKS * CHUNK_BYTES;   // 96 KB: 3 x (hi + lo plane) in the 3-pass mode, 6 x hi plane (a whole
                                                              // tile of gather look-ahead) in the 1-pass mode
constexpr int MAX_SEG_BUFS = 6;
constexpr int PE_CHUNKS = 12;                 // 96-wide per-point tile of layer 3
constexpr int EPI_WARPS = 4, MMA_WARP = 4, LOAD_WARP = 5, PROD_WARP0 = 6, PROD_WARPS = 16;
constexpr int NT = (PROD_WARP0 + PROD_WARPS) * 32;   // 704
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int PTS_PER_GROUP = TP / (PROD_WARPS * 4);   // an 8-lane group owns points g, g+64
constexpr int CLUSTER = 2;                    // CTAs sharing one weight stream through TMA multicast

// shared-memory map (bytes)
constexpr int OFF_SEG = 0;                                         // segment ring
constexpr int OFF_ONES = OFF_SEG + SEG_RING_BYTES;                 // constant (1,1,0..) K-step, 2 chunks
constexpr int OFF_PE = OFF_ONES + 2 * CHUNK_BYTES;                 //  69632
constexpr int OFF_RING = OFF_PE + PE_CHUNKS * CHUNK_BYTES;         //  94208
constexpr int OFF_GEOM = OFF_RING + NUM_SLOTS * SLOT_BYTES;        // 192512  float4[128] (wx,wy,wz,z)   producer-owned
constexpr int OFF_GRID = OFF_GEOM + TP * 16;                       // float4[128] (gx,gy,gz,-)           producer-owned
constexpr int OFF_RAW = OFF_GRID + TP * 16;                        // float4[128] (r,g,b,sigma)          epilogue-owned
constexpr int OFF_Z = OFF_RAW + TP * 16;                           // float[128] z                       epilogue-owned
constexpr int OFF_XF = OFF_Z + TP * 4;                             // FrameXf (producer-owned)
constexpr int OFF_BAR = OFF_XF + 128;
enum { BAR_W_FULL = 0, BAR_W_EMPTY = NUM_SLOTS, BAR_SEG_FULL = 2 * NUM_SLOTS, BAR_SEG_EMPTY = 2 * NUM_SLOTS + MAX_SEG_BUFS,
       BAR_ACC_FULL = 2 * NUM_SLOTS + 2 * MAX_SEG_BUFS, BAR_H_READY, BAR_GEOM_FULL, BAR_GEOM_FREE, NUM_BARS };
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;

constexpr uint32_t TM_ACC = 0, TM_HI = 256, TM_LO = 384;

struct TileCoord { int b, r0, nr; };
__device__ __forceinline__ TileCoord tile_coord(const RenderParams& P, int tile) {
    TileCoord t;
    if (tile >= P.n_groups) { t.b = 0; t.r0 = 0; t.nr = 0; return t; }   // padding tile (keeps a cluster in lockstep)
    t.b = tile / P.groups_per_frame;
    t.r0 = (tile % P.groups_per_frame) * P.rays_per_group;
    t.nr = min(P.rays_per_group, P.n_rays - t.r0);
    return t;
}

template <int NP, typename VT>
__global__ void __launch_bounds__(NT, 1) render_tc_kernel(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = P.n_samples;
    constexpr int NUM_SEG_BUFS = (NP == 3) ? 3 : 6;          // segment ring depth (how far the gather may run ahead)
    constexpr int SEG_BYTES = SEG_RING_BYTES / NUM_SEG_BUFS;  // hi (+ lo) plane of one 64-channel segment

    // ------------------------------------------------------------------ one-time set-up
    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == LOAD_WARP * 32) {
        for (int i = 0; i < NUM_SLOTS; ++i) { tc::mbar_init(&bars[BAR_W_FULL + i], 1); tc::mbar_init(&bars[BAR_W_EMPTY + i], CLUSTER); }
        for (int i = 0; i < NUM_SEG_BUFS; ++i) { tc::mbar_init(&bars[BAR_SEG_FULL + i], PROD_WARPS); tc::mbar_init(&bars[BAR_SEG_EMPTY + i], 1); }
        tc::mbar_init(&bars[BAR_ACC_FULL], 1);
        tc::mbar_init(&bars[BAR_H_READY], EPI_WARPS * 32);
        tc::mbar_init(&bars[BAR_GEOM_FULL], PROD_WARPS);
        tc::mbar_init(&bars[BAR_GEOM_FREE], EPI_WARPS * 32);
        tc::fence_mbar_init();
    }
    if (warp >= PROD_WARP0) {   // constant "ones" K-step of the A operand: chunk 0 = (1,1,0,..), chunk 1 = 0
        const int pt = tid - PROD_WARP0 * 32;
        if (pt < TP) {
            unsigned char* o = smem + OFF_ONES;
            *reinterpret_cast<uint4*>(o + (pt >> 3) * 128 + (pt & 7) * 16) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(o + CHUNK_BYTES + (pt >> 3) * 128 + (pt & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        tc::fence_proxy_async();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) tc::cluster_sync_all();      // peers' mbarriers are initialised before any multicast lands
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // every CTA of a cluster walks the same number of tiles (the weight stream is shared in lockstep)
    const int n_iters = (P.n_groups + gridDim.x - 1) / gridDim.x;
    const int n_tiles = n_iters * gridDim.x;
    const uint32_t crank = CLUSTER > 1 ? tc::cluster_ctarank() : 0u;
    constexpr uint16_t CMASK = (1u << CLUSTER) - 1;

    // ================================================================== PRODUCERS: geometry + gather
    if (warp >= PROD_WARP0) {
        const int pt = tid - PROD_WARP0 * 32;
        const int pw = warp - PROD_WARP0;
        float4* geom = reinterpret_cast<float4*>(smem + OFF_GEOM);
        float4* grid = reinterpret_cast<float4*>(smem + OFF_GRID);
        FrameXf* xf = reinterpret_cast<FrameXf*>(smem + OFF_XF);
        const unsigned char* volbase = reinterpret_cast<const unsigned char*>(P.volume);
        // an 8-lane group owns the tile rows {grp, grp + 64}; lane t of the group owns channels 4t..4t+3 of
        // every 32-channel unit.  Units in K order: u0 = L1, u1-2 = L2, u3-6 = L3, u7-10 = L4; segment s = units 2s, 2s+1.
        const int grp = pw * 4 + (lane >> 3);
        const int t = lane & 7;
        const uint32_t* occ_base = reinterpret_cast<const uint32_t*>(volbase);
        Tracer tr;
        tr.init((pw == 0 && lane == 0) ? P.trace : nullptr, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc_ = tile_coord(P, tile);
            tc::mbar_wait(&bars[BAR_GEOM_FREE], (it & 1) ^ 1);    // epilogue has copied the previous tile's geometry
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
                                             P.t_rand ? P.t_rand + ri * S : nullptr, P.z_user ? P.z_user + ri * S : nullptr);
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
            if (lane == 0) tc::mbar_arrive(&bars[BAR_GEOM_FULL]);
            tr.ev(1);                                            // geometry done

            float4 g[PTS_PER_GROUP];
#pragma unroll
            for (int pp = 0; pp < PTS_PER_GROUP; ++pp) g[pp] = grid[grp + 64 * pp];
            uint32_t coff[PTS_PER_GROUP][8];   // element offset of each corner voxel inside the level
            float cw[PTS_PER_GROUP][8];        // corner weight (0 when out of range)
            bool occupied[PTS_PER_GROUP];      // cell-occupancy bit: false => this level interpolates exact zeros
            const VT* vol = nullptr;
            int cur_lvl = -1;
            for (int seg = 0; seg < NUM_SEGS; ++seg) {
                const uint32_t gseg = (uint32_t)it * NUM_SEGS + seg;
                const uint32_t buf = gseg % NUM_SEG_BUFS;
                tc::mbar_wait(&bars[BAR_SEG_EMPTY + buf], ((gseg / NUM_SEG_BUFS) & 1) ^ 1);
                tr.ev(10 + seg);                                 // segment buffer free
                unsigned char* hi_plane = smem + OFF_SEG + buf * SEG_BYTES;
                unsigned char* lo_plane = hi_plane + SEG_CHUNKS * CHUNK_BYTES;
                const int nunits = (seg == NUM_SEGS - 1) ? 1 : 2;
                for (int uu = 0; uu < nunits; ++uu) {
                    const int unit = 2 * seg + uu;
                    int lvl, c0;
                    if (unit < 1) { lvl = 0; c0 = 0; }
                    else if (unit < 3) { lvl = 1; c0 = (unit - 1) * 32; }
                    else if (unit < 7) { lvl = 2; c0 = (unit - 3) * 32; }
                    else { lvl = 3; c0 = (unit - 7) * 32; }
                    if (lvl != cur_lvl) {
                        cur_lvl = lvl;
                        const int C = P.lvl_C[lvl], D = P.lvl_D[lvl], H = P.lvl_H[lvl], W = P.lvl_W[lvl];
                        vol = reinterpret_cast<const VT*>(volbase + P.lvl_off[lvl]) + (size_t)tc_.b * P.lvl_bstride[lvl];
                        const uint32_t* cellbits = occ_base + P.occ_off[lvl] / 4 + (size_t)tc_.b * P.occ_bstride[lvl];
#pragma unroll
                        for (int pp = 0; pp < PTS_PER_GROUP; ++pp) {
                            Corners cn;
                            corner_setup(unnormalize(g[pp].x, W), unnormalize(g[pp].y, H), unnormalize(g[pp].z, D), W, H, D, cn);
                            occupied[pp] = false;
                            if (cn.x0 != -2) {
                                const uint32_t cell = ((uint32_t)(cn.z0 + 1) * (H + 1) + (cn.y0 + 1)) * (W + 1) + (cn.x0 + 1);
                                occupied[pp] = (__ldg(cellbits + (cell >> 5)) >> (cell & 31)) & 1u;
                            }
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
                                const bool ok = occupied[pp] && corner_valid(cn, dx, dy, dz, W, H, D);
                                cw[pp][c] = ok ? corner_weight(cn, dx, dy, dz) : 0.f;
                                coff[pp][c] = ok ? (uint32_t)((((cn.z0 + dz) * H + (cn.y0 + dy)) * W + (cn.x0 + dx)) * C) : 0u;
                            }
                        }
                    }
#pragma unroll
                    for (int pp = 0; pp < PTS_PER_GROUP; ++pp) {
                        const int p = grp + 64 * pp;
                        float a[4] = {0.f, 0.f, 0.f, 0.f};
                        if (occupied[pp]) {
                            typename Quad<VT>::raw v[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                v[c] = (cw[pp][c] != 0.f) ? Quad<VT>::load(vol + coff[pp][c] + c0 + 4 * t) : Quad<VT>::zero();
#pragma unroll
                            for (int c = 0; c < 8; ++c) Quad<VT>::fma(a, v[c], cw[pp][c]);   // ATen order: x fastest
                        }
                        uint2 hi;
                        hi.x = tc::cvt_f16x2(a[0], a[1]); hi.y = tc::cvt_f16x2(a[2], a[3]);
                        const int so = ((uu * 4 + (t >> 1)) * 16 + (p >> 3)) * 128 + (p & 7) * 16 + (t & 1) * 8;
                        *reinterpret_cast<uint2*>(hi_plane + so) = hi;
                        if (NP == 3) {
                            uint2 lo;
                            lo.x = tc::cvt_f16x2(f16lo_of(a[0], hi.x, 0), f16lo_of(a[1], hi.x, 1));
                            lo.y = tc::cvt_f16x2(f16lo_of(a[2], hi.y, 0), f16lo_of(a[3], hi.y, 1));
                            *reinterpret_cast<uint2*>(lo_plane + so) = lo;
                        }
                    }
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars[BAR_SEG_FULL + buf]);
                tr.ev(20 + seg);                                 // segment gathered (this warp)
            }
        }
    }
    // ================================================================== LOADER: weight stream -> ring
    else if (warp == LOAD_WARP) {
        if (lane == 0) {
            uint32_t cnt = 0;
            const unsigned char* seq = reinterpret_cast<const unsigned char*>(P.wf16);
            // one ring slot = one group: up to two source pieces (the per-frame step of layer 3 lives elsewhere)
            auto push = [&](const unsigned char* src, uint32_t bytes, const unsigned char* src2 = nullptr, uint32_t bytes2 = 0) {
                const uint32_t slot = cnt % NUM_SLOTS, round = cnt / NUM_SLOTS;
                unsigned char* dst = smem + OFF_RING + slot * SLOT_BYTES;
                tc::mbar_wait(&bars[BAR_W_EMPTY + slot], (round & 1) ^ 1);     // every CTA of the cluster has consumed it
                tc::mbar_arrive_expect_tx(&bars[BAR_W_FULL + slot], bytes + bytes2);   // arm OUR barrier (a peer may issue the copy)
                if (CLUSTER == 1) {
                    tc::bulk_g2s(dst, src, bytes, &bars[BAR_W_FULL + slot]);
                    if (bytes2) tc::bulk_g2s(dst + bytes, src2, bytes2, &bars[BAR_W_FULL + slot]);
                } else if (cnt % CLUSTER == crank) {
                    tc::bulk_g2s_multicast(dst, src, bytes, &bars[BAR_W_FULL + slot], CMASK);
                    if (bytes2) tc::bulk_g2s_multicast(dst + bytes, src2, bytes2, &bars[BAR_W_FULL + slot], CMASK);
                }
                ++cnt;
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int b = tile < P.n_groups ? tile / P.groups_per_frame : 0;
                // layers 0-2: per group [hi steps][lo steps] (the 1-pass mode skips the lo halves), then the bias step
                for (int layer = 0; layer < 3; ++layer) {
                    const int nks = layer == 0 ? kKsL0 : kKsL12;
                    const unsigned char* base = seq + 2 * (layer == 0 ? sL0 : layer == 1 ? sL1 : sL2);
                    for (int g0 = 0; g0 < nks; g0 += 4) {
                        const int gs = (nks - g0) < 4 ? (nks - g0) : 4;
                        push(base + 2 * step256_offset(g0, 0, nks), gs * STEP_BYTES);
                        if (NP == 3) push(base + 2 * step256_offset(g0, 1, nks), gs * STEP_BYTES);
                    }
                    push(base + 2 * bias256_offset(nks), STEP_BYTES);
                }
                {   // layer 3: 16 h2 steps in 4 groups, then the 6 per-point-tile steps as 4 + (1 common + 1 per-frame)
                    const uint32_t sb = kStepHalves3 * 2;
                    const unsigned char* l3 = seq + sL3 * 2;
                    for (int g0 = 0; g0 < 20; g0 += 4) push(l3 + (size_t)g0 * sb, 4 * sb);
                    push(l3 + (size_t)20 * sb, sb, reinterpret_cast<const unsigned char*>(P.wframe) + (size_t)b * sb, sb);
                }
                push(seq + sL4 * 2, kStepsL4 * kStepHalves4 * 2);   // layer 4: all 9 steps in one group
            }
        }
    }
    // ================================================================== MMA ISSUER
    else if (warp == MMA_WARP) {
        if (lane == 0) {
            uint32_t cnt = 0, hcnt = 0;
            int it = 0;
            Tracer tr;
            tr.init(P.trace, 1);
            const uint32_t seg_addr = tc::smem_u32(smem + OFF_SEG), pe_addr = tc::smem_u32(smem + OFF_PE);
            const uint32_t ones_addr = tc::smem_u32(smem + OFF_ONES), ring_addr = tc::smem_u32(smem + OFF_RING);
            constexpr uint32_t ID256 = tc::make_idesc_f16(128, 256), ID3 = tc::make_idesc_f16(128, kN3),
                               ID4 = tc::make_idesc_f16(128, kN4);
            auto wait_slot = [&](uint32_t& slot) {
                slot = cnt % NUM_SLOTS;
                tc::mbar_wait(&bars[BAR_W_FULL + slot], (cnt / NUM_SLOTS) & 1);
                tc::tc_fence_after();
            };
            auto release_slot = [&](uint32_t slot) {
                if (CLUSTER == 1) tc::mma_commit(&bars[BAR_W_EMPTY + slot]);
                else tc::mma_commit_multicast(&bars[BAR_W_EMPTY + slot], CMASK);
                ++cnt;
            };
            auto a_desc = [&](uint32_t base, int ks) { return tc::make_smem_desc(base + ks * 2 * CHUNK_BYTES, CHUNK_BYTES, 128); };
            // B operand = step i of the group held by `slot` (steps are N*32 bytes apart)
            auto b_desc = [&](uint32_t slot, int i, int N) {
                return tc::make_smem_desc(ring_addr + slot * SLOT_BYTES + i * N * 32, N * 16, 128);
            };
            auto wait_h = [&]() { tc::mbar_wait(&bars[BAR_H_READY], hcnt & 1); ++hcnt; tc::tc_fence_after(); };

            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                uint32_t slot;
                if (it > 0) wait_h();                 // previous tile's last epilogue has drained the accumulator
                tr.ev(1);                             // tile begin
                // ---- layer 0: A = gathered features, K-pipelined: segment = weight group = 4 K-steps
                for (int seg = 0; seg < NUM_SEGS; ++seg) {
                    const uint32_t gseg = (uint32_t)it * NUM_SEGS + seg;
                    const uint32_t buf = gseg % NUM_SEG_BUFS;
                    tc::mbar_wait(&bars[BAR_SEG_FULL + buf], (gseg / NUM_SEG_BUFS) & 1);
                    tc::tc_fence_after();
                    tr.ev(10 + seg);                  // segment available
                    const uint32_t hi_addr = seg_addr + buf * SEG_BYTES, lo_addr = hi_addr + SEG_CHUNKS * CHUNK_BYTES;
                    const int nks = (seg == NUM_SEGS - 1) ? 2 : 4;
                    wait_slot(slot);
                    for (int ks = 0; ks < nks; ++ks) {
                        tc::mma_ss(tmem + TM_ACC, a_desc(hi_addr, ks), b_desc(slot, ks, 256), ID256, (seg | ks) != 0);
                        if (NP == 3) tc::mma_ss(tmem + TM_ACC, a_desc(lo_addr, ks), b_desc(slot, ks, 256), ID256, true);
                    }
                    release_slot(slot);
                    if (NP == 3) {
                        wait_slot(slot);
                        for (int ks = 0; ks < nks; ++ks)
                            tc::mma_ss(tmem + TM_ACC, a_desc(hi_addr, ks), b_desc(slot, ks, 256), ID256, true);
                        release_slot(slot);
                    }
                    tc::mma_commit(&bars[BAR_SEG_EMPTY + buf]);
                }
                wait_slot(slot);
                tc::mma_ss(tmem + TM_ACC, a_desc(ones_addr, 0), b_desc(slot, 0, 256), ID256, true);
                release_slot(slot);
                tc::mma_commit(&bars[BAR_ACC_FULL]);
                tr.ev(20);                            // layer 0 issued
                // ---- layers 1, 2: A = h (TMEM, in place), 4 groups of 4 K-steps (x passes) + ones step
                for (int layer = 1; layer <= 2; ++layer) {
                    wait_h();
                    tr.ev(30 + layer);                // h ready for this layer
                    for (int g0 = 0; g0 < 16; g0 += 4) {
                        wait_slot(slot);
                        for (int i = 0; i < 4; ++i) {
                            tc::mma_ts(tmem + TM_ACC, tmem + TM_HI + (g0 + i) * 8, b_desc(slot, i, 256), ID256, (g0 | i) != 0);
                            if (NP == 3) tc::mma_ts(tmem + TM_ACC, tmem + TM_LO + (g0 + i) * 8, b_desc(slot, i, 256), ID256, true);
                        }
                        release_slot(slot);
                        if (NP == 3) {
                            wait_slot(slot);
                            for (int i = 0; i < 4; ++i)
                                tc::mma_ts(tmem + TM_ACC, tmem + TM_HI + (g0 + i) * 8, b_desc(slot, i, 256), ID256, true);
                            release_slot(slot);
                        }
                    }
                    wait_slot(slot);
                    tc::mma_ss(tmem + TM_ACC, a_desc(ones_addr, 0), b_desc(slot, 0, 256), ID256, true);
                    release_slot(slot);
                    tc::mma_commit(&bars[BAR_ACC_FULL]);
                    tr.ev(20 + layer);                // layer issued
                }
                // ---- layer 3: N = 144: A = h2 (hi [+ lo]), then the per-point tile (PE | ones) from smem
                wait_h();
                tr.ev(33);
                for (int g0 = 0; g0 < 16; g0 += 4) {
                    wait_slot(slot);
                    for (int i = 0; i < 4; ++i) {
                        tc::mma_ts(tmem + TM_ACC, tmem + TM_HI + (g0 + i) * 8, b_desc(slot, i, kN3), ID3, (g0 | i) != 0);
                        // the lo half of the activations only matters on the density path: rows 128..143 of the step (alpha_fc
                        // hi / lo + padding) -> accumulator columns 128..143; the 128 colour columns take the hi half alone
                        if (NP == 3)
                            tc::mma_ts(tmem + TM_ACC + 128, tmem + TM_LO + (g0 + i) * 8,
                                       tc::make_smem_desc(ring_addr + slot * SLOT_BYTES + i * kN3 * 32 + 128 * 16, kN3 * 16, 128), ID4, true);
                    }
                    release_slot(slot);
                }
                wait_slot(slot);
                for (int i = 0; i < 4; ++i) tc::mma_ss(tmem + TM_ACC, a_desc(pe_addr, i), b_desc(slot, i, kN3), ID3, true);
                release_slot(slot);
                wait_slot(slot);
                for (int i = 0; i < 2; ++i) tc::mma_ss(tmem + TM_ACC, a_desc(pe_addr, 4 + i), b_desc(slot, i, kN3), ID3, true);
                release_slot(slot);
                tc::mma_commit(&bars[BAR_ACC_FULL]);
                tr.ev(23);
                // ---- layer 4: N = 16: A = relu(w) (fp16 in h_hi[0:64), K = 128) + ones step, one group
                wait_h();
                tr.ev(34);
                wait_slot(slot);
                for (int ks = 0; ks < 8; ++ks)
                    tc::mma_ts(tmem + TM_ACC, tmem + TM_HI + ks * 8, b_desc(slot, ks, kN4), ID4, ks > 0);
                tc::mma_ss(tmem + TM_ACC, a_desc(ones_addr, 0), b_desc(slot, 8, kN4), ID4, true);
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
        Tracer tr;
        tr.init(tid == 0 ? P.trace : nullptr, 2);
        auto wait_acc = [&]() { tc::mbar_wait(&bars[BAR_ACC_FULL], acnt & 1); ++acnt; tc::tc_fence_after(); };
        // accumulator columns [0, ncols) -> relu -> fp16 hi (+ lo) pairs -> TMEM h (in place: the layer's MMAs are done)
        // (software-pipelined: the tcgen05.ld of granule c+1 is in flight while granule c is converted and stored)
        auto relu_to_h = [&](int ncols, bool with_lo) {
            const int ng = ncols / 8;
            uint32_t va[8], vb[8];
            auto convert_store = [&](const uint32_t (&v)[8], int c) {
                uint32_t h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = tc::cvt_relu_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                tc::tmem_st4(lane_base + TM_HI + c * 4, h);
                if (NP == 3 && with_lo) {
                    uint32_t l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        l[i] = tc::cvt_f16x2(f16lo_of(fmaxf(__uint_as_float(v[2 * i]), 0.f), h[i], 0),
                                             f16lo_of(fmaxf(__uint_as_float(v[2 * i + 1]), 0.f), h[i], 1));
                    tc::tmem_st4(lane_base + TM_LO + c * 4, l);
                }
            };
            tc::tmem_ld8(lane_base + TM_ACC, va);
            tc::tmem_ld_wait();
            for (int c = 0; c < ng; c += 2) {
                tc::tmem_ld8(lane_base + TM_ACC + (c + 1) * 8, vb);          // ng is even
                convert_store(va, c);
                tc::tmem_ld_wait();
                if (c + 2 < ng) tc::tmem_ld8(lane_base + TM_ACC + (c + 2) * 8, va);
                convert_store(vb, c + 1);
                tc::tmem_ld_wait();
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
            tc::mbar_wait(&bars[BAR_GEOM_FULL], it & 1);
            tr.ev(1);                                  // geometry available
            const float4 gm = geom[row];
            zbuf[row] = gm.w;
            tc::mbar_arrive(&bars[BAR_GEOM_FREE]);
            // ---- per-point tile of layer 3: [PE10(world xyz) 63 | 0 | PE4(viewdir) 27 | 0 | 1 | 1 | 0 | 0]
            // (written while the producers gather; read by the MMA only after three more h_ready hand-offs;
            //  the previous tile's layer-3 MMAs were complete before its last accumulator hand-off)
            {
                __half* peh = reinterpret_cast<__half*>(PE);
                // element (row, k) of the 96-wide tile: chunk k/8, 8x8 core matrix (row/8), row%8, k%8
                auto put = [&](int k, float v) {
                    peh[((k >> 3) * 16 + (row >> 3)) * 64 + (row & 7) * 8 + (k & 7)] = __float2half_rn(v);
                };
                positional_embed_anchored<10, 5>(gm.x, gm.y, gm.z, [&](int j, float v) { put(j, v); });
                put(63, 0.f);
                const float dx = __ldg(P.ray_d + ri * 3), dy = __ldg(P.ray_d + ri * 3 + 1), dz = __ldg(P.ray_d + ri * 3 + 2);
                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                positional_embed_anchored<4, 4>(__fdiv_rn(dx, nrm), __fdiv_rn(dy, nrm), __fdiv_rn(dz, nrm), [&](int j, float v) { put(64 + j, v); });
                put(91, 0.f); put(92, 1.f); put(93, 1.f); put(94, 0.f); put(95, 0.f);
                tc::fence_proxy_async();
            }
            // ---- layers 0, 1, 2 -> h (hi [+ lo]) in place
            tr.ev(2);                                  // PE tile written
            for (int layer = 0; layer < 3; ++layer) {
                wait_acc();
                tr.ev(10 + layer);                     // accumulator of this layer complete
                relu_to_h(256, true);
                h_done();
                tr.ev(20 + layer);                     // epilogue of this layer done
            }
            // ---- layer 3: sigma = acc[128] + acc[129]; colour hidden -> fp16 in h_hi[0:64)
            wait_acc();
            tr.ev(13);
            float sigma;
            {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + TM_ACC + 128, v);
                tc::tmem_ld_wait();
                sigma = __uint_as_float(v[0]) + __uint_as_float(v[1]);
            }
            relu_to_h(128, false);
            h_done();
            tr.ev(23);
            // ---- layer 4: rgb logits = hi rows + lo rows
            wait_acc();
            tr.ev(14);
            {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + TM_ACC, v);
                tc::tmem_ld_wait();
                rawbuf[row] = make_float4(__uint_as_float(v[0]) + __uint_as_float(v[3]), __uint_as_float(v[1]) + __uint_as_float(v[4]),
                                          __uint_as_float(v[2]) + __uint_as_float(v[5]), sigma);
            }
            h_done();                                  // accumulator drained: the next tile's layer 0 may start
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
            named_bar_sync(2, EPI_WARPS * 32);     // rawbuf / zbuf are reused by the next tile
            tr.ev(30);                                 // composite done
        }
    }

    // ------------------------------------------------------------------ teardown
    tc::tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) tc::cluster_sync_all();      // no CTA exits while a peer may still multicast into it
    if (warp == MMA_WARP) {
        __syncwarp();
        tc::tmem_dealloc<512>(tmem);
    }
}

template <int NP, typename VT>
static cudaError_t launch(const RenderParams& p, int grid, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(render_tc_kernel<NP, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, render_tc_kernel<NP, VT>, p);
}

}  // namespace tcr

bool tc_available() { return true; }

int launch_render_tc(const RenderParams& p_in, int volume_dtype, int passes, cudaStream_t stream) {
    RenderParams p = p_in;
    const int S = p.n_samples;
    if (S > tcr::TP) {
        set_error("the tensor-core render kernel supports n_samples <= 128 (got %d); use NB_PRECISION_FP32", S);
        return NB_ERR_UNSUPPORTED;
    }
    p.rays_per_group = tcr::TP / S;
    p.tiles_per_group = 1;
    p.groups_per_frame = (p.n_rays + p.rays_per_group - 1) / p.rays_per_group;
    p.n_groups = p.groups_per_frame * p.batch;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.n_groups < sms ? p.n_groups : sms;
    if (p.n_groups == 0) return NB_OK;
    grid = (grid + tcr::CLUSTER - 1) / tcr::CLUSTER * tcr::CLUSTER;     // whole clusters; padding tiles are no-ops
    if (grid > sms) grid = sms / tcr::CLUSTER * tcr::CLUSTER;
    cudaError_t e;
    if (passes == 3) e = (volume_dtype == NB_DTYPE_F32) ? tcr::launch<3, float>(p, grid, stream) : tcr::launch<3, __half>(p, grid, stream);
    else e = (volume_dtype == NB_DTYPE_F32) ? tcr::launch<1, float>(p, grid, stream) : tcr::launch<1, __half>(p, grid, stream);
    if (e != cudaSuccess) { set_error("render_tc launch failed: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
    return NB_OK;
}

}  // namespace nb
