// Tensor-core fused render kernel with EXACT empty-sample skipping (nb_render_args.skip_empty).
//
// Same tile pipeline as nb_render_tc.cu (producers -> K-segmented layer-0 operand, weight-stream ring,
// tcgen05 MMAs with TMEM-resident activations, epilogue warps), but the 128 rows of a tile are no longer
// "two whole rays": a CTA takes a BLOCK of rays (<= 1024 samples), classifies every sample with the
// cell-occupancy bitmaps built by nb_pack_volume, and packs only the occupied samples into tiles.
//
// Why this is exact.  A sample whose four trilinear cells are all unoccupied interpolates features that are
// exactly 0 (SparseConvNet's .dense() is exactly 0 off the active set), so its density is the per-frame
// constant sigma_empty = alpha_fc(relu(fc_2(relu(fc_1(relu(b_0)))))), computed once by nb_pack_weights.  When
// sigma_empty < 0 (any model with empty space), alpha = 1 - exp(-relu(sigma) * dist) is exactly 0, the sample's
// compositing weight is exactly 0, and its colour never reaches rgb_map / depth_map / acc_map: skipping its MLP
// evaluation changes no output bit (nerf_net_utils.py:27-43).  If sigma_empty is not robustly negative the
// kernel classifies every sample as occupied and degenerates to the dense evaluation.
//
// The weight stream cannot be shared by a CTA pair here (tile counts are data-dependent), so this kernel runs
// without clusters.  Roofline accounting: only executed tiles count (nb_render_args.stats).
#include "nb_tc_common.cuh"

namespace nb {
namespace tcs {

using tcr::Quad;
using tcr::Tracer;
using tcr::named_bar_sync;
using tcr::f16lo_of;

constexpr int TP = 128;
constexpr int NUM_SLOTS = 3;
constexpr int STEP_BYTES = 8192;
constexpr int SLOT_BYTES = 4 * STEP_BYTES;
constexpr int CHUNK_BYTES = 2048;
constexpr int SEG_CHUNKS = 8;
constexpr int NUM_SEGS = 6;
constexpr int SEG_RING_BYTES = 4 * SEG_CHUNKS * CHUNK_BYTES;      // 64 KB: 2 x (hi+lo) or 4 x hi
constexpr int MAX_SEG_BUFS = 4;
constexpr int PE_CHUNKS = 12;
constexpr int EPI_WARPS = 4, MMA_WARP = 4, LOAD_WARP = 5, PROD_WARP0 = 6, PROD_WARPS = 16;
constexpr int NT = (PROD_WARP0 + PROD_WARPS) * 32;               // 704
constexpr int PROD_THREADS = PROD_WARPS * 32;                     // 512
constexpr int PTS_PER_GROUP = TP / (PROD_WARPS * 4);
constexpr int MAXS = 1024;                                         // samples per ray block (2 classification passes of 512)

// shared-memory map (bytes)
constexpr int OFF_SEG = 0;
constexpr int OFF_ONES = OFF_SEG + SEG_RING_BYTES;
constexpr int OFF_PE = OFF_ONES + 2 * CHUNK_BYTES;
constexpr int OFF_RING = OFF_PE + PE_CHUNKS * CHUNK_BYTES;
constexpr int OFF_LIST = OFF_RING + NUM_SLOTS * SLOT_BYTES;        // float4[1024] compact list: (wx,wy,wz,sample) producer-owned
constexpr int OFF_RAWB = OFF_LIST + MAXS * 16;                     // float4[1024] (r,g,b,sigma) per block sample epilogue-owned
constexpr int OFF_ZB = OFF_RAWB + MAXS * 16;                       // float[1024] z per block sample              epilogue-owned
constexpr int OFF_ROWS = OFF_ZB + MAXS * 4;                        // int[128] block sample of each tile row      epilogue-owned
constexpr int OFF_XF = OFF_ROWS + TP * 4;                          // FrameXf
constexpr int OFF_INFO = OFF_XF + 128;                             // TileInfo[2]
constexpr int OFF_WCNT = OFF_INFO + 64;                            // int[16] per-warp occupied counts
constexpr int OFF_BAR = OFF_WCNT + 64;
enum { BAR_W_FULL = 0, BAR_W_EMPTY = NUM_SLOTS, BAR_SEG_FULL = 2 * NUM_SLOTS, BAR_SEG_EMPTY = 2 * NUM_SLOTS + MAX_SEG_BUFS,
       BAR_ACC_FULL = 2 * NUM_SLOTS + 2 * MAX_SEG_BUFS, BAR_H_READY, BAR_MSG_FULL, BAR_MSG_FREE, NUM_BARS };
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

constexpr uint32_t TM_ACC = 0, TM_HI = 256, TM_LO = 384;

// one message from the producers to the MMA / loader / epilogue roles
struct TileInfo {
    int nrows;       // occupied samples in this tile (0 = block without any tile: composite only)
    int list_off;    // first entry of the tile in the block's compact list
    int blk;         // ray block id
    int flags;       // bit0 first message of the block, bit1 last message of the block, bit2 done (no more work)
};

struct BlockCoord { int b, r0, nr; };
__device__ __forceinline__ BlockCoord block_coord(const RenderParams& P, int blk) {
    BlockCoord c;
    c.b = blk / P.groups_per_frame;
    c.r0 = (blk % P.groups_per_frame) * P.rays_per_group;
    c.nr = min(P.rays_per_group, P.n_rays - c.r0);
    return c;
}

template <int NP, typename VT>
__global__ void __launch_bounds__(NT, 1) render_tc_sparse_kernel(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    volatile TileInfo* info = reinterpret_cast<volatile TileInfo*>(smem + OFF_INFO);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = P.n_samples;
    constexpr int NUM_SEG_BUFS = (NP == 3) ? 2 : 4;
    constexpr int SEG_BYTES = SEG_RING_BYTES / NUM_SEG_BUFS;

    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == LOAD_WARP * 32) {
        for (int i = 0; i < NUM_SLOTS; ++i) { tc::mbar_init(&bars[BAR_W_FULL + i], 1); tc::mbar_init(&bars[BAR_W_EMPTY + i], 1); }
        for (int i = 0; i < NUM_SEG_BUFS; ++i) { tc::mbar_init(&bars[BAR_SEG_FULL + i], PROD_WARPS); tc::mbar_init(&bars[BAR_SEG_EMPTY + i], 1); }
        tc::mbar_init(&bars[BAR_ACC_FULL], 1);
        tc::mbar_init(&bars[BAR_H_READY], EPI_WARPS * 32);
        tc::mbar_init(&bars[BAR_MSG_FULL], PROD_WARPS);
        tc::mbar_init(&bars[BAR_MSG_FREE], EPI_WARPS * 32 + 2);       // epilogue threads + MMA thread + loader thread
        tc::fence_mbar_init();
    }
    if (warp >= PROD_WARP0) {
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
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_blocks = P.n_groups;
    const float sigma_empty = __ldg(P.wf32 + oSigmaEmpty);
    const bool can_skip = sigma_empty < -1e-3f;      // robustly negative => empty samples have weight exactly 0

    // ================================================================== PRODUCERS
    if (warp >= PROD_WARP0) {
        const int pt = tid - PROD_WARP0 * 32;          // 0..511 = sample of the block during classification
        const int pw = warp - PROD_WARP0;
        float4* lst = reinterpret_cast<float4*>(smem + OFF_LIST);
        int* wcnt = reinterpret_cast<int*>(smem + OFF_WCNT);
        FrameXf* xf = reinterpret_cast<FrameXf*>(smem + OFF_XF);
        const unsigned char* volbase = reinterpret_cast<const unsigned char*>(P.volume);
        const uint32_t* occ_base = reinterpret_cast<const uint32_t*>(volbase);
        const int grp = pw * 4 + (lane >> 3);
        const int t = lane & 7;
        uint32_t msg = 0, it = 0;                      // messages sent, real tiles sent
        Tracer tr;
        tr.init((pw == 0 && lane == 0) ? P.trace : nullptr, 0);
        unsigned long long n_tiles_done = 0, n_occ = 0;
        auto publish = [&](int nrows, int off, int blk, int flags) {
            tc::mbar_wait(&bars[BAR_MSG_FREE], (msg & 1) ^ 1);        // everybody has read the previous message
            if (pt == 0) {
                volatile TileInfo* ti = &info[msg & 1];
                ti->nrows = nrows; ti->list_off = off; ti->blk = blk; ti->flags = flags;
            }
            named_bar_sync(1, PROD_THREADS);
            if (lane == 0) tc::mbar_arrive(&bars[BAR_MSG_FULL]);
            ++msg;
        };
        for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
            const BlockCoord bc = block_coord(P, blk);
            tr.ev(2);                                   // block begin
            // the previous block's list may be overwritten once the epilogue has copied its last tile's rows
            tc::mbar_wait(&bars[BAR_MSG_FREE], (msg & 1) ^ 1);
            if (pt < 9) xf->R[pt] = __ldg(P.R + bc.b * 9 + pt);
            if (pt < 3) {
                xf->Th[pt] = __ldg(P.Th + bc.b * 3 + pt);
                xf->min_dhw[pt] = __ldg(P.bounds + bc.b * 6 + (2 - pt));
                xf->voxel[pt] = P.voxel_size[pt];
                xf->out_sh[pt] = P.out_sh[pt];
            }
            named_bar_sync(1, PROD_THREADS);
            // ---- classify (one sample per thread, MAXS / 512 passes) + order-preserving compaction of the occupied ones
            int total = 0;
            for (int pass = 0; pass < MAXS / PROD_THREADS; ++pass) {
                // SAMPLE-major enumeration: consecutive list entries are the same depth sample of neighbouring rays, which
                // cross the same cells at every level, so a tile's rows share most of their corner lines in L1
                const int j = pass * PROD_THREADS + pt;
                const int ry = j % P.rays_per_group, s = j / P.rays_per_group;
                bool occ = false;
                float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ry < bc.nr && s < S) {
                    const size_t ri = (size_t)bc.b * P.n_rays + bc.r0 + ry;
                    const float ox = __ldg(P.ray_o + ri * 3), oy = __ldg(P.ray_o + ri * 3 + 1), oz = __ldg(P.ray_o + ri * 3 + 2);
                    const float dx = __ldg(P.ray_d + ri * 3), dy = __ldg(P.ray_d + ri * 3 + 1), dz = __ldg(P.ray_d + ri * 3 + 2);
                    const float z = z_sample(__ldg(P.near + ri), __ldg(P.far + ri), P.t_vals, s, S, P.t_rand ? P.t_rand + ri * S : nullptr, P.z_user ? P.z_user + ri * S : nullptr);
                    gm.x = __fadd_rn(ox, __fmul_rn(dx, z));
                    gm.y = __fadd_rn(oy, __fmul_rn(dy, z));
                    gm.z = __fadd_rn(oz, __fmul_rn(dz, z));
                    float gx, gy, gz;
                    world_to_grid(*xf, gm.x, gm.y, gm.z, gx, gy, gz);
                    // per-level cell occupancy (bit l = the sample's level-l cell holds a non-zero voxel); the gather reuses it
                    uint32_t lm = 0;
                    const bool inside = P.mask_nv == 0 || inside_masks(P, gm.x, gm.y, gm.z);   // f-1 mask views
                    if (inside) {
#pragma unroll
                        for (int lvl = 0; lvl < 4; ++lvl) {
                            const int D = P.lvl_D[lvl], H = P.lvl_H[lvl], W = P.lvl_W[lvl];
                            Corners cn;
                            corner_setup(unnormalize(gx, W), unnormalize(gy, H), unnormalize(gz, D), W, H, D, cn);
                            if (cn.x0 != -2) {
                                const uint32_t* cellbits = occ_base + P.occ_off[lvl] / 4 + (size_t)bc.b * P.occ_bstride[lvl];
                                const uint32_t cell = ((uint32_t)(cn.z0 + 1) * (H + 1) + (cn.y0 + 1)) * (W + 1) + (cn.x0 + 1);
                                lm |= ((__ldg(cellbits + (cell >> 5)) >> (cell & 31)) & 1u) << lvl;
                            }
                        }
                    }
                    occ = inside && (lm != 0u || !can_skip);
                    gm.w = __int_as_float((ry * S + s) | (int)(lm << 16));   // block sample id (ray-major, as rawb / zb) | level bits
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, occ);
                if (lane == 0) wcnt[pw] = __popc(bal);
                named_bar_sync(1, PROD_THREADS);
                int base = total;
#pragma unroll
                for (int w = 0; w < PROD_WARPS; ++w) { const int c = wcnt[w]; base += (w < pw) ? c : 0; total += c; }
                if (occ) lst[base + __popc(bal & ((1u << lane) - 1))] = gm;
                named_bar_sync(1, PROD_THREADS);
            }
            n_occ += (pt == 0) ? total : 0;
            const int ntile = (total + TP - 1) / TP;
            tr.ev(3);                                   // block classified
            if (ntile == 0) publish(0, 0, blk, 3);
            for (int tl = 0; tl < ntile; ++tl) {
                const int off = tl * TP;
                const int nrows = min(TP, total - off);
                publish(nrows, off, blk, (tl == 0 ? 1 : 0) | (tl == ntile - 1 ? 2 : 0));
                tr.ev(1);                               // tile published
                // ---- gather this tile (same unit mapping as the dense kernel; rows >= nrows are zero)
                float4 g[PTS_PER_GROUP];
#pragma unroll
                for (int pp = 0; pp < PTS_PER_GROUP; ++pp) {
                    const int row = grp + 64 * pp;
                    g[pp] = make_float4(-4.f, -4.f, -4.f, 0.f);
                    if (row < nrows) {
                        const float4 e = lst[off + row];
                        world_to_grid(*xf, e.x, e.y, e.z, g[pp].x, g[pp].y, g[pp].z);
                        g[pp].w = __int_as_float(__float_as_int(e.w) >> 16);     // per-level occupancy bits
                    }
                }
                uint32_t coff[PTS_PER_GROUP][8];
                float cw[PTS_PER_GROUP][8];
                bool occupied[PTS_PER_GROUP];
                const VT* vol = nullptr;
                int cur_lvl = -1;
                const uint32_t seg_base = tc::smem_u32(smem + OFF_SEG);
                const uint32_t so0 = (uint32_t)((((t >> 1) * 16 + (grp >> 3)) * 128) + (grp & 7) * 16 + (t & 1) * 8);
                for (int seg = 0; seg < NUM_SEGS; ++seg) {
                    const uint32_t gseg = it * NUM_SEGS + seg;
                    const uint32_t buf = gseg % NUM_SEG_BUFS;
                    tc::mbar_wait(&bars[BAR_SEG_EMPTY + buf], ((gseg / NUM_SEG_BUFS) & 1) ^ 1);
                    tr.ev(10 + seg);
                    // this thread's (row, channel quad) slot of the hi plane; the lo plane follows SEG_CHUNKS chunks later
                    const uint32_t dst = seg_base + buf * SEG_BYTES + so0;
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
                            vol = reinterpret_cast<const VT*>(volbase + P.lvl_off[lvl]) + (size_t)bc.b * P.lvl_bstride[lvl];
                            // lane t of the row's 8-lane group sets up corner t; the group exchanges the 8 (weight, offset) pairs
                            const int ddx = t & 1, ddy = (t >> 1) & 1, ddz = t >> 2;
#pragma unroll
                            for (int pp = 0; pp < PTS_PER_GROUP; ++pp) {
                                Corners cn;
                                corner_setup(unnormalize(g[pp].x, W), unnormalize(g[pp].y, H), unnormalize(g[pp].z, D), W, H, D, cn);
                                occupied[pp] = (__float_as_int(g[pp].w) >> lvl) & 1;
                                const bool ok = occupied[pp] && corner_valid(cn, ddx, ddy, ddz, W, H, D);
                                const float w_own = ok ? __fmul_rn(__fmul_rn(ddx ? cn.wx[1] : cn.wx[0], ddy ? cn.wy[1] : cn.wy[0]),
                                                                   ddz ? cn.wz[1] : cn.wz[0]) : 0.f;
                                // byte offset of the corner vector inside this frame's level; corners that do not contribute
                                // (out of range, zero weight, unoccupied cell) point at voxel 0 and are never accumulated
                                const uint32_t o_own = ok ? (uint32_t)((((cn.z0 + ddz) * H + (cn.y0 + ddy)) * W + (cn.x0 + ddx)) * C) * (uint32_t)sizeof(VT) : 0u;
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    cw[pp][c] = __shfl_sync(0xffffffffu, w_own, c, 8);
                                    coff[pp][c] = __shfl_sync(0xffffffffu, o_own, c, 8) + (uint32_t)(4 * t * sizeof(VT));
                                }
                            }
                        }
#pragma unroll
                        for (int pp = 0; pp < PTS_PER_GROUP; ++pp) {
                            float a[4] = {0.f, 0.f, 0.f, 0.f};
                            if (occupied[pp]) {
                                const unsigned char* ub = reinterpret_cast<const unsigned char*>(vol + c0);   // warp-uniform
                                typename Quad<VT>::raw v[8];
#pragma unroll
                                for (int c = 0; c < 8; ++c) v[c] = Quad<VT>::load_bytes(ub + coff[pp][c]);
#pragma unroll
                                for (int c = 0; c < 8; ++c)
                                    if (cw[pp][c] != 0.f) Quad<VT>::fma(a, v[c], cw[pp][c]);
                            }
                            uint2 hi;
                            hi.x = tc::cvt_f16x2(a[0], a[1]); hi.y = tc::cvt_f16x2(a[2], a[3]);
                            const uint32_t so = dst + (uint32_t)(uu * 4 * 16 * 128 + pp * 8 * 128);   // K-major core-matrix layout
                            tcr::sts_v2(so, hi);
                            if (NP == 3) {
                                uint2 lo;
                                lo.x = tc::cvt_f16x2(f16lo_of(a[0], hi.x, 0), f16lo_of(a[1], hi.x, 1));
                                lo.y = tc::cvt_f16x2(f16lo_of(a[2], hi.y, 0), f16lo_of(a[3], hi.y, 1));
                                tcr::sts_v2(so + SEG_CHUNKS * CHUNK_BYTES, lo);
                            }
                        }
                    }
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bars[BAR_SEG_FULL + buf]);
                    tr.ev(20 + seg);
                }
                ++it;
                ++n_tiles_done;
            }
        }
        publish(0, 0, 0, 4);                            // done
        if (pt == 0 && P.stats) {
            atomicAdd(P.stats + 0, n_tiles_done);
            atomicAdd(P.stats + 1, n_occ);
        }
    }
    // ================================================================== LOADER
    else if (warp == LOAD_WARP) {
        if (lane == 0) {
            uint32_t cnt = 0, msg = 0;
            const unsigned char* seq = reinterpret_cast<const unsigned char*>(P.wf16);
            auto push = [&](const unsigned char* src, uint32_t bytes, const unsigned char* src2 = nullptr, uint32_t bytes2 = 0) {
                const uint32_t slot = cnt % NUM_SLOTS, round = cnt / NUM_SLOTS;
                unsigned char* dst = smem + OFF_RING + slot * SLOT_BYTES;
                tc::mbar_wait(&bars[BAR_W_EMPTY + slot], (round & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&bars[BAR_W_FULL + slot], bytes + bytes2);
                tc::bulk_g2s(dst, src, bytes, &bars[BAR_W_FULL + slot]);
                if (bytes2) tc::bulk_g2s(dst + bytes, src2, bytes2, &bars[BAR_W_FULL + slot]);
                ++cnt;
            };
            for (;;) {
                tc::mbar_wait(&bars[BAR_MSG_FULL], msg & 1);
                const int nrows = info[msg & 1].nrows, flags = info[msg & 1].flags, blk = info[msg & 1].blk;
                tc::mbar_arrive(&bars[BAR_MSG_FREE]);
                ++msg;
                if (flags & 4) break;
                if (nrows == 0) continue;
                const int b = blk / P.groups_per_frame;
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
                {
                    const uint32_t sb = kStepHalves3 * 2;
                    const unsigned char* l3 = seq + sL3 * 2;
                    for (int g0 = 0; g0 < 20; g0 += 4) push(l3 + (size_t)g0 * sb, 4 * sb);
                    push(l3 + (size_t)20 * sb, sb, reinterpret_cast<const unsigned char*>(P.wframe) + (size_t)b * sb, sb);
                }
                push(seq + sL4 * 2, kStepsL4 * kStepHalves4 * 2);
            }
        }
    }
    // ================================================================== MMA ISSUER
    else if (warp == MMA_WARP) {
        if (lane == 0) {
            uint32_t cnt = 0, hcnt = 0, msg = 0, it = 0;
            const uint32_t seg_addr = tc::smem_u32(smem + OFF_SEG), pe_addr = tc::smem_u32(smem + OFF_PE);
            const uint32_t ones_addr = tc::smem_u32(smem + OFF_ONES), ring_addr = tc::smem_u32(smem + OFF_RING);
            constexpr uint32_t ID256 = tc::make_idesc_f16(128, 256), ID3 = tc::make_idesc_f16(128, kN3),
                               ID4 = tc::make_idesc_f16(128, kN4);
            auto wait_slot = [&](uint32_t& slot) {
                slot = cnt % NUM_SLOTS;
                tc::mbar_wait(&bars[BAR_W_FULL + slot], (cnt / NUM_SLOTS) & 1);
                tc::tc_fence_after();
            };
            auto release_slot = [&](uint32_t slot) { tc::mma_commit(&bars[BAR_W_EMPTY + slot]); ++cnt; };
            auto a_desc = [&](uint32_t base, int ks) { return tc::make_smem_desc(base + ks * 2 * CHUNK_BYTES, CHUNK_BYTES, 128); };
            auto b_desc = [&](uint32_t slot, int i, int N) {
                return tc::make_smem_desc(ring_addr + slot * SLOT_BYTES + i * N * 32, N * 16, 128);
            };
            auto wait_h = [&]() { tc::mbar_wait(&bars[BAR_H_READY], hcnt & 1); ++hcnt; tc::tc_fence_after(); };
            Tracer tr;
            tr.init(P.trace, 1);
            for (;;) {
                tc::mbar_wait(&bars[BAR_MSG_FULL], msg & 1);
                const int nrows = info[msg & 1].nrows, flags = info[msg & 1].flags;
                tc::mbar_arrive(&bars[BAR_MSG_FREE]);
                ++msg;
                if (flags & 4) break;
                if (nrows == 0) continue;
                uint32_t slot;
                if (it > 0) wait_h();
                tr.ev(1);
                for (int seg = 0; seg < NUM_SEGS; ++seg) {
                    const uint32_t gseg = it * NUM_SEGS + seg;
                    const uint32_t buf = gseg % NUM_SEG_BUFS;
                    tc::mbar_wait(&bars[BAR_SEG_FULL + buf], (gseg / NUM_SEG_BUFS) & 1);
                    tc::tc_fence_after();
                    tr.ev(10 + seg);
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
                tr.ev(20);
                for (int layer = 1; layer <= 2; ++layer) {
                    wait_h();
                    tr.ev(30 + layer);
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
                    tr.ev(20 + layer);
                }
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
                wait_h();
                tr.ev(34);
                wait_slot(slot);
                for (int ks = 0; ks < 8; ++ks)
                    tc::mma_ts(tmem + TM_ACC, tmem + TM_HI + ks * 8, b_desc(slot, ks, kN4), ID4, ks > 0);
                tc::mma_ss(tmem + TM_ACC, a_desc(ones_addr, 0), b_desc(slot, 8, kN4), ID4, true);
                release_slot(slot);
                tc::mma_commit(&bars[BAR_ACC_FULL]);
                ++it;
            }
        }
    }
    // ================================================================== EPILOGUE (thread = tile row)
    else {
        const int row = tid;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const float4* lst = reinterpret_cast<const float4*>(smem + OFF_LIST);
        float4* rawb = reinterpret_cast<float4*>(smem + OFF_RAWB);
        float* zb = reinterpret_cast<float*>(smem + OFF_ZB);
        unsigned char* PE = smem + OFF_PE;
        uint32_t acnt = 0, msg = 0;
        auto wait_acc = [&]() { tc::mbar_wait(&bars[BAR_ACC_FULL], acnt & 1); ++acnt; tc::tc_fence_after(); };
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
                tc::tmem_ld8(lane_base + TM_ACC + (c + 1) * 8, vb);
                convert_store(va, c);
                tc::tmem_ld_wait();
                if (c + 2 < ng) tc::tmem_ld8(lane_base + TM_ACC + (c + 2) * 8, va);
                convert_store(vb, c + 1);
                tc::tmem_ld_wait();
            }
            tc::tmem_st_wait();
        };
        auto h_done = [&]() { tc::tc_fence_before(); tc::mbar_arrive(&bars[BAR_H_READY]); };

        for (;;) {
            tc::mbar_wait(&bars[BAR_MSG_FULL], msg & 1);
            const int nrows = info[msg & 1].nrows, flags = info[msg & 1].flags, blk = info[msg & 1].blk, off = info[msg & 1].list_off;
            // copy this row's entry of the producer-owned list, then release the message
            float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
            int smp = -1;
            if (row < nrows) { gm = lst[off + row]; smp = __float_as_int(gm.w) & 0xFFFF; }
            tc::mbar_arrive(&bars[BAR_MSG_FREE]);
            ++msg;
            if (flags & 4) break;
            const BlockCoord bc = block_coord(P, blk);
            const int nsmp = bc.nr * S;
            if (flags & 1) {            // first message of the block: every sample starts as "empty"
                for (int j = row; j < nsmp; j += EPI_WARPS * 32) {
                    rawb[j] = make_float4(0.f, 0.f, 0.f, fminf(sigma_empty, 0.f));   // skipped samples: weight exactly 0
                    const size_t ri = (size_t)bc.b * P.n_rays + bc.r0 + j / S;
                    zb[j] = z_sample(__ldg(P.near + ri), __ldg(P.far + ri), P.t_vals, j % S, S, P.t_rand ? P.t_rand + ri * S : nullptr, P.z_user ? P.z_user + ri * S : nullptr);
                }
                named_bar_sync(2, EPI_WARPS * 32);
            }
            if (nrows > 0) {
                const size_t ri = (size_t)bc.b * P.n_rays + bc.r0 + (smp >= 0 ? smp / S : 0);
                {
                    __half* peh = reinterpret_cast<__half*>(PE);
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
                for (int layer = 0; layer < 3; ++layer) {
                    wait_acc();
                    relu_to_h(256, true);
                    h_done();
                }
                wait_acc();
                float sigma;
                {
                    uint32_t v[16];
                    tc::tmem_ld16(lane_base + TM_ACC + 128, v);
                    tc::tmem_ld_wait();
                    sigma = __uint_as_float(v[0]) + __uint_as_float(v[1]);
                }
                relu_to_h(128, false);
                h_done();
                wait_acc();
                {
                    uint32_t v[16];
                    tc::tmem_ld16(lane_base + TM_ACC, v);
                    tc::tmem_ld_wait();
                    if (smp >= 0)
                        rawb[smp] = make_float4(__uint_as_float(v[0]) + __uint_as_float(v[3]), __uint_as_float(v[1]) + __uint_as_float(v[4]),
                                                __uint_as_float(v[2]) + __uint_as_float(v[5]), sigma);
                }
                h_done();
            }
            if (flags & 2) {            // last message of the block: composite its rays (a10), one warp per ray
                named_bar_sync(2, EPI_WARPS * 32);
                for (int g = warp; g < bc.nr; g += EPI_WARPS) {
                    const size_t rg = (size_t)bc.b * P.n_rays + bc.r0 + g;
                    const float dx = __ldg(P.ray_d + rg * 3), dy = __ldg(P.ray_d + rg * 3 + 1), dz = __ldg(P.ray_d + rg * 3 + 2);
                    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                    float* wout = P.weights ? P.weights + rg * S : nullptr;
                    RayOut o = composite_ray(rawb + g * S, zb + g * S, S, nrm, wout, lane);
                    if (P.raw) {
                        float4* rdst = reinterpret_cast<float4*>(P.raw) + rg * S;
                        for (int s = lane; s < S; s += 32) rdst[s] = rawb[g * S + s];
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
                named_bar_sync(2, EPI_WARPS * 32);
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        tc::tmem_dealloc<512>(tmem);
    }
}

template <int NP, typename VT>
static cudaError_t launch(const RenderParams& p, int grid, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(render_tc_sparse_kernel<NP, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    render_tc_sparse_kernel<NP, VT><<<grid, NT, SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tcs

int launch_render_tc_sparse(const RenderParams& p_in, int volume_dtype, int passes, cudaStream_t stream) {
    RenderParams p = p_in;
    const int S = p.n_samples;
    if (S > tcs::TP) {
        set_error("the tensor-core render kernel supports n_samples <= 128 (got %d); use NB_PRECISION_FP32", S);
        return NB_ERR_UNSUPPORTED;
    }
    p.rays_per_group = tcs::MAXS / S;                 // rays per block (<= 512 samples)
    p.tiles_per_group = 0;
    p.groups_per_frame = (p.n_rays + p.rays_per_group - 1) / p.rays_per_group;
    p.n_groups = p.groups_per_frame * p.batch;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.n_groups < sms ? p.n_groups : sms;
    if (grid == 0) return NB_OK;
    cudaError_t e;
    if (passes == 3) e = (volume_dtype == NB_DTYPE_F32) ? tcs::launch<3, float>(p, grid, stream) : tcs::launch<3, __half>(p, grid, stream);
    else e = (volume_dtype == NB_DTYPE_F32) ? tcs::launch<1, float>(p, grid, stream) : tcs::launch<1, __half>(p, grid, stream);
    if (e != cudaSuccess) { set_error("render_tc_sparse launch failed: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
    return NB_OK;
}

}  // namespace nb
