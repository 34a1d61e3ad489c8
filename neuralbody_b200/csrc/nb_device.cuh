// Device helpers shared by the exact-fp32 and the tcgen05 render kernels:
// sample generation, world->SMPL->grid transform, trilinear corner set-up,
// positional encoding and the alpha-compositing warp scan.
//
// Parity-critical arithmetic follows the reference's op sequence in fp32 with
// explicit round-to-nearest intrinsics (no FMA contraction) where upstream issues
// separate PyTorch ops, so z_vals / grid coordinates agree to the last bit or ulp.
#pragma once
#include "nb_internal.h"

namespace nb {

// ---------------------------------------------------------------- a2: get_sampling_points
// lib/networks/renderer/if_clight_renderer.py:13-14.  torch.linspace's CPU/CUDA kernels
// fill symmetrically: start + step*i below the midpoint, end - step*(steps-1-i) above.
__device__ __forceinline__ float linspace01(int i, int steps) {
    if (steps == 1) return 0.f;
    float step = __fdiv_rn(1.f, (float)(steps - 1));
    return (i < steps / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.f, __fmul_rn(step, (float)(steps - 1 - i)));
}

__device__ __forceinline__ float z_plain(float near, float far, float t) {
    // near[..., None] * (1. - t_vals) + far[..., None] * t_vals
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, t)), __fmul_rn(far, t));
}

// z value of sample s, including the stratified jitter of if_clight_renderer.py:16-23
// when t_rand != nullptr (t_rand points at this ray's S uniforms).
// z_user != nullptr (this ray's S caller-supplied depths, nb_render_args.z_vals): used as they are -- the fine pass of
// hierarchical sampling renders sorted(coarse z + importance samples), which no (near, far, t) formula produces.
__device__ __forceinline__ float z_sample(float near, float far, const float* __restrict__ t_vals, int s, int S,
                                          const float* __restrict__ t_rand, const float* __restrict__ z_user = nullptr) {
    if (z_user) return __ldg(z_user + s);
    float tc = t_vals ? __ldg(t_vals + s) : linspace01(s, S);
    float z = z_plain(near, far, tc);
    if (t_rand) {
        float lower = z, upper = z;
        if (s > 0) {
            float tp = t_vals ? __ldg(t_vals + s - 1) : linspace01(s - 1, S);
            lower = __fmul_rn(.5f, __fadd_rn(z, z_plain(near, far, tp)));
        }
        if (s < S - 1) {
            float tn = t_vals ? __ldg(t_vals + s + 1) : linspace01(s + 1, S);
            upper = __fmul_rn(.5f, __fadd_rn(z_plain(near, far, tn), z));
        }
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(t_rand + s)));
    }
    return z;
}

// Per-frame constants of the world -> grid transform, staged once per CTA work item.
struct FrameXf {
    float R[9];        // sp_input['R'][b]   row-major
    float Th[3];       // sp_input['Th'][b]
    float min_dhw[3];  // bounds[b,0,[2,1,0]]
    float voxel[3];    // cfg.voxel_size (dhw)
    float out_sh[3];   // dhw
};

// a5 + a6: latent_xyzc.py:41-60.  Input world point, output grid coords (x,y,z) in [-1,1].
__device__ __forceinline__ void world_to_grid(const FrameXf& f, float wx, float wy, float wz, float& gx, float& gy,
                                              float& gz) {
    float px = __fsub_rn(wx, f.Th[0]), py = __fsub_rn(wy, f.Th[1]), pz = __fsub_rn(wz, f.Th[2]);
    // torch.matmul(pts, R): c_j = sum_i p_i R[i][j]
    float cx = fmaf(pz, f.R[6], fmaf(py, f.R[3], __fmul_rn(px, f.R[0])));
    float cy = fmaf(pz, f.R[7], fmaf(py, f.R[4], __fmul_rn(px, f.R[1])));
    float cz = fmaf(pz, f.R[8], fmaf(py, f.R[5], __fmul_rn(px, f.R[2])));
    float d = __fdiv_rn(__fsub_rn(cz, f.min_dhw[0]), f.voxel[0]);
    float h = __fdiv_rn(__fsub_rn(cy, f.min_dhw[1]), f.voxel[1]);
    float w = __fdiv_rn(__fsub_rn(cx, f.min_dhw[2]), f.voxel[2]);
    d = __fsub_rn(__fmul_rn(__fdiv_rn(d, f.out_sh[0]), 2.f), 1.f);
    h = __fsub_rn(__fmul_rn(__fdiv_rn(h, f.out_sh[1]), 2.f), 1.f);
    w = __fsub_rn(__fmul_rn(__fdiv_rn(w, f.out_sh[2]), 2.f), 1.f);
    gx = w; gy = h; gz = d;   // grid_coords = dhw[..., [2,1,0]]
}

// F.grid_sample(align_corners=True) un-normalisation: ((g + 1) / 2) * (size - 1)
__device__ __forceinline__ float unnormalize(float g, int size) {
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));   // x / 2 == x * 0.5 bit for bit
}

// Trilinear corner set-up for one level (ATen grid_sampler_3d, zeros padding).
struct Corners {
    int x0, y0, z0;        // floor indices (may be -1 or size-1 => partly out of range); -2 = all out
    float wx[2], wy[2], wz[2];
};
__device__ __forceinline__ void corner_setup(float ix, float iy, float iz, int W, int H, int D, Corners& c) {
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    bool ok = (fx >= -1.f) && (fx <= (float)W) && (fy >= -1.f) && (fy <= (float)H) && (fz >= -1.f) && (fz <= (float)D);
    // (NaN coordinates fail the comparisons => all corners skipped => zeros, as ATen's bounds test does)
    c.x0 = ok ? (int)fx : -2; c.y0 = ok ? (int)fy : -2; c.z0 = ok ? (int)fz : -2;
    c.wx[0] = __fsub_rn(__fadd_rn(fx, 1.f), ix); c.wx[1] = __fsub_rn(ix, fx);
    c.wy[0] = __fsub_rn(__fadd_rn(fy, 1.f), iy); c.wy[1] = __fsub_rn(iy, fy);
    c.wz[0] = __fsub_rn(__fadd_rn(fz, 1.f), iz); c.wz[1] = __fsub_rn(iz, fz);
}
__device__ __forceinline__ bool corner_valid(const Corners& c, int dx, int dy, int dz, int W, int H, int D) {
    int x = c.x0 + dx, y = c.y0 + dy, z = c.z0 + dz;
    return (c.x0 != -2) && x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D;
}
__device__ __forceinline__ float corner_weight(const Corners& c, int dx, int dy, int dz) {
    return __fmul_rn(__fmul_rn(c.wx[dx], c.wy[dy]), c.wz[dz]);
}

// ---------------------------------------------------------------- f-1: if_clight_renderer_mmsk.py:12-45
// inside = AND over the mask views of msk[round(v)][round(u)], (u, v) = perspective projection of the world
// point, rounded half-to-even (torch.round) and clamped to the image like upstream.
__device__ __forceinline__ int mask_pixel(float r, int size) {
    // torch: .round().long() then clamp(0, size-1); non-finite / out-of-range casts give LONG_MIN on x86 => 0
    if (!(fabsf(r) < 9.0e18f)) return 0;
    const long long q = (long long)rintf(r);
    return (int)(q < 0 ? 0 : (q > size - 1 ? size - 1 : q));
}
// The single-view variant (if_clight_renderer_msk.py:17-30) first moves the sample into the world of the snapshot frame:
// can = (p - Th) @ R;  q = can @ R0^T + Th0.
__device__ __forceinline__ bool inside_masks(const RenderParams& P, const FrameXf& f, float wx, float wy, float wz) {
    if (P.mask_R0) {
        const float px = __fsub_rn(wx, f.Th[0]), py = __fsub_rn(wy, f.Th[1]), pz = __fsub_rn(wz, f.Th[2]);
        const float cx = fmaf(pz, f.R[6], fmaf(py, f.R[3], __fmul_rn(px, f.R[0])));
        const float cy = fmaf(pz, f.R[7], fmaf(py, f.R[4], __fmul_rn(px, f.R[1])));
        const float cz = fmaf(pz, f.R[8], fmaf(py, f.R[5], __fmul_rn(px, f.R[2])));
        const float* R0 = P.mask_R0;
        wx = __fadd_rn(fmaf(cz, __ldg(R0 + 2), fmaf(cy, __ldg(R0 + 1), __fmul_rn(cx, __ldg(R0 + 0)))), __ldg(P.mask_Th0 + 0));
        wy = __fadd_rn(fmaf(cz, __ldg(R0 + 5), fmaf(cy, __ldg(R0 + 4), __fmul_rn(cx, __ldg(R0 + 3)))), __ldg(P.mask_Th0 + 1));
        wz = __fadd_rn(fmaf(cz, __ldg(R0 + 8), fmaf(cy, __ldg(R0 + 7), __fmul_rn(cx, __ldg(R0 + 6)))), __ldg(P.mask_Th0 + 2));
    }
    for (int v = 0; v < P.mask_nv; ++v) {
        const float* RT = P.mask_RT + v * 12;
        const float* K = P.mask_Ks + v * 9;
        // pts @ R^T + T, then @ K^T
        const float cx = __fadd_rn(fmaf(wz, RT[2], fmaf(wy, RT[1], __fmul_rn(wx, RT[0]))), RT[3]);
        const float cy = __fadd_rn(fmaf(wz, RT[6], fmaf(wy, RT[5], __fmul_rn(wx, RT[4]))), RT[7]);
        const float cz = __fadd_rn(fmaf(wz, RT[10], fmaf(wy, RT[9], __fmul_rn(wx, RT[8]))), RT[11]);
        const float ix = fmaf(cz, K[2], fmaf(cy, K[1], __fmul_rn(cx, K[0])));
        const float iy = fmaf(cz, K[5], fmaf(cy, K[4], __fmul_rn(cx, K[3])));
        const float iz = fmaf(cz, K[8], fmaf(cy, K[7], __fmul_rn(cx, K[6])));
        const int u = mask_pixel(__fdiv_rn(ix, iz), P.mask_W), w = mask_pixel(__fdiv_rn(iy, iz), P.mask_H);
        if (!__ldg(P.mask_msks + ((size_t)v * P.mask_H + w) * P.mask_W + u)) return false;
    }
    return true;
}

// ---------------------------------------------------------------- a9: embedder.py:5-50
// out[0..2] = x; out[3+6f+j] = sin(2^f x_j); out[6+6f+j] = cos(2^f x_j)   (x * freq is exact: freq = 2^f)
template <int L, typename Store>
__device__ __forceinline__ void positional_embed(float x, float y, float z, Store&& store) {
    store(0, x); store(1, y); store(2, z);
    float f = 1.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        float s, c;
        sincosf(x * f, &s, &c); store(3 + 6 * l + 0, s); store(3 + 6 * l + 3, c);
        sincosf(y * f, &s, &c); store(3 + 6 * l + 1, s); store(3 + 6 * l + 4, c);
        sincosf(z * f, &s, &c); store(3 + 6 * l + 2, s); store(3 + 6 * l + 5, c);
        f *= 2.f;
    }
}

// Same layout, cheaper: an accurate sincosf only every ANCHOR-th octave, the octaves in between by the
// double-angle recurrence (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a).  Each doubling at most doubles the
// absolute error, so with ANCHOR <= 5 the values stay within ~2e-6 of sincosf -- far below the fp16 rounding
// (2^-11) they get as tensor-core operands.  Used by the tensor-core kernel only (colour path).
template <int L, int ANCHOR, typename Store>
__device__ __forceinline__ void positional_embed_anchored(float x, float y, float z, Store&& store) {
    store(0, x); store(1, y); store(2, z);
    float f = 1.f;
    float sx = 0.f, cx = 1.f, sy = 0.f, cy = 1.f, sz = 0.f, cz = 1.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        if (l % ANCHOR == 0) {
            sincosf(x * f, &sx, &cx); sincosf(y * f, &sy, &cy); sincosf(z * f, &sz, &cz);
        } else {
            float t;
            t = 2.f * sx * cx; cx = fmaf(-2.f * sx, sx, 1.f); sx = t;
            t = 2.f * sy * cy; cy = fmaf(-2.f * sy, sy, 1.f); sy = t;
            t = 2.f * sz * cz; cz = fmaf(-2.f * sz, sz, 1.f); sz = t;
        }
        store(3 + 6 * l + 0, sx); store(3 + 6 * l + 1, sy); store(3 + 6 * l + 2, sz);
        store(3 + 6 * l + 3, cx); store(3 + 6 * l + 4, cy); store(3 + 6 * l + 5, cz);
        f *= 2.f;
    }
}

// ---------------------------------------------------------------- a10: raw2outputs
// nerf_net_utils.py:6-51, one warp per ray.  raw = (rgb logits x3, sigma) per sample in smem,
// z = perturbed z_vals in smem.  All lanes return the same reduced values.
struct RayOut { float r, g, b, depth, acc; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ RayOut composite_ray(const float4* __restrict__ raw, const float* __restrict__ z, int S,
                                                float norm_d, float* __restrict__ weights_out, int lane) {
    float T_run = 1.f;
    float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
    for (int base = 0; base < S; base += 32) {
        int s = base + lane;
        float alpha = 0.f, fac = 1.f, zr = 0.f;
        float4 rw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < S) {
            rw = raw[s];
            zr = z[s];
            float dist = (s + 1 < S) ? __fsub_rn(z[s + 1], zr) : 1e10f;
            dist = __fmul_rn(dist, norm_d);
            alpha = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(rw.w, 0.f), dist)));   // 1 - exp(-relu(sigma) * dists)
            fac = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);                      // 1 - alpha + 1e-10
        }
        // inclusive product scan over the 32 lanes
        float incl = fac;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= up;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        float w = alpha * (T_run * excl);
        T_run *= __shfl_sync(0xffffffffu, incl, 31);
        if (s < S) {
            if (weights_out) weights_out[s] = w;
            ar += w * (1.f / (1.f + expf(-rw.x)));   // torch.sigmoid
            ag += w * (1.f / (1.f + expf(-rw.y)));
            ab += w * (1.f / (1.f + expf(-rw.z)));
            ad += w * zr;
            aa += w;
        }
    }
    RayOut o;
    o.r = warp_sum(ar); o.g = warp_sum(ag); o.b = warp_sum(ab); o.depth = warp_sum(ad); o.acc = warp_sum(aa);
    return o;
}

// disp_map = 1 / max(1e-10, depth / acc); torch.max propagates the NaN of 0/0 (nerf_net_utils.py:44-45)
__device__ __forceinline__ float disparity(float depth, float acc) {
    float q = __fdiv_rn(depth, acc);
    float m = (q != q) ? q : fmaxf(1e-10f, q);
    return __fdiv_rn(1.f, m);
}

}  // namespace nb
