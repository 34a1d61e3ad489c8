// Layout of the packed decoder-weight blob shared by nb_pack.cu and the render kernels.
//
// Decoder as written upstream (lib/networks/latent_xyzc.py:99-121):
//   h0 = relu(fc_0 f)   h1 = relu(fc_1 h0)   h2 = relu(fc_2 h1)   sigma = alpha_fc h2
//   u  = feature_fc h2 ; v = latent_fc [u (+) latent] ; w = relu(view_fc [v (+) PE(view) (+) PE(xyz)])
//   rgb = rgb_fc w
// There is no activation between feature_fc, latent_fc and view_fc[:, :256], so they fold
// exactly into one 128x256 matrix Wc and a per-frame bias bc:
//   w = relu(Wc h2 + Wx PE(xyz) + Wv PE(view) + bc)
#pragma once
#include <stddef.h>

namespace nb {

constexpr int kFeat = 352;      // 32 + 64 + 128 + 128
constexpr int kHidden = 256;
constexpr int kColor = 128;
constexpr int kXyzPE = 63;
constexpr int kViewPE = 27;
constexpr int kColorK = 320;    // 256 (h2) + 63 (PE xyz) + 1 zero pad

// ---- fp32 section (float offsets). "t" = transposed / K-major: Wt[k][n] = W[n][k]
constexpr size_t oW0t = 0;                                  // [352][256]
constexpr size_t oB0 = oW0t + (size_t)kFeat * kHidden;      // [256]
constexpr size_t oW1t = oB0 + kHidden;                      // [256][256]
constexpr size_t oB1 = oW1t + (size_t)kHidden * kHidden;
constexpr size_t oW2t = oB1 + kHidden;                      // [256][256]
constexpr size_t oB2 = oW2t + (size_t)kHidden * kHidden;
constexpr size_t oAlphaW = oB2 + kHidden;                   // [256]
constexpr size_t oAlphaB = oAlphaW + kHidden;               // [1] (+3 pad)
constexpr size_t oWct = oAlphaB + 4;                        // [320][128]: rows 0..255 Wc^T, 256..318 Wx^T, 319 zero
constexpr size_t oWvt = oWct + (size_t)kColorK * kColor;    // [27][128]  Wv^T (+ pad to 28 rows)
constexpr size_t oRgbW = oWvt + (size_t)28 * kColor;        // [3][128]
constexpr size_t oRgbB = oRgbW + 3 * kColor;                // [3] (+1 pad)
constexpr size_t kF32Floats = oRgbB + 4;

// ---- fp16 section: tcgen05 canonical K-major no-swizzle operand tiles.
// Element (n, k) of an N x K matrix lives at half-offset
//     ((k/8) * (N/8) + n/8) * 64 + (n%8) * 8 + (k%8)
// i.e. 8x8 "core matrices" (8 rows x 16 bytes, 128 B contiguous), core matrices of one
// 8-wide K chunk contiguous over N (stride-byte-offset 128 B), K chunks N*16 B apart
// (leading-byte-offset).  A K=16 MMA step is therefore one contiguous N*32-byte slab.
constexpr size_t kF16ByteOffset = ((kF32Floats * 4 + 255) / 256) * 256;
constexpr size_t hW0 = 0;                                   // N=256, K=352
constexpr size_t hW1 = hW0 + (size_t)kHidden * kFeat;       // N=256, K=256
constexpr size_t hW2 = hW1 + (size_t)kHidden * kHidden;     // N=256, K=256
constexpr size_t hW3 = hW2 + (size_t)kHidden * kHidden;     // N=128, K=320  (Wc | Wx | 0)
constexpr size_t kF16Halves = hW3 + (size_t)kColor * kColorK;

// ---- scratch for the fp64 fold (doubles), then per-frame bias bc (floats)
constexpr size_t kScratchByteOffset = ((kF16ByteOffset + kF16Halves * 2 + 255) / 256) * 256;
constexpr size_t kScratchDoubles = (size_t)kColor * kHidden;   // T = view_fc[:, :256] * latent_fc[:, :256]
constexpr size_t kBcByteOffset = kScratchByteOffset + kScratchDoubles * 8;
// after bc[B][128] floats: u[B][256] doubles (fold scratch)

__host__ __device__ inline size_t packed_weights_bytes(int batch) {
    size_t b = kBcByteOffset + (size_t)batch * kColor * 4;
    b = (b + 255) / 256 * 256;
    return b + (size_t)batch * kHidden * 8;
}
__host__ __device__ inline size_t u_byte_offset(int batch) {
    size_t b = kBcByteOffset + (size_t)batch * kColor * 4;
    return (b + 255) / 256 * 256;
}

__host__ __device__ inline size_t umma_kmajor_offset(int n, int k, int N) {
    return ((size_t)(k >> 3) * (N >> 3) + (n >> 3)) * 64 + (size_t)(n & 7) * 8 + (k & 7);
}

}  // namespace nb
