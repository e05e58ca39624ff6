// On-chip 4x4 / stride-2 SAME patches of an fp32 NHWC 3-channel image (shared by conv_gemm.cu A_IMG and wgrad_gemm.cu IMG).
//
// A producer group stages the 2*bp + 2 image rows a tile needs as raw fp32 (cp.async, double buffered: the next tile's
// rows stream in while this tile is assembled), converts them ONCE to bf16 into a padded row buffer
//     B[r][j] = bf16(row_r[j - 3])   (zeros for j < 3 and j >= 3*w + 3:  the left / right SAME padding)
// and then assembles each pixel's 48 patch values (kh, kw, c) by pure 32-bit word moves: the 12 values of one kh are the
// 24 contiguous bytes B[2*pl + kh][6q .. 6q + 12) -- word offset 3q, no bounds checks, no conversions -- and the row of the
// 128B-swizzled operand tile (K-major A of the forward product == MN-major operand of the weight gradient) is six
// 16-byte chunks of those words.  np = 2 keeps a second buffer with the low parts bf16(x - bf16(x)).
#pragma once
#include "ptx.cuh"

namespace t2i {

// words (bf16 pairs) per padded row: 3*w + 6 entries rounded up to a multiple of 4
__host__ __device__ inline int img_pitch_words(int img_w) { return ((img_w * 3 + 6 + 3) / 4) * 2; }

// raw fp32 rows [n_rows][iw3] -> bf16 padded rows s_hi (and s_lo when two planes are needed); executed by `nthreads`
// threads, this one is `pt`.  The (row, chunk) pairs are walked flat, two per trip with the loads of both issued before
// the first use, and the row / chunk indices advance incrementally (no divisions): the producer warps are few, so what
// matters is the length of the dependent instruction chain per tile.
__device__ __forceinline__ void img_convert_chunk(const float* src, int c, int iw3, float (&v)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = 4 * c - 3 + e;
        v[e] = (j >= 0 && j < iw3) ? src[j] : 0.f;
    }
}
__device__ __forceinline__ void img_store_chunk(uint32_t* s_hi, uint32_t* s_lo, int off, const float (&v)[4]) {
    uint2 hi;
    hi.x = pack_bf16x2(v[0], v[1]);
    hi.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(s_hi + off) = hi;
    if (s_lo != nullptr) {
        uint2 lo;
        lo.x = pack_bf16x2(v[0] - bf16_lo(hi.x), v[1] - bf16_hi(hi.x));
        lo.y = pack_bf16x2(v[2] - bf16_lo(hi.y), v[3] - bf16_hi(hi.y));
        *reinterpret_cast<uint2*>(s_lo + off) = lo;
    }
}
__device__ __forceinline__ void img_rows_convert(const float* raw, int n_rows, int iw3, uint32_t* s_hi, uint32_t* s_lo,
                                                 int pitchw, int pt, int nthreads) {
    const int cpr = pitchw >> 1;           // 4-entry chunks per row
    int r0 = 0, c0 = pt;                   // first chunk of this thread; the second one is nthreads further
    while (c0 >= cpr) { c0 -= cpr; ++r0; }
    int r1 = r0, c1 = c0 + nthreads;
    while (c1 >= cpr) { c1 -= cpr; ++r1; }
    // a step of 2 * nthreads chunks = dr rows and dc chunks
    int dr = 0, dc = 2 * nthreads;
    while (dc >= cpr) { dc -= cpr; ++dr; }
    while (r0 < n_rows) {
        float a[4], b[4];
        const bool two = r1 < n_rows;
        img_convert_chunk(raw + r0 * iw3, c0, iw3, a);
        if (two) img_convert_chunk(raw + r1 * iw3, c1, iw3, b);
        img_store_chunk(s_hi, s_lo, r0 * pitchw + 2 * c0, a);
        if (two) img_store_chunk(s_hi, s_lo, r1 * pitchw + 2 * c1, b);
        r0 += dr; c0 += dc;
        if (c0 >= cpr) { c0 -= cpr; ++r0; }
        r1 += dr; c1 += dc;
        if (c1 >= cpr) { c1 -= cpr; ++r1; }
    }
}

// The raw rows of one tile: n_rows x iw3 floats starting at image row ih0 of `base` (rows outside [0, img_h) are zero
// filled = SAME padding), streamed with 16-byte cp.async; exactly the copies, the caller commits the group.
__device__ __forceinline__ void img_rows_prefetch(const float* base, int ih0, int img_h, int n_rows, int iw3, float* dst_rows,
                                                  const float* any_valid, int pt, int nthreads) {
    const int iw3q = iw3 >> 2;
    int r = 0, c4 = pt;
    while (c4 >= iw3q) { c4 -= iw3q; ++r; }
    int dr = 0, dc = nthreads;
    while (dc >= iw3q) { dc -= iw3q; ++dr; }
    float4* dst = reinterpret_cast<float4*>(dst_rows);
    while (r < n_rows) {
        const int ih = ih0 + r;
        const bool ok = ih >= 0 && ih < img_h;
        cp_async16(dst + r * iw3q + c4, ok ? base + static_cast<long long>(ih) * iw3 + c4 * 4 : any_valid, ok);
        r += dr; c4 += dc;
        if (c4 >= iw3q) { c4 -= iw3q; ++r; }
    }
}

// One operand-tile row (pixel (pl, q) of the tile, tile row `row`): 48 bf16 = chunks 0..5 of the 128-byte row at `dst`
// (16-byte chunk j of row r lives at chunk j ^ (r & 7)); zero_tail also clears chunks 6, 7 (columns 48..63).
__device__ __forceinline__ void img_patch_row(const uint32_t* s_bf, int pitchw, int pl, int q, uint8_t* dst, int row,
                                              bool zero_tail) {
    uint32_t w[4][6];
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
        const uint32_t* src = s_bf + (2 * pl + kh) * pitchw + 3 * q;
#pragma unroll
        for (int i = 0; i < 6; ++i) w[kh][i] = src[i];
    }
    const int sw = row & 7;
    *reinterpret_cast<uint4*>(dst + ((0 ^ sw) * 16)) = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
    *reinterpret_cast<uint4*>(dst + ((1 ^ sw) * 16)) = make_uint4(w[0][4], w[0][5], w[1][0], w[1][1]);
    *reinterpret_cast<uint4*>(dst + ((2 ^ sw) * 16)) = make_uint4(w[1][2], w[1][3], w[1][4], w[1][5]);
    *reinterpret_cast<uint4*>(dst + ((3 ^ sw) * 16)) = make_uint4(w[2][0], w[2][1], w[2][2], w[2][3]);
    *reinterpret_cast<uint4*>(dst + ((4 ^ sw) * 16)) = make_uint4(w[2][4], w[2][5], w[3][0], w[3][1]);
    *reinterpret_cast<uint4*>(dst + ((5 ^ sw) * 16)) = make_uint4(w[3][2], w[3][3], w[3][4], w[3][5]);
    if (zero_tail) {
        *reinterpret_cast<uint4*>(dst + ((6 ^ sw) * 16)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dst + ((7 ^ sw) * 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

}  // namespace t2i
