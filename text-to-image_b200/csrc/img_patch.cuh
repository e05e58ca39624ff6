// On-chip 4x4 / stride-2 SAME patches of an fp32 NHWC 3-channel image (shared by conv_gemm.cu A_IMG and wgrad_gemm.cu IMG).
//
// The image reaches these kernels as PADDED bf16 ROWS (t2i_img_to_rows, one small pass per image: 6 bytes per pixel
// against the 128 bytes per patch row of a materialised patch matrix):
//     B[n][r][j] = bf16(row_r[j - 3])   (zeros for j < 3 and j >= 3*w + 3:  the left / right SAME padding)
// planes [np][n][h][pitch] like every other activation (np = 2: hi + lo).  A producer group streams the 2*bp + 2 rows a
// tile needs into a shared-memory ring with bulk copies (several tiles ahead) and assembles each pixel's 48 patch values
// (kh, kw, c) by pure 32-bit word moves: the 12 values of one kh are the 24 contiguous bytes B[2*pl + kh][6q .. 6q + 12)
// -- word offset 3q, no bounds checks, no conversions -- and the row of the 128B-swizzled operand tile (K-major A of the
// forward product == MN-major operand of the weight gradient) is six 16-byte chunks of those words.
#pragma once
#include "ptx.cuh"

namespace t2i {

// words (bf16 pairs) per padded row: 3*w + 6 entries rounded up to a multiple of 4
__host__ __device__ inline int img_pitch_words(int img_w) { return ((img_w * 3 + 6 + 3) / 4) * 2; }

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The padded bf16 rows of one tile (image rows ih0 .. ih0 + n_rows - 1 of every plane) -> ring slot `dst`
// ([plane][n_rows][pitchw] words).  Rows outside the image (SAME padding; at most the first and the last row of a tile)
// are zero-filled by all `nthreads` producers (nthreads >= np); then one bulk copy per plane brings the valid rows.
// Call it after a barrier that guarantees nobody still reads the slot.
__device__ __forceinline__ void img_rows_fetch(const uint32_t* rows, long long plane_words, long long sample_words, int np,
                                               int n, int ih0, int img_h, int n_rows, int pitchw, uint32_t* dst,
                                               uint64_t* bar, int pt, int nthreads) {
    const int lo = ih0 < 0 ? -ih0 : 0;                                   // rows [0, lo) and [hi, n_rows) are outside
    const int hi = (ih0 + n_rows > img_h) ? img_h - ih0 : n_rows;
    if (lo > 0 || hi < n_rows) {
        for (int pl = 0; pl < np; ++pl)
            for (int r = 0; r < n_rows; ++r)
                if (r < lo || r >= hi)
                    for (int c = pt; c >= 0 && c < pitchw; c += nthreads) dst[(pl * n_rows + r) * pitchw + c] = 0u;
    }
    // The valid rows of a plane are CONTIGUOUS in global memory and in the slot: one bulk copy per plane (issuing a copy
    // costs the issuing thread ~0.1 us whatever its size -- one copy per row made the fetch the longest part of a trip).
    // Thread 0 arms the barrier; a copy may complete before the expect_tx lands: the transaction count is signed, the
    // phase needs thread 0's arrival too.
    if (pt == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(np * (hi - lo) * pitchw * 4));
    if (pt >= 0 && pt < np && hi > lo)
        bulk_load_1d(dst + (pt * n_rows + lo) * pitchw,
                     rows + pt * plane_words + n * sample_words + static_cast<long long>(ih0 + lo) * pitchw,
                     static_cast<uint32_t>((hi - lo) * pitchw * 4), bar);
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
